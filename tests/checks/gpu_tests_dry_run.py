"""Dry run of the `-m gpu` test files WITHOUT a GPU: `_lib.Context` is replaced by the oracle-backed stand-in
(tests/standin.py), so the tests' own Python and the package's host logic execute while nothing is measured or proven
about the device.  Catches attribute / signature breakage before GPU time is spent.  Tests that need CUDA tensors, the
library's error contract or device statistics fail here by construction (14 of 109 at the time of writing); everything
that goes through `Clustering(...)` passes.

    python tests/checks/gpu_tests_dry_run.py [extra pytest arguments]
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
sys.path[:0] = [TESTS, ROOT]


class _Plugin:
    @staticmethod
    def pytest_configure(config):
        from standin import StandInContext

        from commonnnclustering_b200 import _lib

        import numpy as np
        import torch

        from oracle import port

        # "cuda" tensors of the sweep driver live on the CPU here
        real_device, real_empty = torch.device, torch.empty
        torch.device = lambda *a, **k: real_device("cpu")
        torch.empty = lambda *a, **k: real_empty(*a, **{key: v for key, v in k.items() if key != "pin_memory"})
        torch.cuda.synchronize = lambda *a, **k: None

        class Ctx(StandInContext):
            def __init__(self, device=0, **_):
                super().__init__()

            def close(self):
                pass

            def neighbours_from_points(self, x, radius_cutoff, metric="euclidean_r", method="auto"):
                return super().neighbours_from_points(x.numpy() if hasattr(x, "numpy") else x, radius_cutoff, metric, method)

            # the count-mode route of sweep.fit_sweep: counts once per radius, then one component pass per cell
            def cn_count(self, n_member_cutoff, out=None, rows=None):
                if out is None or not hasattr(out, "numpy"):
                    return super().cn_count(n_member_cutoff, out, rows)
                counts, edges = super().cn_count(n_member_cutoff)
                out.numpy()[:counts.shape[0]] = counts.view(np.int32)
                return out, edges

            def components_begin(self, n_member_cutoff, similarity_cutoff, consider=None):
                self._cell = (n_member_cutoff, similarity_cutoff)

            def components_from_counts(self, counts):
                pass

            def components_labels_sorted(self, labels, current_start=1, member_cutoff=None, max_clusters=None):
                lab = port.fit_bfs(self.row_ptr, self.col_idx, *self._cell, current_start)
                found = int(lab.max())
                port.sort_by_size(lab, member_cutoff, max_clusters)
                self._remember(lab)
                labels.numpy()[:] = lab
                return found

        shared = {}
        _lib.Context = Ctx
        _lib.default_context = lambda device=0: shared.setdefault(device, Ctx())


if __name__ == "__main__":
    import pytest

    slow = ["c1_full", "c1_r019_c37", "c2_full", "c1_through"]     # minutes on the CPU oracle
    args = [os.path.join(TESTS, "test_gpu_parity.py"), os.path.join(TESTS, "test_plugin.py"), "-m", "gpu", "-q",
            "-k", " and ".join(f"not {s}" for s in slow)] + sys.argv[1:]
    sys.exit(pytest.main(args, plugins=[_Plugin()]))
