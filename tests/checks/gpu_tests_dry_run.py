"""Dry run of the `-m gpu` test files WITHOUT a GPU: `_lib.Context` is replaced by the oracle-backed stand-in
(tests/standin.py), so the tests' own Python and the package's host logic execute while nothing is measured or proven
about the device.  Catches attribute / signature breakage before GPU time is spent.  Tests that need CUDA tensors, the
library's error contract or device statistics fail here by construction (17 of 109 at the time of writing); everything
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

        class Ctx(StandInContext):
            def __init__(self, device=0, **_):
                super().__init__()

            def close(self):
                pass

        shared = {}
        _lib.Context = Ctx
        _lib.default_context = lambda device=0: shared.setdefault(device, Ctx())


if __name__ == "__main__":
    import pytest

    slow = ["c1_full", "c1_r019_c37", "c2_full", "c1_through"]     # minutes on the CPU oracle
    args = [os.path.join(TESTS, "test_gpu_parity.py"), os.path.join(TESTS, "test_plugin.py"), "-m", "gpu", "-q",
            "-k", " and ".join(f"not {s}" for s in slow)] + sys.argv[1:]
    sys.exit(pytest.main(args, plugins=[_Plugin()]))
