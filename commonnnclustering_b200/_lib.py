"""ctypes binding of ``libcnnb200.so`` (the C-ABI declared in ``include/cnnb200.h``).

There is no CPU fallback: if the library is missing or cannot be loaded, importing the compute
path raises.  ``__graft_entry__.build()`` / ``python -m commonnnclustering_b200.build`` compile it.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcnnb200.so")

HOST, DEVICE = 0, 1
METRIC_IDS = {"euclidean_r": 0, "euclidean": 1, "euclidean_periodic_r": 2}
METHOD_IDS = {"auto": 0, "grid": 1, "tiled": 2, "tensor": 3}

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_size_t, C.c_void_p)
FREE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)

_P = C.c_void_p
_I64 = C.c_int64

# name -> (argtypes); every function returns int except the two noted below
SIGNATURES = {
    "cnnb200_ctx_create": [C.c_int, C.POINTER(_P)],
    "cnnb200_ctx_destroy": [_P],
    "cnnb200_ctx_set_stream": [_P, _P],
    "cnnb200_ctx_set_allocator": [_P, ALLOC_FN, FREE_FN, _P],
    "cnnb200_ctx_synchronize": [_P],
    "cnnb200_ctx_set_periodic_bounds": [_P, _P, _I64],
    "cnnb200_ctx_release": [_P],
    "cnnb200_neighbours_from_points": [_P, _P, C.c_int, _I64, _I64, C.c_double, C.c_int, C.c_int, C.POINTER(_I64)],
    "cnnb200_grid_begin": [_P, _P, C.c_int, _I64, _I64, C.c_double, C.c_int],
    "cnnb200_grid_balance": [_P, C.c_int, _P],
    "cnnb200_grid_halo": [_P, _I64, _I64, C.POINTER(_I64), C.POINTER(_I64)],
    "cnnb200_grid_degrees": [_P, _I64, _I64, _P],
    "cnnb200_grid_fill": [_P, _I64, _I64, _P, _P],
    "cnnb200_neighbours_adopt": [_P, _P, _P, _I64, _I64, _I64, C.c_int],
    "cnnb200_neighbours_from_csr": [_P, _P, _P, C.c_int, C.c_int, C.c_int, _I64, _I64],
    "cnnb200_neighbours_from_padded": [_P, _P, _P, C.c_int, _I64, _I64],
    "cnnb200_neighbours_from_distances": [_P, _P, C.c_int, _I64, C.c_double],
    "cnnb200_neighbours_from_condensed": [_P, _P, C.c_int, _I64, C.c_double],
    "cnnb200_neighbours_sort_rows": [_P, C.POINTER(C.c_int)],
    "cnnb200_neighbours_asymmetry": [_P, C.POINTER(_I64)],
    "cnnb200_neighbours_info": [_P, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)],
    "cnnb200_neighbours_export": [_P, _P, _P, C.c_int],
    "cnnb200_neighbours_device_view": [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)],
    "cnnb200_cn_count": [_P, _I64, _P, C.c_int, C.POINTER(_I64)],
    "cnnb200_cn_count_rows": [_P, _I64, _I64, _I64, _P, C.c_int, C.POINTER(_I64)],
    "cnnb200_cn_check_pairs": [_P, _P, _I64, _I64, _P, C.c_int],
    "cnnb200_fit": [_P, _I64, _I64, _I64, _P, _P, C.c_int, C.POINTER(_I64)],
    "cnnb200_components_begin": [_P, _I64, _I64, _P, C.c_int],
    "cnnb200_components_edges": [_P, _I64, _I64],
    "cnnb200_components_phase": [_P, _I64, _I64, C.c_int],
    "cnnb200_components_from_counts": [_P, _P],
    "cnnb200_components_parents": [_P, C.POINTER(_P)],
    "cnnb200_components_merge": [_P, _P, _I64],
    "cnnb200_components_labels": [_P, _I64, _P, C.c_int, C.c_int, C.POINTER(_I64)],
    "cnnb200_components_labels_sorted": [_P, _I64, _I64, _I64, _P, C.c_int, C.POINTER(_I64)],
    "cnnb200_fit_sharded_begin": [_P, _P, C.c_int, _I64, _I64, C.c_double, C.c_int, _I64, _I64, C.c_int, C.c_int, _P,
                                  C.POINTER(_P)],
    "cnnb200_fit_sharded_end": [_P, _P, _I64, C.c_int, _I64, _I64, _I64, _P, C.c_int, C.POINTER(_I64)],
    "cnnb200_fit_sharded_ranges": [_P, _I64, _P],
    "cnnb200_fit_sharded_end_windows": [_P, _P, _I64, _P, _I64, C.c_int, _I64, _I64, _I64, _P, C.c_int, C.POINTER(_I64)],
    "cnnb200_neighbours_other_from_points": [_P, _P, C.c_int, _I64, _I64, C.POINTER(_I64)],
    "cnnb200_neighbours_other_from_csr": [_P, _P, _P, C.c_int, _I64, _I64],
    "cnnb200_predict": [_P, _I64, _I64, _P, _P, _I64, _P, _P, C.c_int],
    "cnnb200_sort_by_size": [_P, _P, C.c_int, _I64, _I64, _I64, _P, _I64],
    "cnnb200_fit_points": [_P, _P, C.c_int, _I64, _I64, C.c_double, C.c_int, C.c_int, _I64, _I64, _I64,
                           C.c_int, _I64, _I64, _P, C.POINTER(_I64)],
    "cnnb200_fit_sorted": [_P, _I64, _I64, _I64, _I64, _I64, _P, C.c_int, C.POINTER(_I64)],
    "cnnb200_label_summary": [_P, _P],
    "cnnb200_labels_group": [_P, _P, _P, C.c_int, _I64, _P, _P],
    "cnnb200_host_register": [_P, C.c_size_t],
    "cnnb200_host_unregister": [_P],
    "cnnb200_stats": [_P, _P, _I64],
    "cnnb200_ctx_set_timing": [_P, C.c_int],
    "cnnb200_timings": [_P, _P, _P, _I64, C.POINTER(_I64), C.c_int],
}

TIME_SLOTS = ["h2d", "grid_build", "search_count", "scan", "search_fill", "ingest", "comp_init", "comp_sample",
              "comp_edges", "comp_flatten", "labels", "sort_by_size", "d2h", "cn_count"]

_lib = None


class CnnB200Error(RuntimeError):
    pass


def load():
    """Loads the shared library (once).  Raises ImportError when it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built "
            "(run `python -m commonnnclustering_b200.build`); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.cnnb200_version.restype = C.c_int
    lib.cnnb200_version.argtypes = []
    lib.cnnb200_last_error.restype = C.c_char_p
    lib.cnnb200_last_error.argtypes = []
    for name, argtypes in SIGNATURES.items():
        f = getattr(lib, name)
        f.restype = C.c_int
        f.argtypes = argtypes
    _lib = lib
    return lib


_ERRORS = {-1: ValueError, -3: MemoryError}


def check(rc: int):
    if rc != 0:
        msg = load().cnnb200_last_error().decode(errors="replace")
        raise _ERRORS.get(rc, CnnB200Error)(f"cnnb200 error {rc}: {msg}")


def _ptr_loc(a):
    """(raw pointer, HOST|DEVICE) of a numpy array or a torch tensor."""
    if a is None:
        return None, HOST
    if isinstance(a, np.ndarray):
        if not a.flags.c_contiguous:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data, HOST
    # torch tensor (duck-typed so that numpy-only callers never import torch)
    if not a.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return a.data_ptr(), (DEVICE if a.is_cuda else HOST)


class Context:
    """One ``cnnb200_ctx``: a CUDA device, a stream and the device scratch of the fit path."""

    def __init__(self, device: int = 0, stream=None, torch_allocator: bool = False):
        self._lib = load()
        h = _P()
        check(self._lib.cnnb200_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self._keep = []
        if stream is not None:
            self.set_stream(stream)
        if torch_allocator:
            self.use_torch_allocator()

    # -- plumbing ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.cnnb200_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream):
        ptr = getattr(stream, "cuda_stream", stream)
        check(self._lib.cnnb200_ctx_set_stream(self._h, _P(int(ptr) if ptr else None)))

    def use_torch_allocator(self):
        """Route device scratch through torch's caching allocator (PyTorch as buffer carrier)."""
        import torch

        dev = self.device

        def _alloc(nbytes, _user):
            try:
                return torch.cuda.caching_allocator_alloc(int(nbytes), dev)
            except Exception:
                return None

        delete = torch.cuda.caching_allocator_delete  # bound now: module globals vanish at shutdown

        def _free(ptr, _user):
            try:
                delete(ptr)
            except Exception:
                pass

        a, f = ALLOC_FN(_alloc), FREE_FN(_free)
        self._keep = [a, f]
        check(self._lib.cnnb200_ctx_set_allocator(self._h, a, f, None))

    def set_periodic_bounds(self, bounds):
        b = np.ascontiguousarray(bounds, dtype=np.float64)
        check(self._lib.cnnb200_ctx_set_periodic_bounds(self._h, b.ctypes.data, b.shape[0]))

    def synchronize(self):
        check(self._lib.cnnb200_ctx_synchronize(self._h))

    def release(self):
        check(self._lib.cnnb200_ctx_release(self._h))

    # -- neighbourhoods ---------------------------------------------------------------------
    def neighbours_from_points(self, x, radius_cutoff, metric="euclidean_r", method="auto"):
        p, loc = _ptr_loc(x)
        n, d = x.shape
        nnz = _I64()
        check(self._lib.cnnb200_neighbours_from_points(
            self._h, p, loc, n, d, float(radius_cutoff), METRIC_IDS[metric], METHOD_IDS[method], C.byref(nnz)))
        return nnz.value

    # -- row-sharded construction (multi-GPU building blocks) ------------------------------------
    def grid_begin(self, x, radius_cutoff, metric="euclidean_r"):
        p, loc = _ptr_loc(x)
        n, d = x.shape
        check(self._lib.cnnb200_grid_begin(self._h, p, loc, n, d, float(radius_cutoff), METRIC_IDS[metric]))

    def grid_balance(self, parts: int):
        b = np.zeros(parts + 1, dtype=np.int64)
        check(self._lib.cnnb200_grid_balance(self._h, int(parts), b.ctypes.data))
        return b

    def grid_halo(self, row_begin, row_end):
        lo, hi = _I64(), _I64()
        check(self._lib.cnnb200_grid_halo(self._h, int(row_begin), int(row_end), C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def grid_degrees(self, row_begin, row_end, deg):
        p, loc = _ptr_loc(deg)
        assert loc == DEVICE
        check(self._lib.cnnb200_grid_degrees(self._h, int(row_begin), int(row_end), p))

    def grid_fill(self, row_begin, row_end, row_ptr, col_idx):
        rp, loc = _ptr_loc(row_ptr)
        cp, loc2 = _ptr_loc(col_idx)
        assert loc == DEVICE and loc2 == DEVICE
        check(self._lib.cnnb200_grid_fill(self._h, int(row_begin), int(row_end), rp, cp))

    def neighbours_adopt(self, row_ptr, col_idx, max_deg, grid_perm):
        rp, loc = _ptr_loc(row_ptr)
        cp, loc2 = _ptr_loc(col_idx)
        assert loc == DEVICE and loc2 == DEVICE
        self._borrowed = (row_ptr, col_idx)
        check(self._lib.cnnb200_neighbours_adopt(self._h, rp, cp, row_ptr.shape[0] - 1, col_idx.shape[0],
                                                 int(max_deg), int(bool(grid_perm))))

    def neighbours_from_csr(self, row_ptr, col_idx, borrow=False):
        rp, loc = _ptr_loc(row_ptr)
        cp, loc2 = _ptr_loc(col_idx)
        if loc != loc2:
            raise ValueError("row_ptr and col_idx must live in the same memory space")
        n = row_ptr.shape[0] - 1
        nnz = col_idx.shape[0]
        isz = 8 if "64" in str(col_idx.dtype) else 4
        if borrow:
            self._borrowed = (row_ptr, col_idx)
        check(self._lib.cnnb200_neighbours_from_csr(self._h, rp, cp, isz, loc, int(bool(borrow)), n, nnz))

    def neighbours_from_padded(self, padded, counts):
        pp, loc = _ptr_loc(padded)
        cp, loc2 = _ptr_loc(counts)
        if loc != loc2:
            raise ValueError("padded and counts must live in the same memory space")
        n, max_deg = padded.shape
        check(self._lib.cnnb200_neighbours_from_padded(self._h, pp, cp, loc, n, max_deg))

    def neighbours_from_distances(self, dist, radius_cutoff):
        """Neighbourhoods from a precomputed f64 [n, n] distance matrix (``dist[i, j] <= radius_cutoff``)."""
        if len(dist.shape) != 2 or dist.shape[0] != dist.shape[1]:
            raise ValueError("the distance matrix must be square")
        if "float64" not in str(dist.dtype):
            raise ValueError("the distance matrix must be float64")
        dp, loc = _ptr_loc(dist)
        check(self._lib.cnnb200_neighbours_from_distances(self._h, dp, loc, dist.shape[0], float(radius_cutoff)))

    def neighbours_from_condensed(self, dist, n, radius_cutoff):
        """Neighbourhoods from condensed f64 distances ``[n (n - 1) / 2]`` (row-by-row upper triangle, the layout of
        ``scipy.spatial.distance.pdist`` and of the reference's ``InputDataExtDistancesLinearMemoryview``)."""
        n = int(n)
        if len(dist.shape) != 1 or dist.shape[0] != n * (n - 1) // 2:
            raise ValueError("condensed distances: expected a vector of n (n - 1) / 2 values")
        if "float64" not in str(dist.dtype):
            raise ValueError("the distances must be float64")
        dp, loc = _ptr_loc(dist) if dist.shape[0] else (None, HOST)
        check(self._lib.cnnb200_neighbours_from_condensed(self._h, dp, loc, n, float(radius_cutoff)))

    def neighbours_sort_rows(self) -> bool:
        """Sorts the held rows ascending when they are not; True when a sort was needed."""
        flag = C.c_int(0)
        check(self._lib.cnnb200_neighbours_sort_rows(self._h, C.byref(flag)))
        return bool(flag.value)

    def neighbours_asymmetry(self) -> int:
        """Number of entries (i, j) of the held neighbourhoods whose mirror (j, i) is missing."""
        k = _I64()
        check(self._lib.cnnb200_neighbours_asymmetry(self._h, C.byref(k)))
        return k.value

    def neighbours_info(self):
        n, nnz, md = _I64(), _I64(), _I64()
        check(self._lib.cnnb200_neighbours_info(self._h, C.byref(n), C.byref(nnz), C.byref(md)))
        return n.value, nnz.value, md.value

    def neighbours_export(self):
        """(row_ptr int64[n+1], col_idx int64[nnz]) in the caller's index space, rows ascending."""
        n, nnz, _ = self.neighbours_info()
        row_ptr = np.empty(n + 1, dtype=np.int64)
        col_idx = np.empty(max(nnz, 1), dtype=np.int64)
        check(self._lib.cnnb200_neighbours_export(self._h, row_ptr.ctypes.data, col_idx.ctypes.data, HOST))
        return row_ptr, col_idx[:nnz]

    def neighbours_export_device(self):
        """The same CSR (caller's index space, rows ascending) as CUDA tensors: row_ptr int64 [n+1], col_idx int32 [nnz]."""
        import torch

        n, nnz, _ = self.neighbours_info()
        dev = torch.device("cuda", self.device)
        row_ptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
        col64 = torch.empty(max(nnz, 1), dtype=torch.int64, device=dev)
        check(self._lib.cnnb200_neighbours_export(self._h, row_ptr.data_ptr(), col64.data_ptr(), DEVICE))
        return row_ptr, col64[:nnz].to(torch.int32)

    def neighbours_device_view(self):
        a, b, c = _P(), _P(), _P()
        check(self._lib.cnnb200_neighbours_device_view(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    # -- similarity ---------------------------------------------------------------------------
    def cn_count(self, n_member_cutoff, out=None, rows=None):
        """uint32 common-neighbour count per CSR slot (internal CSR order); 0xFFFFFFFF = not an edge.
        ``rows=(begin, end)`` restricts the pass to a row block (row-sharded runs)."""
        n, nnz, _ = self.neighbours_info()
        if out is None:
            out = np.empty(max(nnz, 1), dtype=np.uint32)
        p, loc = _ptr_loc(out)
        ne = _I64()
        b, e = (0, n) if rows is None else rows
        check(self._lib.cnnb200_cn_count_rows(self._h, int(n_member_cutoff), int(b), int(e), p, loc, C.byref(ne)))
        return out[:nnz], ne.value

    def cn_check_pairs(self, pairs, similarity_cutoff):
        """Early-exit check of ``pairs`` (int32 [k, 2], internal row indices) -> uint8 [k]: bit 0 = per-lane merge,
        bit 1 = whole-warp window form; both say |N(i) ∩ N(j)| >= similarity_cutoff."""
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        out = np.zeros(pairs.shape[0], dtype=np.uint8)
        check(self._lib.cnnb200_cn_check_pairs(self._h, pairs.ctypes.data, pairs.shape[0], int(similarity_cutoff),
                                               out.ctypes.data, HOST))
        return out

    # -- components -----------------------------------------------------------------------------
    def fit(self, n_member_cutoff, similarity_cutoff, current_start, labels, consider=None):
        lp, loc = _ptr_loc(labels)
        cp, loc2 = _ptr_loc(consider)
        if consider is not None and loc != loc2:
            raise ValueError("labels and consider must live in the same memory space")
        nc = _I64()
        check(self._lib.cnnb200_fit(self._h, int(n_member_cutoff), int(similarity_cutoff), int(current_start),
                                    lp, cp, loc, C.byref(nc)))
        return nc.value

    def components_begin(self, n_member_cutoff, similarity_cutoff, consider=None):
        cp, loc = _ptr_loc(consider)
        check(self._lib.cnnb200_components_begin(self._h, int(n_member_cutoff), int(similarity_cutoff), cp, loc))

    def components_edges(self, row_begin, row_end):
        check(self._lib.cnnb200_components_edges(self._h, int(row_begin), int(row_end)))

    def components_phase(self, row_begin, row_end, phase):
        check(self._lib.cnnb200_components_phase(self._h, int(row_begin), int(row_end), int(phase)))

    def components_from_counts(self, counts):
        p, loc = _ptr_loc(counts)
        if loc != DEVICE:
            raise ValueError("counts must be a device tensor")
        check(self._lib.cnnb200_components_from_counts(self._h, p))

    def components_parents_ptr(self):
        p = _P()
        check(self._lib.cnnb200_components_parents(self._h, C.byref(p)))
        return p.value

    def components_merge(self, others, count):
        p, loc = _ptr_loc(others)
        if loc != DEVICE:
            raise ValueError("foreign parent arrays must be device tensors")
        check(self._lib.cnnb200_components_merge(self._h, p, int(count)))

    def fit_sharded_begin(self, x, radius_cutoff, n_member_cutoff, similarity_cutoff, rank, world, metric="euclidean_r"):
        """-> (block bounds int64 [world + 1], device pointer of this rank's int32 [n] parent array)."""
        p, loc = _ptr_loc(x)
        n, d = x.shape
        bounds = np.zeros(world + 1, dtype=np.int64)
        parents = _P()
        check(self._lib.cnnb200_fit_sharded_begin(self._h, p, loc, n, d, float(radius_cutoff), METRIC_IDS[metric],
                                                  int(n_member_cutoff), int(similarity_cutoff), int(rank), int(world),
                                                  bounds.ctypes.data, C.byref(parents)))
        return bounds, parents.value

    def fit_sharded_end(self, others, world, labels, sort_by_size=True, current_start=1, member_cutoff=None,
                        max_clusters=None):
        """``labels`` None: merge only (this rank's caller does not want the labels)."""
        op, loc = _ptr_loc(others)
        if loc != DEVICE:
            raise ValueError("the gathered parent arrays must be a device tensor")
        lp, lloc = _ptr_loc(labels)
        nc = _I64()
        check(self._lib.cnnb200_fit_sharded_end(
            self._h, op, int(world), int(bool(sort_by_size)), int(current_start),
            -1 if member_cutoff is None else int(member_cutoff), -1 if max_clusters is None else int(max_clusters),
            lp, lloc, C.byref(nc)))
        return nc.value

    def fit_sharded_ranges(self, world):
        """-> int64 [world, 2]: the rows [lo, hi) every rank built in the last ``fit_sharded_begin`` (block + halo)."""
        r = np.zeros((int(world), 2), dtype=np.int64)
        check(self._lib.cnnb200_fit_sharded_ranges(self._h, int(world), r.ctypes.data))
        return r

    def fit_sharded_end_windows(self, others, world, starts, length, labels, sort_by_size=True, current_start=1,
                                member_cutoff=None, max_clusters=None):
        """``fit_sharded_end`` on the gathered WINDOWS ``[world, length]`` of the parent arrays (window g starts at row
        ``starts[g]``)."""
        op, loc = _ptr_loc(others)
        if loc != DEVICE:
            raise ValueError("the gathered parent windows must be a device tensor")
        lp, lloc = _ptr_loc(labels)
        st = np.ascontiguousarray(starts, dtype=np.int64)
        nc = _I64()
        check(self._lib.cnnb200_fit_sharded_end_windows(
            self._h, op, int(world), st.ctypes.data, int(length), int(bool(sort_by_size)), int(current_start),
            -1 if member_cutoff is None else int(member_cutoff), -1 if max_clusters is None else int(max_clusters),
            lp, lloc, C.byref(nc)))
        return nc.value

    def components_labels(self, current_start, labels, fresh=True):
        lp, loc = _ptr_loc(labels)
        nc = _I64()
        check(self._lib.cnnb200_components_labels(self._h, int(current_start), lp, loc, int(bool(fresh)), C.byref(nc)))
        return nc.value

    def components_labels_sorted(self, labels, current_start=1, member_cutoff=None, max_clusters=None):
        """Fresh labels of the finished forest, size-sorted (= components_labels + sort_by_size); returns the number
        of clusters before the size sort."""
        lp, loc = _ptr_loc(labels)
        nc = _I64()
        check(self._lib.cnnb200_components_labels_sorted(
            self._h, int(current_start), -1 if member_cutoff is None else int(member_cutoff),
            -1 if max_clusters is None else int(max_clusters), lp, loc, C.byref(nc)))
        return nc.value

    # -- prediction ---------------------------------------------------------------------------------
    def neighbours_other_from_points(self, x_other):
        p, loc = _ptr_loc(x_other)
        m, d = x_other.shape
        nnz = _I64()
        check(self._lib.cnnb200_neighbours_other_from_points(self._h, p, loc, m, d, C.byref(nnz)))
        return nnz.value

    def neighbours_other_from_csr(self, row_ptr, col_idx):
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
        col_idx = np.ascontiguousarray(col_idx, dtype=np.int64)
        check(self._lib.cnnb200_neighbours_other_from_csr(self._h, row_ptr.ctypes.data, col_idx.ctypes.data, HOST,
                                                          row_ptr.shape[0] - 1, col_idx.shape[0]))

    def predict(self, n_member_cutoff, similarity_cutoff, labels, label_ok, other_labels, other_consider):
        """In-place update of ``other_labels`` / ``other_consider`` (NumPy arrays)."""
        lp, loc = _ptr_loc(labels)
        op, loc2 = _ptr_loc(other_labels)
        cp, loc3 = _ptr_loc(other_consider)
        kp, loc4 = _ptr_loc(label_ok)
        if not (loc == loc2 == loc3 == loc4):
            raise ValueError("predict: all arrays must live in the same memory space")
        check(self._lib.cnnb200_predict(self._h, int(n_member_cutoff), int(similarity_cutoff), lp, kp,
                                        int(label_ok.shape[0] - 1), op, cp, loc))

    def sort_by_size(self, labels, member_cutoff=None, max_clusters=None, return_map=False, max_label=None):
        lp, loc = _ptr_loc(labels)
        n = labels.shape[0]
        mc = -1 if member_cutoff is None else int(member_cutoff)
        mx = -1 if max_clusters is None else int(max_clusters)
        if return_map:
            if max_label is None:
                max_label = int(labels.max()) if n else 0
            m = np.zeros(max_label + 1, dtype=np.int64)
            check(self._lib.cnnb200_sort_by_size(self._h, lp, loc, n, mc, mx, m.ctypes.data, m.shape[0]))
            return m
        check(self._lib.cnnb200_sort_by_size(self._h, lp, loc, n, mc, mx, None, 0))
        return None

    def fit_points(self, x, radius_cutoff, n_member_cutoff, similarity_cutoff, labels, current_start=1,
                   metric="euclidean_r", method="auto", sort_by_size=True, member_cutoff=None, max_clusters=None):
        xp, loc = _ptr_loc(x)
        lp, loc2 = _ptr_loc(labels)
        if loc != loc2:
            raise ValueError("x and labels must live in the same memory space")
        n, d = x.shape
        nc = _I64()
        check(self._lib.cnnb200_fit_points(
            self._h, xp, loc, n, d, float(radius_cutoff), METRIC_IDS[metric], METHOD_IDS[method],
            int(n_member_cutoff), int(similarity_cutoff), int(current_start), int(bool(sort_by_size)),
            -1 if member_cutoff is None else int(member_cutoff),
            -1 if max_clusters is None else int(max_clusters), lp, C.byref(nc)))
        return nc.value

    def labels_group(self, prev, cur):
        """(order, rank) int64 [n]: points sorted by (prev, cur, index) and their position among the points with the
        same previous label (``cnnb200_labels_group``)."""
        prev = np.ascontiguousarray(prev, dtype=np.int64)
        cur = np.ascontiguousarray(cur, dtype=np.int64)
        n = prev.shape[0]
        order = np.empty(n, dtype=np.int64)
        rank = np.empty(n, dtype=np.int64)
        check(self._lib.cnnb200_labels_group(self._h, prev.ctypes.data, cur.ctypes.data, HOST, n, order.ctypes.data,
                                             rank.ctypes.data))
        return order, rank

    def fit_sorted(self, n_member_cutoff, similarity_cutoff, labels, current_start=1, member_cutoff=None, max_clusters=None):
        """Fresh fit on the held neighbourhoods + sort_by_size in one call; returns the number of clusters found."""
        lp, loc = _ptr_loc(labels)
        nc = _I64()
        check(self._lib.cnnb200_fit_sorted(self._h, int(n_member_cutoff), int(similarity_cutoff), int(current_start),
                                           -1 if member_cutoff is None else int(member_cutoff),
                                           -1 if max_clusters is None else int(max_clusters), lp, loc, C.byref(nc)))
        return nc.value

    def label_summary(self):
        """(clusters kept, members of the largest one, noise points) of the last size-sorted result."""
        out = np.zeros(3, dtype=np.int64)
        check(self._lib.cnnb200_label_summary(self._h, out.ctypes.data))
        return int(out[0]), int(out[1]), int(out[2])

    def set_timing(self, enabled: bool):
        check(self._lib.cnnb200_ctx_set_timing(self._h, int(bool(enabled))))

    def timings(self, reset: bool = True):
        """Per-stage device time (CUDA events on the context's stream) accumulated since the last
        reset: {stage: (total ms, calls)} plus the number of hand-written kernels launched."""
        k = len(TIME_SLOTS)
        ms = np.zeros(k, dtype=np.float64)
        calls = np.zeros(k, dtype=np.int64)
        launches = _I64()
        check(self._lib.cnnb200_timings(self._h, ms.ctypes.data, calls.ctypes.data, k, C.byref(launches), int(reset)))
        return {name: (float(ms[i]), int(calls[i])) for i, name in enumerate(TIME_SLOTS)}, launches.value

    def stats(self):
        s = np.zeros(8, dtype=np.int64)
        check(self._lib.cnnb200_stats(self._h, s.ctypes.data, 8))
        return {"candidate_edges": int(s[0]), "checks": int(s[1]), "accepted": int(s[2]),
                "skipped": int(s[3]), "distance_evals": int(s[4]), "rechecks": int(s[5]),
                "window_fallbacks": int(s[6]), "search_method": {0: None, 1: "grid", 2: "tiled", 3: "tensor"}.get(int(s[7]) & 15),
                "count_packed": bool(int(s[7]) & 16)}


def host_register(array) -> bool:
    """Page-locks a NumPy array in place; False when the driver refuses (then copies stay pageable)."""
    try:
        check(load().cnnb200_host_register(array.ctypes.data, array.nbytes))
        return True
    except CnnB200Error:
        return False


def host_unregister(ptr: int):
    load().cnnb200_host_unregister(ptr)


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    ctx = _default_ctx.get(device)
    if ctx is None:
        ctx = _default_ctx[device] = Context(device)
    return ctx
