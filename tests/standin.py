"""TEST INFRASTRUCTURE: a stand-in for `commonnnclustering_b200._lib.Context` that answers the context's contract on
the CPU from the oracle (`oracle/port.py`) and NumPy.  It exists so that the package's HOST logic -- recipes, input data
types, getters, the fitters' parameter handling, `Clustering.fit / predict / isolate / reel / fit_hierarchical`,
`Labels.sort_by_size`'s bookkeeping -- runs under `-m "not gpu"` and can be compared with the reference there.  It is
injected by the tests (`fitter._ctx = StandInContext()`); nothing under `commonnnclustering_b200/` knows it, and the
product has no CPU path: without a GPU `_lib.Context()` raises (tests/test_abi.py::test_fails_loudly_without_a_gpu)."""
import numpy as np

from oracle import port

NOT_AN_EDGE = 0xFFFFFFFF


def _sorted_rows(row_ptr, col_idx):
    out = np.array(col_idx, dtype=np.int64, copy=True)
    changed = False
    for i in range(row_ptr.shape[0] - 1):
        row = out[row_ptr[i]:row_ptr[i + 1]]
        if row.size > 1 and np.any(row[1:] < row[:-1]):
            row.sort()
            changed = True
    return out, changed


class StandInContext:
    device = 0

    def __init__(self):
        self.row_ptr = self.col_idx = None
        self.other = None
        self.bounds = None
        self.search = None            # (x, radius_cutoff, metric) of the last search: prediction reuses them
        self.summary = (0, 0, 0)
        self.calls = []               # names of the context calls made, in order (tests look at the route taken)

    # ---- ingest ------------------------------------------------------------------------------------------------
    def _adopt(self, row_ptr, col_idx, n):
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
        col_idx = np.ascontiguousarray(col_idx, dtype=np.int64)
        if row_ptr.shape[0] != n + 1 or row_ptr[0] != 0 or np.any(np.diff(row_ptr) < 0) or row_ptr[-1] != col_idx.shape[0]:
            raise ValueError("row offsets are not a monotone prefix of the index array")
        if col_idx.size and (col_idx.min() < 0 or col_idx.max() >= n):
            raise ValueError("neighbour index out of range")
        self.col_idx, self.unsorted_on_arrival = _sorted_rows(row_ptr, col_idx)   # every ingest path checks and sorts
        self.row_ptr = row_ptr

    def set_periodic_bounds(self, bounds):
        self.calls.append("set_periodic_bounds")
        self.bounds = np.ascontiguousarray(bounds, dtype=np.float64)

    def neighbours_from_points(self, x, radius_cutoff, metric="euclidean_r", method="auto"):
        self.calls.append("neighbours_from_points")
        x = np.ascontiguousarray(x, dtype=np.float64)
        if metric == "euclidean_periodic_r":
            if self.bounds is None or self.bounds.shape[0] != x.shape[1]:
                raise RuntimeError("periodic metric without bounds for every dimension")
            rp, ci = port.neighbours_periodic(x, radius_cutoff, self.bounds)
        else:
            rp, ci = port.neighbours(x, radius_cutoff, metric, "brute")
        self._adopt(rp, ci, x.shape[0])
        self.search = (x, float(radius_cutoff), metric)
        return int(rp[-1])

    def neighbours_from_distances(self, dist, radius_cutoff):
        self.calls.append("neighbours_from_distances")
        rp, ci = port.neighbours_from_distances(np.asarray(dist), radius_cutoff)
        self._adopt(rp, ci, dist.shape[0])

    def neighbours_from_condensed(self, dist, n, radius_cutoff):
        self.calls.append("neighbours_from_condensed")
        square = np.zeros((n, n))
        iu = np.triu_indices(n, k=1)
        square[iu] = np.asarray(dist)
        square[(iu[1], iu[0])] = np.asarray(dist)
        rp, ci = port.neighbours_from_distances(square, radius_cutoff)
        self._adopt(rp, ci, n)

    def neighbours_from_csr(self, row_ptr, col_idx, borrow=False):
        self.calls.append("neighbours_from_csr")
        self._adopt(np.asarray(row_ptr), np.asarray(col_idx), np.asarray(row_ptr).shape[0] - 1)

    def neighbours_from_padded(self, padded, counts):
        self.calls.append("neighbours_from_padded")
        padded, counts = np.asarray(padded), np.asarray(counts, dtype=np.int64)
        if counts.size and (counts.min() < 0 or counts.max() > padded.shape[1]):
            raise ValueError("neighbour count out of range")
        row_ptr = np.zeros(counts.shape[0] + 1, dtype=np.int64)
        np.cumsum(counts, out=row_ptr[1:])
        keep = np.arange(padded.shape[1])[None, :] < counts[:, None]
        self._adopt(row_ptr, padded[keep], counts.shape[0])

    def neighbours_sort_rows(self):
        self.calls.append("neighbours_sort_rows")
        return bool(self.unsorted_on_arrival)

    def neighbours_asymmetry(self):
        rows = np.repeat(np.arange(self.row_ptr.shape[0] - 1), np.diff(self.row_ptr))
        have = set(zip(rows.tolist(), self.col_idx.tolist()))
        return sum((j, i) not in have for i, j in have)

    def neighbours_info(self):
        deg = np.diff(self.row_ptr)
        return int(deg.shape[0]), int(self.row_ptr[-1]), int(deg.max()) if deg.size else 0

    def neighbours_export(self):
        return self.row_ptr.copy(), self.col_idx.copy()

    def _row(self, i):
        return self.col_idx[self.row_ptr[i]:self.row_ptr[i + 1]]

    # ---- similarity --------------------------------------------------------------------------------------------
    def cn_count(self, n_member_cutoff, out=None, rows=None):
        deg = np.diff(self.row_ptr)
        counts = np.full(max(int(self.row_ptr[-1]), 1), NOT_AN_EDGE, dtype=np.uint32)
        edges = 0
        for i in range(deg.shape[0]):
            if not deg[i] > n_member_cutoff:
                continue
            for k in range(self.row_ptr[i], self.row_ptr[i + 1]):
                j = int(self.col_idx[k])
                if j > i and deg[j] > n_member_cutoff:
                    counts[k] = port.count_common(self._row(i), self._row(j))
                    edges += 1
        return counts[:int(self.row_ptr[-1])], edges

    def cn_check_pairs(self, pairs, similarity_cutoff):
        pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
        return np.array([3 if port.check(self._row(i), self._row(j), similarity_cutoff) else 0 for i, j in pairs],
                        dtype=np.uint8)

    # ---- fit ---------------------------------------------------------------------------------------------------
    def fit(self, n_member_cutoff, similarity_cutoff, current_start, labels, consider=None):
        self.calls.append("fit")
        n = self.row_ptr.shape[0] - 1
        if consider is None:          # fresh: the labels are fully overwritten
            labels[:] = 0
            consider = np.ones(n, dtype=np.uint8)
        before = int(labels.max()) if n else 0
        port.fit_bfs(self.row_ptr, self.col_idx, n_member_cutoff, similarity_cutoff, current_start, labels=labels,
                     consider=consider)
        return max(int(labels.max()) - max(before, current_start - 1), 0) if n else 0

    def _remember(self, labels):
        counts = np.bincount(labels)
        self.summary = (int(np.count_nonzero(counts[1:])), int(counts[1:].max()) if counts.size > 1 else 0, int(counts[0]))

    def fit_sorted(self, n_member_cutoff, similarity_cutoff, labels, current_start=1, member_cutoff=None, max_clusters=None):
        self.calls.append("fit_sorted")
        found = self.fit(n_member_cutoff, similarity_cutoff, current_start, labels, None)
        self.calls.pop()
        port.sort_by_size(labels, member_cutoff, max_clusters)
        self._remember(labels)
        return found

    def fit_points(self, x, radius_cutoff, n_member_cutoff, similarity_cutoff, labels, current_start=1,
                   metric="euclidean_r", method="auto", sort_by_size=True, member_cutoff=None, max_clusters=None):
        self.neighbours_from_points(x, radius_cutoff, metric, method)
        self.calls[-1] = "fit_points"
        found = self.fit(n_member_cutoff, similarity_cutoff, current_start, labels, None)
        self.calls.pop()
        if sort_by_size:
            port.sort_by_size(labels, member_cutoff, max_clusters)
        self._remember(labels)
        return found

    def label_summary(self):
        return self.summary

    def sort_by_size(self, labels, member_cutoff=None, max_clusters=None, return_map=False, max_label=None):
        self.calls.append("sort_by_size")
        old = labels.copy()
        port.sort_by_size(labels, member_cutoff, max_clusters)
        self._remember(labels)
        if not return_map:
            return None
        reassign = np.zeros((int(old.max()) if max_label is None else max_label) + 1, dtype=np.int64)
        reassign[old] = labels
        return reassign

    def labels_group(self, prev, cur):
        self.calls.append("labels_group")
        n = prev.shape[0]
        order = np.lexsort((np.arange(n), cur, prev))
        rank = np.empty(n, dtype=np.int64)
        for p in np.unique(prev):
            idx = np.flatnonzero(prev == p)
            rank[idx] = np.arange(idx.size)
        return order.astype(np.int64), rank[order]

    def stats(self):
        return {"candidate_edges": 0, "checks": 0, "accepted": 0, "skipped": 0, "distance_evals": 0, "rechecks": 0,
                "window_fallbacks": 0, "search_method": None, "count_packed": False}

    # ---- prediction --------------------------------------------------------------------------------------------
    def neighbours_other_from_points(self, x_other):
        self.calls.append("neighbours_other_from_points")
        x, radius_cutoff, metric = self.search
        self.other = port.neighbours_other(x, x_other, radius_cutoff, metric)
        return int(self.other[0][-1])

    def neighbours_other_from_csr(self, row_ptr, col_idx):
        self.calls.append("neighbours_other_from_csr")
        self.other = (np.ascontiguousarray(row_ptr, dtype=np.int64), np.ascontiguousarray(col_idx, dtype=np.int64))

    def predict(self, n_member_cutoff, similarity_cutoff, labels, label_ok, other_labels, other_consider):
        self.calls.append("predict")
        # the FIRST member in the caller's list order decides; the check itself works on ascending copies
        o_rp, o_ci = self.other
        clusters = set(np.flatnonzero(label_ok).tolist())
        for p in range(o_rp.shape[0] - 1):
            if not other_consider[p]:
                continue
            members = o_ci[o_rp[p]:o_rp[p + 1]]
            if not members.size > n_member_cutoff:
                continue
            mine = np.sort(members)
            for member in members:
                label = int(labels[member])
                if label not in clusters:
                    continue
                row = self._row(member)
                if not row.size > n_member_cutoff:
                    continue
                if similarity_cutoff == 0 or port.count_common(mine, row) >= similarity_cutoff:
                    other_consider[p] = 0
                    other_labels[p] = label
                    break
