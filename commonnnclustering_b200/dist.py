"""Row-sharded fit over the GPUs of one box: one process per GPU, ``torch.distributed`` for the
plumbing (NCCL on the box, gloo in the CPU tests).  SURVEY.md section 8(e):

* coordinates are replicated, query rows (points in cell-sorted order) are split into contiguous
  blocks that balance the number of candidate pairs;
* coordinates input: no CSR exchange at all -- every rank searches its block plus the halo of rows its
  block's neighbour lists refer to (contiguous in the cell-sorted order); rows outside stay empty;
* precomputed neighbourhoods (and ``exchange=True``): exchange 1 -- the CSR: degrees by an all-reduce
  of the row-indexed int32 array (each rank fills only its block), then every rank writes its rows
  straight into its slice of the gathered ``col_idx`` buffer and the slices are broadcast;
* every rank settles the candidate edges of its own row block on a private parent array;
* exchange 2 -- the parent arrays: all-gather of ``G x N`` int32, then each rank unions
  ``(i, parent_g[i])`` for all g.  Hooking always goes towards the smaller root, so the merged
  forest -- and with it every label -- is identical on all ranks and independent of ``G``.

The collectives are ordinary ``torch.distributed`` calls on tensors that alias the buffers the C
library works on; the compute between them is the same C-ABI used on a single GPU.
"""
import os

import numpy as np
import torch
import torch.distributed as dist


class _DevPtr:
    """Exposes a raw device pointer to torch through ``__cuda_array_interface__``."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def tensor_from_ptr(ptr, shape, typestr, device):
    return torch.as_tensor(_DevPtr(ptr, shape, typestr), device=device)


# ---- backend-agnostic plumbing (also exercised with gloo on CPU tensors) ---------------------------
def block_of(bounds, rank):
    return int(bounds[rank]), int(bounds[rank + 1])


def exchange_degrees(deg, group=None):
    """deg: row-indexed int32 [n+1], zero outside the caller's row block -> full degrees on every rank;
    returns row_ptr int64 [n+1]."""
    dist.all_reduce(deg, op=dist.ReduceOp.SUM, group=group)
    row_ptr = torch.zeros(deg.shape[0], dtype=torch.int64, device=deg.device)
    torch.cumsum(deg[:-1], 0, dtype=torch.int64, out=row_ptr[1:])
    return row_ptr


def exchange_slices(col_idx, slice_bounds, group=None):
    """col_idx [nnz] holds the caller's slice [slice_bounds[rank], slice_bounds[rank+1]) -> all slices."""
    world = dist.get_world_size(group)
    works = []
    for g in range(world):
        s, e = int(slice_bounds[g]), int(slice_bounds[g + 1])
        if e > s:
            src = dist.get_global_rank(group, g) if group is not None else g
            works.append(dist.broadcast(col_idx[s:e], src=src, group=group, async_op=True))
    for w in works:
        w.wait()


def chunk_of(total, world):
    """Length of the equal flat chunks an array of ``total`` elements is cut into for ``world`` ranks."""
    return -(-max(int(total), 1) // world)


def allgather_chunks(full, chunk, group=None):
    """``full`` [world * chunk] holds the caller's chunk ``[rank * chunk, (rank + 1) * chunk)`` -> all chunks, in place
    (NCCL: ``all_gather_into_tensor`` on a view of the buffer itself; gloo: per-chunk views)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = full[rank * chunk:(rank + 1) * chunk]
    if full.is_cuda:
        dist.all_gather_into_tensor(full, mine, group=group)
    else:
        dist.all_gather(list(full.view(world, chunk).unbind(0)), mine.clone(), group=group)


def plan_windows(ranges, n):
    """ranges int64 [world, 2]: the rows [lo, hi) every rank built -> (length, starts): the ranks exchange
    ``parents[starts[r] : starts[r] + length]``; every window covers its rank's range and stays inside [0, n)."""
    ranges = np.asarray(ranges, dtype=np.int64)
    length = max(int((ranges[:, 1] - ranges[:, 0]).max()), 1)
    starts = np.maximum(np.minimum(ranges[:, 0], n - length), 0)
    return length, starts


def gather_parents(parents, group=None):
    """parents int32 [n] of this rank -> int32 [G, n] of all ranks."""
    world = dist.get_world_size(group)
    out = torch.empty((world, parents.shape[0]), dtype=parents.dtype, device=parents.device)
    if parents.is_cuda:
        dist.all_gather_into_tensor(out, parents, group=group)
    else:
        dist.all_gather(list(out.unbind(0)), parents, group=group)
    return out


class ShardedFit:
    """``Clustering.fit`` hot path on ``world_size`` GPUs; every rank ends up with the full labels."""

    def __init__(self, ctx, group=None):
        self.ctx = ctx
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.device("cuda", ctx.device)
        ctx.set_stream(torch.cuda.current_stream(self.device))  # collectives are ordered against this stream
        self.last = {}
        self._registered = {}
        # exchange windows of the parent arrays instead of whole arrays (fit_points): pays from 8 ranks on
        env = os.environ.get("CNNB200_DIST_WINDOWS")
        self.windows = (env == "1") if env is not None else self.world >= 8

    # -- neighbourhoods from coordinates, rows sharded ---------------------------------------------
    def neighbours_from_points(self, x, radius_cutoff, metric="euclidean_r", exchange=False):
        """Coordinates are replicated, so a rank can build every row it will ever read by itself: its
        row block plus the halo of rows that occur in the block's neighbour lists (one contiguous range
        of the cell-sorted order, ``grid_halo``).  No CSR travels; rows outside the range stay empty.
        ``exchange=True`` builds the block only and all-gathers the CSR instead (the layout used for
        precomputed neighbourhoods, kept for comparison)."""
        ctx, G = self.ctx, self.world
        n = x.shape[0]
        ctx.grid_begin(x, radius_cutoff, metric)          # replicated: every rank sorts all points
        bounds = ctx.grid_balance(G)                      # identical on every rank (same data)
        b, e = block_of(bounds, self.rank)
        deg = torch.zeros(n + 1, dtype=torch.int32, device=self.device)
        if not exchange:
            lo, hi = ctx.grid_halo(b, e)
            ctx.grid_degrees(lo, hi, deg)
            row_ptr = torch.zeros(n + 1, dtype=torch.int64, device=self.device)
            torch.cumsum(deg[:-1], 0, dtype=torch.int64, out=row_ptr[1:])
            host = torch.stack([row_ptr[-1], deg.max().to(torch.int64)]).cpu().numpy()
            nnz, max_deg = int(host[0]), int(host[1])
            col_idx = torch.empty(max(nnz, 1), dtype=torch.int32, device=self.device)[:nnz]
            ctx.grid_fill(lo, hi, row_ptr, col_idx)
            ctx.neighbours_adopt(row_ptr, col_idx, max_deg, grid_perm=True)
            self.last = dict(bounds=bounds, halo=(lo, hi), nnz=nnz)
            return bounds
        ctx.grid_degrees(b, e, deg)
        row_ptr = exchange_degrees(deg, self.group)
        bt = torch.as_tensor(bounds, device=self.device)
        host = torch.cat([row_ptr[bt], deg.max().to(torch.int64).view(1)]).cpu().numpy()
        slice_bounds, max_deg = host[:-1], int(host[-1])
        nnz = int(slice_bounds[-1])
        col_idx = torch.empty(max(nnz, 1), dtype=torch.int32, device=self.device)[:nnz]
        ctx.grid_fill(b, e, row_ptr, col_idx)             # lands directly in this rank's slice
        exchange_slices(col_idx, slice_bounds, self.group)
        ctx.neighbours_adopt(row_ptr, col_idx, max_deg, grid_perm=True)
        self.last = dict(bounds=bounds, slice_bounds=slice_bounds, nnz=nnz)
        return bounds

    # -- neighbourhoods from a precomputed CSR: each rank uploads only its row block ------------------
    def upload_csr_block(self, row_ptr_host, col_idx_host):
        """Host CSR (row_ptr int64 [n+1], col_idx int32 [nnz]; every rank sees the arrays, e.g. memory-mapped) ->
        1/G of the bytes in HBM on every rank.  Both arrays are cut into G EQUAL flat chunks (a chunk need not end at a
        row: the upload and the exchange move bytes, only the component pass works on row blocks), so the exchange is
        one in-place ``all_gather_into_tensor`` per array.  The arrays are page-locked in place once; the G uploads run
        concurrently, each over its own PCIe link."""
        G = self.world
        n = row_ptr_host.shape[0] - 1
        nnz = int(row_ptr_host[-1])
        from . import _lib

        for a in (row_ptr_host, col_idx_host):
            key = (a.ctypes.data, a.nbytes)
            if key not in self._registered:
                self._registered[key] = _lib.host_register(a)
        key = (row_ptr_host.ctypes.data, n, nnz, G)
        plan = getattr(self, "_csr_plan", None)
        if plan is None or plan[0] != key:     # bounds depend on the offsets only: computed once per structure
            targets = (np.arange(G + 1, dtype=np.float64) * nnz / G).astype(np.int64)
            bounds = np.searchsorted(row_ptr_host, targets, side="left").astype(np.int64)
            bounds[0], bounds[-1] = 0, n
            bounds = np.maximum.accumulate(bounds)                     # row blocks balanced by entries
            chunk_r, chunk_c = chunk_of(n + 1, G), chunk_of(nnz, G)       # flat chunks of the two arrays
            row_ptr = torch.empty(G * chunk_r, dtype=torch.int64, device=self.device)
            col_idx = torch.empty(G * chunk_c, dtype=torch.int32, device=self.device)
            plan = self._csr_plan = (key, bounds, chunk_r, chunk_c, row_ptr, col_idx)
        _, bounds, chunk_r, chunk_c, row_ptr, col_idx = plan
        g = self.rank
        a, b = g * chunk_r, min((g + 1) * chunk_r, n + 1)
        if b > a:
            row_ptr[a:b].copy_(torch.from_numpy(row_ptr_host[a:b]), non_blocking=True)
        a, b = g * chunk_c, min((g + 1) * chunk_c, nnz)
        if b > a:
            if col_idx_host.dtype == np.int32:
                col_idx[a:b].copy_(torch.from_numpy(col_idx_host[a:b]), non_blocking=True)
            else:
                col_idx[a:b].copy_(torch.from_numpy(np.ascontiguousarray(col_idx_host[a:b])).to(self.device, non_blocking=True))
        return dict(n=n, nnz=nnz, bounds=bounds, chunk_r=chunk_r, chunk_c=chunk_c, row_ptr=row_ptr, col_idx=col_idx)

    def exchange_csr(self, blk):
        """Exchange 1 of SURVEY 8e on bytes that already sit in HBM: all-gather of the G equal chunks of ``row_ptr``
        and of ``col_idx`` (in place, NCCL over NVLink); the gathered CSR is adopted.  Returns the row-block bounds."""
        allgather_chunks(blk["row_ptr"], blk["chunk_r"], self.group)
        allgather_chunks(blk["col_idx"], blk["chunk_c"], self.group)
        self.ctx.neighbours_adopt(blk["row_ptr"][:blk["n"] + 1], blk["col_idx"][:blk["nnz"]], -1, grid_perm=False)
        self.last = dict(bounds=blk["bounds"], nnz=blk["nnz"], two_phase=True)
        return blk["bounds"]

    def neighbours_from_csr_host(self, row_ptr_host, col_idx_host):
        return self.exchange_csr(self.upload_csr_block(row_ptr_host, col_idx_host))

    # -- components on the gathered CSR, rows sharded ----------------------------------------------------
    def fit(self, bounds, n_member_cutoff, similarity_cutoff, labels, current_start=1, sort_by_size=True,
            member_cutoff=None, max_clusters=None):
        """Components of the adopted (gathered) CSR: every rank settles the candidate edges of its row block,
        all-gather of the parent arrays, merge + ranking + labels (``cnnb200_fit_sharded_end``); ``labels`` may be
        ``None`` on ranks whose caller does not want them."""
        ctx = self.ctx
        n = int(bounds[-1])
        b, e = block_of(bounds, self.rank)
        ctx.components_begin(n_member_cutoff, similarity_cutoff)
        pptr = ctx.components_parents_ptr()
        key = (pptr, n)
        if getattr(self, "_parents_key", None) != key:
            self._parents_key, self._parents = key, tensor_from_ptr(pptr, (n,), "<i4", self.device)
        gathered = self._gather_buffer(n)

        def gather():
            if gathered.is_cuda:
                dist.all_gather_into_tensor(gathered, self._parents, group=self.group)
            else:
                dist.all_gather(list(gathered.unbind(0)), self._parents, group=self.group)

        if self.last.get("two_phase") and self.world > 1:
            # lists without index locality: most candidate edges cross the row blocks, so the forests of the sampling
            # rounds are exchanged and merged BEFORE the scan -- each rank then skips what any rank already joined
            ctx.components_phase(b, e, 1)
            gather()
            ctx.components_merge(gathered, self.world)
            ctx.components_phase(b, e, 2)
        else:
            ctx.components_edges(b, e)
        gather()
        return ctx.fit_sharded_end(gathered, self.world, labels, sort_by_size, current_start, member_cutoff, max_clusters)

    def upload_points(self, x):
        """Host coordinates (every rank sees the same array) -> replicated device copy.  Each rank uploads only its
        1/G of the rows over its own PCIe link and the slices are all-gathered over NVLink: G concurrent uploads of
        the whole array compete for the host's memory bandwidth (C2 on 8 GPUs: 2.96 ms end to end that way)."""
        from . import _lib

        G, n, d = self.world, x.shape[0], x.shape[1]
        key = (x.ctypes.data, x.nbytes)
        if key not in self._registered:          # page-locked in place, once per array
            self._registered[key] = _lib.host_register(x)
        rows = (n + G - 1) // G
        xd = torch.empty((rows * G, d), dtype=torch.float64, device=self.device)
        r0 = min(self.rank * rows, n)
        r1 = min(r0 + rows, n)
        if r1 > r0:
            xd[r0:r1].copy_(torch.from_numpy(x[r0:r1]), non_blocking=True)
        mine = xd[self.rank * rows:(self.rank + 1) * rows]
        dist.all_gather_into_tensor(xd, mine, group=self.group)    # in place: `mine` is this rank's slice of `xd`
        return xd[:n]

    def _window_buffer(self, length):
        buf = getattr(self, "_windows", None)
        if buf is None or buf.shape[1] < length or buf.shape[1] > 2 * length:
            buf = self._windows = torch.empty((self.world, length + length // 8), dtype=torch.int32, device=self.device)
        # a [world, length] view must be contiguous for the collective: carve it from the flat storage
        return buf.view(-1)[:self.world * length].view(self.world, length)

    def _gather_buffer(self, n):
        buf = getattr(self, "_gathered", None)
        if buf is None or buf.shape[1] != n:
            buf = self._gathered = torch.empty((self.world, n), dtype=torch.int32, device=self.device)
        return buf

    def fit_points(self, x, radius_cutoff, n_member_cutoff, similarity_cutoff, labels, current_start=1,
                   metric="euclidean_r", sort_by_size=True, member_cutoff=None, max_clusters=None, exchange=False, **_):
        """The whole row-sharded fit: two C calls per rank (``cnnb200_fit_sharded_begin`` / ``_end``) with ONE
        collective between them, the all-gather of the parent arrays.  ``labels`` may be ``None`` on ranks whose
        caller does not want them (they only take part in the merge); ``exchange=True`` runs the CSR all-gather
        layout of the precomputed-neighbourhood path on coordinates instead (kept for comparison)."""
        if isinstance(x, np.ndarray) and self.world > 1 and x.flags.c_contiguous and x.dtype == np.float64:
            x = self.upload_points(x)
        if exchange:
            bounds = self.neighbours_from_points(x, radius_cutoff, metric, exchange=True)
            return self.fit(bounds, n_member_cutoff, similarity_cutoff, labels, current_start, sort_by_size,
                            member_cutoff, max_clusters)
        ctx = self.ctx
        n = x.shape[0]
        bounds, pptr = ctx.fit_sharded_begin(x, radius_cutoff, n_member_cutoff, similarity_cutoff, self.rank, self.world,
                                             metric)
        key = (pptr, n)
        if getattr(self, "_parents_key", None) != key:
            self._parents_key, self._parents = key, tensor_from_ptr(pptr, (n,), "<i4", self.device)
        if not self.windows:
            gathered = self._gather_buffer(n)
            dist.all_gather_into_tensor(gathered, self._parents, group=self.group)
            self.last = dict(bounds=bounds)
            return ctx.fit_sharded_end(gathered, self.world, labels, sort_by_size, current_start, member_cutoff, max_clusters)
        # windows=True: a rank's parent array says something only for the rows it built (block + halo), so the ranks
        # exchange windows of the longest such range instead of whole arrays (8 GPUs on C2: 0.19 n per rank).  Same-box
        # A/B of bench.py on C2: 8 GPUs 1.00 / 1.02 ms against 1.10 ms per fit (merge + flatten 0.087 -> 0.064 ms, time
        # outside any stage 0.24 -> 0.17 ms); 2 GPUs 1.63 against 1.60 ms (the windows are 0.6 n there and the extra
        # host work between the two C calls costs more than the bytes) -- hence the default: from 8 ranks on.
        ranges = ctx.fit_sharded_ranges(self.world)
        plan = getattr(self, "_win_plan", None)
        if plan is None or plan[0] != key or not np.array_equal(plan[1], ranges):   # repeated fits: nothing to redo
            length, starts = plan_windows(ranges, n)
            gathered = self._gather_buffer(n) if length == n else self._window_buffer(length)
            s0 = int(starts[self.rank])
            plan = self._win_plan = (key, ranges, length, starts, gathered, self._parents[s0:s0 + length])
        _, _, length, starts, gathered, mine = plan
        dist.all_gather_into_tensor(gathered, mine, group=self.group)
        self.last = dict(bounds=bounds, ranges=ranges, window=length)
        return ctx.fit_sharded_end_windows(gathered, self.world, starts, length, labels, sort_by_size, current_start,
                                           member_cutoff, max_clusters)
