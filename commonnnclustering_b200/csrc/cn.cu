// cn.cu -- K5/K6: common-nearest-neighbour similarity over sorted CSR rows + concurrent
// union-find over the accepted edges.
//
// Bulk replacement of FitterExtBFS._fit (reference _fit.pyx:408-491) and of the per-pair
// SimilarityCheckerExt*._check / ._get (_types.pyx:2580-2610, :2654-2706, :2708-2753).
//
// The reference grows one cluster at a time with a queue and a `consider` byte mask; its result
// is exactly "connected components of the graph whose edges are the pairs of core points (i, j)
// with j in N(i) and |N(i) ∩ N(j)| >= similarity_cutoff" (validated against
// the reference).  Components do not depend on the order edges are examined in, so here all
// candidate edges are examined concurrently and merged with a lock-free union-find that always
// hooks the larger root under the smaller one (the final root of a component is its smallest
// member, independent of scheduling and of how rows are sharded over GPUs).
//
// Like the reference's BFS -- which never re-tests a point once `consider` is cleared
// (_fit.pyx:458-459) -- an edge whose endpoints are already connected is not evaluated:
// a sampling pass links every core point with a few of its neighbours first (Afforest-style),
// after which most candidate edges fall inside one tree and cost two parent look-ups.
#include "common.cuh"
#include "uf.cuh"

// Warp-cooperative |A ∩ B| over ascending int32 lists in global memory (L1/L2 resident: a row
// and the rows of its neighbours sit next to each other in the cell-sorted CSR).
//   CHECK = true : early exit, returns 1 as soon as `cutoff` common elements are found, 0 when
//                  that has become impossible (SimilarityChecker._check semantics).
//   CHECK = false: returns the full count (SimilarityChecker._get semantics).
// The shorter list is streamed 32 consecutive entries at a time (one coalesced load); each lane
// binary-searches its entry in the not-yet-consumed tail of the longer list.
template <bool CHECK>
__device__ __forceinline__ int warp_common(const int *__restrict__ a, int na, const int *__restrict__ b,
                                           int nb, int cutoff, int lane) {
    if (CHECK && cutoff <= 0) return 1;
    if (na > nb) {
        const int *t = a; a = b; b = t;
        int tn = na; na = nb; nb = tn;
    }
    if (CHECK && na < cutoff) return 0;
    int count = 0, lo = 0;
    for (int base = 0; base < na; base += 32) {
        int k = base + lane;
        int v = k < na ? a[k] : 0x7fffffff;
        // lower_bound of v in b[lo, nb)
        int l = lo, h = nb;
        while (l < h) {
            int mid = (l + h) >> 1;
            if (b[mid] < v) l = mid + 1; else h = mid;
        }
        bool found = k < na && l < nb && b[l] == v;
        count += __popc(__ballot_sync(0xffffffffu, found));
        if (CHECK) {
            if (count >= cutoff) return 1;
            if (count + (na - base - 32) < cutoff) return 0;
        }
        lo = __shfl_sync(0xffffffffu, l, 31);
        if (lo >= nb) break;
    }
    return CHECK ? 0 : count;
}

// One-thread form of the check for short lists: a branch-free two-pointer merge with early exit,
// the direct analogue of SimilarityCheckerExtScreensorted._check (_types.pyx:2654-2706).  Used
// where every lane of a warp has its OWN pair of rows to compare (32 independent merges in
// flight per warp hide the load latency that a single cooperative intersection cannot).
__device__ __forceinline__ int lane_check(const int *__restrict__ a, int na, const int *__restrict__ b, int nb,
                                          int cutoff) {
    if (cutoff <= 0) return 1;
    if (na < cutoff || nb < cutoff) return 0;
    int ia = 0, ib = 0, count = 0;
    if (na > 192 || nb > 192) {
        // long lists (C1: 10^3 entries per row): start the merge where the index ranges begin to overlap --
        // one binary search instead of a walk of dependent loads through the non-overlapping head
        const int a0 = a[0], b0 = b[0];
        const int *s = a0 < b0 ? a : b;
        const int target = a0 < b0 ? b0 : a0;
        int l = 0, h = a0 < b0 ? na : nb;
        while (l < h) {
            const int mid = (l + h) >> 1;
            if (s[mid] < target) l = mid + 1; else h = mid;
        }
        if (a0 < b0) ia = l; else ib = l;
    }
    while (ia < na && ib < nb) {
        const int va = a[ia], vb = b[ib];
        count += va == vb ? 1 : 0;
        ia += va <= vb ? 1 : 0;
        ib += vb <= va ? 1 : 0;
        if (count >= cutoff) return 1;
    }
    return 0;
}

// the same on bit-packed rows (packed.cu): ascending (word index, mask) pairs -- 6-7x fewer steps on cell-sorted data
__device__ __forceinline__ int lane_check_packed(const int2 *__restrict__ a, int na, const int2 *__restrict__ b, int nb,
                                                 int cutoff) {
    int ia = 0, ib = 0, count = 0;
    while (ia < na && ib < nb) {
        const int2 ea = a[ia], eb = b[ib];
        if (ea.x == eb.x) count += __popc((unsigned)ea.y & (unsigned)eb.y);
        ia += ea.x <= eb.x ? 1 : 0;
        ib += eb.x <= ea.x ? 1 : 0;
        if (count >= cutoff) return 1;
    }
    return 0;
}

// parent[p] = p for core points, -1 for the others (so one load answers "core?" and "which tree?")
__global__ void comp_init_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ perm,
                                 const uint8_t *__restrict__ consider, ll n, ll n_member_cutoff,
                                 int *parent, uint8_t *core) {
    ll p = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (p >= n) return;
    ll deg = row_ptr[p + 1] - row_ptr[p];
    bool c = deg > n_member_cutoff;  // `_enough`: n_points > cutoff (_types.pyx:1408-1411)
    if (c && consider) c = consider[perm ? perm[p] : p] != 0;
    core[p] = c ? 1 : 0;
    parent[p] = c ? (int)p : -1;
}

#define MAX_SAMPLES 2

// Sampling pass (Afforest-style).  A warp takes 32 consecutive rows; everything that is a chain of
// dependent loads per row (core flag, row bounds, picking the sampled neighbour, its core flag and
// row bounds) is done with one row per LANE, so 32 such chains are in flight per warp; only the
// intersection itself is warp-cooperative.  Sample s = 0 is the first entry above i (the nearest
// point in the cell-sorted order), s = 1 the last entry of the row.  Accepted edges are recorded
// and linked afterwards by comp_union_kernel with one thread per edge -- no union-find traffic here.
__global__ void __launch_bounds__(256)
comp_sample_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col_idx, ll row_begin, ll row_end,
                   int cutoff, int n_samples, int *parent, int *__restrict__ edge_out,
                   int *__restrict__ above, unsigned long long *stats, int fuse_union,
                   const int2 *__restrict__ wrows, const int *__restrict__ wcnt, int rpw) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    unsigned long long n_checks = 0, n_acc = 0;
    // rpw (<= 32) rows per warp tile: fewer when the rows are few and long, so that every SM still gets tiles
    for (ll i0 = row_begin + warp * rpw; i0 < row_end; i0 += nwarps * rpw) {
        const ll i = i0 + lane;
        const bool valid = lane < rpw && i < row_end;
        const bool core_i = valid && parent[i] >= 0;
        const ll rb = valid ? row_ptr[i] : 0;
        const int m = valid ? (int)(row_ptr[i + 1] - rb) : 0;
        int first_above = -1;
        int jacc[MAX_SAMPLES];
#pragma unroll
        for (int smp = 0; smp < MAX_SAMPLES; ++smp) jacc[smp] = -1;
#pragma unroll
        for (int smp = 0; smp < MAX_SAMPLES; ++smp) {
            if (smp >= n_samples) break;
            int j = -1;
            if (core_i && m > 0) {
                if (smp == 0) {
                    int l = 0, h = m;  // first entry > i
                    while (l < h) {
                        int mid = (l + h) >> 1;
                        if (col_idx[rb + mid] <= (int)i) l = mid + 1; else h = mid;
                    }
                    if (l < m) j = col_idx[rb + l];
                    first_above = j;
                    above[i] = l;  // position of the first entry > i: the edge pass starts there
                } else {
                    j = col_idx[rb + m - 1];
                    if (j <= (int)i || j == first_above) j = -1;
                }
            }
            bool ok = j >= 0 && parent[j] >= 0;
            const ll jb = ok ? row_ptr[j] : 0;
            const int mj = ok ? (int)(row_ptr[j + 1] - jb) : 0;
            int acc = 0;
            if (ok) {
                if (cutoff <= 0) acc = 1;
                else if (m < cutoff || mj < cutoff) acc = 0;
                else if (wrows) acc = lane_check_packed(wrows + rb, wcnt[i], wrows + jb, wcnt[j], cutoff);
                else acc = lane_check(col_idx + rb, m, col_idx + jb, mj, cutoff);
            }
            n_checks += __popc(__ballot_sync(full, ok));
            n_acc += __popc(__ballot_sync(full, acc != 0));
            jacc[smp] = acc ? j : -1;
            if (!fuse_union && valid) edge_out[(i - row_begin) * n_samples + smp] = jacc[smp];
        }
        if (fuse_union) {
            // link right here (see comp_union_kernel for the run-head trick): other warps keep the SM busy
            // while this one waits for its atomics
            const unsigned linked = __ballot_sync(full, jacc[0] == (int)i + 1 && lane < 31);
            const unsigned below = ~linked & ((1u << lane) - 1u);
            const int head = below ? 32 - __clz(below) : 0;
            const int hi = (int)(i - lane + head);
#pragma unroll
            for (int smp = 0; smp < MAX_SAMPLES; ++smp)
                if (jacc[smp] >= 0) uf_union(parent, hi, jacc[smp]);
        }
    }
    if (lane == 0) {
        if (n_checks) atomicAdd(stats + ST_CHECKS, n_checks);
        if (n_acc) atomicAdd(stats + ST_ACCEPTED, n_acc);
    }
}

// One thread per row, all samples of the row.  Sample 0 is usually the next row (j == i + 1); linking
// every such pair concurrently would build parent chains as long as a dense run of rows, which the
// later finds then have to walk.  A run of consecutive linked rows inside the warp's 32-row tile is
// therefore collapsed first: every edge of a row is applied to the run's FIRST row (its `head`, found
// with a ballot), which is connected to the row by the accepted edges of the lanes in between.
__global__ void comp_union_kernel(const int *__restrict__ edge_out, ll row_begin, ll rows, int n_samples,
                                  int *parent) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    ll t = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    const bool valid = t < rows;
    const ll i = row_begin + t;
    const int j0 = valid ? edge_out[t * n_samples] : -1;
    const unsigned linked = __ballot_sync(full, j0 == (int)i + 1 && lane < 31);  // bit l: row l ~ row l+1
    // lowest lane h with bits h .. lane-1 all set
    const unsigned below = ~linked & ((1u << lane) - 1u);
    const int head = below ? 32 - __clz(below) : 0;
    if (!valid) return;
    const int hi = (int)(i - lane + head);
    if (j0 >= 0) uf_union(parent, hi, j0);
    for (int s = 1; s < n_samples; ++s) {
        const int j = edge_out[t * n_samples + s];
        if (j >= 0) uf_union(parent, hi, j);
    }
}

// Second sampling round, on flattened parents: the first round leaves chains of consecutive rows (trees that follow
// the x-rows of cells); what joins them are edges to rows far away in the index order.  Trying such an edge from
// EVERY row of a chain repeats the same union hundreds of times in dense regions (C2: 0.67 M checks for 0.03 M
// unions), so only every `stride`-th row of a tree (counted from its root) tries, and only when the far neighbour
// sits in another tree: the LAST entry of the row (the highest cell row of its window) and the entry a quarter into
// the upper part (the next cell row of the same layer).  The edges go to the work list and are evaluated by
// comp_worklist_kernel (whole-warp intersections: the two rows of such an edge overlap only partly and are not in L1,
// a per-lane merge would walk them with one dependent global load per step).
__global__ void __launch_bounds__(256)
comp_far_emit_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col_idx, ll row_begin, ll row_end, int stride,
                     const int *__restrict__ parent, const int *__restrict__ above, int2 *__restrict__ worklist,
                     unsigned long long *wl_count, int n_far) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const ll i = row_begin + blockIdx.x * (ll)blockDim.x + threadIdx.x;
    const int ri = i < row_end ? parent[i] : -1;
    const bool sel = ri >= 0 && ((int)i - ri) % stride == 0;
    if (!__any_sync(full, sel)) return;
    const ll rb = sel ? row_ptr[i] : 0;
    const int m = sel ? (int)(row_ptr[i + 1] - rb) : 0;
    const int l = sel ? above[i] : 0;
    int j0 = -1, j1 = -1;
    if (sel && l < m) {
        j0 = col_idx[rb + m - 1];
        if (n_far > 1) j1 = col_idx[rb + l + ((m - l) >> 2)];
        if (j1 == j0) j1 = -1;
    }
    const int r0 = j0 >= 0 ? parent[j0] : -1, r1 = j1 >= 0 ? parent[j1] : -1;
    const bool e0 = r0 >= 0 && r0 != ri, e1 = r1 >= 0 && r1 != ri && r1 != r0;
    const unsigned b0 = __ballot_sync(full, e0), b1 = __ballot_sync(full, e1);
    const int tot = __popc(b0) + __popc(b1);
    if (!tot) return;
    unsigned long long pos = 0;
    if (lane == 0) pos = atomicAdd(wl_count, (unsigned long long)tot);
    pos = __shfl_sync(full, pos, 0);
    // the two edges of a row next to each other: the work-list kernel builds the row's window once for both
    const unsigned long long mine = pos + __popc(b0 & lt) + __popc(b1 & lt);
    if (e0) worklist[mine] = make_int2((int)i, j0);
    if (e1) worklist[mine + (e0 ? 1 : 0)] = make_int2((int)i, j1);
}

__global__ void comp_compress_kernel(int *parent, ll n) {
    ll p = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (p >= n) return;
    if (ld_parent(parent, (int)p) < 0) return;
    int r = uf_find_ro(parent, (int)p);
    parent[p] = r;
}

// ---- shared-memory row window -----------------------------------------------------------------------
// Rows of a cell-sorted CSR are clustered: a row with 10^2 entries spans a few 10^4 indices.  A warp
// that has to intersect ONE row N(i) with MANY rows N(j) turns N(i) into a bitmap over the index
// window [first, last] of the row in shared memory; every element of a streamed N(j) then costs one
// probe (subtract, range check, word load, shift) instead of a binary search, and the loads of N(j)
// are coalesced.  The kernels built on it are bound by memory latency, not by issue slots, so the
// first WIN_PF chunks of the NEXT row to stream are fetched into registers while the current one is
// being probed.  A row whose window does not fit (span == 0) falls back to warp_common.
#define WIN_BITS 65536  // window bits per warp (8 KB)
#define WIN_WARPS 4     // warps per CTA (32 KB of windows)
#define WIN_PF 5        // chunks (of 32 entries) prefetched per streamed row

struct RowWin {
    unsigned *bits;  // cap / 32 words owned by the warp; all zero between rows
    int first;
    unsigned span;  // 0: no window
    unsigned cap = WIN_BITS;  // window bits available to the warp
};

__device__ __forceinline__ void win_set(RowWin &w, const int *__restrict__ row, int m, int lane) {
    w.first = row[0];
    const unsigned span = (unsigned)(row[m - 1] - w.first) + 1u;
    w.span = span <= w.cap ? span : 0u;
    if (w.span)
        for (int k = lane; k < m; k += 32) {
            const unsigned idx = (unsigned)(row[k] - w.first);
            atomicOr(&w.bits[idx >> 5], 1u << (idx & 31));
        }
    __syncwarp();
}

__device__ __forceinline__ void win_clear(RowWin &w, const int *__restrict__ row, int m, int lane) {
    __syncwarp();
    if (w.span)
        for (int k = lane; k < m; k += 32) w.bits[(unsigned)(row[k] - w.first) >> 5] = 0;
    w.span = 0;
    __syncwarp();
}

__device__ __forceinline__ bool win_ready(const RowWin &w) { return w.span != 0; }

// 1 when v is an entry of the windowed row.  v = -1 (padding) never is: first >= 0 puts it out of range.
__device__ __forceinline__ int win_probe(const RowWin &w, int v) {
    const unsigned idx = (unsigned)(v - w.first);
    return idx < w.span ? (int)((w.bits[idx >> 5] >> (idx & 31)) & 1u) : 0;
}

__device__ __forceinline__ void row_prefetch(int (&buf)[WIN_PF], const int *__restrict__ rj, int mj, int lane) {
#pragma unroll
    for (int u = 0; u < WIN_PF; ++u) {
        const int t = lane + 32 * u;
        buf[u] = t < mj ? rj[t] : -1;
    }
}

// |windowed row ∩ N(j)|, N(j)'s first WIN_PF chunks already in registers.  (Measured and dropped, r2: keeping a row
// WITHOUT index locality as a sorted array in the window words and probing it by binary search in shared memory --
// on C4 the global-memory search of warp_common, which shrinks its range as it goes and stops early, is faster:
// 9.5 ms against 11.2 ms per fit.)
__device__ __forceinline__ int row_count(const int (&buf)[WIN_PF], const int *__restrict__ rj, int mj, int lane,
                                         const RowWin &w) {
    int c = 0;
#pragma unroll
    for (int u = 0; u < WIN_PF; ++u) c += win_probe(w, buf[u]);
#pragma unroll 4
    for (int t = lane + 32 * WIN_PF; t < mj; t += 32) c += win_probe(w, rj[t]);
    return __reduce_add_sync(0xffffffffu, c);
}

// The same as a CHECK (|.| >= cutoff?) with the reference's early exits (SimilarityCheckerExtScreensorted._check,
// _types.pyx:2654-2706): after the WIN_PF prefetched chunks the answer is often known -- enough common neighbours
// already (long rows of dense regions overlap almost completely), or too few entries left to get there -- and the
// rest of N(j), the part that still has to come from memory, is never read.
__device__ __forceinline__ int row_check(const int (&buf)[WIN_PF], const int *__restrict__ rj, int mj, int lane,
                                         const RowWin &w, int cutoff) {
    int c = 0;
#pragma unroll
    for (int u = 0; u < WIN_PF; ++u) c += win_probe(w, buf[u]);
    if (mj <= 32 * WIN_PF) return __reduce_add_sync(0xffffffffu, c) >= cutoff;
    int total = __reduce_add_sync(0xffffffffu, c);
    if (total >= cutoff) return 1;
    if (total + (mj - 32 * WIN_PF) < cutoff) return 0;
    c = 0;
#pragma unroll 4
    for (int t = lane + 32 * WIN_PF; t < mj; t += 32) c += win_probe(w, rj[t]);
    return total + __reduce_add_sync(0xffffffffu, c) >= cutoff;
}

// One streamed row in flight: where it is, which lane of the current 32-edge block it belongs to, and
// its first WIN_PF chunks.  Three of them rotate (fetch two ahead, consume the oldest), so that two
// memory round trips overlap the probing of one row.
struct RowFetch {
    int buf[WIN_PF];
    const int *rj;
    int mj;
    int src;  // lane that owns the edge; -1: nothing fetched
};

// takes the next edge out of `todo` (bj / mj are the per-lane row bounds of the block's edges)
__device__ __forceinline__ void fetch_next(RowFetch &f, unsigned &todo, const int *__restrict__ col_idx, ll bj, int mj,
                                           bool windowed, int lane) {
    if (!todo) { f.src = -1; return; }
    f.src = __ffs(todo) - 1;
    todo &= todo - 1;
    f.rj = col_idx + __shfl_sync(0xffffffffu, bj, f.src);
    f.mj = __shfl_sync(0xffffffffu, mj, f.src);
    if (windowed) row_prefetch(f.buf, f.rj, f.mj, lane);
}

// Edge scan.  Runs on flattened (root-pointing) parents, which do not change during this kernel:
// a candidate edge (i < j, both core) is settled for free when parent[i] == parent[j]; the others
// go to a work list.  A warp takes 32 consecutive rows (bounds, root and the position of the first
// entry above i are fetched one row per lane) and then streams the upper part of each core row,
// 32 entries per step.
__global__ void __launch_bounds__(256)
comp_scan_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col_idx, ll row_begin, ll row_end,
                 const int *__restrict__ above, const int *__restrict__ parent, int2 *__restrict__ worklist,
                 unsigned long long *wl_count, unsigned long long *stats, int tile_rows) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    unsigned long long n_edges = 0, n_need = 0;
    constexpr int R = 8;  // chunks in flight per warp
    // tile_rows (<= 32) rows per warp tile: fewer when the rows are long, so that few long rows (C1: 10^4 rows of
    // 10^3 entries) still spread over all SMs
    for (ll i0 = row_begin + warp * tile_rows; i0 < row_end; i0 += nwarps * tile_rows) {
        const ll il = i0 + lane;
        const bool lv = il < row_end && lane < tile_rows;
        // independent loads (no branch on the root): one memory round trip for the whole tile header
        const int rl = lv ? parent[il] : -1;
        const ll e_raw = lv ? row_ptr[il + 1] : 0, b_raw = lv ? row_ptr[il] : 0;
        const int ab = lv ? above[il] : 0;
        ll bl = 0, el = 0;
        if (rl >= 0) {
            el = e_raw;
            bl = b_raw + ab;
        }
        // the tile's work items are its (row, chunk of 32 entries) pairs; incl = inclusive prefix of the
        // chunk counts, so item t belongs to row popc(ballot(incl <= t))
        const int nch = (int)((el - bl + 31) >> 5);
        int incl = nch;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(full, incl, o);
            if (lane >= o) incl += t;
        }
        const int excl = incl - nch;
        const int total = __shfl_sync(full, incl, 31);
        for (int t0 = 0; t0 < total; t0 += R) {
            int rr[R], r_i[R], jv[R], pjv[R];
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const int t = t0 + u;
                rr[u] = -1;
                jv[u] = -1;
                if (t < total) {
                    const int r = __popc(__ballot_sync(full, incl <= t));
                    rr[u] = r;
                    r_i[u] = __shfl_sync(full, rl, r);
                    const ll k = __shfl_sync(full, bl, r) + 32 * (t - __shfl_sync(full, excl, r)) + lane;
                    if (k < __shfl_sync(full, el, r)) jv[u] = col_idx[k];
                }
            }
#pragma unroll
            for (int u = 0; u < R; ++u) pjv[u] = jv[u] >= 0 ? parent[jv[u]] : -1;
            // one reservation in the work list for the whole batch of R chunks
            unsigned nmv[R];
            int tot = 0;
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const bool cand = rr[u] >= 0 && pjv[u] >= 0;
                nmv[u] = __ballot_sync(full, cand && pjv[u] != r_i[u]);
                n_edges += __popc(__ballot_sync(full, cand));
                tot += __popc(nmv[u]);
            }
            if (tot) {
                unsigned long long pos = 0;
                if (lane == 0) pos = atomicAdd(wl_count, (unsigned long long)tot);
                pos = __shfl_sync(full, pos, 0);
#pragma unroll
                for (int u = 0; u < R; ++u) {
                    if ((nmv[u] >> lane) & 1u) worklist[pos + __popc(nmv[u] & lt)] = make_int2((int)i0 + rr[u], jv[u]);
                    pos += __popc(nmv[u]);
                }
                n_need += tot;
            }
        }
    }
    if (lane == 0) {
        if (n_edges) atomicAdd(stats + ST_EDGES, n_edges);
        if (n_edges - n_need) atomicAdd(stats + ST_SKIPPED, n_edges - n_need);
    }
}

// Work-list pass: the edges the scan could not settle, 32 per warp step.  The finds (with path
// halving) and the row bounds are fetched one edge per LANE; edges whose ends meanwhile share a root
// are dropped.  The rest is evaluated by the whole warp, one edge after the other, against the
// shared-memory window of row i -- the scan emits the work list row by row, so consecutive entries
// mostly share i and the window is rebuilt only when i changes -- while the entries of the next
// N(j) are already in flight.
template <int WBITS, int MINB>
__global__ void __launch_bounds__(WIN_WARPS * 32, MINB)
comp_worklist_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col_idx, const int2 *__restrict__ worklist,
                     const unsigned long long *__restrict__ wl_count, unsigned long long *chunk_counter, int chunk,
                     int cutoff, int *parent, unsigned long long *stats) {
    __shared__ unsigned win_all[WIN_WARPS][WBITS / 32];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    RowWin w;
    w.bits = win_all[threadIdx.x >> 5];
    w.first = 0;
    w.span = 0;
    w.cap = WBITS;
    for (int t = lane; t < WBITS / 32; t += 32) w.bits[t] = 0;
    __syncwarp();
    const ll count = (ll)*wl_count;
    {   // short lists: at least four chunks per warp
        const ll per_warp = count / (((ll)gridDim.x * blockDim.x >> 5) * 32 * 4);
        chunk = per_warp < 1 ? 1 : (per_warp < chunk ? (int)per_warp : chunk);
    }
    unsigned long long n_checks = 0, n_acc = 0, n_skip = 0;
    // A warp takes `chunk` consecutive batches of 32 edges at a time (handed out in list order by a counter) and works
    // through them one after the other with fresh finds per batch.  The scan emits the list row by row, so the many
    // edges between the same two trees sit next to each other: taken by ONE warp in sequence, the first accepted edge
    // settles the rest (two finds each); spread over concurrently running warps they would all be evaluated.
    for (;;) {
        ll c0 = 0;
        if (lane == 0) c0 = (ll)atomicAdd(chunk_counter, 1ull);
        c0 = __shfl_sync(full, c0, 0) * chunk * 32;
        if (c0 >= count) break;
        const ll c1 = c0 + (ll)chunk * 32 < count ? c0 + (ll)chunk * 32 : count;
    for (ll e0 = c0; e0 < c1; e0 += 32) {
        const ll e = e0 + lane;
        const bool valid = e < count;
        const int2 ed = valid ? worklist[e] : make_int2(0, 0);
        const int ri = valid ? uf_find(parent, ed.x) : 0;
        const int rj = valid ? uf_find(parent, ed.y) : 0;
        const bool need = valid && ri != rj;
        const ll bi = need ? row_ptr[ed.x] : 0, bj = need ? row_ptr[ed.y] : 0;
        const int mi = need ? (int)(row_ptr[ed.x + 1] - bi) : 0, mj = need ? (int)(row_ptr[ed.y + 1] - bj) : 0;
        int acc = need && cutoff <= 0 ? 1 : 0;
        unsigned work = __ballot_sync(full, need && cutoff > 0 && mi >= cutoff && mj >= cutoff);
        const unsigned nd = __ballot_sync(full, need);
        n_skip += __popc(__ballot_sync(full, valid)) - __popc(nd);
        n_checks += __popc(nd);
        // software pipeline over the edges of `work`
        int cur_x = -1, cur_m = 0;
        const int *cur_row = nullptr;
        // an accepted edge settles every other edge of the batch between the same two trees (a row's upper neighbours
        // mostly sit in a few trees): those are dropped without a check
        bool dead = false;
        auto consume = [&](const RowFetch &f) {
            if (__shfl_sync(full, dead, f.src)) return;
            const int x = __shfl_sync(full, ed.x, f.src);
            if (x != cur_x) {
                if (cur_x >= 0) win_clear(w, cur_row, cur_m, lane);
                cur_x = x;
                cur_row = col_idx + __shfl_sync(full, bi, f.src);
                cur_m = __shfl_sync(full, mi, f.src);
                win_set(w, cur_row, cur_m, lane);
            }
            int ok;
            if (win_ready(w)) ok = row_check(f.buf, f.rj, f.mj, lane, w, cutoff);
            else ok = warp_common<true>(cur_row, cur_m, f.rj, f.mj, cutoff, lane);
            if (lane == f.src) acc = ok;
            if (ok) {
                const int ra = __shfl_sync(full, ri, f.src), rb = __shfl_sync(full, rj, f.src);
                if (lane != f.src && need && ((ri == ra && rj == rb) || (ri == rb && rj == ra))) dead = true;
            }
        };
        RowFetch f0, f1, f2;
        fetch_next(f0, work, col_idx, bj, mj, true, lane);
        fetch_next(f1, work, col_idx, bj, mj, true, lane);
        for (;;) {
            if (f0.src < 0) break;
            fetch_next(f2, work, col_idx, bj, mj, true, lane);
            consume(f0);
            if (f1.src < 0) break;
            fetch_next(f0, work, col_idx, bj, mj, true, lane);
            consume(f1);
            if (f2.src < 0) break;
            fetch_next(f1, work, col_idx, bj, mj, true, lane);
            consume(f2);
        }
        if (cur_x >= 0) win_clear(w, cur_row, cur_m, lane);
        if (acc) uf_union(parent, ri, rj);
        n_acc += __popc(__ballot_sync(full, acc != 0));
        const int n_dead = __popc(__ballot_sync(full, dead && acc == 0));
        n_checks -= n_dead;
        n_skip += n_dead;
    }
    }
    if (lane == 0) {
        if (n_checks) atomicAdd(stats + ST_CHECKS, n_checks);
        if (n_acc) atomicAdd(stats + ST_ACCEPTED, n_acc);
        if (n_skip) atomicAdd(stats + ST_SKIPPED, n_skip);
    }
}

// above[i] = position of the first entry > i in row i (only needed when the sampling pass is off)
__global__ void comp_above_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col_idx, ll row_begin,
                                  ll row_end, int *__restrict__ above) {
    ll i = row_begin + blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (i >= row_end) return;
    const ll rb = row_ptr[i];
    int l = 0, h = (int)(row_ptr[i + 1] - rb);
    while (l < h) {
        int mid = (l + h) >> 1;
        if (col_idx[rb + mid] <= (int)i) l = mid + 1; else h = mid;
    }
    above[i] = l;
}

// ---- components from precomputed common-neighbour counts (parameter sweeps) ---------------------------
// counts[k] of cnnb200_cn_count answer every similarity cutoff at once: slot k (row i, column j > i)
// links i and j when both have more than n_member_cutoff entries and counts[k] >= similarity_cutoff.
// Same two-step shape as the check-mode pass: link every core row with its first accepted neighbour
// above it, flatten, then look only at the candidate edges whose ends still sit in different trees.
__device__ __forceinline__ bool count_accepts(unsigned cnt, unsigned cutoff) {
    return cutoff == 0 || (cnt != 0xffffffffu && cnt >= cutoff);
}

__global__ void __launch_bounds__(256)
counts_sample_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col_idx, const unsigned *__restrict__ counts,
                     ll n, unsigned cutoff, const int *__restrict__ parent, int *__restrict__ edge_out,
                     int *__restrict__ above, int n_samples) {
    ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int j_out = -1, j_last = -1;
    if (parent[i] >= 0) {
        const ll rb = row_ptr[i];
        const int m = (int)(row_ptr[i + 1] - rb);
        int l = 0, h = m;  // first entry > i
        while (l < h) {
            int mid = (l + h) >> 1;
            if (col_idx[rb + mid] <= (int)i) l = mid + 1; else h = mid;
        }
        above[i] = l;
        int k_first = m;
        for (int k = l; k < m; ++k) {
            const int j = col_idx[rb + k];
            if (parent[j] >= 0 && count_accepts(counts[rb + k], cutoff)) { j_out = j; k_first = k; break; }
        }
        // second sample: the LAST accepted entry of the row (the far end of the neighbourhood in the cell-sorted order)
        for (int k = m - 1; n_samples > 1 && k > k_first; --k) {
            const int j = col_idx[rb + k];
            if (parent[j] >= 0 && count_accepts(counts[rb + k], cutoff)) { j_last = j; break; }
        }
    }
    edge_out[2 * i] = j_out;
    edge_out[2 * i + 1] = j_last;
}

__global__ void __launch_bounds__(256)
counts_scan_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col_idx, const unsigned *__restrict__ counts,
                   ll n, const int *__restrict__ above, unsigned cutoff, int *parent, unsigned long long *stats) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    unsigned long long n_edges = 0, n_acc = 0, n_looked = 0;
    constexpr int R = 8;
    for (ll i0 = warp * 32; i0 < n; i0 += nwarps * 32) {
        const ll il = i0 + lane;
        const int rl = il < n ? ld_parent(parent, (int)il) : -1;
        ll bl = 0, el = 0;
        if (rl >= 0) {
            el = row_ptr[il + 1];
            bl = row_ptr[il] + above[il];
        }
        const int nch = (int)((el - bl + 31) >> 5);
        int incl = nch;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(full, incl, o);
            if (lane >= o) incl += t;
        }
        const int excl = incl - nch;
        const int total = __shfl_sync(full, incl, 31);
        for (int t0 = 0; t0 < total; t0 += R) {
            int rr[R], r_i[R], jv[R], pjv[R];
            unsigned cv[R];  // the counts travel with the columns (same index, coalesced): no dependent load later
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const int t = t0 + u;
                rr[u] = -1;
                jv[u] = -1;
                cv[u] = 0xffffffffu;
                if (t < total) {
                    const int r = __popc(__ballot_sync(full, incl <= t));
                    rr[u] = r;
                    r_i[u] = __shfl_sync(full, rl, r);
                    const ll k = __shfl_sync(full, bl, r) + 32 * (t - __shfl_sync(full, excl, r)) + lane;
                    if (k < __shfl_sync(full, el, r)) { jv[u] = col_idx[k]; cv[u] = counts[k]; }
                }
            }
#pragma unroll
            for (int u = 0; u < R; ++u) pjv[u] = jv[u] >= 0 ? ld_parent(parent, jv[u]) : -1;
#pragma unroll
            for (int u = 0; u < R; ++u) {
                if (rr[u] < 0) break;
                const bool cand = pjv[u] >= 0;
                // parents move while this kernel runs: a stale mismatch only costs a look at the count
                const bool look = cand && pjv[u] != r_i[u];
                bool acc = false;
                if (look) {
                    acc = count_accepts(cv[u], cutoff);
                    if (acc) uf_union(parent, (int)i0 + rr[u], jv[u]);
                }
                n_edges += __popc(__ballot_sync(full, cand));
                n_looked += __popc(__ballot_sync(full, look));
                n_acc += __popc(__ballot_sync(full, acc));
            }
        }
    }
    if (lane == 0) {
        if (n_edges) atomicAdd(stats + ST_EDGES, n_edges);
        if (n_looked) atomicAdd(stats + ST_CHECKS, n_looked);
        if (n_acc) atomicAdd(stats + ST_ACCEPTED, n_acc);
        if (n_edges - n_looked) atomicAdd(stats + ST_SKIPPED, n_edges - n_looked);
    }
}

// A rank that built only the rows around its block sees every other row as empty, hence not core.
// Before merging, rows that ANY rank reports as core (parent >= 0) become singleton trees here too.
__global__ void comp_adopt_core_kernel(const int *__restrict__ others, ll n, ll count, int *parent, uint8_t *core) {
    ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (i >= n || parent[i] >= 0) return;
    for (ll g = 0; g < count; ++g)
        if (others[g * n + i] >= 0) {
            parent[i] = (int)i;
            core[i] = 1;
            return;
        }
}

// merge step of the row-sharded multi-GPU run: union (i, other[i]) for each foreign parent array
__global__ void comp_merge_kernel(const int *__restrict__ others, ll n, ll count, int *parent) {
    ll t = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    ll stride = (ll)gridDim.x * blockDim.x;
    for (; t < n * count; t += stride) {
        int i = (int)(t % n);
        int o = others[t];
        // most entries say what this rank's forest says already (identical after an earlier merge, or the rank's own
        // slice): one comparison instead of two finds
        if (o >= 0 && o != i && o != ld_parent(parent, i)) uf_union(parent, i, o);
    }
}

// The same two steps on WINDOWS of the foreign parent arrays: rank g built the rows [lo_g, hi_g) only (its block plus
// the halo), everything else in its array is -1 and says nothing -- so the ranks exchange `len` entries from
// starts[g] on (len = the longest range, starts[g] <= lo_g), G x len instead of G x n values (8 GPUs on C2: 0.19 n).
__global__ void comp_adopt_core_win_kernel(const int *__restrict__ others, const ll *__restrict__ starts, ll len, ll n,
                                           ll count, int *parent, uint8_t *core) {
    ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (i >= n || parent[i] >= 0) return;
    for (ll g = 0; g < count; ++g) {
        const ll k = i - starts[g];
        if (k >= 0 && k < len && others[g * len + k] >= 0) {
            parent[i] = (int)i;
            core[i] = 1;
            return;
        }
    }
}

__global__ void comp_merge_win_kernel(const int *__restrict__ others, const ll *__restrict__ starts, ll len, ll count,
                                      int *parent) {
    ll t = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    ll stride = (ll)gridDim.x * blockDim.x;
    for (; t < len * count; t += stride) {
        const ll g = t / len;
        const int i = (int)(starts[g] + (t - g * len));
        const int o = others[t];
        if (o >= 0 && o != i && o != ld_parent(parent, i)) uf_union(parent, i, o);
    }
}

// ---- count mode -----------------------------------------------------------------------------------
// |N(i) ∩ N(j)| for every candidate edge (SimilarityChecker._get semantics, no early exit).
// One warp per row i: N(i) becomes the shared-memory window, every N(j) with j in N(i), j > i is
// streamed through it (see above).  The row bounds of the 32 edges of a block are fetched by 32
// lanes at once, the counts of the block are written with one coalesced store.  Rows are handed out
// in small batches by an atomic counter: concurrently running warps work on neighbouring rows
// (their N(j) overlap in L1/L2) and the heavy rows of dense regions do not serialise behind a static
// partition.
#define CNW_BATCH 4

__global__ void __launch_bounds__(WIN_WARPS * 32)
cn_count_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col_idx, ll row_begin, ll n, ll n_member_cutoff,
                unsigned *__restrict__ counts, unsigned long long *stats, unsigned long long *row_counter) {
    __shared__ unsigned win_all[WIN_WARPS][WIN_BITS / 32];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    RowWin w;
    w.bits = win_all[threadIdx.x >> 5];
    w.first = 0;
    w.span = 0;
    for (int t = lane; t < WIN_BITS / 32; t += 32) w.bits[t] = 0;
    __syncwarp();
    unsigned long long n_edges = 0, n_fallback = 0;
    for (;;) {
        ll i0 = 0;
        if (lane == 0) i0 = row_begin + (ll)atomicAdd(row_counter, (unsigned long long)CNW_BATCH);
        i0 = __shfl_sync(full, i0, 0);
        if (i0 >= n) break;  // n = end of the row range
        const ll i1 = i0 + CNW_BATCH < n ? i0 + CNW_BATCH : n;
        for (ll i = i0; i < i1; ++i) {
            const ll b = row_ptr[i];
            const int m = (int)(row_ptr[i + 1] - b);
            const int *row = col_idx + b;
            if (m <= n_member_cutoff) {  // not a core row (`_enough`): no slot of it is an edge
                for (int k = lane; k < m; k += 32) counts[b + k] = 0xffffffffu;
                continue;
            }
            win_set(w, row, m, lane);
            if (!win_ready(w)) ++n_fallback;
            for (int k0 = 0; k0 < m; k0 += 32) {
                const int k = k0 + lane;
                const int j = k < m ? row[k] : -1;
                const bool valid = j > (int)i;
                const ll bj = valid ? row_ptr[j] : 0;
                const int mj = valid ? (int)(row_ptr[j + 1] - bj) : 0;
                unsigned todo = __ballot_sync(full, valid && mj > n_member_cutoff);
                unsigned out = 0xffffffffu;
                n_edges += __popc(todo);
                const bool windowed = win_ready(w);
                auto consume = [&](const RowFetch &f) {
                    const int c = windowed ? row_count(f.buf, f.rj, f.mj, lane, w)
                                           : warp_common<false>(row, m, f.rj, f.mj, 0, lane);
                    if (lane == f.src) out = (unsigned)c;
                };
                RowFetch f0, f1, f2;
                fetch_next(f0, todo, col_idx, bj, mj, windowed, lane);
                fetch_next(f1, todo, col_idx, bj, mj, windowed, lane);
                for (;;) {
                    if (f0.src < 0) break;
                    fetch_next(f2, todo, col_idx, bj, mj, windowed, lane);
                    consume(f0);
                    if (f1.src < 0) break;
                    fetch_next(f0, todo, col_idx, bj, mj, windowed, lane);
                    consume(f1);
                    if (f2.src < 0) break;
                    fetch_next(f1, todo, col_idx, bj, mj, windowed, lane);
                    consume(f2);
                }
                if (k < m) counts[b + k] = out;
            }
            win_clear(w, row, m, lane);
        }
    }
    if (lane == 0) {
        if (n_edges) atomicAdd(stats + ST_EDGES, n_edges);
        if (n_fallback) atomicAdd(stats + ST_FALLBACK, n_fallback);
    }
}

// ---- host side ----------------------------------------------------------------------------------
// resident CTAs per SM of a persistent kernel (grid = SMs x this), asked once per kernel
template <typename K>
static int resident_ctas(K kernel, int threads) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    return per_sm;
}

static int fetch_stats(cnnb200_ctx *ctx) {
    unsigned long long *h;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&h));
    CNN_CUDA(cudaMemcpyAsync(h, ctx->b_stats.p, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CNN_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 4; ++k) ctx->stats[k] = (ll)h[k];
    return CNNB200_OK;
}

int cnn_fetch_comp_stats(cnnb200_ctx *ctx) { return fetch_stats(ctx); }

extern "C" int cnnb200_components_begin(cnnb200_ctx *ctx, int64_t n_member_cutoff, int64_t similarity_cutoff,
                                        const uint8_t *consider, int loc) {
    if (!ctx) CNN_FAIL(CNNB200_E_INVALID, "ctx is NULL");

    if (!ctx->have_nb) CNN_FAIL(CNNB200_E_STATE, "components: no neighbourhoods held by this context");
    if (n_member_cutoff < 0 || similarity_cutoff < 0)
        CNN_FAIL(CNNB200_E_INVALID, "components: cutoffs must be >= 0");
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ll n = ctx->n;
    CNN_TRY(cnn_ensure(ctx, ctx->b_parent, sizeof(int) * (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_core, (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_stats, sizeof(unsigned long long) * ST_N));
    CNN_CUDA(cudaMemsetAsync(ctx->b_stats.p, 0, sizeof(unsigned long long) * ST_N, st));
    const uint8_t *d_consider = consider;
    if (consider && loc == CNNB200_HOST) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_consider, (size_t)n));
        CNN_CUDA(cudaMemcpyAsync(ctx->b_consider.p, consider, (size_t)n, cudaMemcpyHostToDevice, st));
        d_consider = ctx->b_consider.as<uint8_t>();
    }
    {
        Span sp(ctx, T_COMP_INIT);
        comp_init_kernel<<<cnn_blocks(n, 256), 256, 0, st>>>(ctx->row_ptr, ctx->perm, d_consider, n,
                                                              n_member_cutoff, ctx->b_parent.as<int>(),
                                                              ctx->b_core.as<uint8_t>());
        CNN_LAUNCHED(ctx, 1);
    }
    CNN_KERNEL_CHECK();
    ctx->n_member_cutoff = n_member_cutoff;
    ctx->similarity_cutoff = similarity_cutoff;
    ctx->comp_packed = false;
    if (ctx->pack_in_fit && similarity_cutoff > 0) CNN_TRY(cnn_pack_rows(ctx, &ctx->comp_packed));
    ctx->have_comp = true;
    return CNNB200_OK;
}

// phases: 1 = the sampling rounds (first entry above every row; far entries from every k-th row of a tree),
//         2 = the scan of all candidate edges + the work list.  A row-sharded run may exchange and merge the forests
//         between the two, so that the scan sees what the other ranks' samples already joined.
static int components_pass(cnnb200_ctx *ctx, int64_t row_begin, int64_t row_end, int phases) {
    if (!ctx) CNN_FAIL(CNNB200_E_INVALID, "ctx is NULL");
    if (!ctx->have_comp) CNN_FAIL(CNNB200_E_STATE, "components_edges: call components_begin first");
    if (row_begin < 0 || row_end > ctx->n || row_begin > row_end)
        CNN_FAIL(CNNB200_E_INVALID, "components_edges: bad row range");
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ll rows = row_end - row_begin;
    if (rows == 0) return CNNB200_OK;
    // a similarity cutoff beyond any possible intersection size links nothing
    int cutoff = ctx->similarity_cutoff > 0x7fffffffll ? 0x7fffffff : (int)ctx->similarity_cutoff;
    unsigned long long *stats = (unsigned long long *)ctx->b_stats.p;
    const int ns = ctx->n_samples < MAX_SAMPLES ? ctx->n_samples : MAX_SAMPLES;
    int *parent = ctx->b_parent.as<int>();
    CNN_TRY(cnn_ensure(ctx, ctx->b_above, sizeof(int) * (size_t)ctx->n));
    int *above = ctx->b_above.as<int>();
    if (!(phases & 1)) {
    } else if (ns > 0) {
        Span sp(ctx, T_COMP_SAMPLE);
        ll slots = rows * ns;
        CNN_TRY(cnn_ensure(ctx, ctx->b_edges, sizeof(int) * (size_t)slots));
        int rpw = 32;
        while (rpw > 8 && rows / rpw < (ll)ctx->sm_count * 16) rpw >>= 1;
        int sblocks = (int)((rows + 8 * rpw - 1) / (8 * rpw));
        if (sblocks > ctx->sm_count * 32) sblocks = ctx->sm_count * 32;
        comp_sample_kernel<<<sblocks, 256, 0, st>>>(ctx->row_ptr, ctx->col_idx, row_begin, row_end, cutoff, ns, parent,
                                                    ctx->b_edges.as<int>(), above, stats, ctx->fuse_union ? 1 : 0,
                                                    ctx->comp_packed ? ctx->b_wrows.as<int2>() : nullptr,
                                                    ctx->comp_packed ? ctx->b_wcnt.as<int>() : nullptr, rpw);
        if (!ctx->fuse_union)
            comp_union_kernel<<<cnn_blocks(rows, 256), 256, 0, st>>>(ctx->b_edges.as<int>(), row_begin, rows, ns, parent);
        CNN_LAUNCHED(ctx, ctx->fuse_union ? 1 : 2);
    } else {
        comp_above_kernel<<<cnn_blocks(rows, 256), 256, 0, st>>>(ctx->row_ptr, ctx->col_idx, row_begin, row_end, above);
        CNN_LAUNCHED(ctx, 1);
    }
    if (phases & 1) {
        Span sp(ctx, T_COMP_FLATTEN);
        comp_compress_kernel<<<cnn_blocks(ctx->n, 256), 256, 0, st>>>(parent, ctx->n);
    }
    ll cap_edges = ctx->nnz + 64;  // work list: at most one entry per CSR entry of the row block
    CNN_TRY(cnn_ensure(ctx, ctx->b_worklist, sizeof(int2) * (size_t)cap_edges));
    unsigned long long *wl_count = stats + ST_WL_COUNT;  // + ST_COUNTER right behind it
    // work-list kernel variant: window bits per warp / resident CTAs per SM asked for (CNNB200_WL_VARIANT: 0 = 64 Ki bits,
    // 1 = 32 Ki bits and 8 CTAs per SM -- more warps in flight, rows spanning more than 32 Ki indices take the
    // binary-search path)
    static const int wl_variant = getenv("CNNB200_WL_VARIANT") ? atoi(getenv("CNNB200_WL_VARIANT")) : 0;
    static const int wl_ctas = wl_variant == 1 ? resident_ctas(comp_worklist_kernel<32768, 8>, WIN_WARPS * 32)
                                               : resident_ctas(comp_worklist_kernel<WIN_BITS, 6>, WIN_WARPS * 32);
    auto launch_worklist = [&](int chunk_) {
        if (wl_variant == 1)
            comp_worklist_kernel<32768, 8><<<ctx->sm_count * wl_ctas, WIN_WARPS * 32, 0, st>>>(
                ctx->row_ptr, ctx->col_idx, ctx->b_worklist.as<int2>(), wl_count, stats + ST_COUNTER, chunk_, cutoff, parent, stats);
        else
            comp_worklist_kernel<WIN_BITS, 6><<<ctx->sm_count * wl_ctas, WIN_WARPS * 32, 0, st>>>(
                ctx->row_ptr, ctx->col_idx, ctx->b_worklist.as<int2>(), wl_count, stats + ST_COUNTER, chunk_, cutoff, parent, stats);
    };
    if ((phases & 1) && ns > 0 && ctx->far_stride > 0 && !ctx->comp_packed) {
        {
            Span sp(ctx, T_COMP_SAMPLE);
            CNN_CUDA(cudaMemsetAsync(wl_count, 0, 2 * sizeof(unsigned long long), st));
            comp_far_emit_kernel<<<cnn_blocks(rows, 256), 256, 0, st>>>(ctx->row_ptr, ctx->col_idx, row_begin, row_end,
                                                                        ctx->far_stride, parent, above,
                                                                        ctx->b_worklist.as<int2>(), wl_count, ctx->n_far);
            launch_worklist(ctx->wl_chunk);
        }
        Span sp(ctx, T_COMP_FLATTEN);
        comp_compress_kernel<<<cnn_blocks(ctx->n, 256), 256, 0, st>>>(parent, ctx->n);
        CNN_LAUNCHED(ctx, 3);
    }
    if (phases & 2) {
        Span sp(ctx, T_COMP_EDGES);
        CNN_CUDA(cudaMemsetAsync(wl_count, 0, 2 * sizeof(unsigned long long), st));
        int tile_rows = 32;  // about 8k entries per warp tile
        const ll mean_deg = ctx->nnz / (ctx->n > 0 ? ctx->n : 1);
        while (tile_rows > 1 && mean_deg * tile_rows > 8192) tile_rows >>= 1;
        int sblocks = (int)((rows + 8 * tile_rows - 1) / (8 * tile_rows));
        if (sblocks > ctx->sm_count * 32) sblocks = ctx->sm_count * 32;
        comp_scan_kernel<<<sblocks, 256, 0, st>>>(ctx->row_ptr, ctx->col_idx, row_begin, row_end, above, parent,
                                                  ctx->b_worklist.as<int2>(), wl_count, stats, tile_rows);
        if (ctx->comp_packed) {
            CNN_TRY(cnn_worklist_packed(ctx, ctx->b_worklist.as<int2>(), wl_count, cutoff, parent));
        } else {
            launch_worklist(ctx->wl_chunk);
        }
        CNN_LAUNCHED(ctx, 2);
    }
    if (phases & 2) {
        Span sp(ctx, T_COMP_FLATTEN);
        comp_compress_kernel<<<cnn_blocks(ctx->n, 256), 256, 0, st>>>(parent, ctx->n);
    }
    CNN_LAUNCHED(ctx, 2);
    CNN_KERNEL_CHECK();
    return CNNB200_OK;
}

extern "C" int cnnb200_components_edges(cnnb200_ctx *ctx, int64_t row_begin, int64_t row_end) {
    return components_pass(ctx, row_begin, row_end, 3);
}

extern "C" int cnnb200_components_phase(cnnb200_ctx *ctx, int64_t row_begin, int64_t row_end, int phase) {
    if (phase != 1 && phase != 2) CNN_FAIL(CNNB200_E_INVALID, "components_phase: phase must be 1 (sampling rounds) or 2 (scan + work list)");
    return components_pass(ctx, row_begin, row_end, phase);
}

extern "C" int cnnb200_components_from_counts(cnnb200_ctx *ctx, const uint32_t *counts) {
    if (!ctx || !counts) CNN_FAIL(CNNB200_E_INVALID, "components_from_counts: NULL argument");
    if (!ctx->have_comp) CNN_FAIL(CNNB200_E_STATE, "components_from_counts: call components_begin first");
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const ll n = ctx->n;
    if (n == 0) return CNNB200_OK;
    const unsigned cutoff = ctx->similarity_cutoff > 0xfffffffell ? 0xfffffffeu : (unsigned)ctx->similarity_cutoff;
    unsigned long long *stats = (unsigned long long *)ctx->b_stats.p;
    int *parent = ctx->b_parent.as<int>();
    CNN_TRY(cnn_ensure(ctx, ctx->b_above, sizeof(int) * (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_edges, sizeof(int) * (size_t)n * 2));
    {
        Span sp(ctx, T_COMP_SAMPLE);
        counts_sample_kernel<<<cnn_blocks(n, 256), 256, 0, st>>>(ctx->row_ptr, ctx->col_idx, counts, n, cutoff, parent,
                                                                 ctx->b_edges.as<int>(), ctx->b_above.as<int>(), ctx->n_samples);
        comp_union_kernel<<<cnn_blocks(n, 256), 256, 0, st>>>(ctx->b_edges.as<int>(), 0, n, 2, parent);
    }
    {
        Span sp(ctx, T_COMP_FLATTEN);
        comp_compress_kernel<<<cnn_blocks(n, 256), 256, 0, st>>>(parent, n);
    }
    {
        Span sp(ctx, T_COMP_EDGES);
        int sblocks = (int)((n + 255) / 256);
        if (sblocks > ctx->sm_count * 32) sblocks = ctx->sm_count * 32;
        counts_scan_kernel<<<sblocks, 256, 0, st>>>(ctx->row_ptr, ctx->col_idx, counts, n, ctx->b_above.as<int>(), cutoff,
                                                    parent, stats);
    }
    {
        Span sp(ctx, T_COMP_FLATTEN);
        comp_compress_kernel<<<cnn_blocks(n, 256), 256, 0, st>>>(parent, n);
    }
    CNN_LAUNCHED(ctx, 5);
    CNN_KERNEL_CHECK();
    return CNNB200_OK;
}

extern "C" int cnnb200_components_parents(cnnb200_ctx *ctx, int32_t **parents) {
    if (!ctx || !parents) CNN_FAIL(CNNB200_E_INVALID, "components_parents: NULL argument");
    if (!ctx->have_comp) CNN_FAIL(CNNB200_E_STATE, "components_parents: call components_begin first");
    *parents = ctx->b_parent.as<int>();
    return CNNB200_OK;
}

extern "C" int cnnb200_components_merge(cnnb200_ctx *ctx, const int32_t *others, int64_t count) {
    if (!ctx || !others) CNN_FAIL(CNNB200_E_INVALID, "components_merge: NULL argument");
    if (!ctx->have_comp) CNN_FAIL(CNNB200_E_STATE, "components_merge: call components_begin first");
    if (count < 1) return CNNB200_OK;
    CNN_CUDA(cudaSetDevice(ctx->device));
    Span sp(ctx, T_COMP_FLATTEN);
    CNN_LAUNCHED(ctx, 3);
    comp_adopt_core_kernel<<<cnn_blocks(ctx->n, 256), 256, 0, ctx->stream>>>(others, ctx->n, count, ctx->b_parent.as<int>(),
                                                                           ctx->b_core.as<uint8_t>());
    comp_merge_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(others, ctx->n, count, ctx->b_parent.as<int>());
    comp_compress_kernel<<<cnn_blocks(ctx->n, 256), 256, 0, ctx->stream>>>(ctx->b_parent.as<int>(), ctx->n);
    CNN_KERNEL_CHECK();
    return CNNB200_OK;
}

extern "C" int cnnb200_cn_count(cnnb200_ctx *ctx, int64_t n_member_cutoff, uint32_t *counts, int loc,
                                int64_t *n_edges_out) {
    if (!ctx) CNN_FAIL(CNNB200_E_INVALID, "cn_count: NULL argument");
    if (!ctx->have_nb) CNN_FAIL(CNNB200_E_STATE, "cn_count: no neighbourhoods held by this context");
    return cnnb200_cn_count_rows(ctx, n_member_cutoff, 0, ctx->n, counts, loc, n_edges_out);
}

extern "C" int cnnb200_cn_count_rows(cnnb200_ctx *ctx, int64_t n_member_cutoff, int64_t row_begin, int64_t row_end,
                                     uint32_t *counts, int loc, int64_t *n_edges_out) {
    if (!ctx || !counts) CNN_FAIL(CNNB200_E_INVALID, "cn_count: NULL argument");
    if (!ctx->have_nb) CNN_FAIL(CNNB200_E_STATE, "cn_count: no neighbourhoods held by this context");
    if (row_begin < 0 || row_end > ctx->n || row_begin > row_end) CNN_FAIL(CNNB200_E_INVALID, "cn_count: bad row range");
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ll nnz = ctx->nnz;
    const ll rows = row_end - row_begin;
    CNN_TRY(cnn_ensure(ctx, ctx->b_stats, sizeof(unsigned long long) * ST_N));
    CNN_CUDA(cudaMemsetAsync(ctx->b_stats.p, 0, sizeof(unsigned long long) * ST_N, st));
    unsigned *d_counts = counts;
    if (loc == CNNB200_HOST) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_tmp0, sizeof(unsigned) * (size_t)nnz));
        d_counts = ctx->b_tmp0.as<unsigned>();
    }
    bool packed = false;
    CNN_TRY(cnn_pack_rows(ctx, &packed));
    ctx->stats[ST_PATH] = (ctx->stats[ST_PATH] & ~16ll) | (packed ? 16 : 0);
    if (packed) {
        CNN_CUDA(cudaMemsetAsync(ctx->b_stats.p, 0, sizeof(unsigned long long) * ST_N, st));
        CNN_TRY(cnn_count_packed(ctx, n_member_cutoff, row_begin, row_end, d_counts));
        if (loc == CNNB200_HOST)
            CNN_CUDA(cudaMemcpyAsync(counts, d_counts, sizeof(unsigned) * (size_t)nnz, cudaMemcpyDeviceToHost, st));
        if (loc == CNNB200_HOST || n_edges_out) {
            CNN_TRY(fetch_stats(ctx));
            if (n_edges_out) *n_edges_out = ctx->stats[ST_EDGES];
        }
        return CNNB200_OK;
    }
    // persistent grid: as many CTAs as fit (32 KB of window bitmaps each), rows handed out dynamically
    static const int cn_ctas = resident_ctas(cn_count_kernel, WIN_WARPS * 32);
    int blocks = ctx->sm_count * cn_ctas;
    ll want = (rows + WIN_WARPS * CNW_BATCH - 1) / (WIN_WARPS * CNW_BATCH);
    if (blocks > want) blocks = (int)want;
    if (blocks < 1) blocks = 1;
    unsigned long long *row_counter = (unsigned long long *)ctx->b_stats.p + ST_COUNTER;
    {
        Span sp(ctx, T_CN_COUNT);
        cn_count_kernel<<<blocks, WIN_WARPS * 32, 0, st>>>(ctx->row_ptr, ctx->col_idx, row_begin, row_end, n_member_cutoff, d_counts,
                                                          (unsigned long long *)ctx->b_stats.p, row_counter);
        CNN_LAUNCHED(ctx, 1);
    }
    CNN_KERNEL_CHECK();
    if (loc == CNNB200_HOST)
        CNN_CUDA(cudaMemcpyAsync(counts, d_counts, sizeof(unsigned) * (size_t)nnz, cudaMemcpyDeviceToHost, st));
    if (loc == CNNB200_HOST || n_edges_out) {
        CNN_TRY(fetch_stats(ctx));
        if (n_edges_out) *n_edges_out = ctx->stats[ST_EDGES];
    }
    return CNNB200_OK;
}

static int sharded_finish(cnnb200_ctx *ctx, int sort_by_size, int64_t current_start, int64_t member_cutoff, int64_t max_clusters,
                          int64_t *labels, int loc, int64_t *n_clusters_out);

// Second half of the row-sharded fit: `others` = the parent arrays of all ranks ([world, n] int32, all-gathered by the
// caller; this rank's own slice included or not -- a union with oneself is a no-op).  Rows another rank knows to be
// core become singleton trees here too, the forests are merged (hooking towards the smaller root makes the result
// the same on every rank), then the components are ranked and the labels written: labels == NULL skips the labels
// (ranks whose caller does not want them), loc = HOST copies them back.
extern "C" int cnnb200_fit_sharded_end(cnnb200_ctx *ctx, const int32_t *others, int64_t world, int sort_by_size,
                                       int64_t current_start, int64_t member_cutoff, int64_t max_clusters,
                                       int64_t *labels, int loc, int64_t *n_clusters_out) {
    if (!ctx || !others) CNN_FAIL(CNNB200_E_INVALID, "fit_sharded_end: NULL argument");
    if (!ctx->have_comp) CNN_FAIL(CNNB200_E_STATE, "fit_sharded_end: call fit_sharded_begin first");
    CNN_TRY(cnnb200_components_merge(ctx, others, world));
    return sharded_finish(ctx, sort_by_size, current_start, member_cutoff, max_clusters, labels, loc, n_clusters_out);
}

extern "C" int cnnb200_fit_sharded_ranges(cnnb200_ctx *ctx, int64_t world, int64_t *ranges) {
    if (!ctx || !ranges) CNN_FAIL(CNNB200_E_INVALID, "fit_sharded_ranges: NULL argument");
    if (!ctx->have_comp || ctx->shard_world != world) CNN_FAIL(CNNB200_E_STATE, "fit_sharded_ranges: call fit_sharded_begin first");
    for (int64_t r = 0; r < 2 * world; ++r) ranges[r] = ctx->shard_ranges[r];
    return CNNB200_OK;
}

extern "C" int cnnb200_fit_sharded_end_windows(cnnb200_ctx *ctx, const int32_t *others, int64_t world, const int64_t *starts,
                                               int64_t len, int sort_by_size, int64_t current_start, int64_t member_cutoff,
                                               int64_t max_clusters, int64_t *labels, int loc, int64_t *n_clusters_out) {
    if (!ctx || !others || !starts) CNN_FAIL(CNNB200_E_INVALID, "fit_sharded_end_windows: NULL argument");
    if (!ctx->have_comp) CNN_FAIL(CNNB200_E_STATE, "fit_sharded_end_windows: call fit_sharded_begin first");
    if (world < 1 || world > 1024 || len < 0 || len > ctx->n) CNN_FAIL(CNNB200_E_INVALID, "fit_sharded_end_windows: bad world / len");
    for (int64_t g = 0; g < world; ++g)
        if (starts[g] < 0 || starts[g] + len > ctx->n) CNN_FAIL(CNNB200_E_INVALID, "fit_sharded_end_windows: window %lld leaves [0, n)", (long long)g);
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    {
        Span sp(ctx, T_COMP_FLATTEN);
        CNN_TRY(cnn_ensure(ctx, ctx->b_bbox, sizeof(ll) * (size_t)(3 * world + 1)));
        ll *h;
        CNN_TRY(cnn_pinned(ctx, sizeof(ll) * (size_t)(3 * world + 1), (void **)&h));
        for (int64_t g = 0; g < world; ++g) h[g] = starts[g];
        CNN_CUDA(cudaMemcpyAsync(ctx->b_bbox.p, h, sizeof(ll) * (size_t)world, cudaMemcpyHostToDevice, st));
        const ll *d_starts = ctx->b_bbox.as<ll>();
        if (len > 0) {
            comp_adopt_core_win_kernel<<<cnn_blocks(ctx->n, 256), 256, 0, st>>>(others, d_starts, len, ctx->n, world,
                                                                               ctx->b_parent.as<int>(), ctx->b_core.as<uint8_t>());
            comp_merge_win_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(others, d_starts, len, world, ctx->b_parent.as<int>());
        }
        comp_compress_kernel<<<cnn_blocks(ctx->n, 256), 256, 0, st>>>(ctx->b_parent.as<int>(), ctx->n);
        CNN_LAUNCHED(ctx, 3);
        CNN_KERNEL_CHECK();
    }
    return sharded_finish(ctx, sort_by_size, current_start, member_cutoff, max_clusters, labels, loc, n_clusters_out);
}

static int sharded_finish(cnnb200_ctx *ctx, int sort_by_size, int64_t current_start, int64_t member_cutoff, int64_t max_clusters,
                          int64_t *labels, int loc, int64_t *n_clusters_out) {
    const ll n = ctx->n;
    cudaStream_t st = ctx->stream;
    if (!labels) {
        if (n_clusters_out) *n_clusters_out = -1;
        return CNNB200_OK;
    }
    if (!sort_by_size) {
        if (loc == CNNB200_HOST) {  // fresh labels: nothing to upload
            CNN_TRY(cnn_ensure(ctx, ctx->b_labels, sizeof(ll) * (size_t)n));
            CNN_CUDA(cudaMemsetAsync(ctx->b_labels.p, 0, sizeof(ll) * (size_t)n, st));
            CNN_TRY(cnnb200_components_labels(ctx, current_start, ctx->b_labels.as<ll>(), CNNB200_DEVICE, 1, n_clusters_out));
            CNN_CUDA(cudaMemcpyAsync(labels, ctx->b_labels.p, sizeof(ll) * (size_t)n, cudaMemcpyDeviceToHost, st));
            CNN_CUDA(cudaStreamSynchronize(st));
            return CNNB200_OK;
        }
        return cnnb200_components_labels(ctx, current_start, labels, loc, 1, n_clusters_out);
    }
    CNN_TRY(cnn_ensure(ctx, ctx->b_labels, sizeof(ll) * (size_t)n));
    ll *d_labels = loc == CNNB200_HOST ? ctx->b_labels.as<ll>() : labels;
    return cnn_labels_sorted(ctx, current_start, member_cutoff, max_clusters, labels, d_labels, loc, n_clusters_out);
}

// ---- the early-exit checks on arbitrary pairs (SimilarityChecker.check in bulk; also what the tests pin) ---------------
// out[k] bit 0: lane_check (the per-lane two-pointer merge of the sampling round, SimilarityCheckerExtScreensorted._check,
// _types.pyx:2654-2706); bit 1: the whole-warp form of the work list (row window of N(i), N(j) streamed through it, or
// warp_common<true> when the row has no window).  Both must equal |N(i) ∩ N(j)| >= cutoff (cutoff 0: always true).
__global__ void __launch_bounds__(WIN_WARPS * 32)
cn_check_pairs_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col_idx, const int2 *__restrict__ pairs,
                      ll n_pairs, int cutoff, uint8_t *__restrict__ out) {
    __shared__ unsigned win_all[WIN_WARPS][WIN_BITS / 32];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    RowWin w;
    w.bits = win_all[threadIdx.x >> 5];
    w.first = 0;
    w.span = 0;
    for (int t = lane; t < WIN_BITS / 32; t += 32) w.bits[t] = 0;
    __syncwarp();
    const ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5, nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll k0 = warp * 32; k0 < n_pairs; k0 += nwarps * 32) {
        const ll k = k0 + lane;
        const bool valid = k < n_pairs;
        const int2 pr = valid ? pairs[k] : make_int2(0, 0);
        const ll bi = row_ptr[pr.x], bj = row_ptr[pr.y];
        const int mi = (int)(row_ptr[pr.x + 1] - bi), mj = (int)(row_ptr[pr.y + 1] - bj);
        int res = valid ? lane_check(col_idx + bi, mi, col_idx + bj, mj, cutoff) : 0;
        for (int src = 0; src < 32; ++src) {
            if (k0 + src >= n_pairs) break;
            const int *ri = col_idx + __shfl_sync(full, bi, src), *rj = col_idx + __shfl_sync(full, bj, src);
            const int ni = __shfl_sync(full, mi, src), nj = __shfl_sync(full, mj, src);
            int ok;
            if (cutoff <= 0) ok = 1;
            else if (ni < cutoff || nj < cutoff) ok = 0;
            else {
                win_set(w, ri, ni, lane);
                if (w.span) {
                    int buf[WIN_PF];
                    row_prefetch(buf, rj, nj, lane);
                    ok = row_check(buf, rj, nj, lane, w, cutoff);
                } else {
                    ok = warp_common<true>(ri, ni, rj, nj, cutoff, lane);
                }
                win_clear(w, ri, ni, lane);
            }
            if (lane == src) res |= ok ? 2 : 0;
        }
        if (valid) out[k] = (uint8_t)res;
    }
}

extern "C" int cnnb200_cn_check_pairs(cnnb200_ctx *ctx, const int32_t *pairs, int64_t n_pairs, int64_t similarity_cutoff,
                                      uint8_t *out, int loc) {
    if (!ctx || (!pairs && n_pairs > 0) || (!out && n_pairs > 0)) CNN_FAIL(CNNB200_E_INVALID, "cn_check_pairs: NULL argument");
    if (!ctx->have_nb) CNN_FAIL(CNNB200_E_STATE, "cn_check_pairs: no neighbourhoods held by this context");
    if (n_pairs < 0 || similarity_cutoff < 0) CNN_FAIL(CNNB200_E_INVALID, "cn_check_pairs: negative argument");
    if (n_pairs == 0) return CNNB200_OK;
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int2 *d_pairs = (const int2 *)pairs;
    uint8_t *d_out = out;
    if (loc == CNNB200_HOST) {
        for (ll k = 0; k < 2 * n_pairs; ++k)
            if (pairs[k] < 0 || pairs[k] >= ctx->n) CNN_FAIL(CNNB200_E_INVALID, "cn_check_pairs: index outside [0, n)");
        CNN_TRY(cnn_ensure(ctx, ctx->b_edges, sizeof(int2) * (size_t)n_pairs));
        CNN_TRY(cnn_ensure(ctx, ctx->b_flags, (size_t)n_pairs + 64));
        CNN_CUDA(cudaMemcpyAsync(ctx->b_edges.p, pairs, sizeof(int2) * (size_t)n_pairs, cudaMemcpyHostToDevice, st));
        d_pairs = ctx->b_edges.as<int2>();
        d_out = ctx->b_flags.as<uint8_t>();
    }
    const int cutoff = similarity_cutoff > 0x7fffffffll ? 0x7fffffff : (int)similarity_cutoff;
    ll blocks = (n_pairs + WIN_WARPS * 32 - 1) / (WIN_WARPS * 32);
    if (blocks > (ll)ctx->sm_count * 8) blocks = (ll)ctx->sm_count * 8;
    cn_check_pairs_kernel<<<(int)blocks, WIN_WARPS * 32, 0, st>>>(ctx->row_ptr, ctx->col_idx, d_pairs, n_pairs, cutoff, d_out);
    CNN_LAUNCHED(ctx, 1);
    CNN_KERNEL_CHECK();
    if (loc == CNNB200_HOST) {
        CNN_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n_pairs, cudaMemcpyDeviceToHost, st));
        CNN_CUDA(cudaStreamSynchronize(st));
    }
    return CNNB200_OK;
}
