// grid.cu -- K1/K2: uniform-grid (cell list) radius-neighbour search for d <= 3 -> sorted CSR.
//
// Bulk replacement of NeighboursGetterExtBruteForce._get (reference _types.pyx:1692-1710).
// The acceptance test is the reference's own f64 expression (Metric._calc_distance,
// _types.pyx:2253-2267: sequential sum of squared differences, `<=` against the adjusted
// radius), evaluated with explicitly rounded intrinsics so that nvcc cannot contract to FMA.
// The grid only generates candidates, so the neighbour SETS equal the brute-force ones.
//
// Index space: points are reordered by cell (x fastest); the CSR is emitted in that
// cell-sorted space (rows ascending without any per-row sort because cells are visited in
// ascending id) together with perm[] (sorted position -> caller index).
#include <cub/cub.cuh>
#include <math.h>

#include <memory>

#include "common.cuh"

struct GridDesc {
    double lo[3];
    double h;
    int dims[3];
    int d;
};

// ---- bounding box -------------------------------------------------------------------------
__global__ void bbox_partial_kernel(const double *__restrict__ x, ll n, int d, double *partial) {
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x)
        for (int k = 0; k < d; ++k) {
            double v = x[i * d + k];
            mn[k] = fmin(mn[k], v);
            mx[k] = fmax(mx[k], v);
        }
    __shared__ double s[6][8];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o; o >>= 1) {
            mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
        if (lane == 0) { s[k][w] = mn[k]; s[3 + k][w] = mx[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        int k = threadIdx.x;
        double v = s[k][0];
        for (int j = 1; j < (int)(blockDim.x >> 5); ++j) v = k < 3 ? fmin(v, s[k][j]) : fmax(v, s[k][j]);
        partial[blockIdx.x * 6 + k] = v;
    }
}

__global__ void bbox_final_kernel(const double *partial, int nblocks, double *out) {
    int k = threadIdx.x;
    if (k >= 6) return;
    double v = partial[k];
    for (int b = 1; b < nblocks; ++b) v = k < 3 ? fmin(v, partial[b * 6 + k]) : fmax(v, partial[b * 6 + k]);
    out[k] = v;
}

// ---- cell assignment ------------------------------------------------------------------------
__device__ __forceinline__ int cell_coord(double v, double lo, double h, int dim) {
    int c = (int)floor(__ddiv_rn(__dsub_rn(v, lo), h));
    return min(max(c, 0), dim - 1);
}

__global__ void cell_id_kernel(const double *__restrict__ x, ll n, GridDesc g, int *cell_id, int *idx,
                               int *cell_count) {
    ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c[3] = {0, 0, 0};
    for (int k = 0; k < g.d; ++k) c[k] = cell_coord(x[i * g.d + k], g.lo[k], g.h, g.dims[k]);
    int cid = (c[2] * g.dims[1] + c[1]) * g.dims[0] + c[0];
    cell_id[i] = cid;
    idx[i] = (int)i;
    atomicAdd(&cell_count[cid], 1);
}

// gather coordinates into cell-sorted SoA: xs[k][p] = x[perm[p]][k]
__global__ void gather_kernel(const double *__restrict__ x, const int *__restrict__ perm, ll n, int d,
                              double *xs) {
    ll p = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (p >= n) return;
    ll src = perm[p];
    for (int k = 0; k < d; ++k) xs[k * n + p] = x[src * d + k];
}

// ---- K2: one warp per query row ---------------------------------------------------------------
template <int D, bool SQRT>
__device__ __forceinline__ bool within(double qx, double qy, double qz, const double *__restrict__ xs,
                                       ll n, int q, double rc) {
    double t = __dsub_rn(qx, xs[q]);
    double s = __dmul_rn(t, t);
    if (D > 1) {
        t = __dsub_rn(qy, xs[n + q]);
        s = __dadd_rn(s, __dmul_rn(t, t));
    }
    if (D > 2) {
        t = __dsub_rn(qz, xs[2 * n + q]);
        s = __dadd_rn(s, __dmul_rn(t, t));
    }
    if (SQRT) s = __dsqrt_rn(s);
    return s <= rc;
}

// One warp per tile of 32 consecutive cell-sorted points.  Within the tile, points of the same
// cell form a group of g contiguous lanes that shares one candidate list (the 3^D cells around
// it = 3^(D-1) contiguous index runs).  The 32 lanes are laid out as gp x k (gp = g rounded up to
// a power of two, k = 32 / gp): lane = sub * gp + jq tests query jq against candidate q0 + sub,
// so dense cells (g = 32, k = 1: one candidate broadcast to 32 queries) and isolated points
// (g = 1, k = 32: one query against 32 candidates) both keep the lanes busy.
//   count pass: per-lane hit counts, summed over `sub` by shuffles -> deg[]
//   fill pass : hits of one step are ranked with a ballot restricted to the lanes of the same
//               query (mask rep << jq); candidates ascend with the lane index, so rows come out
//               ascending in the cell-sorted index space.
__constant__ unsigned c_rep[6] = {0xffffffffu, 0x55555555u, 0x11111111u, 0x01010101u, 0x00010001u, 0x00000001u};

// MODE 0: count pass (deg[] only)            } the two-pass form, used when the single-pass
// MODE 1: fill pass (recomputes the distances) } scratch would not fit
// MODE 2: single pass -- hits are written straight into a gapped buffer in which row p owns
//         cap[p] slots (cap = size of its candidate list >= its degree) and deg[] comes out of
//         the same pass; compact_rows_kernel then squeezes the rows into the CSR.
template <int D, bool SQRT, int MODE>
__global__ void __launch_bounds__(256)
grid_rows_kernel(const double *__restrict__ xs, ll n, const int *__restrict__ cell_sorted,
                 const int *__restrict__ cell_start, GridDesc g, double rc, int *__restrict__ deg,
                 const ll *__restrict__ row_ptr, int *__restrict__ col_idx) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll p0 = warp * 32; p0 < n; p0 += nwarps * 32) {
        ll p = p0 + lane;
        int cid = p < n ? cell_sorted[p] : -1;
        unsigned todo = __ballot_sync(full, p < n);
        while (todo) {
            const int a = __ffs(todo) - 1;  // first lane of the next group
            const int c = __shfl_sync(full, cid, a);
            const unsigned grp = __ballot_sync(full, cid == c) & todo;
            todo &= ~grp;
            const int gsz = __popc(grp);
            const int lg = gsz <= 1 ? 0 : 32 - __clz(gsz - 1);
            const int gp = 1 << lg, k = 32 >> lg;
            const int jq = lane & (gp - 1), sub = lane >> lg;
            const bool qvalid = jq < gsz;
            const ll pq = p0 + a + (qvalid ? jq : 0);
            const double qx = xs[pq], qy = D > 1 ? xs[n + pq] : 0.0, qz = D > 2 ? xs[2 * n + pq] : 0.0;
            const unsigned qm = c_rep[lg] << jq;
            const ll base = MODE ? row_ptr[pq] : 0;
            const int cx = c % g.dims[0];
            const int t = c / g.dims[0];
            const int cy = t % g.dims[1];
            const int cz = t / g.dims[1];
            const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dims[0] - 1);
            const int y0 = D > 1 ? max(cy - 1, 0) : 0, y1 = D > 1 ? min(cy + 1, g.dims[1] - 1) : 0;
            const int z0 = D > 2 ? max(cz - 1, 0) : 0, z1 = D > 2 ? min(cz + 1, g.dims[2] - 1) : 0;
            int cnt = 0;
            for (int z = z0; z <= z1; ++z)
                for (int y = y0; y <= y1; ++y) {
                    const int cb = (z * g.dims[1] + y) * g.dims[0];
                    const int s = cell_start[cb + x0], e = cell_start[cb + x1 + 1];
#pragma unroll 4
                    for (int q0 = s; q0 < e; q0 += k) {
                        const int q = q0 + sub;
                        const bool inb = q < e;
                        // branch-free: out-of-range lanes re-test entry s (always valid here) and are masked
                        const bool hit = within<D, SQRT>(qx, qy, qz, xs, n, inb ? q : s, rc) && inb && qvalid;
                        if (MODE) {
                            const unsigned m = __ballot_sync(full, hit) & qm;
                            if (hit) col_idx[base + cnt + __popc(m & lt)] = q;
                            cnt += __popc(m);
                        } else {
                            cnt += hit ? 1 : 0;
                        }
                    }
                }
            if (MODE == 0)
                for (int off = gp; off < 32; off <<= 1) cnt += __shfl_xor_sync(full, cnt, off);
            if (MODE != 1 && sub == 0 && qvalid) deg[pq] = cnt;
        }
    }
}

// cap[p] = size of the candidate list of point p (all points of the 3^D cells around its cell)
template <int D>
__global__ void grid_cap_kernel(ll n, const int *__restrict__ cell_sorted, const int *__restrict__ cell_start,
                                GridDesc g, int *__restrict__ cap) {
    ll p = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (p > n) return;
    if (p == n) { cap[n] = 0; return; }
    const int c = cell_sorted[p];
    const int cx = c % g.dims[0];
    const int t = c / g.dims[0];
    const int cy = t % g.dims[1];
    const int cz = t / g.dims[1];
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dims[0] - 1);
    const int y0 = D > 1 ? max(cy - 1, 0) : 0, y1 = D > 1 ? min(cy + 1, g.dims[1] - 1) : 0;
    const int z0 = D > 2 ? max(cz - 1, 0) : 0, z1 = D > 2 ? min(cz + 1, g.dims[2] - 1) : 0;
    int tot = 0;
    for (int z = z0; z <= z1; ++z)
        for (int y = y0; y <= y1; ++y) {
            const int cb = (z * g.dims[1] + y) * g.dims[0];
            tot += cell_start[cb + x1 + 1] - cell_start[cb + x0];
        }
    cap[p] = tot;
}

// one warp per row: gapped[goff[p] .. +deg[p]) -> col_idx[row_ptr[p] .. +deg[p])
__global__ void __launch_bounds__(256)
compact_rows_kernel(const int *__restrict__ gapped, const ll *__restrict__ goff, const ll *__restrict__ row_ptr,
                    ll n, int *__restrict__ col_idx) {
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll p = warp; p < n; p += nwarps) {
        const ll src = goff[p], dst = row_ptr[p];
        const int m = (int)(row_ptr[p + 1] - dst);
        for (int k = lane; k < m; k += 32) col_idx[dst + k] = gapped[src + k];
    }
}

// ---- exclusive scan deg -> row_ptr (shared with the other producers) ------------------------
int cnn_scan_deg_to_row_ptr(cnnb200_ctx *ctx, const int32_t *deg, ll n, ll *row_ptr, ll *nnz_out,
                            ll *max_deg_out) {
    // deg has n+1 entries, the last one 0, so that row_ptr[n] = nnz comes out of the scan
    // int32 in, int64 out: the accumulator type follows the int64 initial value
    Span sp(ctx, T_SCAN);
    size_t need1 = 0, need2 = 0;
    CNN_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, need1, deg, row_ptr, ::cuda::std::plus<>{}, (ll)0, (int)(n + 1), ctx->stream));
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp2, 64));
    int *d_max = ctx->b_tmp2.as<int>() + 8;  // slots 0..7 belong to the callers' flags
    CNN_CUDA(cub::DeviceReduce::Max(nullptr, need2, deg, d_max, (int)(n + 1), ctx->stream));
    CNN_TRY(cnn_ensure(ctx, ctx->b_cub, need1 > need2 ? need1 : need2));
    size_t cap = ctx->b_cub.cap;
    CNN_CUDA(cub::DeviceScan::ExclusiveScan(ctx->b_cub.p, cap, deg, row_ptr, ::cuda::std::plus<>{}, (ll)0, (int)(n + 1), ctx->stream));
    cap = ctx->b_cub.cap;
    CNN_CUDA(cub::DeviceReduce::Max(ctx->b_cub.p, cap, deg, d_max, (int)(n + 1), ctx->stream));
    void *h;
    CNN_TRY(cnn_pinned(ctx, 64, &h));
    CNN_CUDA(cudaMemcpyAsync(h, row_ptr + n, sizeof(ll), cudaMemcpyDeviceToHost, ctx->stream));
    CNN_CUDA(cudaMemcpyAsync((char *)h + 8, d_max, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CNN_CUDA(cudaStreamSynchronize(ctx->stream));
    *nnz_out = *(ll *)h;
    *max_deg_out = *(int *)((char *)h + 8);
    return CNNB200_OK;
}

template <int D, bool SQRT>
static int launch_rows(cnnb200_ctx *ctx, ll n, const GridDesc &g, double rc, int mode, const ll *offsets, int *out) {
    const double *xs = ctx->b_xs.as<double>();
    const int *cs = ctx->b_cell_id_sorted.as<int>();
    const int *cstart = ctx->b_cell_start.as<int>();
    int blocks = (int)((n + 255) / 256);  // one warp per tile of 32 points, 8 warps per CTA
    int cap = ctx->sm_count * 64;
    if (blocks > cap) blocks = cap;
    int *deg = ctx->b_deg.as<int>();
    if (mode == 0)
        grid_rows_kernel<D, SQRT, 0><<<blocks, 256, 0, ctx->stream>>>(xs, n, cs, cstart, g, rc, deg, nullptr, nullptr);
    else if (mode == 1)
        grid_rows_kernel<D, SQRT, 1><<<blocks, 256, 0, ctx->stream>>>(xs, n, cs, cstart, g, rc, nullptr, offsets, out);
    else
        grid_rows_kernel<D, SQRT, 2><<<blocks, 256, 0, ctx->stream>>>(xs, n, cs, cstart, g, rc, deg, offsets, out);
    CNN_LAUNCHED(ctx, 1);
    CNN_KERNEL_CHECK();
    return CNNB200_OK;
}

static int launch_rows_any(cnnb200_ctx *ctx, ll n, const GridDesc &g, double rc, int metric, int mode,
                           const ll *offsets, int *out) {
    bool sq = metric == CNNB200_METRIC_EUCLIDEAN;
    switch (g.d * 2 + (sq ? 1 : 0)) {
        case 2: return launch_rows<1, false>(ctx, n, g, rc, mode, offsets, out);
        case 3: return launch_rows<1, true>(ctx, n, g, rc, mode, offsets, out);
        case 4: return launch_rows<2, false>(ctx, n, g, rc, mode, offsets, out);
        case 5: return launch_rows<2, true>(ctx, n, g, rc, mode, offsets, out);
        case 6: return launch_rows<3, false>(ctx, n, g, rc, mode, offsets, out);
        default: return launch_rows<3, true>(ctx, n, g, rc, mode, offsets, out);
    }
}

static int scan_int_to_ll(cnnb200_ctx *ctx, const int *in, ll *out, ll count) {
    size_t need = 0;
    CNN_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, need, in, out, ::cuda::std::plus<>{}, (ll)0, (int)count, ctx->stream));
    CNN_TRY(cnn_ensure(ctx, ctx->b_cub, need));
    size_t cap = ctx->b_cub.cap;
    CNN_CUDA(cub::DeviceScan::ExclusiveScan(ctx->b_cub.p, cap, in, out, ::cuda::std::plus<>{}, (ll)0, (int)count, ctx->stream));
    return CNNB200_OK;
}

int cnn_grid_search(cnnb200_ctx *ctx, const double *d_x, ll n, ll d, double radius_cutoff, int metric) {
    cudaStream_t st = ctx->stream;
    std::unique_ptr<Span> sp_build(new Span(ctx, T_GRID_BUILD));
    // --- bounding box --------------------------------------------------------------------
    int rb = ctx->sm_count * 2;
    if (rb > (int)((n + 255) / 256)) rb = (int)((n + 255) / 256);
    CNN_TRY(cnn_ensure(ctx, ctx->b_bbox, sizeof(double) * (size_t)(rb + 1) * 6));
    double *partial = ctx->b_bbox.as<double>();
    bbox_partial_kernel<<<rb, 256, 0, st>>>(d_x, n, (int)d, partial + 6);
    bbox_final_kernel<<<1, 32, 0, st>>>(partial + 6, rb, partial);
    CNN_LAUNCHED(ctx, 2);
    CNN_KERNEL_CHECK();
    double *hb;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&hb));
    CNN_CUDA(cudaMemcpyAsync(hb, partial, sizeof(double) * 6, cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < (int)d; ++k)
        if (!(isfinite(hb[k]) && isfinite(hb[3 + k])))
            CNN_FAIL(CNNB200_E_INVALID, "neighbour search: coordinates contain NaN or infinity");

    // --- grid geometry --------------------------------------------------------------------
    GridDesc g;
    g.d = (int)d;
    double r_true = metric == CNNB200_METRIC_EUCLIDEAN ? radius_cutoff : sqrt(radius_cutoff);
    // cell edge a little larger than the radius: rounding in the cell coordinate can then never
    // put two points that are within the radius more than one cell apart
    g.h = r_true * (1.0 + 1e-6);
    double span = 0;
    for (int k = 0; k < 3; ++k) {
        g.lo[k] = k < d ? hb[k] : 0.0;
        g.dims[k] = 1;
        if (k < d && hb[3 + k] - hb[k] > span) span = hb[3 + k] - hb[k];
    }
    if (!(g.h > 0)) g.h = span > 0 ? span : 1.0;  // radius 0: any cell size works
    const double max_cells = fmax(8.0 * (double)n, 1048576.0);
    for (;;) {
        double tot = 1;
        bool ok = true;
        for (int k = 0; k < (int)d; ++k) {
            double dk = floor((hb[3 + k] - hb[k]) / g.h) + 1;
            if (dk > 1048576.0) ok = false;
            tot *= dk;
            g.dims[k] = dk > 1048576.0 ? 1 : (int)dk;
        }
        if (ok && tot <= max_cells) break;
        g.h *= 1.26;  // sparse bounding box: coarser cells are still a valid candidate superset
    }
    ll ncell = (ll)g.dims[0] * g.dims[1] * g.dims[2];

    // --- K1: cell ids, sort by cell, cell starts, gather --------------------------------------
    CNN_TRY(cnn_ensure(ctx, ctx->b_cell_id, sizeof(int) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_cell_id_sorted, sizeof(int) * (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_idx, sizeof(int) * (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_perm, sizeof(int) * (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_cell_start, sizeof(int) * (size_t)(ncell + 2)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp0, sizeof(int) * (size_t)(ncell + 2)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_xs, sizeof(double) * (size_t)n * (size_t)d));
    CNN_TRY(cnn_ensure(ctx, ctx->b_deg, sizeof(int) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_row_ptr, sizeof(ll) * (size_t)(n + 1)));
    int *cell_count = ctx->b_tmp0.as<int>();
    CNN_CUDA(cudaMemsetAsync(cell_count, 0, sizeof(int) * (size_t)(ncell + 1), st));
    cell_id_kernel<<<cnn_blocks(n, 256), 256, 0, st>>>(d_x, n, g, ctx->b_cell_id.as<int>(),
                                                         ctx->b_idx.as<int>(), cell_count);
    CNN_LAUNCHED(ctx, 1);
    CNN_KERNEL_CHECK();
    int bits = 1;
    while ((1ll << bits) < ncell) ++bits;
    size_t need_sort = 0, need_scan = 0;
    CNN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need_sort, ctx->b_cell_id.as<int>(),
                                             ctx->b_cell_id_sorted.as<int>(), ctx->b_idx.as<int>(),
                                             ctx->b_perm.as<int>(), (int)n, 0, bits, st));
    CNN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need_scan, cell_count, ctx->b_cell_start.as<int>(),
                                           (int)(ncell + 1), st));
    CNN_TRY(cnn_ensure(ctx, ctx->b_cub, need_sort > need_scan ? need_sort : need_scan));
    size_t cap = ctx->b_cub.cap;
    CNN_CUDA(cub::DeviceRadixSort::SortPairs(ctx->b_cub.p, cap, ctx->b_cell_id.as<int>(),
                                             ctx->b_cell_id_sorted.as<int>(), ctx->b_idx.as<int>(),
                                             ctx->b_perm.as<int>(), (int)n, 0, bits, st));
    cap = ctx->b_cub.cap;
    CNN_CUDA(cub::DeviceScan::ExclusiveSum(ctx->b_cub.p, cap, cell_count, ctx->b_cell_start.as<int>(),
                                           (int)(ncell + 1), st));
    gather_kernel<<<cnn_blocks(n, 256), 256, 0, st>>>(d_x, ctx->b_perm.as<int>(), n, (int)d,
                                                        ctx->b_xs.as<double>());
    CNN_LAUNCHED(ctx, 1);
    CNN_KERNEL_CHECK();
    sp_build.reset();

    // --- K2: single pass into a gapped buffer + compaction, or count / scan / fill -----------------
    CNN_CUDA(cudaMemsetAsync(ctx->b_deg.as<int>() + n, 0, sizeof(int), st));
    ll nnz = 0, max_deg = 0;
    ll gapped_total = -1;
    if (ctx->single_pass_limit > 0) {
        // candidate-list sizes -> offsets of the gapped rows
        Span sp(ctx, T_SCAN);
        CNN_TRY(cnn_ensure(ctx, ctx->b_tmp1, sizeof(ll) * (size_t)(n + 1)));
        // the cell ids in input order are dead after the sort: reuse that buffer for cap[] (n+1 ints)
        int *cap = ctx->b_cell_id.as<int>();
        int nb1 = cnn_blocks(n + 1, 256);
        if (d == 1) grid_cap_kernel<1><<<nb1, 256, 0, st>>>(n, ctx->b_cell_id_sorted.as<int>(), ctx->b_cell_start.as<int>(), g, cap);
        else if (d == 2) grid_cap_kernel<2><<<nb1, 256, 0, st>>>(n, ctx->b_cell_id_sorted.as<int>(), ctx->b_cell_start.as<int>(), g, cap);
        else grid_cap_kernel<3><<<nb1, 256, 0, st>>>(n, ctx->b_cell_id_sorted.as<int>(), ctx->b_cell_start.as<int>(), g, cap);
        CNN_LAUNCHED(ctx, 1);
        CNN_KERNEL_CHECK();
        CNN_TRY(scan_int_to_ll(ctx, cap, ctx->b_tmp1.as<ll>(), n + 1));
        ll *h;
        CNN_TRY(cnn_pinned(ctx, 64, (void **)&h));
        CNN_CUDA(cudaMemcpyAsync(h, ctx->b_tmp1.as<ll>() + n, sizeof(ll), cudaMemcpyDeviceToHost, st));
        CNN_CUDA(cudaStreamSynchronize(st));
        gapped_total = h[0];
    }
    if (gapped_total >= 0 && gapped_total * 4 <= ctx->single_pass_limit) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_tmp0, sizeof(int) * (size_t)gapped_total));
        {
            Span sp(ctx, T_SEARCH_COUNT);
            CNN_TRY(launch_rows_any(ctx, n, g, radius_cutoff, metric, 2, ctx->b_tmp1.as<ll>(), ctx->b_tmp0.as<int>()));
        }
        CNN_TRY(cnn_scan_deg_to_row_ptr(ctx, ctx->b_deg.as<int>(), n, ctx->b_row_ptr.as<ll>(), &nnz, &max_deg));
        CNN_TRY(cnn_ensure(ctx, ctx->b_col_idx, sizeof(int) * (size_t)nnz));
        {
            Span sp(ctx, T_SEARCH_FILL);
            int blocks = (int)((n + 7) / 8);
            if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
            compact_rows_kernel<<<blocks, 256, 0, st>>>(ctx->b_tmp0.as<int>(), ctx->b_tmp1.as<ll>(), ctx->b_row_ptr.as<ll>(),
                                                       n, ctx->b_col_idx.as<int>());
            CNN_LAUNCHED(ctx, 1);
            CNN_KERNEL_CHECK();
        }
    } else {
        {
            Span sp(ctx, T_SEARCH_COUNT);
            CNN_TRY(launch_rows_any(ctx, n, g, radius_cutoff, metric, 0, nullptr, nullptr));
        }
        CNN_TRY(cnn_scan_deg_to_row_ptr(ctx, ctx->b_deg.as<int>(), n, ctx->b_row_ptr.as<ll>(), &nnz, &max_deg));
        CNN_TRY(cnn_ensure(ctx, ctx->b_col_idx, sizeof(int) * (size_t)nnz));
        {
            Span sp(ctx, T_SEARCH_FILL);
            CNN_TRY(launch_rows_any(ctx, n, g, radius_cutoff, metric, 1, ctx->b_row_ptr.as<ll>(), ctx->b_col_idx.as<int>()));
        }
    }

    ctx->n = n;
    ctx->nnz = nnz;
    ctx->max_deg = max_deg;
    ctx->row_ptr = ctx->b_row_ptr.as<ll>();
    ctx->col_idx = ctx->b_col_idx.as<int>();
    ctx->perm = ctx->b_perm.as<int>();
    return CNNB200_OK;
}
