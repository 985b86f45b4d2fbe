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
                 const ll *__restrict__ row_ptr, int *__restrict__ col_idx, ll p_begin, ll p_end) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll p0 = p_begin + warp * 32; p0 < p_end; p0 += nwarps * 32) {
        ll p = p0 + lane;
        int cid = p < p_end ? cell_sorted[p] : -1;
        unsigned todo = __ballot_sync(full, p < p_end);
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
                    ll p_begin, ll p_end, int *__restrict__ col_idx) {
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll p = p_begin + warp; p < p_end; p += nwarps) {
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
static int launch_rows(cnnb200_ctx *ctx, ll pb, ll pe, int mode, int *deg, const ll *offsets, int *out) {
    const GridDesc &g = ctx->grid;
    const ll n = ctx->grid_n;
    const double rc = ctx->grid_rc;
    const double *xs = ctx->b_xs.as<double>();
    const int *cs = ctx->b_cell_id_sorted.as<int>();
    const int *cstart = ctx->b_cell_start.as<int>();
    int blocks = (int)((pe - pb + 255) / 256);  // one warp per tile of 32 points, 8 warps per CTA
    int cap = ctx->sm_count * 64;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (mode == 0)
        grid_rows_kernel<D, SQRT, 0><<<blocks, 256, 0, ctx->stream>>>(xs, n, cs, cstart, g, rc, deg, nullptr, nullptr, pb, pe);
    else if (mode == 1)
        grid_rows_kernel<D, SQRT, 1><<<blocks, 256, 0, ctx->stream>>>(xs, n, cs, cstart, g, rc, nullptr, offsets, out, pb, pe);
    else
        grid_rows_kernel<D, SQRT, 2><<<blocks, 256, 0, ctx->stream>>>(xs, n, cs, cstart, g, rc, deg, offsets, out, pb, pe);
    CNN_LAUNCHED(ctx, 1);
    CNN_KERNEL_CHECK();
    return CNNB200_OK;
}

static int launch_rows_any(cnnb200_ctx *ctx, ll pb, ll pe, int mode, int *deg, const ll *offsets, int *out) {
    bool sq = ctx->grid_metric == CNNB200_METRIC_EUCLIDEAN;
    switch (ctx->grid.d * 2 + (sq ? 1 : 0)) {
        case 2: return launch_rows<1, false>(ctx, pb, pe, mode, deg, offsets, out);
        case 3: return launch_rows<1, true>(ctx, pb, pe, mode, deg, offsets, out);
        case 4: return launch_rows<2, false>(ctx, pb, pe, mode, deg, offsets, out);
        case 5: return launch_rows<2, true>(ctx, pb, pe, mode, deg, offsets, out);
        case 6: return launch_rows<3, false>(ctx, pb, pe, mode, deg, offsets, out);
        default: return launch_rows<3, true>(ctx, pb, pe, mode, deg, offsets, out);
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

// ---- K1: grid geometry, sort by cell, candidate-list prefix --------------------------------------
int cnn_grid_build(cnnb200_ctx *ctx, const double *d_x, ll n, ll d, double radius_cutoff, int metric) {
    cudaStream_t st = ctx->stream;
    ctx->have_grid = false;
    Span sp_build(ctx, T_GRID_BUILD);
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

    // --- cell ids, sort by cell, cell starts, gather ------------------------------------------
    CNN_TRY(cnn_ensure(ctx, ctx->b_cell_id, sizeof(int) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_cell_id_sorted, sizeof(int) * (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_idx, sizeof(int) * (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_perm, sizeof(int) * (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_cell_start, sizeof(int) * (size_t)(ncell + 2)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_cell_count, sizeof(int) * (size_t)(ncell + 2)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_xs, sizeof(double) * (size_t)n * (size_t)d));
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp1, sizeof(ll) * (size_t)(n + 1)));
    int *cell_count = ctx->b_cell_count.as<int>();
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
    // candidate-list sizes (the cell ids in input order are dead after the sort: reuse that buffer)
    // and their prefix = offsets of the gapped rows / the work measure used to balance row blocks
    int *capv = ctx->b_cell_id.as<int>();
    int nb1 = cnn_blocks(n + 1, 256);
    const int *cs = ctx->b_cell_id_sorted.as<int>(), *cst = ctx->b_cell_start.as<int>();
    if (d == 1) grid_cap_kernel<1><<<nb1, 256, 0, st>>>(n, cs, cst, g, capv);
    else if (d == 2) grid_cap_kernel<2><<<nb1, 256, 0, st>>>(n, cs, cst, g, capv);
    else grid_cap_kernel<3><<<nb1, 256, 0, st>>>(n, cs, cst, g, capv);
    CNN_LAUNCHED(ctx, 1);
    CNN_KERNEL_CHECK();
    CNN_TRY(scan_int_to_ll(ctx, capv, ctx->b_tmp1.as<ll>(), n + 1));
    ctx->grid = g;
    ctx->grid_n = n;
    ctx->grid_rc = radius_cutoff;
    ctx->grid_metric = metric;
    ctx->have_grid = true;
    ctx->perm = ctx->b_perm.as<int>();
    return CNNB200_OK;
}

// ---- K2 on a row block [pb, pe) of the cell-sorted points: degrees (and, single pass, the hits) -------
int cnn_grid_rows(cnnb200_ctx *ctx, ll pb, ll pe, int *deg) {
    cudaStream_t st = ctx->stream;
    ll *h;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&h));
    const ll *goff = ctx->b_tmp1.as<ll>();
    CNN_CUDA(cudaMemcpyAsync(h, goff + pb, sizeof(ll), cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaMemcpyAsync(h + 1, goff + pe, sizeof(ll), cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaStreamSynchronize(st));
    const ll g0 = h[0], gapped = h[1] - h[0];
    ctx->blk_begin = pb;
    ctx->blk_end = pe;
    ctx->blk_goff0 = g0;
    ctx->blk_single = ctx->single_pass_limit > 0 && gapped * 4 <= ctx->single_pass_limit;
    ctx->stats[ST_DIST] = gapped;
    Span sp(ctx, T_SEARCH_COUNT);
    if (ctx->blk_single) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_tmp0, sizeof(int) * (size_t)gapped));
        // row p writes at gapped[goff[p] - g0 ...]: shift the base pointer instead of the offsets
        CNN_TRY(launch_rows_any(ctx, pb, pe, 2, deg, goff, ctx->b_tmp0.as<int>() - g0));
    } else {
        CNN_TRY(launch_rows_any(ctx, pb, pe, 0, deg, nullptr, nullptr));
    }
    return CNNB200_OK;
}

// ---- K2 fill of the same row block into a CSR whose row_ptr covers all rows ----------------------------
int cnn_grid_fill(cnnb200_ctx *ctx, ll pb, ll pe, const ll *row_ptr, int *col_idx) {
    if (pb != ctx->blk_begin || pe != ctx->blk_end)
        CNN_FAIL(CNNB200_E_STATE, "grid fill: row block differs from the one the degrees were computed for");
    Span sp(ctx, T_SEARCH_FILL);
    if (ctx->blk_single) {
        int blocks = (int)((pe - pb + 7) / 8);
        if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
        if (blocks < 1) blocks = 1;
        compact_rows_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->b_tmp0.as<int>() - ctx->blk_goff0, ctx->b_tmp1.as<ll>(),
                                                            row_ptr, pb, pe, col_idx);
        CNN_LAUNCHED(ctx, 1);
        CNN_KERNEL_CHECK();
    } else {
        CNN_TRY(launch_rows_any(ctx, pb, pe, 1, nullptr, row_ptr, col_idx));
    }
    return CNNB200_OK;
}

int cnn_grid_search(cnnb200_ctx *ctx, const double *d_x, ll n, ll d, double radius_cutoff, int metric) {
    cudaStream_t st = ctx->stream;
    CNN_TRY(cnn_grid_build(ctx, d_x, n, d, radius_cutoff, metric));
    CNN_TRY(cnn_ensure(ctx, ctx->b_deg, sizeof(int) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_row_ptr, sizeof(ll) * (size_t)(n + 1)));
    CNN_CUDA(cudaMemsetAsync(ctx->b_deg.as<int>() + n, 0, sizeof(int), st));
    CNN_TRY(cnn_grid_rows(ctx, 0, n, ctx->b_deg.as<int>()));
    ll nnz = 0, max_deg = 0;
    CNN_TRY(cnn_scan_deg_to_row_ptr(ctx, ctx->b_deg.as<int>(), n, ctx->b_row_ptr.as<ll>(), &nnz, &max_deg));
    CNN_TRY(cnn_ensure(ctx, ctx->b_col_idx, sizeof(int) * (size_t)nnz));
    CNN_TRY(cnn_grid_fill(ctx, 0, n, ctx->b_row_ptr.as<ll>(), ctx->b_col_idx.as<int>()));
    ctx->n = n;
    ctx->nnz = nnz;
    ctx->max_deg = max_deg;
    ctx->row_ptr = ctx->b_row_ptr.as<ll>();
    ctx->col_idx = ctx->b_col_idx.as<int>();
    ctx->perm = ctx->b_perm.as<int>();
    return CNNB200_OK;
}

// ---- row-sharded construction (multi-GPU): C-ABI -----------------------------------------------------
__global__ void balance_kernel(const ll *__restrict__ goff, ll n, int parts, ll *bounds) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > parts) return;
    if (k == 0) { bounds[0] = 0; return; }
    if (k == parts) { bounds[parts] = n; return; }
    // first row whose candidate prefix reaches k/parts of the total, rounded to a multiple of 32 rows
    const ll total = goff[n];
    const ll target = (ll)((double)total * k / parts);
    ll lo = 0, hi = n;
    while (lo < hi) {
        ll mid = (lo + hi) >> 1;
        if (goff[mid] < target) lo = mid + 1; else hi = mid;
    }
    lo = (lo + 16) / 32 * 32;
    bounds[k] = lo > n ? n : lo;
}

extern "C" int cnnb200_grid_begin(cnnb200_ctx *ctx, const double *x, int loc, int64_t n, int64_t d,
                                  double radius_cutoff, int metric) {
    if (!ctx || !x) CNN_FAIL(CNNB200_E_INVALID, "grid_begin: NULL argument");
    if (n < 1 || d < 1 || d > 3) CNN_FAIL(CNNB200_E_INVALID, "grid_begin: need n >= 1 and 1 <= d <= 3");
    if (n >= (1ll << 31) - 1) CNN_FAIL(CNNB200_E_RANGE, "grid_begin: n must be < 2^31-1");
    if (metric != CNNB200_METRIC_EUCLIDEAN_REDUCED && metric != CNNB200_METRIC_EUCLIDEAN)
        CNN_FAIL(CNNB200_E_INVALID, "grid_begin: unknown metric %d", metric);
    if (!(radius_cutoff >= 0.0)) CNN_FAIL(CNNB200_E_INVALID, "grid_begin: radius_cutoff must be >= 0");
    CNN_CUDA(cudaSetDevice(ctx->device));
    ctx->have_nb = false;
    ctx->have_comp = false;
    const double *d_x = x;
    if (loc == CNNB200_HOST) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_x, sizeof(double) * (size_t)n * (size_t)d));
        Span sp(ctx, T_H2D);
        CNN_CUDA(cudaMemcpyAsync(ctx->b_x.p, x, sizeof(double) * (size_t)n * (size_t)d, cudaMemcpyHostToDevice, ctx->stream));
        d_x = ctx->b_x.as<double>();
    }
    return cnn_grid_build(ctx, d_x, n, d, radius_cutoff, metric);
}

extern "C" int cnnb200_grid_balance(cnnb200_ctx *ctx, int parts, int64_t *bounds) {
    if (!ctx || !bounds || parts < 1 || parts > 1024) CNN_FAIL(CNNB200_E_INVALID, "grid_balance: bad argument");
    if (!ctx->have_grid) CNN_FAIL(CNNB200_E_STATE, "grid_balance: call grid_begin first");
    CNN_CUDA(cudaSetDevice(ctx->device));
    CNN_TRY(cnn_ensure(ctx, ctx->b_bbox, sizeof(ll) * (size_t)(parts + 1)));
    balance_kernel<<<cnn_blocks(parts + 1, 128), 128, 0, ctx->stream>>>(ctx->b_tmp1.as<ll>(), ctx->grid_n, parts,
                                                                         ctx->b_bbox.as<ll>());
    CNN_KERNEL_CHECK();
    ll *h;
    CNN_TRY(cnn_pinned(ctx, sizeof(ll) * (size_t)(parts + 1), (void **)&h));
    CNN_CUDA(cudaMemcpyAsync(h, ctx->b_bbox.p, sizeof(ll) * (size_t)(parts + 1), cudaMemcpyDeviceToHost, ctx->stream));
    CNN_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k <= parts; ++k) bounds[k] = h[k];
    for (int k = 1; k <= parts; ++k)
        if (bounds[k] < bounds[k - 1]) bounds[k] = bounds[k - 1];
    return CNNB200_OK;
}

extern "C" int cnnb200_grid_degrees(cnnb200_ctx *ctx, int64_t row_begin, int64_t row_end, int32_t *deg) {
    if (!ctx || !deg) CNN_FAIL(CNNB200_E_INVALID, "grid_degrees: NULL argument");
    if (!ctx->have_grid) CNN_FAIL(CNNB200_E_STATE, "grid_degrees: call grid_begin first");
    if (row_begin < 0 || row_end > ctx->grid_n || row_begin > row_end)
        CNN_FAIL(CNNB200_E_INVALID, "grid_degrees: bad row range");
    CNN_CUDA(cudaSetDevice(ctx->device));
    return cnn_grid_rows(ctx, row_begin, row_end, deg);
}

extern "C" int cnnb200_grid_fill(cnnb200_ctx *ctx, int64_t row_begin, int64_t row_end, const int64_t *row_ptr,
                                 int32_t *col_idx) {
    if (!ctx || !row_ptr || !col_idx) CNN_FAIL(CNNB200_E_INVALID, "grid_fill: NULL argument");
    if (!ctx->have_grid) CNN_FAIL(CNNB200_E_STATE, "grid_fill: call grid_begin first");
    CNN_CUDA(cudaSetDevice(ctx->device));
    return cnn_grid_fill(ctx, row_begin, row_end, (const ll *)row_ptr, col_idx);
}

extern "C" int cnnb200_neighbours_adopt(cnnb200_ctx *ctx, const int64_t *row_ptr, const int32_t *col_idx, int64_t n,
                                        int64_t nnz, int64_t max_deg, int grid_perm) {
    if (!ctx || !row_ptr || (!col_idx && nnz > 0)) CNN_FAIL(CNNB200_E_INVALID, "neighbours_adopt: NULL argument");
    if (grid_perm && (!ctx->have_grid || ctx->grid_n != n))
        CNN_FAIL(CNNB200_E_STATE, "neighbours_adopt: no grid of that size (call grid_begin first)");
    ctx->n = n;
    ctx->nnz = nnz;
    ctx->max_deg = max_deg;
    ctx->row_ptr = (const ll *)row_ptr;
    ctx->col_idx = col_idx;
    ctx->perm = grid_perm ? ctx->b_perm.as<int>() : nullptr;
    ctx->have_nb = true;
    ctx->have_comp = false;
    return CNNB200_OK;
}
