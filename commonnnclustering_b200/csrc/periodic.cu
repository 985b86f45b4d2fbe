// periodic.cu -- cell-list radius search under the periodic reduced Euclidean metric (d <= 3).
//
// Reference: MetricExtEuclideanPeriodicReduced._calc_distance (_types.pyx:2290-2318): per dimension
// t = |a - b|; with a period L > 0: t = fmod(t, L), t = L - t when t > L/2; distance = sum of t*t, compared with r*r.
//
// The non-periodic cell list (grid.cu) is reused as a FILTER on an augmented point set: every point wrapped into
// [0, L), plus its periodic images that lie within r' of the box (r' = r plus a rounding margin; needs 2 r' <= L in
// every periodic dimension, otherwise the tiled all-pairs kernel serves the request).  The rows of the real points are
// then re-tested with the reference's own arithmetic on the ORIGINAL coordinates, images are mapped back to the points
// they stand for, and the rows are brought into ascending order -- the neighbour sets are exactly the reference's.
#include <cub/cub.cuh>
#include <math.h>

#include "common.cuh"

struct PBox {
    double L[3];
    double rp;  // r': filter radius (not squared)
    int d;
};

__device__ __forceinline__ double pg_wrap(double v, double L) {
    if (!(L > 0.0)) return v;
    double w = fmod(v, L);
    if (w < 0.0) w += L;
    return w;
}

// per dimension: bit 0 = an image at w + L is needed (the point sits within r' of the lower face),
//                bit 1 = an image at w - L (within r' of the upper face)
__device__ __forceinline__ int pg_faces(double w, double L, double rp) {
    if (!(L > 0.0)) return 0;
    return (w < rp ? 1 : 0) | (w >= L - rp ? 2 : 0);
}

__device__ __forceinline__ int pg_options(int faces) { return 1 + (faces & 1) + ((faces >> 1) & 1); }

// the o-th choice of shift for a dimension with the given faces: 0 first, then +L, then -L
__device__ __forceinline__ double pg_shift(int faces, int o, double L) {
    if (o == 0) return 0.0;
    if (o == 1) return (faces & 1) ? L : -L;
    return -L;
}

__global__ void pg_absmax_kernel(const double *__restrict__ x, ll count, unsigned long long *out) {
    double m = 0.0;
    for (ll k = blockIdx.x * (ll)blockDim.x + threadIdx.x; k < count; k += (ll)gridDim.x * blockDim.x) {
        const double v = fabs(x[k]);
        if (v > m) m = v;  // a NaN is caught by the grid build
    }
    unsigned long long bits = (unsigned long long)__double_as_longlong(m);  // order-preserving for v >= 0
    for (int o = 16; o; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, bits, o);
        if (t > bits) bits = t;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(out, bits);
}

__global__ void pg_image_count_kernel(const double *__restrict__ x, ll n, PBox box, int *__restrict__ cnt) {
    ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (i > n) return;
    int c = 1;
    if (i < n)
        for (int k = 0; k < box.d; ++k) c *= pg_options(pg_faces(pg_wrap(x[i * box.d + k], box.L[k]), box.L[k], box.rp));
    cnt[i] = i < n ? c - 1 : 0;
}

// aug[0..n) = the wrapped points themselves, aug[n + off[i] ...] = the images of point i; orig[] = who they stand for
__global__ void pg_image_fill_kernel(const double *__restrict__ x, ll n, PBox box, const ll *__restrict__ off,
                                     double *__restrict__ aug, int *__restrict__ orig) {
    ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double w[3];
    int f[3], no[3];
    for (int k = 0; k < 3; ++k) { w[k] = 0.0; f[k] = 0; no[k] = 1; }
    for (int k = 0; k < box.d; ++k) {
        w[k] = pg_wrap(x[i * box.d + k], box.L[k]);
        f[k] = pg_faces(w[k], box.L[k], box.rp);
        no[k] = pg_options(f[k]);
        aug[i * box.d + k] = w[k];
    }
    orig[i] = (int)i;
    ll slot = n + off[i];
    for (int o2 = 0; o2 < no[2]; ++o2)
        for (int o1 = 0; o1 < no[1]; ++o1)
            for (int o0 = 0; o0 < no[0]; ++o0) {
                if (o0 == 0 && o1 == 0 && o2 == 0) continue;
                const int o[3] = {o0, o1, o2};
                for (int k = 0; k < box.d; ++k) aug[slot * box.d + k] = w[k] + pg_shift(f[k], o[k], box.L[k]);
                orig[slot] = (int)i;
                ++slot;
            }
}

// internal (cell-sorted) position p of the augmented search: real point or image?
__global__ void pg_flag_kernel(const int *__restrict__ perm, ll m, ll n, int *__restrict__ flag, int *__restrict__ inv) {
    ll p = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (p > m) return;
    if (p == m) { flag[p] = 0; return; }
    const int a = perm[p];
    flag[p] = a < n ? 1 : 0;
    if (a < n) inv[a] = (int)p;
}

// For every position q of the augmented search: mf[q] = position of the point behind q in the final (real points only)
// order, bit 31 set when q is an image; xq[q] = ORIGINAL coordinates of that point (what the exact test reads -- rows
// refer to nearby positions, so these gathers stay in L1/L2).  new_perm: final position -> caller index.
#define PG_IMG 0x80000000u
__global__ void pg_map_kernel(const double *__restrict__ x, const int *__restrict__ perm, const int *__restrict__ orig,
                              const int *__restrict__ rank, const int *__restrict__ inv, ll m, ll n, int d,
                              unsigned *__restrict__ mf, double *__restrict__ xq, int *__restrict__ new_perm) {
    ll q = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (q >= m) return;
    const int a = perm[q];
    const int o = orig[a];
    mf[q] = (unsigned)rank[inv[o]] | (a < n ? 0u : PG_IMG);
    for (int k = 0; k < d; ++k) xq[q * d + k] = x[(ll)o * d + k];
    if (a < n) new_perm[rank[q]] = a;
}

// the reference's arithmetic, operation by operation (no FMA: the library is built with --fmad=false)
__device__ __forceinline__ bool pg_exact(const double *__restrict__ xi, const double *__restrict__ xj, const PBox &box,
                                         double rc) {
    double total = 0.0;
    for (int k = 0; k < box.d; ++k) {
        double t = fabs(__dsub_rn(xi[k], xj[k]));
        const double L = box.L[k];
        if (L > 0.0) {
            if (t >= L) t = fmod(t, L);  // fmod(t, L) == t below L: same value, without the division loop
            if (t > L * 0.5) t = __dsub_rn(L, t);
        }
        total = __dadd_rn(total, __dmul_rn(t, t));
    }
    return total <= rc;
}

// Warp per internal row of a real point: the exact test on every entry the filter found.  The count pass gives the
// degrees (and how many entries of a row are images), the fill passes write the surviving entries, mapped to their
// final positions.  Positions of real points map monotonically, so a row without images
// stays ascending and is written directly.  A row with images (points near a periodic face) goes through shared
// memory: its real entries R (ascending) from the front, its image entries I (any order) from the back, and every
// entry lands at its rank = (entries of R below it) + (entries of I below it).  Rows longer than PG_CAP raise
// *needs_sort and are ordered by the generic row sort afterwards.
#define PG_CAP 2048
#define PG_WARPS 4
// MODE 0: count; 1: fill the rows that need no staging (no images, or too long); 2: fill the staged rows
template <int MODE>
__global__ void __launch_bounds__(PG_WARPS * 32)
pg_rows_kernel(const double *__restrict__ xq, const ll *__restrict__ row_ptr, const int *__restrict__ col_idx,
               const int *__restrict__ flag, const int *__restrict__ rank, const unsigned *__restrict__ mf, ll m, PBox box,
               double rc, int *__restrict__ deg_new, int *__restrict__ img_new, const ll *__restrict__ new_row_ptr,
               int *__restrict__ new_col, int *needs_sort) {
    __shared__ int stage_all[MODE == 2 ? PG_WARPS : 1][MODE == 2 ? PG_CAP : 1];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int *stage = stage_all[MODE == 2 ? (threadIdx.x >> 5) : 0];
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll p = warp; p < m; p += nwarps) {
        if (!flag[p]) continue;
        const int row = rank[p];
        ll base = 0;
        if (MODE != 0) {
            base = new_row_ptr[row];
            const int deg = (int)(new_row_ptr[row + 1] - base);
            const int n_img_row = img_new[row];
            const bool staged = n_img_row > 0 && deg <= PG_CAP;
            if (staged != (MODE == 2)) continue;
            if (MODE == 1 && n_img_row > 0 && lane == 0) *needs_sort = 1;
        }
        double xi[3];
        for (int k = 0; k < box.d; ++k) xi[k] = xq[p * box.d + k];
        const ll b = row_ptr[p], e = row_ptr[p + 1];
        ll pos = base;
        int count = 0, n_img = 0, nr = 0;
        for (ll k0 = b; k0 < e; k0 += 32) {
            const ll k = k0 + lane;
            unsigned v = 0;
            bool ok = false;
            if (k < e) {
                const int q = col_idx[k];
                ok = pg_exact(xi, xq + (ll)q * box.d, box, rc);
                v = mf[q];
            }
            const bool img = ok && (v & PG_IMG);
            const unsigned hits = __ballot_sync(full, ok);
            const unsigned imgs = __ballot_sync(full, img);
            if (MODE == 0) {
                count += __popc(hits);
                n_img += __popc(imgs);
            } else if (MODE == 1) {
                if (ok) new_col[pos + __popc(hits & lt)] = (int)(v & ~PG_IMG);
                pos += __popc(hits);
            } else {
                const unsigned reals = hits & ~imgs;
                if (ok && !img) stage[nr + __popc(reals & lt)] = (int)v;
                if (img) stage[PG_CAP - 1 - (n_img + __popc(imgs & lt))] = (int)(v & ~PG_IMG);
                nr += __popc(reals);
                n_img += __popc(imgs);
            }
        }
        if (MODE == 0 && lane == 0) {
            deg_new[row] = count;
            img_new[row] = n_img;
        }
        if (MODE == 2) {
            __syncwarp();
            const int *I = stage + PG_CAP - n_img;
            for (int t = lane; t < nr; t += 32) {
                const int val = stage[t];
                int c = 0;
                for (int u = 0; u < n_img; ++u) c += I[u] < val ? 1 : 0;
                new_col[base + t + c] = val;
            }
            for (int u = lane; u < n_img; u += 32) {
                const int val = I[u];
                int l = 0, h = nr;
                while (l < h) {
                    const int mid = (l + h) >> 1;
                    if (stage[mid] < val) l = mid + 1; else h = mid;
                }
                int c = l;
                for (int w = 0; w < n_img; ++w) c += I[w] < val ? 1 : 0;
                new_col[base + c] = val;
            }
            __syncwarp();
        }
    }
}

static inline void swap_buf(Buf &a, Buf &b) {
    Buf t = a;
    a = b;
    b = t;
}

// *served = false: the request does not fit the image construction (a period shorter than 2 r'); nothing was changed
int cnn_grid_search_periodic(cnnb200_ctx *ctx, const double *d_x, ll n, ll d, double radius_cutoff, bool *served) {
    cudaStream_t st = ctx->stream;
    *served = false;
    if (d < 1 || d > 3 || ctx->n_bounds != d) return CNNB200_OK;
    PBox box;
    box.d = (int)d;
    double Lmax = 0.0;
    for (int k = 0; k < 3; ++k) {
        box.L[k] = k < d ? ctx->bounds_h[k] : 0.0;
        if (box.L[k] > 0.0) {
            if (box.L[k] > Lmax) Lmax = box.L[k];
        } else if (box.L[k] < 0.0 || box.L[k] != box.L[k]) {
            box.L[k] = 0.0;  // the reference treats bound <= 0 as "not periodic"
        }
    }
    // largest |coordinate|: the reference's |a - b| carries a rounding error of eps * |a - b| into fmod
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp2, 64));
    unsigned long long *d_absmax = (unsigned long long *)(ctx->b_tmp2.as<char>() + 48);
    CNN_CUDA(cudaMemsetAsync(d_absmax, 0, sizeof(unsigned long long), st));
    {
        Span sp(ctx, T_GRID_BUILD);
        pg_absmax_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(d_x, n * d, d_absmax);
        CNN_LAUNCHED(ctx, 1);
    }
    CNN_KERNEL_CHECK();
    unsigned long long *h;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&h));
    CNN_CUDA(cudaMemcpyAsync(h, d_absmax, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaStreamSynchronize(st));
    double absmax;
    memcpy(&absmax, h, sizeof(double));
    if (!(absmax < 1e300)) return CNNB200_OK;  // inf / overflow: let the exact all-pairs kernel decide
    const double r = sqrt(radius_cutoff);
    const double eps = 2.220446049250313e-16;
    box.rp = r * (1.0 + 1e-12) + 64.0 * eps * (2.0 * absmax + 4.0 * Lmax);
    for (int k = 0; k < d; ++k)
        if (box.L[k] > 0.0 && !(2.0 * box.rp < box.L[k])) return CNNB200_OK;  // strict: no point may meet a point and its image

    // ---- augmented set ------------------------------------------------------------------------------------------------
    CNN_TRY(cnn_ensure(ctx, ctx->b_deg, sizeof(int) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_row_ptr, sizeof(ll) * (size_t)(n + 1)));
    ll n_img = 0, max_img = 0;
    {
        Span sp(ctx, T_GRID_BUILD);
        pg_image_count_kernel<<<cnn_blocks(n + 1, 256), 256, 0, st>>>(d_x, n, box, ctx->b_deg.as<int>());
        CNN_LAUNCHED(ctx, 1);
    }
    CNN_KERNEL_CHECK();
    CNN_TRY(cnn_scan_deg_to_row_ptr(ctx, ctx->b_deg.as<int>(), n, ctx->b_row_ptr.as<ll>(), &n_img, &max_img));
    const ll m = n + n_img;
    if (m >= (1ll << 31) - 1) return CNNB200_OK;
    CNN_TRY(cnn_ensure(ctx, ctx->b_pg_x, sizeof(double) * (size_t)(m * d)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_pg_orig, sizeof(int) * (size_t)m));
    {
        Span sp(ctx, T_GRID_BUILD);
        pg_image_fill_kernel<<<cnn_blocks(n, 256), 256, 0, st>>>(d_x, n, box, ctx->b_row_ptr.as<ll>(),
                                                                  ctx->b_pg_x.as<double>(), ctx->b_pg_orig.as<int>());
        CNN_LAUNCHED(ctx, 1);
    }
    CNN_KERNEL_CHECK();

    // ---- filter: the non-periodic cell list on the augmented set ---------------------------------------------------------
    CNN_TRY(cnn_grid_search(ctx, ctx->b_pg_x.as<double>(), m, d, box.rp * box.rp, CNNB200_METRIC_EUCLIDEAN_REDUCED));
    const ll filter_nnz = ctx->nnz;

    // ---- exact test on the original coordinates, images mapped back, real rows only ----------------------------------------
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp0, sizeof(int) * (size_t)(2 * (m + 1))));
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp1, sizeof(int) * (size_t)(n + m)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_pg_xq, sizeof(double) * (size_t)(m * d)));
    int *flag = ctx->b_tmp0.as<int>(), *rank = flag + (m + 1);
    int *inv = ctx->b_tmp1.as<int>();
    unsigned *mf = (unsigned *)(inv + n);
    double *xq = ctx->b_pg_xq.as<double>();
    CNN_TRY(cnn_ensure(ctx, ctx->b_pg_perm, sizeof(int) * (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_pg_deg, sizeof(int) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_pg_row_ptr, sizeof(ll) * (size_t)(n + 1)));
    const ll *f_row_ptr = ctx->row_ptr;
    const int *f_col = ctx->col_idx, *f_perm = ctx->perm;
    CNN_TRY(cnn_ensure(ctx, ctx->b_pg_img, sizeof(int) * (size_t)(n + 1)));
    int *img_new = ctx->b_pg_img.as<int>();
    int *needs_sort = ctx->b_tmp2.as<int>() + 14;  // byte 56: free between the scans (slots 0..7 flags, 8 max, 10 rechecks)
    CNN_CUDA(cudaMemsetAsync(needs_sort, 0, sizeof(int), st));
    ll blocks = (m + PG_WARPS - 1) / PG_WARPS;
    if (blocks > (ll)ctx->sm_count * 32) blocks = (ll)ctx->sm_count * 32;
    {
        Span sp(ctx, T_SEARCH_FILL);
        pg_flag_kernel<<<cnn_blocks(m + 1, 256), 256, 0, st>>>(f_perm, m, n, flag, inv);
        size_t need = 0;
        CNN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, flag, rank, (int)(m + 1), st));
        CNN_TRY(cnn_ensure(ctx, ctx->b_cub, need));
        size_t cap = ctx->b_cub.cap;
        CNN_CUDA(cub::DeviceScan::ExclusiveSum(ctx->b_cub.p, cap, flag, rank, (int)(m + 1), st));
        pg_map_kernel<<<cnn_blocks(m, 256), 256, 0, st>>>(d_x, f_perm, ctx->b_pg_orig.as<int>(), rank, inv, m, n, (int)d, mf, xq,
                                                         ctx->b_pg_perm.as<int>());
        CNN_CUDA(cudaMemsetAsync(ctx->b_pg_deg.as<int>() + n, 0, sizeof(int), st));
        pg_rows_kernel<0><<<(int)blocks, PG_WARPS * 32, 0, st>>>(xq, f_row_ptr, f_col, flag, rank, mf, m, box, radius_cutoff,
                                                                 ctx->b_pg_deg.as<int>(), img_new, nullptr, nullptr, nullptr);
        CNN_LAUNCHED(ctx, 3);
    }
    CNN_KERNEL_CHECK();
    ll nnz = 0, max_deg = 0;
    CNN_TRY(cnn_scan_deg_to_row_ptr(ctx, ctx->b_pg_deg.as<int>(), n, ctx->b_pg_row_ptr.as<ll>(), &nnz, &max_deg));
    CNN_TRY(cnn_ensure(ctx, ctx->b_pg_col, sizeof(int) * (size_t)(nnz > 0 ? nnz : 1)));
    {
        Span sp(ctx, T_SEARCH_FILL);
        pg_rows_kernel<1><<<(int)blocks, PG_WARPS * 32, 0, st>>>(xq, f_row_ptr, f_col, flag, rank, mf, m, box, radius_cutoff,
                                                                 nullptr, img_new, ctx->b_pg_row_ptr.as<ll>(),
                                                                 ctx->b_pg_col.as<int>(), needs_sort);
        pg_rows_kernel<2><<<(int)blocks, PG_WARPS * 32, 0, st>>>(xq, f_row_ptr, f_col, flag, rank, mf, m, box, radius_cutoff,
                                                                 nullptr, img_new, ctx->b_pg_row_ptr.as<ll>(),
                                                                 ctx->b_pg_col.as<int>(), needs_sort);
        CNN_LAUNCHED(ctx, 2);
    }
    CNN_KERNEL_CHECK();
    int *h_flag;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&h_flag));
    CNN_CUDA(cudaMemcpyAsync(h_flag, needs_sort, sizeof(int), cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaStreamSynchronize(st));
    const bool long_rows_with_images = h_flag[0] != 0;
    swap_buf(ctx->b_row_ptr, ctx->b_pg_row_ptr);
    swap_buf(ctx->b_col_idx, ctx->b_pg_col);
    swap_buf(ctx->b_perm, ctx->b_pg_perm);
    ctx->n = n;
    ctx->nnz = nnz;
    ctx->max_deg = max_deg;
    ctx->row_ptr = ctx->b_row_ptr.as<ll>();
    ctx->col_idx = ctx->b_col_idx.as<int>();
    ctx->perm = ctx->b_perm.as<int>();
    ctx->have_grid = false;  // the grid describes the augmented set, not the caller's points
    ctx->stats[ST_DIST] += filter_nnz;
    if (long_rows_with_images) {  // rows beyond the shared-memory staging: generic row sort
        ctx->have_nb = true;
        int sorted = 0;
        CNN_TRY(cnnb200_neighbours_sort_rows(ctx, &sorted));
        ctx->have_nb = false;
    }
    *served = true;
    return CNNB200_OK;
}
