// dense.cu -- K3 (exact variant): tiled all-pairs radius search in f64 for any dimension.
//
// Bulk replacement of NeighboursGetterExtBruteForce._get (reference _types.pyx:1692-1710) with
// MetricExtEuclideanReduced / MetricExtEuclidean (_types.pyx:2253-2267, :2173-2187): for every
// pair the sum of squared differences is accumulated sequentially over the dimension index in
// f64 without FMA contraction, exactly as the reference does, so the sets are bit-identical.
//
// A CTA owns 64 consecutive rows and walks over all 64-column tiles in ascending order; the
// hits of a tile are collected as one 64-bit mask per row in shared memory and popped in order,
// so each CSR row comes out ascending with no sort.  Index space = caller's (perm = identity).
#include "common.cuh"

#define DT 64   // tile edge (rows and columns)
#define DK 16   // dimension chunk staged per step

// METRIC: 0 = MetricExtEuclideanReduced, 1 = MetricExtEuclidean, 2 = MetricExtEuclideanPeriodicReduced
// (_types.pyx:2290-2318: |a - b| folded into [0, bound/2] per dimension when bound > 0, then squared and summed)
template <int METRIC, bool FILL>
__global__ void __launch_bounds__(256)
dense_rows_kernel(const double *__restrict__ x, ll n, int d, double rc, const double *__restrict__ bounds,
                  int *__restrict__ deg, const ll *__restrict__ row_ptr, int *__restrict__ col_idx) {
    constexpr bool SQRT = METRIC == 1;
    __shared__ double sa[DK][DT];
    __shared__ double sb[DK][DT];
    __shared__ unsigned long long smask[DT];
    __shared__ int scnt[DT];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const ll i0 = (ll)blockIdx.x * DT;
    if (tid < DT) scnt[tid] = 0;

    for (ll j0 = 0; j0 < n; j0 += DT) {
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
        if (tid < DT) smask[tid] = 0ull;

        for (int k0 = 0; k0 < d; k0 += DK) {
            __syncthreads();
            // stage [DT points] x [DK dims] of both operands, zero padded ((0-0)^2 adds exactly 0)
            for (int e = tid; e < DT * DK; e += 256) {
                int r = e / DK, kk = e % DK;
                ll gi = i0 + r, gj = j0 + r;
                int k = k0 + kk;
                sa[kk][r] = (gi < n && k < d) ? x[gi * d + k] : 0.0;
                sb[kk][r] = (gj < n && k < d) ? x[gj * d + k] : 0.0;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < DK; ++kk) {
                double av[4], bv[4];
                const double bound = (METRIC == 2 && k0 + kk < d) ? bounds[k0 + kk] : 0.0;
#pragma unroll
                for (int a = 0; a < 4; ++a) av[a] = sa[kk][ty * 4 + a];
#pragma unroll
                for (int b = 0; b < 4; ++b) bv[b] = sb[kk][tx + 16 * b];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        double t = __dsub_rn(av[a], bv[b]);
                        if (METRIC == 2) {
                            t = fabs(t);
                            if (bound > 0.0) {
                                t = fmod(t, bound);  // exact, as in C
                                if (t > bound * 0.5) t = __dsub_rn(bound, t);
                            }
                        }
                        acc[a][b] = __dadd_rn(acc[a][b], __dmul_rn(t, t));
                    }
            }
        }
        // threshold -> per-row bit masks
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            unsigned long long bits = 0ull;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                double s = acc[a][b];
                if (SQRT) s = __dsqrt_rn(s);
                ll gi = i0 + ty * 4 + a, gj = j0 + tx + 16 * b;
                if (gi < n && gj < n && s <= rc) bits |= 1ull << (tx + 16 * b);
            }
            if (bits) atomicOr(&smask[ty * 4 + a], bits);
        }
        __syncthreads();
        if (tid < DT) {
            unsigned long long m = smask[tid];
            int c = scnt[tid];
            if (FILL) {
                ll base = (i0 + tid < n) ? row_ptr[i0 + tid] : 0;
                while (m) {
                    int b = __ffsll((long long)m) - 1;
                    col_idx[base + c++] = (int)(j0 + b);
                    m &= m - 1;
                }
            } else {
                c += __popcll(m);
            }
            scnt[tid] = c;
        }
        // the next iteration's first __syncthreads orders smask/scnt reuse
    }
    __syncthreads();
    if (!FILL && tid < DT && i0 + tid < n) deg[i0 + tid] = scnt[tid];
}

int cnn_tensor_search(cnnb200_ctx *ctx, const double *d_x, ll n, ll d, double radius_cutoff, int metric);

int cnn_dense_search(cnnb200_ctx *ctx, const double *d_x, ll n, ll d, double radius_cutoff, int metric,
                     int method) {
    if (d > (1 << 20)) CNN_FAIL(CNNB200_E_RANGE, "dense search: dimension too large");
    if (method == 3) return cnn_tensor_search(ctx, d_x, n, d, radius_cutoff, metric);
    cudaStream_t st = ctx->stream;
    CNN_TRY(cnn_ensure(ctx, ctx->b_deg, sizeof(int) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_row_ptr, sizeof(ll) * (size_t)(n + 1)));
    CNN_CUDA(cudaMemsetAsync(ctx->b_deg.as<int>() + n, 0, sizeof(int), st));
    int blocks = (int)((n + DT - 1) / DT);
    const double *bounds = nullptr;
    if (metric == CNNB200_METRIC_EUCLIDEAN_PERIODIC_REDUCED) {
        if (ctx->n_bounds != d) CNN_FAIL(CNNB200_E_STATE, "periodic metric: set %lld bounds with cnnb200_ctx_set_periodic_bounds first", (long long)d);
        bounds = ctx->b_bounds.as<double>();
    }
    {
    Span sp(ctx, T_SEARCH_COUNT);
    if (metric == CNNB200_METRIC_EUCLIDEAN)
        dense_rows_kernel<1, false><<<blocks, 256, 0, st>>>(d_x, n, (int)d, radius_cutoff, bounds,
                                                            ctx->b_deg.as<int>(), nullptr, nullptr);
    else if (metric == CNNB200_METRIC_EUCLIDEAN_PERIODIC_REDUCED)
        dense_rows_kernel<2, false><<<blocks, 256, 0, st>>>(d_x, n, (int)d, radius_cutoff, bounds,
                                                            ctx->b_deg.as<int>(), nullptr, nullptr);
    else
        dense_rows_kernel<0, false><<<blocks, 256, 0, st>>>(d_x, n, (int)d, radius_cutoff, bounds,
                                                            ctx->b_deg.as<int>(), nullptr, nullptr);
    }
    CNN_LAUNCHED(ctx, 1);
    CNN_KERNEL_CHECK();
    ll nnz = 0, max_deg = 0;
    CNN_TRY(cnn_scan_deg_to_row_ptr(ctx, ctx->b_deg.as<int>(), n, ctx->b_row_ptr.as<ll>(), &nnz, &max_deg));
    CNN_TRY(cnn_ensure(ctx, ctx->b_col_idx, sizeof(int) * (size_t)nnz));
    {
    Span sp(ctx, T_SEARCH_FILL);
    if (metric == CNNB200_METRIC_EUCLIDEAN)
        dense_rows_kernel<1, true><<<blocks, 256, 0, st>>>(d_x, n, (int)d, radius_cutoff, bounds, nullptr,
                                                           ctx->b_row_ptr.as<ll>(), ctx->b_col_idx.as<int>());
    else if (metric == CNNB200_METRIC_EUCLIDEAN_PERIODIC_REDUCED)
        dense_rows_kernel<2, true><<<blocks, 256, 0, st>>>(d_x, n, (int)d, radius_cutoff, bounds, nullptr,
                                                           ctx->b_row_ptr.as<ll>(), ctx->b_col_idx.as<int>());
    else
        dense_rows_kernel<0, true><<<blocks, 256, 0, st>>>(d_x, n, (int)d, radius_cutoff, bounds, nullptr,
                                                           ctx->b_row_ptr.as<ll>(), ctx->b_col_idx.as<int>());
    }
    CNN_LAUNCHED(ctx, 1);
    CNN_KERNEL_CHECK();
    ctx->n = n;
    ctx->nnz = nnz;
    ctx->max_deg = max_deg;
    ctx->row_ptr = ctx->b_row_ptr.as<ll>();
    ctx->col_idx = ctx->b_col_idx.as<int>();
    ctx->perm = nullptr;
    return CNNB200_OK;
}
