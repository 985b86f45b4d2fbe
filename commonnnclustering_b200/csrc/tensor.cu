// tensor.cu -- K3: radius-neighbour search for high-dimensional points on the 5th-generation
// tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA) with an exact f64
// recheck of every pair near the cutoff.
//
// Bulk replacement of NeighboursGetterExtBruteForce._get (reference _types.pyx:1692-1710) with
// MetricExtEuclideanReduced / MetricExtEuclidean (_types.pyx:2253-2267, :2173-2187).  The
// reference evaluates N^2 distances in f64, one virtual call each; here the N^2 part is a dense
// contraction d2(i,j) = |x_i|^2 + |x_j|^2 - 2 x_i.x_j and runs as an fp16 GEMM:
//
//   * the (mean-centred, power-of-two scaled) coordinates are split x = hi + lo into two fp16
//     parts; the three products lo.hi + hi.lo + hi.hi are ONE MMA chain over the concatenated K
//     axis A' = [lo | hi | hi], B' = [hi | lo | hi] (K = 3 D, fp32 accumulation in TMEM);
//   * the epilogue reads the accumulator tile with tcgen05.ld, forms d2~ and classifies:
//       d2~ <= R - eps  -> neighbour,   d2~ > R + eps -> not a neighbour,   otherwise borderline;
//     the result leaves the SM as hit BITMAPS (32 rows x 128 columns per warp and tile, only where
//     there is a hit), borderline bits in a second bitmap;
//   * borderline pairs are decided by tile_resolve_kernel with the reference's own f64 expression
//     (sequential sum, no FMA, sqrt for the non-reduced metric); eps is a bound on |d2~ - d2|
//     derived from the largest norm, so the neighbour SETS are exactly the reference's;
//   * the CSR follows from the bitmaps by position (tile_rows_kernel: popcount -> scan -> expand):
//     rows come out ascending with no pair list and no sort.
//
// Warp roles (320 threads): warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (one
// elected lane) + TMEM allocation, warps 2-9 = epilogue (two row tiles x four TMEM lane quarters).
// Pipelines: smem B stages full/empty (TMA <-> MMA), four TMEM accumulator stages full/empty
// (MMA <-> epilogue).
#include <cub/cub.cuh>
#include <cuda.h>
#include <string.h>
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"

#define TT 128            // tile edge (rows of A' = UMMA M, rows of B' = UMMA N)
#define KATOM 64          // 16-bit elements per 128-byte swizzle atom
#define ATOM_BYTES (TT * 128)
#define MAX_ATOMS 4

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 in, f32 accumulate (kind::f16)
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}
// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;             // leading byte offset (not used by swizzled K-major layouts)
    d |= (uint64_t)(1024 >> 4) << 32;   // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;             // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;             // SWIZZLE_128B
    return d;
}
// 32 lanes x 32 columns of 32-bit: thread l of the warp gets columns [c, c+32) of its TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t *r = reinterpret_cast<uint32_t *>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
        " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// one lane of a converged warp (the others skip the block, nobody diverges for good)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------------
// Column sums, minima and maxima with whole rows read coalesced (a warp per row, lanes over the dimensions).  The
// order of the atomic adds is not fixed; the mean only centres the data (any centre gives the same neighbour sets:
// the filter's band is derived from the centred norms and every borderline pair is decided in f64 on the original
// coordinates).  Minima / maxima travel as order-preserving 64-bit codes so that atomicMin / atomicMax apply.
#define CM_MAXD 128
__device__ __forceinline__ unsigned long long f64_code(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_decode(unsigned long long c) {
    return __longlong_as_double((long long)((c >> 63) ? (c & 0x7fffffffffffffffull) : ~c));
}

__global__ void __launch_bounds__(256)
col_sum_kernel(const double *__restrict__ x, ll n, int d, double *sum, unsigned long long *mm /*[2][CM_MAXD]*/) {
    __shared__ double sh[8][CM_MAXD];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double acc[CM_MAXD / 32] = {0.0, 0.0, 0.0, 0.0};
    double lo[CM_MAXD / 32], hi[CM_MAXD / 32];
#pragma unroll
    for (int u = 0; u < CM_MAXD / 32; ++u) { lo[u] = INFINITY; hi[u] = -INFINITY; }
    const ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5, nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll i = warp; i < n; i += nwarps)
#pragma unroll
        for (int u = 0; u < CM_MAXD / 32; ++u)
            if (lane + 32 * u < d) {
                const double v = x[i * d + lane + 32 * u];
                acc[u] += v;
                lo[u] = fmin(lo[u], v);
                hi[u] = fmax(hi[u], v);
            }
    for (int pass = 0; pass < 3; ++pass) {
        __syncthreads();
#pragma unroll
        for (int u = 0; u < CM_MAXD / 32; ++u) sh[w][lane + 32 * u] = pass == 0 ? acc[u] : pass == 1 ? lo[u] : hi[u];
        __syncthreads();
        if ((int)threadIdx.x < d) {
            double t = sh[0][threadIdx.x];
            for (int k = 1; k < 8; ++k) {
                const double o = sh[k][threadIdx.x];
                t = pass == 0 ? t + o : pass == 1 ? fmin(t, o) : fmax(t, o);
            }
            if (pass == 0) atomicAdd(sum + threadIdx.x, t);
            else if (pass == 1) atomicMin(mm + threadIdx.x, f64_code(t));
            else atomicMax(mm + CM_MAXD + threadIdx.x, f64_code(t));
        }
    }
}

// mean[k] = sum[k] / n and the power of two that brings the largest centred coordinate into (2^13, 2^14]: fp16 keeps
// 11 significant bits down to 2^-14, so after the scaling every coordinate within 2^-27 of the largest one still
// splits into two normal halves (smaller ones lose relative, not absolute, precision: DESIGN.md)
__global__ void col_mean_kernel(double *mean, const unsigned long long *mm, int d, ll n, int *scale_e) {
    __shared__ double dev[CM_MAXD];
    double v = 0.0;
    if ((int)threadIdx.x < d) {
        const double m = mean[threadIdx.x] / (double)n;
        mean[threadIdx.x] = m;
        v = fmax(f64_decode(mm[CM_MAXD + threadIdx.x]) - m, m - f64_decode(mm[threadIdx.x]));
    }
    dev[threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double mx = 0.0;
        for (int k = 0; k < d; ++k) mx = fmax(mx, dev[k]);
        int e = 0;
        if (mx > 0.0 && isfinite(mx)) {
            int ex;
            frexp(mx, &ex);  // mx = f * 2^ex, 0.5 <= f < 1
            e = 14 - ex;
        }
        *scale_e = e;
    }
}

// X = 2^e (x - mean) split into two fp16 parts, X = hi + lo + delta with |lo| <= 2^-11 |X|, |delta| <= 2^-22 |X|
// (+ 2^-25 absolute where a part is subnormal).  A'[i] = [lo | hi | hi | 0..], B'[i] = [hi | lo | hi | 0..] (row
// length kt): the MMA chain over the concatenated K axis accumulates lo.hi + hi.lo FIRST, while the accumulator is
// still 2^-10 of its final size, and hi.hi last -- the roundings of the fp32 accumulation scale with what is being
// added to.  norm[i] = |X_i|^2 (f64 sum, rounded once); padding rows (i >= n) are zero with an infinite norm so
// that they never come near the cutoff.
// One warp per row: lanes stride over the dimensions, so the f64 row comes in as one coalesced read and every fp16
// segment goes out as contiguous 2-byte stores.
template <bool PAIR>
__global__ void __launch_bounds__(256)
split_kernel(const double *__restrict__ x, const double *__restrict__ mean, const int *__restrict__ scale_e, ll n,
             ll n_pad, int d, int kt, __half *A, __half *B, float *norm, int *maxnorm_bits) {
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    const __half zero = __float2half_rn(0.f);
    const int e = *scale_e;
    for (ll i = warp; i < n_pad; i += nwarps) {
        __half *a = A + i * kt, *b = B + i * kt;
        if (i >= n) {
            for (int k = 2 * lane; k < kt; k += 64) {  // kt is a multiple of 64
                *reinterpret_cast<__half2 *>(a + k) = __halves2half2(zero, zero);
                *reinterpret_cast<__half2 *>(b + k) = __halves2half2(zero, zero);
            }
            if (lane == 0) norm[i] = INFINITY;
            continue;
        }
        double s = 0;
        if (PAIR) {  // d even: two adjacent dimensions per lane, 4-byte stores (a thread per dimension: 2-byte stores)
            for (int k = 2 * lane; k < d; k += 64) {
                const double X0 = scalbn(x[i * d + k] - mean[k], e), X1 = scalbn(x[i * d + k + 1] - mean[k + 1], e);
                s += X0 * X0;
                s += X1 * X1;
                const __half h0 = __double2half(X0), h1 = __double2half(X1);
                const __half l0 = __double2half(X0 - (double)__half2float(h0)), l1 = __double2half(X1 - (double)__half2float(h1));
                const __half2 hi = __halves2half2(h0, h1), lo = __halves2half2(l0, l1);
                *reinterpret_cast<__half2 *>(a + k) = lo;
                *reinterpret_cast<__half2 *>(a + d + k) = hi;
                *reinterpret_cast<__half2 *>(a + 2 * d + k) = hi;
                *reinterpret_cast<__half2 *>(b + k) = hi;
                *reinterpret_cast<__half2 *>(b + d + k) = lo;
                *reinterpret_cast<__half2 *>(b + 2 * d + k) = hi;
            }
            for (int k = 3 * d + 2 * lane; k < kt; k += 64) {
                *reinterpret_cast<__half2 *>(a + k) = __halves2half2(zero, zero);
                *reinterpret_cast<__half2 *>(b + k) = __halves2half2(zero, zero);
            }
        } else {
            for (int k = lane; k < d; k += 32) {
                const double X = scalbn(x[i * d + k] - mean[k], e);
                s += X * X;
                const __half hi = __double2half(X);
                const __half lo = __double2half(X - (double)__half2float(hi));
                a[k] = lo; a[d + k] = hi; a[2 * d + k] = hi;
                b[k] = hi; b[d + k] = lo; b[2 * d + k] = hi;
            }
            for (int k = 3 * d + lane; k < kt; k += 32) { a[k] = zero; b[k] = zero; }
        }
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            const float sf = (float)s;
            norm[i] = sf;
            atomicMax(maxnorm_bits, __float_as_int(sf));  // NaN sorts above +inf: caught by the set-up kernel
        }
    }
}

// thresholds of the filter in the scaled units of the operands: thr[0] = R_s - eps, thr[1] = R_s + eps (rounded
// away from R_s, clamped to finite values so that an infinite column norm never meets an infinite threshold)
__global__ void tensor_setup_kernel(const int *maxnorm_bits, const int *scale_e, double R, double coef, float *thr, int *bad) {
    const float mf = __int_as_float(*maxnorm_bits);
    if (!isfinite(mf)) { *bad = 1; thr[0] = -INFINITY; thr[1] = -INFINITY; return; }
    *bad = 0;
    const double M = (double)mf * (1.0 + ldexp(1.0, -20));
    const double Rs = scalbn(R, 2 * *scale_e);
    const double eps = M * coef + Rs * ldexp(1.0, -22) + 1e-300;
    float lo = (float)(Rs - eps), hi = (float)(Rs + eps);
    lo = nextafterf(lo, -INFINITY);
    hi = nextafterf(hi, INFINITY);
    const float big = 3.0e38f;
    thr[0] = fminf(fmaxf(lo, -big), big);
    thr[1] = fminf(fmaxf(hi, -big), big);
}

// ---------------------------------------------------------------------------------------------------
// reordering and tile pruning
// ---------------------------------------------------------------------------------------------------
// High-dimensional data that is worth clustering is clustered: after sorting the points by (nearest pivot, distance
// to it) the 128-row tiles are spatially compact, and a tile pair whose bounding balls are further apart than the
// radius cannot contain a neighbour pair (triangle inequality) -- it is skipped by all three warp roles.  The CSR
// is produced in the sorted index space together with perm[] (sorted position -> caller index), as in the grid path.
#define N_PIVOTS 64

// A warp takes FOUR points at a time: the rows come in with coalesced reads (lanes stride over the dimensions); each lane
// then owns two pivots (n_piv <= 64) and walks the dimensions, the points' coordinates broadcast by shuffles, the pivots in
// shared memory as floats [d][n_piv].  Only the inner products are accumulated (|x - p|^2 = |x|^2 - 2 x.p + |p|^2: 14
// instructions per dimension for 4 points x 2 pivots); one arg-min per point at the end.  fp32 is enough: whatever the
// assignment, the tile pruning works from the bounding balls of the tiles that result.
__global__ void __launch_bounds__(256)
pivot_assign_kernel(const double *__restrict__ x, ll n, int d, int n_piv, unsigned long long *keys, int *idx) {
    extern __shared__ float piv[];  // [d][n_piv] | [n_piv] squared norms
    float *pn = piv + (size_t)d * n_piv;
    const unsigned full = 0xffffffffu;
    for (int t = threadIdx.x; t < n_piv * d; t += blockDim.x) {
        const int p = t / d, k = t - p * d;
        piv[k * n_piv + p] = (float)x[(ll)((double)p * (double)n / (double)n_piv) * d + k];
    }
    __syncthreads();
    for (int p = threadIdx.x; p < n_piv; p += blockDim.x) {
        float sq = 0.f;
        for (int k = 0; k < d; ++k) sq = fmaf(piv[k * n_piv + p], piv[k * n_piv + p], sq);
        pn[p] = sq;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const bool has0 = lane < n_piv, has1 = lane + 32 < n_piv;
    const float pn0 = has0 ? pn[lane] : 0.f, pn1 = has1 ? pn[lane + 32] : 0.f;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll i0 = warp * 4; i0 < n; i0 += nwarps * 4) {
        float xi[4][3];  // d <= 85 < 96: at most 3 coordinates per lane and point
        float xx[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const ll i = i0 + p < n ? i0 + p : n - 1;
            float sq = 0.f;
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                xi[p][u] = lane + 32 * u < d ? (float)x[i * d + lane + 32 * u] : 0.f;
                sq = fmaf(xi[p][u], xi[p][u], sq);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(full, sq, o);
            xx[p] = sq;
        }
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int kn = d - 32 * u < 32 ? d - 32 * u : 32;
            for (int kk = 0; kk < kn; ++kk) {
                const float *row = piv + (32 * u + kk) * n_piv;
                const float r0 = has0 ? row[lane] : 0.f, r1 = has1 ? row[lane + 32] : 0.f;
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float xv = __shfl_sync(full, xi[p][u], kk);
                    s0[p] = fmaf(xv, r0, s0[p]);
                    s1[p] = fmaf(xv, r1, s1[p]);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float best = has0 ? fmaf(-2.f, s0[p], pn0) : INFINITY;
            int arg = lane;
            const float b1 = fmaf(-2.f, s1[p], pn1);
            if (has1 && b1 < best) { best = b1; arg = lane + 32; }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const float ob = __shfl_xor_sync(full, best, o);
                const int oa = __shfl_xor_sync(full, arg, o);
                if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
            }
            if (lane == 0 && i0 + p < n) {
                const float dist = fmaxf(best + xx[p], 0.f);  // >= 0: the bits are ordered
                keys[i0 + p] = ((unsigned long long)arg << 32) | (unsigned long long)__float_as_uint(dist);
                idx[i0 + p] = (int)(i0 + p);
            }
        }
    }
}

__global__ void gather_rows_kernel(const double *__restrict__ x, const int *__restrict__ perm, ll n, int d, double *out) {
    ll t = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (t >= n * d) return;
    const ll p = t / d;
    const int k = (int)(t - p * d);
    out[t] = x[(ll)perm[p] * d + k];
}

// One block per tile: bounding ball(s) of its live rows -- centre (mean) and radius (largest distance of a row to the
// centre, rounded up).  A tile that straddles two pivot cells holds points of two unrelated regions; one ball around
// both would reach every other tile (C3: 62 such tiles made 43 % of all tile visits and the longest CTA 7x the
// average), so the rows from the first change of pivot on get a ball of their own (radius2 < 0: no second ball).
__global__ void tile_ball_kernel(const double *__restrict__ x, const unsigned long long *__restrict__ keys, ll n, int d,
                                 double *centre, double *radius, double *centre2, double *radius2) {
    __shared__ double c[2][128];
    __shared__ double rmax[TT];
    __shared__ int split;
    const ll T = blockIdx.x;
    const ll r0 = T * TT;
    const int rows = (int)((n - r0) < TT ? (n - r0) : TT);
    if (threadIdx.x == 0) split = rows;
    __syncthreads();
    if ((int)threadIdx.x < rows && (unsigned)(keys[r0 + threadIdx.x] >> 32) != (unsigned)(keys[r0] >> 32)) atomicMin(&split, (int)threadIdx.x);
    __syncthreads();
    const int sp = split;
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        for (int r = 0; r < sp; ++r) s1 += x[(r0 + r) * d + k];
        for (int r = sp; r < rows; ++r) s2 += x[(r0 + r) * d + k];
        c[0][k] = s1 / sp;
        c[1][k] = sp < rows ? s2 / (rows - sp) : 0.0;
        centre[(ll)k * gridDim.x + T] = c[0][k];  // [d][n_tiles]: tile_skip_kernel reads 32 consecutive tiles per dimension
        centre2[(ll)k * gridDim.x + T] = c[1][k];
    }
    __syncthreads();
    const int grp = (int)threadIdx.x >= sp ? 1 : 0;
    double dist = 0.0;
    if ((int)threadIdx.x < rows) {
        double s = 0.0;
        for (int k = 0; k < d; ++k) {
            const double t = x[(r0 + threadIdx.x) * d + k] - c[grp][k];
            s += t * t;
        }
        dist = sqrt(s);
    }
    for (int pass = 0; pass < 2; ++pass) {
        __syncthreads();
        rmax[threadIdx.x] = ((int)threadIdx.x < rows && grp == pass) ? dist : 0.0;
        __syncthreads();
        for (int o = TT / 2; o; o >>= 1) {
            if ((int)threadIdx.x < o) rmax[threadIdx.x] = fmax(rmax[threadIdx.x], rmax[threadIdx.x + o]);
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const double rr = rmax[0] * (1.0 + 1e-12) + 1e-300;
            if (pass == 0) radius[T] = rr;
            else radius2[T] = sp < rows ? rr : -1.0;
        }
    }
}

__device__ __forceinline__ bool balls_apart(const double *__restrict__ ca, ll a, const double *__restrict__ cb, ll b, ll n_tiles,
                                            int d, double ra, double rb, double reach) {
    double s = 0.0;
    for (int k = 0; k < d; ++k) {
        const double t = ca[(ll)k * n_tiles + a] - cb[(ll)k * n_tiles + b];
        s += t * t;
    }
    // any i in ball a, j in ball b: |x_i - x_j| >= |c_a - c_b| - rho_a - rho_b
    return sqrt(s) * (1.0 - 1e-12) - ra - rb > reach;
}

// bit J of row I set <=> tile pair (I, J) can be skipped; one warp per 32 column tiles.  Only the words at or above the
// diagonal are ever read (TileIter starts at the word of I0), the others are left alone.  This pass looks at the FIRST
// ball of every tile only; tile_skip_split_kernel takes back what that got wrong for the few tiles with two balls
// (testing all four combinations here made every warp with one such column walk the dimensions twice: 0.17 ms
// against 0.055 ms on C3).
__global__ void tile_skip_kernel(const double *__restrict__ centre, const double *__restrict__ radius, ll n_tiles, int d,
                                 int words, double r_true, unsigned *skip) {
    const ll gw = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= n_tiles * words) return;
    const ll I = gw / words;
    const ll wJ = gw - I * words;
    if (wJ * 32 + 31 < (I & ~1ll)) return;  // below the diagonal of the CTA's first row tile (I0 = I & ~1)
    const ll J = wJ * 32 + lane;
    bool sk = false;
    if (J < n_tiles) sk = balls_apart(centre, I, centre, J, n_tiles, d, radius[I], radius[J], r_true * (1.0 + 1e-9));
    const unsigned m = __ballot_sync(0xffffffffu, sk);
    if (lane == 0) skip[gw] = m;
}

// One row of blocks per tile S; only the tiles with a second ball do anything: against every other tile X the pair may be skipped
// only if ALL combinations of their balls are apart -- where they are not, the bits (S, X) and (X, S) are cleared.
__global__ void tile_skip_split_kernel(const double *__restrict__ centre, const double *__restrict__ radius,
                                       const double *__restrict__ centre2, const double *__restrict__ radius2, ll n_tiles, int d,
                                       int words, double r_true, unsigned *skip) {
    const ll S = blockIdx.x;
    const double rS2 = radius2[S];
    if (rS2 < 0.0) return;
    const double reach = r_true * (1.0 + 1e-9);
    {   // blockIdx.y: which 128 tiles X (one per thread: a block per tile S walking all X took 0.32 ms on 62 blocks)
        const ll X = blockIdx.y * (ll)blockDim.x + threadIdx.x;
        if (X >= n_tiles) return;
        const double rX2 = radius2[X];
        bool apart = balls_apart(centre, S, centre, X, n_tiles, d, radius[S], radius[X], reach) &&
                     balls_apart(centre2, S, centre, X, n_tiles, d, rS2, radius[X], reach);
        if (apart && rX2 >= 0.0)
            apart = balls_apart(centre, S, centre2, X, n_tiles, d, radius[S], rX2, reach) &&
                    balls_apart(centre2, S, centre2, X, n_tiles, d, rS2, rX2, reach);
        if (!apart) {
            atomicAnd(skip + S * words + (X >> 5), ~(1u << (X & 31)));
            atomicAnd(skip + X * words + (S >> 5), ~(1u << (S & 31)));
        }
    }
}

// the column tiles a CTA has to visit: J >= I0 that are needed by at least one of its row tiles I0, I0 + 1
struct TileIter {
    const unsigned *m0, *m1;  // skip rows of the two row tiles (nullptr: nothing is skipped)
    ll n_tiles;
    int w, words;
    unsigned cur;
    __device__ void init(const unsigned *skip, ll I0, ll n_tiles_, int words_) {
        n_tiles = n_tiles_;
        words = words_;
        m0 = skip ? skip + I0 * words : nullptr;
        m1 = (skip && I0 + 1 < n_tiles) ? skip + (I0 + 1) * words : m0;
        w = (int)(I0 >> 5);
        cur = load(w) & (0xffffffffu << (I0 & 31));
    }
    __device__ unsigned load(int ww) const {
        unsigned need = m0 ? ~(m0[ww] & m1[ww]) : 0xffffffffu;
        const ll base = (ll)ww * 32;
        if (base + 32 > n_tiles) need &= (1u << (n_tiles - base)) - 1u;  // 1 .. 31 tiles in the last word
        return need;
    }
    __device__ ll next() {  // -1 when exhausted
        while (!cur) {
            if (++w >= words) return -1;
            cur = load(w);
        }
        const int b = __ffs(cur) - 1;
        cur &= cur - 1;
        return (ll)w * 32 + b;
    }
};

// ---------------------------------------------------------------------------------------------------
// the tile kernel
// ---------------------------------------------------------------------------------------------------
#define NSB 8   // barrier slots for the shared-memory stages of the B operand; a stage is ONE 64-wide K atom (16 KB) of a
                // column tile: the loads run two tiles and more ahead of the MMAs (whole-tile stages: only two fit)
#define NST 4   // TMEM accumulator stages of 128 columns (MMA <-> epilogue): the whole 512-column TMEM
#define EPI_WARPS 8   // epilogue warps: one per (row tile, TMEM lane quarter), all 128 columns each
#define NUM_THREADS ((2 + EPI_WARPS) * 32)
#define BAND_FLAG 0x80000000u
#define SLAB_POOL 16  // slab slots a warp reserves at a time

// Output of the tile pass: hit BITMAPS.  The unit is a *slab*: the 32 rows of one TMEM lane quarter (row quarter
// g = 4 I + q of row tile I) against the 128 columns of column tile J, one uint4 per row (bit c of word w = column
// 32 w + c), 512 bytes written by one warp with one coalesced store.  Only slabs that hold a hit exist; they are
// found through two directories, dir[g][J] and its transpose dirT[J][g] (slot + 1, bit 31 = the slab has
// borderline pairs, whose bits sit in the parallel array bslabs).  Bits are kept for i < j only.
// Everything downstream is positional: the row of i is {bits of column i in the slabs above it} (ascending row
// quarters: ascending j < i), i itself, {bits of row i} (ascending column tiles) -- no pair list, no sort.
struct TensorParams {
    const float *norm;    // [n_pad] |X|^2 in the scaled units of the operands
    const float *thr;     // device: {R_s - eps, R_s + eps}
    ll n, n_tiles;
    int atoms, nsb;
    const unsigned *skip; // [n_tiles][words] tile pairs that cannot hold a neighbour pair (nullptr: visit all)
    int words;
    uint4 *slabs, *bslabs;      // [slab_cap][32]
    unsigned *dir, *dirT;       // [4 n_tiles][n_tiles], [n_tiles][4 n_tiles]
    int2 *info;                 // [slab_cap] (row quarter + 1, column tile | bit 30: holds borderline pairs) of the slabs in use, else (0, 0)
    unsigned *slab_count;       // slots handed out (may exceed slab_cap: then the pass is repeated)
    unsigned slab_cap;
    int dbg;  // measurement only (CNNB200_TENSOR_DBG): 1 = the epilogue skips the classification, 2 = no MMA is issued
};

// 32 accumulator elements of one row against their column norms.  w = |X_j|^2 - 2 X_i.X_j is formed in place and its
// minimum tracked (1.5 instructions per element); only when some row of the warp comes within reach of the cutoff are
// the two sign masks collected (4 more): bit c of mh <=> w_c <= thr_hi (neighbour or borderline), of ml <=> w_c <=
// thr_lo (neighbour for certain).  Two independent shift chains per mask.
__device__ __forceinline__ void classify32(float (&v)[32], const float *nc, float thr_hi, float thr_lo, unsigned &mh,
                                           unsigned &ml) {
    float mn = INFINITY;
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
        const float4 n4 = reinterpret_cast<const float4 *>(nc)[c4];
        v[4 * c4 + 0] = fmaf(-2.f, v[4 * c4 + 0], n4.x);
        v[4 * c4 + 1] = fmaf(-2.f, v[4 * c4 + 1], n4.y);
        v[4 * c4 + 2] = fmaf(-2.f, v[4 * c4 + 2], n4.z);
        v[4 * c4 + 3] = fmaf(-2.f, v[4 * c4 + 3], n4.w);
        mn = fminf(mn, fminf(v[4 * c4 + 0], v[4 * c4 + 1]));
        mn = fminf(mn, fminf(v[4 * c4 + 2], v[4 * c4 + 3]));
    }
    mh = 0;
    ml = 0;
    if (__any_sync(0xffffffffu, mn <= thr_hi)) {
        unsigned ha = 0, hb = 0, la = 0, lb = 0;  // sign bit set <=> w > thr
#pragma unroll
        for (int k = 15; k >= 0; --k) {
            ha = __funnelshift_l(__float_as_uint(thr_hi - v[16 + k]), ha, 1);
            hb = __funnelshift_l(__float_as_uint(thr_hi - v[k]), hb, 1);
            la = __funnelshift_l(__float_as_uint(thr_lo - v[16 + k]), la, 1);
            lb = __funnelshift_l(__float_as_uint(thr_lo - v[k]), lb, 1);
        }
        mh = ~((ha << 16) | hb);
        ml = ~((la << 16) | lb);
    }
}

// SINGLE pass over the upper triangle of the tile grid.  CTA c owns the TWO adjacent row tiles 2c and 2c+1 (256 rows,
// both A tiles stay in shared memory) and walks the column tiles J >= 2c: every B tile that comes in by TMA feeds two
// MMA groups, which halves the L2 -> shared-memory traffic per flop.  CTAs are launched in order of decreasing work.
// Warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocation, warps 2-5 = epilogue of row tile 2c, warps 6-9 of row
// tile 2c+1 (one TMEM lane quarter each).  Nothing depends on the order in which tiles are visited.
template <int ATOMS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tensor_tiles_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const TensorParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: 2 A tiles | nsb B atom stages | barriers | tmem slot | per-warp column norms
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic keeps the address space
    constexpr int atoms = ATOMS;
    const int nsb = P.nsb;
    const uint32_t tile_bytes = (uint32_t)atoms * ATOM_BYTES;
    uint8_t *sA = smem;
    uint8_t *sB = smem + 2 * tile_bytes;
    uint64_t *bars = (uint64_t *)(smem + 2 * tile_bytes + nsb * ATOM_BYTES);
    uint64_t *a_full = bars, *full = bars + 2, *empty = full + NSB, *t_full = empty + NSB, *t_empty = t_full + NST;
    uint32_t *tmem_slot = (uint32_t *)(t_empty + NST);
    float *ncs = (float *)(bars + 32);                  // [EPI_WARPS][2][TT], 16-byte aligned

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const ll n_tiles = P.n_tiles;
    const ll I0 = 2 * (ll)blockIdx.x;  // first row tile; the second one (I0 + 1) may lie past the end (all rows dead)

    if (warp == 0 && lane == 0) {
        mbar_init(a_full, 1);
        for (int s = 0; s < nsb; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < NST; ++s) { mbar_init(t_full + s, 1); mbar_init(t_empty + s, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(NST * TT));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (whole warp in the loop, one elected lane issues) =====
        if (elect_one()) {
            mbar_expect_tx(a_full, 2 * tile_bytes);   // rows past the end of the operand are zero-filled
            for (int rt = 0; rt < 2; ++rt)
                for (int a = 0; a < atoms; ++a)
                    tma_load_2d(sA + rt * tile_bytes + a * ATOM_BYTES, &mapA, a * KATOM, (int)((I0 + rt) * TT), a_full);
        }
        __syncwarp();
        int s = 0;
        uint32_t ph = 1;  // parity of the "empty" phase a fresh barrier passes
        TileIter it;
        it.init(P.skip, I0, n_tiles, P.words);
        for (ll j = it.next(); j >= 0; j = it.next())
#pragma unroll
            for (int a = 0; a < ATOMS; ++a) {
                mbar_wait(empty + s, ph);
                if (elect_one()) {
                    mbar_expect_tx(full + s, ATOM_BYTES);
                    tma_load_2d(sB + s * ATOM_BYTES, &mapB, a * KATOM, (int)(j * TT), full + s);
                }
                __syncwarp();
                if (++s == nsb) { s = 0; ph ^= 1; }
            }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // The WHOLE warp walks the loop (waits included) and one elected lane issues: addresses, descriptors and the
        // loop state then live in uniform registers and the 12 tcgen05.mma of a tile go out back to back.  (With
        // "if (lane == 0)" around the loop the compiler wrapped every MMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY
        // sequence, ~23 instructions of one thread each: 160 cycles per MMA against 64 in the tensor pipe -- the tile pass
        // of C3 took 0.70 ms with an empty epilogue.)
        // instruction descriptor: D f32, A/B fp16 (format 0), both K-major, N = 128, M = 128
        const uint32_t idesc = (1u << 4) | ((uint32_t)(TT >> 3) << 17) | ((uint32_t)(TT >> 4) << 24);
        const uint64_t desc_a0 = umma_desc(smem_u32(sA)), desc_a1 = umma_desc(smem_u32(sA) + tile_bytes);
        mbar_wait(a_full, 0);
        ll t = 0;
        int s = 0;
        uint32_t ph = 0;
        TileIter it;
        it.init(P.skip, I0, n_tiles, P.words);
        for (ll j = it.next(); j >= 0; j = it.next(), ++t) {
            // the tile (I0 + 1, I0) below the diagonal is computed too (its bits are dropped by the epilogue):
            // every accumulator stage is then used on every second step and the phases stay in lock step
            const int ts0 = (int)((t & 1) << 1);
#pragma unroll
            for (int a = 0; a < ATOMS; ++a) {
                mbar_wait(full + s, ph);          // TMA has landed this atom of the B tile
                tc_fence_after();
                const uint64_t desc_b = umma_desc(smem_u32(sB + s * ATOM_BYTES));
#pragma unroll
                for (int rt = 0; rt < 2; ++rt) {
                    if (a == 0) {
                        mbar_wait(t_empty + ts0 + rt, (uint32_t)((t >> 1) & 1) ^ 1);   // epilogue has drained this accumulator stage
                        tc_fence_after();
                    }
                    const uint32_t d_tmem = tmem_base + (uint32_t)(ts0 + rt) * TT;
                    const uint64_t desc_a = (rt ? desc_a1 : desc_a0) + (uint64_t)(a * (ATOM_BYTES >> 4));  // 16-byte units
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4 && !(P.dbg & 2); ++k)   // UMMA K = 16 halves = 32 bytes inside the swizzle atom
                            tc_mma(d_tmem, desc_a + (uint64_t)(k * 2), desc_b + (uint64_t)(k * 2), idesc, a > 0 || k > 0);
                        if (a == ATOMS - 1) tc_commit(t_full + ts0 + rt);  // accumulator ready for the epilogue
                    }
                    __syncwarp();
                }
                if (elect_one()) tc_commit(empty + s);        // atom stage free once these MMAs have read it
                __syncwarp();
                if (++s == nsb) { s = 0; ph ^= 1; }
            }
        }
        if (elect_one() && t) atomicAdd(P.slab_count + 1, (unsigned)(2 * t));  // 128 x 128 tiles this CTA put through the MMA
    } else {
        // ===== epilogue: TMEM -> registers -> classify -> hit bitmaps =====
        const unsigned fullm = 0xffffffffu;
        const int ew = warp - 2;             // 0 .. EPI_WARPS-1
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int rt = (ew >> 2) & 1;        // which of the CTA's two row tiles
        const int row = q * 32 + lane;       // accumulator row = TMEM lane
        float *ncw = ncs + ew * 2 * TT;
        const ll It = I0 + rt;
        const ll gi = It * TT + row;
        const ll g = It * 4 + q;             // row quarter
        const bool live = gi < P.n;
        const float nr = live ? P.norm[gi] : 0.f;
        // d2~ = nr + nc - 2 acc <= thr   <=>   w = nc - 2 acc <= thr - nr ; dead rows never pass
        const float thr_hi = live ? P.thr[1] - nr : -INFINITY, thr_lo = live ? P.thr[0] - nr : -INFINITY;
        unsigned pool_next = 0, pool_end = 0;
        TileIter it;
        it.init(P.skip, I0, n_tiles, P.words);
        ll j = it.next();
        float4 nn = make_float4(0.f, 0.f, 0.f, 0.f);  // column norms of the next tile: 4 per lane
        if (j >= 0) nn = reinterpret_cast<const float4 *>(P.norm + j * TT)[lane];
        for (ll t = 0; j >= 0; ++t) {
            const ll j_next = it.next();
            const int ts = (int)((t & 1) << 1) | rt;
            float *nc = ncw + (int)(t & 1) * TT;
            reinterpret_cast<float4 *>(nc)[lane] = nn;
            if (j_next >= 0) nn = reinterpret_cast<const float4 *>(P.norm + j_next * TT)[lane];
            __syncwarp();
            mbar_wait(t_full + ts, (uint32_t)((t >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ts * TT);
            unsigned hi[4], lo[4];
#pragma unroll
            for (int hp = 0; hp < 2; ++hp) {  // 64 columns at a time: two tcgen05.ld in flight
                float va[32], vb[32];
                tmem_ld32(taddr + hp * 64, va);
                tmem_ld32(taddr + hp * 64 + 32, vb);
                tmem_ld_wait();
                if (hp == 1) {  // this warp's part of the accumulator stage is in registers: hand it back
                    tc_fence_before();
                    mbar_arrive(t_empty + ts);
                }
                if (P.dbg & 1) { hi[2 * hp] = hi[2 * hp + 1] = lo[2 * hp] = lo[2 * hp + 1] = __float_as_uint(va[0] + vb[31]) == 0x12345u; continue; }
                classify32(va, nc + hp * 64, thr_hi, thr_lo, hi[2 * hp], lo[2 * hp]);
                classify32(vb, nc + hp * 64 + 32, thr_hi, thr_lo, hi[2 * hp + 1], lo[2 * hp + 1]);
            }
            if (j <= It) {  // diagonal tile: only columns above the row are kept; the tile below the diagonal: nothing
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    unsigned keep = 0;
                    if (j == It) keep = w < q ? 0u : (w > q ? 0xffffffffu : (lane == 31 ? 0u : (0xfffffffeu << lane)));
                    hi[w] &= keep;
                    lo[w] &= keep;
                }
            }
            const unsigned any = hi[0] | hi[1] | hi[2] | hi[3];
            if (__any_sync(fullm, any != 0)) {
                if (pool_next == pool_end) {
                    unsigned base = 0;
                    if (lane == 0) base = atomicAdd(P.slab_count, (unsigned)SLAB_POOL);
                    pool_next = __shfl_sync(fullm, base, 0);
                    pool_end = pool_next + SLAB_POOL;
                }
                const unsigned slot = pool_next++;
                if (slot < P.slab_cap) {
                    const uint4 bd = make_uint4(hi[0] & ~lo[0], hi[1] & ~lo[1], hi[2] & ~lo[2], hi[3] & ~lo[3]);
                    const bool band = __any_sync(fullm, (bd.x | bd.y | bd.z | bd.w) != 0);
                    P.slabs[(size_t)slot * 32 + lane] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    if (band) P.bslabs[(size_t)slot * 32 + lane] = bd;
                    if (lane == 0) {
                        const unsigned v = (slot + 1) | (band ? BAND_FLAG : 0u);
                        P.dir[g * n_tiles + j] = v;
                        P.dirT[j * (4 * n_tiles) + g] = v;
                        P.info[slot] = make_int2((int)g + 1, (int)j | (band ? 0x40000000 : 0));
                    }
                }
            }
            j = j_next;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NST * TT));
    }
}

// lane l holds row l of a 32 x 32 bit block (bit c = column c); afterwards lane c holds column c (bit l = row l)
__device__ __forceinline__ unsigned transpose32(unsigned x, int lane) {
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int k = 16; k; k >>= 1) {
        // m: bits whose index has bit k clear
        const unsigned m = k == 16 ? 0x0000ffffu : k == 8 ? 0x00ff00ffu : k == 4 ? 0x0f0f0f0fu : k == 2 ? 0x33333333u : 0x55555555u;
        const unsigned y = __shfl_xor_sync(full, x, k);
        x = (lane & k) ? ((x & ~m) | ((y & ~m) >> k)) : ((x & m) | ((y & m) << k));
    }
    return x;
}

// ---------------------------------------------------------------------------------------------------
// borderline pairs, decided with the reference's own f64 expression (_types.pyx:2253-2267)
// ---------------------------------------------------------------------------------------------------
// A warp walks the slabs that hold borderline pairs (slots are handed out strided, so the dense regions of the tile grid
// spread over all warps) and collects their pairs in a list in shared memory; whenever the list is nearly full -- and at
// the end -- the pairs are evaluated 32 at a time, one per LANE, in the reference's order (left to right, no FMA).  The
// coordinates pass through shared memory 8 dimensions at a time: the 64 rows of the 32 pairs come in as 64-byte
// pieces (4 lanes x 16 bytes: whole sectors, every byte used once) and each lane reads its own pair back as
// [dimension][row] pairs of adjacent doubles.  A rejected pair loses its bit in the hit slab (atomicAnd: the pairs of
// one row may sit in different lanes).
// Measured on C3 (10^6 borderline pairs): four pairs at a time with 8 lanes each and the running sum handed from lane
// to lane 0.57 ms (every DADD of the chain issued for the warp with four lanes active); one pair per lane reading
// its rows directly 1.26 ms (8 of every 32 bytes fetched used per request, rows evicted from L1 between uses).
#define RL_CAP 256
#define RL_WARPS 4
#define RL_PITCH 66   // doubles per dimension row of the stage: 64 rows + 2 (bank spread of the 4 column groups on store)
template <bool SQRT>
__global__ void __launch_bounds__(RL_WARPS * 32)
tile_resolve_kernel(const double *__restrict__ x, int d, const int2 *__restrict__ info, const unsigned *__restrict__ slab_count,
                    unsigned slab_cap, unsigned *slab_words, const uint4 *__restrict__ bslabs, double rc,
                    unsigned long long *n_band, unsigned long long *n_rejected) {
    __shared__ int4 s_list[RL_WARPS][RL_CAP];
    __shared__ __align__(16) double s_stage[RL_WARPS][8][RL_PITCH];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int4 *list = s_list[threadIdx.x >> 5];
    double (*stage)[RL_PITCH] = s_stage[threadIdx.x >> 5];
    const unsigned lt = (1u << lane) - 1u;
    const unsigned nwarps = (unsigned)(((ll)gridDim.x * blockDim.x) >> 5);
    unsigned n_slots = *slab_count;
    if (n_slots > slab_cap) n_slots = slab_cap;
    unsigned long long my_band = 0, my_rej = 0;
    int n_list = 0;  // warp-uniform
    auto drain = [&]() {
        __syncwarp();
        for (int p0 = 0; p0 < n_list; p0 += 32) {
            const bool have = p0 + lane < n_list;
            int4 e = make_int4(0, 0, 0, 0);
            if (have) e = list[p0 + lane];
            // load slot t of this lane: row rr = 8 t + lane / 4 of the 64 (pair rr / 2, its i or j), doubles 2 (lane & 3), + 1
            const double *src[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int rr = 8 * t + (lane >> 2);
                const int pi = __shfl_sync(full, e.x, rr >> 1), pj = __shfl_sync(full, e.y, rr >> 1);
                src[t] = x + (ll)((rr & 1) ? pj : pi) * d + 2 * (lane & 3);
            }
            double sum = 0.0;
            for (int k0 = 0; k0 < d; k0 += 8) {
                const int kk = k0 + 2 * (lane & 3);
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int rr = 8 * t + (lane >> 2);
                    const double v0 = kk < d ? src[t][k0] : 0.0, v1 = kk + 1 < d ? src[t][k0 + 1] : 0.0;
                    stage[2 * (lane & 3)][rr] = v0;       // a pair that does not exist reads row 0: harmless
                    stage[2 * (lane & 3) + 1][rr] = v1;
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const double2 ab = *reinterpret_cast<const double2 *>(&stage[k][2 * lane]);
                    const double t = __dsub_rn(ab.x, ab.y);   // 0 - 0 past the last dimension: adds +0.0
                    sum = __dadd_rn(sum, __dmul_rn(t, t));
                }
                __syncwarp();
            }
            if (SQRT) sum = __dsqrt_rn(sum);
            if (have && !(sum <= rc)) {
                atomicAnd(slab_words + (unsigned)e.z, ~(1u << e.w));
                ++my_rej;
            }
        }
        n_list = 0;
        __syncwarp();
    };
    for (unsigned slot = (unsigned)((blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5); slot < n_slots; slot += nwarps) {
        const int2 inf = info[slot];
        if (inf.x == 0 || !(inf.y & 0x40000000)) continue;
        const ll g = inf.x - 1, J = inf.y & 0x3fffffff;
        uint4 bd = bslabs[(size_t)slot * 32 + lane];
        my_band += __popc(bd.x) + __popc(bd.y) + __popc(bd.z) + __popc(bd.w);
        for (;;) {
            // this lane's next pending pair
            int w = -1, bit = 0;
            if (bd.x) { w = 0; bit = __ffs(bd.x) - 1; bd.x &= bd.x - 1; }
            else if (bd.y) { w = 1; bit = __ffs(bd.y) - 1; bd.y &= bd.y - 1; }
            else if (bd.z) { w = 2; bit = __ffs(bd.z) - 1; bd.z &= bd.z - 1; }
            else if (bd.w) { w = 3; bit = __ffs(bd.w) - 1; bd.w &= bd.w - 1; }
            const unsigned todo = __ballot_sync(full, w >= 0);
            if (!todo) break;
            if (n_list + 32 > RL_CAP) drain();
            if (w >= 0)
                list[n_list + __popc(todo & lt)] = make_int4((int)(g * 32 + lane), (int)(J * TT + 32 * w + bit),
                                                            (int)((slot * 32u + (unsigned)lane) * 4u + (unsigned)w), bit);
            n_list += __popc(todo);
        }
    }
    drain();
    for (int o = 16; o; o >>= 1) {
        my_band += __shfl_xor_sync(full, my_band, o);
        my_rej += __shfl_xor_sync(full, my_rej, o);
    }
    if (lane == 0 && my_band) atomicAdd(n_band, my_band);
    if (lane == 0 && my_rej) atomicAdd(n_rejected, my_rej);
}

// ---------------------------------------------------------------------------------------------------
// bitmaps -> CSR
// ---------------------------------------------------------------------------------------------------
// Degrees, slab by slab (a warp per slab, slots strided): the rows of the slab add their popcounts to up[] (entries
// above the diagonal), its columns -- after the 32 x 32 bit transposes, which are kept for the fill pass -- to low[]
// (entries below the diagonal).  Atomics on 4-byte counters, no return value, only for non-zero counts.
__global__ void __launch_bounds__(256)
slab_count_kernel(const int2 *__restrict__ info, const unsigned *__restrict__ slab_count, unsigned slab_cap,
                  const uint4 *__restrict__ slabs, unsigned *__restrict__ tslabs, int *__restrict__ up, int *__restrict__ low) {
    const int lane = threadIdx.x & 31;
    const unsigned nwarps = (unsigned)(((ll)gridDim.x * blockDim.x) >> 5);
    unsigned n_slots = *slab_count;
    if (n_slots > slab_cap) n_slots = slab_cap;
    for (unsigned slot = (unsigned)((blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5); slot < n_slots; slot += nwarps) {
        const int2 inf = info[slot];
        if (inf.x == 0) continue;
        const ll g = inf.x - 1, J = inf.y & 0x3fffffff;
        const uint4 h = slabs[(size_t)slot * 32 + lane];
        const int c = __popc(h.x) + __popc(h.y) + __popc(h.z) + __popc(h.w);
        if (c) atomicAdd(up + g * 32 + lane, c);
        unsigned *ts = tslabs + (size_t)slot * 128;
        const unsigned wds[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const unsigned m = transpose32(wds[w], lane);
            ts[32 * w + lane] = m;
            if (m) atomicAdd(low + J * TT + 32 * w + lane, __popc(m));
        }
    }
}

__global__ void deg_sum_kernel(const int *__restrict__ up, const int *__restrict__ low, ll n, int *__restrict__ deg) {
    const ll r = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (r <= n) deg[r] = r < n ? up[r] + low[r] + 1 : 0;  // deg[n] = 0: the scan leaves nnz in row_ptr[n]
}

// One CTA per row tile T; warps 0-3 write the part of the rows below the diagonal (one row quarter each, one row per
// lane): bits of the row's COLUMN in the transposed slabs above it (dirT[T][.], row quarters ascending, the blocks
// slab_count_kernel left behind); warps 4-7 the row itself and the part above: the row's own slabs (dir[g][.], column tiles
// ascending).  Eight slabs are fetched at a time and the next 32 directory entries are under way while these are served (the
// walk is a chain of dependent L2 reads: directory -> slab).  low[r] (entries below the diagonal, from the count pass)
// tells the upper warps where their part starts.  Rows come out ascending.
#define RB 8   // slabs fetched at a time
__global__ void __launch_bounds__(256)
tile_rows_kernel(const unsigned *__restrict__ dir, const unsigned *__restrict__ dirT, const uint4 *__restrict__ slabs,
                 const unsigned *__restrict__ tslabs, ll n, ll n_tiles, const int *__restrict__ low,
                 const ll *__restrict__ row_ptr, int *__restrict__ col_idx) {
    const unsigned full = 0xffffffffu;
    const ll T = blockIdx.x;
    const int warp = threadIdx.x >> 5, q = warp & 3, lane = threadIdx.x & 31;
    const bool upper = warp >= 4;
    const ll g = T * 4 + q, r = g * 32 + lane, G = 4 * n_tiles;
    int cnt = 0;
    int *out = nullptr;  // dead rows (r >= n) have no bits anywhere: nothing is ever written through it
    if (r < n) out = col_idx + row_ptr[r] + (upper ? low[r] : 0);
    if (!upper) {
        const unsigned *dt = dirT + T * G;
        unsigned v = lane <= g ? dt[lane] : 0u;
        for (ll g0 = 0; g0 <= g; g0 += 32) {
            const ll gpn = g0 + 32 + lane;
            const unsigned vn = gpn <= g ? dt[gpn] : 0u;
            unsigned nz = __ballot_sync(full, v != 0u);
            while (nz) {
                int b[RB];
                unsigned xw[RB];
#pragma unroll
                for (int k = 0; k < RB; ++k) {
                    b[k] = nz ? __ffs(nz) - 1 : -1;
                    nz &= nz - 1;  // no-op at 0
                }
#pragma unroll
                for (int k = 0; k < RB; ++k) {
                    const unsigned slot = (__shfl_sync(full, v, b[k] < 0 ? 0 : b[k]) & ~BAND_FLAG) - 1u;
                    xw[k] = b[k] < 0 ? 0u : tslabs[(size_t)slot * 128 + q * 32 + lane];
                }
#pragma unroll
                for (int k = 0; k < RB; ++k) {
                    if (b[k] < 0) continue;  // warp-uniform
                    unsigned m = xw[k];      // bit l: row l of quarter g0 + b has column r
                    const int base = (int)((g0 + b[k]) * 32);
                    while (m) {
                        const int bit = __ffs(m) - 1;
                        m &= m - 1;
                        out[cnt++] = base + bit;
                    }
                }
            }
            v = vn;
        }
    } else {
        if (r < n) {
            out[0] = (int)r;
            cnt = 1;
        }
        const unsigned *dr = dir + g * n_tiles;
        const ll Jb = T & ~31ll;
        unsigned v = (Jb + lane >= T && Jb + lane < n_tiles) ? dr[Jb + lane] : 0u;
        for (ll J0 = Jb; J0 < n_tiles; J0 += 32) {
            const ll Jn = J0 + 32 + lane;
            const unsigned vn = Jn < n_tiles ? dr[Jn] : 0u;
            unsigned nz = __ballot_sync(full, v != 0u);
            while (nz) {
                int b[RB];
                uint4 m[RB];
#pragma unroll
                for (int k = 0; k < RB; ++k) {
                    b[k] = nz ? __ffs(nz) - 1 : -1;
                    nz &= nz - 1;
                }
#pragma unroll
                for (int k = 0; k < RB; ++k) {
                    const unsigned slot = (__shfl_sync(full, v, b[k] < 0 ? 0 : b[k]) & ~BAND_FLAG) - 1u;
                    m[k] = b[k] < 0 ? make_uint4(0u, 0u, 0u, 0u) : slabs[(size_t)slot * 32 + lane];
                }
#pragma unroll
                for (int k = 0; k < RB; ++k) {
                    if (b[k] < 0) continue;
                    const int base = (int)((J0 + b[k]) * TT);
                    const unsigned wds[4] = {m[k].x, m[k].y, m[k].z, m[k].w};
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        unsigned xw = wds[w];
                        while (xw) {
                            const int bit = __ffs(xw) - 1;
                            xw &= xw - 1;
                            out[cnt++] = base + 32 * w + bit;
                        }
                    }
                }
            }
            v = vn;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static int make_map(CUtensorMap *map, void *base, ll n_pad, int kt) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) CNN_FAIL(CNNB200_E_CUDA, "tensor search: cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)kt, (cuuint64_t)n_pad};
    cuuint64_t strides[1] = {(cuuint64_t)kt * 2};
    cuuint32_t box[2] = {KATOM, TT};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) CNN_FAIL(CNNB200_E_CUDA, "tensor search: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CNNB200_OK;
}

int cnn_tensor_search(cnnb200_ctx *ctx, const double *d_x, ll n, ll d, double radius_cutoff, int metric) {
    cudaStream_t st = ctx->stream;
    const int atoms = (int)((3 * d + KATOM - 1) / KATOM);
    if (atoms > MAX_ATOMS)
        CNN_FAIL(CNNB200_E_RANGE, "tensor search supports up to %d dimensions (got %lld); use the tiled method",
                 MAX_ATOMS * KATOM / 3, (long long)d);
    const int kt = atoms * KATOM;
    const ll n_tiles = (n + TT - 1) / TT, n_pad = n_tiles * TT;
    if (n_tiles > 16384)  // the two slab directories take 32 n_tiles^2 bytes
        CNN_FAIL(CNNB200_E_RANGE, "tensor search serves up to %d points (got %lld); use the tiled method", 16384 * TT, (long long)n);
    const bool sq = metric == CNNB200_METRIC_EUCLIDEAN;
    const double R = sq ? radius_cutoff * radius_cutoff : radius_cutoff;

    // ---- sort the points by (nearest pivot, distance to it); everything below works in that order ---------------
    const bool reorder = n >= 4096 && ctx->tensor_reorder;
    const unsigned *skip = nullptr;
    const int words = (int)((n_tiles + 31) / 32);
    if (reorder) {
        Span sp(ctx, T_GRID_BUILD);
        const int n_piv = N_PIVOTS;
        CNN_TRY(cnn_ensure(ctx, ctx->b_perm, sizeof(int) * (size_t)n));
        CNN_TRY(cnn_ensure(ctx, ctx->b_idx, sizeof(int) * (size_t)n));
        CNN_TRY(cnn_ensure(ctx, ctx->b_lab_keys, sizeof(unsigned long long) * (size_t)n));
        CNN_TRY(cnn_ensure(ctx, ctx->b_lab_keys2, sizeof(unsigned long long) * (size_t)n));
        CNN_TRY(cnn_ensure(ctx, ctx->b_xperm, sizeof(double) * (size_t)n * (size_t)d));
        unsigned long long *pk = ctx->b_lab_keys.as<unsigned long long>(), *pk2 = ctx->b_lab_keys2.as<unsigned long long>();
        pivot_assign_kernel<<<ctx->sm_count * 8, 256, sizeof(float) * (size_t)n_piv * (size_t)(d + 1), st>>>(d_x, n, (int)d, n_piv, pk,
                                                                                                      ctx->b_idx.as<int>());
        size_t need = 0;
        CNN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, pk, pk2, ctx->b_idx.as<int>(), ctx->b_perm.as<int>(), (int)n, 0, 40, st));
        CNN_TRY(cnn_ensure(ctx, ctx->b_cub, need));
        size_t cap = ctx->b_cub.cap;
        CNN_CUDA(cub::DeviceRadixSort::SortPairs(ctx->b_cub.p, cap, pk, pk2, ctx->b_idx.as<int>(), ctx->b_perm.as<int>(), (int)n, 0, 40, st));
        gather_rows_kernel<<<cnn_blocks(n * d, 256), 256, 0, st>>>(d_x, ctx->b_perm.as<int>(), n, (int)d, ctx->b_xperm.as<double>());
        d_x = ctx->b_xperm.as<double>();
        // bounding balls of the tiles and the pairs of tiles that are too far apart to matter
        CNN_TRY(cnn_ensure(ctx, ctx->b_cell_start, sizeof(double) * 2 * (size_t)n_tiles * (size_t)(d + 1)));
        CNN_TRY(cnn_ensure(ctx, ctx->b_cell_count, sizeof(unsigned) * (size_t)n_tiles * (size_t)words));
        double *centre = ctx->b_cell_start.as<double>(), *radius = centre + n_tiles * d;
        double *centre2 = radius + n_tiles, *radius2 = centre2 + n_tiles * d;
        tile_ball_kernel<<<(int)n_tiles, TT, 0, st>>>(d_x, pk2, n, (int)d, centre, radius, centre2, radius2);
        const double r_true = sq ? radius_cutoff : sqrt(radius_cutoff);
        tile_skip_kernel<<<cnn_blocks(n_tiles * words * 32, 256), 256, 0, st>>>(centre, radius, n_tiles, (int)d, words, r_true,
                                                                               ctx->b_cell_count.as<unsigned>());
        tile_skip_split_kernel<<<dim3((unsigned)n_tiles, (unsigned)((n_tiles + 127) / 128)), 128, 0, st>>>(centre, radius, centre2, radius2, n_tiles, (int)d, words, r_true,
                                                             ctx->b_cell_count.as<unsigned>());
        skip = ctx->b_cell_count.as<unsigned>();
        CNN_LAUNCHED(ctx, 5);
        CNN_KERNEL_CHECK();
    }

    // ---- operands -----------------------------------------------------------------------------
    CNN_TRY(cnn_ensure(ctx, ctx->b_xs, sizeof(__half) * (size_t)n_pad * kt * 2));
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp1, sizeof(float) * (size_t)n_pad));
    CNN_TRY(cnn_ensure(ctx, ctx->b_bbox, sizeof(double) * (size_t)(CM_MAXD + 2 * CM_MAXD + 16)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_deg, sizeof(int) * (size_t)(n_pad + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_row_ptr, sizeof(ll) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_stats, sizeof(unsigned long long) * ST_N));
    __half *A = ctx->b_xs.as<__half>(), *B = A + (size_t)n_pad * kt;
    float *norm = ctx->b_tmp1.as<float>();
    double *mean = ctx->b_bbox.as<double>();
    unsigned long long *mm = (unsigned long long *)(mean + CM_MAXD);  // [2][CM_MAXD] codes of the column minima / maxima
    int *maxbits = (int *)(mm + 2 * CM_MAXD);
    int *scale_e = maxbits + 1, *bad = maxbits + 2;
    float *thr = (float *)(maxbits + 4);
    unsigned *slab_count = (unsigned *)(maxbits + 6);
    unsigned long long *band_count = (unsigned long long *)(maxbits + 8), *rej_count = band_count + 1;

    // |d2~ - d2| in units of the largest squared norm M ("Exactness of the tensor-core filter"):
    //   6.1 * 2^-22                      the products the split drops (lo.lo) and the residuals of the halves;
    //   68 * 2^-23 * (n_hh * 1.002 + n_small * 2^-10.9)
    //                                    fp32 accumulation in the tensor core: per K = 16 instruction 17 addends, each
    //                                    off by at most one ulp of the largest magnitude in play -- M for the
    //                                    instructions that hold hi.hi products (and the zero padding behind them),
    //                                    2^-10.9 M before that --, twice for 2 x_i.x_j, and a factor 2 of head-room;
    //   2^-20                            the norms rounded to f32, the roundings of nc - 2 acc and thr - nr.
    const int n_instr = atoms * 4, n_small = (int)(2 * d / 16), n_hh = n_instr - n_small;
    double coef = 6.1 * ldexp(1.0, -22) + 68.0 * ldexp(1.0, -23) * (n_hh * 1.002 + n_small * 5.3e-4) + ldexp(1.0, -20);
    if (const char *e = getenv("CNNB200_TENSOR_EPS_SCALE")) coef *= atof(e);  // measurement only: how much head-room the bound has
    {
        Span sp(ctx, T_GRID_BUILD);
        CNN_CUDA(cudaMemsetAsync(ctx->b_stats.p, 0, sizeof(unsigned long long) * ST_N, st));
        CNN_CUDA(cudaMemsetAsync(mean, 0, sizeof(double) * (size_t)(CM_MAXD + 2 * CM_MAXD + 16), st));
        CNN_CUDA(cudaMemsetAsync(mm, 0xff, sizeof(unsigned long long) * CM_MAXD, st));  // minima start at the largest code
        col_sum_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(d_x, n, (int)d, mean, mm);
        col_mean_kernel<<<1, CM_MAXD, 0, st>>>(mean, mm, (int)d, n, scale_e);
        const int sb = cnn_blocks(n_pad, 8) < ctx->sm_count * 32 ? cnn_blocks(n_pad, 8) : ctx->sm_count * 32;
        if (d % 2 == 0) split_kernel<true><<<sb, 256, 0, st>>>(d_x, mean, scale_e, n, n_pad, (int)d, kt, A, B, norm, maxbits);
        else split_kernel<false><<<sb, 256, 0, st>>>(d_x, mean, scale_e, n, n_pad, (int)d, kt, A, B, norm, maxbits);
        tensor_setup_kernel<<<1, 1, 0, st>>>(maxbits, scale_e, R, coef, thr, bad);
        CNN_LAUNCHED(ctx, 4);
        CNN_KERNEL_CHECK();
    }

    CUtensorMap mapA, mapB;
    CNN_TRY(make_map(&mapA, A, n_pad, kt));
    CNN_TRY(make_map(&mapB, B, n_pad, kt));

    const size_t dir_entries = (size_t)4 * n_tiles * n_tiles;
    CNN_TRY(cnn_ensure(ctx, ctx->b_tdir, sizeof(unsigned) * 2 * dir_entries));
    unsigned *dir = ctx->b_tdir.as<unsigned>(), *dirT = dir + dir_entries;

    TensorParams P;
    memset(&P, 0, sizeof(P));
    P.norm = norm;
    P.thr = thr;
    P.n = n;
    P.n_tiles = n_tiles;
    P.atoms = atoms;
    {   // 2 A tiles + nsb atom stages of B + 12 KB of barriers / norms / alignment must fit 227 KB
        const int room = (227 * 1024 - 12 * 1024 - 2 * atoms * ATOM_BYTES) / ATOM_BYTES;
        P.nsb = room < NSB ? room : NSB;
    }
    P.skip = skip;
    P.words = words;
    P.dir = dir;
    P.dirT = dirT;
    P.slab_count = slab_count;
    P.dbg = getenv("CNNB200_TENSOR_DBG") ? atoi(getenv("CNNB200_TENSOR_DBG")) : 0;

    const size_t smem = (size_t)(2 * atoms + P.nsb) * ATOM_BYTES + 1024 /*alignment*/ + 32 * 8 +
                        EPI_WARPS * 2 * TT * sizeof(float) + 64;
    void (*tiles_kernel)(const CUtensorMap, const CUtensorMap, const TensorParams) =
        atoms == 1 ? tensor_tiles_kernel<1> : atoms == 2 ? tensor_tiles_kernel<2> : atoms == 3 ? tensor_tiles_kernel<3> : tensor_tiles_kernel<4>;
    CNN_CUDA(cudaFuncSetAttribute(tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    // ---- the pass over the upper triangle, the exact decisions, the degrees (all of it repeated once with more
    //      room should the slabs not fit: the host learns that together with nnz) -------------------------------------
    ll slab_cap = ctx->tensor_band_cap > 0 ? ctx->tensor_band_cap : (256 * n_tiles > 4096 ? 256 * n_tiles : 4096);
    unsigned *hb;
    CNN_TRY(cnn_pinned(ctx, 256, (void **)&hb));
    hb += 32;  // the first 128 bytes belong to cnn_scan_deg_to_row_ptr
    ll nnz = 0, max_deg = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (slab_cap > (1ll << 24))  // word offsets into the slab store travel as 32-bit values (tile_resolve_kernel)
            CNN_FAIL(CNNB200_E_RANGE, "tensor search: more than 2^24 hit slabs (%lld); use the tiled method", (long long)slab_cap);
        CNN_TRY(cnn_ensure(ctx, ctx->b_tslab, sizeof(uint4) * 32 * (size_t)slab_cap));
        CNN_TRY(cnn_ensure(ctx, ctx->b_tband, sizeof(uint4) * 32 * 2 * (size_t)slab_cap));  // borderline slabs | transposed hit slabs
        P.slabs = ctx->b_tslab.as<uint4>();
        P.bslabs = ctx->b_tband.as<uint4>();
        unsigned *tslabs = (unsigned *)(P.bslabs + 32 * (size_t)slab_cap);
        P.slab_cap = (unsigned)slab_cap;
        CNN_TRY(cnn_ensure(ctx, ctx->b_tmp0, sizeof(int2) * (size_t)slab_cap + sizeof(int) * 2 * (size_t)n_pad));
        P.info = ctx->b_tmp0.as<int2>();
        int *low = (int *)(P.info + slab_cap), *up = low + n_pad;
        CNN_CUDA(cudaMemsetAsync(P.info, 0, sizeof(int2) * (size_t)slab_cap + sizeof(int) * 2 * (size_t)n_pad, st));
        CNN_CUDA(cudaMemsetAsync(dir, 0, sizeof(unsigned) * 2 * dir_entries, st));
        CNN_CUDA(cudaMemsetAsync(slab_count, 0, 2 * sizeof(unsigned) + 2 * sizeof(unsigned long long), st));
        CNN_CUDA(cudaMemsetAsync(ctx->b_deg.as<int>() + n, 0, sizeof(int), st));
        {
            Span sp(ctx, T_SEARCH_COUNT);
            tiles_kernel<<<(int)((n_tiles + 1) / 2), NUM_THREADS, smem, st>>>(mapA, mapB, P);
            CNN_LAUNCHED(ctx, 1);
            CNN_KERNEL_CHECK();
        }
        {
            Span sp(ctx, T_SCAN);
            const int rb = ctx->sm_count * 12;
            if (sq)
                tile_resolve_kernel<true><<<rb, RL_WARPS * 32, 0, st>>>(d_x, (int)d, P.info, slab_count, P.slab_cap, (unsigned *)P.slabs, P.bslabs,
                                                              radius_cutoff, band_count, rej_count);
            else
                tile_resolve_kernel<false><<<rb, RL_WARPS * 32, 0, st>>>(d_x, (int)d, P.info, slab_count, P.slab_cap, (unsigned *)P.slabs, P.bslabs,
                                                               radius_cutoff, band_count, rej_count);
            slab_count_kernel<<<rb, 256, 0, st>>>(P.info, slab_count, P.slab_cap, P.slabs, tslabs, up, low);
            deg_sum_kernel<<<cnn_blocks(n + 1, 256), 256, 0, st>>>(up, low, n, ctx->b_deg.as<int>());
            CNN_LAUNCHED(ctx, 3);
            CNN_KERNEL_CHECK();
        }
        CNN_CUDA(cudaMemcpyAsync(hb, bad, 4 * sizeof(int) + 2 * sizeof(unsigned) + 2 * sizeof(unsigned long long),
                                 cudaMemcpyDeviceToHost, st));
        CNN_TRY(cnn_scan_deg_to_row_ptr(ctx, ctx->b_deg.as<int>(), n, ctx->b_row_ptr.as<ll>(), &nnz, &max_deg));  // synchronises
        if (((int *)hb)[0]) CNN_FAIL(CNNB200_E_INVALID, "neighbour search: coordinates contain NaN or infinity");
        const ll used = (ll)hb[4];
        if (used <= slab_cap) break;
        if (attempt == 1) CNN_FAIL(CNNB200_E_STATE, "tensor search: slab store overflowed twice");
        slab_cap = used + used / 16 + 1024;
    }
    ctx->tensor_band_cap = slab_cap;
    if (nnz >= (1ll << 31) - 1) CNN_FAIL(CNNB200_E_RANGE, "tensor search: more than 2^31 neighbour entries");
    CNN_TRY(cnn_ensure(ctx, ctx->b_col_idx, sizeof(int) * (size_t)nnz));
    {
        Span sp(ctx, T_SEARCH_FILL);
        tile_rows_kernel<<<(int)n_tiles, 256, 0, st>>>(dir, dirT, P.slabs, (const unsigned *)(P.bslabs + 32 * (size_t)slab_cap), n, n_tiles,
                                                       (const int *)(ctx->b_tmp0.as<int2>() + slab_cap), ctx->b_row_ptr.as<ll>(),
                                                       ctx->b_col_idx.as<int>());
        CNN_LAUNCHED(ctx, 1);
        CNN_KERNEL_CHECK();
    }
    ctx->stats[ST_RECHECK] = (ll)((unsigned long long *)(hb + 6))[0];
    ctx->stats[ST_DIST] = (ll)hb[5] * TT * TT;  // pairs in the tiles that were visited (the others were ruled out by the bounding balls)
    ctx->n = n;
    ctx->nnz = nnz;
    ctx->max_deg = max_deg;
    ctx->row_ptr = ctx->b_row_ptr.as<ll>();
    ctx->col_idx = ctx->b_col_idx.as<int>();
    ctx->perm = reorder ? ctx->b_perm.as<int>() : nullptr;
    return CNNB200_OK;
}
