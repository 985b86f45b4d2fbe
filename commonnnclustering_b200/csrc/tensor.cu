// tensor.cu -- K3: radius-neighbour search for high-dimensional points on the 5th-generation
// tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA) with an exact f64
// recheck of every pair near the cutoff.
//
// Bulk replacement of NeighboursGetterExtBruteForce._get (reference _types.pyx:1692-1710) with
// MetricExtEuclideanReduced / MetricExtEuclidean (_types.pyx:2253-2267, :2173-2187).  The
// reference evaluates N^2 distances in f64, one virtual call each; here the N^2 part is a dense
// contraction d2(i,j) = |x_i|^2 + |x_j|^2 - 2 x_i.x_j and runs as a bf16 GEMM:
//
//   * the (mean-centred) coordinates are split x = hi + lo into two bf16 parts; the three
//     products hi.hi + hi.lo + lo.hi are ONE MMA chain over the concatenated K axis
//     A' = [hi | hi | lo], B' = [hi | lo | hi] (K = 3 D, fp32 accumulation in TMEM);
//   * the epilogue reads the accumulator tile with tcgen05.ld, forms d2~ and classifies:
//       d2~ <= R - eps  -> neighbour,   d2~ > R + eps -> not a neighbour,
//       otherwise       -> the pair is re-evaluated right there with the reference's own f64
//                          expression (sequential sum, no FMA, sqrt for the non-reduced metric);
//     eps is a bound on |d2~ - d2| derived from the largest norm, so the neighbour
//     SETS are exactly the reference's;
//   * a CTA owns 128 rows and walks over all 128-column tiles in ascending order, so each CSR
//     row is emitted ascending (count pass -> scan -> fill pass, as for the other producers).
//
// Warp roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (one
// elected lane) + TMEM allocation, warps 2-5 = epilogue (one TMEM lane quarter each).
// Pipelines: smem B stages full/empty (TMA <-> MMA), TMEM accumulator stages full/empty
// (MMA <-> epilogue), two stages each.
#include <cub/cub.cuh>
#include <cuda.h>
#include <string.h>
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"

#define TT 128            // tile edge (rows of A' = UMMA M, rows of B' = UMMA N)
#define KATOM 64          // bf16 elements per 128-byte swizzle atom
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
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 in, f32 accumulate (kind::f16)
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------------
// column sums with whole rows read coalesced (a warp per row, lanes over the dimensions; a block per dimension read
// x with a stride of one row: 16x the bytes, 0.15 ms on C3), then mean = sum / n.  The order of the atomic adds
// is not fixed; the mean only centres the data (any centre gives the same neighbour sets: the filter's band is
// derived from the centred norms and every borderline pair is decided in f64 on the original coordinates).
#define CM_MAXD 128
__global__ void __launch_bounds__(256)
col_sum_kernel(const double *__restrict__ x, ll n, int d, double *sum) {
    __shared__ double sh[8][CM_MAXD];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double acc[CM_MAXD / 32] = {0.0, 0.0, 0.0, 0.0};
    const ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5, nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll i = warp; i < n; i += nwarps)
#pragma unroll
        for (int u = 0; u < CM_MAXD / 32; ++u)
            if (lane + 32 * u < d) acc[u] += x[i * d + lane + 32 * u];
#pragma unroll
    for (int u = 0; u < CM_MAXD / 32; ++u) sh[w][lane + 32 * u] = acc[u];
    __syncthreads();
    if ((int)threadIdx.x < d) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += sh[k][threadIdx.x];
        atomicAdd(sum + threadIdx.x, t);
    }
}

__global__ void col_mean_kernel(double *mean, int d, ll n) {
    if ((int)threadIdx.x < d) mean[threadIdx.x] /= (double)n;
}

// A'[i] = [hi | hi | lo | 0..], B'[i] = [hi | lo | hi | 0..] (bf16, row length kt), norm[i] = |x_i - mean|^2;
// padding rows (i >= n) are zero with an infinite norm so that they never come near the cutoff.
// One warp per row: lanes stride over the dimensions, so the f64 row comes in as one coalesced read and every bf16
// segment goes out as contiguous 2-byte stores (a thread per row wrote 2 bytes at a stride of one row: 1.1 ms on C3).
__global__ void __launch_bounds__(256)
split_kernel(const double *__restrict__ x, const double *__restrict__ mean, ll n, ll n_pad, int d,
             int kt, __nv_bfloat16 *A, __nv_bfloat16 *B, float *norm, int *maxnorm_bits) {
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
    for (ll i = warp; i < n_pad; i += nwarps) {
        __nv_bfloat16 *a = A + i * kt, *b = B + i * kt;
        if (i >= n) {
            for (int k = lane; k < kt; k += 32) { a[k] = zero; b[k] = zero; }
            if (lane == 0) norm[i] = INFINITY;
            continue;
        }
        double s = 0;
        for (int k = lane; k < d; k += 32) {
            const double xc = x[i * d + k] - mean[k];
            s += xc * xc;
            const __nv_bfloat16 hi = __float2bfloat16_rn((float)xc);
            const __nv_bfloat16 lo = __float2bfloat16_rn((float)(xc - (double)__bfloat162float(hi)));
            a[k] = hi; a[d + k] = hi; a[2 * d + k] = lo;
            b[k] = hi; b[d + k] = lo; b[2 * d + k] = hi;
        }
        for (int k = 3 * d + lane; k < kt; k += 32) { a[k] = zero; b[k] = zero; }
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            const float sf = (float)s;
            norm[i] = sf;
            atomicMax(maxnorm_bits, __float_as_int(sf));
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// exact f64 distance, the reference's expression (_types.pyx:2253-2267)
// ---------------------------------------------------------------------------------------------------
template <bool SQRT>
__device__ __noinline__ bool exact_within(const double *__restrict__ x, int d, ll i, ll j, double rc) {
    const double *a = x + i * d, *b = x + j * d;
    double s = 0.0;
    for (int k = 0; k < d; ++k) {
        double t = __dsub_rn(a[k], b[k]);
        s = __dadd_rn(s, __dmul_rn(t, t));
    }
    if (SQRT) s = __dsqrt_rn(s);
    return s <= rc;
}

// ---------------------------------------------------------------------------------------------------
// reordering and tile pruning
// ---------------------------------------------------------------------------------------------------
// High-dimensional data that is worth clustering is clustered: after sorting the points by (nearest pivot, distance
// to it) the 128-row tiles are spatially compact, and a tile pair whose bounding balls are further apart than the
// radius cannot contain a neighbour pair (triangle inequality) -- it is skipped by all three warp roles.  The CSR
// is produced in the sorted index space together with perm[] (sorted position -> caller index), as in the grid path.
#define N_PIVOTS 64

// one warp per point: the row comes in with one coalesced read (lanes stride over the dimensions); each lane then owns
// two pivots (n_piv <= 64) and walks the dimensions, the point's coordinate broadcast by a shuffle, the pivots in shared
// memory as floats [d][n_piv] -- no reduction per pivot, one arg-min at the end
__global__ void __launch_bounds__(256)
pivot_assign_kernel(const double *__restrict__ x, ll n, int d, int n_piv, unsigned long long *keys, int *idx) {
    extern __shared__ float piv[];  // [d][n_piv]
    const unsigned full = 0xffffffffu;
    for (int t = threadIdx.x; t < n_piv * d; t += blockDim.x) {
        const int p = t / d, k = t - p * d;
        piv[k * n_piv + p] = (float)x[(ll)((double)p * (double)n / (double)n_piv) * d + k];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const bool has0 = lane < n_piv, has1 = lane + 32 < n_piv;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll i = warp; i < n; i += nwarps) {
        float xi[3];  // d <= 85 < 96: at most 3 coordinates per lane
#pragma unroll
        for (int u = 0; u < 3; ++u) xi[u] = lane + 32 * u < d ? (float)x[i * d + lane + 32 * u] : 0.f;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int kn = d - 32 * u < 32 ? d - 32 * u : 32;
            for (int kk = 0; kk < kn; ++kk) {
                const float xv = __shfl_sync(full, xi[u], kk);
                const float *row = piv + (32 * u + kk) * n_piv;
                const float t0 = xv - (has0 ? row[lane] : 0.f), t1 = xv - (has1 ? row[lane + 32] : 0.f);
                s0 = fmaf(t0, t0, s0);
                s1 = fmaf(t1, t1, s1);
            }
        }
        float best = has0 ? s0 : INFINITY;
        int arg = lane;
        if (has1 && s1 < best) { best = s1; arg = lane + 32; }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ob = __shfl_xor_sync(full, best, o);
            const int oa = __shfl_xor_sync(full, arg, o);
            if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        if (lane == 0) {
            keys[i] = ((unsigned long long)arg << 32) | (unsigned long long)__float_as_uint(best);  // best >= 0: bits are ordered
            idx[i] = (int)i;
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

// one block per tile: centre (mean of its live rows) and radius (largest distance of a row to the centre, rounded up)
__global__ void tile_ball_kernel(const double *__restrict__ x, ll n, int d, double *centre, double *radius) {
    __shared__ double c[128];
    __shared__ double rmax[TT];
    const ll T = blockIdx.x;
    const ll r0 = T * TT;
    const int rows = (int)((n - r0) < TT ? (n - r0) : TT);
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < rows; ++r) s += x[(r0 + r) * d + k];
        c[k] = s / rows;
        centre[(ll)k * gridDim.x + T] = c[k];  // [d][n_tiles]: tile_skip_kernel reads 32 consecutive tiles per dimension
    }
    __syncthreads();
    double dist = 0.0;
    if ((int)threadIdx.x < rows) {
        double s = 0.0;
        for (int k = 0; k < d; ++k) {
            const double t = x[(r0 + threadIdx.x) * d + k] - c[k];
            s += t * t;
        }
        dist = sqrt(s);
    }
    rmax[threadIdx.x] = dist;
    __syncthreads();
    for (int o = TT / 2; o; o >>= 1) {
        if ((int)threadIdx.x < o) rmax[threadIdx.x] = fmax(rmax[threadIdx.x], rmax[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) radius[T] = rmax[0] * (1.0 + 1e-12) + 1e-300;
}

// bit J of row I set <=> tile pair (I, J) can be skipped; one warp per 32 column tiles.  Only the words at or above the
// diagonal are ever read (TileIter starts at the word of I0), the others are left alone.
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
    if (J < n_tiles) {
        double s = 0.0;
        for (int k = 0; k < d; ++k) {
            const double t = centre[(ll)k * n_tiles + I] - centre[(ll)k * n_tiles + J];
            s += t * t;
        }
        // any i in I, j in J: |x_i - x_j| >= |c_I - c_J| - rho_I - rho_J
        sk = sqrt(s) * (1.0 - 1e-12) - radius[I] - radius[J] > r_true * (1.0 + 1e-9);
    }
    const unsigned m = __ballot_sync(0xffffffffu, sk);
    if (lane == 0) skip[gw] = m;
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
#define NSB 3   // barrier slots reserved for the shared-memory stages of the B operand (2 are used; 1 for 4-atom tiles)
#define NST 4   // TMEM accumulator stages of 128 columns (MMA <-> epilogue): the whole 512-column TMEM
#ifndef CNN_EPI_WARPS
#define CNN_EPI_WARPS 8   // measured r2 on C3: 16 warps 3.64 ms, 8 warps 3.48 ms
#endif
#define EPI_WARPS CNN_EPI_WARPS  // epilogue warps: 8 = one per (row tile, TMEM lane quarter), all 128 columns each;
                                 // 16 = two per (row tile, lane quarter), 64 columns each (more warps to hide the
                                 // tcgen05.ld / shared-memory latency of an epilogue that, not the MMA, bounds the kernel)
#define PAIR_BUF (128 / EPI_WARPS)  // pairs a thread collects before it reserves space in the global list
#define NUM_THREADS ((2 + EPI_WARPS) * 32)
#define BAND_FLAG 0x80000000u

struct TensorParams {
    const float *norm;    // [n_pad]
    ll n, n_tiles;
    int atoms, nsb;
    float r2_lo, r2_hi;   // R - eps, R + eps in squared-distance units
    const unsigned *skip; // [n_tiles][words] tile pairs that cannot hold a neighbour pair (nullptr: visit all)
    int words;
    int2 *pairs;          // out: (i, j) with i < j; j carries BAND_FLAG when the pair is inside the +-eps band
    unsigned long long *pair_count;  // running length of `pairs` (may exceed pair_cap: then nothing is stored)
    ll pair_cap;
};

// SINGLE pass over the upper triangle of the tile grid.  CTA c owns the TWO adjacent row tiles 2c and 2c+1 (256 rows,
// both A tiles stay in shared memory) and walks the column tiles J >= 2c: every B tile that comes in by TMA feeds two
// MMA groups, which halves the L2 -> shared-memory traffic per flop (a 128-row CTA needs ~12 TB/s of B tiles to keep
// the tensor pipe busy at K = 192).  CTAs are launched in order of decreasing work.
// The epilogue (warps 2-5: row tile 2c, warps 6-9: row tile 2c+1, one TMEM lane quarter each) classifies every
// accumulator element -- d2~ <= R - eps: neighbours, d2~ > R + eps: not, otherwise borderline -- and only EMITS pairs
// (i < j) into a global list (per-thread buffers of PAIR_BUF pairs, one atomic per flush).  Borderline pairs are
// decided afterwards by pair_resolve_kernel with the reference's own f64 expression, fully parallel, instead of
// stalling an epilogue warp (and, through it, the MMA pipeline) for ~1 us each; degrees, row offsets and the two
// directed entries of every pair follow from the list, and one radix sort puts everything in CSR order.  Nothing
// depends on the order in which tiles are visited.
__global__ void __launch_bounds__(NUM_THREADS, 1)
tensor_pairs_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const TensorParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: 2 A tiles | nsb B stages | barriers | tmem slot | per-warp column norms | per-thread pair buffers
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int atoms = P.atoms, nsb = P.nsb;
    const uint32_t tile_bytes = (uint32_t)atoms * ATOM_BYTES;
    uint8_t *sA = smem;
    uint8_t *sB = smem + 2 * tile_bytes;
    uint64_t *bars = (uint64_t *)(smem + (2 + nsb) * tile_bytes);
    uint64_t *a_full = bars, *full = bars + 2, *empty = full + NSB, *t_full = empty + NSB, *t_empty = t_full + NST;
    uint32_t *tmem_slot = (uint32_t *)(t_empty + NST);
    float *ncs = (float *)(bars + 20);                  // [EPI_WARPS][2][TT], 16-byte aligned
    int *pbuf = (int *)(ncs + EPI_WARPS * 2 * TT);      // [EPI_WARPS * 32][PAIR_BUF]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const ll n_tiles = P.n_tiles;
    const ll I0 = 2 * (ll)blockIdx.x;  // first row tile; the second one (I0 + 1) may lie past the end (all rows dead)

    if (warp == 0 && lane == 0) {
        mbar_init(a_full, 1);
        for (int s = 0; s < nsb; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < NST; ++s) { mbar_init(t_full + s, 1); mbar_init(t_empty + s, 16 * EPI_WARPS); }
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
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_expect_tx(a_full, 2 * tile_bytes);   // rows past the end of the operand are zero-filled
            for (int rt = 0; rt < 2; ++rt)
                for (int a = 0; a < atoms; ++a)
                    tma_load_2d(sA + rt * tile_bytes + a * ATOM_BYTES, &mapA, a * KATOM, (int)((I0 + rt) * TT), a_full);
            ll t = 0;
            TileIter it;
            it.init(P.skip, I0, n_tiles, P.words);
            for (ll j = it.next(); j >= 0; j = it.next(), ++t) {
                const int s = (int)(t % nsb);
                mbar_wait(empty + s, (uint32_t)((t / nsb) & 1) ^ 1);
                mbar_expect_tx(full + s, tile_bytes);
                for (int a = 0; a < atoms; ++a)
                    tma_load_2d(sB + s * tile_bytes + a * ATOM_BYTES, &mapB, a * KATOM, (int)(j * TT), full + s);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor: D f32, A/B bf16, both K-major, N = 128, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TT >> 3) << 17) | ((uint32_t)(TT >> 4) << 24);
            const uint32_t a_addr = smem_u32(sA);
            mbar_wait(a_full, 0);
            ll t = 0;
            TileIter it;
            it.init(P.skip, I0, n_tiles, P.words);
            for (ll j = it.next(); j >= 0; j = it.next(), ++t) {
                const int s = (int)(t % nsb);
                mbar_wait(full + s, (uint32_t)((t / nsb) & 1));          // TMA has landed this B tile
                const uint32_t b_addr = smem_u32(sB + s * tile_bytes);
                // the tile (I0 + 1, I0) below the diagonal is computed too (its pairs are dropped by the epilogue):
                // every accumulator stage is then used on every second step and the phases stay in lock step
                for (int rt = 0; rt < 2; ++rt) {
                    const int ts = (int)((t & 1) << 1) | rt;
                    mbar_wait(t_empty + ts, (uint32_t)((t >> 1) & 1) ^ 1);   // epilogue has drained this accumulator stage
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)ts * TT;
                    for (int k = 0; k < atoms * 4; ++k) {  // UMMA K = 16 bf16 = 32 bytes inside the swizzle atom
                        const uint32_t off = (uint32_t)(k >> 2) * ATOM_BYTES + (uint32_t)(k & 3) * 32;
                        tc_mma(d_tmem, umma_desc(a_addr + rt * tile_bytes + off), umma_desc(b_addr + off), idesc, k > 0);
                    }
                    tc_commit(t_full + ts);  // accumulator ready for the epilogue
                }
                tc_commit(empty + s);        // smem stage free once these MMAs have read it
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> classify -> emit pairs =====
        const int ew = warp - 2;             // 0 .. EPI_WARPS-1
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int rt = (ew >> 2) & 1;        // which of the CTA's two row tiles
        const int ch = ew >> 3;              // 16 epilogue warps: which 64 columns of the tile (0 with 8 warps)
        const int row = q * 32 + lane;       // accumulator row = TMEM lane
        float *ncw = ncs + ew * 2 * TT;
        int *mybuf = pbuf + (ew * 32 + lane) * PAIR_BUF;
        int n_buf = 0;
        const ll gi = (I0 + rt) * TT + row;
        const bool live = gi < P.n;
        const float nr = live ? P.norm[gi] : 0.f;
        // d2~ = nr + nc - 2 acc <= thr   <=>   w = nc - 2 acc <= thr - nr ; dead rows never pass
        const float thr_hi = live ? P.r2_hi - nr : -INFINITY, thr_lo = live ? P.r2_lo - nr : -INFINITY;
        auto flush = [&]() {
            const ll pos = (ll)atomicAdd(P.pair_count, (unsigned long long)n_buf);
            if (pos + n_buf <= P.pair_cap)
                for (int k = 0; k < n_buf; ++k) P.pairs[pos + k] = make_int2((int)gi, mybuf[k]);
            n_buf = 0;
        };
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
#pragma unroll
            for (int hp = EPI_WARPS == 16 ? ch : 0; hp < (EPI_WARPS == 16 ? ch + 1 : 2); ++hp) {  // 64 columns at a time: two tcgen05.ld in flight
                float va[32], vb[32];
                tmem_ld32(taddr + hp * 64, va);
                tmem_ld32(taddr + hp * 64 + 32, vb);
                tmem_ld_wait();
                if (EPI_WARPS == 16 || hp == 1) {  // this warp's part of the accumulator stage is in registers: hand it back
                    tc_fence_before();
                    mbar_arrive(t_empty + ts);
                }
#pragma unroll
                for (int chunk = 0; chunk < 2; ++chunk) {
                    const float *cur = chunk ? vb : va;
                    const int cbase = hp * 64 + chunk * 32;
                    // two bit masks per lane, element c at bit 31 - c: funnel-shift the sign of (thr - w) in
                    unsigned m_hi = 0, m_lo = 0;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float w = fmaf(-2.f, cur[c], nc[cbase + c]);
                        m_hi = __funnelshift_l(__float_as_uint(thr_hi - w), m_hi, 1);
                        m_lo = __funnelshift_l(__float_as_uint(thr_lo - w), m_lo, 1);
                    }
                    m_hi = ~m_hi;  // bit set <=> w <= thr (the difference is never NaN: norms are finite or +inf)
                    m_lo = ~m_lo;
                    if (!live) m_hi = 0;
                    const int col0 = (int)(j * TT) + cbase;
                    while (m_hi) {
                        const int c = __clz(m_hi);
                        const unsigned bit = 0x80000000u >> c;
                        m_hi &= ~bit;
                        const int gj = col0 + c;
                        if (gj <= (int)gi) continue;  // diagonal and sub-diagonal tiles: only j > i is emitted
                        mybuf[n_buf++] = (m_lo & bit) ? gj : (int)((unsigned)gj | BAND_FLAG);
                        if (n_buf == PAIR_BUF) flush();
                    }
                }
            }
            j = j_next;
        }
        if (n_buf) flush();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NST * TT));
    }
}

// borderline pairs are decided exactly (rejected ones get i = -1 and are counted: the degrees and row offsets follow
// from the sorted keys afterwards -- two atomics per pair on the degree array cost more than the whole exact test).
// Only ~1 pair in 6 of the list is borderline, so a thread per pair would run the 2 x d-element row walk with a few
// live lanes per warp and fetch every 32-byte sector for 8 of its bytes.  Instead the warp collects its borderline
// pairs by ballot and evaluates them four at a time, 8 lanes per pair: a lane loads 8 consecutive coordinates of both
// rows (whole sectors) and forms its 8 squared differences; the sum is then accumulated in the reference's order
// (_types.pyx:2253-2267: left to right, no FMA) by passing the running value from lane to lane of the group.
template <bool SQRT>
__global__ void __launch_bounds__(256)
pair_resolve_kernel(const double *__restrict__ x, int d, int2 *__restrict__ pairs, ll count, double rc,
                    unsigned long long *n_band, unsigned long long *n_rejected) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 3, gl = lane & 7;
    const ll k = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    int2 pr = make_int2(-1, -1);
    bool band = false;
    if (k < count) {
        pr = pairs[k];
        band = ((unsigned)pr.y & BAND_FLAG) != 0;
        pr.y = (int)((unsigned)pr.y & ~BAND_FLAG);
    }
    unsigned todo = __ballot_sync(full, band);
    if (lane == 0 && todo) atomicAdd(n_band, (unsigned long long)__popc(todo));
    bool ok = k < count;
    while (todo) {
        int src[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            src[q] = todo ? __ffs(todo) - 1 : -1;
            todo &= todo - 1;  // no-op at 0
        }
        const int mine = g == 0 ? src[0] : g == 1 ? src[1] : g == 2 ? src[2] : src[3];
        const int from = mine < 0 ? 0 : mine;
        const ll pi = __shfl_sync(full, pr.x, from), pj = __shfl_sync(full, pr.y, from);
        const double *a = x + pi * d, *b = x + pj * d;
        double sum = 0.0;
        for (int k0 = 0; k0 < d; k0 += 64) {
            double t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int kk = k0 + gl * 8 + u;
                double df = 0.0;
                if (mine >= 0 && kk < d) df = __dsub_rn(a[kk], b[kk]);
                t[u] = __dmul_rn(df, df);  // +0.0 past the end: adding it changes nothing
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (gl == r) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) sum = __dadd_rn(sum, t[u]);
                }
                sum = __shfl_sync(full, sum, (lane & ~7) | r);
            }
        }
        if (SQRT) sum = __dsqrt_rn(sum);
        const bool within = sum <= rc;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool wq = __shfl_sync(full, within, q * 8);
            if (lane == src[q]) ok = wq;
        }
    }
    if (band) pairs[k] = ok ? pr : make_int2(-1, -1);
    const unsigned rej = __ballot_sync(full, band && !ok);
    if (lane == 0 && rej) atomicAdd(n_rejected, (unsigned long long)__popc(rej));
}

// Both directed entries of every surviving pair and the self entry of every row as keys (row << bits | column) with
// 2^bits > n; rejected pairs get row n and sort behind everything.  After one radix sort over 2 * bits key bits (5
// passes instead of 7 for N = 2*10^5) the low parts of the first nnz keys ARE col_idx (rows ascending, columns
// ascending within a row).
__global__ void pair_keys_kernel(const int2 *__restrict__ pairs, ll count, ll n, int bits, unsigned long long *__restrict__ keys) {
    ll k = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (k < n) keys[2 * count + k] = ((unsigned long long)k << bits) | (unsigned long long)k;
    if (k >= count) return;
    const int2 pr = pairs[k];
    const bool ok = pr.x >= 0;
    const unsigned long long dead = (unsigned long long)n << bits;
    keys[2 * k] = ok ? ((unsigned long long)(unsigned)pr.x << bits) | (unsigned)pr.y : dead;
    keys[2 * k + 1] = ok ? ((unsigned long long)(unsigned)pr.y << bits) | (unsigned)pr.x : dead;
}

// low words of the sorted keys = col_idx; the first key of every row gives row_ptr (every row holds its self entry,
// so no row is empty)
__global__ void keys_to_cols_kernel(const unsigned long long *__restrict__ keys, ll nnz, ll n, int bits,
                                    int *__restrict__ col_idx, ll *__restrict__ row_ptr) {
    ll k = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const unsigned long long key = keys[k];
    col_idx[k] = (int)(unsigned)(key & ((1ull << bits) - 1ull));
    const unsigned row = (unsigned)(key >> bits);
    if (k == 0 || row != (unsigned)(keys[k - 1] >> bits)) row_ptr[row] = k;
    if (k == nnz - 1) row_ptr[n] = nnz;
}

__global__ void max_degree_kernel(const ll *__restrict__ row_ptr, ll n, int *out) {
    ll r = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    int dg = r < n ? (int)(row_ptr[r + 1] - row_ptr[r]) : 0;
    for (int o = 16; o; o >>= 1) dg = max(dg, __shfl_xor_sync(0xffffffffu, dg, o));
    if ((threadIdx.x & 31) == 0 && dg) atomicMax(out, dg);
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
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
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
        pivot_assign_kernel<<<ctx->sm_count * 8, 256, sizeof(float) * (size_t)n_piv * (size_t)d, st>>>(d_x, n, (int)d, n_piv, pk,
                                                                                                      ctx->b_idx.as<int>());
        size_t need = 0;
        CNN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, pk, pk2, ctx->b_idx.as<int>(), ctx->b_perm.as<int>(), (int)n, 0, 40, st));
        CNN_TRY(cnn_ensure(ctx, ctx->b_cub, need));
        size_t cap = ctx->b_cub.cap;
        CNN_CUDA(cub::DeviceRadixSort::SortPairs(ctx->b_cub.p, cap, pk, pk2, ctx->b_idx.as<int>(), ctx->b_perm.as<int>(), (int)n, 0, 40, st));
        gather_rows_kernel<<<cnn_blocks(n * d, 256), 256, 0, st>>>(d_x, ctx->b_perm.as<int>(), n, (int)d, ctx->b_xperm.as<double>());
        d_x = ctx->b_xperm.as<double>();
        // bounding balls of the tiles and the pairs of tiles that are too far apart to matter
        CNN_TRY(cnn_ensure(ctx, ctx->b_cell_start, sizeof(double) * (size_t)n_tiles * (size_t)(d + 1)));
        CNN_TRY(cnn_ensure(ctx, ctx->b_cell_count, sizeof(unsigned) * (size_t)n_tiles * (size_t)words));
        double *centre = ctx->b_cell_start.as<double>(), *radius = centre + n_tiles * d;
        tile_ball_kernel<<<(int)n_tiles, TT, 0, st>>>(d_x, n, (int)d, centre, radius);
        const double r_true = sq ? radius_cutoff : sqrt(radius_cutoff);
        tile_skip_kernel<<<cnn_blocks(n_tiles * words * 32, 256), 256, 0, st>>>(centre, radius, n_tiles, (int)d, words, r_true,
                                                                               ctx->b_cell_count.as<unsigned>());
        skip = ctx->b_cell_count.as<unsigned>();
        CNN_LAUNCHED(ctx, 4);
        CNN_KERNEL_CHECK();
    }

    // ---- operands -----------------------------------------------------------------------------
    CNN_TRY(cnn_ensure(ctx, ctx->b_xs, sizeof(__nv_bfloat16) * (size_t)n_pad * kt * 2));
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp1, sizeof(float) * (size_t)n_pad));
    CNN_TRY(cnn_ensure(ctx, ctx->b_bbox, sizeof(double) * (size_t)(d + 2)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_deg, sizeof(int) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_row_ptr, sizeof(ll) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_stats, sizeof(unsigned long long) * ST_N));
    __nv_bfloat16 *A = ctx->b_xs.as<__nv_bfloat16>(), *B = A + (size_t)n_pad * kt;
    float *norm = ctx->b_tmp1.as<float>();
    double *mean = ctx->b_bbox.as<double>();
    int *maxbits = (int *)(mean + d);
    {
        Span sp(ctx, T_GRID_BUILD);
        CNN_CUDA(cudaMemsetAsync(maxbits, 0, sizeof(int), st));
        CNN_CUDA(cudaMemsetAsync(ctx->b_stats.p, 0, sizeof(unsigned long long) * ST_N, st));
        CNN_CUDA(cudaMemsetAsync(mean, 0, sizeof(double) * (size_t)d, st));
        col_sum_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(d_x, n, (int)d, mean);
        col_mean_kernel<<<1, CM_MAXD, 0, st>>>(mean, (int)d, n);
        split_kernel<<<cnn_blocks(n_pad, 8) < ctx->sm_count * 32 ? cnn_blocks(n_pad, 8) : ctx->sm_count * 32, 256, 0, st>>>(d_x, mean, n, n_pad, (int)d, kt, A, B, norm, maxbits);
        CNN_LAUNCHED(ctx, 2);
        CNN_KERNEL_CHECK();
    }
    float *hm;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&hm));
    CNN_CUDA(cudaMemcpyAsync(hm, maxbits, sizeof(float), cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaStreamSynchronize(st));
    const double maxnorm = hm[0];
    if (!isfinite(maxnorm)) CNN_FAIL(CNNB200_E_INVALID, "neighbour search: coordinates contain NaN or infinity");
    // |d2~ - d2| <= 2 * maxnorm * (2^-15 [bf16 split: dropped lo.lo and residuals] + K * 2^-22 [fp32
    // accumulation in the tensor core]) + 2^-20 * maxnorm [norms rounded to f32, the f32 roundings of
    // nc - 2 acc and of t - nr in the epilogue]; see DESIGN.md
    const double eps = 2.0 * maxnorm * (ldexp(1.0, -15) + (double)kt * ldexp(1.0, -22)) + ldexp(maxnorm, -20) + 1e-30;

    CUtensorMap mapA, mapB;
    CNN_TRY(make_map(&mapA, A, n_pad, kt));
    CNN_TRY(make_map(&mapB, B, n_pad, kt));

    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp2, 128));
    unsigned long long *pair_count = (unsigned long long *)(ctx->b_tmp2.as<char>() + 48);
    unsigned long long *band_count = pair_count + 1;
    unsigned long long *rej_count = pair_count + 2;
    int *d_max_deg = (int *)(pair_count + 3);

    TensorParams P;
    memset(&P, 0, sizeof(P));
    P.norm = norm;
    P.n = n;
    P.n_tiles = n_tiles;
    P.atoms = atoms;
    P.nsb = atoms >= 4 ? 1 : 2;  // 2 A tiles + nsb B tiles must fit 227 KB
    P.r2_lo = (float)(R - eps);
    P.r2_hi = (float)(R + eps);
    // keep the f32 thresholds on the safe side of the f64 ones
    P.r2_lo = nextafterf(P.r2_lo, -INFINITY);
    P.r2_hi = nextafterf(P.r2_hi, INFINITY);
    P.pair_count = pair_count;
    P.skip = skip;
    P.words = words;

    const size_t smem = (size_t)(2 + P.nsb) * atoms * ATOM_BYTES + 1024 /*alignment*/ + 20 * 8 +
                        EPI_WARPS * 2 * TT * sizeof(float) + EPI_WARPS * 32 * PAIR_BUF * sizeof(int) + 64;
    CNN_CUDA(cudaFuncSetAttribute(tensor_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    // ---- the pass over the upper triangle (rerun once with a larger list should the pairs not fit) ------------
    ll n_pairs = 0;
    ll pair_cap = ctx->tensor_band_cap > 0 ? ctx->tensor_band_cap : (128 * n > (1ll << 20) ? 128 * n : (1ll << 20));
    unsigned long long *hb;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&hb));
    for (int attempt = 0; attempt < 2; ++attempt) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_worklist, sizeof(int2) * (size_t)pair_cap));
        P.pairs = ctx->b_worklist.as<int2>();
        P.pair_cap = pair_cap;
        CNN_CUDA(cudaMemsetAsync(pair_count, 0, 2 * sizeof(unsigned long long), st));
        {
            Span sp(ctx, T_SEARCH_COUNT);
            tensor_pairs_kernel<<<(int)((n_tiles + 1) / 2), NUM_THREADS, smem, st>>>(mapA, mapB, P);
            CNN_LAUNCHED(ctx, 1);
            CNN_KERNEL_CHECK();
        }
        CNN_CUDA(cudaMemcpyAsync(hb, pair_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CNN_CUDA(cudaStreamSynchronize(st));
        n_pairs = (ll)hb[0];
        if (n_pairs <= pair_cap) break;
        if (attempt == 1) CNN_FAIL(CNNB200_E_STATE, "tensor search: pair list overflowed twice");
        pair_cap = n_pairs + n_pairs / 16 + 1024;
    }
    ctx->tensor_band_cap = pair_cap;

    // ---- borderline pairs decided exactly; the survivors are counted, not yet attributed to rows ------------------------
    CNN_CUDA(cudaMemsetAsync(rej_count, 0, 2 * sizeof(unsigned long long), st));  // + d_max_deg
    {
        Span sp(ctx, T_SCAN);
        if (n_pairs > 0) {
            if (sq)
                pair_resolve_kernel<true><<<cnn_blocks(n_pairs, 256), 256, 0, st>>>(d_x, (int)d, P.pairs, n_pairs, radius_cutoff,
                                                                                band_count, rej_count);
            else
                pair_resolve_kernel<false><<<cnn_blocks(n_pairs, 256), 256, 0, st>>>(d_x, (int)d, P.pairs, n_pairs, radius_cutoff,
                                                                                 band_count, rej_count);
        }
        CNN_LAUNCHED(ctx, 1);
        CNN_KERNEL_CHECK();
    }
    CNN_CUDA(cudaMemcpyAsync(hb, band_count, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaStreamSynchronize(st));
    const ll n_band = (ll)hb[0], n_rej = (ll)hb[1];
    const ll nnz = 2 * (n_pairs - n_rej) + n;
    if (nnz >= (1ll << 31) - 1) CNN_FAIL(CNNB200_E_RANGE, "tensor search: more than 2^31 neighbour entries");
    // ---- (row : column) keys of both directions + self, one radix sort: low words = col_idx, row changes = row_ptr ------
    CNN_TRY(cnn_ensure(ctx, ctx->b_col_idx, sizeof(int) * (size_t)nnz));
    const ll n_keys = 2 * n_pairs + n;
    if (n_keys >= (1ll << 31) - 1) CNN_FAIL(CNNB200_E_RANGE, "tensor search: more than 2^31 sort keys");
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp0, sizeof(unsigned long long) * (size_t)n_keys * 2));
    unsigned long long *keys = ctx->b_tmp0.as<unsigned long long>(), *keys2 = keys + n_keys;
    {
        Span sp(ctx, T_SEARCH_FILL);
        const ll work = n_pairs > n ? n_pairs : n;
        int bits = 1;
        while ((1ll << bits) <= n) ++bits;  // 2^bits > n: row n (rejected pairs) fits as well
        pair_keys_kernel<<<cnn_blocks(work, 256), 256, 0, st>>>(P.pairs, n_pairs, n, bits, keys);
        size_t need = 0;
        CNN_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, need, keys, keys2, (int)n_keys, 0, 64, st));
        CNN_TRY(cnn_ensure(ctx, ctx->b_cub, need));
        size_t cap = ctx->b_cub.cap;
        CNN_CUDA(cub::DeviceRadixSort::SortKeys(ctx->b_cub.p, cap, keys, keys2, (int)n_keys, 0, 2 * bits, st));
        keys_to_cols_kernel<<<cnn_blocks(nnz, 256), 256, 0, st>>>(keys2, nnz, n, bits, ctx->b_col_idx.as<int>(),
                                                                  ctx->b_row_ptr.as<ll>());
        max_degree_kernel<<<cnn_blocks(n, 256), 256, 0, st>>>(ctx->b_row_ptr.as<ll>(), n, d_max_deg);
        CNN_LAUNCHED(ctx, 3);
        CNN_KERNEL_CHECK();
    }
    int *hmax = (int *)(hb + 4);
    CNN_CUDA(cudaMemcpyAsync(hmax, d_max_deg, sizeof(int), cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaStreamSynchronize(st));
    const ll max_deg = hmax[0];
    hb[0] = (unsigned long long)n_band;
    ctx->stats[ST_RECHECK] = (ll)hb[0];
    ctx->stats[ST_DIST] = n * (n - 1) / 2;
    ctx->n = n;
    ctx->nnz = nnz;
    ctx->max_deg = max_deg;
    ctx->row_ptr = ctx->b_row_ptr.as<ll>();
    ctx->col_idx = ctx->b_col_idx.as<int>();
    ctx->perm = reorder ? ctx->b_perm.as<int>() : nullptr;
    return CNNB200_OK;
}
