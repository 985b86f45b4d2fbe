// distmat.cu -- neighbourhoods from a precomputed distance matrix, and ascending order for rows that arrive unsorted.
//
// Reference: the "distances" recipe (recipes.py:329-339) runs NeighboursGetterExtBruteForce._get (_types.pyx:1692-1710)
// with MetricExtPrecomputed (_types.pyx:2108-2127): point j is a neighbour of i when matrix[i, j] <= radius_cutoff
// (the radius is not adjusted; a NaN never passes; the diagonal decides whether a point counts itself).
//
// The matrix is read exactly once (8*N^2 bytes, the compulsory traffic): dist_mask_kernel turns every row into a bit
// mask (N^2/8 bytes, 1/64 of the input) and its popcount, a scan gives row_ptr, dist_fill_kernel expands the masks
// into ascending int32 rows.  A host matrix is streamed through a staging buffer in row chunks.
#include <cub/cub.cuh>

#include "common.cuh"

#define DM_COLS 1024  // columns per warp item: 32 mask words

// one warp per (row, 1024-column block): 32 coalesced 256-byte loads, one ballot each
__global__ void __launch_bounds__(256)
dist_mask_kernel(const double *__restrict__ dist, ll row0, ll rows, ll n, ll words, double rc,
                 unsigned *__restrict__ mask, int *__restrict__ deg) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const ll blocks_per_row = (n + DM_COLS - 1) / DM_COLS;
    const ll items = rows * blocks_per_row;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll it = warp; it < items; it += nwarps) {
        const ll r = it / blocks_per_row, blk = it - r * blocks_per_row;
        const double *src = dist + r * n;
        const ll c0 = blk * DM_COLS + lane;
        unsigned word = 0;
#pragma unroll
        for (int k0 = 0; k0 < 32; k0 += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const ll c = c0 + (ll)(k0 + u) * 32;
                v[u] = c < n ? __ldcs(src + c) : __longlong_as_double(0x7ff8000000000000ll);  // NaN: never a neighbour
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned b = __ballot_sync(full, v[u] <= rc);
                if (lane == k0 + u) word = b;
            }
        }
        const ll w = blk * 32 + lane;
        if (w < words) mask[(row0 + r) * words + w] = word;
        const int c = __reduce_add_sync(full, __popc(word));
        if (lane == 0 && c) atomicAdd(deg + row0 + r, c);
    }
}

// warp per row: the set bits of the row's mask words become ascending column indices
__global__ void __launch_bounds__(256)
dist_fill_kernel(const unsigned *__restrict__ mask, ll n, ll words, const ll *__restrict__ row_ptr,
                 int *__restrict__ col_idx) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll r = warp; r < n; r += nwarps) {
        const unsigned *m = mask + r * words;
        ll base = row_ptr[r];
        for (ll w0 = 0; w0 < words; w0 += 32) {
            const ll w = w0 + lane;
            unsigned word = w < words ? m[w] : 0u;
            const int c = __popc(word);
            int incl = c;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const int t = __shfl_up_sync(full, incl, s);
                if (lane >= s) incl += t;
            }
            ll pos = base + incl - c;
            while (word) {
                const int b = __ffs(word) - 1;
                word &= word - 1;
                col_idx[pos++] = (int)(w * 32 + b);
            }
            base += __shfl_sync(full, incl, 31);
        }
    }
}

// degrees -> offsets, masks -> ascending rows, adoption by the context
static int rows_from_masks(cnnb200_ctx *ctx, const unsigned *mask, int *deg, ll n, ll words) {
    cudaStream_t st = ctx->stream;
    ll nnz = 0, md = 0;
    {
        Span sp(ctx, T_SCAN);
        CNN_TRY(cnn_scan_deg_to_row_ptr(ctx, deg, n, ctx->b_row_ptr.as<ll>(), &nnz, &md));
    }
    CNN_TRY(cnn_ensure(ctx, ctx->b_col_idx, sizeof(int) * (size_t)(nnz > 0 ? nnz : 1)));
    {
        Span sp(ctx, T_SEARCH_FILL);
        ll blocks = (n + 7) / 8;
        if (blocks > (ll)ctx->sm_count * 32) blocks = (ll)ctx->sm_count * 32;
        dist_fill_kernel<<<(int)blocks, 256, 0, st>>>(mask, n, words, ctx->b_row_ptr.as<ll>(), ctx->b_col_idx.as<int>());
        CNN_LAUNCHED(ctx, 1);
    }
    CNN_KERNEL_CHECK();
    ctx->stats[ST_RECHECK] = 0;
    return cnn_finish_csr(ctx, n, nnz, md, ctx->b_row_ptr.as<ll>(), ctx->b_col_idx.as<int>());
}

extern "C" int cnnb200_neighbours_from_distances(cnnb200_ctx *ctx, const double *dist, int loc, int64_t n,
                                                 double radius_cutoff) {
    if (!ctx || !dist) CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_distances: NULL argument");
    if (n < 1) CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_distances: need n >= 1");
    if (n >= (1ll << 31) - 1) CNN_FAIL(CNNB200_E_RANGE, "neighbours_from_distances: n must be < 2^31-1");
    if (radius_cutoff != radius_cutoff) CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_distances: radius_cutoff is NaN");
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->have_nb = false;
    ctx->have_points = false;
    ctx->have_other = false;
    ctx->have_comp = false;
    const ll words = (n + 31) / 32;
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp1, sizeof(unsigned) * (size_t)(n * words)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_deg, sizeof(int) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_row_ptr, sizeof(ll) * (size_t)(n + 1)));
    CNN_CUDA(cudaMemsetAsync(ctx->b_deg.p, 0, sizeof(int) * (size_t)(n + 1), st));
    unsigned *mask = ctx->b_tmp1.as<unsigned>();
    int *deg = ctx->b_deg.as<int>();
    const ll blocks_per_row = (n + DM_COLS - 1) / DM_COLS;
    auto launch = [&](const double *d_rows, ll row0, ll rows) {
        ll blocks = (rows * blocks_per_row + 7) / 8;
        if (blocks > (ll)ctx->sm_count * 32) blocks = (ll)ctx->sm_count * 32;
        if (blocks < 1) blocks = 1;
        dist_mask_kernel<<<(int)blocks, 256, 0, st>>>(d_rows, row0, rows, n, words, radius_cutoff, mask, deg);
        CNN_LAUNCHED(ctx, 1);
    };
    if (loc == CNNB200_DEVICE) {
        Span sp(ctx, T_SEARCH_COUNT);
        launch(dist, 0, n);
    } else {
        // row chunks of at most 256 MB through two staging halves: the copy of chunk k+1 is queued behind the
        // kernel of chunk k on the same stream (the path is bound by the host link: 8*N^2 bytes over PCIe)
        ll rows_per_chunk = (256ll << 20) / (8 * n);
        if (rows_per_chunk < 1) rows_per_chunk = 1;
        if (rows_per_chunk > n) rows_per_chunk = n;
        CNN_TRY(cnn_ensure(ctx, ctx->b_tmp0, sizeof(double) * (size_t)(rows_per_chunk * n)));
        for (ll r0 = 0; r0 < n; r0 += rows_per_chunk) {
            const ll rows = n - r0 < rows_per_chunk ? n - r0 : rows_per_chunk;
            {
                Span sp(ctx, T_H2D);
                CNN_CUDA(cudaMemcpyAsync(ctx->b_tmp0.p, dist + r0 * n, sizeof(double) * (size_t)(rows * n),
                                         cudaMemcpyHostToDevice, st));
            }
            Span sp(ctx, T_SEARCH_COUNT);
            launch(ctx->b_tmp0.as<double>(), r0, rows);
        }
    }
    CNN_KERNEL_CHECK();
    ctx->stats[ST_DIST] = n * n;
    return rows_from_masks(ctx, mask, deg, n, words);
}

// ---- condensed distances ----------------------------------------------------------------------------------------
// InputDataExtDistancesLinearMemoryview (_types.pyx:1112-1150) stores the N (N - 1) / 2 distances d(a, b), a < b, row
// by row: d(a, b) sits at a (N - 1) - a (a - 1) / 2 + (b - a - 1); d(a, a) = 0.  The upper triangle of the bit matrix
// follows from ONE coalesced pass over the vector (4 N^2 bytes, the compulsory traffic); the lower triangle is its
// transpose, formed on the bits (N^2 / 8 bytes) by 32 x 32 block transposes.

// one warp per (row, 1024-column block) at or right of the diagonal
__global__ void __launch_bounds__(256)
cond_mask_kernel(const double *__restrict__ cond, ll n, ll words, double rc, unsigned *__restrict__ mask) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const ll blocks_per_row = (n + DM_COLS - 1) / DM_COLS;
    const ll items = n * blocks_per_row;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    for (ll it = warp; it < items; it += nwarps) {
        const ll r = it / blocks_per_row, blk = it - r * blocks_per_row;
        const ll w = blk * 32 + lane;
        if ((blk + 1) * DM_COLS <= r) {  // wholly left of the diagonal: filled by the transposition
            if (w < words) mask[r * words + w] = 0u;
            continue;
        }
        const double *src = cond + (r * (n - 1) - r * (r - 1) / 2) - r - 1;  // src[c] = d(r, c) for c > r
        const ll c0 = blk * DM_COLS + lane;
        unsigned word = 0;
#pragma unroll
        for (int k0 = 0; k0 < 32; k0 += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const ll c = c0 + (ll)(k0 + u) * 32;
                v[u] = c >= n || c < r ? nan : (c == r ? 0.0 : __ldcs(src + c));  // NaN: never a neighbour
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned b = __ballot_sync(full, v[u] <= rc);
                if (lane == k0 + u) word = b;
            }
        }
        if (w < words) mask[r * words + w] = word;
    }
}

// lane l holds row l of a 32 x 32 bit block (bit c = column c); afterwards lane c holds column c (bit l = row l)
__device__ __forceinline__ unsigned bit_transpose32(unsigned x, int lane) {
#pragma unroll
    for (int k = 16; k; k >>= 1) {
        const unsigned m = k == 16 ? 0x0000ffffu : k == 8 ? 0x00ff00ffu : k == 4 ? 0x0f0f0f0fu : k == 2 ? 0x33333333u : 0x55555555u;
        const unsigned y = __shfl_xor_sync(0xffffffffu, x, k);
        x = (lane & k) ? ((x & ~m) | ((y & ~m) >> k)) : ((x & m) | ((y & m) << k));
    }
    return x;
}

// one warp per 32 x 32 bit block (I, J), J >= I, of the upper triangle: its transpose becomes block (J, I)
__global__ void __launch_bounds__(256)
mask_mirror_kernel(unsigned *__restrict__ mask, ll n, ll words) {
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll it = warp; it < words * words; it += nwarps) {
        const ll I = it / words, J = it - I * words;
        if (J < I) continue;
        const ll r = 32 * I + lane;
        const unsigned x = r < n ? mask[r * words + J] : 0u;
        const unsigned t = bit_transpose32(x, lane);
        if (J == I) {
            if (r < n) mask[r * words + I] = x | t;
        } else {
            const ll rt = 32 * J + lane;
            if (rt < n) mask[rt * words + I] = t;
        }
    }
}

__global__ void __launch_bounds__(256)
mask_degree_kernel(const unsigned *__restrict__ mask, ll n, ll words, int *__restrict__ deg) {
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll r = warp; r < n; r += nwarps) {
        int c = 0;
        for (ll w = lane; w < words; w += 32) c += __popc(mask[r * words + w]);
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) deg[r] = c;
    }
}

extern "C" int cnnb200_neighbours_from_condensed(cnnb200_ctx *ctx, const double *dist, int loc, int64_t n,
                                                 double radius_cutoff) {
    if (!ctx || (!dist && n > 1)) CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_condensed: NULL argument");
    if (n < 1) CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_condensed: need n >= 1");
    if (n >= (1ll << 31) - 1) CNN_FAIL(CNNB200_E_RANGE, "neighbours_from_condensed: n must be < 2^31-1");
    if (radius_cutoff != radius_cutoff) CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_condensed: radius_cutoff is NaN");
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->have_nb = false;
    ctx->have_points = false;
    ctx->have_other = false;
    ctx->have_comp = false;
    const ll words = (n + 31) / 32, len = n * (n - 1) / 2;
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp1, sizeof(unsigned) * (size_t)(n * words)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_deg, sizeof(int) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_row_ptr, sizeof(ll) * (size_t)(n + 1)));
    CNN_CUDA(cudaMemsetAsync(ctx->b_deg.p, 0, sizeof(int) * (size_t)(n + 1), st));
    unsigned *mask = ctx->b_tmp1.as<unsigned>();
    int *deg = ctx->b_deg.as<int>();
    const double *d_cond = dist;
    if (loc != CNNB200_DEVICE && len > 0) {
        Span sp(ctx, T_H2D);
        CNN_TRY(cnn_ensure(ctx, ctx->b_tmp0, sizeof(double) * (size_t)len));
        CNN_CUDA(cudaMemcpyAsync(ctx->b_tmp0.p, dist, sizeof(double) * (size_t)len, cudaMemcpyHostToDevice, st));
        d_cond = ctx->b_tmp0.as<double>();
    }
    {
        Span sp(ctx, T_SEARCH_COUNT);
        const ll blocks_per_row = (n + DM_COLS - 1) / DM_COLS;
        ll blocks = (n * blocks_per_row + 7) / 8;
        if (blocks > (ll)ctx->sm_count * 32) blocks = (ll)ctx->sm_count * 32;
        cond_mask_kernel<<<(int)blocks, 256, 0, st>>>(d_cond, n, words, radius_cutoff, mask);
        ll mb = (words * words + 7) / 8;
        if (mb > (ll)ctx->sm_count * 32) mb = (ll)ctx->sm_count * 32;
        mask_mirror_kernel<<<(int)mb, 256, 0, st>>>(mask, n, words);
        ll db = (n + 7) / 8;
        if (db > (ll)ctx->sm_count * 32) db = (ll)ctx->sm_count * 32;
        mask_degree_kernel<<<(int)db, 256, 0, st>>>(mask, n, words, deg);
        CNN_LAUNCHED(ctx, 3);
    }
    CNN_KERNEL_CHECK();
    ctx->stats[ST_DIST] = len;
    return rows_from_masks(ctx, mask, deg, n, words);
}

// ---- ascending rows ------------------------------------------------------------------------------------------------
// The reference's "neighbourhoods" recipe (recipes.py:340-348) accepts neighbour lists in any order (its containers
// are hash sets); every kernel here works on ascending rows.  The labels do not depend on the order inside a list
// (a cluster is a connected component, numbered by its smallest member), so unsorted input is sorted once on arrival.
__global__ void rows_unsorted_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col_idx, ll n, int *flag) {
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    bool bad = false;
    for (ll r = warp; r < n; r += nwarps) {
        const ll b = row_ptr[r], e = row_ptr[r + 1];
        for (ll k = b + lane; k + 1 < e; k += 32) bad |= col_idx[k] > col_idx[k + 1];
    }
    if (bad) *flag = 1;
}

__global__ void row_keys_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col_idx, ll n,
                                unsigned long long *__restrict__ keys) {
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll r = warp; r < n; r += nwarps) {
        const ll b = row_ptr[r], e = row_ptr[r + 1];
        for (ll k = b + lane; k < e; k += 32) keys[k] = ((unsigned long long)r << 32) | (unsigned)col_idx[k];
    }
}

__global__ void keys_low_kernel(const unsigned long long *__restrict__ keys, ll nnz, int *__restrict__ col_idx) {
    ll k = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (k < nnz) col_idx[k] = (int)(unsigned)(keys[k] & 0xffffffffull);
}

// Rows of a CSR (device: row_ptr [n+1], col [nnz]) ascending: one radix sort of (row : column) keys, the low words of
// the sorted keys are the ascending rows.  Writes into `out` (grown as needed; may not alias `col`).
int cnn_sort_csr_rows(cnnb200_ctx *ctx, const ll *row_ptr, const int *col, ll n, ll nnz, Buf &out) {
    cudaStream_t st = ctx->stream;
    if (nnz >= (1ll << 31) - 1)
        CNN_FAIL(CNNB200_E_RANGE, "row sort: more than 2^31-2 entries; pass ascending rows");
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp0, sizeof(unsigned long long) * (size_t)nnz));
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp1, sizeof(unsigned long long) * (size_t)nnz));
    unsigned long long *keys = ctx->b_tmp0.as<unsigned long long>(), *keys2 = ctx->b_tmp1.as<unsigned long long>();
    ll blocks = (n + 7) / 8;
    if (blocks > (ll)ctx->sm_count * 32) blocks = (ll)ctx->sm_count * 32;
    row_keys_kernel<<<(int)blocks, 256, 0, st>>>(row_ptr, col, n, keys);
    int bits = 1;
    while ((1ll << bits) < n) ++bits;
    size_t need = 0;
    CNN_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, need, keys, keys2, (int)nnz, 0, 32 + bits, st));
    CNN_TRY(cnn_ensure(ctx, ctx->b_cub, need));
    size_t cap = ctx->b_cub.cap;
    CNN_CUDA(cub::DeviceRadixSort::SortKeys(ctx->b_cub.p, cap, keys, keys2, (int)nnz, 0, 32 + bits, st));
    CNN_TRY(cnn_ensure(ctx, out, sizeof(int) * (size_t)nnz));
    keys_low_kernel<<<cnn_blocks(nnz, 256), 256, 0, st>>>(keys2, nnz, out.as<int>());
    CNN_LAUNCHED(ctx, 2);
    CNN_KERNEL_CHECK();
    return CNNB200_OK;
}

// 1 in *unsorted when some row of the CSR is not ascending (synchronises)
int cnn_csr_rows_unsorted(cnnb200_ctx *ctx, const ll *row_ptr, const int *col, ll n, int *unsorted) {
    cudaStream_t st = ctx->stream;
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp2, 64));
    int *flag = ctx->b_tmp2.as<int>();
    CNN_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
    ll blocks = (n + 7) / 8;
    if (blocks > (ll)ctx->sm_count * 32) blocks = (ll)ctx->sm_count * 32;
    rows_unsorted_kernel<<<(int)blocks, 256, 0, st>>>(row_ptr, col, n, flag);
    CNN_LAUNCHED(ctx, 1);
    CNN_KERNEL_CHECK();
    int *h;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&h));
    CNN_CUDA(cudaMemcpyAsync(h, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaStreamSynchronize(st));
    *unsorted = h[0];
    return CNNB200_OK;
}

// The intersection kernels (cn.cu: shared-memory row window, lower_bound of `above`) REQUIRE ascending rows; every
// ingest path ends here, so an unsorted list costs a sort instead of an out-of-bounds shared-memory write.
int cnn_ensure_sorted_rows(cnnb200_ctx *ctx, bool known_unsorted, int *sorted_out) {
    const ll n = ctx->n, nnz = ctx->nnz;
    if (sorted_out) *sorted_out = 0;
    if (nnz < 2) return CNNB200_OK;
    int unsorted = known_unsorted ? 1 : 0;
    if (!known_unsorted) CNN_TRY(cnn_csr_rows_unsorted(ctx, ctx->row_ptr, ctx->col_idx, n, &unsorted));
    if (!unsorted) return CNNB200_OK;
    if (ctx->col_idx == ctx->b_col_idx.as<int>()) {
        // sort through the spare buffer and swap (the sort may not run in place)
        CNN_TRY(cnn_sort_csr_rows(ctx, ctx->row_ptr, ctx->col_idx, n, nnz, ctx->b_col_spare));
        Buf t = ctx->b_col_idx; ctx->b_col_idx = ctx->b_col_spare; ctx->b_col_spare = t;
    } else {
        CNN_TRY(cnn_sort_csr_rows(ctx, ctx->row_ptr, ctx->col_idx, n, nnz, ctx->b_col_idx));  // borrowed memory stays untouched
    }
    ctx->col_idx = ctx->b_col_idx.as<int>();
    ++ctx->nb_version;
    ctx->have_comp = false;
    if (sorted_out) *sorted_out = 1;
    return CNNB200_OK;
}

extern "C" int cnnb200_neighbours_sort_rows(cnnb200_ctx *ctx, int *sorted_out) {
    if (!ctx) CNN_FAIL(CNNB200_E_INVALID, "ctx is NULL");
    if (!ctx->have_nb) CNN_FAIL(CNNB200_E_STATE, "neighbours_sort_rows: no neighbourhoods held");
    CNN_CUDA(cudaSetDevice(ctx->device));
    Span sp(ctx, T_INGEST);
    return cnn_ensure_sorted_rows(ctx, false, sorted_out);
}
