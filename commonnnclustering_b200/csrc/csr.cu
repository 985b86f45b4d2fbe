// csr.cu -- K4: ingest of precomputed neighbourhoods (CSR or the reference's zero-padded int64
// matrix) and export of the held neighbourhoods in the caller's index space.
//
// Bulk replacement of NeighboursGetterExtLookup._get (reference _types.pyx:1806-1817) over
// InputDataExtNeighbourhoodsMemoryview (_types.pyx:1153-1217): instead of copying one padded
// row into a scratch container per visit, all rows are converted once to int32 CSR in HBM.
#include <cub/cub.cuh>

#include "common.cuh"

__global__ void widen_kernel(const int *__restrict__ in, ll *__restrict__ out, ll m) {
    ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    ll stride = (ll)gridDim.x * blockDim.x;
    for (; i < m; i += stride) out[i] = in[i];
}

// narrow int64 -> int32 with a range check (flag[0] set when an index is outside [0, n))
__global__ void narrow_kernel(const ll *__restrict__ in, int *__restrict__ out, ll m, ll n, int *flag) {
    ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    ll stride = (ll)gridDim.x * blockDim.x;
    bool bad = false;
    for (; i < m; i += stride) {
        ll v = in[i];
        bad |= (v < 0 || v >= n);
        out[i] = (int)v;
    }
    if (bad) *flag = 1;
}

__global__ void range_check_kernel(const int *__restrict__ in, ll m, ll n, int *flag) {
    ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    ll stride = (ll)gridDim.x * blockDim.x;
    bool bad = false;
    for (; i < m; i += stride) {
        int v = in[i];
        bad |= (v < 0 || v >= n);
    }
    if (bad) *flag = 1;
}

// warp per row: indices inside [0, n) (flag[0]) and rows ascending (flag[2]); one read of col_idx for both.  Rows whose
// bounds are not monotone are skipped here (deg_from_row_ptr_kernel reports them).
__global__ void rows_check_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col, ll n, ll nnz, int *flag) {
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    bool bad = false, unsorted = false;
    for (ll r = warp; r < n; r += nwarps) {
        const ll b = row_ptr[r], e = row_ptr[r + 1];
        if (b < 0 || e > nnz || b > e) continue;
        for (ll k = b + lane; k < e; k += 32) {
            const int v = col[k];
            bad |= v < 0 || v >= n;
            if (k + 1 < e) unsorted |= v > col[k + 1];
        }
    }
    if (bad) flag[0] = 1;
    if (unsorted) flag[2] = 1;
}

// The same checks with one THREAD per row, for short rows (a warp per row spends most of its time waiting for two
// dependent loads per 50 entries): offsets monotone (flag[0]), indices inside [0, n) (flag[0]), rows ascending (flag[2]),
// largest degree (flag[1]).  Consecutive threads walk consecutive rows, so the sectors they touch are shared through L1.
__global__ void __launch_bounds__(256)
rows_check_short_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col, ll n, ll nnz, int *flag) {
    const ll r = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    bool bad = false, unsorted = false;
    int deg = 0;
    if (r < n) {
        const ll b = row_ptr[r], e = row_ptr[r + 1];
        if (b < 0 || e > nnz || b > e || e - b > 0x7fffffffll) {
            bad = true;
        } else {
            deg = (int)(e - b);
            int prev = -1;
            for (ll k = b; k < e; ++k) {
                const int v = col[k];
                bad |= v < 0 || v >= n;
                unsorted |= v < prev;
                prev = v;
            }
        }
    }
    deg = __reduce_max_sync(0xffffffffu, deg);
    if ((threadIdx.x & 31) == 0 && deg > 0) atomicMax(flag + 1, deg);
    if (bad) flag[0] = 1;
    if (unsorted) flag[2] = 1;
}

__global__ void deg_from_row_ptr_kernel(const ll *__restrict__ row_ptr, ll n, int *deg, int *flag) {
    ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { deg[n] = 0; return; }
    ll dlt = row_ptr[i + 1] - row_ptr[i];
    if (dlt < 0 || dlt > 0x7fffffffll) { *flag = 1; dlt = 0; }
    deg[i] = (int)dlt;
}

__global__ void counts_to_deg_kernel(const ll *__restrict__ counts, ll n, ll max_deg, int *deg, int *flag) {
    ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { deg[n] = 0; return; }
    ll c = counts[i];
    if (c < 0 || c > max_deg) { *flag = 1; c = 0; }
    deg[i] = (int)c;
}

// one warp per row: copy the valid prefix of a padded int64 row into its CSR slot as int32
__global__ void padded_rows_kernel(const ll *__restrict__ padded, ll row0, ll rows, ll max_deg, ll n,
                                   const ll *__restrict__ row_ptr, int *__restrict__ col_idx, int *flag) {
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    bool bad = false;
    for (ll r = warp; r < rows; r += nwarps) {
        ll b = row_ptr[row0 + r];
        int m = (int)(row_ptr[row0 + r + 1] - b);
        const ll *src = padded + r * max_deg;
        for (int k = lane; k < m; k += 32) {
            ll v = src[k];
            bad |= (v < 0 || v >= n);
            col_idx[b + k] = (int)v;
        }
    }
    if (bad) *flag = 1;
}

int cnn_finish_csr(cnnb200_ctx *ctx, ll n, ll nnz, ll max_deg, const ll *row_ptr, const int *col_idx) {
    ctx->n = n;
    ctx->nnz = nnz;
    ctx->max_deg = max_deg;
    ctx->row_ptr = row_ptr;
    ctx->col_idx = col_idx;
    ctx->perm = nullptr;
    ctx->have_nb = true;
    ++ctx->nb_version;
    ctx->have_comp = false;
    return CNNB200_OK;
}

static int read_flag(cnnb200_ctx *ctx, int *d_flag, int *out) {
    int *h;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&h));
    CNN_CUDA(cudaMemcpyAsync(h, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CNN_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = *h;
    return CNNB200_OK;
}

extern "C" int cnnb200_neighbours_from_csr(cnnb200_ctx *ctx, const int64_t *row_ptr, const void *col_idx,
                                           int idx_bytes, int loc, int borrow, int64_t n, int64_t nnz) {
    if (!ctx || !row_ptr || (!col_idx && nnz > 0)) CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_csr: NULL argument");
    if (n < 1 || nnz < 0) CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_csr: need n >= 1, nnz >= 0");
    if (n >= (1ll << 31) - 1) CNN_FAIL(CNNB200_E_RANGE, "neighbours_from_csr: n must be < 2^31-1");
    if (idx_bytes != 4 && idx_bytes != 8) CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_csr: idx_bytes must be 4 or 8");
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->have_nb = false;
    ctx->have_points = false;
    ctx->have_other = false;
    ctx->have_comp = false;
    cudaMemcpyKind kind = loc == CNNB200_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    Span sp(ctx, T_INGEST);
    if (loc == CNNB200_HOST && (row_ptr[0] != 0 || row_ptr[n] != nnz))
        CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_csr: row_ptr[0] must be 0 and row_ptr[n] == nnz");

    const ll *d_row_ptr;
    const int *d_col;
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp2, 64));
    int *flag = ctx->b_tmp2.as<int>();
    CNN_CUDA(cudaMemsetAsync(flag, 0, 4 * sizeof(int), st));
    if (loc == CNNB200_DEVICE && borrow && idx_bytes == 4) {
        d_row_ptr = (const ll *)row_ptr;
        d_col = (const int *)col_idx;
    } else {
        CNN_TRY(cnn_ensure(ctx, ctx->b_row_ptr, sizeof(ll) * (size_t)(n + 1)));
        CNN_TRY(cnn_ensure(ctx, ctx->b_col_idx, sizeof(int) * (size_t)nnz));
        CNN_CUDA(cudaMemcpyAsync(ctx->b_row_ptr.p, row_ptr, sizeof(ll) * (size_t)(n + 1), kind, st));
        if (idx_bytes == 4) {
            CNN_CUDA(cudaMemcpyAsync(ctx->b_col_idx.p, col_idx, sizeof(int) * (size_t)nnz, kind, st));
        } else {
            // stage int64 indices in chunks and narrow them on the device
            const ll chunk = 32ll << 20;  // 32 Mi entries = 256 MB
            const ll *src = (const ll *)col_idx;
            if (loc == CNNB200_DEVICE) {
                narrow_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(src, ctx->b_col_idx.as<int>(), nnz, n, flag);
            } else {
                CNN_TRY(cnn_ensure(ctx, ctx->b_tmp0, sizeof(ll) * (size_t)(nnz < chunk ? nnz : chunk)));
                for (ll o = 0; o < nnz; o += chunk) {
                    ll m = nnz - o < chunk ? nnz - o : chunk;
                    CNN_CUDA(cudaMemcpyAsync(ctx->b_tmp0.p, src + o, sizeof(ll) * (size_t)m, kind, st));
                    narrow_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->b_tmp0.as<ll>(),
                                                                     ctx->b_col_idx.as<int>() + o, m, n, flag);
                }
            }
        }
        d_row_ptr = ctx->b_row_ptr.as<ll>();
        d_col = ctx->b_col_idx.as<int>();
    }
    CNN_KERNEL_CHECK();
    if (nnz / n < 96) {
        // short rows: offsets, index range, ascending order and the largest degree in ONE pass, a thread per row
        rows_check_short_kernel<<<cnn_blocks(n, 256), 256, 0, st>>>(d_row_ptr, d_col, n, nnz, flag);
        CNN_LAUNCHED(ctx, 1);
        CNN_KERNEL_CHECK();
    } else {
        // degrees (for max_deg and a monotonicity check of row_ptr), then index range + ascending rows, a warp per row
        CNN_TRY(cnn_ensure(ctx, ctx->b_deg, sizeof(int) * (size_t)(n + 1)));
        deg_from_row_ptr_kernel<<<cnn_blocks(n + 1, 256), 256, 0, st>>>(d_row_ptr, n, ctx->b_deg.as<int>(), flag);
        {
            ll blocks = (n + 7) / 8;
            if (blocks > (ll)ctx->sm_count * 32) blocks = (ll)ctx->sm_count * 32;
            rows_check_kernel<<<(int)blocks, 256, 0, st>>>(d_row_ptr, d_col, n, nnz, flag);
        }
        CNN_KERNEL_CHECK();
        size_t need = 0;
        CNN_CUDA(cub::DeviceReduce::Max(nullptr, need, ctx->b_deg.as<int>(), flag + 1, (int)(n + 1), st));
        CNN_TRY(cnn_ensure(ctx, ctx->b_cub, need));
        size_t cap = ctx->b_cub.cap;
        CNN_CUDA(cub::DeviceReduce::Max(ctx->b_cub.p, cap, ctx->b_deg.as<int>(), flag + 1, (int)(n + 1), st));
    }
    int *h;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&h));
    CNN_CUDA(cudaMemcpyAsync(h, flag, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaStreamSynchronize(st));
    if (h[0]) CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_csr: row_ptr not monotone or an index outside [0, n)");
    const bool unsorted = h[2] != 0;
    CNN_TRY(cnn_finish_csr(ctx, n, nnz, h[1], d_row_ptr, d_col));
    if (unsorted) CNN_TRY(cnn_ensure_sorted_rows(ctx, true, nullptr));  // lists in any order: sorted on arrival
    return CNNB200_OK;
}

extern "C" int cnnb200_neighbours_from_padded(cnnb200_ctx *ctx, const int64_t *padded, const int64_t *counts,
                                              int loc, int64_t n, int64_t max_deg) {
    if (!ctx || !counts || (!padded && max_deg > 0)) CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_padded: NULL argument");
    if (n < 1 || max_deg < 0) CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_padded: need n >= 1, max_deg >= 0");
    if (n >= (1ll << 31) - 1) CNN_FAIL(CNNB200_E_RANGE, "neighbours_from_padded: n must be < 2^31-1");
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->have_nb = false;
    ctx->have_points = false;
    ctx->have_other = false;
    ctx->have_comp = false;
    cudaMemcpyKind kind = loc == CNNB200_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    Span sp(ctx, T_INGEST);
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp2, 64));
    int *flag = ctx->b_tmp2.as<int>();
    CNN_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
    CNN_TRY(cnn_ensure(ctx, ctx->b_deg, sizeof(int) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_row_ptr, sizeof(ll) * (size_t)(n + 1)));
    const ll *d_counts = (const ll *)counts;
    if (loc == CNNB200_HOST) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_tmp1, sizeof(ll) * (size_t)n));
        CNN_CUDA(cudaMemcpyAsync(ctx->b_tmp1.p, counts, sizeof(ll) * (size_t)n, kind, st));
        d_counts = ctx->b_tmp1.as<ll>();
    }
    counts_to_deg_kernel<<<cnn_blocks(n + 1, 256), 256, 0, st>>>(d_counts, n, max_deg, ctx->b_deg.as<int>(), flag);
    CNN_KERNEL_CHECK();
    ll nnz = 0, md = 0;
    CNN_TRY(cnn_scan_deg_to_row_ptr(ctx, ctx->b_deg.as<int>(), n, ctx->b_row_ptr.as<ll>(), &nnz, &md));
    int bad = 0;
    CNN_TRY(read_flag(ctx, flag, &bad));
    if (bad) CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_padded: a count is negative or larger than max_deg");
    CNN_TRY(cnn_ensure(ctx, ctx->b_col_idx, sizeof(int) * (size_t)nnz));
    if (max_deg > 0) {
        if (loc == CNNB200_DEVICE) {
            padded_rows_kernel<<<ctx->sm_count * 8, 256, 0, st>>>((const ll *)padded, 0, n, max_deg, n,
                                                                  ctx->b_row_ptr.as<ll>(), ctx->b_col_idx.as<int>(), flag);
        } else {
            ll rows_per_chunk = (256ll << 20) / (8 * max_deg);
            if (rows_per_chunk < 1) rows_per_chunk = 1;
            if (rows_per_chunk > n) rows_per_chunk = n;
            CNN_TRY(cnn_ensure(ctx, ctx->b_tmp0, sizeof(ll) * (size_t)(rows_per_chunk * max_deg)));
            for (ll r0 = 0; r0 < n; r0 += rows_per_chunk) {
                ll rows = n - r0 < rows_per_chunk ? n - r0 : rows_per_chunk;
                CNN_CUDA(cudaMemcpyAsync(ctx->b_tmp0.p, padded + r0 * max_deg,
                                         sizeof(ll) * (size_t)(rows * max_deg), kind, st));
                padded_rows_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->b_tmp0.as<ll>(), r0, rows, max_deg, n,
                                                                      ctx->b_row_ptr.as<ll>(), ctx->b_col_idx.as<int>(), flag);
            }
        }
        CNN_KERNEL_CHECK();
    }
    CNN_TRY(read_flag(ctx, flag, &bad));
    if (bad) CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_padded: a neighbour index is outside [0, n)");
    CNN_TRY(cnn_finish_csr(ctx, n, nnz, md, ctx->b_row_ptr.as<ll>(), ctx->b_col_idx.as<int>()));
    return cnn_ensure_sorted_rows(ctx, false, nullptr);  // lists in any order: sorted on arrival
}

// ---- export in the caller's index space ----------------------------------------------------------
__global__ void export_deg_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ perm, ll n, int *deg_o) {
    ll p = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (p > n) return;
    if (p == n) { deg_o[n] = 0; return; }
    deg_o[perm[p]] = (int)(row_ptr[p + 1] - row_ptr[p]);
}

__global__ void export_rows_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col_idx,
                                   const int *__restrict__ perm, ll n, const ll *__restrict__ row_ptr_o,
                                   int *__restrict__ keys) {
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll p = warp; p < n; p += nwarps) {
        ll b = row_ptr[p];
        int m = (int)(row_ptr[p + 1] - b);
        ll o = row_ptr_o[perm[p]];
        for (int k = lane; k < m; k += 32) keys[o + k] = perm[col_idx[b + k]];
    }
}

extern "C" int cnnb200_neighbours_export(cnnb200_ctx *ctx, int64_t *row_ptr, int64_t *col_idx, int loc) {
    if (!ctx) CNN_FAIL(CNNB200_E_INVALID, "ctx is NULL");
    if (!ctx->have_nb) CNN_FAIL(CNNB200_E_STATE, "no neighbourhoods held by this context");
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ll n = ctx->n, nnz = ctx->nnz;
    if (nnz >= (1ll << 31)) CNN_FAIL(CNNB200_E_RANGE, "neighbours_export: nnz too large for the export path");
    cudaMemcpyKind kind = loc == CNNB200_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    const ll *rp = ctx->row_ptr;
    const int *ci = ctx->col_idx;
    if (ctx->perm) {
        // rows go to their caller-space position, columns are mapped through perm and re-sorted per row
        CNN_TRY(cnn_ensure(ctx, ctx->b_tmp0, sizeof(int) * (size_t)(n + 1)));
        CNN_TRY(cnn_ensure(ctx, ctx->b_tmp1, sizeof(ll) * (size_t)(n + 1)));
        CNN_TRY(cnn_ensure(ctx, ctx->b_lab_keys, sizeof(int) * (size_t)nnz));
        CNN_TRY(cnn_ensure(ctx, ctx->b_lab_keys2, sizeof(int) * (size_t)nnz));
        export_deg_kernel<<<cnn_blocks(n + 1, 256), 256, 0, st>>>(rp, ctx->perm, n, ctx->b_tmp0.as<int>());
        CNN_KERNEL_CHECK();
        ll nnz2 = 0, md = 0;
        CNN_TRY(cnn_scan_deg_to_row_ptr(ctx, ctx->b_tmp0.as<int>(), n, ctx->b_tmp1.as<ll>(), &nnz2, &md));
        export_rows_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(rp, ci, ctx->perm, n, ctx->b_tmp1.as<ll>(),
                                                              ctx->b_lab_keys.as<int>());
        CNN_KERNEL_CHECK();
        size_t need = 0;
        CNN_CUDA(cub::DeviceSegmentedSort::SortKeys(nullptr, need, ctx->b_lab_keys.as<int>(),
                                                    ctx->b_lab_keys2.as<int>(), (int)nnz, (int)n,
                                                    ctx->b_tmp1.as<ll>(), ctx->b_tmp1.as<ll>() + 1, st));
        CNN_TRY(cnn_ensure(ctx, ctx->b_cub, need));
        size_t cap = ctx->b_cub.cap;
        CNN_CUDA(cub::DeviceSegmentedSort::SortKeys(ctx->b_cub.p, cap, ctx->b_lab_keys.as<int>(),
                                                    ctx->b_lab_keys2.as<int>(), (int)nnz, (int)n,
                                                    ctx->b_tmp1.as<ll>(), ctx->b_tmp1.as<ll>() + 1, st));
        rp = ctx->b_tmp1.as<ll>();
        ci = ctx->b_lab_keys2.as<int>();
    }
    if (row_ptr) CNN_CUDA(cudaMemcpyAsync(row_ptr, rp, sizeof(ll) * (size_t)(n + 1), kind, st));
    if (col_idx && nnz > 0) {
        ll *dst = col_idx;
        if (loc == CNNB200_HOST) {
            CNN_TRY(cnn_ensure(ctx, ctx->b_lab_vals, sizeof(ll) * (size_t)nnz));
            dst = ctx->b_lab_vals.as<ll>();
        }
        widen_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ci, dst, nnz);
        CNN_KERNEL_CHECK();
        if (loc == CNNB200_HOST)
            CNN_CUDA(cudaMemcpyAsync(col_idx, dst, sizeof(ll) * (size_t)nnz, kind, st));
    }
    CNN_CUDA(cudaStreamSynchronize(st));
    return CNNB200_OK;
}

// ---- symmetry of the lists (debug / validation aid) ------------------------------------------------------------------
__global__ void asymmetry_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col, ll n, unsigned long long *count) {
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    unsigned missing = 0;
    for (ll i = warp; i < n; i += nwarps) {
        const ll b = row_ptr[i], e = row_ptr[i + 1];
        for (ll k = b + lane; k < e; k += 32) {
            const int j = col[k];
            ll lo = row_ptr[j], hi = row_ptr[j + 1];
            while (lo < hi) {  // rows are ascending: lower bound of i in row j
                const ll mid = (lo + hi) >> 1;
                if (col[mid] < (int)i) lo = mid + 1; else hi = mid;
            }
            if (!(lo < row_ptr[j + 1] && col[lo] == (int)i)) ++missing;
        }
    }
    missing = __reduce_add_sync(0xffffffffu, missing);
    if (lane == 0 && missing) atomicAdd(count, (unsigned long long)missing);
}

extern "C" int cnnb200_neighbours_asymmetry(cnnb200_ctx *ctx, int64_t *one_sided_out) {
    if (!ctx || !one_sided_out) CNN_FAIL(CNNB200_E_INVALID, "neighbours_asymmetry: NULL argument");
    if (!ctx->have_nb) CNN_FAIL(CNNB200_E_STATE, "neighbours_asymmetry: no neighbourhoods held by this context");
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp2, 64));
    unsigned long long *cnt = ctx->b_tmp2.as<unsigned long long>() + 4;
    CNN_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), st));
    ll blocks = (ctx->n + 7) / 8;
    if (blocks > (ll)ctx->sm_count * 32) blocks = (ll)ctx->sm_count * 32;
    if (blocks < 1) blocks = 1;
    asymmetry_kernel<<<(int)blocks, 256, 0, st>>>(ctx->row_ptr, ctx->col_idx, ctx->n, cnt);
    CNN_LAUNCHED(ctx, 1);
    CNN_KERNEL_CHECK();
    ll *h;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&h));
    CNN_CUDA(cudaMemcpyAsync(h, cnt, sizeof(ll), cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaStreamSynchronize(st));
    *one_sided_out = h[0];
    return CNNB200_OK;
}
