// predict.cu -- label prediction for points outside the clustered set.
//
// Bulk replacement of PredictorFirstmatch.predict (reference _fit.pyx:1231-1291) with
// NeighboursGetterExtBruteForce._get_other (_types.pyx:1712-1731) / NeighboursGetterExtLookup._get_other
// (:1819-1832): a point p of the OTHER set gets the label of the first member m (ascending index in the clustered
// set) of its neighbourhood N_o(p) in the clustered set that (i) carries a label of the consider set, (ii) has
// more than n_member_cutoff neighbours itself and (iii) shares at least similarity_cutoff entries of N(m) with
// N_o(p); p needs more than n_member_cutoff neighbours in the first place.
//
// The reference walks the members in ascending caller index and stops at the first match; here every member is
// tested by a warp (N_o(p) as a shared-memory window, N(m) streamed through it -- the machinery of the fit's work
// list) and the match with the smallest caller index wins, which is the same label.
#include <cub/cub.cuh>

#include "common.cuh"

// ---- neighbourhoods of the other points in the clustered set ------------------------------------------------------
__device__ __forceinline__ int cell_of(double v, double lo, double h, int dim) {
    int c = (int)floor(__ddiv_rn(__dsub_rn(v, lo), h));
    return min(max(c, 0), dim - 1);
}

template <int D, bool SQRT>
__device__ __forceinline__ bool within_other(const double *q, const double *__restrict__ xs, ll n, int c, double rc) {
    double t = __dsub_rn(q[0], xs[c]);
    double s = __dmul_rn(t, t);
    if (D > 1) { t = __dsub_rn(q[1], xs[n + c]); s = __dadd_rn(s, __dmul_rn(t, t)); }
    if (D > 2) { t = __dsub_rn(q[2], xs[2 * n + c]); s = __dadd_rn(s, __dmul_rn(t, t)); }
    if (SQRT) s = __dsqrt_rn(s);
    return s <= rc;
}

// one warp per other point; candidates = the 3^D cells around the (clamped) cell of the point in the clustered
// set's grid; FILL = false counts, FILL = true writes the (cell-sorted) indices ascending
template <int D, bool SQRT, bool FILL>
__global__ void __launch_bounds__(256)
other_rows_kernel(const double *__restrict__ xo, ll m, const double *__restrict__ xs, ll n,
                  const int *__restrict__ cell_start, GridDesc g, double rc, int *__restrict__ deg,
                  const ll *__restrict__ row_ptr, int *__restrict__ col_idx) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll p = warp; p < m; p += nwarps) {
        double q[3] = {0.0, 0.0, 0.0};
        int c[3] = {0, 0, 0};
        bool reach = true;  // a point further than one cell outside the grid has no neighbour
        for (int k = 0; k < D; ++k) {
            q[k] = xo[p * D + k];
            const double u = floor(__ddiv_rn(__dsub_rn(q[k], g.lo[k]), g.h));
            if (u < -1.0 || u > (double)g.dims[k]) reach = false;
            c[k] = cell_of(q[k], g.lo[k], g.h, g.dims[k]);
        }
        int cnt = 0;
        const ll base = FILL ? row_ptr[p] : 0;
        if (reach) {
            const int x0 = max(c[0] - 1, 0), x1 = min(c[0] + 1, g.dims[0] - 1);
            const int y0 = D > 1 ? max(c[1] - 1, 0) : 0, y1 = D > 1 ? min(c[1] + 1, g.dims[1] - 1) : 0;
            const int z0 = D > 2 ? max(c[2] - 1, 0) : 0, z1 = D > 2 ? min(c[2] + 1, g.dims[2] - 1) : 0;
            for (int z = z0; z <= z1; ++z)
                for (int y = y0; y <= y1; ++y) {
                    const int cb = (z * g.dims[1] + y) * g.dims[0];
                    const int s = cell_start[cb + x0], e = cell_start[cb + x1 + 1];
                    for (int c0 = s; c0 < e; c0 += 32) {
                        const int cc = c0 + lane;
                        const bool hit = cc < e && within_other<D, SQRT>(q, xs, n, cc, rc);
                        const unsigned mk = __ballot_sync(full, hit);
                        if (FILL && hit) col_idx[base + cnt + __popc(mk & lt)] = cc;
                        cnt += __popc(mk);
                    }
                }
        }
        if (!FILL && lane == 0) deg[p] = cnt;
    }
}

// any dimension: all clustered points are candidates (row-major coordinates, internal order = caller order)
template <bool SQRT, bool FILL>
__global__ void __launch_bounds__(256)
other_rows_dense_kernel(const double *__restrict__ xo, ll m, const double *__restrict__ x, ll n, int d, double rc,
                        int *__restrict__ deg, const ll *__restrict__ row_ptr, int *__restrict__ col_idx) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll p = warp; p < m; p += nwarps) {
        const double *q = xo + p * d;
        int cnt = 0;
        const ll base = FILL ? row_ptr[p] : 0;
        for (ll c0 = 0; c0 < n; c0 += 32) {
            const ll cc = c0 + lane;
            bool hit = false;
            if (cc < n) {
                const double *b = x + cc * d;
                double s = 0.0;
                for (int k = 0; k < d; ++k) {
                    const double t = __dsub_rn(q[k], b[k]);
                    s = __dadd_rn(s, __dmul_rn(t, t));
                }
                if (SQRT) s = __dsqrt_rn(s);
                hit = s <= rc;
            }
            const unsigned mk = __ballot_sync(full, hit);
            if (FILL && hit) col_idx[base + cnt + __popc(mk & lt)] = (int)cc;
            cnt += __popc(mk);
        }
        if (!FILL && lane == 0) deg[p] = cnt;
    }
}

// ---- the prediction -------------------------------------------------------------------------------------------------
#define PW_BITS 65536
#define PW_WARPS 4

__device__ __forceinline__ int pr_common(const int *__restrict__ a, int na, const int *__restrict__ b, int nb, int lane) {
    // |A ∩ B| for ascending lists (fallback when the window does not fit): lanes binary-search the entries of A in B
    int count = 0;
    for (int base = 0; base < na; base += 32) {
        const int k = base + lane;
        bool found = false;
        if (k < na) {
            const int v = a[k];
            int l = 0, h = nb;
            while (l < h) {
                const int mid = (l + h) >> 1;
                if (b[mid] < v) l = mid + 1; else h = mid;
            }
            found = l < nb && b[l] == v;
        }
        count += __popc(__ballot_sync(0xffffffffu, found));
    }
    return count;
}

__global__ void __launch_bounds__(PW_WARPS * 32)
predict_kernel(const ll *__restrict__ o_row_ptr, const int *__restrict__ o_col_idx, ll m, const ll *__restrict__ row_ptr,
               const int *__restrict__ col_idx, const int *__restrict__ perm, const ll *__restrict__ labels,
               const uint8_t *__restrict__ label_ok, ll max_label, ll n_member_cutoff, ll similarity_cutoff,
               ll *__restrict__ other_labels, uint8_t *__restrict__ other_consider, unsigned long long *row_counter,
               const int *__restrict__ o_col_orig) {
    __shared__ unsigned win_all[PW_WARPS][PW_BITS / 32];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    unsigned *win = win_all[threadIdx.x >> 5];
    for (int t = lane; t < PW_BITS / 32; t += 32) win[t] = 0;
    __syncwarp();
    for (;;) {
        ll p = 0;
        if (lane == 0) p = (ll)atomicAdd(row_counter, 1ull);
        p = __shfl_sync(full, p, 0);
        if (p >= m) break;
        if (other_consider[p] == 0) continue;
        const ll ob = o_row_ptr[p];
        const int mo = (int)(o_row_ptr[p + 1] - ob);
        if (mo <= n_member_cutoff) continue;  // `enough`: n_points > cutoff (_types.pyx:1408-1411)
        const int *orow = o_col_idx + ob;
        const int first = orow[0];
        const unsigned span = (unsigned)(orow[mo - 1] - first) + 1u;
        const bool windowed = span <= (unsigned)PW_BITS;
        if (windowed) {
            for (int k = lane; k < mo; k += 32) {
                const unsigned idx = (unsigned)(orow[k] - first);
                atomicOr(&win[idx >> 5], 1u << (idx & 31));
            }
            __syncwarp();
        }
        // The reference takes the FIRST matching member in the order of the predictand's list (_fit.pyx:1264-1291).
        // Lists built here from coordinates are ascending in the internal order: the first match of the reference's
        // ascending caller-index list is the matching member with the smallest caller index.  Lists the caller gave
        // (o_col_orig: their original order, when they had to be sorted for the window) are walked in that order and
        // the first match ends the walk.
        const int *scan = o_col_orig ? o_col_orig + ob : orow;
        int best = 0x7fffffff;  // smallest caller index among the matching members
        for (int k = 0; k < mo; ++k) {
            const int mm = scan[k];
            const int caller = perm ? perm[mm] : mm;
            if (caller >= best) continue;
            const ll lab = labels[caller];
            if (lab < 0 || lab > max_label || !label_ok[lab]) continue;
            const ll mb = row_ptr[mm];
            const int nm = (int)(row_ptr[mm + 1] - mb);
            if (nm <= n_member_cutoff) continue;
            const int *mrow = col_idx + mb;
            int c = 0;
            if (similarity_cutoff > 0) {
                if (windowed) {
                    for (int t = lane; t < nm; t += 32) {
                        const unsigned idx = (unsigned)(mrow[t] - first);
                        if (idx < span) c += (win[idx >> 5] >> (idx & 31)) & 1u;
                    }
                    c = __reduce_add_sync(full, c);
                } else {
                    c = pr_common(orow, mo, mrow, nm, lane);
                }
            }
            if ((ll)c >= similarity_cutoff) {
                best = caller;
                if (o_col_orig || !perm) break;  // list order is the reference's order: the first match decides
            }
        }
        if (windowed) {
            __syncwarp();
            for (int k = lane; k < mo; k += 32) win[(unsigned)(orow[k] - first) >> 5] = 0;
            __syncwarp();
        }
        if (best != 0x7fffffff && lane == 0) {
            other_labels[p] = labels[best];
            other_consider[p] = 0;
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------
static int other_scan(cnnb200_ctx *ctx, ll m, ll *nnz_out) {
    size_t need = 0;
    int *deg = ctx->b_o_deg.as<int>();
    ll *rp = ctx->b_o_row_ptr.as<ll>();
    CNN_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, need, deg, rp, ::cuda::std::plus<>{}, (ll)0, (int)(m + 1), ctx->stream));
    CNN_TRY(cnn_ensure(ctx, ctx->b_cub, need));
    size_t cap = ctx->b_cub.cap;
    CNN_CUDA(cub::DeviceScan::ExclusiveScan(ctx->b_cub.p, cap, deg, rp, ::cuda::std::plus<>{}, (ll)0, (int)(m + 1), ctx->stream));
    ll *h;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&h));
    CNN_CUDA(cudaMemcpyAsync(h, rp + m, sizeof(ll), cudaMemcpyDeviceToHost, ctx->stream));
    CNN_CUDA(cudaStreamSynchronize(ctx->stream));
    *nnz_out = h[0];
    return CNNB200_OK;
}

extern "C" int cnnb200_neighbours_other_from_points(cnnb200_ctx *ctx, const double *x_other, int loc, int64_t m, int64_t d,
                                                    int64_t *nnz_out) {
    if (!ctx || !x_other) CNN_FAIL(CNNB200_E_INVALID, "neighbours_other_from_points: NULL argument");
    if (!ctx->have_nb || !ctx->have_points)
        CNN_FAIL(CNNB200_E_STATE, "neighbours_other_from_points: call neighbours_from_points for the clustered set first");
    if (d != ctx->pts_d) CNN_FAIL(CNNB200_E_INVALID, "neighbours_other_from_points: dimension differs from the clustered set");
    if (m < 0 || m >= (1ll << 31) - 1) CNN_FAIL(CNNB200_E_RANGE, "neighbours_other_from_points: m out of range");
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->have_other = false;
    const double *d_xo = x_other;
    if (loc == CNNB200_HOST) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_o_x, sizeof(double) * (size_t)(m > 0 ? m : 1) * (size_t)d));
        CNN_CUDA(cudaMemcpyAsync(ctx->b_o_x.p, x_other, sizeof(double) * (size_t)m * (size_t)d, cudaMemcpyHostToDevice, st));
        d_xo = ctx->b_o_x.as<double>();
    }
    CNN_TRY(cnn_ensure(ctx, ctx->b_o_deg, sizeof(int) * (size_t)(m + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_o_row_ptr, sizeof(ll) * (size_t)(m + 1)));
    CNN_CUDA(cudaMemsetAsync(ctx->b_o_deg.p, 0, sizeof(int) * (size_t)(m + 1), st));
    int blocks = (int)((m + 7) / 8);
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    if (blocks < 1) blocks = 1;
    const bool sq = ctx->pts_metric == CNNB200_METRIC_EUCLIDEAN;
    const double rc = ctx->pts_rc;
    const bool grid = ctx->perm != nullptr && ctx->have_grid && ctx->grid_n == ctx->n && ctx->pts_grid;
    const ll n = ctx->n;
    ll nnz = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int *deg = ctx->b_o_deg.as<int>();
        const ll *rp = ctx->b_o_row_ptr.as<ll>();
        int *ci = pass ? ctx->b_o_col_idx.as<int>() : nullptr;
        if (grid) {
            const double *xs = ctx->b_xs.as<double>();
            const int *cst = ctx->b_cell_start.as<int>();
            const GridDesc g = ctx->grid;
#define LAUNCH_O(DD, SQ)                                                                                                  \
    do {                                                                                                                  \
        if (pass) other_rows_kernel<DD, SQ, true><<<blocks, 256, 0, st>>>(d_xo, m, xs, n, cst, g, rc, deg, rp, ci);       \
        else other_rows_kernel<DD, SQ, false><<<blocks, 256, 0, st>>>(d_xo, m, xs, n, cst, g, rc, deg, rp, ci);          \
    } while (0)
            switch (g.d * 2 + (sq ? 1 : 0)) {
                case 2: LAUNCH_O(1, false); break;
                case 3: LAUNCH_O(1, true); break;
                case 4: LAUNCH_O(2, false); break;
                case 5: LAUNCH_O(2, true); break;
                case 6: LAUNCH_O(3, false); break;
                default: LAUNCH_O(3, true); break;
            }
#undef LAUNCH_O
        } else {
            if (ctx->perm) CNN_FAIL(CNNB200_E_STATE, "neighbours_other_from_points: unsupported neighbour structure (use method grid or tiled)");
            const double *x = ctx->pts_x;
            if (sq) {
                if (pass) other_rows_dense_kernel<true, true><<<blocks, 256, 0, st>>>(d_xo, m, x, n, (int)d, rc, deg, rp, ci);
                else other_rows_dense_kernel<true, false><<<blocks, 256, 0, st>>>(d_xo, m, x, n, (int)d, rc, deg, rp, ci);
            } else {
                if (pass) other_rows_dense_kernel<false, true><<<blocks, 256, 0, st>>>(d_xo, m, x, n, (int)d, rc, deg, rp, ci);
                else other_rows_dense_kernel<false, false><<<blocks, 256, 0, st>>>(d_xo, m, x, n, (int)d, rc, deg, rp, ci);
            }
        }
        CNN_LAUNCHED(ctx, 1);
        CNN_KERNEL_CHECK();
        if (pass == 0) {
            CNN_TRY(other_scan(ctx, m, &nnz));
            CNN_TRY(cnn_ensure(ctx, ctx->b_o_col_idx, sizeof(int) * (size_t)(nnz > 0 ? nnz : 1)));
        }
    }
    ctx->o_m = m;
    ctx->o_nnz = nnz;
    ctx->o_col_orig = nullptr;
    ctx->have_other = true;
    if (nnz_out) *nnz_out = nnz;
    return CNNB200_OK;
}

__global__ void other_narrow_kernel(const ll *__restrict__ in, ll count, int *__restrict__ out) {
    ll k = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (k < count) out[k] = (in[k] < 0 || in[k] > 0x7fffffffll) ? -1 : (int)in[k];  // out of range: caught by the check
}

// offsets monotone, first 0, last nnz (flag[0]); indices inside [0, n) (flag[0]); rows ascending (flag[2])
__global__ void other_check_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col, ll m, ll nnz, ll n, int *flag) {
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    bool bad = false, unsorted = false;
    if (warp == 0 && lane == 0) bad = row_ptr[0] != 0 || row_ptr[m] != nnz;
    for (ll r = warp; r < m; r += nwarps) {
        const ll b = row_ptr[r], e = row_ptr[r + 1];
        if (b < 0 || e > nnz || b > e) { bad = true; continue; }
        for (ll k = b + lane; k < e; k += 32) {
            const int v = col[k];
            bad |= v < 0 || v >= n;
            if (k + 1 < e) unsorted |= v > col[k + 1];
        }
    }
    if (bad) flag[0] = 1;
    if (unsorted) flag[2] = 1;
}

extern "C" int cnnb200_neighbours_other_from_csr(cnnb200_ctx *ctx, const int64_t *row_ptr, const int64_t *col_idx, int loc,
                                                 int64_t m, int64_t nnz) {
    if (!ctx || !row_ptr || (!col_idx && nnz > 0)) CNN_FAIL(CNNB200_E_INVALID, "neighbours_other_from_csr: NULL argument");
    if (!ctx->have_nb) CNN_FAIL(CNNB200_E_STATE, "neighbours_other_from_csr: no neighbourhoods of the clustered set held");
    if (ctx->perm) CNN_FAIL(CNNB200_E_STATE, "neighbours_other_from_csr: the clustered set must be held in the caller's index space (precomputed neighbourhoods)");
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->have_other = false;
    CNN_TRY(cnn_ensure(ctx, ctx->b_o_row_ptr, sizeof(ll) * (size_t)(m + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_o_col_idx, sizeof(int) * (size_t)(nnz > 0 ? nnz : 1)));
    const cudaMemcpyKind kind = loc == CNNB200_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    CNN_CUDA(cudaMemcpyAsync(ctx->b_o_row_ptr.p, row_ptr, sizeof(ll) * (size_t)(m + 1), kind, st));
    if (nnz > 0) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_tmp0, sizeof(ll) * (size_t)nnz));
        CNN_CUDA(cudaMemcpyAsync(ctx->b_tmp0.p, col_idx, sizeof(ll) * (size_t)nnz, kind, st));
        other_narrow_kernel<<<cnn_blocks(nnz, 256), 256, 0, st>>>(ctx->b_tmp0.as<ll>(), nnz, ctx->b_o_col_idx.as<int>());
        CNN_LAUNCHED(ctx, 1);
        CNN_KERNEL_CHECK();
    }
    // the lists index the clustered set: offsets monotone and ending at nnz, indices inside [0, n); the kernel's window
    // needs ascending rows -- lists in any order (NeighboursGetterExtLookup with is_sorted = False) are sorted into a
    // second array and the original order is kept for the first-match walk
    ctx->o_col_orig = nullptr;
    if (m > 0) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_tmp2, 64));
        int *flag = ctx->b_tmp2.as<int>();
        CNN_CUDA(cudaMemsetAsync(flag, 0, 4 * sizeof(int), st));
        other_check_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->b_o_row_ptr.as<ll>(), ctx->b_o_col_idx.as<int>(), m, nnz, ctx->n,
                                                             flag);
        CNN_LAUNCHED(ctx, 1);
        CNN_KERNEL_CHECK();
        int *h;
        CNN_TRY(cnn_pinned(ctx, 64, (void **)&h));
        CNN_CUDA(cudaMemcpyAsync(h, flag, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
        CNN_CUDA(cudaStreamSynchronize(st));
        if (h[0]) CNN_FAIL(CNNB200_E_INVALID, "neighbours_other_from_csr: row_ptr not monotone / not ending at nnz, or an index outside [0, n)");
        if (h[2]) {
            CNN_TRY(cnn_sort_csr_rows(ctx, ctx->b_o_row_ptr.as<ll>(), ctx->b_o_col_idx.as<int>(), m, nnz, ctx->b_o_col_sorted));
            ctx->o_col_orig = ctx->b_o_col_idx.as<int>();
        }
    }
    ctx->o_m = m;
    ctx->o_nnz = nnz;
    ctx->have_other = true;
    return CNNB200_OK;
}

extern "C" int cnnb200_predict(cnnb200_ctx *ctx, int64_t n_member_cutoff, int64_t similarity_cutoff, const int64_t *labels,
                               const uint8_t *label_ok, int64_t max_label, int64_t *other_labels, uint8_t *other_consider,
                               int loc) {
    if (!ctx || !labels || !label_ok || !other_labels || !other_consider)
        CNN_FAIL(CNNB200_E_INVALID, "predict: NULL argument");
    if (!ctx->have_nb || !ctx->have_other)
        CNN_FAIL(CNNB200_E_STATE, "predict: needs the neighbourhoods of the clustered set and of the other set");
    if (n_member_cutoff < 0 || similarity_cutoff < 0 || max_label < 0) CNN_FAIL(CNNB200_E_INVALID, "predict: negative argument");
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const ll n = ctx->n, m = ctx->o_m;
    if (m == 0) return CNNB200_OK;
    const ll *d_labels = labels;
    const uint8_t *d_ok = label_ok;
    ll *d_ol = other_labels;
    uint8_t *d_oc = other_consider;
    if (loc == CNNB200_HOST) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_labels, sizeof(ll) * (size_t)n));
        CNN_TRY(cnn_ensure(ctx, ctx->b_o_labels, sizeof(ll) * (size_t)m + (size_t)m + (size_t)(max_label + 1) + 64));
        d_labels = ctx->b_labels.as<ll>();
        d_ol = ctx->b_o_labels.as<ll>();
        d_oc = (uint8_t *)(d_ol + m);
        uint8_t *ok = d_oc + m;
        d_ok = ok;
        CNN_CUDA(cudaMemcpyAsync((void *)d_labels, labels, sizeof(ll) * (size_t)n, cudaMemcpyHostToDevice, st));
        CNN_CUDA(cudaMemcpyAsync(d_ol, other_labels, sizeof(ll) * (size_t)m, cudaMemcpyHostToDevice, st));
        CNN_CUDA(cudaMemcpyAsync(d_oc, other_consider, (size_t)m, cudaMemcpyHostToDevice, st));
        CNN_CUDA(cudaMemcpyAsync(ok, label_ok, (size_t)(max_label + 1), cudaMemcpyHostToDevice, st));
    }
    CNN_TRY(cnn_ensure(ctx, ctx->b_stats, sizeof(unsigned long long) * ST_N));
    unsigned long long *row_counter = (unsigned long long *)ctx->b_stats.p + ST_COUNTER;
    CNN_CUDA(cudaMemsetAsync(row_counter, 0, sizeof(unsigned long long), st));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, predict_kernel, PW_WARPS * 32, 0) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    ll want = (m + PW_WARPS - 1) / PW_WARPS;
    int blocks = ctx->sm_count * per_sm;
    if (blocks > want) blocks = (int)want;
    const int *o_sorted = ctx->o_col_orig ? ctx->b_o_col_sorted.as<int>() : ctx->b_o_col_idx.as<int>();
    predict_kernel<<<blocks, PW_WARPS * 32, 0, st>>>(ctx->b_o_row_ptr.as<ll>(), o_sorted, m, ctx->row_ptr,
                                                    ctx->col_idx, ctx->perm, d_labels, d_ok, max_label, n_member_cutoff,
                                                    similarity_cutoff, d_ol, d_oc, row_counter, ctx->o_col_orig);
    CNN_LAUNCHED(ctx, 1);
    CNN_KERNEL_CHECK();
    if (loc == CNNB200_HOST) {
        CNN_CUDA(cudaMemcpyAsync(other_labels, d_ol, sizeof(ll) * (size_t)m, cudaMemcpyDeviceToHost, st));
        CNN_CUDA(cudaMemcpyAsync(other_consider, d_oc, (size_t)m, cudaMemcpyDeviceToHost, st));
        CNN_CUDA(cudaStreamSynchronize(st));
    }
    return CNNB200_OK;
}
