// packed.cu -- bit-packed neighbour rows and the intersection kernels that run on them.
//
// In a cell-sorted CSR the entries of a row come in a few dense runs of consecutive indices, so a row of 10^2 entries
// touches only ~15 aligned 32-index words.  pack_rows_kernel rewrites every row as its ascending list of
// (word index = v >> 5, 32-bit mask) pairs (in place of the row: same offsets, wcnt[i] pairs).  An intersection
// |N(i) ∩ N(j)| then costs one AND + POPC per word of N(j) -- 6-7x fewer probes than one per entry -- against the
// same shared-memory window of N(i), which is now simply the row's words stored at (word index - first word index).
// Inputs whose rows do not compress (scattered indices) keep the per-entry kernels of cn.cu.
#include "common.cuh"
#include "uf.cuh"

#define PK_WORDS (65536 / 32)  // window words per warp (8 KB): same index span as the per-entry window
#define PK_WARPS 4
#define PK_PF 2                // chunks (of 32 word pairs) prefetched per streamed row
#define PK_BATCH 4

// ---- CSR rows -> packed rows ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_rows_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col_idx, ll n, int2 *__restrict__ wrows,
                 int *__restrict__ wcnt, unsigned long long *total) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    unsigned long long n_words = 0;
    // the bounds and the first chunk of the warp's NEXT row are fetched while the current one is being packed
    ll i = warp;
    ll b = i < n ? row_ptr[i] : 0;
    int m = i < n ? (int)(row_ptr[i + 1] - b) : 0;
    int v_first = lane < m ? col_idx[b + lane] : 0;
    for (; i < n; i += nwarps) {
        const ll i_next = i + nwarps;
        ll b_next = 0;
        int m_next = 0, v_next = 0;
        if (i_next < n) {
            b_next = row_ptr[i_next];
            m_next = (int)(row_ptr[i_next + 1] - b_next);
            v_next = lane < m_next ? col_idx[b_next + lane] : 0;
        }
        int2 *out = wrows + b;
        int wc = 0, ckey = -1;
        unsigned cmask = 0;
        for (int k0 = 0; k0 < m; k0 += 32) {
            const int k = k0 + lane;
            const bool valid = k < m;
            const int v = k0 == 0 ? v_first : (valid ? col_idx[b + k] : 0);
            const int key = valid ? (v >> 5) : -2 - lane;  // invalid lanes never match anything
            const unsigned bit = valid ? 1u << (v & 31) : 0u;
            const int prev = __shfl_up_sync(full, key, 1);
            const bool leader = valid && (lane == 0 || key != prev);
            const unsigned leaders = __ballot_sync(full, leader);
            // the lanes of my word: from the last leader at or below me to the lane before the next leader (keys are
            // ascending, so a word is a run of lanes; padding lanes carry bit 0 and may ride along with the last run)
            const unsigned le = leaders & (full >> (31 - lane));
            const int start = 31 - __clz(le);
            const unsigned above = lane == 31 ? 0u : leaders & (full << (lane + 1));
            const int end = above ? __ffs(above) - 2 : 31;
            const unsigned peers = (full >> (31 - end)) & (full << start);
            const unsigned gmask = __reduce_or_sync(peers, bit);
            const int first_key = __shfl_sync(full, key, 0);
            const bool merge = ckey == first_key;
            if (ckey >= 0 && !merge) {  // the carried word is complete
                if (lane == 0) out[wc] = make_int2(ckey, (int)cmask);
                ++wc;
            }
            const int n_groups = __popc(leaders);
            const int last_leader = 31 - __clz(leaders);
            const int rank = __popc(leaders & lt);
            // every group but the last is complete; the last one is carried into the next chunk
            if (leader && lane != last_leader) out[wc + rank] = make_int2(key, (int)(gmask | ((merge && rank == 0) ? cmask : 0u)));
            wc += n_groups - 1;
            const unsigned lastmask = __shfl_sync(full, gmask, last_leader);
            cmask = lastmask | ((merge && n_groups == 1) ? cmask : 0u);
            ckey = __shfl_sync(full, key, last_leader);
        }
        if (ckey >= 0) {
            if (lane == 0) out[wc] = make_int2(ckey, (int)cmask);
            ++wc;
        }
        if (lane == 0) wcnt[i] = wc;
        n_words += wc;
        b = b_next;
        m = m_next;
        v_first = v_next;
    }
    if (lane == 0 && n_words) atomicAdd(total, n_words);
}

// ---- window of one packed row -----------------------------------------------------------------------------------------
struct PkWin {
    unsigned *words;  // PK_WORDS words owned by the warp; all zero between rows
    int first;        // word index of slot 0
    unsigned span;    // slots in use; 0: the row does not fit
};

__device__ __forceinline__ void pk_set(PkWin &w, const int2 *__restrict__ row, int wc, int lane) {
    w.first = row[0].x;
    const unsigned span = (unsigned)(row[wc - 1].x - w.first) + 1u;
    w.span = span <= (unsigned)PK_WORDS ? span : 0u;
    if (w.span)
        for (int k = lane; k < wc; k += 32) {
            const int2 e = row[k];
            w.words[e.x - w.first] = (unsigned)e.y;
        }
    __syncwarp();
}

__device__ __forceinline__ void pk_clear(PkWin &w, const int2 *__restrict__ row, int wc, int lane) {
    __syncwarp();
    if (w.span)
        for (int k = lane; k < wc; k += 32) w.words[row[k].x - w.first] = 0;
    w.span = 0;
    __syncwarp();
}

__device__ __forceinline__ int pk_probe(const PkWin &w, int2 e) {
    const unsigned idx = (unsigned)(e.x - w.first);
    return idx < w.span ? __popc(w.words[idx] & (unsigned)e.y) : 0;
}

struct PkFetch {
    int2 buf[PK_PF];
    const int2 *rj;
    int wj;
    int src;
};

// padding pairs carry mask 0: they never contribute
__device__ __forceinline__ void pk_fetch(PkFetch &f, unsigned &todo, const int2 *__restrict__ wrows, ll bj, int wj, int lane) {
    if (!todo) { f.src = -1; return; }
    f.src = __ffs(todo) - 1;
    todo &= todo - 1;
    f.rj = wrows + __shfl_sync(0xffffffffu, bj, f.src);
    f.wj = __shfl_sync(0xffffffffu, wj, f.src);
#pragma unroll
    for (int u = 0; u < PK_PF; ++u) {
        const int t = lane + 32 * u;
        f.buf[u] = t < f.wj ? f.rj[t] : make_int2(0, 0);
    }
}

__device__ __forceinline__ int pk_count(const PkFetch &f, int lane, const PkWin &w) {
    int c = 0;
#pragma unroll
    for (int u = 0; u < PK_PF; ++u) c += pk_probe(w, f.buf[u]);
    for (int t = lane + 32 * PK_PF; t < f.wj; t += 32) c += pk_probe(w, f.rj[t]);
    return __reduce_add_sync(0xffffffffu, c);
}

// |A ∩ B| for two packed rows by binary search (rows whose window does not fit)
__device__ __forceinline__ int pk_common_search(const int2 *__restrict__ a, int na, const int2 *__restrict__ b, int nb, int lane) {
    int c = 0;
    for (int k = lane; k < na; k += 32) {
        const int2 e = a[k];
        int l = 0, h = nb;
        while (l < h) {
            const int mid = (l + h) >> 1;
            if (b[mid].x < e.x) l = mid + 1; else h = mid;
        }
        if (l < nb && b[l].x == e.x) c += __popc((unsigned)b[l].y & (unsigned)e.y);
    }
    return __reduce_add_sync(0xffffffffu, c);
}

// ---- count mode on packed rows (same contract as cn_count_kernel) -------------------------------------------------
__global__ void __launch_bounds__(PK_WARPS * 32)
cn_count_packed_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col_idx, const int2 *__restrict__ wrows,
                       const int *__restrict__ wcnt, ll row_begin, ll row_end, ll n_member_cutoff,
                       unsigned *__restrict__ counts, unsigned long long *stats, unsigned long long *row_counter) {
    __shared__ unsigned win_all[PK_WARPS][PK_WORDS];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    PkWin w;
    w.words = win_all[threadIdx.x >> 5];
    w.first = 0;
    w.span = 0;
    for (int t = lane; t < PK_WORDS; t += 32) w.words[t] = 0;
    __syncwarp();
    unsigned long long n_edges = 0, n_fallback = 0;
    for (;;) {
        ll i0 = 0;
        if (lane == 0) i0 = row_begin + (ll)atomicAdd(row_counter, (unsigned long long)PK_BATCH);
        i0 = __shfl_sync(full, i0, 0);
        if (i0 >= row_end) break;
        const ll i1 = i0 + PK_BATCH < row_end ? i0 + PK_BATCH : row_end;
        for (ll i = i0; i < i1; ++i) {
            const ll b = row_ptr[i];
            const int m = (int)(row_ptr[i + 1] - b);
            const int *row = col_idx + b;
            if (m <= n_member_cutoff) {  // not a core row (`_enough`): no slot of it is an edge
                for (int k = lane; k < m; k += 32) counts[b + k] = 0xffffffffu;
                continue;
            }
            const int2 *wi = wrows + b;
            const int wci = wcnt[i];
            pk_set(w, wi, wci, lane);
            if (!w.span) ++n_fallback;
            // slots at or below the diagonal are not edges; the first entry above i starts the candidates
            int above = m;
            for (int k0 = 0; k0 < m; k0 += 32) {
                const int k = k0 + lane;
                const bool up = k < m && row[k] > (int)i;
                const unsigned um = __ballot_sync(full, up);
                if (um) { above = k0 + __ffs(um) - 1; break; }
            }
            for (int k = lane; k < above; k += 32) counts[b + k] = 0xffffffffu;
            // (one edge per lane -- every lane walking the words of its own N(j) against the shared window -- was
            // tried and is slower: uncoalesced 8-byte loads and bank conflicts cost more than the votes save)
            for (int k0 = above; k0 < m; k0 += 32) {
                const int k = k0 + lane;
                const int j = k < m ? row[k] : -1;
                const bool valid = j >= 0;
                const ll bj = valid ? row_ptr[j] : 0;
                const int mj = valid ? (int)(row_ptr[j + 1] - bj) : 0;
                const bool cand = valid && mj > n_member_cutoff;
                const int wj = cand ? wcnt[j] : 0;
                unsigned todo = __ballot_sync(full, cand);
                unsigned out = 0xffffffffu;
                n_edges += __popc(todo);
                auto consume = [&](const PkFetch &f) {
                    const int c = w.span ? pk_count(f, lane, w) : pk_common_search(wi, wci, f.rj, f.wj, lane);
                    if (lane == f.src) out = (unsigned)c;
                };
                PkFetch f0, f1, f2;
                pk_fetch(f0, todo, wrows, bj, wj, lane);
                pk_fetch(f1, todo, wrows, bj, wj, lane);
                for (;;) {
                    if (f0.src < 0) break;
                    pk_fetch(f2, todo, wrows, bj, wj, lane);
                    consume(f0);
                    if (f1.src < 0) break;
                    pk_fetch(f0, todo, wrows, bj, wj, lane);
                    consume(f1);
                    if (f2.src < 0) break;
                    pk_fetch(f1, todo, wrows, bj, wj, lane);
                    consume(f2);
                }
                if (k < m) counts[b + k] = out;
            }
            pk_clear(w, wi, wci, lane);
        }
    }
    if (lane == 0) {
        if (n_edges) atomicAdd(stats + ST_EDGES, n_edges);
        if (n_fallback) atomicAdd(stats + ST_FALLBACK, n_fallback);
    }
}

// ---- work-list pass on packed rows (same contract as comp_worklist_kernel in cn.cu) ----------------------------------
__global__ void __launch_bounds__(PK_WARPS * 32)
comp_worklist_packed_kernel(const ll *__restrict__ row_ptr, const int2 *__restrict__ wrows, const int *__restrict__ wcnt,
                            const int2 *__restrict__ worklist, const unsigned long long *__restrict__ wl_count, int cutoff,
                            int *parent, unsigned long long *stats) {
    __shared__ unsigned win_all[PK_WARPS][PK_WORDS];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    PkWin w;
    w.words = win_all[threadIdx.x >> 5];
    w.first = 0;
    w.span = 0;
    for (int t = lane; t < PK_WORDS; t += 32) w.words[t] = 0;
    __syncwarp();
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    const ll count = (ll)*wl_count;
    unsigned long long n_checks = 0, n_acc = 0, n_skip = 0;
    for (ll e0 = warp * 32; e0 < count; e0 += nwarps * 32) {
        const ll e = e0 + lane;
        const bool valid = e < count;
        const int2 ed = valid ? worklist[e] : make_int2(0, 0);
        const int ri = valid ? uf_find(parent, ed.x) : 0;
        const int rj = valid ? uf_find(parent, ed.y) : 0;
        const bool need = valid && ri != rj;
        const ll bi = need ? row_ptr[ed.x] : 0, bj = need ? row_ptr[ed.y] : 0;
        const int mi = need ? (int)(row_ptr[ed.x + 1] - bi) : 0, mj = need ? (int)(row_ptr[ed.y + 1] - bj) : 0;
        int acc = need && cutoff <= 0 ? 1 : 0;
        const bool coop = need && cutoff > 0 && mi >= cutoff && mj >= cutoff;
        const int wi = coop ? wcnt[ed.x] : 0, wj = coop ? wcnt[ed.y] : 0;
        unsigned work = __ballot_sync(full, coop);
        const unsigned nd = __ballot_sync(full, need);
        n_skip += __popc(__ballot_sync(full, valid)) - __popc(nd);
        n_checks += __popc(nd);
        int cur_x = -1, cur_w = 0;
        const int2 *cur_row = nullptr;
        auto consume = [&](const PkFetch &f) {
            const int x = __shfl_sync(full, ed.x, f.src);
            if (x != cur_x) {  // the scan emits the work list row by row: the window is rebuilt only when i changes
                if (cur_x >= 0) pk_clear(w, cur_row, cur_w, lane);
                cur_x = x;
                cur_row = wrows + __shfl_sync(full, bi, f.src);
                cur_w = __shfl_sync(full, wi, f.src);
                pk_set(w, cur_row, cur_w, lane);
            }
            const int c = w.span ? pk_count(f, lane, w) : pk_common_search(cur_row, cur_w, f.rj, f.wj, lane);
            if (lane == f.src) acc = c >= cutoff;
        };
        PkFetch f0, f1, f2;
        pk_fetch(f0, work, wrows, bj, wj, lane);
        pk_fetch(f1, work, wrows, bj, wj, lane);
        for (;;) {
            if (f0.src < 0) break;
            pk_fetch(f2, work, wrows, bj, wj, lane);
            consume(f0);
            if (f1.src < 0) break;
            pk_fetch(f0, work, wrows, bj, wj, lane);
            consume(f1);
            if (f2.src < 0) break;
            pk_fetch(f1, work, wrows, bj, wj, lane);
            consume(f2);
        }
        if (cur_x >= 0) pk_clear(w, cur_row, cur_w, lane);
        if (acc) uf_union(parent, ri, rj);
        n_acc += __popc(__ballot_sync(full, acc != 0));
    }
    if (lane == 0) {
        if (n_checks) atomicAdd(stats + ST_CHECKS, n_checks);
        if (n_acc) atomicAdd(stats + ST_ACCEPTED, n_acc);
        if (n_skip) atomicAdd(stats + ST_SKIPPED, n_skip);
    }
}

int cnn_worklist_packed(cnnb200_ctx *ctx, const int2 *worklist, const unsigned long long *wl_count, int cutoff, int *parent) {
    static int ctas = 0;
    if (!ctas) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, comp_worklist_packed_kernel, PK_WARPS * 32, 0) != cudaSuccess || ctas < 1) {
            cudaGetLastError();
            ctas = 1;
        }
    }
    comp_worklist_packed_kernel<<<ctx->sm_count * ctas, PK_WARPS * 32, 0, ctx->stream>>>(
        ctx->row_ptr, ctx->b_wrows.as<int2>(), ctx->b_wcnt.as<int>(), worklist, wl_count, cutoff, parent,
        (unsigned long long *)ctx->b_stats.p);
    CNN_KERNEL_CHECK();
    return CNNB200_OK;
}

// ---- host side -----------------------------------------------------------------------------------------------------------
// Packed rows of the CSR the context holds (built once per neighbourhood structure); *usable_out = false when the
// rows do not compress enough to pay for the wider pairs.
int cnn_pack_rows(cnnb200_ctx *ctx, bool *usable_out) {
    if (ctx->packed_version != ctx->nb_version) {
        cudaStream_t st = ctx->stream;
        const ll n = ctx->n, nnz = ctx->nnz;
        CNN_TRY(cnn_ensure(ctx, ctx->b_wrows, sizeof(int2) * (size_t)(nnz > 0 ? nnz : 1)));
        CNN_TRY(cnn_ensure(ctx, ctx->b_wcnt, sizeof(int) * (size_t)(n + 1)));
        CNN_TRY(cnn_ensure(ctx, ctx->b_tmp2, 64));
        unsigned long long *total = (unsigned long long *)(ctx->b_tmp2.as<char>() + 56);
        CNN_CUDA(cudaMemsetAsync(total, 0, sizeof(unsigned long long), st));
        int blocks = (int)((n + 7) / 8);
        if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
        if (blocks < 1) blocks = 1;
        {
            Span sp(ctx, T_INGEST);
            pack_rows_kernel<<<blocks, 256, 0, st>>>(ctx->row_ptr, ctx->col_idx, n, ctx->b_wrows.as<int2>(), ctx->b_wcnt.as<int>(), total);
            CNN_LAUNCHED(ctx, 1);
        }
        CNN_KERNEL_CHECK();
        unsigned long long *h;
        CNN_TRY(cnn_pinned(ctx, 64, (void **)&h));
        CNN_CUDA(cudaMemcpyAsync(h, total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CNN_CUDA(cudaStreamSynchronize(st));
        ctx->packed_words = (ll)h[0];
        ctx->packed_version = ctx->nb_version;
    }
    // 8 bytes per word against 4 per entry: worth it from a compression of ~2.5 on
    *usable_out = ctx->use_packed && ctx->packed_words * 5 <= ctx->nnz * 2;
    return CNNB200_OK;
}

int cnn_count_packed(cnnb200_ctx *ctx, ll n_member_cutoff, ll row_begin, ll row_end, unsigned *d_counts) {
    cudaStream_t st = ctx->stream;
    static int ctas = 0;
    if (!ctas) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, cn_count_packed_kernel, PK_WARPS * 32, 0) != cudaSuccess || ctas < 1) {
            cudaGetLastError();
            ctas = 1;
        }
    }
    int blocks = ctx->sm_count * ctas;
    const ll want = (row_end - row_begin + PK_WARPS * PK_BATCH - 1) / (PK_WARPS * PK_BATCH);
    if (blocks > want) blocks = (int)want;
    if (blocks < 1) blocks = 1;
    unsigned long long *stats = (unsigned long long *)ctx->b_stats.p;
    Span sp(ctx, T_CN_COUNT);
    cn_count_packed_kernel<<<blocks, PK_WARPS * 32, 0, st>>>(ctx->row_ptr, ctx->col_idx, ctx->b_wrows.as<int2>(),
                                                            ctx->b_wcnt.as<int>(), row_begin, row_end, n_member_cutoff, d_counts,
                                                            stats, stats + ST_COUNTER);
    CNN_LAUNCHED(ctx, 1);
    CNN_KERNEL_CHECK();
    return CNNB200_OK;
}
