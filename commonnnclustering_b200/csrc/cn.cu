// cn.cu -- K5/K6: common-nearest-neighbour similarity over sorted CSR rows + concurrent
// union-find over the accepted edges.
//
// Bulk replacement of FitterExtBFS._fit (reference _fit.pyx:408-491) and of the per-pair
// SimilarityCheckerExt*._check / ._get (_types.pyx:2580-2610, :2654-2706, :2708-2753).
//
// The reference grows one cluster at a time with a queue and a `consider` byte mask; its result
// is exactly "connected components of the graph whose edges are the pairs of core points (i, j)
// with j in N(i) and |N(i) ∩ N(j)| >= similarity_cutoff" (validated against
// the reference).  Components do not depend on the order edges are examined in, so here all
// candidate edges are examined concurrently and merged with a lock-free union-find that always
// hooks the larger root under the smaller one (the final root of a component is its smallest
// member, independent of scheduling and of how rows are sharded over GPUs).
//
// Like the reference's BFS -- which never re-tests a point once `consider` is cleared
// (_fit.pyx:458-459) -- an edge whose endpoints are already connected is not evaluated:
// a sampling pass links every core point with a few of its neighbours first (Afforest-style),
// after which most candidate edges fall inside one tree and cost two parent look-ups.
#include "common.cuh"

__device__ __forceinline__ int ld_parent(const int *parent, int x) { return __ldcg(parent + x); }

// find with path halving; parents only ever move towards smaller indices
__device__ __forceinline__ int uf_find(int *parent, int x) {
    int p = ld_parent(parent, x);
    while (p != x) {
        int gp = ld_parent(parent, p);
        if (gp != p) parent[x] = gp;
        x = p;
        p = gp;
    }
    return x;
}

__device__ __forceinline__ int uf_find_ro(const int *parent, int x) {
    int p = ld_parent(parent, x);
    while (p != x) {
        x = p;
        p = ld_parent(parent, x);
    }
    return x;
}

// returns true when the call linked two previously separate trees
__device__ __forceinline__ bool uf_union(int *parent, int a, int b) {
    int ra = uf_find(parent, a), rb = uf_find(parent, b);
    while (ra != rb) {
        if (ra < rb) { int t = ra; ra = rb; rb = t; }
        int old = atomicCAS(parent + ra, ra, rb);
        if (old == ra) return true;
        ra = uf_find(parent, ra);
        rb = uf_find(parent, rb);
    }
    return false;
}

// Warp-cooperative |A ∩ B| over ascending int32 lists.
//   CHECK = true : early exit, returns 1 as soon as `cutoff` common elements are found, 0 when
//                  that has become impossible (SimilarityChecker._check semantics).
//   CHECK = false: returns the full count (SimilarityChecker._get semantics).
// Lanes take 32 consecutive elements of the shorter list and binary-search them in the
// not-yet-consumed tail of the longer one.
template <bool CHECK>
__device__ __forceinline__ int warp_common(const int *__restrict__ a, int na, const int *__restrict__ b,
                                           int nb, int cutoff, int lane) {
    if (CHECK && cutoff <= 0) return 1;
    if (na > nb) {
        const int *t = a; a = b; b = t;
        int tn = na; na = nb; nb = tn;
    }
    if (CHECK && na < cutoff) return 0;
    int count = 0, lo = 0;
    for (int base = 0; base < na; base += 32) {
        int k = base + lane;
        int v = k < na ? a[k] : 0x7fffffff;
        // lower_bound of v in b[lo, nb)
        int l = lo, h = nb;
        while (l < h) {
            int mid = (l + h) >> 1;
            if (b[mid] < v) l = mid + 1; else h = mid;
        }
        bool found = k < na && l < nb && b[l] == v;
        count += __popc(__ballot_sync(0xffffffffu, found));
        if (CHECK) {
            if (count >= cutoff) return 1;
            if (count + (na - base - 32) < cutoff) return 0;
        }
        lo = __shfl_sync(0xffffffffu, l, 31);
        if (lo >= nb) break;
    }
    return CHECK ? 0 : count;
}

__global__ void comp_init_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ perm,
                                 const uint8_t *__restrict__ consider, ll n, ll n_member_cutoff,
                                 int *parent, uint8_t *core) {
    ll p = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (p >= n) return;
    ll deg = row_ptr[p + 1] - row_ptr[p];
    bool c = deg > n_member_cutoff;  // `_enough`: n_points > cutoff (_types.pyx:1408-1411)
    if (c && consider) c = consider[perm ? perm[p] : p] != 0;
    core[p] = c ? 1 : 0;
    parent[p] = (int)p;
}

// Sampling pass: every core row links with the core neighbours that sit next to it in its own
// row (closest in the cell-sorted order) and with the first and last core entries of the row.
__global__ void __launch_bounds__(256)
comp_sample_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col_idx,
                   const uint8_t *__restrict__ core, ll row_begin, ll row_end, int cutoff, int *parent,
                   unsigned long long *stats) {
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    unsigned long long n_checks = 0, n_acc = 0;
    for (ll i = row_begin + warp; i < row_end; i += nwarps) {
        if (!core[i]) continue;
        ll b = row_ptr[i];
        int m = (int)(row_ptr[i + 1] - b);
        const int *row = col_idx + b;
        // position of the first entry > i (rows are ascending)
        int l = 0, h = m;
        while (l < h) {
            int mid = (l + h) >> 1;
            if (row[mid] <= (int)i) l = mid + 1; else h = mid;
        }
        // candidates: next after self, previous before self, first, last
        int cand[4];
        cand[0] = l < m ? row[l] : -1;
        int ps = l - 1;
        if (ps >= 0 && row[ps] == (int)i) --ps;
        cand[1] = ps >= 0 ? row[ps] : -1;
        cand[2] = m > 0 ? row[0] : -1;
        cand[3] = m > 0 ? row[m - 1] : -1;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            int j = cand[t];
            if (j < 0 || j == (int)i || !core[j]) continue;
            bool dup = false;
            for (int u = 0; u < t; ++u) dup |= cand[u] == j;
            if (dup) continue;
            int same = 0;
            if (lane == 0) same = uf_find(parent, (int)i) == uf_find(parent, j);
            same = __shfl_sync(0xffffffffu, same, 0);
            if (same) continue;
            ll bj = row_ptr[j];
            int mj = (int)(row_ptr[j + 1] - bj);
            ++n_checks;
            if (warp_common<true>(row, m, col_idx + bj, mj, cutoff, lane)) {
                if (lane == 0) uf_union(parent, (int)i, j);
                ++n_acc;
            }
        }
    }
    if (lane == 0) {
        if (n_checks) atomicAdd(stats + ST_CHECKS, n_checks);
        if (n_acc) atomicAdd(stats + ST_ACCEPTED, n_acc);
    }
}

__global__ void comp_compress_kernel(int *parent, ll n) {
    ll p = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (p >= n) return;
    int r = uf_find_ro(parent, (int)p);
    parent[p] = r;
}

// Full pass: every candidate edge (i < j, both core) of the row block is settled -- skipped when
// both ends already share a root, otherwise evaluated and linked when similar enough.
__global__ void __launch_bounds__(256)
comp_edges_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col_idx,
                  const uint8_t *__restrict__ core, ll row_begin, ll row_end, int cutoff, int *parent,
                  unsigned long long *stats) {
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    unsigned long long n_edges = 0, n_checks = 0, n_acc = 0, n_skip = 0;
    for (ll i = row_begin + warp; i < row_end; i += nwarps) {
        if (!core[i]) continue;
        ll b = row_ptr[i];
        int m = (int)(row_ptr[i + 1] - b);
        const int *row = col_idx + b;
        int ri = 0;
        if (lane == 0) ri = uf_find(parent, (int)i);
        ri = __shfl_sync(0xffffffffu, ri, 0);
        for (int base = 0; base < m; base += 32) {
            int k = base + lane;
            int j = k < m ? row[k] : -1;
            bool cand = j > (int)i && core[j];
            bool need = cand && uf_find_ro(parent, j) != ri;
            unsigned cm = __ballot_sync(0xffffffffu, cand);
            unsigned nm = __ballot_sync(0xffffffffu, need);
            n_edges += __popc(cm);
            n_skip += __popc(cm & ~nm);
            while (nm) {
                int src = __ffs(nm) - 1;
                nm &= nm - 1;
                int jj = __shfl_sync(0xffffffffu, j, src);
                int same = 0;
                if (lane == 0) {
                    ri = uf_find(parent, (int)i);
                    same = uf_find(parent, jj) == ri;
                }
                same = __shfl_sync(0xffffffffu, same, 0);
                if (same) { ++n_skip; continue; }
                ll bj = row_ptr[jj];
                int mj = (int)(row_ptr[jj + 1] - bj);
                ++n_checks;
                if (warp_common<true>(row, m, col_idx + bj, mj, cutoff, lane)) {
                    if (lane == 0) uf_union(parent, (int)i, jj);
                    ++n_acc;
                }
            }
            if (lane == 0) ri = uf_find(parent, (int)i);
            ri = __shfl_sync(0xffffffffu, ri, 0);
        }
    }
    if (lane == 0) {
        if (n_edges) atomicAdd(stats + ST_EDGES, n_edges);
        if (n_checks) atomicAdd(stats + ST_CHECKS, n_checks);
        if (n_acc) atomicAdd(stats + ST_ACCEPTED, n_acc);
        if (n_skip) atomicAdd(stats + ST_SKIPPED, n_skip);
    }
}

// merge step of the row-sharded multi-GPU run: union (i, other[i]) for each foreign parent array
__global__ void comp_merge_kernel(const int *__restrict__ others, ll n, ll count, int *parent) {
    ll t = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    ll stride = (ll)gridDim.x * blockDim.x;
    for (; t < n * count; t += stride) {
        int i = (int)(t % n);
        int o = others[t];
        if (o != i) uf_union(parent, i, o);
    }
}

// ---- count mode -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cn_count_kernel(const ll *__restrict__ row_ptr, const int *__restrict__ col_idx, ll n, ll n_member_cutoff,
                unsigned *__restrict__ counts, unsigned long long *stats) {
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    unsigned long long n_edges = 0;
    for (ll i = warp; i < n; i += nwarps) {
        ll b = row_ptr[i];
        int m = (int)(row_ptr[i + 1] - b);
        const int *row = col_idx + b;
        bool core_i = m > n_member_cutoff;
        for (int k = 0; k < m; ++k) {
            int j = row[k];
            unsigned out = 0xffffffffu;
            if (core_i && j > (int)i) {
                ll bj = row_ptr[j];
                int mj = (int)(row_ptr[j + 1] - bj);
                if (mj > n_member_cutoff) {
                    out = (unsigned)warp_common<false>(row, m, col_idx + bj, mj, 0, lane);
                    ++n_edges;
                }
            }
            if (lane == 0) counts[b + k] = out;
        }
    }
    if (lane == 0 && n_edges) atomicAdd(stats + ST_EDGES, n_edges);
}

// ---- host side ----------------------------------------------------------------------------------
static int fetch_stats(cnnb200_ctx *ctx) {
    unsigned long long *h;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&h));
    CNN_CUDA(cudaMemcpyAsync(h, ctx->b_stats.p, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CNN_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 4; ++k) ctx->stats[k] = (ll)h[k];
    return CNNB200_OK;
}

int cnn_fetch_comp_stats(cnnb200_ctx *ctx) { return fetch_stats(ctx); }

extern "C" int cnnb200_components_begin(cnnb200_ctx *ctx, int64_t n_member_cutoff, int64_t similarity_cutoff,
                                        const uint8_t *consider, int loc) {
    if (!ctx) CNN_FAIL(CNNB200_E_INVALID, "ctx is NULL");
    if (!ctx->have_nb) CNN_FAIL(CNNB200_E_STATE, "components: no neighbourhoods held by this context");
    if (n_member_cutoff < 0 || similarity_cutoff < 0)
        CNN_FAIL(CNNB200_E_INVALID, "components: cutoffs must be >= 0");
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ll n = ctx->n;
    CNN_TRY(cnn_ensure(ctx, ctx->b_parent, sizeof(int) * (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_core, (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_stats, sizeof(unsigned long long) * ST_N));
    CNN_CUDA(cudaMemsetAsync(ctx->b_stats.p, 0, sizeof(unsigned long long) * ST_N, st));
    const uint8_t *d_consider = consider;
    if (consider && loc == CNNB200_HOST) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_consider, (size_t)n));
        CNN_CUDA(cudaMemcpyAsync(ctx->b_consider.p, consider, (size_t)n, cudaMemcpyHostToDevice, st));
        d_consider = ctx->b_consider.as<uint8_t>();
    }
    {
        Span sp(ctx, T_COMP_INIT);
        comp_init_kernel<<<cnn_blocks(n, 256), 256, 0, st>>>(ctx->row_ptr, ctx->perm, d_consider, n,
                                                              n_member_cutoff, ctx->b_parent.as<int>(),
                                                              ctx->b_core.as<uint8_t>());
        CNN_LAUNCHED(ctx, 1);
    }
    CNN_KERNEL_CHECK();
    ctx->n_member_cutoff = n_member_cutoff;
    ctx->similarity_cutoff = similarity_cutoff;
    ctx->have_comp = true;
    return CNNB200_OK;
}

extern "C" int cnnb200_components_edges(cnnb200_ctx *ctx, int64_t row_begin, int64_t row_end) {
    if (!ctx) CNN_FAIL(CNNB200_E_INVALID, "ctx is NULL");
    if (!ctx->have_comp) CNN_FAIL(CNNB200_E_STATE, "components_edges: call components_begin first");
    if (row_begin < 0 || row_end > ctx->n || row_begin > row_end)
        CNN_FAIL(CNNB200_E_INVALID, "components_edges: bad row range");
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ll rows = row_end - row_begin;
    if (rows == 0) return CNNB200_OK;
    // a similarity cutoff beyond any possible intersection size links nothing
    int cutoff = ctx->similarity_cutoff > 0x7fffffffll ? 0x7fffffff : (int)ctx->similarity_cutoff;
    int blocks = (int)((rows + 7) / 8);
    int cap = ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    unsigned long long *stats = (unsigned long long *)ctx->b_stats.p;
    {
        Span sp(ctx, T_COMP_SAMPLE);
        comp_sample_kernel<<<blocks, 256, 0, st>>>(ctx->row_ptr, ctx->col_idx, ctx->b_core.as<uint8_t>(), row_begin,
                                                   row_end, cutoff, ctx->b_parent.as<int>(), stats);
    }
    {
        Span sp(ctx, T_COMP_FLATTEN);
        comp_compress_kernel<<<cnn_blocks(ctx->n, 256), 256, 0, st>>>(ctx->b_parent.as<int>(), ctx->n);
    }
    {
        Span sp(ctx, T_COMP_EDGES);
        comp_edges_kernel<<<blocks, 256, 0, st>>>(ctx->row_ptr, ctx->col_idx, ctx->b_core.as<uint8_t>(), row_begin,
                                                  row_end, cutoff, ctx->b_parent.as<int>(), stats);
    }
    {
        Span sp(ctx, T_COMP_FLATTEN);
        comp_compress_kernel<<<cnn_blocks(ctx->n, 256), 256, 0, st>>>(ctx->b_parent.as<int>(), ctx->n);
    }
    CNN_LAUNCHED(ctx, 4);
    CNN_KERNEL_CHECK();
    return CNNB200_OK;
}

extern "C" int cnnb200_components_parents(cnnb200_ctx *ctx, int32_t **parents) {
    if (!ctx || !parents) CNN_FAIL(CNNB200_E_INVALID, "components_parents: NULL argument");
    if (!ctx->have_comp) CNN_FAIL(CNNB200_E_STATE, "components_parents: call components_begin first");
    *parents = ctx->b_parent.as<int>();
    return CNNB200_OK;
}

extern "C" int cnnb200_components_merge(cnnb200_ctx *ctx, const int32_t *others, int64_t count) {
    if (!ctx || !others) CNN_FAIL(CNNB200_E_INVALID, "components_merge: NULL argument");
    if (!ctx->have_comp) CNN_FAIL(CNNB200_E_STATE, "components_merge: call components_begin first");
    if (count < 1) return CNNB200_OK;
    CNN_CUDA(cudaSetDevice(ctx->device));
    Span sp(ctx, T_COMP_FLATTEN);
    CNN_LAUNCHED(ctx, 2);
    comp_merge_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(others, ctx->n, count, ctx->b_parent.as<int>());
    comp_compress_kernel<<<cnn_blocks(ctx->n, 256), 256, 0, ctx->stream>>>(ctx->b_parent.as<int>(), ctx->n);
    CNN_KERNEL_CHECK();
    return CNNB200_OK;
}

extern "C" int cnnb200_cn_count(cnnb200_ctx *ctx, int64_t n_member_cutoff, uint32_t *counts, int loc,
                                int64_t *n_edges_out) {
    if (!ctx || !counts) CNN_FAIL(CNNB200_E_INVALID, "cn_count: NULL argument");
    if (!ctx->have_nb) CNN_FAIL(CNNB200_E_STATE, "cn_count: no neighbourhoods held by this context");
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ll n = ctx->n, nnz = ctx->nnz;
    CNN_TRY(cnn_ensure(ctx, ctx->b_stats, sizeof(unsigned long long) * ST_N));
    CNN_CUDA(cudaMemsetAsync(ctx->b_stats.p, 0, sizeof(unsigned long long) * ST_N, st));
    unsigned *d_counts = counts;
    if (loc == CNNB200_HOST) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_tmp0, sizeof(unsigned) * (size_t)nnz));
        d_counts = ctx->b_tmp0.as<unsigned>();
    }
    int blocks = (int)((n + 7) / 8);
    int cap = ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    {
        Span sp(ctx, T_CN_COUNT);
        cn_count_kernel<<<blocks, 256, 0, st>>>(ctx->row_ptr, ctx->col_idx, n, n_member_cutoff, d_counts,
                                                (unsigned long long *)ctx->b_stats.p);
        CNN_LAUNCHED(ctx, 1);
    }
    CNN_KERNEL_CHECK();
    if (loc == CNNB200_HOST)
        CNN_CUDA(cudaMemcpyAsync(counts, d_counts, sizeof(unsigned) * (size_t)nnz, cudaMemcpyDeviceToHost, st));
    if (loc == CNNB200_HOST || n_edges_out) {
        CNN_TRY(fetch_stats(ctx));
        if (n_edges_out) *n_edges_out = ctx->stats[ST_EDGES];
    }
    return CNNB200_OK;
}
