// labels.cu -- K7: cluster numbering as FitterExtBFS produces it and Labels.sort_by_size.
//
// (1) Numbering.  The reference's outer loop (_fit.pyx:434-491) visits points in ascending index
//     and opens a new cluster number whenever it meets an unassigned core point, so cluster k
//     (counting from current_start) is the component with the k-th smallest minimum member
//     index.  Here: per component the smallest caller-space member index (atomicMin through
//     perm[]), a flag per such index, an exclusive scan of the flags = the rank.
// (2) Labels.sort_by_size (_types.pyx:212-286): Counter(labels).most_common() is a stable sort by
//     count descending over first-occurrence order; rank > max_clusters -> 0, size < member_cutoff
//     -> 0, survivors renumbered 1, 2, ...  Here: per-label count and first occurrence by
//     warp-aggregated atomics, one 64-bit key (inverted count : first index) per label, radix
//     sort, survivor scan, gather through the map.
#include <cub/cub.cuh>
#include <string.h>

#include <memory>

#include "common.cuh"

int cnn_fetch_comp_stats(cnnb200_ctx *ctx);

__global__ void fill_int_kernel(int *a, ll n, int v) {
    ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

__global__ void minorig_kernel(const int *__restrict__ parent, const uint8_t *__restrict__ core,
                               const int *__restrict__ perm, ll n, int *minorig) {
    ll p = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    bool active = p < n && core[p];
    unsigned am = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    // one atomic per distinct root per warp (a giant cluster would otherwise serialise on one address)
    int r = parent[p];
    unsigned peers = __match_any_sync(am, r);
    int mn = __reduce_min_sync(peers, perm[p]);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicMin(minorig + r, mn);
}

// flags[min caller index of each component] = 1
__global__ void root_flag_kernel(const int *__restrict__ parent, const uint8_t *__restrict__ core,
                                 const int *__restrict__ minorig, ll n, int *flags) {
    ll p = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (p >= n || !core[p] || parent[p] != (int)p) return;
    flags[minorig ? minorig[p] : (int)p] = 1;
}

__global__ void write_labels_kernel(const int *__restrict__ parent, const uint8_t *__restrict__ core,
                                    const int *__restrict__ perm, const int *__restrict__ minorig,
                                    const int *__restrict__ rank, ll n, ll current_start, ll *labels,
                                    int fresh) {
    ll p = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (p >= n) return;
    ll o = perm ? perm[p] : p;
    if (core[p]) {
        int r = parent[p];
        labels[o] = current_start + rank[minorig ? minorig[r] : r];
    } else if (fresh) {
        labels[o] = 0;
    }
}

extern "C" int cnnb200_components_labels(cnnb200_ctx *ctx, int64_t current_start, int64_t *labels, int loc,
                                         int fresh, int64_t *n_clusters_out) {
    if (!ctx || !labels) CNN_FAIL(CNNB200_E_INVALID, "components_labels: NULL argument");
    if (!ctx->have_comp) CNN_FAIL(CNNB200_E_STATE, "components_labels: call components_begin first");
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ll n = ctx->n;
    const int *parent = ctx->b_parent.as<int>();
    const uint8_t *core = ctx->b_core.as<uint8_t>();
    int nb = cnn_blocks(n, 256);
    std::unique_ptr<Span> sp(new Span(ctx, T_LABELS));
    CNN_LAUNCHED(ctx, ctx->perm ? 4 : 2);
    int *minorig = nullptr;
    if (ctx->perm) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_minorig, sizeof(int) * (size_t)n));
        minorig = ctx->b_minorig.as<int>();
        fill_int_kernel<<<nb, 256, 0, st>>>(minorig, n, 0x7fffffff);
        minorig_kernel<<<nb, 256, 0, st>>>(parent, core, ctx->perm, n, minorig);
    }
    CNN_TRY(cnn_ensure(ctx, ctx->b_flags, sizeof(int) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_rank, sizeof(int) * (size_t)(n + 1)));
    CNN_CUDA(cudaMemsetAsync(ctx->b_flags.p, 0, sizeof(int) * (size_t)(n + 1), st));
    root_flag_kernel<<<nb, 256, 0, st>>>(parent, core, minorig, n, ctx->b_flags.as<int>());
    CNN_KERNEL_CHECK();
    size_t need = 0;
    CNN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, ctx->b_flags.as<int>(), ctx->b_rank.as<int>(),
                                           (int)(n + 1), st));
    CNN_TRY(cnn_ensure(ctx, ctx->b_cub, need));
    size_t cap = ctx->b_cub.cap;
    CNN_CUDA(cub::DeviceScan::ExclusiveSum(ctx->b_cub.p, cap, ctx->b_flags.as<int>(), ctx->b_rank.as<int>(),
                                           (int)(n + 1), st));
    ll *d_labels = labels;
    if (loc == CNNB200_HOST) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_labels, sizeof(ll) * (size_t)n));
        d_labels = ctx->b_labels.as<ll>();
        if (!fresh) CNN_CUDA(cudaMemcpyAsync(d_labels, labels, sizeof(ll) * (size_t)n, cudaMemcpyHostToDevice, st));
    }
    write_labels_kernel<<<nb, 256, 0, st>>>(parent, core, ctx->perm, minorig, ctx->b_rank.as<int>(), n,
                                            current_start, d_labels, fresh);
    CNN_KERNEL_CHECK();
    sp.reset();
    if (loc == CNNB200_HOST) {
        Span sd(ctx, T_D2H);
        CNN_CUDA(cudaMemcpyAsync(labels, d_labels, sizeof(ll) * (size_t)n, cudaMemcpyDeviceToHost, st));
    }
    if (n_clusters_out || loc == CNNB200_HOST) {
        int *h;
        CNN_TRY(cnn_pinned(ctx, 128, (void **)&h));
        CNN_CUDA(cudaMemcpyAsync(h + 24, ctx->b_rank.as<int>() + n, sizeof(int), cudaMemcpyDeviceToHost, st));
        CNN_TRY(cnn_fetch_comp_stats(ctx));  // synchronises
        if (n_clusters_out) *n_clusters_out = h[24];
    }
    return CNNB200_OK;
}

extern "C" int cnnb200_fit(cnnb200_ctx *ctx, int64_t n_member_cutoff, int64_t similarity_cutoff,
                           int64_t current_start, int64_t *labels, uint8_t *consider, int loc,
                           int64_t *n_clusters_out) {
    if (!ctx || !labels) CNN_FAIL(CNNB200_E_INVALID, "fit: NULL argument");
    if (!ctx->have_nb) CNN_FAIL(CNNB200_E_STATE, "fit: no neighbourhoods held by this context");
    CNN_TRY(cnnb200_components_begin(ctx, n_member_cutoff, similarity_cutoff, consider, loc));
    CNN_TRY(cnnb200_components_edges(ctx, 0, ctx->n));
    int fresh = (consider == nullptr && loc == CNNB200_HOST) ? 1 : 0;
    CNN_TRY(cnnb200_components_labels(ctx, current_start, labels, loc, fresh, n_clusters_out));
    if (consider) {  // every point has been visited: consider[:] = 0 (_fit.pyx:434-437)
        if (loc == CNNB200_HOST)
            memset(consider, 0, (size_t)ctx->n);
        else
            CNN_CUDA(cudaMemsetAsync(consider, 0, (size_t)ctx->n, ctx->stream));
    }
    return CNNB200_OK;
}

// ---- Labels.sort_by_size ------------------------------------------------------------------------
__global__ void label_hist_kernel(const ll *__restrict__ labels, ll n, int *count, int *first, int *flag,
                                  ll max_label) {
    ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    bool active = i < n;
    ll l = active ? labels[i] : 0;
    if (active && (l < 0 || l > max_label)) { *flag = 1; active = false; }
    active = active && l > 0;
    // warp-aggregate equal labels: one atomic per distinct label per warp
    unsigned am = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    unsigned peers = __match_any_sync(am, l);
    int leader = __ffs(peers) - 1;
    if ((int)(threadIdx.x & 31) == leader) {
        atomicAdd(count + l, __popc(peers));
        atomicMin(first + l, (int)i);  // lanes ascend with i: the leader holds the smallest index
    }
}

// key = (n - size) : first occurrence, kb bits each (n < 2^kb): ascending keys = size descending, ties by first
// occurrence; labels that do not occur get the largest key.  Only 2*kb bits are sorted.
__global__ void label_keys_kernel(const int *__restrict__ count, const int *__restrict__ first, ll L, ll n, int kb,
                                  unsigned long long *keys, int *vals) {
    ll l = blockIdx.x * (ll)blockDim.x + threadIdx.x + 1;  // labels 1 .. L-1
    if (l >= L) return;
    int c = count[l];
    keys[l - 1] = ((unsigned long long)(n - c) << kb) | (unsigned long long)(c > 0 ? first[l] : 0);
    vals[l - 1] = (int)l;
}

// survivors in rank order: t-th entry (rank t+1) survives if it exists, rank <= max_clusters and
// size >= member_cutoff
__global__ void label_survivor_kernel(const unsigned long long *__restrict__ keys, ll m, ll n, int kb,
                                      ll member_cutoff, ll max_clusters, int *surv) {
    ll t = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (t > m) return;
    if (t == m) { surv[m] = 0; return; }
    const ll size = n - (ll)(keys[t] >> kb);
    bool ok = size > 0;
    ok = ok && (max_clusters < 0 || t + 1 <= max_clusters) && size >= member_cutoff;
    surv[t] = ok ? 1 : 0;
}

// what Clustering.fit records about the relabelled result (cluster.pyx:366-390): clusters kept, members of the
// largest one, noise points -- from the sorted sizes, so that the host needs no pass over the labels
__global__ void label_summary_kernel(const unsigned long long *__restrict__ keys, const int *__restrict__ surv, ll m, ll n,
                                     int kb, unsigned long long *out) {
    // out[0] += clusters kept, out[1] = max size, out[2] += members of kept clusters (noise = n - out[2], on the host)
    ll sum = 0, mx = 0, cnt = 0;
    for (ll t = blockIdx.x * (ll)blockDim.x + threadIdx.x; t < m; t += (ll)gridDim.x * blockDim.x)
        if (surv[t]) {
            const ll size = n - (ll)(keys[t] >> kb);
            sum += size;
            mx = size > mx ? size : mx;
            ++cnt;
        }
    for (int o = 16; o; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        const ll om = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = om > mx ? om : mx;
    }
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicAdd(out + 0, (unsigned long long)cnt);
        atomicMax(out + 1, (unsigned long long)mx);
        atomicAdd(out + 2, (unsigned long long)sum);
    }
}

__global__ void label_map_kernel(const int *__restrict__ vals, const int *__restrict__ surv,
                                 const int *__restrict__ surv_rank, ll m, ll *map) {
    ll t = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (t >= m) return;
    map[vals[t]] = surv[t] ? (ll)surv_rank[t] + 1 : 0;
}

__global__ void label_apply_kernel(ll *labels, ll n, const ll *__restrict__ map, ll max_label) {
    ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (i >= n) return;
    ll l = labels[i];
    if (l >= 0 && l <= max_label) labels[i] = map[l];  // negative labels are reported through the flag
}

extern "C" int cnnb200_sort_by_size(cnnb200_ctx *ctx, int64_t *labels, int loc, int64_t n,
                                    int64_t member_cutoff, int64_t max_clusters, int64_t *map_out,
                                    int64_t map_len) {
    if (!ctx || !labels) CNN_FAIL(CNNB200_E_INVALID, "sort_by_size: NULL argument");
    if (n < 0 || n >= (1ll << 31) - 1) CNN_FAIL(CNNB200_E_RANGE, "sort_by_size: n out of range");
    if (n == 0) return CNNB200_OK;
    if (member_cutoff < 0) member_cutoff = 2;  // None -> 2 (_types.pyx:235-237)
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ll *d_labels = labels;
    if (loc == CNNB200_HOST) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_labels, sizeof(ll) * (size_t)n));
        d_labels = ctx->b_labels.as<ll>();
        Span sh(ctx, T_H2D);
        CNN_CUDA(cudaMemcpyAsync(d_labels, labels, sizeof(ll) * (size_t)n, cudaMemcpyHostToDevice, st));
    }
    ctx->have_lab_summary = false;
    ctx->lab_summary_cached = false;
    ctx->lab_summary_n = n;
    std::unique_ptr<Span> sp(new Span(ctx, T_SORT_BY_SIZE));
    // largest label -> table size (known to the caller after a fresh fit: ctx->known_max_label)
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp2, 64));
    size_t cap = 0;
    ll *h;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&h));
    ll max_label = ctx->known_max_label;
    ctx->known_max_label = -1;
    if (max_label < 0) {
        ll *d_max = ctx->b_tmp2.as<ll>() + 2;
        size_t need = 0;
        CNN_CUDA(cub::DeviceReduce::Max(nullptr, need, d_labels, d_max, (int)n, st));
        CNN_TRY(cnn_ensure(ctx, ctx->b_cub, need));
        cap = ctx->b_cub.cap;
        CNN_CUDA(cub::DeviceReduce::Max(ctx->b_cub.p, cap, d_labels, d_max, (int)n, st));
        CNN_CUDA(cudaMemcpyAsync(h, d_max, sizeof(ll), cudaMemcpyDeviceToHost, st));
        CNN_CUDA(cudaStreamSynchronize(st));
        max_label = h[0];
    }
    if (max_label < 0) CNN_FAIL(CNNB200_E_INVALID, "sort_by_size: negative labels");
    if (max_label >= (1ll << 31) - 2) CNN_FAIL(CNNB200_E_RANGE, "sort_by_size: label %lld too large", max_label);
    ll L = max_label + 1, m = L - 1;
    if (m > 0) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_lab_count, sizeof(int) * (size_t)L));
        CNN_TRY(cnn_ensure(ctx, ctx->b_lab_first, sizeof(int) * (size_t)L));
        CNN_TRY(cnn_ensure(ctx, ctx->b_lab_keys, sizeof(ll) * (size_t)m));
        CNN_TRY(cnn_ensure(ctx, ctx->b_lab_keys2, sizeof(ll) * (size_t)m));
        CNN_TRY(cnn_ensure(ctx, ctx->b_lab_vals, sizeof(int) * (size_t)(m + 1)));
        CNN_TRY(cnn_ensure(ctx, ctx->b_lab_vals2, sizeof(int) * (size_t)(m + 1)));
        CNN_TRY(cnn_ensure(ctx, ctx->b_lab_map, sizeof(ll) * (size_t)L));
        CNN_TRY(cnn_ensure(ctx, ctx->b_flags, sizeof(int) * (size_t)(m + 1)));
        CNN_TRY(cnn_ensure(ctx, ctx->b_rank, sizeof(int) * (size_t)(m + 1)));
        int *count = ctx->b_lab_count.as<int>(), *first = ctx->b_lab_first.as<int>();
        int *flag = ctx->b_tmp2.as<int>();
        CNN_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
        CNN_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)L, st));
        fill_int_kernel<<<cnn_blocks(L, 256), 256, 0, st>>>(first, L, 0x7fffffff);
        label_hist_kernel<<<cnn_blocks(n, 256), 256, 0, st>>>(d_labels, n, count, first, flag, max_label);
        unsigned long long *keys = ctx->b_lab_keys.as<unsigned long long>();
        unsigned long long *keys2 = ctx->b_lab_keys2.as<unsigned long long>();
        int *vals = ctx->b_lab_vals.as<int>(), *vals2 = ctx->b_lab_vals2.as<int>();
        int kb = 1;
        while ((1ll << kb) <= n) ++kb;
        label_keys_kernel<<<cnn_blocks(m, 256), 256, 0, st>>>(count, first, L, n, kb, keys, vals);
        CNN_KERNEL_CHECK();
        size_t need_sort = 0, need_scan = 0;
        CNN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need_sort, keys, keys2, vals, vals2, (int)m, 0, 2 * kb, st));
        CNN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need_scan, ctx->b_flags.as<int>(), ctx->b_rank.as<int>(),
                                               (int)(m + 1), st));
        CNN_TRY(cnn_ensure(ctx, ctx->b_cub, need_sort > need_scan ? need_sort : need_scan));
        cap = ctx->b_cub.cap;
        CNN_CUDA(cub::DeviceRadixSort::SortPairs(ctx->b_cub.p, cap, keys, keys2, vals, vals2, (int)m, 0, 2 * kb, st));
        label_survivor_kernel<<<cnn_blocks(m + 1, 256), 256, 0, st>>>(keys2, m, n, kb, member_cutoff, max_clusters,
                                                                        ctx->b_flags.as<int>());
        cap = ctx->b_cub.cap;
        CNN_CUDA(cub::DeviceScan::ExclusiveSum(ctx->b_cub.p, cap, ctx->b_flags.as<int>(), ctx->b_rank.as<int>(),
                                               (int)(m + 1), st));
        ll *map = ctx->b_lab_map.as<ll>();
        CNN_CUDA(cudaMemsetAsync(map, 0, sizeof(ll), st));
        label_map_kernel<<<cnn_blocks(m, 256), 256, 0, st>>>(vals2, ctx->b_flags.as<int>(), ctx->b_rank.as<int>(), m, map);
        label_apply_kernel<<<cnn_blocks(n, 256), 256, 0, st>>>(d_labels, n, map, max_label);
        CNN_TRY(cnn_ensure(ctx, ctx->b_lab_summary, sizeof(ll) * 4));
        CNN_CUDA(cudaMemsetAsync(ctx->b_lab_summary.p, 0, sizeof(ll) * 4, st));
        label_summary_kernel<<<cnn_blocks(m, 256) < 64 ? cnn_blocks(m, 256) : 64, 256, 0, st>>>(
            keys2, ctx->b_flags.as<int>(), m, n, kb, ctx->b_lab_summary.as<unsigned long long>());
        ctx->have_lab_summary = true;
        CNN_LAUNCHED(ctx, 7);
        CNN_KERNEL_CHECK();
        if (map_out && map_len > 0) {
            ll cnt = map_len < L ? map_len : L;
            CNN_CUDA(cudaMemcpyAsync(map_out, map, sizeof(ll) * (size_t)cnt, cudaMemcpyDeviceToHost, st));
            for (ll k = cnt; k < map_len; ++k) map_out[k] = 0;
        }
    } else if (map_out && map_len > 0) {
        for (ll k = 0; k < map_len; ++k) map_out[k] = 0;
    }
    sp.reset();
    if (loc == CNNB200_HOST) {
        Span sd(ctx, T_D2H);
        CNN_CUDA(cudaMemcpyAsync(labels, d_labels, sizeof(ll) * (size_t)n, cudaMemcpyDeviceToHost, st));
    }
    if (loc == CNNB200_HOST || map_out) {
        int *hf = (int *)h + 8;
        *hf = 0;
        if (m > 0) CNN_CUDA(cudaMemcpyAsync(hf, ctx->b_tmp2.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        CNN_CUDA(cudaStreamSynchronize(st));
        if (*hf) CNN_FAIL(CNNB200_E_INVALID, "sort_by_size: negative labels");
    }
    return CNNB200_OK;
}

// (clusters kept, members of the largest cluster, noise points) of the last cnnb200_sort_by_size / cnnb200_fit_points
extern "C" int cnnb200_label_summary(cnnb200_ctx *ctx, int64_t *out3) {
    if (!ctx || !out3) CNN_FAIL(CNNB200_E_INVALID, "label_summary: NULL argument");
    CNN_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->have_lab_summary) {  // only noise (or nothing relabelled yet)
        out3[0] = 0; out3[1] = 0; out3[2] = ctx->lab_summary_n;
        return CNNB200_OK;
    }
    if (ctx->lab_summary_cached) {  // cnnb200_fit_points read them with its final synchronisation
        for (int k = 0; k < 3; ++k) out3[k] = ctx->lab_summary_host[k];
        return CNNB200_OK;
    }
    ll *h;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&h));
    CNN_CUDA(cudaMemcpyAsync(h, ctx->b_lab_summary.p, sizeof(ll) * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CNN_CUDA(cudaStreamSynchronize(ctx->stream));
    out3[0] = h[0]; out3[1] = h[1]; out3[2] = ctx->lab_summary_n - h[2];
    return CNNB200_OK;
}

// ---- fresh fit + sort_by_size in one go (cnnb200_fit_points) ------------------------------------------------------
// After a fresh fit every label comes from this fit, so FitterExtBFS's numbering followed by Labels.sort_by_size
// (_fit.pyx:434-491, _types.pyx:212-286) collapses to: rank the components by (size descending, smallest member index
// ascending -- the first occurrence of the BFS label), drop those below member_cutoff / beyond max_clusters, number the
// rest 1, 2, ...  Components below member_cutoff never influence the rank of the others (they sort behind every
// survivor), so only the survivors are sorted: a few hundred keys on C2 instead of 48 751, in ONE CTA.
//   fused_sizes_kernel    size and smallest caller index per root (warp-aggregated atomics)
//   fused_roots_kernel    roots -> compact list of survivor keys (size : min index), counts all components
//   fused_rank_kernel     one CTA: bitonic sort of up to FUSED_MAX_KEYS keys in shared memory, rank -> root_label[],
//                         summary (clusters kept, largest, members kept); more keys: flag, the caller falls back
//   fused_write_kernel    labels[perm[p]] = root_label[parent[p]]
#define FUSED_MAX_KEYS 16384   // 12 bytes of (dynamic) shared memory per key: 192 KB

__global__ void fused_sizes_kernel(const int *__restrict__ parent, const int *__restrict__ perm, ll n, int *size, int *minorig) {
    ll p = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    const int r = p < n ? parent[p] : -1;
    const bool active = r >= 0;
    const unsigned am = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    const unsigned peers = __match_any_sync(am, r);
    const int mn = __reduce_min_sync(peers, perm ? perm[p] : (int)p);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) {
        atomicAdd(size + r, __popc(peers));
        atomicMin(minorig + r, mn);
    }
}

// counters: [0] components, [1] survivor keys, [2] overflow flag, [3] kept, [4] largest, [5] members kept
__global__ void fused_roots_kernel(const int *__restrict__ parent, const int *__restrict__ size, const int *__restrict__ minorig,
                                   ll n, int kb, ll member_cutoff, unsigned long long *keys, int *vals, int *root_label,
                                   unsigned long long *counters) {
    ll p = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    const bool root = p < n && parent[p] == (int)p;
    const int sz = root ? size[p] : 0;
    const bool surv = root && sz >= member_cutoff;
    const unsigned rm = __ballot_sync(0xffffffffu, root), sm = __ballot_sync(0xffffffffu, surv);
    const int lane = threadIdx.x & 31;
    if (!rm) return;
    unsigned long long base = 0;
    if (lane == 0) {
        atomicAdd(counters + 0, (unsigned long long)__popc(rm));
        if (sm) base = atomicAdd(counters + 1, (unsigned long long)__popc(sm));
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (root) root_label[p] = 0;
    if (surv) {
        const unsigned long long k = base + __popc(sm & ((1u << lane) - 1u));
        if (k < FUSED_MAX_KEYS) {
            keys[k] = ((unsigned long long)(n - sz) << kb) | (unsigned long long)minorig[p];
            vals[k] = (int)p;
        }
    }
}

__global__ void __launch_bounds__(1024)
fused_rank_kernel(const unsigned long long *__restrict__ keys, const int *__restrict__ vals, ll n, int kb, ll max_clusters,
                  int *root_label, unsigned long long *counters) {
    extern __shared__ unsigned long long s_fused[];
    unsigned long long *s_key = s_fused;
    int *s_val = (int *)(s_fused + FUSED_MAX_KEYS);
    const ll m = (ll)counters[1];
    if (m > FUSED_MAX_KEYS) {
        if (threadIdx.x == 0) counters[2] = 1;
        return;
    }
    int np2 = 1;
    while (np2 < m) np2 <<= 1;
    for (int t = threadIdx.x; t < np2; t += blockDim.x) {
        s_key[t] = t < m ? keys[t] : ~0ull;
        s_val[t] = t < m ? vals[t] : -1;
    }
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < np2; t += blockDim.x) {
                const int u = t ^ j;
                if (u > t) {
                    const bool up = (t & k) == 0;
                    const unsigned long long a = s_key[t], b = s_key[u];
                    if ((a > b) == up) {
                        s_key[t] = b; s_key[u] = a;
                        const int v = s_val[t]; s_val[t] = s_val[u]; s_val[u] = v;
                    }
                }
            }
            __syncthreads();
        }
    // keys are distinct (the smallest member index is): ascending keys = size descending, ties by first occurrence
    const ll kept = max_clusters >= 0 && max_clusters < m ? max_clusters : m;
    unsigned long long members = 0;
    for (int t = threadIdx.x; t < m; t += blockDim.x) {
        root_label[s_val[t]] = t < kept ? t + 1 : 0;
        if (t < kept) members += (unsigned long long)(n - (ll)(s_key[t] >> kb));
    }
    for (int o = 16; o; o >>= 1) members += __shfl_xor_sync(0xffffffffu, members, o);
    if ((threadIdx.x & 31) == 0 && members) atomicAdd(counters + 5, members);
    if (threadIdx.x == 0) {
        counters[3] = (unsigned long long)kept;
        counters[4] = kept > 0 ? (unsigned long long)(n - (ll)(s_key[0] >> kb)) : 0ull;
    }
}

__global__ void fused_write_kernel(const int *__restrict__ parent, const int *__restrict__ perm, const int *__restrict__ root_label,
                                   ll n, ll *labels) {
    ll p = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int r = parent[p];
    labels[perm ? perm[p] : p] = r >= 0 ? (ll)root_label[r] : 0;
}

// Enqueues the four kernels; `counters_out` (device, 6 words) is read by the caller together with its final
// synchronisation: [0] number of components = clusters the BFS would have numbered, [2] != 0: more survivors than one
// CTA sorts -- the caller repeats with cnnb200_components_labels + cnnb200_sort_by_size.
int cnn_labels_fused(cnnb200_ctx *ctx, ll member_cutoff, ll max_clusters, ll *d_labels, unsigned long long **counters_out) {
    cudaStream_t st = ctx->stream;
    const ll n = ctx->n;
    if (member_cutoff < 0) member_cutoff = 2;  // None -> 2 (_types.pyx:235-237)
    const int *parent = ctx->b_parent.as<int>();
    Span sp(ctx, T_LABELS);
    CNN_TRY(cnn_ensure(ctx, ctx->b_minorig, sizeof(int) * (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_lab_count, sizeof(int) * (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_rank, sizeof(int) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_lab_keys, sizeof(ll) * FUSED_MAX_KEYS));
    CNN_TRY(cnn_ensure(ctx, ctx->b_lab_vals, sizeof(int) * FUSED_MAX_KEYS));
    CNN_TRY(cnn_ensure(ctx, ctx->b_lab_summary, sizeof(ll) * 8));
    int *size = ctx->b_lab_count.as<int>(), *minorig = ctx->b_minorig.as<int>(), *root_label = ctx->b_rank.as<int>();
    unsigned long long *counters = ctx->b_lab_summary.as<unsigned long long>();
    CNN_CUDA(cudaMemsetAsync(size, 0, sizeof(int) * (size_t)n, st));
    CNN_CUDA(cudaMemsetAsync(minorig, 0x7f, sizeof(int) * (size_t)n, st));
    CNN_CUDA(cudaMemsetAsync(counters, 0, sizeof(ll) * 8, st));
    int kb = 1;
    while ((1ll << kb) <= n) ++kb;
    const int nb = cnn_blocks(n, 256);
    fused_sizes_kernel<<<nb, 256, 0, st>>>(parent, ctx->perm, n, size, minorig);
    fused_roots_kernel<<<nb, 256, 0, st>>>(parent, size, minorig, n, kb, member_cutoff, ctx->b_lab_keys.as<unsigned long long>(),
                                           ctx->b_lab_vals.as<int>(), root_label, counters);
    static const bool smem_ok = cudaFuncSetAttribute(fused_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     FUSED_MAX_KEYS * 12) == cudaSuccess;
    if (!smem_ok) CNN_FAIL(CNNB200_E_CUDA, "fused ranking: cannot opt in to %d bytes of shared memory", FUSED_MAX_KEYS * 12);
    fused_rank_kernel<<<1, 1024, FUSED_MAX_KEYS * 12, st>>>(ctx->b_lab_keys.as<unsigned long long>(), ctx->b_lab_vals.as<int>(), n, kb,
                                                           max_clusters, root_label, counters);
    fused_write_kernel<<<nb, 256, 0, st>>>(parent, ctx->perm, root_label, n, d_labels);
    CNN_LAUNCHED(ctx, 4);
    CNN_KERNEL_CHECK();
    *counters_out = counters;
    return CNNB200_OK;
}

// ---- (previous label, current label) grouping for the hierarchy driver --------------------------------------------------
// HierarchicalFitterRepeat._fit (_fit.pyx:1110-1184) walks all points once per step to collect, for every cluster of
// the previous step, the current labels of its members in point order (`parent_labels_map`), and Bundle.isolate then
// groups each of those lists by label.  Both groupings for ALL parents at once:
//   order[k]  point indices sorted by (previous label, current label, index)      -- two stable radix sorts
//   rank[k]   position of point order[k] among the points with the same previous label, in index order
// so that the members of (parent p, child c) are one contiguous slice of `order`, their places in the parent's own
// label array are the same slice of `rank`, and the parent's label array is cur[order] scattered through rank.
__global__ void group_keys_kernel(const ll *__restrict__ prev, const ll *__restrict__ cur, ll n, unsigned long long *k_prev,
                                  unsigned long long *k_both, int *idx, int *flag) {
    ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ll p = prev[i], c = cur[i];
    if (p < 0 || c < 0 || p > 0x7fffffffll || c > 0x7fffffffll) *flag = 1;
    k_prev[i] = (unsigned long long)p;
    k_both[i] = ((unsigned long long)p << 32) | (unsigned long long)(unsigned)c;
    idx[i] = (int)i;
}

// sorted by previous label (stable): rank = position - first position of the label's run
__global__ void group_rank_kernel(const unsigned long long *__restrict__ k_sorted, const int *__restrict__ idx_sorted, ll n,
                                  int *run_start, int *rank_of_point) {
    ll k = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (k >= n) return;
    // first position of my run: binary search for the lower bound of my key
    const unsigned long long key = k_sorted[k];
    ll lo = 0, hi = k;
    while (lo < hi) {
        const ll mid = (lo + hi) >> 1;
        if (k_sorted[mid] < key) lo = mid + 1; else hi = mid;
    }
    rank_of_point[idx_sorted[k]] = (int)(k - lo);
    (void)run_start;
}

__global__ void group_out_kernel(const int *__restrict__ idx_sorted, const int *__restrict__ rank_of_point, ll n, ll *order,
                                 ll *rank) {
    ll k = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int i = idx_sorted[k];
    order[k] = i;
    rank[k] = rank_of_point[i];
}

extern "C" int cnnb200_labels_group(cnnb200_ctx *ctx, const int64_t *prev, const int64_t *cur, int loc, int64_t n,
                                    int64_t *order, int64_t *rank) {
    if (!ctx || !prev || !cur || !order || !rank) CNN_FAIL(CNNB200_E_INVALID, "labels_group: NULL argument");
    if (n < 0 || n >= (1ll << 31) - 1) CNN_FAIL(CNNB200_E_RANGE, "labels_group: n out of range");
    if (n == 0) return CNNB200_OK;
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t b8 = sizeof(ll) * (size_t)n;
    const ll *d_prev = prev, *d_cur = cur;
    ll *d_order = order, *d_rank = rank;
    if (loc == CNNB200_HOST) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_labels, b8));
        CNN_TRY(cnn_ensure(ctx, ctx->b_lab_map, b8));
        CNN_TRY(cnn_ensure(ctx, ctx->b_o_labels, 2 * b8));
        CNN_CUDA(cudaMemcpyAsync(ctx->b_labels.p, prev, b8, cudaMemcpyHostToDevice, st));
        CNN_CUDA(cudaMemcpyAsync(ctx->b_lab_map.p, cur, b8, cudaMemcpyHostToDevice, st));
        d_prev = ctx->b_labels.as<ll>();
        d_cur = ctx->b_lab_map.as<ll>();
        d_order = ctx->b_o_labels.as<ll>();
        d_rank = d_order + n;
    }
    CNN_TRY(cnn_ensure(ctx, ctx->b_lab_keys, b8));
    CNN_TRY(cnn_ensure(ctx, ctx->b_lab_keys2, b8));
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp0, b8));
    CNN_TRY(cnn_ensure(ctx, ctx->b_lab_vals, sizeof(int) * (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_lab_vals2, sizeof(int) * (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_rank, sizeof(int) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp2, 64));
    unsigned long long *k_prev = ctx->b_lab_keys.as<unsigned long long>(), *k_sorted = ctx->b_lab_keys2.as<unsigned long long>();
    unsigned long long *k_both = ctx->b_tmp0.as<unsigned long long>();
    int *idx = ctx->b_lab_vals.as<int>(), *idx_sorted = ctx->b_lab_vals2.as<int>(), *rank_of_point = ctx->b_rank.as<int>();
    int *flag = ctx->b_tmp2.as<int>();
    CNN_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
    const int nb = cnn_blocks(n, 256);
    group_keys_kernel<<<nb, 256, 0, st>>>(d_prev, d_cur, n, k_prev, k_both, idx, flag);
    size_t need = 0;
    CNN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, k_both, k_sorted, idx, idx_sorted, (int)n, 0, 64, st));
    CNN_TRY(cnn_ensure(ctx, ctx->b_cub, need));
    size_t cap = ctx->b_cub.cap;
    CNN_CUDA(cub::DeviceRadixSort::SortPairs(ctx->b_cub.p, cap, k_prev, k_sorted, idx, idx_sorted, (int)n, 0, 32, st));
    group_rank_kernel<<<nb, 256, 0, st>>>(k_sorted, idx_sorted, n, nullptr, rank_of_point);
    cap = ctx->b_cub.cap;
    CNN_CUDA(cub::DeviceRadixSort::SortPairs(ctx->b_cub.p, cap, k_both, k_sorted, idx, idx_sorted, (int)n, 0, 64, st));
    group_out_kernel<<<nb, 256, 0, st>>>(idx_sorted, rank_of_point, n, d_order, d_rank);
    CNN_LAUNCHED(ctx, 3);
    CNN_KERNEL_CHECK();
    int *h;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&h));
    CNN_CUDA(cudaMemcpyAsync(h, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (loc == CNNB200_HOST) {
        CNN_CUDA(cudaMemcpyAsync(order, d_order, b8, cudaMemcpyDeviceToHost, st));
        CNN_CUDA(cudaMemcpyAsync(rank, d_rank, b8, cudaMemcpyDeviceToHost, st));
    }
    CNN_CUDA(cudaStreamSynchronize(st));
    if (h[0]) CNN_FAIL(CNNB200_E_INVALID, "labels_group: labels must lie in [0, 2^31)");
    return CNNB200_OK;
}
