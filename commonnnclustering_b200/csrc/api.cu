// api.cu -- context management, scratch allocation, error reporting, host-buffer pipeline.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void cnn_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int cnnb200_version(void) { return CNNB200_VERSION; }
extern "C" const char *cnnb200_last_error(void) { return g_err; }

static void *default_alloc(size_t bytes, void *) {
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
static void default_free(void *p, void *) { cudaFree(p); }

int cnn_ensure(cnnb200_ctx *ctx, Buf &b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return CNNB200_OK;
    if (b.p) {
        // the old block may still be in use by work enqueued on the stream
        CNN_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->free_(b.p, ctx->alloc_user);
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8;  // head-room so that slightly larger inputs do not realloc
    want = (want + 255) & ~(size_t)255;
    void *p = ctx->alloc(want, ctx->alloc_user);
    if (!p) {
        want = (bytes + 255) & ~(size_t)255;
        p = ctx->alloc(want, ctx->alloc_user);
    }
    if (!p) CNN_FAIL(CNNB200_E_NOMEM, "device allocation of %zu bytes failed", want);
    b.p = p;
    b.cap = want;
    return CNNB200_OK;
}

void cnn_release(cnnb200_ctx *ctx, Buf &b) {
    if (b.p) ctx->free_(b.p, ctx->alloc_user);
    b.p = nullptr;
    b.cap = 0;
}

int cnn_pinned(cnnb200_ctx *ctx, size_t bytes, void **out) {
    if (ctx->h_pinned_cap < bytes) {
        if (ctx->h_pinned) {
            CNN_CUDA(cudaStreamSynchronize(ctx->stream));
            cudaFreeHost(ctx->h_pinned);
            ctx->h_pinned = nullptr;
            ctx->h_pinned_cap = 0;
        }
        size_t want = bytes < 4096 ? 4096 : bytes;
        CNN_CUDA(cudaMallocHost(&ctx->h_pinned, want));
        ctx->h_pinned_cap = want;
    }
    *out = ctx->h_pinned;
    return CNNB200_OK;
}

extern "C" int cnnb200_ctx_create(int device, cnnb200_ctx **out) {
    if (!out) CNN_FAIL(CNNB200_E_INVALID, "ctx_create: out is NULL");
    int count = 0;
    CNN_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count)
        CNN_FAIL(CNNB200_E_INVALID, "ctx_create: device %d out of range (%d devices)", device, count);
    CNN_CUDA(cudaSetDevice(device));
    cnnb200_ctx *ctx = new (std::nothrow) cnnb200_ctx();
    if (!ctx) CNN_FAIL(CNNB200_E_NOMEM, "ctx_create: host allocation failed");
    ctx->device = device;
    ctx->alloc = default_alloc;
    ctx->free_ = default_free;
    cudaDeviceProp prop;
    CNN_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    if (const char *e = getenv("CNNB200_SAMPLES")) ctx->n_samples = atoi(e);
    if (const char *e = getenv("CNNB200_FAR_STRIDE")) ctx->far_stride = atoi(e);
    if (const char *e = getenv("CNNB200_FAR_SAMPLES")) ctx->n_far = atoi(e);
    if (const char *e = getenv("CNNB200_WL_CHUNK")) ctx->wl_chunk = atoi(e) > 0 ? atoi(e) : 1;
    if (const char *e = getenv("CNNB200_BALANCE_ROW")) ctx->balance_per_row = atoll(e) >= 0 ? atoll(e) : 0;
    if (const char *e = getenv("CNNB200_FUSED_UNION")) ctx->fuse_union = atoi(e) != 0;
    if (const char *e = getenv("CNNB200_FILTER")) ctx->use_filter = atoi(e) != 0;
    if (const char *e = getenv("CNNB200_PACKED")) ctx->use_packed = atoi(e) != 0;
    if (const char *e = getenv("CNNB200_PACK_FIT")) ctx->pack_in_fit = atoi(e) != 0;
    if (const char *e = getenv("CNNB200_TENSOR_REORDER")) ctx->tensor_reorder = atoi(e) != 0;
    if (const char *e = getenv("CNNB200_TENSOR_PAIR_CAP")) ctx->tensor_band_cap = atoll(e);  // initial slab capacity of the tensor-core search (tests)
    if (const char *e = getenv("CNNB200_SINGLE_PASS_MB")) ctx->single_pass_limit = (ll)atoll(e) << 20;
    *out = ctx;
    return CNNB200_OK;
}

#define FOR_EACH_BUF(c, F)                                                                          \
    F(c->b_row_ptr) F(c->b_col_idx) F(c->b_col_spare) F(c->b_o_col_sorted) F(c->b_perm) F(c->b_deg) F(c->b_x) F(c->b_wrows) F(c->b_wcnt) F(c->b_lab_summary) F(c->b_bounds) F(c->b_xs) F(c->b_xf) F(c->b_xperm) F(c->b_o_x) F(c->b_o_deg) F(c->b_o_row_ptr) F(c->b_o_col_idx) F(c->b_o_labels) F(c->b_cell_id)   \
    F(c->b_cell_id_sorted) F(c->b_idx) F(c->b_cell_start) F(c->b_bbox) F(c->b_cub) F(c->b_tmp0)     \
    F(c->b_tmp1) F(c->b_tmp2) F(c->b_parent) F(c->b_core) F(c->b_above) F(c->b_edges) F(c->b_worklist) F(c->b_cell_count) F(c->b_minorig) F(c->b_flags)             \
    F(c->b_rank) F(c->b_labels) F(c->b_consider) F(c->b_stats) F(c->b_lab_count) F(c->b_lab_first)  \
    F(c->b_lab_keys) F(c->b_lab_keys2) F(c->b_lab_vals) F(c->b_lab_vals2) F(c->b_lab_map) \
    F(c->b_tile_off) F(c->b_masks) F(c->b_pg_x) F(c->b_pg_orig) F(c->b_pg_xq) F(c->b_pg_perm) F(c->b_pg_deg) F(c->b_pg_img) F(c->b_pg_row_ptr) F(c->b_pg_col) F(c->b_tslab) F(c->b_tband) F(c->b_tdir)

extern "C" int cnnb200_ctx_release(cnnb200_ctx *ctx) {
    if (!ctx) CNN_FAIL(CNNB200_E_INVALID, "ctx is NULL");
    CNN_CUDA(cudaSetDevice(ctx->device));
    CNN_CUDA(cudaStreamSynchronize(ctx->stream));
#define REL(b) cnn_release(ctx, b);
    FOR_EACH_BUF(ctx, REL)
#undef REL
    ctx->have_nb = false;
    ctx->have_comp = false;
    ctx->have_grid = false;
    ctx->have_points = false;
    ctx->have_other = false;
    ctx->n_bounds = 0;
    ctx->row_ptr = nullptr;
    ctx->col_idx = nullptr;
    ctx->perm = nullptr;
    return CNNB200_OK;
}

extern "C" int cnnb200_ctx_destroy(cnnb200_ctx *ctx) {
    if (!ctx) return CNNB200_OK;
    int rc = cnnb200_ctx_release(ctx);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    for (int k = 0; k < ctx->n_events_made; ++k) {
        cudaEventDestroy(ctx->span_ev[k][0]);
        cudaEventDestroy(ctx->span_ev[k][1]);
    }
    delete ctx;
    return rc;
}

extern "C" int cnnb200_ctx_set_stream(cnnb200_ctx *ctx, void *stream) {
    if (!ctx) CNN_FAIL(CNNB200_E_INVALID, "ctx is NULL");
    ctx->stream = (cudaStream_t)stream;
    return CNNB200_OK;
}

extern "C" int cnnb200_ctx_set_allocator(cnnb200_ctx *ctx, cnnb200_alloc_fn alloc, cnnb200_free_fn free_,
                                         void *user) {
    if (!ctx) CNN_FAIL(CNNB200_E_INVALID, "ctx is NULL");
    CNN_TRY(cnnb200_ctx_release(ctx));  // blocks from the old allocator must go back to it
    if (alloc && free_) {
        ctx->alloc = alloc;
        ctx->free_ = free_;
        ctx->alloc_user = user;
    } else {
        ctx->alloc = default_alloc;
        ctx->free_ = default_free;
        ctx->alloc_user = nullptr;
    }
    return CNNB200_OK;
}

extern "C" int cnnb200_ctx_set_periodic_bounds(cnnb200_ctx *ctx, const double *bounds, int64_t d) {
    if (!ctx || !bounds || d < 1) CNN_FAIL(CNNB200_E_INVALID, "set_periodic_bounds: bad argument");
    CNN_CUDA(cudaSetDevice(ctx->device));
    CNN_TRY(cnn_ensure(ctx, ctx->b_bounds, sizeof(double) * (size_t)d));
    CNN_CUDA(cudaMemcpyAsync(ctx->b_bounds.p, bounds, sizeof(double) * (size_t)d, cudaMemcpyHostToDevice, ctx->stream));
    CNN_CUDA(cudaStreamSynchronize(ctx->stream));  // the host array may go away
    ctx->n_bounds = d;
    for (int k = 0; k < 3; ++k) ctx->bounds_h[k] = k < d ? bounds[k] : 0.0;
    return CNNB200_OK;
}

// Page-locks a caller array in place (cudaHostRegister) so that repeated fits on the same coordinates copy at PCIe
// speed instead of through the driver's pageable staging; nothing is copied or cached.
extern "C" int cnnb200_host_register(void *ptr, size_t bytes) {
    if (!ptr || !bytes) CNN_FAIL(CNNB200_E_INVALID, "host_register: bad argument");
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return CNNB200_OK; }
    if (e != cudaSuccess) { cudaGetLastError(); CNN_FAIL(CNNB200_E_CUDA, "host_register: %s", cudaGetErrorString(e)); }
    return CNNB200_OK;
}
extern "C" int cnnb200_host_unregister(void *ptr) {
    if (!ptr) return CNNB200_OK;
    if (cudaHostUnregister(ptr) != cudaSuccess) cudaGetLastError();
    return CNNB200_OK;
}

extern "C" int cnnb200_ctx_synchronize(cnnb200_ctx *ctx) {
    if (!ctx) CNN_FAIL(CNNB200_E_INVALID, "ctx is NULL");
    CNN_CUDA(cudaStreamSynchronize(ctx->stream));
    return CNNB200_OK;
}

extern "C" int cnnb200_stats(cnnb200_ctx *ctx, int64_t *stats, int64_t len) {
    if (!ctx || !stats) CNN_FAIL(CNNB200_E_INVALID, "stats: NULL argument");
    for (int64_t i = 0; i < len; ++i) stats[i] = i < ST_N ? ctx->stats[i] : 0;
    return CNNB200_OK;
}

extern "C" int cnnb200_ctx_set_timing(cnnb200_ctx *ctx, int enabled) {
    if (!ctx) CNN_FAIL(CNNB200_E_INVALID, "ctx is NULL");
    ctx->timing = enabled != 0;
    return CNNB200_OK;
}

extern "C" int cnnb200_timings(cnnb200_ctx *ctx, double *ms, int64_t *calls, int64_t len, int64_t *launches,
                               int reset) {
    if (!ctx) CNN_FAIL(CNNB200_E_INVALID, "ctx is NULL");
    CNN_CUDA(cudaSetDevice(ctx->device));
    CNN_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < ctx->n_spans; ++k) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, ctx->span_ev[k][0], ctx->span_ev[k][1]) == cudaSuccess) {
            ctx->t_ms[ctx->span_slot[k]] += t;
            ctx->t_calls[ctx->span_slot[k]] += 1;
        } else {
            cudaGetLastError();
        }
    }
    ctx->n_spans = 0;
    for (int64_t i = 0; i < len; ++i) {
        if (ms) ms[i] = i < T_N ? ctx->t_ms[i] : 0.0;
        if (calls) calls[i] = i < T_N ? ctx->t_calls[i] : 0;
    }
    if (launches) *launches = ctx->n_launches;
    if (reset) {
        for (int i = 0; i < T_N; ++i) { ctx->t_ms[i] = 0; ctx->t_calls[i] = 0; }
        ctx->n_launches = 0;
    }
    return CNNB200_OK;
}

extern "C" int cnnb200_neighbours_info(cnnb200_ctx *ctx, int64_t *n, int64_t *nnz, int64_t *max_deg) {
    if (!ctx) CNN_FAIL(CNNB200_E_INVALID, "ctx is NULL");
    if (!ctx->have_nb) CNN_FAIL(CNNB200_E_STATE, "no neighbourhoods held by this context");
    if (n) *n = ctx->n;
    if (nnz) *nnz = ctx->nnz;
    if (max_deg) *max_deg = ctx->max_deg;
    return CNNB200_OK;
}

extern "C" int cnnb200_neighbours_device_view(cnnb200_ctx *ctx, const int64_t **row_ptr,
                                              const int32_t **col_idx, const int32_t **perm) {
    if (!ctx) CNN_FAIL(CNNB200_E_INVALID, "ctx is NULL");
    if (!ctx->have_nb) CNN_FAIL(CNNB200_E_STATE, "no neighbourhoods held by this context");
    if (row_ptr) *row_ptr = (const int64_t *)ctx->row_ptr;
    if (col_idx) *col_idx = ctx->col_idx;
    if (perm) *perm = ctx->perm;
    return CNNB200_OK;
}

// ------------------------------------------------------------------------------------------
// Radius search front end: stage the coordinates on the device, dispatch on method.
// ------------------------------------------------------------------------------------------
extern "C" int cnnb200_neighbours_from_points(cnnb200_ctx *ctx, const double *x, int loc, int64_t n,
                                              int64_t d, double radius_cutoff, int metric, int method,
                                              int64_t *nnz_out) {
    if (!ctx || !x) CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_points: NULL argument");
    if (n < 1 || d < 1) CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_points: need n >= 1 and d >= 1");
    if (n >= (1ll << 31) - 1) CNN_FAIL(CNNB200_E_RANGE, "neighbours_from_points: n must be < 2^31-1");
    if (metric != CNNB200_METRIC_EUCLIDEAN_REDUCED && metric != CNNB200_METRIC_EUCLIDEAN &&
        metric != CNNB200_METRIC_EUCLIDEAN_PERIODIC_REDUCED)
        CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_points: unknown metric %d", metric);
    const bool periodic = metric == CNNB200_METRIC_EUCLIDEAN_PERIODIC_REDUCED;
    if (periodic) {
        if (method == 3) CNN_FAIL(CNNB200_E_INVALID, "the periodic metric is served by the grid (d <= 3) and tiled methods");
        if (ctx->n_bounds != d)
            CNN_FAIL(CNNB200_E_STATE, "periodic metric: set %lld bounds with cnnb200_ctx_set_periodic_bounds first", (long long)d);
        // cell list on the image-augmented set from a few thousand points on; the all-pairs kernel below that
        if (method == 0) method = (d <= 3 && n >= 4096) ? 1 : 2;
    }
    if (!(radius_cutoff >= 0.0)) CNN_FAIL(CNNB200_E_INVALID, "neighbours_from_points: radius_cutoff must be >= 0");
    // auto: cell list up to three dimensions; above that the tcgen05 filter + f64 recheck (the same neighbour sets as
    // the f64 all-pairs kernel, NeighboursGetterExtBruteForce._get for any D) from a few thousand points on
    const bool automatic = method == 0;
    if (method == 0) method = d <= 3 ? 1 : ((d <= 85 && n >= 4096) ? 3 : 2);
    if (method == 1 && d > 3) CNN_FAIL(CNNB200_E_INVALID, "grid method needs d <= 3 (got %lld)", (ll)d);
    if (method < 1 || method > 3) CNN_FAIL(CNNB200_E_INVALID, "unknown method %d", method);
    CNN_CUDA(cudaSetDevice(ctx->device));
    ctx->have_nb = false;
    ctx->have_comp = false;
    ctx->have_points = false;
    const double *d_x = x;
    if (loc == CNNB200_HOST) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_x, sizeof(double) * (size_t)n * (size_t)d));
        Span sp(ctx, T_H2D);
        CNN_CUDA(cudaMemcpyAsync(ctx->b_x.p, x, sizeof(double) * (size_t)n * (size_t)d,
                                 cudaMemcpyHostToDevice, ctx->stream));
        d_x = ctx->b_x.as<double>();
    }
    if (periodic && method == 1) {
        bool served = false;
        CNN_TRY(cnn_grid_search_periodic(ctx, d_x, n, d, radius_cutoff, &served));
        if (!served) method = 2;  // a period shorter than two radii: images would overlap, the all-pairs kernel decides
    }
    if (periodic && method == 1) {
    } else if (method == 1)
        CNN_TRY(cnn_grid_search(ctx, d_x, n, d, radius_cutoff, metric));
    else {
        int rc = cnn_dense_search(ctx, d_x, n, d, radius_cutoff, metric, method);
        if (rc == CNNB200_E_RANGE && automatic && method == 3) {  // more pairs than the pair list can index: exact kernel
            method = 2;
            rc = cnn_dense_search(ctx, d_x, n, d, radius_cutoff, metric, method);
        }
        CNN_TRY(rc);
    }
    ctx->have_nb = true;
    ctx->stats[ST_PATH] = (ctx->stats[ST_PATH] & ~15ll) | method;
    ++ctx->nb_version;
    ctx->have_other = false;
    ctx->have_points = true;  // prediction for other points needs the coordinates, the cutoff and the metric again
    ctx->pts_grid = method == 1 && !periodic;
    ctx->pts_x = d_x;
    ctx->pts_d = d;
    ctx->pts_rc = radius_cutoff;
    ctx->pts_metric = metric;
    if (nnz_out) *nnz_out = ctx->nnz;
    return CNNB200_OK;
}

// Ranking of the components + labels for a forest that is complete (after cnnb200_components_edges /
// _from_counts / _merge) when every label comes from this fit: cnn_labels_fused, one D2H copy and ONE synchronisation;
// the general chain (numbering, then the size sort over all labels) when more clusters survive than one CTA ranks.
// `labels` is the caller's array (host or device), `d_labels` the device array the kernels write.
int cnn_labels_sorted(cnnb200_ctx *ctx, int64_t current_start, int64_t member_cutoff, int64_t max_clusters, int64_t *labels,
                      int64_t *d_labels, int loc, int64_t *n_clusters_out) {
    static const bool fused_labels = !getenv("CNNB200_FUSED_LABELS") || atoi(getenv("CNNB200_FUSED_LABELS")) != 0;
    const ll n = ctx->n;
    unsigned long long *counters = nullptr;
    CNN_TRY(cnn_labels_fused(ctx, member_cutoff, max_clusters, d_labels, &counters));
    unsigned long long *h;
    CNN_TRY(cnn_pinned(ctx, 128, (void **)&h));
    if (loc == CNNB200_HOST) {
        Span sp(ctx, T_D2H);
        CNN_CUDA(cudaMemcpyAsync(labels, d_labels, sizeof(ll) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CNN_CUDA(cudaMemcpyAsync(h, counters, sizeof(ll) * 6, cudaMemcpyDeviceToHost, ctx->stream));
    CNN_CUDA(cudaMemcpyAsync(h + 8, ctx->b_stats.p, sizeof(ll) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CNN_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 4; ++k) ctx->stats[k] = (ll)h[8 + k];
    const int64_t n_clusters = (int64_t)h[0];
    if (n_clusters_out) *n_clusters_out = n_clusters;
    if (h[2] || !fused_labels) {
        // more surviving clusters than one CTA ranks: the general chain (numbering, then the size sort over all labels)
        if (loc == CNNB200_HOST) CNN_CUDA(cudaMemsetAsync(d_labels, 0, sizeof(ll) * (size_t)n, ctx->stream));
        int64_t nc2 = 0;
        CNN_TRY(cnnb200_components_labels(ctx, current_start, d_labels, CNNB200_DEVICE, /*fresh=*/1, &nc2));
        ctx->known_max_label = nc2 > 0 ? current_start + nc2 - 1 : 0;
        CNN_TRY(cnnb200_sort_by_size(ctx, d_labels, CNNB200_DEVICE, n, member_cutoff, max_clusters, nullptr, 0));
        if (loc == CNNB200_HOST) {
            CNN_CUDA(cudaMemcpyAsync(labels, d_labels, sizeof(ll) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
            CNN_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        return CNNB200_OK;
    }
    ctx->have_lab_summary = true;
    ctx->lab_summary_n = n;
    ctx->lab_summary_host[0] = (ll)h[3];
    ctx->lab_summary_host[1] = (ll)h[4];
    ctx->lab_summary_host[2] = n - (ll)h[5];
    ctx->lab_summary_cached = true;
    return CNNB200_OK;
}

extern "C" int cnnb200_components_labels_sorted(cnnb200_ctx *ctx, int64_t current_start, int64_t member_cutoff,
                                                int64_t max_clusters, int64_t *labels, int loc, int64_t *n_clusters_out) {
    if (!ctx || !labels) CNN_FAIL(CNNB200_E_INVALID, "components_labels_sorted: NULL argument");
    if (!ctx->have_comp) CNN_FAIL(CNNB200_E_STATE, "components_labels_sorted: call components_begin first");
    CNN_CUDA(cudaSetDevice(ctx->device));
    CNN_TRY(cnn_ensure(ctx, ctx->b_labels, sizeof(ll) * (size_t)ctx->n));
    int64_t *d_labels = loc == CNNB200_HOST ? (int64_t *)ctx->b_labels.p : labels;
    return cnn_labels_sorted(ctx, current_start, member_cutoff, max_clusters, labels, d_labels, loc, n_clusters_out);
}

extern "C" int cnnb200_fit_points(cnnb200_ctx *ctx, const double *x, int loc, int64_t n, int64_t d,
                                  double radius_cutoff, int metric, int method, int64_t n_member_cutoff,
                                  int64_t similarity_cutoff, int64_t current_start, int sort_by_size,
                                  int64_t member_cutoff, int64_t max_clusters, int64_t *labels,
                                  int64_t *n_clusters_out) {
    if (!labels) CNN_FAIL(CNNB200_E_INVALID, "fit_points: labels is NULL");
    CNN_TRY(cnnb200_neighbours_from_points(ctx, x, loc, n, d, radius_cutoff, metric, method, nullptr));
    if (!sort_by_size)
        return cnnb200_fit(ctx, n_member_cutoff, similarity_cutoff, current_start, labels, nullptr, loc,
                           n_clusters_out);
    return cnnb200_fit_sorted(ctx, n_member_cutoff, similarity_cutoff, current_start, member_cutoff, max_clusters, labels, loc,
                              n_clusters_out);
}

// A fresh fit on the neighbourhoods the context holds (whatever built them) followed by sort_by_size: the labels stay
// on the device, numbering and size sort collapse into one ranking of the components (cnn_labels_fused), one D2H copy
// and ONE synchronisation at the end.
extern "C" int cnnb200_fit_sorted(cnnb200_ctx *ctx, int64_t n_member_cutoff, int64_t similarity_cutoff, int64_t current_start,
                                  int64_t member_cutoff, int64_t max_clusters, int64_t *labels, int loc,
                                  int64_t *n_clusters_out) {
    if (!ctx || !labels) CNN_FAIL(CNNB200_E_INVALID, "fit_sorted: NULL argument");
    if (!ctx->have_nb) CNN_FAIL(CNNB200_E_STATE, "fit_sorted: no neighbourhoods held by this context");
    const ll n = ctx->n;
    CNN_CUDA(cudaSetDevice(ctx->device));
    CNN_TRY(cnn_ensure(ctx, ctx->b_labels, sizeof(ll) * (size_t)n));
    int64_t *d_labels = loc == CNNB200_HOST ? (int64_t *)ctx->b_labels.p : labels;
    CNN_TRY(cnnb200_components_begin(ctx, n_member_cutoff, similarity_cutoff, nullptr, CNNB200_DEVICE));
    CNN_TRY(cnnb200_components_edges(ctx, 0, n));
    return cnn_labels_sorted(ctx, current_start, member_cutoff, max_clusters, labels, d_labels, loc, n_clusters_out);
}
