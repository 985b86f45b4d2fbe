// common.cuh -- context, scratch buffers and error plumbing shared by all translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "cnnb200.h"

typedef int64_t ll;  // the ABI index type (long on LP64)

void cnn_set_error(const char *fmt, ...);

#define CNN_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            cnn_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return CNNB200_E_CUDA;                                                          \
        }                                                                                   \
    } while (0)

#define CNN_TRY(call)                  \
    do {                               \
        int rc__ = (call);             \
        if (rc__ != CNNB200_OK) return rc__; \
    } while (0)

#define CNN_KERNEL_CHECK() CNN_CUDA(cudaGetLastError())

#define CNN_FAIL(code, ...)        \
    do {                           \
        cnn_set_error(__VA_ARGS__); \
        return (code);             \
    } while (0)

struct GridDesc {
    double lo[3];
    double h;
    int dims[3];
    int d;
};

// A growable device scratch buffer owned by the context.
struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

enum StatSlot {
    ST_EDGES = 0,
    ST_CHECKS = 1,
    ST_ACCEPTED = 2,
    ST_SKIPPED = 3,
    ST_DIST = 4,
    ST_RECHECK = 5,
    ST_FALLBACK = 6,  // count mode: rows whose index window did not fit the shared-memory bitmap
    ST_PATH = 7,      // which kernels served the last calls: bits 0-3 search method (1 grid, 2 tiled f64, 3 tensor), bit 4 packed count mode
    ST_WL_COUNT = 8,  // device-side counters that live next to the statistics: work-list length,
    ST_COUNTER = 9,   // dynamic work distribution (rows / work-list chunks handed out so far)
    ST_N = 10
};

// timing slots (cnnb200_timings): device time of the stages of the path, by CUDA events
enum TimeSlot {
    T_H2D = 0,      // host -> device copies of inputs
    T_GRID_BUILD,   // K1: bounding box, cell ids, sort by cell, cell starts, gather
    T_SEARCH_COUNT, // K2/K3 count pass
    T_SCAN,         // degree scan (CUB)
    T_SEARCH_FILL,  // K2/K3 fill pass
    T_INGEST,       // K4: CSR / padded ingest
    T_COMP_INIT,    // core flags + parent init
    T_COMP_SAMPLE,  // K5/K6 sampling pass
    T_COMP_EDGES,   // K5/K6 full edge pass
    T_COMP_FLATTEN, // pointer-jumping passes + multi-GPU merge
    T_LABELS,       // K7 numbering
    T_SORT_BY_SIZE, // K7 Labels.sort_by_size
    T_D2H,          // device -> host copies of results
    T_CN_COUNT,     // K5 count mode
    T_N
};
#define CNN_MAX_SPANS 256

struct cnnb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cnnb200_alloc_fn alloc = nullptr;
    cnnb200_free_fn free_ = nullptr;
    void *alloc_user = nullptr;
    int sm_count = 148;
    ll single_pass_limit = 32ll << 30;  // bytes of gapped scratch the single-pass search may use (0 = two-pass)
    bool fuse_union = true;  // sampling pass links its accepted edges itself (CNNB200_FUSED_UNION=0: separate kernel)
    int n_samples = 1;  // edges per core row tried by the first sampling round (tunable: CNNB200_SAMPLES)
    int far_stride = 8; // second round: every far_stride-th row of a tree tries two far neighbours (0: off; CNNB200_FAR_STRIDE)
    int n_far = 2;      // far entries such a row tries: the last one and the one a quarter into the upper part (CNNB200_FAR_SAMPLES)
    int wl_chunk = 16;  // work-list batches of 32 edges a warp takes at a time (CNNB200_WL_CHUNK)
    ll balance_per_row = 250;  // row-block balance: cost = candidate pairs + this * rows (CNNB200_BALANCE_ROW)

    // ---- neighbourhood structure (internal index space) -------------------------------
    bool have_nb = false;
    ll n = 0, nnz = 0, max_deg = 0;
    const ll *row_ptr = nullptr;       // [n+1]  (points into b_row_ptr or borrowed memory)
    const int32_t *col_idx = nullptr;  // [nnz]
    const int32_t *perm = nullptr;     // [n] internal -> caller index; nullptr = identity
    Buf b_row_ptr, b_col_idx, b_col_spare, b_perm, b_deg;

    // ---- uniform grid of the last cnn_grid_build + the row block being searched ---------------
    bool have_grid = false;
    GridDesc grid;
    ll grid_n = 0;
    double grid_rc = 0;
    int grid_metric = 0;
    ll blk_begin = 0, blk_end = 0, blk_goff0 = 0;
    bool blk_single = false, blk_masks = false;
    bool have_cap = false, grid_filter = false;  // candidate-list prefix computed / the fp32 filter serves this grid
    bool mask_words_pending = false;             // the mask pass ran before the host knew how many words it needs
    ll mask_words_needed = 0;
    const ll *mask_total = nullptr;              // device: total words of the last plan
    Buf b_tile_off, b_masks;       // fp32 filter: offset of every warp tile's hit masks, the masks
    Buf b_cell_count;

    // ---- neighbour search scratch ------------------------------------------------------
    Buf b_x;         // f64 [n, d] as given by the caller (device copy)
    Buf b_xs;        // f64 SoA [d][n] in cell-sorted order
    Buf b_xperm;     // f64 [n, d]: coordinates in the pivot-sorted order of the tensor-core search
    bool tensor_reorder = true;  // CNNB200_TENSOR_REORDER=0: keep the caller's order, visit every tile
    Buf b_xf;        // float4 [n]: cell-local fp32 coordinates (x on lattices A and B, y, z) for the filter kernel
    bool use_filter = true;  // fp32 filter + exact recheck in the cell-list search (CNNB200_FILTER=0: f64 kernel)
    bool grid_recheck_pending = false;
    Buf b_cell_id, b_cell_id_sorted, b_idx, b_cell_start, b_bbox, b_cub;
    Buf b_tmp0, b_tmp1, b_tmp2;

    // ---- the point set the neighbourhoods were built from + the OTHER set of a prediction --------------------------
    bool have_points = false, pts_grid = false;
    const double *pts_x = nullptr;   // row-major coordinates on the device (caller order), valid while have_points
    ll pts_d = 0;
    double pts_rc = 0;
    int pts_metric = 0;
    bool have_other = false;
    ll o_m = 0, o_nnz = 0;
    Buf b_o_x, b_o_deg, b_o_row_ptr, b_o_col_idx, b_o_col_sorted, b_o_labels;
    const int *o_col_orig = nullptr;  // the predictand lists in the order they were given (first-match order), when they had to be sorted

    // ---- bit-packed rows (packed.cu): valid while packed_version == nb_version --------------------------------------
    ll nb_version = 0, packed_version = -1, packed_words = 0;
    bool use_packed = true;  // CNNB200_PACKED=0: per-entry kernels only
    bool pack_in_fit = false; // CNNB200_PACK_FIT=1: the fit's edge passes run on packed rows (measured slower on C2: packing costs 0.73 ms)
    bool comp_packed = false;
    Buf b_wrows, b_wcnt;

    // ---- components ----------------------------------------------------------------------
    bool have_comp = false;
    ll n_member_cutoff = 0, similarity_cutoff = 0;
    Buf b_parent;  // int32 [n]
    Buf b_core;    // uint8 [n]
    Buf b_above;     // int32 [n]: position of the first entry > i in row i
    Buf b_edges;     // int32 [rows * samples]: edges accepted by the sampling pass
    Buf b_worklist;  // int2 [<= nnz/2]: candidate edges the scan could not settle
    Buf b_minorig, b_flags, b_rank, b_labels, b_consider, b_stats;
    Buf b_lab_count, b_lab_first, b_lab_keys, b_lab_keys2, b_lab_vals, b_lab_vals2, b_lab_map;

    // ---- pinned host staging ---------------------------------------------------------------
    void *h_pinned = nullptr;
    size_t h_pinned_cap = 0;

    ll stats[ST_N] = {0};
    Buf b_bounds;     // f64 [n_bounds]: periods of the periodic metric
    ll n_bounds = 0;
    double bounds_h[3] = {0, 0, 0};  // host copy of the first three (periodic cell list, periodic.cu)
    Buf b_pg_x, b_pg_orig, b_pg_xq, b_pg_perm, b_pg_deg, b_pg_img, b_pg_row_ptr, b_pg_col;  // periodic cell list: augmented set + results
    Buf b_lab_summary;
    bool have_lab_summary = false, lab_summary_cached = false;  // cached: the three numbers already sit in lab_summary_host
    ll lab_summary_host[3] = {0, 0, 0};
    ll lab_summary_n = 0;
    ll known_max_label = -1;
    int shard_world = 0;            // row-sharded fit: ranks of the last fit_sharded_begin and the rows each of them built
    ll shard_ranges[2 * 1024] = {0};  // [lo, hi) per rank (block + halo)
    ll tensor_band_cap = 0;  // capacity (slabs) the hit bitmaps of the tensor-core search needed last time
    Buf b_tslab, b_tband, b_tdir;  // tensor-core search: hit slabs, borderline slabs, the two slab directories

    // ---- measurement ---------------------------------------------------------------------------
    bool timing = false;
    int n_spans = 0;
    int span_slot[CNN_MAX_SPANS];
    cudaEvent_t span_ev[CNN_MAX_SPANS][2];
    int n_events_made = 0;
    double t_ms[T_N] = {0};
    ll t_calls[T_N] = {0};
    ll n_launches = 0;  // hand-written kernels launched (library kernels such as CUB are not counted)
};

// RAII span: records an event pair around a stage when timing is enabled
struct Span {
    cnnb200_ctx *c;
    int k;
    Span(cnnb200_ctx *ctx, int slot) : c(ctx), k(-1) {
        if (!c->timing || c->n_spans >= CNN_MAX_SPANS) return;
        k = c->n_spans++;
        if (k >= c->n_events_made) {
            cudaEventCreate(&c->span_ev[k][0]);
            cudaEventCreate(&c->span_ev[k][1]);
            c->n_events_made = k + 1;
        }
        c->span_slot[k] = slot;
        cudaEventRecord(c->span_ev[k][0], c->stream);
    }
    ~Span() {
        if (k >= 0) cudaEventRecord(c->span_ev[k][1], c->stream);
    }
};
#define CNN_LAUNCHED(ctx, k) ((ctx)->n_launches += (k))

int cnn_ensure(cnnb200_ctx *ctx, Buf &b, size_t bytes);
void cnn_release(cnnb200_ctx *ctx, Buf &b);
int cnn_pinned(cnnb200_ctx *ctx, size_t bytes, void **out);

static inline int cnn_blocks(ll work_items, int per_block) {
    ll b = (work_items + per_block - 1) / per_block;
    if (b < 1) b = 1;
    return (int)b;
}

// --- stage entry points implemented across the .cu files ----------------------------------
int cnn_grid_build(cnnb200_ctx *ctx, const double *d_x, ll n, ll d, double radius_cutoff, int metric, bool want_cap);
int cnn_grid_cap(cnnb200_ctx *ctx);
int cnn_grid_rows(cnnb200_ctx *ctx, ll pb, ll pe, int *deg, bool speculative);
int cnn_grid_fill(cnnb200_ctx *ctx, ll pb, ll pe, const ll *row_ptr, int *col_idx);
int cnn_grid_search(cnnb200_ctx *ctx, const double *d_x, ll n, ll d, double radius_cutoff, int metric);
int cnn_grid_search_periodic(cnnb200_ctx *ctx, const double *d_x, ll n, ll d, double radius_cutoff, bool *served);
int cnn_dense_search(cnnb200_ctx *ctx, const double *d_x, ll n, ll d, double radius_cutoff, int metric,
                     int method);
int cnn_labels_sorted(cnnb200_ctx *ctx, int64_t current_start, int64_t member_cutoff, int64_t max_clusters, int64_t *labels,
                      int64_t *d_labels, int loc, int64_t *n_clusters_out);
int cnn_labels_fused(cnnb200_ctx *ctx, ll member_cutoff, ll max_clusters, ll *d_labels, unsigned long long **counters_out);
int cnn_sort_csr_rows(cnnb200_ctx *ctx, const ll *row_ptr, const int *col, ll n, ll nnz, Buf &out);
int cnn_csr_rows_unsorted(cnnb200_ctx *ctx, const ll *row_ptr, const int *col, ll n, int *unsorted);
int cnn_ensure_sorted_rows(cnnb200_ctx *ctx, bool known_unsorted, int *sorted_out);
int cnn_finish_csr(cnnb200_ctx *ctx, ll n, ll nnz, ll max_deg, const ll *row_ptr, const int *col_idx);
int cnn_pack_rows(cnnb200_ctx *ctx, bool *usable_out);
int cnn_worklist_packed(cnnb200_ctx *ctx, const int2 *worklist, const unsigned long long *wl_count, int cutoff, int *parent);
int cnn_count_packed(cnnb200_ctx *ctx, ll n_member_cutoff, ll row_begin, ll row_end, unsigned *d_counts);
int cnn_scan_deg_to_row_ptr(cnnb200_ctx *ctx, const int32_t *deg, ll n, ll *row_ptr, ll *nnz_out,
                            ll *max_deg_out);
