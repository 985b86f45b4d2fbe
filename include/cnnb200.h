/*
 * cnnb200.h -- C-ABI of the B200-native common-nearest-neighbour clustering fit path.
 *
 * This is the drop-in boundary: what a maintainer of the reference (cnnclustering 0.5.0,
 * Python/Cython) would bind with ctypes/cffi/Cython `cdef extern` to run `Clustering.fit`
 * on a B200.  Plain pointers and sizes only; no torch / C++ types.  Citations below name the
 * reference interface (path relative to the reference root) each entry point replaces.
 *
 * The reference drives its components through per-point virtual calls
 * (`_get(index)`, `_check(na, nb)`; src/cnnclustering/_types.pxd:144-251).  That granularity
 * cannot survive on a GPU, so every entry point here is a BULK version of a reference role:
 * the whole neighbourhood structure of a data set, all candidate edges, all labels.
 *
 * Conventions
 *   - every function returns 0 on success, a negative CNNB200_E_* code otherwise;
 *     cnnb200_last_error() returns a thread-local message for the last failure.
 *   - `loc` arguments say where a caller buffer lives: CNNB200_HOST or CNNB200_DEVICE.
 *   - element types at the boundary are the reference's (_primitive_types.pxd:4-7):
 *     coordinates double, indices/labels int64_t, consider uint8_t.
 *   - a context is bound to one CUDA device and one stream; it is not thread-safe.
 *     All work is enqueued on the context's stream; calls that return data in HOST
 *     buffers synchronise that stream before returning.
 *   - device scratch (CSR, parents, ...) is owned by the context and reused across calls;
 *     it is obtained through the allocator hook (cudaMalloc by default; the Python host
 *     layer routes it to torch's caching allocator).
 */
#ifndef CNNB200_H
#define CNNB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNNB200_VERSION 100 /* 0.1.0 */

#define CNNB200_HOST 0
#define CNNB200_DEVICE 1

/* Metric ids: MetricExtEuclideanReduced (_types.pyx:2250-2287, recipe key "euclidean_r") and
 * MetricExtEuclidean (_types.pyx:2170-2207, "euclidean"). */
#define CNNB200_METRIC_EUCLIDEAN_REDUCED 0
#define CNNB200_METRIC_EUCLIDEAN 1
/* MetricExtEuclideanPeriodicReduced (_types.pyx:2290-2348, "euclidean_periodic_r"): per-dimension period `bounds[k]`
 * (0 = not periodic), set with cnnb200_ctx_set_periodic_bounds.  Methods: grid (d <= 3: the cell list on the point set
 * augmented by its periodic images as a filter, then the reference's arithmetic on the original coordinates; needs every
 * period >= two radii, else the request falls through to tiled) and tiled (all pairs, any d); auto = grid from 4096
 * points on when d <= 3. */
#define CNNB200_METRIC_EUCLIDEAN_PERIODIC_REDUCED 2

#define CNNB200_OK 0
#define CNNB200_E_INVALID -1   /* bad argument                                  */
#define CNNB200_E_CUDA -2      /* a CUDA runtime call failed                    */
#define CNNB200_E_NOMEM -3     /* device or host allocation failed              */
#define CNNB200_E_STATE -4     /* call order: no neighbourhoods / components yet */
#define CNNB200_E_RANGE -5     /* size exceeds an implementation limit          */

typedef struct cnnb200_ctx cnnb200_ctx;

/* Allocator hook: alloc returns a device pointer for `bytes` on the context's device
 * (NULL on failure); free releases it.  `user` is passed through. */
typedef void *(*cnnb200_alloc_fn)(size_t bytes, void *user);
typedef void (*cnnb200_free_fn)(void *ptr, void *user);

int cnnb200_version(void);
const char *cnnb200_last_error(void);

int cnnb200_ctx_create(int device, cnnb200_ctx **out);
int cnnb200_ctx_destroy(cnnb200_ctx *ctx);
/* `stream` is a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); NULL = default. */
int cnnb200_ctx_set_stream(cnnb200_ctx *ctx, void *stream);
int cnnb200_ctx_set_allocator(cnnb200_ctx *ctx, cnnb200_alloc_fn alloc, cnnb200_free_fn free_,
                              void *user);
int cnnb200_ctx_synchronize(cnnb200_ctx *ctx);
/* bounds: HOST double [d], the constructor argument of MetricExtEuclideanPeriodicReduced (_types.pyx:2293-2294). */
int cnnb200_ctx_set_periodic_bounds(cnnb200_ctx *ctx, const double *bounds, int64_t d);
/* Releases all device scratch held by the context (it is re-acquired on demand). */
int cnnb200_ctx_release(cnnb200_ctx *ctx);

/* ---------------------------------------------------------------------------------------
 * Radius-neighbour search -> sorted CSR held by the context.
 * Replaces NeighboursGetterExtBruteForce._get (_types.pyx:1692-1710) over
 * DistanceGetterExtMetric._get_single (:1979-1985) and Metric._calc_distance (:2253-2267 /
 * :2173-2187) for ALL points at once.  `radius_cutoff` is already metric-adjusted
 * (ClusterParameters.radius_cutoff, _types.pxd:119-125: r*r for the reduced metric).
 * Neighbour sets are exactly the reference's: the acceptance test is the same f64
 * expression (sequential sum of squared differences, no FMA, `<=`); self is included.
 *
 * x: row-major double [n, d] (InputDataExtComponentsMemoryview layout, _types.pyx:1040-1059).
 * method: 0 = auto (grid for d <= 3, tiled otherwise), 1 = uniform-grid cell list (d <= 3),
 *         2 = tiled all-pairs in f64, 3 = tiled tensor-core filter + f64 boundary recheck.
 * The CSR lives in an internal index space (cell-sorted for the grid method) together with
 * the permutation back to the caller's indices; cnnb200_neighbours_export converts.
 * nnz_out (optional): total number of neighbour entries.
 * ------------------------------------------------------------------------------------- */
int cnnb200_neighbours_from_points(cnnb200_ctx *ctx, const double *x, int loc, int64_t n,
                                   int64_t d, double radius_cutoff, int metric, int method,
                                   int64_t *nnz_out);

/* ---------------------------------------------------------------------------------------
 * Row-sharded form of the cell-list search, for one process per GPU (coordinates replicated,
 * query rows sharded).  Rows are the points in cell-sorted order.
 *   grid_begin    every rank: grid geometry, sort by cell, candidate-list prefix
 *   grid_balance  row-block boundaries bounds[0..parts] (HOST) that balance the candidate pairs
 *   grid_halo     [lo, hi) = the block plus every row that can occur in a neighbour list of the block
 *                 (contiguous in the cell-sorted order).  A rank that searches [lo, hi) itself holds all
 *                 rows its block's common-neighbour checks read and needs no CSR exchange at all
 *                 (rows outside get degree 0; cnnb200_components_merge adopts their core flags).
 *   grid_degrees  this rank's block: deg[row_begin..row_end) (DEVICE int32, indexed by row)
 *   (all ranks exchange the degrees; row_ptr = exclusive prefix sum over all rows)
 *   grid_fill     this rank's block written into col_idx at row_ptr[row] (DEVICE buffers that
 *                 span ALL rows: the block lands directly in its slice of the gathered CSR)
 *   (all ranks exchange the col_idx slices)
 *   neighbours_adopt  use the gathered CSR (borrowed DEVICE buffers); grid_perm != 0 attaches
 *                 the cell-sort permutation of this context's grid.
 * ------------------------------------------------------------------------------------- */
int cnnb200_grid_begin(cnnb200_ctx *ctx, const double *x, int loc, int64_t n, int64_t d,
                       double radius_cutoff, int metric);
int cnnb200_grid_balance(cnnb200_ctx *ctx, int parts, int64_t *bounds);
int cnnb200_grid_halo(cnnb200_ctx *ctx, int64_t row_begin, int64_t row_end, int64_t *lo, int64_t *hi);
int cnnb200_grid_degrees(cnnb200_ctx *ctx, int64_t row_begin, int64_t row_end, int32_t *deg);
int cnnb200_grid_fill(cnnb200_ctx *ctx, int64_t row_begin, int64_t row_end, const int64_t *row_ptr,
                      int32_t *col_idx);
int cnnb200_neighbours_adopt(cnnb200_ctx *ctx, const int64_t *row_ptr, const int32_t *col_idx,
                             int64_t n, int64_t nnz, int64_t max_deg, int grid_perm);

/* Precomputed neighbourhoods, CSR form (row_ptr int64 [n+1]; col_idx int32 or int64 [nnz],
 * `idx_bytes` = 4 or 8), rows ascending.  Replaces NeighboursGetterExtLookup._get
 * (_types.pyx:1806-1817) for all rows.  With loc == CNNB200_DEVICE, idx_bytes == 4 and
 * borrow != 0 the buffers are used in place (caller keeps them alive); otherwise copied. */
int cnnb200_neighbours_from_csr(cnnb200_ctx *ctx, const int64_t *row_ptr, const void *col_idx,
                                int idx_bytes, int loc, int borrow, int64_t n, int64_t nnz);

/* Precomputed neighbourhoods in the reference's own container layout: zero-padded int64
 * [n, max_deg] plus int64 counts [n] (InputDataExtNeighbourhoodsMemoryview,
 * _types.pyx:1153-1217; produced by recipes.prepare_neighbourhoods, recipes.py:150-179). */
int cnnb200_neighbours_from_padded(cnnb200_ctx *ctx, const int64_t *padded,
                                   const int64_t *counts, int loc, int64_t n, int64_t max_deg);

/* Neighbourhoods from a precomputed distance matrix, f64 [n, n] row-major, HOST or DEVICE:
 * j is a neighbour of i when dist[i*n + j] <= radius_cutoff (NaN never passes; the diagonal decides
 * whether a point counts itself).  Replaces NeighboursGetterExtBruteForce._get (_types.pyx:1692-1710)
 * over MetricExtPrecomputed._calc_distance (_types.pyx:2108-2127), i.e. the reference's "distances"
 * recipe (recipes.py:329-339), for all rows at once.  The matrix is read once. */
int cnnb200_neighbours_from_distances(cnnb200_ctx *ctx, const double *dist, int loc, int64_t n,
                                      double radius_cutoff);

/* The same from CONDENSED distances: the N (N - 1) / 2 values d(a, b), a < b, row by row (d(a, b) at
 * a (N - 1) - a (a - 1) / 2 + (b - a - 1), d(a, a) = 0) -- the layout of InputDataExtDistancesLinearMemoryview
 * (_types.pyx:1112-1150, `_get_distance`).  dist: f64 [n (n - 1) / 2] in `loc` (may be NULL for n = 1); read once. */
int cnnb200_neighbours_from_condensed(cnnb200_ctx *ctx, const double *dist, int loc, int64_t n,
                                      double radius_cutoff);

/* Brings the rows of the held neighbourhood structure into ascending order when they are not
 * (the reference's "neighbourhoods" recipe, recipes.py:340-348, takes lists in any order:
 * NeighboursGetterExtLookup with is_sorted=False, _types.pyx:1791-1804).  A borrowed col_idx is left
 * untouched (the sorted copy lives in the context).  *sorted_out (optional) = 1 when a sort was needed.
 * Limit: nnz < 2^31-1 for unsorted input. */
int cnnb200_neighbours_sort_rows(cnnb200_ctx *ctx, int *sorted_out);

/* Sizes of the neighbourhood structure currently held. */
/* Number of entries (i, j) whose mirror (j, i) is missing.  The component pass reads candidate edges as the entries
 * j > i of row i, which equals the reference's BFS (_fit.pyx:434-491) only for SYMMETRIC lists -- radius
 * neighbourhoods are; for lists of other origin (kNN) this is the check to run first. */
int cnnb200_neighbours_asymmetry(cnnb200_ctx *ctx, int64_t *one_sided_out);
int cnnb200_neighbours_info(cnnb200_ctx *ctx, int64_t *n, int64_t *nnz, int64_t *max_deg);

/* Copies the neighbourhoods out in the CALLER's index space, rows ascending:
 * row_ptr int64 [n+1], col_idx int64 [nnz] (either may be NULL to skip). */
int cnnb200_neighbours_export(cnnb200_ctx *ctx, int64_t *row_ptr, int64_t *col_idx, int loc);

/* Raw view of the internal CSR (device pointers, valid until the next neighbours_* call):
 * row_ptr int64 [n+1], col_idx int32 [nnz], perm int32 [n] (internal -> caller index; NULL
 * when the internal space is the caller's). */
int cnnb200_neighbours_device_view(cnnb200_ctx *ctx, const int64_t **row_ptr,
                                   const int32_t **col_idx, const int32_t **perm);

/* ---------------------------------------------------------------------------------------
 * Common-neighbour similarity, bulk.  Replaces SimilarityCheckerExt*._get
 * (_types.pyx:2541-2556, :2612-2634, :2708-2753): counts[k] = |N(i) ∩ N(j)| for the k-th CSR
 * entry (row i, column j) when i < j and both rows have more than n_member_cutoff entries
 * (`_enough`, _types.pyx:1408-1411); every other slot gets 0xFFFFFFFF.
 * counts: uint32 [nnz] in the internal CSR order, DEVICE or HOST.
 * n_edges_out (optional): number of evaluated (i<j, both core) edges.
 * ------------------------------------------------------------------------------------- */
int cnnb200_cn_count(cnnb200_ctx *ctx, int64_t n_member_cutoff, uint32_t *counts, int loc,
                     int64_t *n_edges_out);
/* The same for the rows [row_begin, row_end) only (the slots of the other rows are left untouched): one rank's share
 * of a row-sharded run -- the rows of its block and of every neighbour of them must be held (halo construction). */
int cnnb200_cn_count_rows(cnnb200_ctx *ctx, int64_t n_member_cutoff, int64_t row_begin, int64_t row_end,
                          uint32_t *counts, int loc, int64_t *n_edges_out);
/* SimilarityChecker._check (early exit) for arbitrary pairs of rows of the held neighbourhoods, in the INTERNAL index
 * space: pairs int32 [n_pairs][2], out uint8 [n_pairs].  Bit 0 = the per-lane two-pointer merge the sampling round
 * uses (SimilarityCheckerExtScreensorted._check, _types.pyx:2654-2706), bit 1 = the whole-warp window form of the work
 * list; both equal |N(i) ∩ N(j)| >= similarity_cutoff (0: always true, _types.pyx:2589-2590). */
int cnnb200_cn_check_pairs(cnnb200_ctx *ctx, const int32_t *pairs, int64_t n_pairs, int64_t similarity_cutoff,
                           uint8_t *out, int loc);

/* ---------------------------------------------------------------------------------------
 * Density-connected components.  Replaces FitterExtBFS._fit (_fit.pyx:408-491) including its
 * use of SimilarityCheckerExt*._check (_types.pyx:2580-2610 / :2654-2706) and `_enough`.
 *
 * cnnb200_fit: labels int64 [n] in/out, consider uint8 [n] in/out (NULL = all points
 *   considered = a fresh fit).  Semantics are exactly the reference's: point i is a core
 *   point when consider[i] and |N(i)| > n_member_cutoff; core points i,j are linked when j is
 *   in N(i) and |N(i) ∩ N(j)| >= similarity_cutoff (similarity_cutoff == 0 links always);
 *   components are numbered current_start, current_start+1, ... by ascending smallest member
 *   index and written to labels[]; labels of non-core points are left as they are;
 *   consider[] is zeroed.  With consider == NULL and loc == HOST, labels need not be
 *   initialised (treated as all zero, as Clustering.fit does, cluster.pyx:316-326).
 *   n_clusters_out (optional): number of components numbered.
 *
 * The three-step form exists for row-sharded multi-GPU runs:
 *   components_begin   -> every rank (full neighbourhoods on every rank)
 *   components_edges   -> each rank on its own row block [row_begin, row_end)
 *   (all-gather of the parent arrays, cnnb200_components_parents / _merge)
 *   components_labels  -> numbering + label write-out
 * ------------------------------------------------------------------------------------- */
int cnnb200_fit(cnnb200_ctx *ctx, int64_t n_member_cutoff, int64_t similarity_cutoff,
                int64_t current_start, int64_t *labels, uint8_t *consider, int loc,
                int64_t *n_clusters_out);

int cnnb200_components_begin(cnnb200_ctx *ctx, int64_t n_member_cutoff,
                             int64_t similarity_cutoff, const uint8_t *consider, int loc);
int cnnb200_components_edges(cnnb200_ctx *ctx, int64_t row_begin, int64_t row_end);
/* The two halves of components_edges on their own: phase 1 = the sampling rounds (every row tries its first entry
 * above it, every k-th row of a tree two far entries), phase 2 = the scan of all candidate edges of the rows + the
 * work list.  A row-sharded run on lists WITHOUT index locality (precomputed neighbourhoods in the caller's order:
 * most candidate edges cross the row blocks) all-gathers and merges the forests between the two
 * (cnnb200_components_merge), so that each rank's scan skips what any rank's samples already joined. */
int cnnb200_components_phase(cnnb200_ctx *ctx, int64_t row_begin, int64_t row_end, int phase);
/* Parameter sweeps (the reference's scan loop calls Clustering.fit once per (radius_cutoff, cnn_cutoff)
 * cell, docsrc/benchmark/scans.ipynb; HierarchicalFitterRepeat._fit does the same per step,
 * _fit.pyx:1058-1187).  For one radius the counts of cnnb200_cn_count answer EVERY cnn_cutoff:
 * after components_begin(n_member_cutoff, similarity_cutoff) this links the slots with
 * counts >= similarity_cutoff between rows with more than n_member_cutoff entries -- the same components
 * components_edges finds.  counts: DEVICE uint32 [nnz], computed with a member cutoff <= n_member_cutoff. */
int cnnb200_components_from_counts(cnnb200_ctx *ctx, const uint32_t *counts);
/* Device pointer to the int32 [n] parent array (internal index space). */
int cnnb200_components_parents(cnnb200_ctx *ctx, int32_t **parents);
/* Unions (i, other[i]) for every i of `count` foreign parent arrays laid out back to back
 * (int32 [count, n], DEVICE).  A row some rank reports as a core point (other[i] >= 0) that this
 * context holds as empty (halo construction) becomes a core point here as well. */
int cnnb200_components_merge(cnnb200_ctx *ctx, const int32_t *others, int64_t count);
int cnnb200_components_labels(cnnb200_ctx *ctx, int64_t current_start, int64_t *labels, int loc,
                              int fresh, int64_t *n_clusters_out);
/* components_labels on fresh labels followed by Labels.sort_by_size (_types.pyx:212-286) as ONE ranking of the
 * components (sizes, smallest members, the survivors of member_cutoff sorted by one CTA); what a cell of a parameter
 * scan and the end of a row-sharded fit call.  n_clusters_out: clusters before the size sort. */
int cnnb200_components_labels_sorted(cnnb200_ctx *ctx, int64_t current_start, int64_t member_cutoff,
                                     int64_t max_clusters, int64_t *labels, int loc, int64_t *n_clusters_out);

/* ---------------------------------------------------------------------------------------
 * The row-sharded fit on coordinates (d <= 3) as two calls per rank with ONE collective between them
 * (replaces FitterExtBFS._fit, _fit.pyx:408-491, plus Labels.sort_by_size,
 * _types.pyx:212-286, for one process per GPU):
 *   fit_sharded_begin  cell sort (replicated), row blocks bounds[0..world] (HOST out) balanced by candidate
 *                      pairs, block `rank` plus its halo searched, the block's candidate edges settled on a
 *                      private parent array; *parents = DEVICE int32 [n] to all-gather.
 *   (all-gather of the parent arrays: [world, n] int32)
 *   fit_sharded_end    merge of the forests (identical on every rank), then -- where `labels` is not NULL --
 *                      the ranking of the components and the labels (sort_by_size != 0: size-sorted as
 *                      cnnb200_fit_points delivers them; 0: numbered from current_start as the BFS does).
 * ------------------------------------------------------------------------------------- */
int cnnb200_fit_sharded_begin(cnnb200_ctx *ctx, const double *x, int loc, int64_t n, int64_t d,
                              double radius_cutoff, int metric, int64_t n_member_cutoff,
                              int64_t similarity_cutoff, int rank, int world, int64_t *bounds,
                              int32_t **parents);
int cnnb200_fit_sharded_end(cnnb200_ctx *ctx, const int32_t *others, int64_t world, int sort_by_size,
                            int64_t current_start, int64_t member_cutoff, int64_t max_clusters,
                            int64_t *labels, int loc, int64_t *n_clusters_out);
/* The exchange restricted to what the ranks really built.  Rank r's parent array says something only for the rows
 * [lo_r, hi_r) of its block plus halo (everything else is -1):
 *   fit_sharded_ranges       ranges[2 r], ranges[2 r + 1] = lo_r, hi_r of ALL ranks (HOST out, identical on every rank)
 *                            after fit_sharded_begin; the caller picks len = max (hi_r - lo_r) and
 *                            starts[r] = min(lo_r, n - len) and all-gathers parents[starts[rank] .. + len).
 *   fit_sharded_end_windows  as fit_sharded_end with others = [world, len] int32 (DEVICE) and starts (HOST [world]). */
int cnnb200_fit_sharded_ranges(cnnb200_ctx *ctx, int64_t world, int64_t *ranges);
int cnnb200_fit_sharded_end_windows(cnnb200_ctx *ctx, const int32_t *others, int64_t world, const int64_t *starts,
                                    int64_t len, int sort_by_size, int64_t current_start, int64_t member_cutoff,
                                    int64_t max_clusters, int64_t *labels, int loc, int64_t *n_clusters_out);

/* ---------------------------------------------------------------------------------------
 * Prediction for points OUTSIDE the clustered set.  Replaces PredictorFirstmatch.predict
 * (_fit.pyx:1231-1291) with NeighboursGetterExtBruteForce._get_other (_types.pyx:1712-1731) or
 * NeighboursGetterExtLookup._get_other (:1819-1832).
 *
 *   neighbours_other_from_points  N_o(p) = clustered points within the cutoff of other point p, for all p.
 *       Needs cnnb200_neighbours_from_points on the clustered set first (same cutoff and metric are used;
 *       methods grid and tiled; with device input the caller keeps the coordinates alive).
 *   neighbours_other_from_csr     the same, precomputed: row_ptr int64 [m+1], col_idx int64 [nnz] (indices
 *       into the clustered set, ascending); the clustered set's own neighbourhoods were given by
 *       cnnb200_neighbours_from_csr / _from_padded.
 *   predict   labels int64 [n] of the clustered set; label_ok uint8 [max_label+1] marks the labels of the
 *       consider set (Clustering.predict's `clusters`, cluster.pyx:621-624).  other_labels int64 [m] and
 *       other_consider uint8 [m] are updated in place exactly like the reference: a point with
 *       other_consider != 0 and more than n_member_cutoff neighbours gets the label of the first member m
 *       (ascending index) of N_o(p) whose label is in the consider set, which has more than n_member_cutoff
 *       neighbours itself and |N_o(p) ∩ N(m)| >= similarity_cutoff; its other_consider becomes 0.
 * ------------------------------------------------------------------------------------- */
int cnnb200_neighbours_other_from_points(cnnb200_ctx *ctx, const double *x_other, int loc, int64_t m, int64_t d,
                                         int64_t *nnz_out);
int cnnb200_neighbours_other_from_csr(cnnb200_ctx *ctx, const int64_t *row_ptr, const int64_t *col_idx, int loc,
                                      int64_t m, int64_t nnz);
int cnnb200_predict(cnnb200_ctx *ctx, int64_t n_member_cutoff, int64_t similarity_cutoff, const int64_t *labels,
                    const uint8_t *label_ok, int64_t max_label, int64_t *other_labels, uint8_t *other_consider,
                    int loc);

/* ---------------------------------------------------------------------------------------
 * Labels.sort_by_size (_types.pyx:212-286), in place on labels int64 [n] (HOST or DEVICE):
 * clusters ranked by size descending, ties by first occurrence (Counter.most_common order);
 * rank > max_clusters -> 0; size < member_cutoff -> 0; survivors renumbered 1, 2, ...
 * member_cutoff < 0 means None (2); max_clusters < 0 means None.
 * map_out (optional, HOST int64 [max_label+1], sized by the caller via map_len): old -> new.
 * ------------------------------------------------------------------------------------- */
int cnnb200_sort_by_size(cnnb200_ctx *ctx, int64_t *labels, int loc, int64_t n,
                         int64_t member_cutoff, int64_t max_clusters, int64_t *map_out,
                         int64_t map_len);

/* ---------------------------------------------------------------------------------------
 * One call for the whole of `Clustering.fit` on coordinates (cluster.pyx:256-392 with the
 * "coordinates" recipe, recipes.py:318-328): neighbour search, fit, optional sort_by_size.
 * x and labels in HOST or DEVICE memory (`loc`).  Fresh fit (labels fully overwritten).
 * ------------------------------------------------------------------------------------- */
int cnnb200_fit_points(cnnb200_ctx *ctx, const double *x, int loc, int64_t n, int64_t d,
                       double radius_cutoff, int metric, int method, int64_t n_member_cutoff,
                       int64_t similarity_cutoff, int64_t current_start, int sort_by_size,
                       int64_t member_cutoff, int64_t max_clusters, int64_t *labels,
                       int64_t *n_clusters_out);
/* The same for the neighbourhoods the context already holds (cnnb200_neighbours_from_csr / _from_padded /
 * _from_distances / _from_points): FitterExtBFS.fit on fresh labels + Labels.sort_by_size as one call. */
int cnnb200_fit_sorted(cnnb200_ctx *ctx, int64_t n_member_cutoff, int64_t similarity_cutoff, int64_t current_start,
                       int64_t member_cutoff, int64_t max_clusters, int64_t *labels, int loc,
                       int64_t *n_clusters_out);

/* (clusters kept, members of the largest cluster, noise points) after the last cnnb200_sort_by_size /
 * cnnb200_fit_points with size sorting: what Clustering.fit stores in its Record (cluster.pyx:366-390), HOST int64 [3]. */
int cnnb200_label_summary(cnnb200_ctx *ctx, int64_t *out3);

/* Grouping of the points by (previous label, current label) for the hierarchy driver: replaces the per-point
 * loop of HierarchicalFitterRepeat._fit (_fit.pyx:1150-1160, `parent_labels_map`) and the per-parent grouping of
 * Bundle.isolate (_bundle.pyx:275-352) for ALL parents of a step at once.
 *   order[k]  point indices sorted by (prev, cur, index)
 *   rank[k]   position of point order[k] among the points with the same previous label, in index order
 * The members of (parent p, child c) are one contiguous slice of order; the same slice of rank gives their places in
 * the parent's own label array.  prev / cur / order / rank: int64 [n], all in `loc`. */
int cnnb200_labels_group(cnnb200_ctx *ctx, const int64_t *prev, const int64_t *cur, int loc, int64_t n,
                         int64_t *order, int64_t *rank);

/* Page-lock / release a caller array in place (no copy is kept): repeated fits on the same coordinates then
 * transfer at PCIe speed. */
int cnnb200_host_register(void *ptr, size_t bytes);
int cnnb200_host_unregister(void *ptr);

/* Counters of the last fit / search, for measurement (all optional, HOST):
 *   stats[0] candidate edges (i<j, both core)      stats[1] CN checks executed
 *   stats[2] edges accepted by a check             stats[3] edges skipped (already connected)
 *   stats[4] distance evaluations (neighbour search; tcgen05 filter: the pairs of the 128 x 128 tiles that were visited)
 *   stats[5] pairs re-checked in f64 (fp32 / tensor-core filter paths)
 *   stats[6] count mode: rows whose index window did not fit the shared-memory bitmap
 *   stats[7] which kernels served the last calls: bits 0-3 = method of the last cnnb200_neighbours_from_points
 *            (1 cell list, 2 tiled f64, 3 tcgen05 filter), bit 4 = the last count-mode pass ran on bit-packed rows */
int cnnb200_stats(cnnb200_ctx *ctx, int64_t *stats, int64_t len);

/* Stage timing by CUDA events on the context's stream (off by default).  cnnb200_timings
 * synchronises the stream, folds the event pairs recorded since the last call into per-stage sums
 * and returns: ms[k] / calls[k] for stage k (CNNB200_T_* below, `len` entries each, either may be
 * NULL) and the number of hand-written kernels launched since the last reset. */
#define CNNB200_T_H2D 0
#define CNNB200_T_GRID_BUILD 1
#define CNNB200_T_SEARCH_COUNT 2
#define CNNB200_T_SCAN 3
#define CNNB200_T_SEARCH_FILL 4
#define CNNB200_T_INGEST 5
#define CNNB200_T_COMP_INIT 6
#define CNNB200_T_COMP_SAMPLE 7
#define CNNB200_T_COMP_EDGES 8
#define CNNB200_T_COMP_FLATTEN 9
#define CNNB200_T_LABELS 10
#define CNNB200_T_SORT_BY_SIZE 11
#define CNNB200_T_D2H 12
#define CNNB200_T_CN_COUNT 13
#define CNNB200_T_COUNT 14
int cnnb200_ctx_set_timing(cnnb200_ctx *ctx, int enabled);
int cnnb200_timings(cnnb200_ctx *ctx, double *ms, int64_t *calls, int64_t len, int64_t *launches,
                    int reset);

#ifdef __cplusplus
}
#endif
#endif /* CNNB200_H */
