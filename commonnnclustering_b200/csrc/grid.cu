// grid.cu -- K1/K2: uniform-grid (cell list) radius-neighbour search for d <= 3 -> sorted CSR.
//
// Bulk replacement of NeighboursGetterExtBruteForce._get (reference _types.pyx:1692-1710).
// The acceptance test is the reference's own f64 expression (Metric._calc_distance,
// _types.pyx:2253-2267: sequential sum of squared differences, `<=` against the adjusted
// radius), evaluated with explicitly rounded intrinsics so that nvcc cannot contract to FMA.
// The grid only generates candidates, so the neighbour SETS equal the brute-force ones.
//
// Index space: points are reordered by cell (x fastest); the CSR is emitted in that
// cell-sorted space (rows ascending without any per-row sort because cells are visited in
// ascending id) together with perm[] (sorted position -> caller index).
#include <cub/cub.cuh>
#include <math.h>

#include <memory>

#include "common.cuh"


// ---- bounding box -------------------------------------------------------------------------
__global__ void bbox_partial_kernel(const double *__restrict__ x, ll n, int d, double *partial) {
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x)
        for (int k = 0; k < d; ++k) {
            double v = x[i * d + k];
            mn[k] = fmin(mn[k], v);
            mx[k] = fmax(mx[k], v);
        }
    __shared__ double s[6][8];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o; o >>= 1) {
            mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
        if (lane == 0) { s[k][w] = mn[k]; s[3 + k][w] = mx[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        int k = threadIdx.x;
        double v = s[k][0];
        for (int j = 1; j < (int)(blockDim.x >> 5); ++j) v = k < 3 ? fmin(v, s[k][j]) : fmax(v, s[k][j]);
        partial[blockIdx.x * 6 + k] = v;
    }
}

// one warp per component (3 minima, 3 maxima)
__global__ void bbox_final_kernel(const double *partial, int nblocks, double *out) {
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (k >= 6) return;
    double v = k < 3 ? INFINITY : -INFINITY;
    for (int b = lane; b < nblocks; b += 32) v = k < 3 ? fmin(v, partial[b * 6 + k]) : fmax(v, partial[b * 6 + k]);
    for (int o = 16; o; o >>= 1) {
        const double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = k < 3 ? fmin(v, w) : fmax(v, w);
    }
    if (lane == 0) out[k] = v;
}

// ---- cell assignment ------------------------------------------------------------------------
__device__ __forceinline__ int cell_coord(double v, double lo, double h, int dim) {
    int c = (int)floor(__ddiv_rn(__dsub_rn(v, lo), h));
    return min(max(c, 0), dim - 1);
}

// ORDERED: idx[i] = i for the stable radix sort that follows (the cell-sorted order is then a function of the data
// alone -- what the ranks of a row-sharded run need to agree on); otherwise idx[i] = the point's arrival slot in its
// cell, and scatter_kernel places it at cell_start[cell] + slot: a counting sort, no radix passes.  The order inside
// a cell then depends on the schedule; neighbour sets, intersections and labels do not (rows are ascending in
// whatever internal order there is, only the numbering maps back through perm[]).
template <bool ORDERED>
__global__ void cell_id_kernel(const double *__restrict__ x, ll n, GridDesc g, int *cell_id, int *idx,
                               int *cell_count) {
    ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c[3] = {0, 0, 0};
    for (int k = 0; k < g.d; ++k) c[k] = cell_coord(x[i * g.d + k], g.lo[k], g.h, g.dims[k]);
    int cid = (c[2] * g.dims[1] + c[1]) * g.dims[0] + c[0];
    cell_id[i] = cid;
    const int slot = atomicAdd(&cell_count[cid], 1);
    idx[i] = ORDERED ? (int)i : slot;
}

// Origin of cell index c along one axis.  The SAME expression is used wherever a cell-local
// coordinate is formed, so that differences of cell-local coordinates are differences of the true
// coordinates up to the stated rounding (see the filter kernel below).
__device__ __forceinline__ double cell_origin(double lo, double h, int c) {
    return __dadd_rn(lo, __dmul_rn((double)c, h));
}

// gather coordinates into cell-sorted SoA: xs[k][p] = x[perm[p]][k]; and the fp32 cell-local copy used by
// the filter kernel: xf[p] = (x relative to lattice A, x relative to lattice B, y and z relative to the
// point's own cell).  Lattices A and B have period 16 cells along x and are offset by 8 cells: any run of
// up to 8 neighbouring cells lies inside one period of A or of B, so that all candidates of a run and
// the queries share one origin and their fp32 differences are accurate to ~2^-24 of 16h.
__global__ void gather_kernel(const double *__restrict__ x, const int *__restrict__ perm,
                              const int *__restrict__ cell_sorted, ll n, GridDesc g, double *xs, float4 *xf) {
    ll p = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int d = g.d;
    ll src = perm[p];
    double v[3] = {0.0, 0.0, 0.0};
    for (int k = 0; k < d; ++k) {
        v[k] = x[src * d + k];
        xs[k * n + p] = v[k];
    }
    if (!xf) return;
    const int c = cell_sorted[p];
    const int cx = c % g.dims[0];
    const int t = c / g.dims[0];
    const int cy = t % g.dims[1];
    const int cz = t / g.dims[1];
    float4 f;
    f.x = (float)__dsub_rn(v[0], cell_origin(g.lo[0], g.h, cx & ~15));
    f.y = (float)__dsub_rn(v[0], cell_origin(g.lo[0], g.h, ((cx + 8) & ~15) - 8));
    f.z = d > 1 ? (float)__dsub_rn(v[1], cell_origin(g.lo[1], g.h, cy)) : 0.f;
    f.w = d > 2 ? (float)__dsub_rn(v[2], cell_origin(g.lo[2], g.h, cz)) : 0.f;
    xf[p] = f;
}

// counting-sort form of the same: thread i (caller order, coalesced reads of x) writes point i to its place
__global__ void scatter_kernel(const double *__restrict__ x, const int *__restrict__ cell_id, const int *__restrict__ slot,
                               const int *__restrict__ cell_start, ll n, GridDesc g, int *__restrict__ perm,
                               int *__restrict__ cell_sorted, double *xs, float4 *xf) {
    ll i = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = g.d;
    const int c = cell_id[i];
    const ll p = (ll)cell_start[c] + slot[i];
    perm[p] = (int)i;
    cell_sorted[p] = c;
    double v[3] = {0.0, 0.0, 0.0};
    for (int k = 0; k < d; ++k) {
        v[k] = x[i * d + k];
        xs[k * n + p] = v[k];
    }
    if (!xf) return;
    const int cx = c % g.dims[0];
    const int t = c / g.dims[0];
    const int cy = t % g.dims[1];
    const int cz = t / g.dims[1];
    float4 f;
    f.x = (float)__dsub_rn(v[0], cell_origin(g.lo[0], g.h, cx & ~15));
    f.y = (float)__dsub_rn(v[0], cell_origin(g.lo[0], g.h, ((cx + 8) & ~15) - 8));
    f.z = d > 1 ? (float)__dsub_rn(v[1], cell_origin(g.lo[1], g.h, cy)) : 0.f;
    f.w = d > 2 ? (float)__dsub_rn(v[2], cell_origin(g.lo[2], g.h, cz)) : 0.f;
    xf[p] = f;
}

// ---- K2: one warp per query row ---------------------------------------------------------------
template <int D, bool SQRT>
__device__ __forceinline__ bool within(double qx, double qy, double qz, const double *__restrict__ xs,
                                       ll n, int q, double rc) {
    double t = __dsub_rn(qx, xs[q]);
    double s = __dmul_rn(t, t);
    if (D > 1) {
        t = __dsub_rn(qy, xs[n + q]);
        s = __dadd_rn(s, __dmul_rn(t, t));
    }
    if (D > 2) {
        t = __dsub_rn(qz, xs[2 * n + q]);
        s = __dadd_rn(s, __dmul_rn(t, t));
    }
    if (SQRT) s = __dsqrt_rn(s);
    return s <= rc;
}

// One warp per tile of 32 consecutive cell-sorted points.  Within the tile, points of the same
// cell form a group of g contiguous lanes that shares one candidate list (the 3^D cells around
// it = 3^(D-1) contiguous index runs).  The 32 lanes are laid out as gp x k (gp = g rounded up to
// a power of two, k = 32 / gp): lane = sub * gp + jq tests query jq against candidate q0 + sub,
// so dense cells (g = 32, k = 1: one candidate broadcast to 32 queries) and isolated points
// (g = 1, k = 32: one query against 32 candidates) both keep the lanes busy.
//   count pass: per-lane hit counts, summed over `sub` by shuffles -> deg[]
//   fill pass : hits of one step are ranked with a ballot restricted to the lanes of the same
//               query (mask rep << jq); candidates ascend with the lane index, so rows come out
//               ascending in the cell-sorted index space.
__constant__ unsigned c_rep[6] = {0xffffffffu, 0x55555555u, 0x11111111u, 0x01010101u, 0x00010001u, 0x00000001u};

// MODE 0: count pass (deg[] only)            } the two-pass form, used when the single-pass
// MODE 1: fill pass (recomputes the distances) } scratch would not fit
// MODE 2: single pass -- hits are written straight into a gapped buffer in which row p owns
//         cap[p] slots (cap = size of its candidate list >= its degree) and deg[] comes out of
//         the same pass; compact_rows_kernel then squeezes the rows into the CSR.
template <int D, bool SQRT, int MODE>
__global__ void __launch_bounds__(256)
grid_rows_kernel(const double *__restrict__ xs, ll n, const int *__restrict__ cell_sorted,
                 const int *__restrict__ cell_start, GridDesc g, double rc, int *__restrict__ deg,
                 const ll *__restrict__ row_ptr, int *__restrict__ col_idx, ll p_begin, ll p_end) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll p0 = p_begin + warp * 32; p0 < p_end; p0 += nwarps * 32) {
        ll p = p0 + lane;
        int cid = p < p_end ? cell_sorted[p] : -1;
        unsigned todo = __ballot_sync(full, p < p_end);
        while (todo) {
            const int a = __ffs(todo) - 1;  // first lane of the next group
            const int c = __shfl_sync(full, cid, a);
            const unsigned grp = __ballot_sync(full, cid == c) & todo;
            todo &= ~grp;
            const int gsz = __popc(grp);
            const int lg = gsz <= 1 ? 0 : 32 - __clz(gsz - 1);
            const int gp = 1 << lg, k = 32 >> lg;
            const int jq = lane & (gp - 1), sub = lane >> lg;
            const bool qvalid = jq < gsz;
            const ll pq = p0 + a + (qvalid ? jq : 0);
            const double qx = xs[pq], qy = D > 1 ? xs[n + pq] : 0.0, qz = D > 2 ? xs[2 * n + pq] : 0.0;
            const unsigned qm = c_rep[lg] << jq;
            const ll base = MODE ? row_ptr[pq] : 0;
            const int cx = c % g.dims[0];
            const int t = c / g.dims[0];
            const int cy = t % g.dims[1];
            const int cz = t / g.dims[1];
            const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dims[0] - 1);
            const int y0 = D > 1 ? max(cy - 1, 0) : 0, y1 = D > 1 ? min(cy + 1, g.dims[1] - 1) : 0;
            const int z0 = D > 2 ? max(cz - 1, 0) : 0, z1 = D > 2 ? min(cz + 1, g.dims[2] - 1) : 0;
            int cnt = 0;
            for (int z = z0; z <= z1; ++z)
                for (int y = y0; y <= y1; ++y) {
                    const int cb = (z * g.dims[1] + y) * g.dims[0];
                    const int s = cell_start[cb + x0], e = cell_start[cb + x1 + 1];
#pragma unroll 4
                    for (int q0 = s; q0 < e; q0 += k) {
                        const int q = q0 + sub;
                        const bool inb = q < e;
                        // branch-free: out-of-range lanes re-test entry s (always valid here) and are masked
                        const bool hit = within<D, SQRT>(qx, qy, qz, xs, n, inb ? q : s, rc) && inb && qvalid;
                        if (MODE) {
                            const unsigned m = __ballot_sync(full, hit) & qm;
                            if (hit) col_idx[base + cnt + __popc(m & lt)] = q;
                            cnt += __popc(m);
                        } else {
                            cnt += hit ? 1 : 0;
                        }
                    }
                }
            if (MODE == 0)
                for (int off = gp; off < 32; off <<= 1) cnt += __shfl_xor_sync(full, cnt, off);
            if (MODE != 1 && sub == 0 && qvalid) deg[pq] = cnt;
        }
    }
}

// ---- K2, fp32 filter form (the default) ------------------------------------------------------------------
// Same tiles as above (a warp = up to 32 consecutive cell-sorted points).  A GROUP is a run of lanes whose points lie
// in the same x-row of cells and at most GF_XSPAN cells apart; its candidate list is the union of the 3^D windows of
// its cells (still 3^(D-1) contiguous index runs), so that neighbouring sparse cells are served as one well-filled
// group.  Every lane of the group owns ONE query; the candidates of a run go through shared memory 32 at a time (one
// coalesced float4 per lane) and come back as broadcasts.
//
// Two kernels share the group structure (gf_next_group), so nothing about it is stored:
//   grid_mask_kernel   the distance tests: per (query, 32-candidate chunk) ONE 32-bit hit mask, written as a coalesced
//                      row of `gp` words per chunk (gp = group size rounded up to a power of two), plus deg[]; a tile
//                      reserves its words from a running counter and leaves the offset in tile_off[]
//   grid_expand_kernel after the degree scan: masks -> ascending column indices, whole rows at a time through a
//                      shared-memory row buffer, so that every store instruction writes 32 consecutive words
// The hits therefore reach HBM once (4*nnz bytes, coalesced) plus ~100 bytes of masks per point, instead of 4-byte
// scattered stores into gapped rows that a compaction pass had to read back (round 1: 1.15 GB of traffic, now ~0.5).
//
// The test is an fp32 FILTER with a proven error bound, in the expanded form |c|^2 - 2 q.c <= R - |q|^2 (3 FMA per
// pair instead of 6 operations): coordinates are taken relative to the CENTRE of the group's candidate box (|x| <= 3.5h,
// |y|,|z| <= 1.5h with h the cell edge), so the cancellation error stays below 2^-14 h^2 (derivation in DESIGN.md);
// pairs with |t| <= band (t = computed w - T) are decided by the reference's own f64 expression on the original
// coordinates (`within`), so the neighbour sets are exactly those of the f64 kernel.  A lane tracks the smallest |t|
// it meets in a chunk and, if that lies inside the band, redoes the chunk with the exact test for those pairs.
// Per candidate the loop issues 7 instructions for up to 32 tests: LDS.128, 3 FFMA, FADD, FMNMX, SHF (sign bit -> mask).
#define GF_WARPS 8
#define GF_XSPAN 3            // a group spans at most GF_XSPAN + 1 cells along x

struct FilterParams {
    float r2;      // R (squared radius) rounded to fp32
    float band;    // |t| <= band: the pair is decided by the exact f64 test
    double R;      // the threshold centre in f64 (per-lane T = fl32(R - |q|^2))
};

struct GfGroup {
    int a, gsz, lg;            // first lane, number of lanes, log2 of the mask row width gp
    int nruns, total, nch;     // candidate runs, candidates over all runs, chunks of 32 of the CONCATENATED runs
    int cx0, cx1, cy, cz;      // candidate cells along x; the queries' cell row
    bool use_b;                // the run straddles a period boundary of lattice A: take x from lattice B
    int incl, off;             // per lane (lane < nruns): candidates up to and including run `lane`; index of the
                               // run's first candidate minus its position in the concatenation (index = off + position)
    int ry, rz;                // per lane: cell row of run `lane`
};

// next group of the tile: the run of lanes, starting at the first open one, whose points lie in the same x-row of cells
// and at most GF_XSPAN cells apart (a query also meets candidates outside its own window; the test is exact, that only
// costs time).  Warp-uniform except rs/re/ry/rz.
template <int D>
__device__ __forceinline__ void gf_next_group(unsigned &todo, int my_cz, int my_row, int my_cx, const GridDesc &g,
                                              const int *__restrict__ cell_start, int lane, GfGroup &G) {
    const unsigned full = 0xffffffffu;
    G.a = __ffs(todo) - 1;
    // the cell coordinates of the group's first point come from its lane (gf_tile_header divided once per tile;
    // dividing here again, per group and per pass, was 5 integer divisions of ~20 instructions each)
    const int row_a = __shfl_sync(full, my_row, G.a), cx = __shfl_sync(full, my_cx, G.a);
    const unsigned grp = __ballot_sync(full, my_row == row_a && my_cx - cx <= GF_XSPAN) & todo;
    todo &= ~grp;
    G.gsz = __popc(grp);
    G.lg = G.gsz <= 1 ? 0 : 32 - __clz(G.gsz - 1);
    const int cx_max = __shfl_sync(full, my_cx, G.a + G.gsz - 1);
    G.cz = __shfl_sync(full, my_cz, G.a);
    G.cy = row_a - G.cz * g.dims[1];
    G.cx0 = max(cx - 1, 0);
    G.cx1 = min(cx_max + 1, g.dims[0] - 1);
    const int y0 = D > 1 ? max(G.cy - 1, 0) : 0, y1 = D > 1 ? min(G.cy + 1, g.dims[1] - 1) : 0;
    const int z0 = D > 2 ? max(G.cz - 1, 0) : 0, z1 = D > 2 ? min(G.cz + 1, g.dims[2] - 1) : 0;
    G.use_b = (G.cx0 >> 4) != (G.cx1 >> 4);
    const int ny = y1 - y0 + 1;
    G.nruns = ny * (z1 - z0 + 1);
    int rs = 0, len = 0;
    G.ry = G.cy;
    G.rz = G.cz;
    if (lane < G.nruns) {
        const int qz = ny == 1 ? lane : (ny == 2 ? lane >> 1 : (lane * 11) >> 5);   // lane / ny for lane < 9, ny <= 3
        G.ry = y0 + lane - qz * ny;
        G.rz = z0 + qz;
        const int cb = (G.rz * g.dims[1] + G.ry) * g.dims[0];
        rs = cell_start[cb + G.cx0];
        len = cell_start[cb + G.cx1 + 1] - rs;
    }
    // the runs are walked as ONE sequence, 32 candidates per chunk whatever run they come from: a sparse group's nine
    // short runs make one chunk instead of nine
    G.incl = len;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const int v = __shfl_up_sync(full, G.incl, o);
        if (lane >= o) G.incl += v;
    }
    G.off = rs - (G.incl - len);
    G.total = __shfl_sync(full, G.incl, G.nruns - 1);
    G.nch = (G.total + 31) >> 5;
}

// Position v of the concatenated candidate list -> the run it lies in.  `run` only moves forward from chunk to chunk;
// all lanes take part (the shuffles inside are warp-wide), lanes with valid == false do not move.
__device__ __forceinline__ void gf_locate(const GfGroup &G, int v, bool valid, int &run) {
    for (;;) {
        const int nxt = __shfl_sync(0xffffffffu, G.incl, run);
        const bool adv = valid && v >= nxt && run < G.nruns - 1;
        if (!__any_sync(0xffffffffu, adv)) break;
        run += adv ? 1 : 0;
    }
}

__device__ __forceinline__ void gf_tile_header(const int *__restrict__ cell_sorted, const GridDesc &g, ll p, bool pvalid,
                                               int &my_cz, int &my_row, int &my_cx) {
    const int c = pvalid ? cell_sorted[p] : -1;
    my_row = pvalid ? c / g.dims[0] : -1;
    my_cx = pvalid ? c - my_row * g.dims[0] : 0;
    my_cz = pvalid ? my_row / g.dims[1] : 0;   // row = cz * dims[1] + cy
}

template <int D, bool SQRT, int MINB>
__global__ void __launch_bounds__(GF_WARPS * 32, MINB)
grid_mask_kernel(const double *__restrict__ xs, const float4 *__restrict__ xf, ll n, const int *__restrict__ cell_sorted,
                 const int *__restrict__ cell_start, GridDesc g, double rc, FilterParams fp, int *__restrict__ deg,
                 ll *__restrict__ tile_off, unsigned *__restrict__ masks, ll p_begin, ll p_end, ll n_tiles,
                 unsigned long long *stats, int qpw, ll cap_words, unsigned long long *bump) {
    __shared__ float4 s_q[GF_WARPS][32];
    __shared__ int s_i[GF_WARPS][32];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float4 *qf = s_q[wid];
    int *qi = s_i[wid];  // candidate indices of the staged chunk (the exact recheck needs them)
    const ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    unsigned n_recheck = 0;
    const float hh = 0.5f * (float)g.h;
    for (ll t = warp; t < n_tiles; t += nwarps) {
        const ll p = p_begin + t * qpw + lane;
        const bool pvalid = lane < qpw && p < p_end;
        int my_cz, my_row, my_cx;
        gf_tile_header(cell_sorted, g, p, pvalid, my_cz, my_row, my_cx);
        const float4 my4 = pvalid ? xf[p] : make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned todo = __ballot_sync(full, pvalid);
        // the tile's mask words are reserved from a running counter (their number follows from the group structure
        // alone); the offset is left in tile_off[] for the expand pass.  A tile that does not fit the buffer is
        // skipped -- the counter still ends at the total, and the host repeats the pass with a larger buffer.
        unsigned *mw;
        {
            unsigned td = todo;
            int words = 0;
            while (td) {
                GfGroup G;
                gf_next_group<D>(td, my_cz, my_row, my_cx, g, cell_start, lane, G);
                words += G.nch << G.lg;
            }
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(bump, (unsigned long long)words);
            base = __shfl_sync(full, base, 0);
            if (lane == 0) tile_off[t] = (ll)base;
            if ((ll)base + words > cap_words) continue;
            mw = masks + base;
        }
        while (todo) {
            GfGroup G;
            gf_next_group<D>(todo, my_cz, my_row, my_cx, g, cell_start, lane, G);
            const int jme = lane - G.a;  // my query slot when I belong to the group
            const bool mine = jme >= 0 && jme < G.gsz;
            const int gp = 1 << G.lg;
            // shifts that move cell-local coordinates to the centre of the group's candidate box
            const int lat = G.use_b ? ((G.cx0 + 8) & ~15) - 8 : (G.cx0 & ~15);
            const double xc = __dadd_rn(cell_origin(g.lo[0], g.h, G.cx0), __dmul_rn(0.5 * (double)(G.cx1 - G.cx0 + 1), g.h));
            const float shx = (float)__dsub_rn(cell_origin(g.lo[0], g.h, lat), xc);
            float shy = 0.f, shz = 0.f;
            if (lane < G.nruns) {
                if (D > 1) shy = (float)__dsub_rn(__dsub_rn(cell_origin(g.lo[1], g.h, G.ry), cell_origin(g.lo[1], g.h, G.cy)), 0.5 * g.h);
                if (D > 2) shz = (float)__dsub_rn(__dsub_rn(cell_origin(g.lo[2], g.h, G.rz), cell_origin(g.lo[2], g.h, G.cz)), 0.5 * g.h);
            }
            // my query relative to the same centre; lanes outside the group carry a query that never hits and is
            // never borderline (q = 0, T = -1e30: t = |c|^2 + 1e30)
            const float qx = (G.use_b ? my4.y : my4.x) + shx;
            const float qy = D > 1 ? my4.z - hh : 0.f, qz = D > 2 ? my4.w - hh : 0.f;
            const float q2x = mine ? -2.f * qx : 0.f, q2y = mine ? -2.f * qy : 0.f, q2z = mine ? -2.f * qz : 0.f;
            const double qn = __dadd_rn(__dadd_rn(__dmul_rn((double)qx, (double)qx), __dmul_rn((double)qy, (double)qy)),
                                        __dmul_rn((double)qz, (double)qz));
            const float T = mine ? (float)__dsub_rn(fp.R, qn) : -1e30f;
            int cnt = 0;
            // candidates of chunk 0 (position v = lane of the concatenated runs), fetched one chunk ahead from then on
            int run = 0;
            gf_locate(G, lane, lane < G.total, run);
            int idx = __shfl_sync(full, G.off, run) + lane;
            float sy = __shfl_sync(full, shy, run), sz = __shfl_sync(full, shz, run);
            float4 c4 = lane < G.total ? xf[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
            for (int v0 = 0; v0 < G.total; v0 += 32) {
                __syncwarp();
                {
                    const float cx = (G.use_b ? c4.y : c4.x) + shx, cy = D > 1 ? c4.z + sy : 0.f, cz = D > 2 ? c4.w + sz : 0.f;
                    qf[lane] = make_float4(cx, cy, cz, fmaf(cz, cz, fmaf(cy, cy, cx * cx)));
                    qi[lane] = idx;
                }
                __syncwarp();
                {   // next chunk's candidates, in flight during this one
                    const int vn = v0 + 32 + lane;
                    const bool valid = vn < G.total;
                    gf_locate(G, vn, valid, run);
                    idx = __shfl_sync(full, G.off, run) + vn;
                    sy = __shfl_sync(full, shy, run);
                    sz = __shfl_sync(full, shz, run);
                    if (valid) c4 = xf[idx];
                }
                const int nc = min(32, G.total - v0);
                unsigned m = 0;
                float near = INFINITY;  // smallest |t| this lane meets in the chunk
#pragma unroll 8
                for (int tq = 0; tq < nc; ++tq) {
                    const float4 cv = qf[tq];
                    const float w = fmaf(q2z, cv.z, fmaf(q2y, cv.y, fmaf(q2x, cv.x, cv.w)));
                    const float tt = w - T;
                    near = fminf(near, fabsf(tt));
                    m = __funnelshift_l(__float_as_uint(tt), m, 1);  // sign bit of t: candidate tq ends up at bit nc-1-tq
                }
                m = __brev(m) >> (32 - nc);
                if (near <= fp.band) {
                    // rare: a pair of this lane is borderline -- decide it with the reference's own f64 test
                    for (int tq = 0; tq < nc; ++tq) {
                        const float4 cv = qf[tq];
                        const float w = fmaf(q2z, cv.z, fmaf(q2y, cv.y, fmaf(q2x, cv.x, cv.w)));
                        const float tt = w - T;
                        if (fabsf(tt) <= fp.band) {
                            const bool hit = within<D, SQRT>(xs[p], D > 1 ? xs[n + p] : 0.0, D > 2 ? xs[2 * n + p] : 0.0, xs, n,
                                                             qi[tq], rc);
                            m = (m & ~(1u << tq)) | ((hit ? 1u : 0u) << tq);
                            ++n_recheck;
                        }
                    }
                }
                if (mine) mw[jme] = m;
                mw += gp;
                cnt += __popc(m);
            }
            if (mine) deg[p] = cnt;
        }
    }
    n_recheck = __reduce_add_sync(full, n_recheck);
    if (lane == 0 && n_recheck) atomicAdd(stats, (unsigned long long)n_recheck);
}

// masks -> column indices.  One warp per tile, one query row per lane: a lane walks the mask words of its row in chunk
// order (the words of a chunk are contiguous over the lanes of the group: coalesced loads, the next word fetched one
// step ahead) and turns set bits into indices in a shared-memory stage of 64 slots per row.  After every chunk the
// rows that hold 32 entries or more are written out by the whole warp -- full 128-byte store instructions -- and the
// rest at the end of the tile.  Stage layout: row-major with a pad of one word (the 32 lanes reading one row hit 32
// banks).  Row positions are kept as 32-bit offsets from the tile's first row.
#define GX_WARPS 4
#define GX_STAGE 64
template <int D>
__global__ void __launch_bounds__(GX_WARPS * 32)
grid_expand_kernel(const int *__restrict__ cell_sorted, const int *__restrict__ cell_start, GridDesc g, ll p_begin, ll p_end,
                   int qpw, ll n_tiles, const ll *__restrict__ tile_off, const unsigned *__restrict__ masks,
                   const ll *__restrict__ row_ptr, int *__restrict__ col_idx) {
    __shared__ int s_stage[GX_WARPS][32 * (GX_STAGE + 1)];
    __shared__ int s_cand[GX_WARPS][32];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int *stage = s_stage[wid];
    int *cand = s_cand[wid];
    int *mine_stage = stage + lane * (GX_STAGE + 1);
    const ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll t = warp; t < n_tiles; t += nwarps) {
        const ll p0 = p_begin + t * qpw;
        const ll p = p0 + lane;
        const bool pvalid = lane < qpw && p < p_end;
        int my_cz, my_row, my_cx;
        gf_tile_header(cell_sorted, g, p, pvalid, my_cz, my_row, my_cx);
        const ll tile_base = row_ptr[p0];
        int *out = col_idx + tile_base;
        int rowpos = pvalid ? (int)(row_ptr[p] - tile_base) : 0;  // where my row's next entry goes, from the tile's base
        int cnt = 0;                                               // staged entries of my row
        // writes the staged entries of the rows in `rows` (all of them: a row is emptied by a flush)
        auto flush = [&](unsigned rows) {
            __syncwarp();
            const int mycnt = cnt;
            if ((rows >> lane) & 1u) { rowpos += cnt; cnt = 0; }
            while (rows) {
                const int j = __ffs(rows) - 1;
                rows &= rows - 1;
                const int c = __shfl_sync(full, mycnt, j);
                const int base = __shfl_sync(full, rowpos, j) - c;
                const int *src = stage + j * (GX_STAGE + 1);
                out[base + lane] = src[lane];  // c >= 32 whenever a flush is triggered inside the tile
                if (lane + 32 < c) out[base + lane + 32] = src[lane + 32];
            }
            __syncwarp();
        };
        const unsigned *mw = masks + tile_off[t];
        unsigned todo = __ballot_sync(full, pvalid);
        while (todo) {
            GfGroup G;
            gf_next_group<D>(todo, my_cz, my_row, my_cx, g, cell_start, lane, G);
            const int jme = lane - G.a;
            const bool mine = jme >= 0 && jme < G.gsz;
            const int gp = 1 << G.lg;
            int left = G.nch;  // chunks of the group still to come; the next chunk's word is fetched one step ahead
            unsigned m_next = (mine && left > 0) ? mw[jme] : 0u;
            int run = 0;
            for (int v0 = 0; v0 < G.total; v0 += 32) {
                // candidate index of every position of the chunk (lane = position), through shared memory
                const int v = v0 + lane;
                gf_locate(G, v, v < G.total, run);
                __syncwarp();
                cand[lane] = __shfl_sync(full, G.off, run) + v;
                __syncwarp();
                unsigned m = m_next;
                mw += gp;
                --left;
                m_next = (mine && left > 0) ? mw[jme] : 0u;
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    mine_stage[cnt++] = cand[b];
                }
                const unsigned rdy = __ballot_sync(full, cnt >= 32);
                if (rdy) flush(rdy);
            }
        }
        // the remainders (fewer than 32 entries per row)
        __syncwarp();
        unsigned rows = __ballot_sync(full, cnt > 0);
        while (rows) {
            const int j = __ffs(rows) - 1;
            rows &= rows - 1;
            const int c = __shfl_sync(full, cnt, j);
            const int base = __shfl_sync(full, rowpos, j);
            if (lane < c) out[base + lane] = stage[j * (GX_STAGE + 1) + lane];
        }
        __syncwarp();
    }
}

// cap[p] = size of the candidate list of point p (all points of the 3^D cells around its cell)
template <int D>
__global__ void grid_cap_kernel(ll n, const int *__restrict__ cell_sorted, const int *__restrict__ cell_start,
                                GridDesc g, int *__restrict__ cap) {
    ll p = blockIdx.x * (ll)blockDim.x + threadIdx.x;
    if (p > n) return;
    if (p == n) { cap[n] = 0; return; }
    const int c = cell_sorted[p];
    const int cx = c % g.dims[0];
    const int t = c / g.dims[0];
    const int cy = t % g.dims[1];
    const int cz = t / g.dims[1];
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dims[0] - 1);
    const int y0 = D > 1 ? max(cy - 1, 0) : 0, y1 = D > 1 ? min(cy + 1, g.dims[1] - 1) : 0;
    const int z0 = D > 2 ? max(cz - 1, 0) : 0, z1 = D > 2 ? min(cz + 1, g.dims[2] - 1) : 0;
    int tot = 0;
    for (int z = z0; z <= z1; ++z)
        for (int y = y0; y <= y1; ++y) {
            const int cb = (z * g.dims[1] + y) * g.dims[0];
            tot += cell_start[cb + x1 + 1] - cell_start[cb + x0];
        }
    cap[p] = tot;
}

// one warp per row: gapped[goff[p] .. +deg[p]) -> col_idx[row_ptr[p] .. +deg[p])
__global__ void __launch_bounds__(256)
compact_rows_kernel(const int *__restrict__ gapped, const ll *__restrict__ goff, const ll *__restrict__ row_ptr,
                    ll p_begin, ll p_end, int *__restrict__ col_idx) {
    const int lane = threadIdx.x & 31;
    ll warp = (blockIdx.x * (ll)blockDim.x + threadIdx.x) >> 5;
    ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll p = p_begin + warp; p < p_end; p += nwarps) {
        const ll src = goff[p], dst = row_ptr[p];
        const int m = (int)(row_ptr[p + 1] - dst);
        for (int k = lane; k < m; k += 32) col_idx[dst + k] = gapped[src + k];
    }
}

// ---- exclusive scan deg -> row_ptr (shared with the other producers) ------------------------
int cnn_scan_deg_to_row_ptr(cnnb200_ctx *ctx, const int32_t *deg, ll n, ll *row_ptr, ll *nnz_out,
                            ll *max_deg_out) {
    // deg has n+1 entries, the last one 0, so that row_ptr[n] = nnz comes out of the scan
    // int32 in, int64 out: the accumulator type follows the int64 initial value
    Span sp(ctx, T_SCAN);
    size_t need1 = 0, need2 = 0;
    CNN_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, need1, deg, row_ptr, ::cuda::std::plus<>{}, (ll)0, (int)(n + 1), ctx->stream));
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp2, 64));
    int *d_max = ctx->b_tmp2.as<int>() + 8;  // slots 0..7 belong to the callers' flags
    CNN_CUDA(cub::DeviceReduce::Max(nullptr, need2, deg, d_max, (int)(n + 1), ctx->stream));
    CNN_TRY(cnn_ensure(ctx, ctx->b_cub, need1 > need2 ? need1 : need2));
    size_t cap = ctx->b_cub.cap;
    CNN_CUDA(cub::DeviceScan::ExclusiveScan(ctx->b_cub.p, cap, deg, row_ptr, ::cuda::std::plus<>{}, (ll)0, (int)(n + 1), ctx->stream));
    cap = ctx->b_cub.cap;
    CNN_CUDA(cub::DeviceReduce::Max(ctx->b_cub.p, cap, deg, d_max, (int)(n + 1), ctx->stream));
    void *h;
    CNN_TRY(cnn_pinned(ctx, 64, &h));
    CNN_CUDA(cudaMemcpyAsync(h, row_ptr + n, sizeof(ll), cudaMemcpyDeviceToHost, ctx->stream));
    CNN_CUDA(cudaMemcpyAsync((char *)h + 8, d_max, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->grid_recheck_pending)  // pairs the fp32 filter handed to the exact test (counter next to d_max)
        CNN_CUDA(cudaMemcpyAsync((char *)h + 16, ctx->b_tmp2.as<char>() + 40, sizeof(ll), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->mask_words_pending)
        CNN_CUDA(cudaMemcpyAsync((char *)h + 24, ctx->mask_total, sizeof(ll), cudaMemcpyDeviceToHost, ctx->stream));
    CNN_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->mask_words_pending) {
        ctx->mask_words_needed = *(ll *)((char *)h + 24);
        ctx->stats[ST_DIST] = ctx->mask_words_needed * 32;
    }
    if (ctx->grid_recheck_pending) {
        ctx->stats[ST_RECHECK] = *(ll *)((char *)h + 16);
        ctx->grid_recheck_pending = false;
    }
    *nnz_out = *(ll *)h;
    *max_deg_out = *(int *)((char *)h + 8);
    return CNNB200_OK;
}

template <int D, bool SQRT>
static int launch_rows(cnnb200_ctx *ctx, ll pb, ll pe, int mode, int *deg, const ll *offsets, int *out) {
    const GridDesc &g = ctx->grid;
    const ll n = ctx->grid_n;
    const double rc = ctx->grid_rc;
    const double *xs = ctx->b_xs.as<double>();
    const int *cs = ctx->b_cell_id_sorted.as<int>();
    const int *cstart = ctx->b_cell_start.as<int>();
    int blocks = (int)((pe - pb + 255) / 256);  // one warp per tile of 32 points, 8 warps per CTA
    int cap = ctx->sm_count * 64;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (mode == 0)
        grid_rows_kernel<D, SQRT, 0><<<blocks, 256, 0, ctx->stream>>>(xs, n, cs, cstart, g, rc, deg, nullptr, nullptr, pb, pe);
    else if (mode == 1)
        grid_rows_kernel<D, SQRT, 1><<<blocks, 256, 0, ctx->stream>>>(xs, n, cs, cstart, g, rc, nullptr, offsets, out, pb, pe);
    else
        grid_rows_kernel<D, SQRT, 2><<<blocks, 256, 0, ctx->stream>>>(xs, n, cs, cstart, g, rc, deg, offsets, out, pb, pe);
    CNN_LAUNCHED(ctx, 1);
    CNN_KERNEL_CHECK();
    return CNNB200_OK;
}

// thresholds of the fp32 filter (see grid_mask_kernel); false when fp32 cannot represent the problem
static bool filter_params(const cnnb200_ctx *ctx, FilterParams *fp) {
    const double h = ctx->grid.h;
    if (!(h > 1e-12 && h < 1e12)) return false;
    const double R = ctx->grid_metric == CNNB200_METRIC_EUCLIDEAN ? ctx->grid_rc * ctx->grid_rc : ctx->grid_rc;
    const double r = sqrt(R);
    // coordinate differences carry delta <= h 2^-18 (cell-local fp32 copies and their shifts): 2 sqrt(3) delta sqrt(d2) +
    // 3 delta^2 near the cutoff; the expanded form adds the roundings of |c|^2, of the three FMAs, of T = R - |q|^2
    // and of t = w - T: below 2^-14 h^2 for |q|, |c| <= 5 h (exactness of the fp32 filter)
    const double eps = ldexp(h * r, -16) + ldexp(h * h, -34) + ldexp(R, -21) + ldexp(h * h, -14);
    fp->R = R;
    fp->r2 = (float)R;
    fp->band = nextafterf((float)(eps * (1.0 + ldexp(1.0, -20))), INFINITY);
    return isfinite(fp->band) && isfinite(fp->r2);
}

static int mask_tiles(const cnnb200_ctx *ctx, ll pb, ll pe, int *qpw_out, ll *n_tiles) {
    // fewer queries per warp tile when there are few points with long candidate lists (C1: 10^4 points, 10^3
    // neighbours each), so that the tiles still cover all SMs
    int qpw = 32;
    while (qpw > 8 && (pe - pb) / qpw < (ll)ctx->sm_count * 16) qpw >>= 1;
    if (const char *e = getenv("CNNB200_QPW")) qpw = atoi(e) >= 1 && atoi(e) <= 32 ? atoi(e) : qpw;
    *qpw_out = qpw;
    *n_tiles = (pe - pb + qpw - 1) / qpw;
    return CNNB200_OK;
}

// mask pass on the row block [pb, pe): deg[] and the hit masks (b_masks at b_tile_off[]).  speculative: launched with
// the buffer at hand, the number of words it asked for travels to the host with the degree scan's synchronisation
// (cnn_scan_deg_to_row_ptr) and the caller repeats the pass non-speculatively when it did not fit; otherwise the
// pass is repeated here until it fits.
template <int D, bool SQRT>
static int launch_mask(cnnb200_ctx *ctx, ll pb, ll pe, FilterParams fp, int *deg, bool speculative) {
    cudaStream_t st = ctx->stream;
    int qpw;
    ll n_tiles;
    mask_tiles(ctx, pb, pe, &qpw, &n_tiles);
    const int *cs = ctx->b_cell_id_sorted.as<int>(), *cst = ctx->b_cell_start.as<int>();
    CNN_TRY(cnn_ensure(ctx, ctx->b_tile_off, sizeof(ll) * (size_t)(n_tiles + 1)));
    if (ctx->b_masks.cap == 0)  // first guess: 32 words per point (C2 needs 21; sparse sets less)
        CNN_TRY(cnn_ensure(ctx, ctx->b_masks, sizeof(unsigned) * (size_t)((pe - pb) * 32 + (1 << 16))));
    ll blocks = (n_tiles + GF_WARPS - 1) / GF_WARPS;
    const ll cap = (ll)ctx->sm_count * 48;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    unsigned long long *recheck = (unsigned long long *)(ctx->b_tmp2.as<char>() + 40);
    unsigned long long *bump = (unsigned long long *)(ctx->b_tmp2.as<char>() + 48);
    static const int minb = getenv("CNNB200_MASK_MINB") ? atoi(getenv("CNNB200_MASK_MINB")) : 3;
    for (int attempt = 0; attempt < 3; ++attempt) {
        CNN_CUDA(cudaMemsetAsync(recheck, 0, 2 * sizeof(unsigned long long), st));  // recheck counter + bump
#define GM_ARGS ctx->b_xs.as<double>(), ctx->b_xf.as<float4>(), ctx->grid_n, cs, cst, ctx->grid, ctx->grid_rc, fp, deg,     \
                ctx->b_tile_off.as<ll>(), ctx->b_masks.as<unsigned>(), pb, pe, n_tiles, recheck, qpw,                 \
                (ll)(ctx->b_masks.cap / sizeof(unsigned)), bump
        if (minb == 4) grid_mask_kernel<D, SQRT, 4><<<(int)blocks, GF_WARPS * 32, 0, st>>>(GM_ARGS);
        else if (minb == 2) grid_mask_kernel<D, SQRT, 2><<<(int)blocks, GF_WARPS * 32, 0, st>>>(GM_ARGS);
        else grid_mask_kernel<D, SQRT, 3><<<(int)blocks, GF_WARPS * 32, 0, st>>>(GM_ARGS);
#undef GM_ARGS
        CNN_LAUNCHED(ctx, 1);
        CNN_KERNEL_CHECK();
        ctx->grid_recheck_pending = true;
        ctx->mask_total = (const ll *)bump;
        if (speculative) {
            ctx->mask_words_pending = true;
            return CNNB200_OK;
        }
        ll *h;
        CNN_TRY(cnn_pinned(ctx, 64, (void **)&h));
        CNN_CUDA(cudaMemcpyAsync(h, bump, sizeof(ll), cudaMemcpyDeviceToHost, st));
        CNN_CUDA(cudaStreamSynchronize(st));
        const ll words = h[0];
        ctx->stats[ST_DIST] = words * 32;
        ctx->mask_words_pending = false;
        if ((size_t)words * sizeof(unsigned) <= ctx->b_masks.cap) return CNNB200_OK;
        CNN_TRY(cnn_ensure(ctx, ctx->b_masks, sizeof(unsigned) * (size_t)words));
    }
    CNN_FAIL(CNNB200_E_STATE, "cell-list search: the hit masks did not fit after two regrowths");
}

static int launch_mask_any(cnnb200_ctx *ctx, ll pb, ll pe, FilterParams fp, int *deg, bool speculative) {
    bool sq = ctx->grid_metric == CNNB200_METRIC_EUCLIDEAN;
    switch (ctx->grid.d * 2 + (sq ? 1 : 0)) {
        case 2: return launch_mask<1, false>(ctx, pb, pe, fp, deg, speculative);
        case 3: return launch_mask<1, true>(ctx, pb, pe, fp, deg, speculative);
        case 4: return launch_mask<2, false>(ctx, pb, pe, fp, deg, speculative);
        case 5: return launch_mask<2, true>(ctx, pb, pe, fp, deg, speculative);
        case 6: return launch_mask<3, false>(ctx, pb, pe, fp, deg, speculative);
        default: return launch_mask<3, true>(ctx, pb, pe, fp, deg, speculative);
    }
}

static int launch_expand(cnnb200_ctx *ctx, ll pb, ll pe, const ll *row_ptr, int *col_idx) {
    int qpw;
    ll n_tiles;
    mask_tiles(ctx, pb, pe, &qpw, &n_tiles);
    ll blocks = (n_tiles + GX_WARPS - 1) / GX_WARPS;
    const ll cap = (ll)ctx->sm_count * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
#define GX_ARGS ctx->b_cell_id_sorted.as<int>(), ctx->b_cell_start.as<int>(), ctx->grid, pb, pe, qpw, n_tiles,           \
                ctx->b_tile_off.as<ll>(), ctx->b_masks.as<unsigned>(), row_ptr, col_idx
    if (ctx->grid.d == 1) grid_expand_kernel<1><<<(int)blocks, GX_WARPS * 32, 0, ctx->stream>>>(GX_ARGS);
    else if (ctx->grid.d == 2) grid_expand_kernel<2><<<(int)blocks, GX_WARPS * 32, 0, ctx->stream>>>(GX_ARGS);
    else grid_expand_kernel<3><<<(int)blocks, GX_WARPS * 32, 0, ctx->stream>>>(GX_ARGS);
#undef GX_ARGS
    CNN_LAUNCHED(ctx, 1);
    CNN_KERNEL_CHECK();
    return CNNB200_OK;
}

static int launch_rows_any(cnnb200_ctx *ctx, ll pb, ll pe, int mode, int *deg, const ll *offsets, int *out) {
    bool sq = ctx->grid_metric == CNNB200_METRIC_EUCLIDEAN;
    switch (ctx->grid.d * 2 + (sq ? 1 : 0)) {
        case 2: return launch_rows<1, false>(ctx, pb, pe, mode, deg, offsets, out);
        case 3: return launch_rows<1, true>(ctx, pb, pe, mode, deg, offsets, out);
        case 4: return launch_rows<2, false>(ctx, pb, pe, mode, deg, offsets, out);
        case 5: return launch_rows<2, true>(ctx, pb, pe, mode, deg, offsets, out);
        case 6: return launch_rows<3, false>(ctx, pb, pe, mode, deg, offsets, out);
        default: return launch_rows<3, true>(ctx, pb, pe, mode, deg, offsets, out);
    }
}

static int scan_int_to_ll(cnnb200_ctx *ctx, const int *in, ll *out, ll count) {
    size_t need = 0;
    CNN_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, need, in, out, ::cuda::std::plus<>{}, (ll)0, (int)count, ctx->stream));
    CNN_TRY(cnn_ensure(ctx, ctx->b_cub, need));
    size_t cap = ctx->b_cub.cap;
    CNN_CUDA(cub::DeviceScan::ExclusiveScan(ctx->b_cub.p, cap, in, out, ::cuda::std::plus<>{}, (ll)0, (int)count, ctx->stream));
    return CNNB200_OK;
}

// candidate-list sizes (the cell ids in input order are dead after the sort: reuse that buffer) and their prefix =
// offsets of the gapped rows of the f64 form / the work measure used to balance row blocks
int cnn_grid_cap(cnnb200_ctx *ctx) {
    if (ctx->have_cap) return CNNB200_OK;
    cudaStream_t st = ctx->stream;
    const ll n = ctx->grid_n;
    const GridDesc &g = ctx->grid;
    int *capv = ctx->b_cell_id.as<int>();
    int nb1 = cnn_blocks(n + 1, 256);
    const int *cs = ctx->b_cell_id_sorted.as<int>(), *cst = ctx->b_cell_start.as<int>();
    if (g.d == 1) grid_cap_kernel<1><<<nb1, 256, 0, st>>>(n, cs, cst, g, capv);
    else if (g.d == 2) grid_cap_kernel<2><<<nb1, 256, 0, st>>>(n, cs, cst, g, capv);
    else grid_cap_kernel<3><<<nb1, 256, 0, st>>>(n, cs, cst, g, capv);
    CNN_LAUNCHED(ctx, 1);
    CNN_KERNEL_CHECK();
    CNN_TRY(scan_int_to_ll(ctx, capv, ctx->b_tmp1.as<ll>(), n + 1));
    ctx->have_cap = true;
    return CNNB200_OK;
}

// ---- K1: grid geometry, sort by cell, candidate-list prefix --------------------------------------
int cnn_grid_build(cnnb200_ctx *ctx, const double *d_x, ll n, ll d, double radius_cutoff, int metric, bool want_cap) {
    cudaStream_t st = ctx->stream;
    ctx->have_grid = false;
    Span sp_build(ctx, T_GRID_BUILD);
    // --- bounding box --------------------------------------------------------------------
    int rb = ctx->sm_count * 2;
    if (rb > (int)((n + 255) / 256)) rb = (int)((n + 255) / 256);
    CNN_TRY(cnn_ensure(ctx, ctx->b_bbox, sizeof(double) * (size_t)(rb + 1) * 6));
    double *partial = ctx->b_bbox.as<double>();
    bbox_partial_kernel<<<rb, 256, 0, st>>>(d_x, n, (int)d, partial + 6);
    bbox_final_kernel<<<1, 192, 0, st>>>(partial + 6, rb, partial);
    CNN_LAUNCHED(ctx, 2);
    CNN_KERNEL_CHECK();
    double *hb;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&hb));
    CNN_CUDA(cudaMemcpyAsync(hb, partial, sizeof(double) * 6, cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < (int)d; ++k)
        if (!(isfinite(hb[k]) && isfinite(hb[3 + k])))
            CNN_FAIL(CNNB200_E_INVALID, "neighbour search: coordinates contain NaN or infinity");

    // --- grid geometry --------------------------------------------------------------------
    GridDesc g;
    g.d = (int)d;
    double r_true = metric == CNNB200_METRIC_EUCLIDEAN ? radius_cutoff : sqrt(radius_cutoff);
    // cell edge a little larger than the radius: rounding in the cell coordinate can then never
    // put two points that are within the radius more than one cell apart
    g.h = r_true * (1.0 + 1e-6);
    double span = 0;
    for (int k = 0; k < 3; ++k) {
        g.lo[k] = k < d ? hb[k] : 0.0;
        g.dims[k] = 1;
        if (k < d && hb[3 + k] - hb[k] > span) span = hb[3 + k] - hb[k];
    }
    if (!(g.h > 0)) g.h = span > 0 ? span : 1.0;  // radius 0: any cell size works
    const double max_cells = fmax(8.0 * (double)n, 1048576.0);
    for (;;) {
        double tot = 1;
        bool ok = true;
        for (int k = 0; k < (int)d; ++k) {
            double dk = floor((hb[3 + k] - hb[k]) / g.h) + 1;
            if (dk > 1048576.0) ok = false;
            tot *= dk;
            g.dims[k] = dk > 1048576.0 ? 1 : (int)dk;
        }
        if (ok && tot <= max_cells) break;
        g.h *= 1.26;  // sparse bounding box: coarser cells are still a valid candidate superset
    }
    ll ncell = (ll)g.dims[0] * g.dims[1] * g.dims[2];

    // --- cell ids, sort by cell, cell starts, gather ------------------------------------------
    CNN_TRY(cnn_ensure(ctx, ctx->b_cell_id, sizeof(int) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_cell_id_sorted, sizeof(int) * (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_idx, sizeof(int) * (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_perm, sizeof(int) * (size_t)n));
    CNN_TRY(cnn_ensure(ctx, ctx->b_cell_start, sizeof(int) * (size_t)(ncell + 2)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_cell_count, sizeof(int) * (size_t)(ncell + 2)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_xs, sizeof(double) * (size_t)n * (size_t)d));
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp1, sizeof(ll) * (size_t)(n + 1)));
    int *cell_count = ctx->b_cell_count.as<int>();
    CNN_CUDA(cudaMemsetAsync(cell_count, 0, sizeof(int) * (size_t)(ncell + 1), st));
    // a row-sharded run (want_cap) needs an order every rank reproduces: stable radix sort by cell; a single GPU takes
    // the counting sort (CNNB200_ORDERED_SORT=1 forces the radix sort)
    static const bool force_ordered = getenv("CNNB200_ORDERED_SORT") && atoi(getenv("CNNB200_ORDERED_SORT")) != 0;
    const bool ordered = want_cap || force_ordered;
    if (ordered)
        cell_id_kernel<true><<<cnn_blocks(n, 256), 256, 0, st>>>(d_x, n, g, ctx->b_cell_id.as<int>(), ctx->b_idx.as<int>(), cell_count);
    else
        cell_id_kernel<false><<<cnn_blocks(n, 256), 256, 0, st>>>(d_x, n, g, ctx->b_cell_id.as<int>(), ctx->b_idx.as<int>(), cell_count);
    CNN_LAUNCHED(ctx, 1);
    CNN_KERNEL_CHECK();
    CNN_TRY(cnn_ensure(ctx, ctx->b_xf, sizeof(float4) * (size_t)n));
    size_t need_sort = 0, need_scan = 0;
    int bits = 1;
    while ((1ll << bits) < ncell) ++bits;
    if (ordered)
        CNN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need_sort, ctx->b_cell_id.as<int>(),
                                                 ctx->b_cell_id_sorted.as<int>(), ctx->b_idx.as<int>(),
                                                 ctx->b_perm.as<int>(), (int)n, 0, bits, st));
    CNN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need_scan, cell_count, ctx->b_cell_start.as<int>(),
                                           (int)(ncell + 1), st));
    CNN_TRY(cnn_ensure(ctx, ctx->b_cub, need_sort > need_scan ? need_sort : need_scan));
    size_t cap = ctx->b_cub.cap;
    if (ordered)
        CNN_CUDA(cub::DeviceRadixSort::SortPairs(ctx->b_cub.p, cap, ctx->b_cell_id.as<int>(),
                                                 ctx->b_cell_id_sorted.as<int>(), ctx->b_idx.as<int>(),
                                                 ctx->b_perm.as<int>(), (int)n, 0, bits, st));
    cap = ctx->b_cub.cap;
    CNN_CUDA(cub::DeviceScan::ExclusiveSum(ctx->b_cub.p, cap, cell_count, ctx->b_cell_start.as<int>(),
                                           (int)(ncell + 1), st));
    if (ordered)
        gather_kernel<<<cnn_blocks(n, 256), 256, 0, st>>>(d_x, ctx->b_perm.as<int>(), ctx->b_cell_id_sorted.as<int>(), n, g,
                                                            ctx->b_xs.as<double>(), ctx->b_xf.as<float4>());
    else
        scatter_kernel<<<cnn_blocks(n, 256), 256, 0, st>>>(d_x, ctx->b_cell_id.as<int>(), ctx->b_idx.as<int>(),
                                                             ctx->b_cell_start.as<int>(), n, g, ctx->b_perm.as<int>(),
                                                             ctx->b_cell_id_sorted.as<int>(), ctx->b_xs.as<double>(),
                                                             ctx->b_xf.as<float4>());
    CNN_LAUNCHED(ctx, 1);
    CNN_KERNEL_CHECK();
    ctx->grid = g;
    ctx->grid_n = n;
    ctx->grid_rc = radius_cutoff;
    ctx->grid_metric = metric;
    ctx->have_cap = false;
    FilterParams fp;
    ctx->grid_filter = ctx->use_filter && filter_params(ctx, &fp);
    if (want_cap || !ctx->grid_filter) CNN_TRY(cnn_grid_cap(ctx));
    ctx->have_grid = true;
    ctx->perm = ctx->b_perm.as<int>();
    return CNNB200_OK;
}

// ---- K2 on a row block [pb, pe) of the cell-sorted points: degrees + what the fill needs (hit masks, or the
// gapped hits of the f64 form) ----------------------------------------------------------------------------------
int cnn_grid_rows(cnnb200_ctx *ctx, ll pb, ll pe, int *deg, bool speculative) {
    cudaStream_t st = ctx->stream;
    ctx->blk_begin = pb;
    ctx->blk_end = pe;
    CNN_TRY(cnn_ensure(ctx, ctx->b_tmp2, 64));
    FilterParams fp;
    if (ctx->grid_filter && filter_params(ctx, &fp)) {
        Span sp(ctx, T_SEARCH_COUNT);
        ctx->blk_masks = true;
        return launch_mask_any(ctx, pb, pe, fp, deg, speculative);
    }
    ctx->blk_masks = false;
    CNN_TRY(cnn_grid_cap(ctx));
    ll *h;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&h));
    const ll *goff = ctx->b_tmp1.as<ll>();
    CNN_CUDA(cudaMemcpyAsync(h, goff + pb, sizeof(ll), cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaMemcpyAsync(h + 1, goff + pe, sizeof(ll), cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaStreamSynchronize(st));
    const ll g0 = h[0], gapped = h[1] - h[0];
    ctx->blk_goff0 = g0;
    ctx->blk_single = ctx->single_pass_limit > 0 && gapped * 4 <= ctx->single_pass_limit;
    ctx->stats[ST_DIST] = gapped;
    Span sp(ctx, T_SEARCH_COUNT);
    if (ctx->blk_single) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_tmp0, sizeof(int) * (size_t)gapped));
        // row p writes at gapped[goff[p] - g0 ...]: shift the base pointer instead of the offsets
        CNN_TRY(launch_rows_any(ctx, pb, pe, 2, deg, goff, ctx->b_tmp0.as<int>() - g0));
    } else {
        CNN_TRY(launch_rows_any(ctx, pb, pe, 0, deg, nullptr, nullptr));
    }
    return CNNB200_OK;
}

// ---- K2 fill of the same row block into a CSR whose row_ptr covers all rows ----------------------------
int cnn_grid_fill(cnnb200_ctx *ctx, ll pb, ll pe, const ll *row_ptr, int *col_idx) {
    if (pb != ctx->blk_begin || pe != ctx->blk_end)
        CNN_FAIL(CNNB200_E_STATE, "grid fill: row block differs from the one the degrees were computed for");
    Span sp(ctx, T_SEARCH_FILL);
    if (ctx->blk_masks) return launch_expand(ctx, pb, pe, row_ptr, col_idx);
    if (ctx->blk_single) {
        int blocks = (int)((pe - pb + 7) / 8);
        if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
        if (blocks < 1) blocks = 1;
        compact_rows_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->b_tmp0.as<int>() - ctx->blk_goff0, ctx->b_tmp1.as<ll>(),
                                                            row_ptr, pb, pe, col_idx);
        CNN_LAUNCHED(ctx, 1);
        CNN_KERNEL_CHECK();
    } else {
        CNN_TRY(launch_rows_any(ctx, pb, pe, 1, nullptr, row_ptr, col_idx));
    }
    return CNNB200_OK;
}

int cnn_grid_search(cnnb200_ctx *ctx, const double *d_x, ll n, ll d, double radius_cutoff, int metric) {
    cudaStream_t st = ctx->stream;
    CNN_TRY(cnn_grid_build(ctx, d_x, n, d, radius_cutoff, metric, false));
    CNN_TRY(cnn_ensure(ctx, ctx->b_deg, sizeof(int) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_row_ptr, sizeof(ll) * (size_t)(n + 1)));
    CNN_CUDA(cudaMemsetAsync(ctx->b_deg.as<int>() + n, 0, sizeof(int), st));
    CNN_TRY(cnn_grid_rows(ctx, 0, n, ctx->b_deg.as<int>(), true));
    ll nnz = 0, max_deg = 0;
    CNN_TRY(cnn_scan_deg_to_row_ptr(ctx, ctx->b_deg.as<int>(), n, ctx->b_row_ptr.as<ll>(), &nnz, &max_deg));
    if (ctx->mask_words_pending) {
        ctx->mask_words_pending = false;
        if ((size_t)ctx->mask_words_needed * sizeof(unsigned) > ctx->b_masks.cap) {  // the masks did not fit: grow, repeat
            CNN_TRY(cnn_grid_rows(ctx, 0, n, ctx->b_deg.as<int>(), false));
            CNN_TRY(cnn_scan_deg_to_row_ptr(ctx, ctx->b_deg.as<int>(), n, ctx->b_row_ptr.as<ll>(), &nnz, &max_deg));
        }
    }
    CNN_TRY(cnn_ensure(ctx, ctx->b_col_idx, sizeof(int) * (size_t)nnz));
    CNN_TRY(cnn_grid_fill(ctx, 0, n, ctx->b_row_ptr.as<ll>(), ctx->b_col_idx.as<int>()));
    ctx->n = n;
    ctx->nnz = nnz;
    ctx->max_deg = max_deg;
    ctx->row_ptr = ctx->b_row_ptr.as<ll>();
    ctx->col_idx = ctx->b_col_idx.as<int>();
    ctx->perm = ctx->b_perm.as<int>();
    return CNNB200_OK;
}

// ---- row-sharded construction (multi-GPU): C-ABI -----------------------------------------------------
// The cost of a row block is modelled as (candidate pairs) + per_row * (rows): the search and the component pass both
// have a part that follows the number of points, not of candidates (C2 on two GPUs, balanced by candidates alone: the
// rank with 58 % of the points took 0.14 ms longer: 1.55 ms against 1.41 ms).  per_row = 250 (same-box A/B of bench.py
// on C2): 8 GPUs 1.12 -> 1.01 ms per fit, 2 GPUs unchanged (1.60 ms; 500 moves too much onto rank 0, which also
// ranks the components and writes the labels: 1.64 ms).
__global__ void balance_kernel(const ll *__restrict__ goff, ll n, int parts, ll per_row, ll *bounds) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > parts) return;
    if (k == 0) { bounds[0] = 0; return; }
    if (k == parts) { bounds[parts] = n; return; }
    // first row whose cost prefix reaches k/parts of the total, rounded to a multiple of 32 rows
    const ll total = goff[n] + per_row * n;
    const ll target = (ll)((double)total * k / parts);
    ll lo = 0, hi = n;
    while (lo < hi) {
        ll mid = (lo + hi) >> 1;
        if (goff[mid] + per_row * mid < target) lo = mid + 1; else hi = mid;
    }
    lo = (lo + 16) / 32 * 32;
    bounds[k] = lo > n ? n : lo;
}

extern "C" int cnnb200_grid_begin(cnnb200_ctx *ctx, const double *x, int loc, int64_t n, int64_t d,
                                  double radius_cutoff, int metric) {
    if (!ctx || !x) CNN_FAIL(CNNB200_E_INVALID, "grid_begin: NULL argument");
    if (n < 1 || d < 1 || d > 3) CNN_FAIL(CNNB200_E_INVALID, "grid_begin: need n >= 1 and 1 <= d <= 3");
    if (n >= (1ll << 31) - 1) CNN_FAIL(CNNB200_E_RANGE, "grid_begin: n must be < 2^31-1");
    if (metric != CNNB200_METRIC_EUCLIDEAN_REDUCED && metric != CNNB200_METRIC_EUCLIDEAN)
        CNN_FAIL(CNNB200_E_INVALID, "grid_begin: unknown metric %d", metric);
    if (!(radius_cutoff >= 0.0)) CNN_FAIL(CNNB200_E_INVALID, "grid_begin: radius_cutoff must be >= 0");
    CNN_CUDA(cudaSetDevice(ctx->device));
    ctx->have_nb = false;
    ctx->have_comp = false;
    ctx->have_points = false;
    ctx->have_other = false;
    const double *d_x = x;
    if (loc == CNNB200_HOST) {
        CNN_TRY(cnn_ensure(ctx, ctx->b_x, sizeof(double) * (size_t)n * (size_t)d));
        Span sp(ctx, T_H2D);
        CNN_CUDA(cudaMemcpyAsync(ctx->b_x.p, x, sizeof(double) * (size_t)n * (size_t)d, cudaMemcpyHostToDevice, ctx->stream));
        d_x = ctx->b_x.as<double>();
    }
    return cnn_grid_build(ctx, d_x, n, d, radius_cutoff, metric, true);
}

extern "C" int cnnb200_grid_balance(cnnb200_ctx *ctx, int parts, int64_t *bounds) {
    if (!ctx || !bounds || parts < 1 || parts > 1024) CNN_FAIL(CNNB200_E_INVALID, "grid_balance: bad argument");
    if (!ctx->have_grid) CNN_FAIL(CNNB200_E_STATE, "grid_balance: call grid_begin first");
    CNN_CUDA(cudaSetDevice(ctx->device));
    CNN_TRY(cnn_ensure(ctx, ctx->b_bbox, sizeof(ll) * (size_t)(parts + 1)));
    balance_kernel<<<cnn_blocks(parts + 1, 128), 128, 0, ctx->stream>>>(ctx->b_tmp1.as<ll>(), ctx->grid_n, parts,
                                                                         ctx->balance_per_row, ctx->b_bbox.as<ll>());
    CNN_KERNEL_CHECK();
    ll *h;
    CNN_TRY(cnn_pinned(ctx, sizeof(ll) * (size_t)(parts + 1), (void **)&h));
    CNN_CUDA(cudaMemcpyAsync(h, ctx->b_bbox.p, sizeof(ll) * (size_t)(parts + 1), cudaMemcpyDeviceToHost, ctx->stream));
    CNN_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k <= parts; ++k) bounds[k] = h[k];
    for (int k = 1; k <= parts; ++k)
        if (bounds[k] < bounds[k - 1]) bounds[k] = bounds[k - 1];
    return CNNB200_OK;
}

// Rows whose neighbourhoods a rank needs in order to settle the candidate edges of its row block
// [row_begin, row_end): the block itself plus every row that can occur in one of its neighbour lists.
// In the cell-sorted order that is one contiguous range -- from the first point of the lowest cell
// adjacent to the block's first cell to the last point of the highest cell adjacent to its last cell.
extern "C" int cnnb200_grid_halo(cnnb200_ctx *ctx, int64_t row_begin, int64_t row_end, int64_t *lo, int64_t *hi) {
    if (!ctx || !lo || !hi) CNN_FAIL(CNNB200_E_INVALID, "grid_halo: NULL argument");
    if (!ctx->have_grid) CNN_FAIL(CNNB200_E_STATE, "grid_halo: call grid_begin first");
    const ll n = ctx->grid_n;
    if (row_begin < 0 || row_end > n || row_begin > row_end) CNN_FAIL(CNNB200_E_INVALID, "grid_halo: bad row range");
    if (row_begin == row_end) { *lo = *hi = row_begin; return CNNB200_OK; }
    CNN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const GridDesc &g = ctx->grid;
    const ll ncell = (ll)g.dims[0] * g.dims[1] * g.dims[2];
    int *h;
    CNN_TRY(cnn_pinned(ctx, 64, (void **)&h));
    const int *cs = ctx->b_cell_id_sorted.as<int>(), *cst = ctx->b_cell_start.as<int>();
    CNN_CUDA(cudaMemcpyAsync(h, cs + row_begin, sizeof(int), cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaMemcpyAsync(h + 1, cs + row_end - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaStreamSynchronize(st));
    // cell ids are x-fastest: the window of cell id c lies within [c - reach, c + reach]
    const ll reach = (g.d > 2 ? (ll)g.dims[0] * g.dims[1] : 0) + (g.d > 1 ? g.dims[0] : 0) + 1;
    ll c_lo = (ll)h[0] - reach, c_hi = (ll)h[1] + reach + 1;
    if (c_lo < 0) c_lo = 0;
    if (c_hi > ncell) c_hi = ncell;
    CNN_CUDA(cudaMemcpyAsync(h + 2, cst + c_lo, sizeof(int), cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaMemcpyAsync(h + 3, cst + c_hi, sizeof(int), cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaStreamSynchronize(st));
    *lo = h[2] < row_begin ? h[2] : row_begin;
    *hi = h[3] > row_end ? h[3] : row_end;
    return CNNB200_OK;
}

extern "C" int cnnb200_grid_degrees(cnnb200_ctx *ctx, int64_t row_begin, int64_t row_end, int32_t *deg) {
    if (!ctx || !deg) CNN_FAIL(CNNB200_E_INVALID, "grid_degrees: NULL argument");
    if (!ctx->have_grid) CNN_FAIL(CNNB200_E_STATE, "grid_degrees: call grid_begin first");
    if (row_begin < 0 || row_end > ctx->grid_n || row_begin > row_end)
        CNN_FAIL(CNNB200_E_INVALID, "grid_degrees: bad row range");
    CNN_CUDA(cudaSetDevice(ctx->device));
    return cnn_grid_rows(ctx, row_begin, row_end, deg, false);
}

extern "C" int cnnb200_grid_fill(cnnb200_ctx *ctx, int64_t row_begin, int64_t row_end, const int64_t *row_ptr,
                                 int32_t *col_idx) {
    if (!ctx || !row_ptr || !col_idx) CNN_FAIL(CNNB200_E_INVALID, "grid_fill: NULL argument");
    if (!ctx->have_grid) CNN_FAIL(CNNB200_E_STATE, "grid_fill: call grid_begin first");
    CNN_CUDA(cudaSetDevice(ctx->device));
    return cnn_grid_fill(ctx, row_begin, row_end, (const ll *)row_ptr, col_idx);
}

extern "C" int cnnb200_neighbours_adopt(cnnb200_ctx *ctx, const int64_t *row_ptr, const int32_t *col_idx, int64_t n,
                                        int64_t nnz, int64_t max_deg, int grid_perm) {
    if (!ctx || !row_ptr || (!col_idx && nnz > 0)) CNN_FAIL(CNNB200_E_INVALID, "neighbours_adopt: NULL argument");
    if (grid_perm && (!ctx->have_grid || ctx->grid_n != n))
        CNN_FAIL(CNNB200_E_STATE, "neighbours_adopt: no grid of that size (call grid_begin first)");
    ctx->n = n;
    ctx->nnz = nnz;
    ctx->max_deg = max_deg;
    ctx->row_ptr = (const ll *)row_ptr;
    ctx->col_idx = col_idx;
    ctx->perm = grid_perm ? ctx->b_perm.as<int>() : nullptr;
    ctx->have_nb = true;
    ++ctx->nb_version;
    ctx->have_comp = false;
    ctx->have_points = false;
    ctx->have_other = false;
    return CNNB200_OK;
}

// ---- the row-sharded fit on coordinates as TWO calls per rank ---------------------------------------------------------
// begin:  cell sort (replicated), row blocks balanced by candidate pairs, this rank's block + halo searched, the
//         candidate edges of the block settled on a private parent array.  Three host synchronisations (bounding box,
//         block/halo bounds, nnz), no library kernels other than CUB, nothing in Python between the stages.
// (the caller all-gathers the parent arrays: cnnb200_components_parents)
// end:    cnnb200_fit_sharded_end (cn.cu): merge, ranking of the components, labels.
__global__ void block_ranges_kernel(const ll *__restrict__ goff, const int *__restrict__ cell_sorted,
                                    const int *__restrict__ cell_start, GridDesc g, ll n, int parts, int rank, ll per_row,
                                    ll *out) {
    // out[0 .. parts] = block bounds; out[parts + 1 + 2 r], out[parts + 2 + 2 r] = block r plus its halo (all ranks: every
    // rank then knows which window of its parent array each of the others has to send)
    __shared__ ll s_bounds[1026];
    for (int k = threadIdx.x; k <= parts; k += blockDim.x) {
        ll b;
        if (k == 0) b = 0;
        else if (k == parts) b = n;
        else {
            const ll total = goff[n] + per_row * n;   // cost model: see balance_kernel
            const ll target = (ll)((double)total * k / parts);
            ll lo = 0, hi = n;
            while (lo < hi) {
                ll mid = (lo + hi) >> 1;
                if (goff[mid] + per_row * mid < target) lo = mid + 1; else hi = mid;
            }
            lo = (lo + 16) / 32 * 32;
            b = lo > n ? n : lo;
        }
        s_bounds[k] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k <= parts; ++k)
            if (s_bounds[k] < s_bounds[k - 1]) s_bounds[k] = s_bounds[k - 1];
        for (int k = 0; k <= parts; ++k) out[k] = s_bounds[k];
    }
    __syncthreads();
    for (int r = threadIdx.x; r < parts; r += blockDim.x) {
        const ll rb = s_bounds[r], re = s_bounds[r + 1];
        ll lo = rb, hi = re;
        if (re > rb) {
            // cell ids are x-fastest: the window of cell id c lies within [c - reach, c + reach]
            const ll ncell = (ll)g.dims[0] * g.dims[1] * g.dims[2];
            const ll reach = (g.d > 2 ? (ll)g.dims[0] * g.dims[1] : 0) + (g.d > 1 ? g.dims[0] : 0) + 1;
            ll c_lo = (ll)cell_sorted[rb] - reach, c_hi = (ll)cell_sorted[re - 1] + reach + 1;
            if (c_lo < 0) c_lo = 0;
            if (c_hi > ncell) c_hi = ncell;
            const ll a = cell_start[c_lo], b = cell_start[c_hi];
            lo = a < rb ? a : rb;
            hi = b > re ? b : re;
        }
        out[parts + 1 + 2 * r] = lo;
        out[parts + 2 + 2 * r] = hi;
    }
}

extern "C" int cnnb200_fit_sharded_begin(cnnb200_ctx *ctx, const double *x, int loc, int64_t n, int64_t d,
                                         double radius_cutoff, int metric, int64_t n_member_cutoff,
                                         int64_t similarity_cutoff, int rank, int world, int64_t *bounds,
                                         int32_t **parents) {
    if (!ctx || !x || !bounds || !parents) CNN_FAIL(CNNB200_E_INVALID, "fit_sharded_begin: NULL argument");
    if (world < 1 || world > 1024 || rank < 0 || rank >= world) CNN_FAIL(CNNB200_E_INVALID, "fit_sharded_begin: bad rank / world");
    CNN_TRY(cnnb200_grid_begin(ctx, x, loc, n, d, radius_cutoff, metric));
    cudaStream_t st = ctx->stream;
    // block bounds and this rank's halo, computed on the device, one small copy back
    CNN_TRY(cnn_ensure(ctx, ctx->b_bbox, sizeof(ll) * (size_t)(3 * world + 1)));
    block_ranges_kernel<<<1, 256, 0, st>>>(ctx->b_tmp1.as<ll>(), ctx->b_cell_id_sorted.as<int>(), ctx->b_cell_start.as<int>(),
                                           ctx->grid, n, world, rank, ctx->balance_per_row, ctx->b_bbox.as<ll>());
    CNN_LAUNCHED(ctx, 1);
    CNN_KERNEL_CHECK();
    ll *h;
    CNN_TRY(cnn_pinned(ctx, sizeof(ll) * (size_t)(3 * world + 1), (void **)&h));
    CNN_CUDA(cudaMemcpyAsync(h, ctx->b_bbox.p, sizeof(ll) * (size_t)(3 * world + 1), cudaMemcpyDeviceToHost, st));
    CNN_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k <= world; ++k) bounds[k] = h[k];
    const ll rb = h[rank], re = h[rank + 1], lo = h[world + 1 + 2 * rank], hi = h[world + 2 + 2 * rank];
    ctx->shard_world = world;
    for (int r = 0; r < 2 * world; ++r) ctx->shard_ranges[r] = h[world + 1 + r];
    // rows [lo, hi): degrees (row-indexed, zero elsewhere), offsets, entries
    CNN_TRY(cnn_ensure(ctx, ctx->b_deg, sizeof(int) * (size_t)(n + 1)));
    CNN_TRY(cnn_ensure(ctx, ctx->b_row_ptr, sizeof(ll) * (size_t)(n + 1)));
    CNN_CUDA(cudaMemsetAsync(ctx->b_deg.p, 0, sizeof(int) * (size_t)(n + 1), st));
    ll nnz = 0, max_deg = 0;
    if (hi > lo) {
        CNN_TRY(cnn_grid_rows(ctx, lo, hi, ctx->b_deg.as<int>(), true));
        CNN_TRY(cnn_scan_deg_to_row_ptr(ctx, ctx->b_deg.as<int>(), n, ctx->b_row_ptr.as<ll>(), &nnz, &max_deg));
        if (ctx->mask_words_pending) {
            ctx->mask_words_pending = false;
            if ((size_t)ctx->mask_words_needed * sizeof(unsigned) > ctx->b_masks.cap) {
                CNN_TRY(cnn_grid_rows(ctx, lo, hi, ctx->b_deg.as<int>(), false));
                CNN_TRY(cnn_scan_deg_to_row_ptr(ctx, ctx->b_deg.as<int>(), n, ctx->b_row_ptr.as<ll>(), &nnz, &max_deg));
            }
        }
        CNN_TRY(cnn_ensure(ctx, ctx->b_col_idx, sizeof(int) * (size_t)nnz));
        CNN_TRY(cnn_grid_fill(ctx, lo, hi, ctx->b_row_ptr.as<ll>(), ctx->b_col_idx.as<int>()));
    } else {
        CNN_CUDA(cudaMemsetAsync(ctx->b_row_ptr.p, 0, sizeof(ll) * (size_t)(n + 1), st));
        CNN_TRY(cnn_ensure(ctx, ctx->b_col_idx, 16));
    }
    ctx->n = n;
    ctx->nnz = nnz;
    ctx->max_deg = max_deg;
    ctx->row_ptr = ctx->b_row_ptr.as<ll>();
    ctx->col_idx = ctx->b_col_idx.as<int>();
    ctx->perm = ctx->b_perm.as<int>();
    ctx->have_nb = true;
    ++ctx->nb_version;
    ctx->stats[ST_PATH] = (ctx->stats[ST_PATH] & ~15ll) | 1;
    CNN_TRY(cnnb200_components_begin(ctx, n_member_cutoff, similarity_cutoff, nullptr, CNNB200_DEVICE));
    if (re > rb) CNN_TRY(cnnb200_components_edges(ctx, rb, re));
    *parents = ctx->b_parent.as<int>();
    return CNNB200_OK;
}
