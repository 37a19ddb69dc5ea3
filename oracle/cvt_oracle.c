/*
 * oracle/cvt_oracle.c — TEST INFRASTRUCTURE ONLY ("port" oracle).
 *
 * A plain-C, single-threaded CPU restatement of the reference's CVT / restricted
 * Voronoi diagram hot path (geogram as vendored in GraphiteThree). It exists to
 * CHECK the CUDA path; nothing under graphitethree_b200/ may call, link or load it.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it.
 *
 * Parity status: PINNED against the reference itself — tests/test_oracle_cpu.py
 * compares every function below with oracle/_ref (the unmodified reference compiled by
 * oracle/Makefile.ref) when that build is present, and tests/golden/ holds vectors
 * generated from the reference by tests/golden/make_golden.py.
 *
 * Each function cites the reference code it restates. G/ = geogram/src/lib/geogram/.
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -frounding-math (no FMA contraction,
 * as geogram/cmake/platforms/Linux-gcc.cmake:43).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORC_MAXDIM 8
#define ORC_MAXPOLY 64
#define ORC_NONE 0xffffffffu

typedef uint32_t u32;
typedef uint64_t u64;

/* ------------------------------------------------------------------------- */
/* Geometry helpers                                                           */
/* ------------------------------------------------------------------------- */

/* G/basic/geometry_nd.h:65-74 — sum over coordinates, in order, of (p2-p1)^2 */
static double distance2(const double* p1, const double* p2, int dim) {
    double r = 0.0;
    for (int i = 0; i < dim; ++i) {
        double d = p2[i] - p1[i];
        r += d * d;
    }
    return r;
}

/* G/basic/geometry_nd.h:143-156 — Heron's formula with max(A2,0) clamp */
static double triangle_area(const double* p1, const double* p2, const double* p3, int dim) {
    double a = sqrt(distance2(p1, p2, dim));
    double b = sqrt(distance2(p2, p3, dim));
    double c = sqrt(distance2(p3, p1, dim));
    double s = 0.5 * (a + b + c);
    double A2 = s * (s - a) * (s - b) * (s - c);
    return sqrt(A2 > 0.0 ? A2 : 0.0);
}

/* ------------------------------------------------------------------------- */
/* Nearest-neighbour search: exact kNN with the reference's result semantics. */
/* The reference uses a balanced kd-tree (G/points/kd_tree.cpp:156-387); any  */
/* exact search returns the same distances, and the same indices except among */
/* exact ties (kd_tree.h:173-195: order among ties = traversal order), which   */
/* are reported through tie flags instead of reproduced.                      */
/* Here: uniform grid over the first 3 coordinates + ring expansion.          */
/* ------------------------------------------------------------------------- */

typedef struct {
    int dim;
    u32 n;
    const double* pts;   /* borrowed, n*dim */
    double lo[3], h, inv_h;
    int res[3];
    u32* cell_start;     /* ncells+1 */
    u32* cell_pts;       /* n, point ids grouped by cell */
} orc_grid;

static int grid_coord(const orc_grid* g, double v, int a) {
    int c = (int)floor((v - g->lo[a]) * g->inv_h);
    if (c < 0) c = 0;
    if (c >= g->res[a]) c = g->res[a] - 1;
    return c;
}

static void grid_free(orc_grid* g) {
    free(g->cell_start);
    free(g->cell_pts);
    g->cell_start = NULL;
    g->cell_pts = NULL;
}

static void grid_build(orc_grid* g, int dim, u32 n, const double* pts) {
    g->dim = dim; g->n = n; g->pts = pts;
    double hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int a = 0; a < 3; ++a) g->lo[a] = DBL_MAX;
    int gd = dim < 3 ? dim : 3;
    for (u32 i = 0; i < n; ++i)
        for (int a = 0; a < gd; ++a) {
            double v = pts[(size_t)i * dim + a];
            if (v < g->lo[a]) g->lo[a] = v;
            if (v > hi[a]) hi[a] = v;
        }
    for (int a = gd; a < 3; ++a) { g->lo[a] = 0.0; hi[a] = 0.0; }
    double ext[3], maxext = 0.0;
    for (int a = 0; a < 3; ++a) { ext[a] = hi[a] - g->lo[a]; if (ext[a] > maxext) maxext = ext[a]; }
    if (maxext <= 0.0) maxext = 1.0;
    /* about 8 cells per point, cubic cells */
    double target = 8.0 * (double)(n > 64 ? n : 64);
    double vol = 1.0;
    for (int a = 0; a < 3; ++a) vol *= (ext[a] > maxext * 1e-3 ? ext[a] : maxext * 1e-3);
    double h = cbrt(vol / target);
    g->h = h; g->inv_h = 1.0 / h;
    size_t nc = 1;
    for (int a = 0; a < 3; ++a) {
        g->res[a] = (int)floor(ext[a] * g->inv_h) + 1;
        if (g->res[a] < 1) g->res[a] = 1;
        nc *= (size_t)g->res[a];
    }
    g->cell_start = (u32*)calloc(nc + 1, sizeof(u32));
    g->cell_pts = (u32*)malloc(sizeof(u32) * (n ? n : 1));
    u32* cellof = (u32*)malloc(sizeof(u32) * (n ? n : 1));
    for (u32 i = 0; i < n; ++i) {
        const double* p = pts + (size_t)i * dim;
        int cx = grid_coord(g, gd > 0 ? p[0] : 0.0, 0);
        int cy = grid_coord(g, gd > 1 ? p[1] : 0.0, 1);
        int cz = grid_coord(g, gd > 2 ? p[2] : 0.0, 2);
        u32 c = (u32)((cz * g->res[1] + cy) * g->res[0] + cx);
        cellof[i] = c;
        g->cell_start[c + 1]++;
    }
    for (size_t c = 0; c < nc; ++c) g->cell_start[c + 1] += g->cell_start[c];
    u32* cur = (u32*)malloc(sizeof(u32) * (nc + 1));
    memcpy(cur, g->cell_start, sizeof(u32) * (nc + 1));
    for (u32 i = 0; i < n; ++i) g->cell_pts[cur[cellof[i]]++] = i;
    free(cur);
    free(cellof);
}

/* sorted insertion, ascending distance then ascending index (deterministic order;
 * the reference's order among exact ties is traversal-dependent, kd_tree.h:173-195) */
static void nn_insert(u32* idx, double* d2, u32* count, u32 k, u32 id, double d) {
    u32 n = *count;
    if (n == k) {
        if (d > d2[k - 1] || (d == d2[k - 1] && id > idx[k - 1])) return;
        n = k - 1;
    }
    u32 i = n;
    while (i > 0 && (d2[i - 1] > d || (d2[i - 1] == d && idx[i - 1] > id))) {
        d2[i] = d2[i - 1]; idx[i] = idx[i - 1]; --i;
    }
    d2[i] = d; idx[i] = id;
    *count = n + 1;
}

/* k nearest points to q (q may be one of the points: it is then returned too, as in
 * KdTree::get_nearest_neighbors, kd_tree.cpp:156-189). Returns the number found
 * (min(k, n)). *tie is set when the k-th and (k+1)-th distances are equal, or two
 * returned distances are equal (order among them is then implementation-defined). */
static u32 grid_knn(const orc_grid* g, const double* q, u32 k, u32* idx, double* d2, int* tie) {
    u32 count = 0;
    if (k > g->n) k = g->n;
    if (tie) *tie = 0;
    if (k == 0) return 0;
    int gd = g->dim < 3 ? g->dim : 3;
    double qq[3] = {0, 0, 0};
    for (int a = 0; a < gd; ++a) qq[a] = q[a];
    int c[3];
    for (int a = 0; a < 3; ++a) c[a] = grid_coord(g, qq[a], a);
    int maxr = 0;
    for (int a = 0; a < 3; ++a) {
        if (c[a] > maxr) maxr = c[a];
        if (g->res[a] - 1 - c[a] > maxr) maxr = g->res[a] - 1 - c[a];
    }
    double next_best = DBL_MAX;   /* smallest distance rejected: for tie detection */
    for (int r = 0; r <= maxr; ++r) {
        for (int z = c[2] - r; z <= c[2] + r; ++z) {
            if (z < 0 || z >= g->res[2]) continue;
            for (int y = c[1] - r; y <= c[1] + r; ++y) {
                if (y < 0 || y >= g->res[1]) continue;
                int on_shell_yz = (z == c[2] - r || z == c[2] + r || y == c[1] - r || y == c[1] + r);
                int xstep = on_shell_yz ? 1 : (2 * r > 0 ? 2 * r : 1);
                for (int x = c[0] - r; x <= c[0] + r; x += xstep) {
                    if (x < 0 || x >= g->res[0]) continue;
                    u32 cell = (u32)((z * g->res[1] + y) * g->res[0] + x);
                    for (u32 s = g->cell_start[cell]; s < g->cell_start[cell + 1]; ++s) {
                        u32 id = g->cell_pts[s];
                        double d = distance2(q, g->pts + (size_t)id * g->dim, g->dim);
                        if (count == k) {
                            /* track the best rejected distance */
                            double worst = d2[k - 1];
                            if (d > worst || (d == worst && id > idx[k - 1])) {
                                if (d < next_best) next_best = d;
                                continue;
                            }
                            if (worst < next_best) next_best = worst;
                        }
                        nn_insert(idx, d2, &count, k, id, d);
                    }
                }
            }
        }
        if (count == k) {
            /* every unvisited point lies outside the (2r+1)^3 block around c */
            double b = DBL_MAX;
            for (int a = 0; a < gd; ++a) {
                double lo = g->lo[a] + (double)(c[a] - r) * g->h;
                double hi = g->lo[a] + (double)(c[a] + r + 1) * g->h;
                double dl = qq[a] - lo, dh = hi - qq[a];
                if (c[a] - r > 0 && dl < b) b = dl;
                if (c[a] + r < g->res[a] - 1 && dh < b) b = dh;
            }
            if (b == DBL_MAX) break;
            if (b > 0.0 && d2[k - 1] < b * b * (1.0 - 1e-12)) break;
        }
    }
    if (tie) {
        if (count == k && next_best == d2[k - 1]) *tie = 1;
        for (u32 i = 1; i < count; ++i) if (d2[i] == d2[i - 1]) *tie = 1;
    }
    return count;
}

/* Delaunay_NearestNeighbors::get_neighbors_internal — G/delaunay/delaunay_nn.cpp:105-145.
 * Asks for nb+1 nearest (including i itself), drops i, applies the duplicate rule
 * (:123-134), keeps at most nb. Returns the number of neighbours stored. */
static u32 neighbors_internal(const orc_grid* g, u32 i, u32 nb, u32* out, double* out_d2,
                              u32* wk_idx, double* wk_d2, int* tie) {
    u32 nq = nb + 1;
    if (nq > g->n) nq = g->n;
    u32 got = grid_knn(g, g->pts + (size_t)i * g->dim, nq, wk_idx, wk_d2, tie);
    u32 nres = 0;
    for (u32 j = 0; j < got; ++j) {
        if (wk_idx[j] != i) {
            if (wk_d2[j] == 0.0) {
                if (wk_idx[j] < i) return 0;
            } else {
                out[nres] = wk_idx[j];
                if (out_d2) out_d2[nres] = wk_d2[j];
                nres++;
                if (nres == nq - 1) break;
            }
        }
    }
    return nres;
}

/* ------------------------------------------------------------------------- */
/* Public: kNN                                                                */
/* ------------------------------------------------------------------------- */

/* Delaunay_NearestNeighbors::set_vertices + store_neighbors_CB
 * (G/delaunay/delaunay_nn.cpp:73-103, G/delaunay/delaunay.cpp:259-274).
 * ksize: optional per-seed list sizes (NULL: k for all) — list sizes are sticky in
 * the reference (delaunay.cpp:260-268). idx is S*kstride, padded with 0xffffffff. */
int orc_knn(int dim, u32 S, const double* x, u32 k, const u32* ksize, u32 kstride,
            u32* idx, u32* cnt, double* sqd, uint8_t* tie_flag) {
    orc_grid g;
    grid_build(&g, dim, S, x);
    u32* wi = (u32*)malloc(sizeof(u32) * (kstride + 2));
    double* wd = (double*)malloc(sizeof(double) * (kstride + 2));
    double* od = (double*)malloc(sizeof(double) * (kstride + 2));
    for (u32 i = 0; i < S; ++i) {
        u32 nb = ksize ? ksize[i] : k;
        if (nb > kstride) nb = kstride;
        if (S >= 1 && nb > S - 1) nb = S - 1;
        int tie = 0;
        u32 n = neighbors_internal(&g, i, nb, idx + (size_t)i * kstride, od, wi, wd, &tie);
        for (u32 j = n; j < kstride; ++j) idx[(size_t)i * kstride + j] = ORC_NONE;
        if (sqd) {
            for (u32 j = 0; j < kstride; ++j) sqd[(size_t)i * kstride + j] = j < n ? od[j] : -1.0;
        }
        cnt[i] = n;
        if (tie_flag) tie_flag[i] = (uint8_t)tie;
    }
    free(wi); free(wd); free(od);
    grid_free(&g);
    return 0;
}

/* Delaunay_NearestNeighbors::nearest_vertex — delaunay_nn.cpp:147-149 */
int orc_nearest(int dim, u32 S, const double* x, u32 nq, const double* q, u32* out) {
    orc_grid g;
    grid_build(&g, dim, S, x);
    for (u32 i = 0; i < nq; ++i) {
        u32 id; double d;
        grid_knn(&g, q + (size_t)i * dim, 1, &id, &d, NULL);
        out[i] = id;
    }
    grid_free(&g);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Polygon clipping                                                           */
/* ------------------------------------------------------------------------- */

typedef struct {
    double p[ORC_MAXDIM];
    double w;
    int32_t adj_facet;   /* facet across the edge STARTING at this vertex */
    int32_t adj_seed;    /* seed across the edge ENDING at this vertex    */
    int32_t sym[3];      /* SymbolicVertex (generic_RVD_vertex.h:376-640): sorted set, -(facet+1) / seed+1; symbolic mode only */
    int nsym;            /* 4 = the small_set overflowed */
} orc_vertex;

/* symbolic mode (RestrictedVoronoiDiagram::set_symbolic): on while the RDT is extracted */
static int orc_symbolic = 0;
static struct { u32* tri; u64 cap; u64 n; } orc_rdt_sink = {0, 0, 0};

/* small_set<signed_index_t,3>::insert (generic_RVD_vertex.h:169-203): sorted, no duplicates */
static void sym_insert(orc_vertex* v, int32_t x) {
    int pos = 0;
    while (pos < v->nsym && pos < 3 && v->sym[pos] < x) ++pos;
    if (pos < v->nsym && pos < 3 && v->sym[pos] == x) return;
    if (v->nsym >= 3) { v->nsym = 4; return; }
    for (int i = v->nsym; i > pos; --i) v->sym[i] = v->sym[i - 1];
    v->sym[pos] = x;
    v->nsym++;
}

/* SymbolicVertex::intersect_symbolic (generic_RVD_vertex.h:582-640): sets_intersect(v1, v2) + bisector E; 1 = ok */
static int sym_intersect(orc_vertex* I, const orc_vertex* v1, const orc_vertex* v2, u32 E) {
    int i = 0, j = 0, n1 = v1->nsym < 3 ? v1->nsym : 3, n2 = v2->nsym < 3 ? v2->nsym : 3;
    I->nsym = 0;
    while (i < n1 && j < n2) {
        if (v1->sym[i] < v2->sym[j]) ++i;
        else if (v2->sym[j] < v1->sym[i]) ++j;
        else { I->sym[I->nsym++] = v1->sym[i]; ++i; ++j; }
    }
    sym_insert(I, (int32_t)E + 1);
    return I->nsym == 3;
}

typedef struct {
    int n;
    orc_vertex v[ORC_MAXPOLY];
} orc_polygon;

typedef struct {
    u64 pairs, planes, plane_vertex, intersections, triangles, nonempty_pairs, sr_exits, exhausted;
} orc_counters;

/* GEOGen::Polygon::clip_by_plane_fast<DIM> — G/voronoi/generic_RVD_polygon.h:241-366.
 * Returns 0, or 1 on polygon overflow. */
static int clip_by_plane(const orc_polygon* in, orc_polygon* out, const double* pi, const double* pj,
                         u32 j, int dim, orc_counters* cn) {
    out->n = 0;
    if (in->n == 0) return 0;
    cn->planes++;
    double d = 0.0;
    for (int c = 0; c < dim; ++c) d += (pi[c] + pj[c]) * (pi[c] - pj[c]);
    const orc_vertex* prev = &in->v[in->n - 1];
    double prev_l = 0.0;
    for (int c = 0; c < dim; ++c) prev_l += prev->p[c] * (pi[c] - pj[c]);
    double t = 2.0 * prev_l - d;
    int prev_status = (t > 0.0) - (t < 0.0);
    for (int k = 0; k < in->n; ++k) {
        const orc_vertex* vk = &in->v[k];
        cn->plane_vertex++;
        double l = 0.0;
        for (int c = 0; c < dim; ++c) l += vk->p[c] * (pi[c] - pj[c]);
        t = 2.0 * l - d;
        int status = (t > 0.0) - (t < 0.0);
        if (status != prev_status && prev_status != 0) {
            if (out->n >= ORC_MAXPOLY) return 1;
            orc_vertex* I = &out->v[out->n++];
            cn->intersections++;
            double denom = 2.0 * (prev_l - l);
            double l1, l2;
            if (fabs(denom) < 1e-20) { l1 = 0.5; l2 = 0.5; }
            else { l1 = (d - 2.0 * l) / denom; l2 = 1.0 - l1; }
            for (int c = 0; c < dim; ++c) I->p[c] = l1 * prev->p[c] + l2 * vk->p[c];
            I->w = l1 * prev->w + l2 * vk->w;
            if (status > 0) { I->adj_facet = prev->adj_facet; I->adj_seed = (int32_t)j; }
            else { I->adj_facet = -1; I->adj_seed = vk->adj_seed; }
            /* symbolic mode (:305-314): on a failed symbolic intersection the previous vertex is copied into the result
             * (the intersection point is then dropped with it); the edge fields set above are kept as the reference
             * sets them after the copy */
            if (orc_symbolic && !sym_intersect(I, prev, vk, j)) {
                int32_t af = I->adj_facet, as = I->adj_seed;
                *I = *prev;
                I->adj_facet = af; I->adj_seed = as;
            }
        }
        if (status > 0) {
            if (out->n >= ORC_MAXPOLY) return 1;
            out->v[out->n++] = *vk;
        }
        prev = vk; prev_status = status; prev_l = l;
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Surface RVD evaluation                                                     */
/* ------------------------------------------------------------------------- */

typedef struct {
    int dim;
    u32 nv; const double* V;      /* nv*dim */
    u32 nt; const u32* T;         /* nt*3 */
    const int32_t* adj;           /* nt*3: facet across edge (lv, lv+1), -1 on border */
    const double* weights;        /* nv or NULL */
    u32 S; const double* x;       /* S*dim */
    orc_grid grid;
    u32 kcap;                     /* capacity of each neighbour list */
    u32* nbr;                     /* S*kcap */
    u32* nbr_n;                   /* S */
    int check_SR;
    orc_counters cn;
    uint8_t* flags;               /* S: bit0 = neighbour list exhausted before the radius test passed
                                        bit1 = kNN tie, bit2 = polygon overflow */
    u32* wk_idx; double* wk_d2;
} orc_rvd;

#define ORC_FLAG_EXHAUSTED 1
#define ORC_FLAG_TIE 2
#define ORC_FLAG_OVERFLOW 4

/* Delaunay_NearestNeighbors::enlarge_neighborhood — delaunay_nn.cpp:57-71 */
static void enlarge_neighborhood(orc_rvd* R, u32 i, u32 nb) {
    if (nb > R->kcap) nb = R->kcap;
    if (nb > R->nbr_n[i]) {
        int tie = 0;
        R->nbr_n[i] = neighbors_internal(&R->grid, i, nb, R->nbr + (size_t)i * R->kcap, NULL,
                                         R->wk_idx, R->wk_d2, &tie);
        if (tie) R->flags[i] |= ORC_FLAG_TIE;
    }
}

/* GEOGen::RestrictedVoronoiDiagram::clip_by_cell_SR — G/voronoi/generic_RVD.h:2134-2199.
 * ping/pong are indices into P[3] (0 = the facet F, 1/2 = the two work polygons,
 * swap_polygons :2082-2091). Returns the index of the result. */
static int clip_by_cell_SR(orc_rvd* R, u32 i, orc_polygon* P) {
    int ping = 0, pong = 2;
    const double* pi = R->x + (size_t)i * R->dim;
    u32 jj = 0, prev_nb = 0, cur_n = 0;
    while (cur_n < R->S - 1) {
        cur_n = R->nbr_n[i];
        if (cur_n == 0) return ping;
        if (prev_nb == cur_n) return ping;
        for (; jj < cur_n; ++jj) {
            u32 j = R->nbr[(size_t)i * R->kcap + jj];
            double R2 = 0.0;
            for (int k = 0; k < P[ping].n; ++k) {
                double dik = distance2(pi, P[ping].v[k].p, R->dim);
                if (dik > R2) R2 = dik;
            }
            const double* pj = R->x + (size_t)j * R->dim;
            double dij = distance2(pi, pj, R->dim);
            if (dij > 4.1 * R2) { R->cn.sr_exits++; return ping; }
            if (clip_by_plane(&P[ping], &P[pong], pi, pj, j, R->dim, &R->cn)) R->flags[i] |= ORC_FLAG_OVERFLOW;
            if (ping == 0) { ping = 2; pong = 1; } else { int t = ping; ping = pong; pong = t; }
        }
        if (!R->check_SR) {
            if (P[ping].n > 0) { R->flags[i] |= ORC_FLAG_EXHAUSTED; R->cn.exhausted++; }
            return ping;
        }
        u32 nb = cur_n;
        prev_nb = nb;
        if (nb > 8) nb += nb / 8; else nb++;
        if (nb > R->S - 1) nb = R->S - 1;
        if (nb > R->kcap) { R->flags[i] |= ORC_FLAG_EXHAUSTED; R->cn.exhausted++; return ping; }
        enlarge_neighborhood(R, i, nb);
    }
    return ping;
}

typedef struct {
    int mode;          /* 0 = centroids (Lloyd), 1 = func+grad (Newton) */
    double* m;         /* S   (mode 0) */
    double* mg;        /* S*dim (mode 0) */
    double* g;         /* S*dim (mode 1) */
    double* f_seed;    /* S (mode 1, optional) */
    double f;          /* traversal-order sum (mode 1) */
} orc_accum;

/* TriangleAction fan (generic_RVD.h:452-463) + ComputeCentroids[Weighted] (RVD.cpp:280-363)
 * or ComputeCVTFuncGrad[Weighted] (RVD.cpp:575-718) */
static void integrate_polygon(orc_rvd* R, orc_accum* A, u32 v, const orc_polygon* P) {
    int dim = R->dim;
    const double* p0 = R->x + (size_t)v * dim;
    for (int i = 1; i + 1 < P->n; ++i) {
        const orc_vertex* v1 = &P->v[0];
        const orc_vertex* v2 = &P->v[i];
        const orc_vertex* v3 = &P->v[i + 1];
        const double *p1 = v1->p, *p2 = v2->p, *p3 = v3->p;
        R->cn.triangles++;
        if (A->mode == 0) {
            if (R->weights == NULL) {
                double cur_m = triangle_area(p1, p2, p3, dim);
                double s = cur_m / 3.0;
                A->m[v] += cur_m;
                for (int c = 0; c < dim; ++c) A->mg[(size_t)v * dim + c] += s * (p1[c] + p2[c] + p3[c]);
            } else {
                /* Geom::triangle_centroid — geometry_nd.h:178-199 */
                double a = v1->w, b = v2->w, cw = v3->w;
                double abc = a + b + cw;
                double area = triangle_area(p1, p2, p3, dim);
                double Vt = area / 3.0 * abc;
                double wp = a + abc, wq = b + abc, wr = cw + abc;
                double s = area / 12.0;
                A->m[v] += Vt;
                for (int c = 0; c < dim; ++c)
                    A->mg[(size_t)v * dim + c] += s * (wp * p1[c] + wq * p2[c] + wr * p3[c]);
            }
        } else {
            double t_area = triangle_area(p1, p2, p3, dim);
            if (R->weights == NULL) {
                double cur_f = 0.0;
                for (int c = 0; c < dim; ++c) {
                    double u0 = p0[c] - p1[c];
                    double u1 = p0[c] - p2[c];
                    double u2 = p0[c] - p3[c];
                    cur_f += u0 * u0;
                    cur_f += u1 * (u0 + u1);
                    cur_f += u2 * (u0 + u1 + u2);
                }
                double df = t_area * cur_f / 6.0;
                A->f += df;
                if (A->f_seed) A->f_seed[v] += df;
                for (int c = 0; c < dim; ++c) {
                    double Gc = (1.0 / 3.0) * (p1[c] + p2[c] + p3[c]);
                    A->g[(size_t)v * dim + c] += (2.0 * t_area) * (p0[c] - Gc);
                }
            } else {
                double rho[3] = {v1->w, v2->w, v3->w};
                double Sp = rho[0] + rho[1] + rho[2];
                double alpha[3] = {Sp + rho[0], Sp + rho[1], Sp + rho[2]};
                double d00 = 0, d10 = 0, d11 = 0, d20 = 0, d21 = 0, d22 = 0;
                for (int c = 0; c < dim; ++c) {
                    double sp0 = p0[c] - p1[c], sp1 = p0[c] - p2[c], sp2 = p0[c] - p3[c];
                    d00 += sp0 * sp0; d10 += sp1 * sp0; d11 += sp1 * sp1;
                    d20 += sp2 * sp0; d21 += sp2 * sp1; d22 += sp2 * sp2;
                }
                double cur_f = 0.0;
                cur_f += (alpha[0] + rho[0]) * d00;
                cur_f += (alpha[1] + rho[0]) * d10;
                cur_f += (alpha[1] + rho[1]) * d11;
                cur_f += (alpha[2] + rho[0]) * d20;
                cur_f += (alpha[2] + rho[1]) * d21;
                cur_f += (alpha[2] + rho[2]) * d22;
                double df = t_area * cur_f / 30.0;
                A->f += df;
                if (A->f_seed) A->f_seed[v] += df;
                for (int c = 0; c < dim; ++c)
                    A->g[(size_t)v * dim + c] += (t_area / 6.0) *
                        (4.0 * Sp * p0[c] - (alpha[0] * p1[c] + alpha[1] * p2[c] + alpha[2] * p3[c]));
            }
        }
    }
}

/* GEOGen::RestrictedVoronoiDiagram::compute_surfacic_with_seeds_priority —
 * G/voronoi/generic_RVD.h:1318-1424: double flood-fill over the facet graph and the
 * Delaunay 1-skeleton. pairs_out (optional): non-empty (seed, facet) pairs. */
static void surfacic_traversal(orc_rvd* R, orc_accum* A, u32* pairs_out, u64 pairs_cap, u64* npairs_out) {
    u32* seed_stamp = (u32*)malloc(sizeof(u32) * (R->S ? R->S : 1));
    memset(seed_stamp, 0xff, sizeof(u32) * (R->S ? R->S : 1));
    uint8_t* facet_marked = (uint8_t*)calloc(R->nt ? R->nt : 1, 1);
    u32 fs_cap = 1024, fs_n = 0;
    u32* fstack = (u32*)malloc(sizeof(u32) * 2 * fs_cap);   /* (facet, seed) */
    u32 ss_cap = 1024, ss_n = 0;
    u32* sstack = (u32*)malloc(sizeof(u32) * ss_cap);
    orc_polygon* P = (orc_polygon*)malloc(sizeof(orc_polygon) * 3);
    u64 npairs = 0;
    int dim = R->dim;

    for (u32 f = 0; f < R->nt; ++f) {
        if (facet_marked[f]) continue;
        facet_marked[f] = 1;
        /* find_seed_near_facet: nearest seed of the facet's first vertex (:2010-2015) */
        u32 s0; double d0;
        grid_knn(&R->grid, R->V + (size_t)R->T[3 * f] * dim, 1, &s0, &d0, NULL);
        fstack[0] = f; fstack[1] = s0; fs_n = 1;
        while (fs_n > 0) {
            --fs_n;
            u32 cf = fstack[2 * fs_n], cs = fstack[2 * fs_n + 1];
            /* Polygon::initialize_from_mesh_facet — generic_RVD_polygon.cpp:137-148 */
            P[0].n = 3;
            for (int lv = 0; lv < 3; ++lv) {
                u32 vid = R->T[3 * cf + lv];
                memcpy(P[0].v[lv].p, R->V + (size_t)vid * dim, sizeof(double) * dim);
                P[0].v[lv].w = R->weights ? R->weights[vid] : 1.0;
                P[0].v[lv].adj_facet = R->adj[3 * cf + lv];
                P[0].v[lv].adj_seed = -1;
                P[0].v[lv].nsym = 0;
            }
            if (orc_symbolic) {
                /* generic_RVD_polygon.cpp:69-100: corner i2 = {facet, facet across (i1,i2), facet across (i2,i2+1)};
                 * border edges get the "virtual" facets nb_facets + corner */
                for (int i1 = 0; i1 < 3; ++i1) {
                    int i2 = (i1 + 1) % 3;
                    orc_vertex* v2 = &P[0].v[i2];
                    int32_t a1 = P[0].v[i1].adj_facet, a2 = v2->adj_facet;
                    sym_insert(v2, -(int32_t)cf - 1);
                    sym_insert(v2, -(a1 >= 0 ? a1 : (int32_t)(R->nt + i1)) - 1);
                    sym_insert(v2, -(a2 >= 0 ? a2 : (int32_t)(R->nt + i2)) - 1);
                }
            }
            seed_stamp[cs] = cf;
            sstack[0] = cs; ss_n = 1;
            while (ss_n > 0) {
                u32 seed = sstack[--ss_n];
                int res = clip_by_cell_SR(R, seed, P);
                const orc_polygon* Q = &P[res];
                R->cn.pairs++;
                if (Q->n >= 3) {
                    R->cn.nonempty_pairs++;
                    if (pairs_out && npairs < pairs_cap) { pairs_out[2 * npairs] = seed; pairs_out[2 * npairs + 1] = cf; }
                    npairs++;
                }
                integrate_polygon(R, A, seed, Q);
                if (orc_symbolic && orc_rdt_sink.tri) {
                    /* PrimalTriangleAction (generic_RVD.h:575-619): a vertex on two bisectors = triangle (seed, bisector(0),
                     * bisector(1)), bisector(0) being the LAST entry of the sorted set (generic_RVD_vertex.h:481-484) */
                    for (int v = 0; v < Q->n; ++v) {
                        const orc_vertex* ve = &Q->v[v];
                        int nbis = 0;
                        for (int q = 0; q < ve->nsym && q < 3; ++q) nbis += ve->sym[q] > 0;
                        if (nbis == 2) {
                            int top = (ve->nsym < 3 ? ve->nsym : 3) - 1;
                            u32 iv2 = (u32)(ve->sym[top] - 1), iv3 = (u32)(ve->sym[top - 1] - 1);
                            if (seed < iv2 && seed < iv3) {
                                if (orc_rdt_sink.n < orc_rdt_sink.cap) {
                                    u32* o = orc_rdt_sink.tri + 3 * orc_rdt_sink.n;
                                    o[0] = seed; o[1] = iv2; o[2] = iv3;
                                }
                                orc_rdt_sink.n++;
                            }
                        }
                    }
                }
                for (int v = 0; v < Q->n; ++v) {
                    int32_t nf = Q->v[v].adj_facet;
                    if (nf >= 0 && (u32)nf != cf && !facet_marked[nf]) {
                        facet_marked[nf] = 1;
                        if (fs_n == fs_cap) { fs_cap *= 2; fstack = (u32*)realloc(fstack, sizeof(u32) * 2 * fs_cap); }
                        fstack[2 * fs_n] = (u32)nf; fstack[2 * fs_n + 1] = seed; fs_n++;
                    }
                    int32_t ns = Q->v[v].adj_seed;
                    if (ns != -1 && seed_stamp[ns] != cf) {
                        seed_stamp[ns] = cf;
                        if (ss_n == ss_cap) { ss_cap *= 2; sstack = (u32*)realloc(sstack, sizeof(u32) * ss_cap); }
                        sstack[ss_n++] = (u32)ns;
                    }
                }
            }
        }
    }
    if (npairs_out) *npairs_out = npairs;
    free(seed_stamp); free(facet_marked); free(fstack); free(sstack); free(P);
}

/* ------------------------------------------------------------------------- */
/* Volumetric mode: GEOGen::ConvexCell ("Polyhedron") in dual form            */
/* G/voronoi/generic_RVD_cell.h, generic_RVD_cell.cpp (DIM = 3 only here)     */
/* ------------------------------------------------------------------------- */

#define CELL_END 0xffffffffu            /* END_OF_LIST / NO_TRIANGLE */
enum { CTRI_USED = 0, CTRI_CONFLICT = 1, CTRI_FREE = 2 };

/* ConvexCell::Triangle (generic_RVD_cell.h:95-131): the dual of a polyhedron vertex */
typedef struct { u32 v[3]; u32 t[3]; u32 next; int status; double p[3]; } orc_ctri;
/* ConvexCell::Vertex (:133-142): the dual of a polyhedron face; id > 0: bisector with seed id-1,
 * id < 0: tet-tet face with tet -id-1, id == 0: mesh border */
typedef struct { int64_t id; int32_t t; } orc_cvert;
typedef struct {
    orc_ctri* tri; u32 nt, tcap;
    orc_cvert* vert; u32 nv, vcap;
    u32 first_free; int dirty;
} orc_cell;

static u32 orc_cell_high_water[2] = {0, 0};   /* most triangle slots / planes one cell ever used (sizing of the GPU cell) */
void orc_cell_high_water_get(u32* out, int reset) { out[0] = orc_cell_high_water[0]; out[1] = orc_cell_high_water[1]; if (reset) orc_cell_high_water[0] = orc_cell_high_water[1] = 0; }
static const u32 c_plus1[3] = {1, 2, 0}, c_minus1[3] = {2, 0, 1};

static void cell_clear(orc_cell* C) { C->first_free = CELL_END; C->nt = 0; C->nv = 0; C->dirty = 0; }   /* :201-207 */
static u32 cell_create_vertex(orc_cell* C) {                                                            /* :863-867 */
    C->dirty = 1;
    if (C->nv == C->vcap) { C->vcap = C->vcap ? 2 * C->vcap : 64; C->vert = (orc_cvert*)realloc(C->vert, sizeof(orc_cvert) * C->vcap); }
    C->vert[C->nv].id = -1; C->vert[C->nv].t = -1;
    if (C->nv + 1 > orc_cell_high_water[1]) orc_cell_high_water[1] = C->nv + 1;
    return C->nv++;
}
static u32 cell_create_triangle(orc_cell* C) {                                                          /* :663-672, grow :1324-1328 */
    if (C->first_free == CELL_END) {
        if (C->nt == C->tcap) { C->tcap = C->tcap ? 2 * C->tcap : 64; C->tri = (orc_ctri*)realloc(C->tri, sizeof(orc_ctri) * C->tcap); }
        orc_ctri* T = &C->tri[C->nt];
        T->next = CELL_END; T->status = CTRI_FREE;
        for (int i = 0; i < 3; ++i) { T->v[i] = CELL_END; T->t[i] = CELL_END; }
        C->first_free = C->nt++;
        if (C->nt > orc_cell_high_water[0]) orc_cell_high_water[0] = C->nt;
    }
    u32 r = C->first_free;
    C->first_free = C->tri[r].next;
    C->tri[r].status = CTRI_USED;
    return r;
}
static u32 cell_find_vertex(const orc_cell* C, u32 t, u32 v) {                                          /* :425-439 */
    return (u32)((C->tri[t].v[1] == v) | ((C->tri[t].v[2] == v) * 2));
}
static u32 cell_adjacent_index(const orc_cell* C, u32 t1, u32 t2) {                                     /* :448-465 */
    return (u32)((C->tri[t1].t[1] == t2) | ((C->tri[t1].t[2] == t2) * 2));
}
static void cell_init_v_to_t(orc_cell* C) {                                                             /* :640-652 */
    C->dirty = 0;
    for (u32 v = 0; v < C->nv; ++v) C->vert[v].t = -1;
    for (u32 t = 0; t < C->nt; ++t)
        if (C->tri[t].status == CTRI_USED)
            for (int iv = 0; iv < 3; ++iv) C->vert[C->tri[t].v[iv]].t = (int32_t)t;
}
static int32_t cell_vertex_triangle(orc_cell* C, u32 v) { if (C->dirty) cell_init_v_to_t(C); return C->vert[v].t; }
typedef struct { u32 t, v; } orc_corner;
static void cell_next_around_vertex(const orc_cell* C, orc_corner* c) {                                 /* :631-636 */
    u32 t2 = C->tri[c->t].t[c_plus1[c->v]];
    u32 v = C->tri[c->t].v[c->v];
    c->v = cell_find_vertex(C, t2, v);
    c->t = t2;
}

/* ConvexCell::initialize_from_mesh_tetrahedron — generic_RVD_cell.cpp:208-282 (non-symbolic part) */
static void cell_init_from_tet(orc_cell* C, const double* V, const u32* T, const int32_t* tadj, u32 t) {
    static const u32 tv[4][3] = {{2, 1, 3}, {3, 0, 2}, {0, 3, 1}, {2, 0, 1}};
    cell_clear(C);
    for (int lf = 0; lf < 4; ++lf) {
        u32 v = cell_create_vertex(C);
        int32_t ta = tadj[4 * (size_t)t + lf];
        C->vert[v].id = (ta < 0) ? 0 : -(int64_t)ta - 1;
    }
    for (int lv = 0; lv < 4; ++lv) {
        u32 k = cell_create_triangle(C);
        for (int i = 0; i < 3; ++i) { C->tri[k].v[i] = tv[lv][i]; C->tri[k].t[i] = tv[lv][i]; }
        memcpy(C->tri[k].p, V + 3 * (size_t)T[4 * (size_t)t + lv], sizeof(double) * 3);
    }
}

/* ConvexCell::clip_by_plane<3>, fast predicates — generic_RVD_cell.h:281-334 with
 * find_furthest_point_linear_scan (:1041-1060), signed_bisector_distance (:1074-1086),
 * propagate_conflict_list (:1105-1143), Vertex::side_fast (generic_RVD_vertex.h:1017-1027),
 * find_triangle_on_border (:1226-1242), triangulate_hole (:894-966), Vertex::intersect_geom
 * (generic_RVD_vertex.h:976-1006), merge_into_free_list (:1310-1322). */
static void cell_clip_by_plane(orc_cell* C, const double* pi, const double* pj, u32 j, orc_counters* cn, u32** stack, u32* stack_cap) {
    u32 new_v = cell_create_vertex(C);
    C->vert[new_v].id = (int64_t)j + 1;
    cn->planes++;
    /* Phase I: furthest point, then flood-fill of the conflict zone */
    u32 furthest = CELL_END; double fd = 0.0;
    for (u32 t = 0; t < C->nt; ++t) {
        if (C->tri[t].status != CTRI_USED) continue;
        const double* q = C->tri[t].p;
        double d = 0.0;
        for (int c = 0; c < 3; ++c) {
            d += (q[c] - pj[c]) * (q[c] - pj[c]);
            d -= (q[c] - pi[c]) * (q[c] - pi[c]);
        }
        cn->plane_vertex++;
        if (d < fd) { furthest = t; fd = d; }
    }
    if (!(fd < 0)) return;
    u32 cbegin = CELL_END, cend = CELL_END;
#define APPEND_CONFLICT(tt) do { C->tri[tt].next = cbegin; C->tri[tt].status = CTRI_CONFLICT; cbegin = (tt); if (cend == CELL_END) cend = (tt); } while (0)
    u32 sn = 0;
    if (*stack_cap == 0) { *stack_cap = 256; *stack = (u32*)malloc(sizeof(u32) * 256); }
    (*stack)[sn++] = furthest;
    APPEND_CONFLICT(furthest);
    while (sn > 0) {
        u32 t = (*stack)[--sn];
        for (u32 e = 0; e < 3; ++e) {
            u32 nb = C->tri[t].t[e];
            if (C->tri[nb].status == CTRI_CONFLICT) continue;
            const double* q = C->tri[nb].p;
            double r = 0.0;
            for (int c = 0; c < 3; ++c) {
                r += (pj[c] - q[c]) * (pj[c] - q[c]);
                r -= (pi[c] - q[c]) * (pi[c] - q[c]);
            }
            cn->intersections++;   /* flood-fill side tests */
            if (r < 0.0) {
                if (sn == *stack_cap) { *stack_cap *= 2; *stack = (u32*)realloc(*stack, sizeof(u32) * *stack_cap); }
                (*stack)[sn++] = nb;
                APPEND_CONFLICT(nb);
            }
        }
    }
#undef APPEND_CONFLICT
    /* Phase II: a conflict triangle with a used neighbour */
    u32 t1 = cbegin, e1 = 0; int found = 0;
    do {
        for (e1 = 0; e1 < 3; ++e1) if (C->tri[C->tri[t1].t[e1]].status == CTRI_USED) { found = 1; break; }
        if (found) break;
        t1 = C->tri[t1].next;
    } while (t1 != CELL_END);
    if (!found) { cell_clear(C); return; }
    /* Phase III: triangulate the hole */
    {
        u32 t = t1, e = e1, t_adj = C->tri[t].t[e];
        u32 new_first = CELL_END, new_prev = CELL_END;
        do {
            u32 v1 = C->tri[t].v[c_plus1[e]], v2 = C->tri[t].v[c_minus1[e]];
            u32 nt = cell_create_triangle(C);
            C->tri[nt].v[0] = new_v; C->tri[nt].v[1] = v1; C->tri[nt].v[2] = v2;
            {   /* intersect_geom(vq1 = dual(t), vq2 = dual(adjacent(t, e)), p1 = pi, p2 = pj) */
                const double* q1 = C->tri[t].p; const double* q2 = C->tri[C->tri[t].t[e]].p;
                double d = 0.0, l1 = 0.0, l2 = 0.0;
                for (int c = 0; c < 3; ++c) {
                    double n = pi[c] - pj[c];
                    d -= n * (pj[c] + pi[c]);
                    l1 += q2[c] * n;
                    l2 += q1[c] * n;
                }
                d = 0.5 * d;
                l1 = fabs(l1 + d); l2 = fabs(l2 + d);
                double l12 = l1 + l2;
                if (l12 > 1e-30) { l1 /= l12; l2 /= l12; } else { l1 = 0.5; l2 = 0.5; }
                for (int c = 0; c < 3; ++c) C->tri[nt].p[c] = l1 * q1[c] + l2 * q2[c];
            }
            C->tri[nt].t[0] = t_adj;
            C->tri[t_adj].t[cell_adjacent_index(C, t_adj, t)] = nt;
            e = c_plus1[e];
            t_adj = C->tri[t].t[e];
            while (C->tri[t_adj].status == CTRI_CONFLICT) {
                t = t_adj;
                e = c_minus1[cell_find_vertex(C, t, v2)];
                t_adj = C->tri[t].t[e];
            }
            if (new_prev == CELL_END) new_first = nt;
            else { C->tri[new_prev].t[1] = nt; C->tri[nt].t[2] = new_prev; }
            new_prev = nt;
        } while (t != t1 || e != e1);
        C->tri[new_prev].t[1] = new_first;
        C->tri[new_first].t[2] = new_prev;
    }
    /* Phase IV: conflict zone -> free list */
    {
        u32 cur = cbegin;
        while (cur != cend) { C->tri[cur].status = CTRI_FREE; cur = C->tri[cur].next; }
        C->tri[cend].status = CTRI_FREE;
        C->tri[cend].next = C->first_free;
        C->first_free = cbegin;
    }
}

/* clip_by_cell_SR(index_t seed, Polyhedron& C) — G/voronoi/generic_RVD.h:2282-2347 */
static void cell_clip_by_cell_SR(orc_rvd* R, u32 i, orc_cell* C, u32** stack, u32* stack_cap) {
    const double* pi = R->x + (size_t)i * 3;
    u32 jj = 0, prev_nb = 0, cur_n = 0;
    while (cur_n < R->S - 1) {
        cur_n = R->nbr_n[i];
        if (cur_n == 0) return;
        if (prev_nb == cur_n) return;
        for (; jj < cur_n; ++jj) {
            u32 j = R->nbr[(size_t)i * R->kcap + jj];
            double R2 = 0.0;
            for (u32 k = 0; k < C->nt; ++k) {
                if (C->tri[k].status != CTRI_USED) continue;
                double dik = distance2(pi, C->tri[k].p, 3);
                if (dik > R2) R2 = dik;
            }
            const double* pj = R->x + (size_t)j * 3;
            double dij = distance2(pi, pj, 3);
            if (dij > 4.1 * R2) { R->cn.sr_exits++; return; }
            cell_clip_by_plane(C, pi, pj, j, &R->cn, stack, stack_cap);
        }
        if (!R->check_SR) {
            int used = 0;
            for (u32 k = 0; k < C->nt; ++k) used |= (C->tri[k].status == CTRI_USED);
            if (used) { R->flags[i] |= ORC_FLAG_EXHAUSTED; R->cn.exhausted++; }
            return;
        }
        u32 nb = cur_n;
        prev_nb = nb;
        if (nb > 8) nb += nb / 8; else nb++;
        if (nb > R->S - 1) nb = R->S - 1;
        if (nb > R->kcap) { R->flags[i] |= ORC_FLAG_EXHAUSTED; R->cn.exhausted++; return; }
        enlarge_neighborhood(R, i, nb);
    }
}

/* Geom::tetra_volume<3> — G/basic/geometry.h:483-524 (vecng.h dot/cross) */
static double tetra_volume3(const double* p1, const double* p2, const double* p3, const double* p4) {
    double U[3], Vv[3], W[3];
    for (int c = 0; c < 3; ++c) { U[c] = p2[c] - p1[c]; Vv[c] = p3[c] - p1[c]; W[c] = p4[c] - p1[c]; }
    double cx = Vv[1] * W[2] - Vv[2] * W[1];
    double cy = Vv[2] * W[0] - Vv[0] * W[2];
    double cz = Vv[0] * W[1] - Vv[1] * W[0];
    return fabs((U[0] * cx + U[1] * cy + U[2] * cz) / 6.0);
}

/* TetrahedronAction (generic_RVD.h:901-978) + ComputeCentroidsVolumetric (RVD.cpp:428-497), or
 * VolumetricIntegrationSimplexAction with visit_inner_tets = false (generic_RVD.h:726-786) +
 * ComputeCVTFuncGradVolumetric (RVD.cpp:791-876). Returns 1 if the cell is non-empty. */
static int integrate_cell(orc_rvd* R, orc_accum* A, u32 v, orc_cell* C) {
    const double* p0 = R->x + (size_t)v * 3;
    u32 t0 = CELL_END;
    for (u32 t = 0; t < C->nt; ++t) if (C->tri[t].status == CTRI_USED) { t0 = t; break; }
    if (t0 == CELL_END) return 0;
    for (u32 cv = 0; cv < C->nv; ++cv) {
        int32_t ct = cell_vertex_triangle(C, cv);
        if (ct == -1) continue;
        int64_t adjacent = C->vert[cv].id;
        if (A->mode == 1 && adjacent < 0) continue;            /* tet-tet face: skipped unless visit_inner_tets */
        orc_corner c1; c1.t = (u32)ct; c1.v = cell_find_vertex(C, (u32)ct, cv);
        if (A->mode == 0) {
            /* facet_is_incident_to_vertex (generic_RVD.h:993-1004) */
            orc_corner cur = c1; int inc = 0;
            do { if (cur.t == t0) { inc = 1; break; } cell_next_around_vertex(C, &cur); } while (cur.t != c1.t || cur.v != c1.v);
            if (inc) continue;
        }
        const double* v1 = C->tri[c1.t].p;
        orc_corner c2 = c1; cell_next_around_vertex(C, &c2);
        orc_corner c3 = c2; cell_next_around_vertex(C, &c3);
        do {
            const double* v2 = C->tri[c2.t].p;
            const double* v3 = C->tri[c3.t].p;
            R->cn.triangles++;
            if (A->mode == 0) {
                const double* q0 = C->tri[t0].p;
                double cur_m = tetra_volume3(q0, v1, v2, v3);
                double s = cur_m / 4.0;
                A->m[v] += cur_m;
                for (int c = 0; c < 3; ++c) A->mg[(size_t)v * 3 + c] += s * (q0[c] + v1[c] + v2[c] + v3[c]);
            } else {
                double mi = tetra_volume3(p0, v1, v2, v3);
                double fi = 0.0;
                for (int c = 0; c < 3; ++c) {
                    double Uc = v1[c] - p0[c], Vc = v2[c] - p0[c], Wc = v3[c] - p0[c];
                    fi += Uc * Uc + Vc * Vc + Wc * Wc;
                    fi += (Uc * Vc + Vc * Wc + Wc * Uc);
                }
                fi *= (mi / 10.0);
                A->f += fi;
                if (A->f_seed) A->f_seed[v] += fi;
                for (int c = 0; c < 3; ++c)
                    A->g[(size_t)v * 3 + c] += 2.0 * mi * (0.75 * p0[c] - 0.25 * v1[c] - 0.25 * v2[c] - 0.25 * v3[c]);
            }
            c2 = c3;
            cell_next_around_vertex(C, &c3);
        } while (c3.t != c1.t || c3.v != c1.v);
    }
    return 1;
}

static struct { int on; u32* tet; u64 cap; u64 n; } orc_rdt_vol_sink = {0, 0, 0, 0};

/* compute_volumetric_with_seeds_priority — G/voronoi/generic_RVD.h:1464-1597. R->T holds 4 vertex ids per tet,
 * R->adj 4 adjacent tets per tet (-1: border; adj[4t+lf] is across the face opposite to local vertex lf). */
static void volumetric_traversal(orc_rvd* R, orc_accum* A, u32* pairs_out, u64 pairs_cap, u64* npairs_out) {
    u32* seed_stamp = (u32*)malloc(sizeof(u32) * (R->S ? R->S : 1));
    memset(seed_stamp, 0xff, sizeof(u32) * (R->S ? R->S : 1));
    uint8_t* tet_marked = (uint8_t*)calloc(R->nt ? R->nt : 1, 1);
    u32 ts_cap = 1024, ts_n = 0;
    u32* tstack = (u32*)malloc(sizeof(u32) * 2 * ts_cap);
    u32 ss_cap = 1024, ss_n = 0;
    u32* sstack = (u32*)malloc(sizeof(u32) * ss_cap);
    u32* fstack = NULL; u32 fstack_cap = 0;
    orc_cell C; memset(&C, 0, sizeof(C)); C.first_free = CELL_END;
    u64 npairs = 0;
    for (u32 t = 0; t < R->nt; ++t) {
        if (tet_marked[t]) continue;
        tet_marked[t] = 1;
        u32 s0; double d0;
        grid_knn(&R->grid, R->V + 3 * (size_t)R->T[4 * (size_t)t], 1, &s0, &d0, NULL);   /* find_seed_near_tet :2024-2028 */
        tstack[0] = t; tstack[1] = s0; ts_n = 1;
        while (ts_n > 0) {
            --ts_n;
            u32 ct = tstack[2 * ts_n], cs = tstack[2 * ts_n + 1];
            seed_stamp[cs] = ct;
            sstack[0] = cs; ss_n = 1;
            while (ss_n > 0) {
                u32 seed = sstack[--ss_n];
                cell_init_from_tet(&C, R->V, R->T, R->adj, ct);
                cell_clip_by_cell_SR(R, seed, &C, &fstack, &fstack_cap);
                R->cn.pairs++;
                if (orc_rdt_vol_sink.on) {
                    /* PrimalTetrahedronAction — G/voronoi/generic_RVD.h:1036-1058: every vertex of the piece that lies on three
                     * bisectors is a Voronoi vertex inside the tet; its Delaunay tet is emitted by its smallest seed */
                    for (u32 it = 0; it < C.nt; ++it) {
                        if (C.tri[it].status != CTRI_USED) continue;
                        int64_t i0 = C.vert[C.tri[it].v[0]].id, i1 = C.vert[C.tri[it].v[1]].id, i2 = C.vert[C.tri[it].v[2]].id;
                        if (i0 > 0 && i1 > 0 && i2 > 0) {
                            u32 v1 = (u32)(i0 - 1), v2 = (u32)(i1 - 1), v3 = (u32)(i2 - 1);
                            /* SymbolicVertex keeps its bisectors sorted (GenericVoronoiDiagram::SymbolicVertex, a sorted
                             * small set): bisector(0) > bisector(1) > bisector(2) in the rows the reference emits */
                            if (v1 < v2) { u32 q = v1; v1 = v2; v2 = q; }
                            if (v2 < v3) { u32 q = v2; v2 = v3; v3 = q; }
                            if (v1 < v2) { u32 q = v1; v1 = v2; v2 = q; }
                            if (seed < v1 && seed < v2 && seed < v3) {
                                if (orc_rdt_vol_sink.n < orc_rdt_vol_sink.cap) {
                                    u32* o = orc_rdt_vol_sink.tet + 4 * orc_rdt_vol_sink.n;
                                    o[0] = seed; o[1] = v1; o[2] = v2; o[3] = v3;
                                }
                                orc_rdt_vol_sink.n++;
                            }
                        }
                    }
                }
                if (integrate_cell(R, A, seed, &C)) {
                    R->cn.nonempty_pairs++;
                    if (pairs_out && npairs < pairs_cap) { pairs_out[2 * npairs] = seed; pairs_out[2 * npairs + 1] = ct; }
                    npairs++;
                }
                for (u32 v = 0; v < C.nv; ++v) {
                    if (cell_vertex_triangle(&C, v) == -1) continue;
                    int64_t id = C.vert[v].id;
                    if (id > 0) {
                        u32 ns = (u32)(id - 1);
                        if (seed_stamp[ns] != ct) {
                            seed_stamp[ns] = ct;
                            if (ss_n == ss_cap) { ss_cap *= 2; sstack = (u32*)realloc(sstack, sizeof(u32) * ss_cap); }
                            sstack[ss_n++] = ns;
                        }
                    } else if (id < 0) {
                        int64_t nt = -id - 1;
                        if (nt < (int64_t)R->nt && nt != (int64_t)ct && !tet_marked[nt]) {
                            tet_marked[nt] = 1;
                            if (ts_n == ts_cap) { ts_cap *= 2; tstack = (u32*)realloc(tstack, sizeof(u32) * 2 * ts_cap); }
                            tstack[2 * ts_n] = (u32)nt; tstack[2 * ts_n + 1] = seed; ts_n++;
                        }
                    }
                }
            }
        }
    }
    if (npairs_out) *npairs_out = npairs;
    free(seed_stamp); free(tet_marked); free(tstack); free(sstack); free(fstack); free(C.tri); free(C.vert);
}

/* tet adjacency as Mesh::cells.connect() produces it: adj[4t+lf] = tet sharing the face opposite to local
 * vertex lf (MeshCellDescriptors tet_descriptor, G/mesh/mesh.cpp), -1 on the border. */
typedef struct { u32 a, b, c, t, lf; } orc_face;
static int face_cmp(const void* x, const void* y) {
    const orc_face* e = (const orc_face*)x; const orc_face* g = (const orc_face*)y;
    if (e->a != g->a) return e->a < g->a ? -1 : 1;
    if (e->b != g->b) return e->b < g->b ? -1 : 1;
    if (e->c != g->c) return e->c < g->c ? -1 : 1;
    if (e->t != g->t) return e->t < g->t ? -1 : 1;
    return 0;
}
int orc_tet_adjacency(u32 nt, const u32* T, int32_t* adj) {
    orc_face* E = (orc_face*)malloc(sizeof(orc_face) * 4 * (size_t)(nt ? nt : 1));
    for (u32 t = 0; t < nt; ++t)
        for (u32 lf = 0; lf < 4; ++lf) {
            u32 q[3]; int n = 0;
            for (u32 lv = 0; lv < 4; ++lv) if (lv != lf) q[n++] = T[4 * (size_t)t + lv];
            if (q[0] > q[1]) { u32 w = q[0]; q[0] = q[1]; q[1] = w; }
            if (q[1] > q[2]) { u32 w = q[1]; q[1] = q[2]; q[2] = w; }
            if (q[0] > q[1]) { u32 w = q[0]; q[0] = q[1]; q[1] = w; }
            orc_face* e = &E[4 * (size_t)t + lf];
            e->a = q[0]; e->b = q[1]; e->c = q[2]; e->t = t; e->lf = lf;
            adj[4 * (size_t)t + lf] = -1;
        }
    qsort(E, 4 * (size_t)nt, sizeof(orc_face), face_cmp);
    for (size_t i = 0; i + 1 < 4 * (size_t)nt; ++i)
        if (E[i].a == E[i + 1].a && E[i].b == E[i + 1].b && E[i].c == E[i + 1].c) {
            adj[4 * (size_t)E[i].t + E[i].lf] = (int32_t)E[i + 1].t;
            adj[4 * (size_t)E[i + 1].t + E[i + 1].lf] = (int32_t)E[i].t;
        }
    free(E);
    return 0;
}

/* RestrictedVoronoiDiagram::set_volumetric for the evaluation entry points below: elements are then tetrahedra
 * (4 vertex ids, 4 adjacent tets each) and dim must be 3. */
static int orc_volumetric_mode = 0;
void orc_set_volumetric(int x) { orc_volumetric_mode = x; }

/* facet adjacency as Mesh::facets.connect() produces it: adj[3f+lv] = facet sharing
 * edge (lv, lv+1), -1 on the border or for non-manifold edges (first match wins). */
typedef struct { u32 a, b, f, lv; } orc_edge;
static int edge_cmp(const void* x, const void* y) {
    const orc_edge* e = (const orc_edge*)x; const orc_edge* g = (const orc_edge*)y;
    if (e->a != g->a) return e->a < g->a ? -1 : 1;
    if (e->b != g->b) return e->b < g->b ? -1 : 1;
    if (e->f != g->f) return e->f < g->f ? -1 : 1;
    return 0;
}
int orc_facet_adjacency(u32 nt, const u32* T, int32_t* adj) {
    orc_edge* E = (orc_edge*)malloc(sizeof(orc_edge) * 3 * (size_t)(nt ? nt : 1));
    for (u32 f = 0; f < nt; ++f)
        for (u32 lv = 0; lv < 3; ++lv) {
            u32 a = T[3 * f + lv], b = T[3 * f + (lv + 1) % 3];
            orc_edge* e = &E[3 * (size_t)f + lv];
            e->a = a < b ? a : b; e->b = a < b ? b : a; e->f = f; e->lv = lv;
            adj[3 * (size_t)f + lv] = -1;
        }
    qsort(E, 3 * (size_t)nt, sizeof(orc_edge), edge_cmp);
    for (size_t i = 0; i + 1 < 3 * (size_t)nt; ++i) {
        if (E[i].a == E[i + 1].a && E[i].b == E[i + 1].b) {
            int manifold = (i + 2 >= 3 * (size_t)nt) || !(E[i + 2].a == E[i].a && E[i + 2].b == E[i].b);
            if (manifold && (i == 0 || !(E[i - 1].a == E[i].a && E[i - 1].b == E[i].b))) {
                adj[3 * (size_t)E[i].f + E[i].lv] = (int32_t)E[i + 1].f;
                adj[3 * (size_t)E[i + 1].f + E[i + 1].lv] = (int32_t)E[i].f;
            }
        }
    }
    free(E);
    return 0;
}

static void rvd_init(orc_rvd* R, int dim, u32 nv, const double* V, u32 nt, const u32* T, const int32_t* adj,
                     const double* weights, u32 S, const double* x, u32 k, u32 kcap, u32* ksize, int check_SR,
                     uint8_t* flags) {
    memset(R, 0, sizeof(*R));
    R->dim = dim; R->nv = nv; R->V = V; R->nt = nt; R->T = T; R->adj = adj; R->weights = weights;
    R->S = S; R->x = x; R->check_SR = check_SR; R->kcap = kcap; R->flags = flags;
    grid_build(&R->grid, dim, S, x);
    R->nbr = (u32*)malloc(sizeof(u32) * (size_t)(S ? S : 1) * kcap);
    R->nbr_n = (u32*)malloc(sizeof(u32) * (S ? S : 1));
    R->wk_idx = (u32*)malloc(sizeof(u32) * (kcap + 2));
    R->wk_d2 = (double*)malloc(sizeof(double) * (kcap + 2));
    for (u32 i = 0; i < S; ++i) {
        u32 nb = ksize ? ksize[i] : k;
        if (nb > kcap) nb = kcap;
        if (nb > S - 1) nb = S - 1;
        int tie = 0;
        R->nbr_n[i] = neighbors_internal(&R->grid, i, nb, R->nbr + (size_t)i * kcap, NULL, R->wk_idx, R->wk_d2, &tie);
        if (tie) flags[i] |= ORC_FLAG_TIE;
    }
}

static void rvd_free(orc_rvd* R, u32* ksize) {
    if (ksize) for (u32 i = 0; i < R->S; ++i) ksize[i] = R->nbr_n[i];
    grid_free(&R->grid);
    free(R->nbr); free(R->nbr_n); free(R->wk_idx); free(R->wk_d2);
}

/* One evaluation = Delaunay::set_vertices (kNN rebuild) + RVD::compute_centroids (mode 0,
 * G/voronoi/RVD.cpp:371-412) or RVD::compute_CVT_func_grad (mode 1, RVD.cpp:726-774).
 * Outputs are ACCUMULATED INTO (caller zeroes), like the reference (CVT.cpp:149-150,328-329).
 * ksize (optional, in/out): sticky per-seed list sizes; counters: 8 u64 (orc_counters). */
int orc_surface_eval(int dim, u32 nv, const double* V, u32 nt, const u32* T, const int32_t* adj,
                     const double* weights, u32 S, const double* x, u32 k, u32 kcap, u32* ksize,
                     int check_SR, int mode, double* m, double* mg, double* f, double* g, double* f_seed,
                     uint8_t* flags, u64* counters, u32* pairs_out, u64 pairs_cap, u64* npairs) {
    if (dim > ORC_MAXDIM) return 1;
    orc_rvd R;
    uint8_t* fl = flags ? flags : (uint8_t*)calloc(S ? S : 1, 1);
    rvd_init(&R, dim, nv, V, nt, T, adj, weights, S, x, k, kcap, ksize, check_SR, fl);
    orc_accum A;
    A.mode = mode; A.m = m; A.mg = mg; A.g = g; A.f_seed = f_seed; A.f = 0.0;
    if (orc_volumetric_mode) {
        if (dim != 3) { rvd_free(&R, ksize); if (!flags) free(fl); return 2; }
        volumetric_traversal(&R, &A, pairs_out, pairs_cap, npairs);
    } else surfacic_traversal(&R, &A, pairs_out, pairs_cap, npairs);
    if (mode == 1 && f) *f += A.f;
    if (counters) memcpy(counters, &R.cn, sizeof(orc_counters));
    rvd_free(&R, ksize);
    if (!flags) free(fl);
    return 0;
}

/* RestrictedVoronoiDiagram::compute_RDT(simplices, embedding, RDTMode(0)) for surfaces — G/voronoi/RVD.cpp:2353-2370:
 * for_each_primal_triangle(GetPrimalTriangles) over the symbolic polygons, check_SR = true as compute_surface sets it
 * (CVT.cpp:194). tri_out: (seed, bisector(0), bisector(1)) rows in traversal order; *n_out may exceed cap. */
int orc_rdt(int dim, u32 nv, const double* V, u32 nt, const u32* T, const int32_t* adj, u32 S, const double* x,
            u32 k, u32 kcap, u32* tri_out, u64 cap, u64* n_out) {
    if (dim > ORC_MAXDIM || orc_volumetric_mode) return 1;
    orc_rvd R;
    uint8_t* fl = (uint8_t*)calloc(S ? S : 1, 1);
    rvd_init(&R, dim, nv, V, nt, T, adj, NULL, S, x, k, kcap, NULL, 1, fl);
    double* m = (double*)calloc(S ? S : 1, sizeof(double));
    double* mg = (double*)calloc((size_t)(S ? S : 1) * dim, sizeof(double));
    orc_accum A;
    A.mode = 0; A.m = m; A.mg = mg; A.g = NULL; A.f_seed = NULL; A.f = 0.0;
    orc_symbolic = 1;
    orc_rdt_sink.tri = tri_out; orc_rdt_sink.cap = cap; orc_rdt_sink.n = 0;
    surfacic_traversal(&R, &A, NULL, 0, NULL);
    orc_symbolic = 0;
    if (n_out) *n_out = orc_rdt_sink.n;
    orc_rdt_sink.tri = NULL; orc_rdt_sink.cap = 0;
    rvd_free(&R, NULL);
    free(m); free(mg); free(fl);
    return 0;
}

/* RestrictedVoronoiDiagram::compute_RDT in volumetric mode — G/voronoi/RVD.cpp:2308-2335: for_each_primal_tetrahedron
 * (GetPrimalTetrahedra), check_SR = true as CentroidalVoronoiTesselation::compute_volume sets it (CVT.cpp:245), then every
 * tet is reoriented with orient_3d (swap of the first two vertices). tet_out: rows (seed, v1, v2, v3) in traversal order;
 * *n_out may exceed cap. The orientation uses a long double determinant (the reference: the exact predicate PCK::orient_3d);
 * *n_uncertain counts the rows whose determinant is below the rounding bound. */
int orc_rdt_volume(u32 nv, const double* V, u32 nt, const u32* T, const int32_t* adj, u32 S, const double* x,
                   u32 k, u32 kcap, u32* tet_out, u64 cap, u64* n_out, u64* n_uncertain) {
    if (!orc_volumetric_mode) return 1;
    orc_rvd R;
    uint8_t* fl = (uint8_t*)calloc(S ? S : 1, 1);
    rvd_init(&R, 3, nv, V, nt, T, adj, NULL, S, x, k, kcap, NULL, 1, fl);
    double* m = (double*)calloc(S ? S : 1, sizeof(double));
    double* mg = (double*)calloc((size_t)(S ? S : 1) * 3, sizeof(double));
    orc_accum A;
    A.mode = 0; A.m = m; A.mg = mg; A.g = NULL; A.f_seed = NULL; A.f = 0.0;
    orc_rdt_vol_sink.on = 1; orc_rdt_vol_sink.tet = tet_out; orc_rdt_vol_sink.cap = cap; orc_rdt_vol_sink.n = 0;
    volumetric_traversal(&R, &A, NULL, 0, NULL);
    orc_rdt_vol_sink.on = 0;
    u64 n = orc_rdt_vol_sink.n, unc = 0;
    for (u64 t = 0; t < n && t < cap; ++t) {
        u32* o = tet_out + 4 * t;
        const double* p1 = x + 3 * (size_t)o[0]; const double* p2 = x + 3 * (size_t)o[1];
        const double* p3 = x + 3 * (size_t)o[2]; const double* p4 = x + 3 * (size_t)o[3];
        long double a[3], b[3], c[3], mag = 0.0L;
        for (int i = 0; i < 3; ++i) {
            a[i] = (long double)p2[i] - p1[i]; b[i] = (long double)p3[i] - p1[i]; c[i] = (long double)p4[i] - p1[i];
        }
        long double det = a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
        mag = fabsl(a[0]) * (fabsl(b[1] * c[2]) + fabsl(b[2] * c[1])) + fabsl(a[1]) * (fabsl(b[0] * c[2]) + fabsl(b[2] * c[0])) +
              fabsl(a[2]) * (fabsl(b[0] * c[1]) + fabsl(b[1] * c[0]));
        if (fabsl(det) <= 1e-17L * mag) unc++;
        /* PCK::orient_3d(p1, p2, p3, p4) = sign det(p2 - p1, p3 - p1, p4 - p1) (numerics/predicates/orient3d.h:4-24) */
        if (det < 0) { u32 tmp = o[0]; o[0] = o[1]; o[1] = tmp; }
    }
    if (n_out) *n_out = n;
    if (n_uncertain) *n_uncertain = unc;
    orc_rdt_vol_sink.tet = NULL; orc_rdt_vol_sink.cap = 0;
    rvd_free(&R, NULL);
    free(m); free(mg); free(fl);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Multinerve RDT: connected components of the restricted Voronoi cells       */
/* ------------------------------------------------------------------------- */

/* FacetSeedMarking (generic_RVD.h): (facet, seed) -> connected component, here an open-addressing table */
typedef struct { u64* key; u32* val; u64 cap, n; } orc_fsmap;
static void fsmap_init(orc_fsmap* M, u64 cap) {
    M->cap = 64; while (M->cap < 2 * cap) M->cap <<= 1;
    M->key = (u64*)malloc(sizeof(u64) * M->cap); M->val = (u32*)malloc(sizeof(u32) * M->cap);
    memset(M->key, 0xff, sizeof(u64) * M->cap); M->n = 0;
}
static void fsmap_grow(orc_fsmap* M);
static u64 fsmap_slot(const orc_fsmap* M, u64 k) {
    u64 h = (k * 0x9E3779B97F4A7C15ull) >> 17;
    u64 i = h & (M->cap - 1);
    while (M->key[i] != ~0ull && M->key[i] != k) i = (i + 1) & (M->cap - 1);
    return i;
}
static int fsmap_get(const orc_fsmap* M, u32 f, u32 s, u32* val) {
    u64 k = ((u64)f << 32) | s, i = fsmap_slot(M, k);
    if (M->key[i] != k) return 0;
    if (val) *val = M->val[i];
    return 1;
}
static void fsmap_put(orc_fsmap* M, u32 f, u32 s, u32 val) {
    if (2 * (M->n + 1) > M->cap) fsmap_grow(M);
    u64 k = ((u64)f << 32) | s, i = fsmap_slot(M, k);
    if (M->key[i] != k) { M->key[i] = k; M->n++; }
    M->val[i] = val;
}
static void fsmap_grow(orc_fsmap* M) {
    orc_fsmap N; fsmap_init(&N, M->cap);
    for (u64 i = 0; i < M->cap; ++i) if (M->key[i] != ~0ull) { u64 j = fsmap_slot(&N, M->key[i]); N.key[j] = M->key[i]; N.val[j] = M->val[i]; N.n++; }
    free(M->key); free(M->val); *M = N;
}

/* GetConnectedComponentsPrimalTriangles (G/voronoi/RVD.cpp:1901-2264) without RDT_SELECT_NEAREST / RDT_PROJECT_ON_SURFACE
 * (CentroidalVoronoiTesselation::compute_surface never sets them, CVT.cpp:180-197) */
#define ORC_UNINITIALIZED 0xffffffffu
#define ORC_MULTI_COMP 0xfffffffeu
#define ORC_ON_BORDER 0xfffffffdu
typedef struct {
    int dim, use_centroids, prefer_seeds;
    const uint8_t* locked;
    u32* tri; u64 tri_cap, tri_n;
    double* vert; u64 vert_cap;          /* rows of dim doubles */
    u32* vert_seed;                       /* optional: seed of every vertex */
    double m; u32 cur_seed, cur_vertex; int on_border;
    u32* seed_to_vertex;
} orc_mn;

static void mn_end_component(orc_mn* A, const orc_rvd* R) {                    /* RVD.cpp:2195-2237 */
    const int dim = A->dim;
    if (A->cur_vertex < A->vert_cap) {
        double* v = A->vert + (size_t)A->cur_vertex * dim;
        if (!A->use_centroids || (A->locked && A->locked[A->cur_seed]) || A->on_border)
            memcpy(v, R->x + (size_t)A->cur_seed * dim, sizeof(double) * dim);
        else {
            double scal = (A->m < 1e-30 ? 0.0 : 1.0 / A->m);
            for (int c = 0; c < dim; ++c) v[c] *= scal;
        }
        if (A->vert_seed) A->vert_seed[A->cur_vertex] = A->cur_seed;
    }
    if (A->prefer_seeds) {
        if (A->on_border) A->seed_to_vertex[A->cur_seed] = ORC_ON_BORDER;
        switch (A->seed_to_vertex[A->cur_seed]) {
        case ORC_UNINITIALIZED: A->seed_to_vertex[A->cur_seed] = A->cur_vertex; break;
        case ORC_ON_BORDER: break;
        default: A->seed_to_vertex[A->cur_seed] = ORC_MULTI_COMP; break;
        }
    }
    A->cur_vertex++;
}

static void mn_action(orc_mn* A, const orc_rvd* R, u32 s1, u32 cf, const orc_polygon* P, int changed, u32 cc, const orc_fsmap* visited) {
    const int dim = A->dim;
    (void)cf;
    if (changed) {                                                               /* RVD.cpp:1961-1967 */
        if (A->cur_seed != 0xffffffffu) mn_end_component(A, R);
        A->cur_seed = s1;                                                        /* begin_connected_component, :2180-2190 */
        if (A->cur_vertex < A->vert_cap) memset(A->vert + (size_t)A->cur_vertex * dim, 0, sizeof(double) * dim);
        A->m = 0.0; A->on_border = 0;
    }
    if (A->prefer_seeds && !A->on_border) {                                     /* :1969-1986 */
        for (int i = 0; i < P->n; ++i) {
            int j = (i + 1) % P->n;
            if (P->v[i].adj_facet == -1 && P->v[j].adj_seed == -1) { A->on_border = 1; break; }
        }
    }
    for (int i = 1; i + 1 < P->n; ++i) {                                        /* :1988-2006 */
        double cur_m = triangle_area(P->v[0].p, P->v[i].p, P->v[i + 1].p, dim);
        if (A->cur_vertex < A->vert_cap) {
            double* v = A->vert + (size_t)A->cur_vertex * dim;
            for (int c = 0; c < dim; ++c) v[c] += cur_m / 3.0 * (P->v[0].p[c] + P->v[i].p[c] + P->v[i + 1].p[c]);
        }
        A->m += cur_m;
    }
    for (int i = 0; i < P->n; ++i) {                                            /* :2017-2039 */
        const orc_vertex* ve = &P->v[i];
        int nbis = 0, nfac = 0;
        for (int q = 0; q < ve->nsym && q < 3; ++q) { nbis += ve->sym[q] > 0; nfac += ve->sym[q] < 0; }
        if (nbis == 2 && nfac >= 1) {
            int top = (ve->nsym < 3 ? ve->nsym : 3) - 1;
            u32 s2 = (u32)(ve->sym[top] - 1), s3 = (u32)(ve->sym[top - 1] - 1);
            u32 f = (u32)(-ve->sym[0] - 1);                                      /* boundary_facet(0) */
            u32 v2, v3;
            if (fsmap_get(visited, f, s2, &v2) && fsmap_get(visited, f, s3, &v3)) {
                if (A->tri_n < A->tri_cap) { u32* o = A->tri + 3 * A->tri_n; o[0] = cc; o[1] = v2; o[2] = v3; }
                A->tri_n++;
            }
        }
    }
}

/* compute_surfacic_with_cnx_priority (G/voronoi/generic_RVD.h:1856-2001) driving the action above, then the tail of the
 * action's destructor (RVD.cpp:2050-2146, prefer-seeds branch). Outputs in the reference's own order: triangles as emitted
 * (duplicates included), one vertex per connected component in order of discovery. */
int orc_rdt_multinerve(int dim, u32 nv, const double* V, u32 nt, const u32* T, const int32_t* adj, u32 S, const double* x,
                       u32 k, u32 kcap, int use_centroids, int prefer_seeds, const uint8_t* locked,
                       u32* tri_out, u64 tri_cap, u64* ntri_out, double* vert_out, u32* vert_seed_out, u64 vert_cap, u64* nvert_out) {
    if (dim > ORC_MAXDIM || orc_volumetric_mode) return 1;
    orc_rvd Rv; orc_rvd* R = &Rv;
    uint8_t* fl = (uint8_t*)calloc(S ? S : 1, 1);
    rvd_init(R, dim, nv, V, nt, T, adj, NULL, S, x, k, kcap, NULL, 1, fl);
    orc_symbolic = 1;
    orc_mn A;
    memset(&A, 0, sizeof(A));
    A.dim = dim; A.use_centroids = use_centroids; A.prefer_seeds = prefer_seeds; A.locked = locked;
    A.tri = tri_out; A.tri_cap = tri_cap; A.vert = vert_out; A.vert_cap = vert_cap; A.vert_seed = vert_seed_out;
    A.cur_seed = 0xffffffffu;
    A.seed_to_vertex = (u32*)malloc(sizeof(u32) * (S ? S : 1));
    memset(A.seed_to_vertex, 0xff, sizeof(u32) * (S ? S : 1));

    u32* facet_stamp = (u32*)malloc(sizeof(u32) * (nt ? nt : 1));
    memset(facet_stamp, 0xff, sizeof(u32) * (nt ? nt : 1));
    orc_fsmap visited; fsmap_init(&visited, (u64)nt * 2 + 64);
    u64 dq_cap = 4096, dq_head = 0, dq_tail = 0;           /* deque<FacetSeed> as a growing array */
    u32* dq = (u32*)malloc(sizeof(u32) * 2 * dq_cap);
    u32 st_cap = 1024, st_n = 0;
    u32* st = (u32*)malloc(sizeof(u32) * st_cap);
    orc_polygon* P = (orc_polygon*)malloc(sizeof(orc_polygon) * 3);
    u32 cc = 0;

    for (u32 f = 0; f < nt; ++f) {
        if (facet_stamp[f] != 0xffffffffu) continue;
        u32 s0; double d0;
        grid_knn(&R->grid, V + (size_t)T[3 * f] * dim, 1, &s0, &d0, NULL);       /* find_seed_near_facet */
        dq_head = dq_tail = 0;
        dq[0] = f; dq[1] = s0; dq_tail = 1;
        while (dq_head < dq_tail) {
            u32 cf = dq[2 * dq_head], cs = dq[2 * dq_head + 1];
            dq_head++;
            if (facet_stamp[cf] == cs) continue;
            if (fsmap_get(&visited, cf, cs, NULL)) continue;
            int changed = 1;
            st[0] = cf; st_n = 1;
            facet_stamp[cf] = cs;
            while (st_n > 0) {
                cf = st[--st_n];
                P[0].n = 3;
                for (int lv = 0; lv < 3; ++lv) {
                    u32 vid = T[3 * cf + lv];
                    memcpy(P[0].v[lv].p, V + (size_t)vid * dim, sizeof(double) * dim);
                    P[0].v[lv].w = 1.0;
                    P[0].v[lv].adj_facet = adj[3 * cf + lv];
                    P[0].v[lv].adj_seed = -1;
                    P[0].v[lv].nsym = 0;
                }
                for (int i1 = 0; i1 < 3; ++i1) {
                    int i2 = (i1 + 1) % 3;
                    orc_vertex* v2 = &P[0].v[i2];
                    int32_t a1 = P[0].v[i1].adj_facet, a2 = v2->adj_facet;
                    sym_insert(v2, -(int32_t)cf - 1);
                    sym_insert(v2, -(a1 >= 0 ? a1 : (int32_t)(nt + i1)) - 1);
                    sym_insert(v2, -(a2 >= 0 ? a2 : (int32_t)(nt + i2)) - 1);
                }
                int res = clip_by_cell_SR(R, cs, P);
                const orc_polygon* Q = &P[res];
                mn_action(&A, R, cs, cf, Q, changed, cc, &visited);
                changed = 0;
                int touches = 0;
                for (int v = 0; v < Q->n; ++v) {
                    int32_t nf = Q->v[v].adj_facet;
                    if (nf >= 0 && (u32)nf < nt && facet_stamp[nf] != cs) {
                        facet_stamp[nf] = cs;
                        if (st_n == st_cap) { st_cap *= 2; st = (u32*)realloc(st, sizeof(u32) * st_cap); }
                        st[st_n++] = (u32)nf;
                    }
                    int32_t ns = Q->v[v].adj_seed;
                    if (ns != -1) {
                        touches = 1;
                        if (!fsmap_get(&visited, cf, (u32)ns, NULL)) {
                            if (dq_tail == dq_cap) { dq_cap *= 2; dq = (u32*)realloc(dq, sizeof(u32) * 2 * dq_cap); }
                            dq[2 * dq_tail] = cf; dq[2 * dq_tail + 1] = (u32)ns; dq_tail++;
                        }
                    }
                }
                if (touches) fsmap_put(&visited, cf, cs, cc);
            }
            ++cc;
        }
    }
    if (A.cur_seed != 0xffffffffu) mn_end_component(&A, R);
    if (A.prefer_seeds) {                                                        /* RVD.cpp:2123-2146 */
        for (u32 s = 0; s < S; ++s) {
            u32 v = A.seed_to_vertex[s];
            if (v != ORC_MULTI_COMP && v != ORC_UNINITIALIZED && v != ORC_ON_BORDER && v < A.vert_cap)
                memcpy(A.vert + (size_t)v * dim, x + (size_t)s * dim, sizeof(double) * dim);
        }
    }
    orc_symbolic = 0;
    if (ntri_out) *ntri_out = A.tri_n;
    if (nvert_out) *nvert_out = A.cur_vertex;
    free(A.seed_to_vertex); free(facet_stamp); free(visited.key); free(visited.val); free(dq); free(st); free(P);
    rvd_free(R, NULL);
    free(fl);
    return 0;
}


/* CentroidalVoronoiTesselation::Lloyd_iterations — G/voronoi/CVT.cpp:133-167 */
int orc_lloyd(int dim, u32 nv, const double* V, u32 nt, const u32* T, const int32_t* adj,
              const double* weights, u32 S, double* x, u32 k, u32 nb_iter, const uint8_t* locked,
              uint8_t* flags_any) {
    double* mg = (double*)malloc(sizeof(double) * (size_t)S * dim);
    double* m = (double*)malloc(sizeof(double) * S);
    uint8_t* fl = (uint8_t*)malloc(S ? S : 1);
    for (u32 it = 0; it < nb_iter; ++it) {
        memset(mg, 0, sizeof(double) * (size_t)S * dim);
        memset(m, 0, sizeof(double) * S);
        memset(fl, 0, S);
        orc_surface_eval(dim, nv, V, nt, T, adj, weights, S, x, k, k, NULL, 0, 0, m, mg, NULL, NULL, NULL,
                         fl, NULL, NULL, 0, NULL);
        if (flags_any) for (u32 j = 0; j < S; ++j) flags_any[j] |= fl[j];
        for (u32 j = 0; j < S; ++j) {
            if (m[j] > 1e-30 && !(locked && locked[j])) {
                double s = 1.0 / m[j];
                for (int c = 0; c < dim; ++c) x[(size_t)j * dim + c] = s * mg[(size_t)j * dim + c];
            }
        }
    }
    free(mg); free(m); free(fl);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* HLBFGS (Yang Liu's HLBFGS 1.2 as vendored: G/third_party/HLBFGS/)          */
/* Only the path geogram takes: INFO[3]=0, INFO[7]=0, INFO[10]=0, INFO[13]=0. */
/* ------------------------------------------------------------------------- */

static double ddot(u32 n, const double* a, const double* b) {   /* HLBFGS_BLAS.cpp:28-40 */
    double r = 0; for (u32 i = 0; i < n; ++i) r += a[i] * b[i]; return r;
}
static void daxpy(u32 n, double al, const double* a, double* y) { /* :42-52 */
    for (u32 i = 0; i < n; ++i) y[i] += al * a[i];
}
static double dnrm2(u32 n, const double* a) {                    /* :54-66 */
    double r = 0; for (u32 i = 0; i < n; ++i) r += a[i] * a[i]; return sqrt(r);
}

/* Moré–Thuente line-search state (the function-statics + keep/rkeep arrays of
 * LineSearch.cpp:10-60, made explicit). */
typedef struct {
    double dg, dgm, dginit, dgtest, dgx, dgxm, dgy, dgym, finit, fm, ftest1, fx, fxm, fy, fym;
    double stmax, stmin, stx, sty, width, width1;
    int infoc, brackt, stage1;
} orc_mcs;

static double dmin(double a, double b) { return a < b ? a : b; }
static double dmax(double a, double b) { return a > b ? a : b; }

/* MCSTEP with SAFE_SEARCH — LineSearch.cpp:232-465 */
static void mcstep(double* stx, double* fx, double* dx, double* sty, double* fy, double* dy,
                   double* stp, const double* fp, const double* dp, int* brackt,
                   const double* stpmin, const double* stpmax, int* info) {
    double p, q, r, s, gama, sgnd, stpc, stpf, stpq, theta, t;
    int bound;
    const double xsafe = .001;
    *info = 0;
    if ((*brackt && (*stp <= dmin(*stx, *sty) || *stp >= dmax(*stx, *sty)))
        || *dx * (*stp - *stx) >= 0. || *stpmax < *stpmin) return;
    sgnd = *dp * (*dx / fabs(*dx));
    if (*fp > *fx) {
        *info = 1; bound = 1;
        theta = (*fx - *fp) * 3 / (*stp - *stx) + *dx + *dp;
        s = dmax(dmax(fabs(theta), fabs(*dx)), fabs(*dp));
        t = theta / s;
        gama = s * sqrt(t * t - *dx / s * (*dp / s));
        if (*stp < *stx) gama = -gama;
        p = gama - *dx + theta;
        q = gama - *dx + gama + *dp;
        r = p / q;
        stpc = *stx + r * (*stp - *stx);
        stpq = *stx + *dx / ((*fx - *fp) / (*stp - *stx) + *dx) / 2 * (*stp - *stx);
        if (fabs(stpc - *stx) < fabs(stpq - *stx)) stpf = stpc;
        else stpf = stpc + (stpq - stpc) / 2;
        if (*stp > *stx) stpf = dmax(*stx + xsafe * (*stp - *stx), stpf);
        else stpf = dmin(*stx + xsafe * (*stp - *stx), stpf);
        *brackt = 1;
    } else if (sgnd < 0.) {
        *info = 2; bound = 0;
        theta = (*fx - *fp) * 3 / (*stp - *stx) + *dx + *dp;
        s = dmax(dmax(fabs(theta), fabs(*dx)), fabs(*dp));
        t = theta / s;
        gama = s * sqrt(t * t - *dx / s * (*dp / s));
        if (*stp > *stx) gama = -gama;
        p = gama - *dp + theta;
        q = gama - *dp + gama + *dx;
        r = p / q;
        stpc = *stp + r * (*stx - *stp);
        stpq = *stp + *dp / (*dp - *dx) * (*stx - *stp);
        if (fabs(stpc - *stp) > fabs(stpq - *stp)) stpf = stpc; else stpf = stpq;
        *brackt = 1;
    } else if (fabs(*dp) < fabs(*dx)) {
        *info = 3; bound = 1;
        theta = (*fx - *fp) * 3 / (*stp - *stx) + *dx + *dp;
        s = dmax(dmax(fabs(theta), fabs(*dx)), fabs(*dp));
        t = theta / s;
        gama = s * sqrt(dmax(0., t * t - *dx / s * (*dp / s)));
        if (*stp > *stx) gama = -gama;
        p = gama - *dp + theta;
        q = gama + (*dx - *dp) + gama;
        r = p / q;
        if (r < 0. && gama != 0.) stpc = *stp + r * (*stx - *stp);
        else if (*stp > *stx) stpc = *stpmax;
        else stpc = *stpmin;
        stpq = *stp + *dp / (*dp - *dx) * (*stx - *stp);
        if (*brackt) { if (fabs(*stp - stpc) < fabs(*stp - stpq)) stpf = stpc; else stpf = stpq; }
        else { if (fabs(*stp - stpc) > fabs(*stp - stpq)) stpf = stpc; else stpf = stpq; }
    } else {
        *info = 4; bound = 0;
        if (*brackt) {
            theta = (*fp - *fy) * 3 / (*sty - *stp) + *dy + *dp;
            s = dmax(dmax(fabs(theta), fabs(*dy)), fabs(*dp));
            t = theta / s;
            gama = s * sqrt(t * t - *dy / s * (*dp / s));
            if (*stp > *sty) gama = -gama;
            p = gama - *dp + theta;
            q = gama - *dp + gama + *dy;
            r = p / q;
            stpc = *stp + r * (*sty - *stp);
            stpf = stpc;
        } else if (*stp > *stx) stpf = *stpmax;
        else stpf = *stpmin;
    }
    sgnd = *dp * (*stx - *stp);     /* SAFE_SEARCH variant (:424-427) */
    if (*fp > *fx) { *sty = *stp; *fy = *fp; *dy = *dp; }
    else {
        if (sgnd < 0.) { *sty = *stx; *fy = *fx; *dy = *dx; }
        *stx = *stp; *fx = *fp; *dx = *dp;
    }
    stpf = dmin(*stpmax, stpf);
    stpf = dmax(*stpmin, stpf);
    *stp = stpf;
    if (*brackt && bound) {
        if (*sty > *stx) *stp = dmin(*stx + (*sty - *stx) * .66, *stp);
        else *stp = dmax(*stx + (*sty - *stx) * .66, *stp);
    }
}

/* MCSRCH — LineSearch.cpp:10-230. info: in -1 = resume after an evaluation, else start.
 * On return info == -1 asks for an evaluation at x. */
static void mcsrch(orc_mcs* L, u32 n, double* x, double f, const double* g, const double* s, double* stp,
                   double ftol, double gtol, double xtol, double stpmin, double stpmax, int maxfev,
                   int* info, int* nfev, double* wa) {
    if (*info != -1) {
        L->infoc = 1;
        if (n == 0 || *stp <= 0. || ftol < 0. || gtol < 0. || xtol < 0. || stpmin < 0. || stpmax < stpmin || maxfev <= 0) return;
        L->dginit = ddot(n, g, s);
        if (L->dginit >= 0.) return;
        L->brackt = 0; L->stage1 = 1; *nfev = 0;
        L->finit = f; L->dgtest = ftol * L->dginit;
        L->width = stpmax - stpmin; L->width1 = L->width / .5;
        memcpy(wa, x, sizeof(double) * n);
        L->stx = 0.; L->fx = L->finit; L->dgx = L->dginit;
        L->sty = 0.; L->fy = L->finit; L->dgy = L->dginit;
    } else {
        *info = 0;
        ++(*nfev);
        L->dg = ddot(n, g, s);
        L->ftest1 = L->finit + *stp * L->dgtest;
        if ((L->brackt && (*stp <= L->stmin || *stp >= L->stmax)) || L->infoc == 0) *info = 6;
        if (*stp == stpmax && f <= L->ftest1 && L->dg <= L->dgtest) *info = 5;
        if (*stp == stpmin && (f > L->ftest1 || L->dg >= L->dgtest)) *info = 4;
        if (*nfev >= maxfev) *info = 3;
        if (L->brackt && L->stmax - L->stmin <= xtol * L->stmax) *info = 2;
        if (f <= L->ftest1 && fabs(L->dg) <= gtol * (-L->dginit)) *info = 1;
        if (*info != 0) return;
        if (L->stage1 && f <= L->ftest1 && L->dg >= dmin(ftol, gtol) * L->dginit) L->stage1 = 0;
        if (L->stage1 && f <= L->fx && f > L->ftest1) {
            L->fm = f - *stp * L->dgtest;
            L->fxm = L->fx - L->stx * L->dgtest;
            L->fym = L->fy - L->sty * L->dgtest;
            L->dgm = L->dg - L->dgtest;
            L->dgxm = L->dgx - L->dgtest;
            L->dgym = L->dgy - L->dgtest;
            mcstep(&L->stx, &L->fxm, &L->dgxm, &L->sty, &L->fym, &L->dgym, stp, &L->fm, &L->dgm,
                   &L->brackt, &L->stmin, &L->stmax, &L->infoc);
            L->fx = L->fxm + L->stx * L->dgtest;
            L->fy = L->fym + L->sty * L->dgtest;
            L->dgx = L->dgxm + L->dgtest;
            L->dgy = L->dgym + L->dgtest;
        } else {
            mcstep(&L->stx, &L->fx, &L->dgx, &L->sty, &L->fy, &L->dgy, stp, &f, &L->dg,
                   &L->brackt, &L->stmin, &L->stmax, &L->infoc);
        }
        if (L->brackt) {
            if (fabs(L->sty - L->stx) >= L->width1 * .66) *stp = L->stx + (L->sty - L->stx) * .5;
            L->width1 = L->width;
            L->width = fabs(L->sty - L->stx);
        }
    }
    /* L30 */
    if (L->brackt) { L->stmin = dmin(L->stx, L->sty); L->stmax = dmax(L->stx, L->sty); }
    else { L->stmin = L->stx; L->stmax = *stp + (*stp - L->stx) * 4.; }
    *stp = dmax(*stp, stpmin);
    *stp = dmin(*stp, stpmax);
    if ((L->brackt && (*stp <= L->stmin || *stp >= L->stmax)) || *nfev >= maxfev - 1 || L->infoc == 0
        || (L->brackt && L->stmax - L->stmin <= xtol * L->stmax)) *stp = L->stx;
    memcpy(x, wa, sizeof(double) * n);
    daxpy(n, *stp, s, x);
    *info = -1;
}

typedef void (*orc_evalfunc)(u32 n, const double* x, double* f, double* g, void* user);
typedef void (*orc_newiter)(u32 iter, u32 nfev, const double* x, double f, const double* g, double gnorm, void* user);

/* HLBFGS() — HLBFGS.cpp:281-587 with the settings of HLBFGSOptimizer::optimize
 * (G/numerics/lbfgs_optimizers.cpp:159-196): PARAMETERS[5]=0, PARAMETERS[6]=epsg,
 * INFO[4]=max_iter. Returns the number of iterations performed. */
int orc_hlbfgs(u32 N, u32 M, double* x, orc_evalfunc evalfunc, orc_newiter newiter, void* user,
               u32 max_iter, double epsg, u32* nfev_total_out) {
    const double ftol = 1.0e-4, xtol = 1.0e-16, gtol = 0.9, stpmin = 1.0e-20, stpmax = 1.0e+20;
    const int maxfev = 20;
    if (N < 1 || max_iter < 1) return 0;
    double* q = (double*)calloc(N, sizeof(double));
    double* g = (double*)calloc(N, sizeof(double));
    double* alpha = (double*)calloc(M ? M : 1, sizeof(double));
    double* rho = (double*)calloc(M ? M : 1, sizeof(double));
    double* s = (double*)calloc((size_t)(M ? M : 1) * N, sizeof(double));
    double* y = (double*)calloc((size_t)(M ? M : 1) * N, sizeof(double));
    double* prev_x = (double*)calloc(N, sizeof(double));
    double* prev_g = (double*)calloc(N, sizeof(double));
    double* wa = (double*)calloc(N, sizeof(double));
    orc_mcs L; memset(&L, 0, sizeof(L));
    double f = 0, stp, gnorm = 0;
    int info, nfev = 0, bound = 0, cur_pos = 0;
    u32 iter = 0, nfev_total = 0;
    for (;;) {
        if (iter == 0) { evalfunc(N, x, &f, g, user); nfev_total++; }
        if (iter > 0 && M > 0) {
            size_t start = (size_t)cur_pos * N;
            for (u32 i = 0; i < N; ++i) { s[start + i] = x[i] - prev_x[i]; y[start + i] = g[i] - prev_g[i]; }
            rho[cur_pos] = 1.0 / ddot(N, &y[start], &s[start]);
        }
        for (u32 i = 0; i < N; ++i) q[i] = -g[i];
        if (iter > 0 && M > 0) {
            bound = iter > M ? (int)M - 1 : (int)iter - 1;
            /* HLBFGS_UPDATE_First_Step — HLBFGS.cpp:157-176 */
            for (int i = bound; i >= 0; --i) {
                int st = iter <= M ? cur_pos - bound + i : (cur_pos - (bound - i) + (int)M) % (int)M;
                alpha[i] = rho[st] * ddot(N, q, &s[(size_t)st * N]);
                daxpy(N, -alpha[i], &y[(size_t)st * N], q);
            }
            /* HLBFGS_UPDATE_Hessian, INFO[3]==0, INFO[12]==1 — HLBFGS.cpp:90-118 */
            {
                size_t start = (size_t)cur_pos * N;
                double ys = ddot(N, &y[start], &s[start]);
                double yy = ddot(N, &y[start], &y[start]);
                double factor = ys / yy;
                for (u32 i = 0; i < N; ++i) q[i] *= factor;
            }
            /* HLBFGS_UPDATE_Second_Step — HLBFGS.cpp:178-196 */
            for (int i = 0; i <= bound; ++i) {
                int st = iter <= M ? i : (cur_pos + 1 + i) % (int)M;
                double tmp = alpha[i] - rho[st] * ddot(N, &y[(size_t)st * N], q);
                daxpy(N, tmp, &s[(size_t)st * N], q);
            }
            cur_pos = (cur_pos + 1) % (int)M;
        }
        memcpy(prev_x, x, sizeof(double) * N);
        memcpy(prev_g, g, sizeof(double) * N);
        if (iter == 0) { gnorm = dnrm2(N, g); stp = 1.0 / gnorm; } else stp = 1;
        info = 0;
        for (;;) {
            mcsrch(&L, N, x, f, g, q, &stp, ftol, gtol, xtol, stpmin, stpmax, maxfev, &info, &nfev, wa);
            if (info != -1) break;
            evalfunc(N, x, &f, g, user);
            nfev_total++;
        }
        gnorm = dnrm2(N, g);
        iter++;
        if (newiter) newiter(iter, nfev_total, x, f, g, gnorm, user);
        double xnorm = dnrm2(N, x);
        xnorm = 1 > xnorm ? 1 : xnorm;
        if (info != 1) break;
        if (gnorm / xnorm <= 0.0) break;
        if (gnorm < epsg) break;
        if (stp < stpmin || stp > stpmax) break;
        if (iter > max_iter) break;
    }
    if (nfev_total_out) *nfev_total_out = nfev_total;
    free(q); free(g); free(alpha); free(rho); free(s); free(y); free(prev_x); free(prev_g); free(wa);
    return (int)iter;
}

/* test_HLBFGS's objective (geogram/src/tests/test_HLBFGS/main.cpp): Rosenbrock */
static void rosenbrock(u32 n, const double* x, double* f, double* g, void* user) {
    (void)user;
    *f = 0.0;
    for (u32 i = 0; i < n; i += 2) {
        double T1 = 1.0 - x[i];
        double T2 = 10.0 * (x[i + 1] - x[i] * x[i]);
        g[i + 1] = 20.0 * T2;
        g[i] = -2.0 * (x[i] * g[i + 1] + T1);
        *f += T1 * T1 + T2 * T2;
    }
}
int orc_hlbfgs_rosenbrock(u32 N, u32 M, double* x, u32 max_iter, double* f_out, u32* nfev) {
    int it = orc_hlbfgs(N, M, x, rosenbrock, NULL, NULL, max_iter, 0.0, nfev);
    double* g = (double*)malloc(sizeof(double) * N);
    rosenbrock(N, x, f_out, g, NULL);
    free(g);
    return it;
}

/* CentroidalVoronoiTesselation::Newton_iterations + funcgrad — G/voronoi/CVT.cpp:272-338 */
typedef struct {
    int dim; u32 nv; const double* V; u32 nt; const u32* T; const int32_t* adj; const double* weights;
    u32 S; u32 k, kcap; u32* ksize; const uint8_t* locked;
    double* f_hist; double* gnorm_hist; u32 hist_cap, hist_n; uint8_t* flags_any;
} orc_newton_ctx;

static void newton_funcgrad(u32 n, const double* x, double* f, double* g, void* user) {
    orc_newton_ctx* C = (orc_newton_ctx*)user;
    memset(g, 0, sizeof(double) * n);
    *f = 0.0;
    uint8_t* fl = (uint8_t*)calloc(C->S ? C->S : 1, 1);
    orc_surface_eval(C->dim, C->nv, C->V, C->nt, C->T, C->adj, C->weights, C->S, x, C->k, C->kcap, C->ksize,
                     1, 1, NULL, NULL, f, g, NULL, fl, NULL, NULL, 0, NULL);
    if (C->flags_any) for (u32 j = 0; j < C->S; ++j) C->flags_any[j] |= fl[j];
    free(fl);
    if (C->locked)   /* constrain_points — CVT.cpp:309-321 */
        for (u32 i = 0; i < C->S; ++i)
            if (C->locked[i]) for (int c = 0; c < C->dim; ++c) g[(size_t)i * C->dim + c] = 0.0;
}
static void newton_newiter(u32 iter, u32 nfev, const double* x, double f, const double* g, double gnorm, void* user) {
    (void)iter; (void)nfev; (void)x; (void)g;
    orc_newton_ctx* C = (orc_newton_ctx*)user;
    if (C->hist_n < C->hist_cap) {
        if (C->f_hist) C->f_hist[C->hist_n] = f;
        if (C->gnorm_hist) C->gnorm_hist[C->hist_n] = gnorm;
    }
    C->hist_n++;
}
int orc_newton(int dim, u32 nv, const double* V, u32 nt, const u32* T, const int32_t* adj, const double* weights,
               u32 S, double* x, u32 k, u32 kcap, u32 nb_iter, u32 m, const uint8_t* locked,
               double* f_hist, double* gnorm_hist, u32 hist_cap, u32* n_iter_out, u32* nfev_out, uint8_t* flags_any) {
    orc_newton_ctx C;
    C.dim = dim; C.nv = nv; C.V = V; C.nt = nt; C.T = T; C.adj = adj; C.weights = weights; C.S = S;
    C.k = k; C.kcap = kcap; C.locked = locked; C.f_hist = f_hist; C.gnorm_hist = gnorm_hist;
    C.hist_cap = hist_cap; C.hist_n = 0; C.flags_any = flags_any;
    C.ksize = (u32*)malloc(sizeof(u32) * (S ? S : 1));
    for (u32 i = 0; i < S; ++i) C.ksize[i] = k;
    int it = orc_hlbfgs(S * (u32)dim, m, x, newton_funcgrad, newton_newiter, &C, nb_iter, 0.0, nfev_out);
    if (n_iter_out) *n_iter_out = (u32)it;
    free(C.ksize);
    return 0;
}


/* ------------------------------------------------------------------------- */
/* Initial sampling: compute_initial_sampling_on_surface / _in_volume          */
/* G/voronoi/RVD.cpp:1658-1698 -> G/mesh/mesh_sampling.h:119-199, 280-360       */
/* ------------------------------------------------------------------------- */

/* std::mt19937_64, default seed 5489 (Numeric::random_reset, G/basic/numeric.cpp:71-73) */
typedef struct { u64 mt[312]; int idx; } orc_mt64;
static void mt64_seed(orc_mt64* g, u64 seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 312; ++i) g->mt[i] = 6364136223846793005ull * (g->mt[i - 1] ^ (g->mt[i - 1] >> 62)) + (u64)i;
    g->idx = 312;
}
static u64 mt64_next(orc_mt64* g) {
    if (g->idx >= 312) {
        for (int i = 0; i < 312; ++i) {
            u64 x = (g->mt[i] & 0xFFFFFFFF80000000ull) | (g->mt[(i + 1) % 312] & 0x7FFFFFFFull);
            u64 xa = x >> 1;
            if (x & 1ull) xa ^= 0xB5026F5AA96619E9ull;
            g->mt[i] = g->mt[(i + 156) % 312] ^ xa;
        }
        g->idx = 0;
    }
    u64 y = g->mt[g->idx++];
    y ^= (y >> 29) & 0x5555555555555555ull;
    y ^= (y << 17) & 0x71D67FFFEDA60000ull;
    y ^= (y << 37) & 0xFFF7EEE000000000ull;
    y ^= y >> 43;
    return y;
}
/* Numeric::random_float64 = std::uniform_real_distribution<double>(0, 1) over mt19937_64 (libstdc++ generate_canonical:
 * one draw, double(x) / 2^64, clamped below 1) */
static double mt64_float64(orc_mt64* g) {
    double r = (double)mt64_next(g) * 0x1p-64;
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    return r;
}
static int cmp_double(const void* a, const void* b) { double x = *(const double*)a, y = *(const double*)b; return (x > y) - (x < y); }

/* mesh_facet_mass<DIM> (mesh_sampling.h:67-96): Geom::triangle_area — the vec3 overload for DIM = 3 (cross product,
 * G/basic/geometry.h:346-372), Heron otherwise (geometry_nd.h:143-156) — or Geom::triangle_mass with vertex weights
 * (geometry_nd.h:237-252: a template defined before geometry.h's vec3 overload is declared, so its qualified call
 * resolves to the Heron template in every dimension). The two area formulas agree to rounding; no test input separates them. */
static double sampling_facet_mass(const double* p1, const double* p2, const double* p3, int dim, const double* w3) {
    double area;
    if (dim == 3 && !w3) {
        double Ux = p2[0] - p1[0], Uy = p2[1] - p1[1], Uz = p2[2] - p1[2];
        double Vx = p3[0] - p1[0], Vy = p3[1] - p1[1], Vz = p3[2] - p1[2];
        double Nx = Uy * Vz - Uz * Vy, Ny = Uz * Vx - Ux * Vz, Nz = Ux * Vy - Uy * Vx;
        area = 0.5 * sqrt(Nx * Nx + Ny * Ny + Nz * Nz);
    } else area = triangle_area(p1, p2, p3, dim);
    if (w3) return area / 3.0 * (sqrt(fabs(w3[0])) + sqrt(fabs(w3[1])) + sqrt(fabs(w3[2])));
    return area;
}
/* mesh_tetra_mass<3> (mesh_sampling.h:213-262): |dot(p2 - p1, cross(p3 - p1, p4 - p1)) / 6| (geometry.h:483-525) */
static double sampling_tet_mass(const double* p1, const double* p2, const double* p3, const double* p4, const double* w4) {
    double a[3], b[3], c[3];
    for (int k = 0; k < 3; ++k) { a[k] = p2[k] - p1[k]; b[k] = p3[k] - p1[k]; c[k] = p4[k] - p1[k]; }
    double cx = b[1] * c[2] - c[1] * b[2], cy = b[2] * c[0] - c[2] * b[0], cz = b[0] * c[1] - c[0] * b[1];
    double v = fabs((a[0] * cx + a[1] * cy + a[2] * cz) / 6.0);
    if (w4) v *= (w4[0] + w4[1] + w4[2] + w4[3]) / 4.0;
    return v;
}

/* mesh_generate_random_samples_on_surface<DIM> / _in_volume<DIM>. per = 3: triangles, 4: tetrahedra (dim 3).
 * x_out: S*dim. elem_out (optional): element of every sample. Returns 0, or 1 when all samples fell into one element. */
int orc_initial_sampling(int dim, u32 nv, const double* V, u32 ne, const u32* E, int per, const double* weights, u32 S,
                         double* x_out, u32* elem_out) {
    (void)nv;
    if (ne == 0 || (per == 4 && dim != 3)) return 2;
    orc_mt64 g; mt64_seed(&g, 5489ull);
    double* s = (double*)malloc(sizeof(double) * (S ? S : 1));
    for (u32 i = 0; i < S; ++i) s[i] = mt64_float64(&g);
    qsort(s, S, sizeof(double), cmp_double);
    double* mass = (double*)malloc(sizeof(double) * ne);
    double Atot = 0.0;
    for (u32 t = 0; t < ne; ++t) {
        const u32* e = E + (size_t)t * per;
        double w[4];
        if (weights) for (int k = 0; k < per; ++k) w[k] = weights[e[k]];
        mass[t] = per == 3 ? sampling_facet_mass(V + (size_t)e[0] * dim, V + (size_t)e[1] * dim, V + (size_t)e[2] * dim, dim, weights ? w : NULL)
                           : sampling_tet_mass(V + (size_t)e[0] * dim, V + (size_t)e[1] * dim, V + (size_t)e[2] * dim, V + (size_t)e[3] * dim, weights ? w : NULL);
        Atot += mass[t];
    }
    u32 first_t = 0xffffffffu, last_t = 0, cur_t = 0;
    double cur_s = mass[0] / Atot;
    for (u32 i = 0; i < S; ++i) {
        while (s[i] > cur_s && cur_t < ne - 1) { cur_t++; cur_s += mass[cur_t] / Atot; }
        if (first_t == 0xffffffffu) first_t = cur_t;
        if (cur_t > last_t) last_t = cur_t;
        const u32* e = E + (size_t)cur_t * per;
        const double* p1 = V + (size_t)e[0] * dim; const double* p2 = V + (size_t)e[1] * dim; const double* p3 = V + (size_t)e[2] * dim;
        if (per == 3) {                                                   /* Geom::random_point_in_triangle, geometry_nd.h:337-349 */
            double l1 = mt64_float64(&g), l2 = mt64_float64(&g);
            if (l1 + l2 > 1.0) { l1 = 1.0 - l1; l2 = 1.0 - l2; }
            double l3 = 1.0 - l1 - l2;
            /* the vec3 overload (G/basic/geometry.h:602-619) weights p2, p3 with the two draws and p1 with the remainder */
            if (dim == 3) for (int c = 0; c < 3; ++c) x_out[(size_t)i * 3 + c] = l3 * p1[c] + l1 * p2[c] + l2 * p3[c];
            else for (int c = 0; c < dim; ++c) x_out[(size_t)i * dim + c] = l1 * p1[c] + l2 * p2[c] + l3 * p3[c];
        } else {                                                          /* Geom::random_point_in_tetra, :363-385 */
            const double* p4 = V + (size_t)e[3] * dim;
            double ss = mt64_float64(&g), tt = mt64_float64(&g), uu = mt64_float64(&g);
            if (ss + tt > 1.0) { ss = 1.0 - ss; tt = 1.0 - tt; }
            if (tt + uu > 1.0) { double tmp = uu; uu = 1.0 - ss - tt; tt = 1.0 - tmp; }
            else if (ss + tt + uu > 1.0) { double tmp = uu; uu = ss + tt + uu - 1.0; ss = 1.0 - tt - tmp; }
            double a = 1.0 - ss - tt - uu;
            for (int c = 0; c < dim; ++c) x_out[(size_t)i * dim + c] = a * p1[c] + ss * p2[c] + tt * p3[c] + uu * p4[c];
        }
        if (elem_out) elem_out[i] = cur_t;
    }
    free(s); free(mass);
    return (ne > 1 && S > 0 && last_t == first_t) ? 1 : 0;
}
