/*
 * oracle/cvt_oracle.c — TEST INFRASTRUCTURE ONLY ("port" oracle).
 *
 * A plain-C, single-threaded CPU restatement of the reference's CVT / restricted
 * Voronoi diagram hot path (geogram as vendored in GraphiteThree). It exists to
 * CHECK the CUDA path; nothing under graphitethree_b200/ may call, link or load it.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it.
 *
 * Parity status: PINNED against the reference itself — tests/test_oracle_vs_reference.py
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
} orc_vertex;

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
    surfacic_traversal(&R, &A, pairs_out, pairs_cap, npairs);
    if (mode == 1 && f) *f += A.f;
    if (counters) memcpy(counters, &R.cn, sizeof(orc_counters));
    rvd_free(&R, ksize);
    if (!flags) free(fl);
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
