/*
 * rt_oracle.c -- CPU ORACLE (test infrastructure, see rt_oracle.h) restating, op for op, the
 * trace! -> segmentize! path of rvignolo/RayTracing.jl v0.2.3.  Every function cites the reference
 * file:line it follows (paths relative to the reference root).  Build: -O2 -ffp-contract=off, no
 * fast-math: Julia does not contract a*b+c into FMA, and +,-,*,/,sqrt are IEEE on both sides.
 *
 * No code is copied from the reference (which is Julia); this is an independent C restatement.
 */
#include "rt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define RTOL_DEFAULT 1.4901161193847656e-8 /* sqrt(eps(Float64)), Base.rtoldefault(Float64) */
#define MAX_ITER 10000                     /* src/track.jl:104 */
#define RUNAWAY_STEPS 200000000L           /* oracle-only guard against a non-terminating walk */

/* ------------------------------------------------------------------------------------------------
 * Julia Base / LinearAlgebra tolerance primitives
 * ---------------------------------------------------------------------------------------------- */

/* Base.isapprox(x::Number, y::Number; atol, rtol) */
int orc_isapprox_scalar(double x, double y, double atol, double rtol) {
    if (x == y) return 1;
    if (!(isfinite(x) && isfinite(y))) return 0;
    double ax = fabs(x), ay = fabs(y);
    double m = ax > ay ? ax : ay;
    double tol = rtol * m;
    if (atol > tol) tol = atol;
    return fabs(x - y) <= tol;
}

static inline double norm2(double a, double b) { return sqrt(a * a + b * b); }

/* LinearAlgebra.isapprox(x::AbstractArray, y::AbstractArray): norm(x-y) <= max(0, rtol*max(norm x, norm y)) */
int orc_isapprox_point(double px, double py, double qx, double qy) {
    double d = norm2(px - qx, py - qy);
    if (!isfinite(d)) return 0; /* nans=false, and Inf distance is never approx for finite inputs */
    double np = norm2(px, py), nq = norm2(qx, qy);
    double m = np > nq ? np : nq;
    double tol = RTOL_DEFAULT * m;
    if (tol < 0.0) tol = 0.0;
    return d <= tol;
}

/* ------------------------------------------------------------------------------------------------
 * geometry primitives
 * ---------------------------------------------------------------------------------------------- */

/* src/intersection.jl:11-18 -- (A,B,C)/norm((A,B,C)), the norm INCLUDES C */
void orc_general_form(double xi, double yi, double xo, double yo, double abc[3]) {
    double A = yi - yo;
    double B = xo - xi;
    double C = xi * yo - xo * yi;
    double n = sqrt(A * A + B * B + C * C);
    abc[0] = A / n;
    abc[1] = B / n;
    abc[2] = C / n;
}

/* src/intersection.jl:127-138 */
int orc_intersection(const double l1[3], const double l2[3], double xy[2]) {
    double a = l1[1] * l2[0];
    double b = l2[1] * l1[0];
    double x = 0.0, y = 0.0;
    int par = orc_isapprox_scalar(a, b, 0.0, RTOL_DEFAULT);
    if (!par) {
        double det = a - b;
        x = (l1[2] * l2[1] - l2[2] * l1[1]) / det;
        y = (l1[0] * l2[2] - l2[0] * l1[2]) / det;
    }
    xy[0] = x;
    xy[1] = y;
    return par;
}

/* src/segment.jl:39-44 */
int orc_point_in_segment(double px, double py, double qx, double qy, double x, double y) {
    double lpx = norm2(px - x, py - y);
    double lqx = norm2(qx - x, qy - y);
    double lpq = norm2(px - qx, py - qy);
    return orc_isapprox_scalar(lpx + lqx, lpq, 0.0, RTOL_DEFAULT);
}

/* ------------------------------------------------------------------------------------------------
 * Mesh (src/mesh.jl:10-83) + an exact nearest-neighbour KD-tree standing in for NearestNeighbors.jl
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
    int32_t lo, hi;      /* range in perm */
    int32_t left, right; /* children or -1 */
    int32_t dim;
    double split;
} kdnode;

struct orc_mesh {
    int32_t n_nodes, n_cells;
    double *xy;
    int32_t *cell_ptrs, *cell_data, *node_cell_ptrs, *node_cell_data;
    double bb_min[2], bb_max[2];
    int32_t *perm; /* 0-based node ids */
    kdnode *kd;
    int32_t n_kd, cap_kd;
};

#define KD_LEAF 8

static void kd_swap(int32_t *a, int32_t *b) {
    int32_t t = *a;
    *a = *b;
    *b = t;
}

/* quickselect on perm[lo,hi) so that perm[mid] has the mid-th smallest coordinate along dim */
static void kd_select(const double *xy, int32_t *perm, int32_t lo, int32_t hi, int32_t mid, int dim) {
    while (hi - lo > 1) {
        int32_t a = lo, b = hi - 1, c = lo + (hi - lo) / 2;
        /* median of three to the end */
        double va = xy[2 * perm[a] + dim], vb = xy[2 * perm[b] + dim], vc = xy[2 * perm[c] + dim];
        int32_t pidx = (va < vb) ? ((vb < vc) ? b : (va < vc ? c : a)) : ((va < vc) ? a : (vb < vc ? c : b));
        double pv = xy[2 * perm[pidx] + dim];
        kd_swap(&perm[pidx], &perm[hi - 1]);
        int32_t s = lo;
        for (int32_t i = lo; i < hi - 1; ++i)
            if (xy[2 * perm[i] + dim] < pv) kd_swap(&perm[i], &perm[s++]);
        kd_swap(&perm[s], &perm[hi - 1]);
        if (s == mid) return;
        if (mid < s)
            hi = s;
        else
            lo = s + 1;
    }
}

static int32_t kd_build(orc_mesh *m, int32_t lo, int32_t hi) {
    if (m->n_kd == m->cap_kd) {
        m->cap_kd = m->cap_kd ? 2 * m->cap_kd : 1024;
        m->kd = (kdnode *)realloc(m->kd, sizeof(kdnode) * (size_t)m->cap_kd);
    }
    int32_t id = m->n_kd++;
    m->kd[id].lo = lo;
    m->kd[id].hi = hi;
    m->kd[id].left = m->kd[id].right = -1;
    m->kd[id].dim = 0;
    m->kd[id].split = 0.0;
    if (hi - lo <= KD_LEAF) return id;
    double mn[2] = {INFINITY, INFINITY}, mx[2] = {-INFINITY, -INFINITY};
    for (int32_t i = lo; i < hi; ++i)
        for (int d = 0; d < 2; ++d) {
            double v = m->xy[2 * m->perm[i] + d];
            if (v < mn[d]) mn[d] = v;
            if (v > mx[d]) mx[d] = v;
        }
    int dim = (mx[1] - mn[1] > mx[0] - mn[0]) ? 1 : 0;
    int32_t mid = lo + (hi - lo) / 2;
    kd_select(m->xy, m->perm, lo, hi, mid, dim);
    double split = m->xy[2 * m->perm[mid] + dim];
    int32_t l = kd_build(m, lo, mid);
    int32_t r = kd_build(m, mid, hi);
    m->kd[id].dim = dim;
    m->kd[id].split = split;
    m->kd[id].left = l;
    m->kd[id].right = r;
    return id;
}

typedef struct {
    int k, n;
    double d2[ORC_MAX_K];
    int32_t id[ORC_MAX_K];
    int32_t skip; /* 0-based id to skip or -1 */
} knn_state;

static inline int cand_less(double d2a, int32_t ida, double d2b, int32_t idb) {
    return d2a < d2b || (d2a == d2b && ida < idb);
}

static void knn_insert(knn_state *s, double d2, int32_t id) {
    if (s->n == s->k && !cand_less(d2, id, s->d2[s->k - 1], s->id[s->k - 1])) return;
    int pos = s->n < s->k ? s->n : s->k - 1;
    while (pos > 0 && cand_less(d2, id, s->d2[pos - 1], s->id[pos - 1])) {
        s->d2[pos] = s->d2[pos - 1];
        s->id[pos] = s->id[pos - 1];
        --pos;
    }
    s->d2[pos] = d2;
    s->id[pos] = id;
    if (s->n < s->k) s->n++;
}

static void kd_search(const orc_mesh *m, int32_t node, double x, double y, knn_state *s) {
    const kdnode *nd = &m->kd[node];
    if (nd->left < 0) {
        for (int32_t i = nd->lo; i < nd->hi; ++i) {
            int32_t id = m->perm[i];
            if (id == s->skip) continue;
            double dx = x - m->xy[2 * id], dy = y - m->xy[2 * id + 1];
            knn_insert(s, dx * dx + dy * dy, id);
        }
        return;
    }
    double diff = (nd->dim == 0 ? x : y) - nd->split;
    int32_t near = diff < 0.0 ? nd->left : nd->right;
    int32_t far = diff < 0.0 ? nd->right : nd->left;
    kd_search(m, near, x, y, s);
    if (s->n < s->k || diff * diff <= s->d2[s->k - 1]) kd_search(m, far, x, y, s);
}

int32_t orc_nn(const orc_mesh *m, double x, double y) {
    knn_state s;
    s.k = 1;
    s.n = 0;
    s.skip = -1;
    kd_search(m, 0, x, y, &s);
    return s.n ? s.id[0] + 1 : -1;
}

/* Diagnostic (off by default, never on in a timed run): does the nearest-node query at (x, y) have two nodes at EXACTLY the same
 * distance?  Only then is the answer of an exact nearest-neighbour search a matter of tie-breaking (lowest node id here and on
 * the device; NearestNeighbors.jl does not document its choice), i.e. only such queries could make find_element (src/mesh.jl:103-146)
 * scan the cells of another node first than the reference does. */
static int g_diag_ties = 0;
void orc_set_diag_ties(int on) { g_diag_ties = on; }

int orc_nn_is_tied(const orc_mesh *m, double x, double y) {
    knn_state s;
    s.k = 2;
    s.n = 0;
    s.skip = -1;
    kd_search(m, 0, x, y, &s);
    return s.n == 2 && s.d2[0] == s.d2[1];
}

int orc_knn(const orc_mesh *m, double x, double y, int k, int32_t skip, int32_t *ids) {
    knn_state s;
    if (k > ORC_MAX_K) k = ORC_MAX_K; /* orc_segmentize rejects k > ORC_MAX_K, so this never truncates a walk */
    s.k = k;
    s.n = 0;
    s.skip = skip - 1;
    kd_search(m, 0, x, y, &s);
    for (int i = 0; i < s.n; ++i) ids[i] = s.id[i] + 1;
    return s.n;
}

int32_t orc_nn_brute(const orc_mesh *m, double x, double y) {
    double best = INFINITY;
    int32_t bi = -1;
    for (int32_t i = 0; i < m->n_nodes; ++i) {
        double dx = x - m->xy[2 * i], dy = y - m->xy[2 * i + 1];
        double d2 = dx * dx + dy * dy;
        if (d2 < best) {
            best = d2;
            bi = i;
        }
    }
    return bi + 1;
}

static void *dup_mem(const void *src, size_t bytes) {
    void *p = malloc(bytes ? bytes : 1);
    if (bytes) memcpy(p, src, bytes);
    return p;
}

/* src/mesh.jl:24-31 (Mesh ctor) and :53-69 (bounding_box = min/max over all node coordinates) */
orc_mesh *orc_mesh_create(int32_t n_nodes, const double *xy, int32_t n_cells, const int32_t *cell_ptrs,
                          const int32_t *cell_data, const int32_t *node_cell_ptrs,
                          const int32_t *node_cell_data) {
    orc_mesh *m = (orc_mesh *)calloc(1, sizeof(orc_mesh));
    m->n_nodes = n_nodes;
    m->n_cells = n_cells;
    m->xy = (double *)dup_mem(xy, sizeof(double) * 2 * (size_t)n_nodes);
    m->cell_ptrs = (int32_t *)dup_mem(cell_ptrs, sizeof(int32_t) * ((size_t)n_cells + 1));
    m->cell_data = (int32_t *)dup_mem(cell_data, sizeof(int32_t) * (size_t)(cell_ptrs[n_cells] - 1));
    m->node_cell_ptrs = (int32_t *)dup_mem(node_cell_ptrs, sizeof(int32_t) * ((size_t)n_nodes + 1));
    m->node_cell_data =
        (int32_t *)dup_mem(node_cell_data, sizeof(int32_t) * (size_t)(node_cell_ptrs[n_nodes] - 1));
    for (int d = 0; d < 2; ++d) {
        double mn = xy[d], mx = xy[d];
        for (int32_t i = 1; i < n_nodes; ++i) {
            double v = xy[2 * i + d];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        m->bb_min[d] = mn;
        m->bb_max[d] = mx;
    }
    m->perm = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_nodes);
    for (int32_t i = 0; i < n_nodes; ++i) m->perm[i] = i;
    kd_build(m, 0, n_nodes);
    return m;
}

void orc_mesh_destroy(orc_mesh *m) {
    if (!m) return;
    free(m->xy);
    free(m->cell_ptrs);
    free(m->cell_data);
    free(m->node_cell_ptrs);
    free(m->node_cell_data);
    free(m->perm);
    free(m->kd);
    free(m);
}

void orc_mesh_bbox(const orc_mesh *m, double bb_min[2], double bb_max[2]) {
    bb_min[0] = m->bb_min[0];
    bb_min[1] = m->bb_min[1];
    bb_max[0] = m->bb_max[0];
    bb_max[1] = m->bb_max[1];
}

/* src/mesh.jl:91-95 -- atol>0 makes Base's default rtol 0 */
int orc_inboundary(const orc_mesh *m, double x, double y, double atol) {
    double rtol = atol > 0.0 ? 0.0 : RTOL_DEFAULT;
    return orc_isapprox_scalar(x, m->bb_max[0], atol, rtol) || orc_isapprox_scalar(x, m->bb_min[0], atol, rtol) ||
           orc_isapprox_scalar(y, m->bb_max[1], atol, rtol) || orc_isapprox_scalar(y, m->bb_min[1], atol, rtol);
}

/* src/mesh.jl:158-176 -- lambda = R \ r with R = [x1 x2 x3; y1 y2 y3; 1 1 1], r = [x, y, 1], using the
 * StaticArrays 3x3 closed form (adjugate rows dotted with r, divided by det = col1 . (col2 x col3));
 * the products with the literal ones are exact and therefore dropped. */
static int pit_nodes(const orc_mesh *m, const int32_t *nid, double x, double y) {
    const double *xy = m->xy;
    double x1 = xy[2 * (nid[0] - 1)], y1 = xy[2 * (nid[0] - 1) + 1];
    double x2 = xy[2 * (nid[1] - 1)], y2 = xy[2 * (nid[1] - 1) + 1];
    double x3 = xy[2 * (nid[2] - 1)], y3 = xy[2 * (nid[2] - 1) + 1];
    double d = x1 * (y2 - y3) + y1 * (x3 - x2) + (x2 * y3 - y2 * x3);
    double l1 = ((y2 - y3) * x + (x3 - x2) * y + (x2 * y3 - x3 * y2)) / d;
    double l2 = ((y3 - y1) * x + (x1 - x3) * y + (x3 * y1 - x1 * y3)) / d;
    double l3 = ((y1 - y2) * x + (x2 - x1) * y + (x1 * y2 - x2 * y1)) / d;
    double lo = 0.0 - RTOL_DEFAULT, hi = 1.0 + RTOL_DEFAULT;
    return (lo <= l1 && l1 <= hi) && (lo <= l2 && l2 <= hi) && (lo <= l3 && l3 <= hi);
}

int orc_point_in_triangle(const orc_mesh *m, int32_t cell, double x, double y) {
    return pit_nodes(m, &m->cell_data[m->cell_ptrs[cell - 1] - 1], x, y);
}

/* src/mesh.jl:184-201 point_in_quadrangle: "look on 4 triangles because we do not know the order of the nodes" -- the triangles
 * (k1, k2, k3) with k_j = mod1(i + j - 1, 4), i = 1..4.  The reference never reaches this function (point_in_element, :149-150,
 * always calls the triangle test, which reads the first three nodes only): SURVEY 8(f)-4 asks for the behaviour it gestures at. */
static int piq_nodes(const orc_mesh *m, const int32_t *nid, double x, double y) {
    for (int i = 0; i < 4; ++i) {
        int32_t t[3];
        for (int j = 0; j < 3; ++j) t[j] = nid[(i + j) % 4];
        if (pit_nodes(m, t, x, y)) return 1;
    }
    return 0;
}

/* point_in_element dispatched on the number of nodes of the cell (the reference's own TODO at src/mesh.jl:148) */
static int pie_cell(const orc_mesh *m, int32_t cell, double x, double y) {
    const int32_t *nid = &m->cell_data[m->cell_ptrs[cell - 1] - 1];
    return (m->cell_ptrs[cell] - m->cell_ptrs[cell - 1] == 4) ? piq_nodes(m, nid, x, y) : pit_nodes(m, nid, x, y);
}

int orc_point_in_element(const orc_mesh *m, int32_t cell, double x, double y) { return pie_cell(m, cell, x, y); }

static int32_t scan_node_cells(const orc_mesh *m, int32_t node, double x, double y) {
    for (int32_t q = m->node_cell_ptrs[node - 1]; q < m->node_cell_ptrs[node]; ++q) {
        int32_t cell = m->node_cell_data[q - 1];
        if (pie_cell(m, cell, x, y)) return cell;
    }
    return -1;
}

/* src/mesh.jl:103-146; *used_knn reports whether the knn branch (:123-132) had to be entered */
static int32_t find_element_ex(const orc_mesh *m, double x, double y, int k, int *used_knn) {
    int32_t nn_id = orc_nn(m, x, y);
    int32_t c = scan_node_cells(m, nn_id, x, y);
    if (c > 0) return c;
    if (used_knn) *used_knn = 1;
    int32_t ids[ORC_MAX_K];
    int n = orc_knn(m, x, y, k, nn_id, ids);
    for (int i = 0; i < n; ++i) {
        c = scan_node_cells(m, ids[i], x, y);
        if (c > 0) return c;
    }
    return -1;
}

int32_t orc_find_element(const orc_mesh *m, double x, double y, int k) {
    return find_element_ex(m, x, y, k, NULL);
}

/* src/intersection.jl:151-159 */
static void order_points(double phi, const double x1[2], int e1, const double x2[2], int e2, double pq[4],
                         int edges[2]) {
    int first;
    if (phi < M_PI / 2)
        first = x1[0] < x2[0];
    else
        first = x1[0] > x2[0];
    const double *a = first ? x1 : x2, *b = first ? x2 : x1;
    pq[0] = a[0];
    pq[1] = a[1];
    pq[2] = b[0];
    pq[3] = b[1];
    edges[0] = first ? e1 : e2;
    edges[1] = first ? e2 : e1;
}

/* src/intersection.jl:34-119 */
int orc_intersections(const orc_mesh *m, int32_t cell, const double abc[3], double phi, double pq[4],
                      int edges[2], int *n_int_out) {
    const int32_t *nid = &m->cell_data[m->cell_ptrs[cell - 1] - 1];
    int nn = m->cell_ptrs[cell] - m->cell_ptrs[cell - 1];
    double ip[4][2];
    int ie[4];
    int n_int = 0, parallel_found = 0;
    for (int i = 0; i < nn; ++i) {
        int j = (i == nn - 1) ? 0 : i + 1;
        double p1x = m->xy[2 * (nid[i] - 1)], p1y = m->xy[2 * (nid[i] - 1) + 1];
        double p2x = m->xy[2 * (nid[j] - 1)], p2y = m->xy[2 * (nid[j] - 1) + 1];
        double L[3], X[2];
        orc_general_form(p1x, p1y, p2x, p2y, L);
        int par = orc_intersection(abc, L, X);
        if (par) {
            parallel_found = 1;
            continue;
        } else if (!orc_point_in_segment(p1x, p1y, p2x, p2y, X[0], X[1])) {
            continue;
        } else {
            if (n_int < 4) {
                ip[n_int][0] = X[0];
                ip[n_int][1] = X[1];
                ie[n_int] = i;
            }
            n_int++;
        }
    }
    if (n_int_out) *n_int_out = n_int;
    edges[0] = edges[1] = -1;
    if (n_int == 3 || n_int == 4) {
        double l = 0.0;
        int s1 = -1, s2 = -1;
        for (int i = 2; i <= n_int; ++i)
            for (int j = i; j <= n_int; ++j) {
                const double *x1 = ip[i - 2], *x2 = ip[j - 1];
                double li = norm2(x1[0] - x2[0], x1[1] - x2[1]);
                if (li > l) {
                    s1 = i - 2;
                    s2 = j - 1;
                    l = li;
                }
            }
        if (s1 < 0) return ORC_ERR_UNDEF; /* x_int1 never assigned: UndefVarError in Julia */
        order_points(phi, ip[s1], ie[s1], ip[s2], ie[s2], pq, edges);
        return 0;
    } else if (n_int == 2 && parallel_found) {
        order_points(phi, ip[0], ie[0], ip[1], ie[1], pq, edges);
        return 0;
    } else if (n_int == 2) {
        if (orc_isapprox_point(ip[0][0], ip[0][1], ip[1][0], ip[1][1])) {
            pq[0] = ip[0][0];
            pq[1] = ip[0][1];
            pq[2] = ip[1][0];
            pq[3] = ip[1][1];
            edges[0] = ie[0];
            edges[1] = ie[1];
        } else {
            order_points(phi, ip[0], ie[0], ip[1], ie[1], pq, edges);
        }
        return 0;
    }
    /* n_int in {0,1}: the caller sees isapprox(p,q) and steps (:114-118) */
    pq[0] = pq[1] = pq[2] = pq[3] = 0.0;
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * TrackGenerator / trace!
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
    double px, py, qx, qy, len;
    int32_t element;
} orc_seg;

typedef struct {
    orc_seg *s;
    int32_t n, cap;
    int32_t status;
    int32_t done;
} orc_segvec;

struct orc_tg {
    const orc_mesh *mesh;
    int n_azim, n2, n4;
    double delta, tiny_step;
    int32_t bcs[4]; /* top,bottom,right,left */
    int64_t *ntx, *nty, *nt, *base; /* per angle; base = uid offset (0-based count of earlier tracks) */
    int64_t n_total;
    /* quadrature */
    double *phis, *deltas, *weights, *dxe, *dye, *sinp, *cosp, *tanp;
    int traced;
    /* tracks SoA (uid-1) */
    int64_t *t_azim, *t_idx, *t_nfwd, *t_nbwd;
    double *t_p, *t_q, *t_phi, *t_len, *t_abc;
    int8_t *t_bcf, *t_bcb, *t_df, *t_db;
    /* segments */
    orc_segvec *segs;
    int64_t stats[8];
};

/* src/azimuthal_quad.jl:21-34 + src/trackgenerator.jl:80-125 */
int orc_tg_create(orc_tg **out, const orc_mesh *m, int n_azim, double delta, const int32_t bcs[4],
                  double tiny_step) {
    *out = NULL;
    if (!(n_azim > 0)) return ORC_E_NAZIM_POS;
    if (n_azim % 4 != 0) return ORC_E_NAZIM_MULT4;
    if (!(delta > 0)) return ORC_E_DELTA_POS;
    orc_tg *t = (orc_tg *)calloc(1, sizeof(orc_tg));
    t->mesh = m;
    t->n_azim = n_azim;
    t->n2 = n_azim / 2;
    t->n4 = n_azim / 4;
    t->delta = delta;
    t->tiny_step = tiny_step;
    memcpy(t->bcs, bcs, sizeof(int32_t) * 4);
    int n2 = t->n2;
    t->ntx = (int64_t *)calloc((size_t)n2, sizeof(int64_t));
    t->nty = (int64_t *)calloc((size_t)n2, sizeof(int64_t));
    t->nt = (int64_t *)calloc((size_t)n2, sizeof(int64_t));
    t->base = (int64_t *)calloc((size_t)n2 + 1, sizeof(int64_t));
    double dx = m->bb_max[0] - m->bb_min[0], dy = m->bb_max[1] - m->bb_min[1]; /* mesh.jl:76,83 */
    for (int i = 1; i <= t->n4; ++i) {
        double phi = M_PI / n2 * (i - 1.0 / 2);
        t->ntx[i - 1] = (int64_t)(floor(dx / delta * fabs(sin(phi))) + 1);
        t->nty[i - 1] = (int64_t)(floor(dy / delta * fabs(cos(phi))) + 1);
        t->nt[i - 1] = t->ntx[i - 1] + t->nty[i - 1];
        int j = n2 - i + 1;
        t->ntx[j - 1] = t->ntx[i - 1];
        t->nty[j - 1] = t->nty[i - 1];
        t->nt[j - 1] = t->nt[i - 1];
    }
    t->n_total = 0;
    for (int i = 0; i < n2; ++i) {
        t->base[i] = t->n_total;
        t->n_total += t->nt[i];
    }
    t->base[n2] = t->n_total;
    t->phis = (double *)calloc((size_t)n2 * 8, sizeof(double));
    t->deltas = t->phis + n2;
    t->weights = t->deltas + n2;
    t->dxe = t->weights + n2;
    t->dye = t->dxe + n2;
    t->sinp = t->dye + n2;
    t->cosp = t->sinp + n2;
    t->tanp = t->cosp + n2;
    *out = t;
    return 0;
}

void orc_seg_free(orc_tg *t) {
    if (!t || !t->segs) return;
    for (int64_t i = 0; i < t->n_total; ++i) free(t->segs[i].s);
    free(t->segs);
    t->segs = NULL;
}

void orc_tg_destroy(orc_tg *t) {
    if (!t) return;
    orc_seg_free(t);
    free(t->ntx);
    free(t->nty);
    free(t->nt);
    free(t->base);
    free(t->phis);
    free(t->t_azim);
    free(t->t_idx);
    free(t->t_nfwd);
    free(t->t_nbwd);
    free(t->t_p);
    free(t->t_q);
    free(t->t_phi);
    free(t->t_len);
    free(t->t_abc);
    free(t->t_bcf);
    free(t->t_bcb);
    free(t->t_df);
    free(t->t_db);
    free(t);
}

int orc_tg_nazim2(const orc_tg *t) { return t->n2; }
int64_t orc_tg_n_total_tracks(const orc_tg *t) { return t->n_total; }
void orc_tg_counts(const orc_tg *t, int64_t *a, int64_t *b, int64_t *c) {
    for (int i = 0; i < t->n2; ++i) {
        if (a) a[i] = t->ntx[i];
        if (b) b[i] = t->nty[i];
        if (c) c[i] = t->nt[i];
    }
}

/* src/boundary.jl:48-63: sides tested in the order top, bottom, right, left */
static int boundary_condition(const orc_tg *t, double x, double y, int32_t *bc) {
    const double *mn = t->mesh->bb_min, *mx = t->mesh->bb_max;
    /* p1=bb_min, p2=(xmin,ymax), p3=bb_max, p4=(xmax,ymin); top=(p2,p3) bottom=(p4,p1) right=(p3,p4)
     * left=(p1,p2)  (src/trackgenerator.jl:172-177) */
    if (orc_point_in_segment(mn[0], mx[1], mx[0], mx[1], x, y))
        *bc = t->bcs[0];
    else if (orc_point_in_segment(mx[0], mn[1], mn[0], mn[1], x, y))
        *bc = t->bcs[1];
    else if (orc_point_in_segment(mx[0], mx[1], mx[0], mn[1], x, y))
        *bc = t->bcs[2];
    else if (orc_point_in_segment(mn[0], mn[1], mn[0], mx[1], x, y))
        *bc = t->bcs[3];
    else
        return ORC_E_NOT_ON_BOUNDARY;
    return 0;
}

/* src/trackgenerator.jl:134-280 (+ next_tracks :282-348, init_weights! azimuthal_quad.jl:35-53) */
int orc_trace(orc_tg *t) {
    const orc_mesh *m = t->mesh;
    int n2 = t->n2, n4 = t->n4;
    double Dx = m->bb_max[0] - m->bb_min[0], Dy = m->bb_max[1] - m->bb_min[1];
    for (int i = 1; i <= n4; ++i) {
        double phi = atan((Dy * (double)t->ntx[i - 1]) / (Dx * (double)t->nty[i - 1]));
        t->phis[i - 1] = phi;
        t->dxe[i - 1] = Dx / (double)t->ntx[i - 1];
        t->dye[i - 1] = Dy / (double)t->nty[i - 1];
        t->deltas[i - 1] = t->dxe[i - 1] * sin(phi);
        int j = n2 - i + 1;
        t->phis[j - 1] = M_PI - phi;
        t->dxe[j - 1] = t->dxe[i - 1];
        t->dye[j - 1] = t->dye[i - 1];
        t->deltas[j - 1] = t->deltas[i - 1];
    }
    /* init_weights!: the isone(i) branch is tested first */
    for (int i = 1; i <= n4; ++i) {
        double w;
        if (i == 1)
            w = t->phis[i] - t->phis[i - 1];
        else if (i == n4)
            w = M_PI - t->phis[i - 1] - t->phis[i - 2];
        else
            w = t->phis[i] - t->phis[i - 2];
        w /= 4 * M_PI;
        t->weights[i - 1] = w;
        t->weights[n2 - i] = w;
    }
    for (int i = 0; i < n2; ++i) {
        t->sinp[i] = sin(t->phis[i]);
        t->cosp[i] = cos(t->phis[i]);
        t->tanp[i] = tan(t->phis[i]);
    }
    size_t n = (size_t)t->n_total;
    if (!t->t_azim) {
        t->t_azim = (int64_t *)malloc(sizeof(int64_t) * n);
        t->t_idx = (int64_t *)malloc(sizeof(int64_t) * n);
        t->t_nfwd = (int64_t *)malloc(sizeof(int64_t) * n);
        t->t_nbwd = (int64_t *)malloc(sizeof(int64_t) * n);
        t->t_p = (double *)malloc(sizeof(double) * 2 * n);
        t->t_q = (double *)malloc(sizeof(double) * 2 * n);
        t->t_phi = (double *)malloc(sizeof(double) * n);
        t->t_len = (double *)malloc(sizeof(double) * n);
        t->t_abc = (double *)malloc(sizeof(double) * 3 * n);
        t->t_bcf = (int8_t *)malloc(n);
        t->t_bcb = (int8_t *)malloc(n);
        t->t_df = (int8_t *)malloc(n);
        t->t_db = (int8_t *)malloc(n);
    }
    const int32_t bc_top = t->bcs[0], bc_bottom = t->bcs[1], bc_right = t->bcs[2], bc_left = t->bcs[3];
    int64_t uid = 1;
    for (int i = 1; i <= n2; ++i) {
        double phi = t->phis[i - 1];
        int right = i <= n4;
        int64_t nx = t->ntx[i - 1], ny = t->nty[i - 1], nt = t->nt[i - 1];
        double dxe = t->dxe[i - 1], dye = t->dye[i - 1];
        for (int64_t j = 1; j <= nt; ++j) {
            double px, py, qx, qy;
            if (j <= nx) {
                px = right ? dxe * ((double)(nx - j) + 1.0 / 2) : dxe * ((double)j - 1.0 / 2);
                py = 0.0;
            } else {
                px = right ? 0.0 : Dx;
                py = dye * ((double)(j - nx) - 1.0 / 2);
            }
            double mm = tan(phi);
            qx = px - (py - Dy) / mm;
            qy = Dy;
            if (!(0 <= qx && qx <= Dx)) {
                if (right) {
                    qx = Dx;
                    qy = py + mm * (Dx - px);
                } else {
                    qx = 0.0;
                    qy = py - mm * px;
                }
                if (!(0 <= qy && qy <= Dy)) return ORC_E_NO_EXIT;
            }
            px += m->bb_min[0];
            py += m->bb_min[1];
            qx += m->bb_min[0];
            qy += m->bb_min[1];
            double len = norm2(px - qx, py - qy);
            double abc[3];
            orc_general_form(px, py, qx, qy, abc);
            int32_t bcf, bcb, bcf1, bcb1;
            if (boundary_condition(t, qx, qy, &bcf)) return ORC_E_NOT_ON_BOUNDARY;
            if (boundary_condition(t, px, py, &bcb)) return ORC_E_NOT_ON_BOUNDARY;
            if (right) {
                bcf1 = j <= ny ? bc_right : bc_top;
                bcb1 = j <= nx ? bc_bottom : bc_left;
            } else {
                bcf1 = j <= ny ? bc_left : bc_top;
                bcb1 = j <= nx ? bc_bottom : bc_right;
            }
            if (bcf != bcf1 || bcb != bcb1) return ORC_E_BC_MISMATCH;
            int8_t df, db;
            if (j <= ny)
                df = 0;
            else
                df = (bcf == 2) ? 0 : 1;
            if (j <= nx)
                db = (bcb == 2) ? 1 : 0;
            else
                db = 1;
            size_t u = (size_t)(uid - 1);
            t->t_azim[u] = i;
            t->t_idx[u] = j;
            t->t_p[2 * u] = px;
            t->t_p[2 * u + 1] = py;
            t->t_q[2 * u] = qx;
            t->t_q[2 * u + 1] = qy;
            t->t_phi[u] = phi;
            t->t_len[u] = len;
            t->t_abc[3 * u] = abc[0];
            t->t_abc[3 * u + 1] = abc[1];
            t->t_abc[3 * u + 2] = abc[2];
            t->t_bcf[u] = (int8_t)bcf;
            t->t_bcb[u] = (int8_t)bcb;
            t->t_df[u] = df;
            t->t_db[u] = db;
            ++uid;
        }
    }
    /* next_track_fwd / next_track_bwd (:294-348); k is the supplementary angle */
    for (int64_t u = 0; u < t->n_total; ++u) {
        int64_t i = t->t_azim[u], j = t->t_idx[u], k = n2 - i + 1;
        int64_t nx = t->ntx[i - 1], ny = t->nty[i - 1], nt = t->nt[i - 1];
        int bcf = t->t_bcf[u], bcb = t->t_bcb[u];
        int64_t ai, aj;
        if (j <= ny) {
            ai = (bcf == 2) ? i : k;
            aj = j + nx;
        } else {
            if (bcf == 2) {
                ai = i;
                aj = j - ny;
            } else {
                ai = k;
                aj = nt + ny - j + 1;
            }
        }
        t->t_nfwd[u] = t->base[ai - 1] + aj;
        if (j <= nx) {
            if (bcb == 2) {
                ai = i;
                aj = j + ny;
            } else {
                ai = k;
                aj = nx - j + 1;
            }
        } else {
            ai = (bcb == 2) ? i : k;
            aj = j - nx;
        }
        t->t_nbwd[u] = t->base[ai - 1] + aj;
    }
    t->traced = 1;
    return 0;
}

void orc_tg_quadrature(const orc_tg *t, double *phis, double *deltas, double *weights) {
    for (int i = 0; i < t->n2; ++i) {
        if (phis) phis[i] = t->phis[i];
        if (deltas) deltas[i] = t->deltas[i];
        if (weights) weights[i] = t->weights[i];
    }
}

void orc_tg_angle_tables(const orc_tg *t, double *s, double *c, double *tn, double *dxe, double *dye) {
    for (int i = 0; i < t->n2; ++i) {
        if (s) s[i] = t->sinp[i];
        if (c) c[i] = t->cosp[i];
        if (tn) tn[i] = t->tanp[i];
        if (dxe) dxe[i] = t->dxe[i];
        if (dye) dye[i] = t->dye[i];
    }
}

#define COPY_IF(dst, src, cnt) \
    if (dst) memcpy(dst, src, sizeof(*(src)) * (size_t)(cnt))

void orc_tg_tracks(const orc_tg *t, int64_t *azim_idx, int64_t *track_idx, double *p, double *q, double *phi,
                   double *len, double *abc, int8_t *bc_fwd, int8_t *bc_bwd, int8_t *dir_fwd, int8_t *dir_bwd,
                   int64_t *next_fwd, int64_t *next_bwd) {
    int64_t n = t->n_total;
    COPY_IF(azim_idx, t->t_azim, n);
    COPY_IF(track_idx, t->t_idx, n);
    COPY_IF(p, t->t_p, 2 * n);
    COPY_IF(q, t->t_q, 2 * n);
    COPY_IF(phi, t->t_phi, n);
    COPY_IF(len, t->t_len, n);
    COPY_IF(abc, t->t_abc, 3 * n);
    COPY_IF(bc_fwd, t->t_bcf, n);
    COPY_IF(bc_bwd, t->t_bcb, n);
    COPY_IF(dir_fwd, t->t_df, n);
    COPY_IF(dir_bwd, t->t_db, n);
    COPY_IF(next_fwd, t->t_nfwd, n);
    COPY_IF(next_bwd, t->t_nbwd, n);
}

/* ------------------------------------------------------------------------------------------------
 * segmentize!
 * ---------------------------------------------------------------------------------------------- */

static void seg_push(orc_segvec *v, double px, double py, double qx, double qy, int32_t element) {
    if (v->n == v->cap) {
        v->cap = v->cap ? 2 * v->cap : 64;
        v->s = (orc_seg *)realloc(v->s, sizeof(orc_seg) * (size_t)v->cap);
    }
    orc_seg *s = &v->s[v->n++];
    s->px = px;
    s->py = py;
    s->qx = qx;
    s->qy = qy;
    s->len = norm2(px - qx, py - qy); /* src/segment.jl:31-33 */
    s->element = element;
}

/* src/track.jl:106-178.  advance_step (src/point.jl:43) is x + step*(cos phi, sin phi). */
static int walk_track(const orc_tg *t, int64_t u, orc_segvec *v, int k, double rtol, int64_t st[8]) {
    const orc_mesh *m = t->mesh;
    double phi = t->t_phi[u];
    int a = (int)t->t_azim[u] - 1;
    double sx = t->tiny_step * t->cosp[a], sy = t->tiny_step * t->sinp[a];
    const double *abc = &t->t_abc[3 * u];
    v->n = 0;
    double xpx = t->t_p[2 * u] + sx, xpy = t->t_p[2 * u + 1] + sy;
    int i = 0;
    int32_t element = -1, prev_element = -1;
    long steps = 0;
    while (i < MAX_ITER) {
        if (++steps > RUNAWAY_STEPS) return ORC_ERR_RUNAWAY;
        st[0]++;
        int used_knn = 0;
        element = find_element_ex(m, xpx, xpy, 2, &used_knn);
        if (g_diag_ties) st[7] += orc_nn_is_tied(m, xpx, xpy);
        if (orc_inboundary(m, xpx, xpy, t->tiny_step)) {
            if (v->n == 0) {
                st[4]++;
                xpx = xpx + sx;
                xpy = xpy + sy;
                continue;
            } else {
                break;
            }
        }
        if (used_knn) st[1]++;
        if (element == -1) {
            st[2]++;
            element = find_element_ex(m, xpx, xpy, k, NULL);
            if (element == -1) return ORC_ERR_TRY_K;
        }
        if (prev_element == element) {
            st[3]++;
            xpx = xpx + sx;
            xpy = xpy + sy;
            continue;
        }
        double pq[4];
        int edges[2], n_int;
        int rc = orc_intersections(m, element, abc, phi, pq, edges, &n_int);
        if (rc) return rc;
        if (n_int >= 3) st[6]++;
        if (orc_isapprox_point(pq[0], pq[1], pq[2], pq[3])) {
            st[5]++;
            xpx = xpx + sx;
            xpy = xpy + sy;
            continue;
        }
        seg_push(v, pq[0], pq[1], pq[2], pq[3], element);
        xpx = pq[2] + sx;
        xpy = pq[3] + sy;
        prev_element = element;
        i += 1;
    }
    /* sum(l.(segments)) -- sequential left-to-right (only feeds an rtol ~1.5e-8 comparison) */
    double sum = 0.0;
    for (int32_t s = 0; s < v->n; ++s) sum += v->s[s].len;
    if (!orc_isapprox_scalar(t->t_len[u], sum, 0.0, rtol)) return ORC_ERR_LENGTH;
    return ORC_OK;
}

int orc_segmentize(orc_tg *t, int k, double rtol, int64_t uid_begin, int64_t uid_end, int nthreads,
                   int64_t *n_segments, int64_t *first_bad_uid) {
    if (!t->traced) return ORC_E_NOT_TRACED; /* src/trackgenerator.jl:360-361 */
    if (k < 1 || k > ORC_MAX_K) return ORC_E_BAD_K;
    if (!t->segs) t->segs = (orc_segvec *)calloc((size_t)t->n_total, sizeof(orc_segvec));
    if (uid_begin < 1) uid_begin = 1;
    if (uid_end > t->n_total + 1) uid_end = t->n_total + 1;
    memset(t->stats, 0, sizeof(t->stats));
    int64_t total = 0;
    int64_t stats[8] = {0};
    (void)nthreads;
#ifdef _OPENMP
    if (nthreads > 1) {
#pragma omp parallel num_threads(nthreads)
        {
            int64_t st[8] = {0};
            int64_t loc = 0;
#pragma omp for schedule(dynamic, 1)
            for (int64_t uid = uid_begin; uid < uid_end; ++uid) {
                orc_segvec *v = &t->segs[uid - 1];
                v->status = walk_track(t, uid - 1, v, k, rtol, st);
                v->done = 1;
                loc += v->n;
            }
#pragma omp critical
            {
                total += loc;
                for (int q = 0; q < 8; ++q) stats[q] += st[q];
            }
        }
    } else
#endif
    {
        for (int64_t uid = uid_begin; uid < uid_end; ++uid) {
            orc_segvec *v = &t->segs[uid - 1];
            v->status = walk_track(t, uid - 1, v, k, rtol, stats);
            v->done = 1;
            total += v->n;
        }
    }
    memcpy(t->stats, stats, sizeof(stats));
    if (n_segments) *n_segments = total;
    int rc = 0;
    if (first_bad_uid) *first_bad_uid = 0;
    for (int64_t uid = uid_begin; uid < uid_end; ++uid)
        if (t->segs[uid - 1].status != ORC_OK) {
            rc = t->segs[uid - 1].status;
            if (first_bad_uid) *first_bad_uid = uid;
            break;
        }
    return rc;
}

void orc_seg_counts(const orc_tg *t, int64_t uid_begin, int64_t uid_end, int64_t *counts, int32_t *status) {
    for (int64_t uid = uid_begin; uid < uid_end; ++uid) {
        const orc_segvec *v = &t->segs[uid - 1];
        if (counts) counts[uid - uid_begin] = v->n;
        if (status) status[uid - uid_begin] = v->status;
    }
}

void orc_seg_copy(const orc_tg *t, int64_t uid_begin, int64_t uid_end, double *px, double *py, double *qx,
                  double *qy, double *len, int32_t *element) {
    size_t o = 0;
    for (int64_t uid = uid_begin; uid < uid_end; ++uid) {
        const orc_segvec *v = &t->segs[uid - 1];
        for (int32_t s = 0; s < v->n; ++s, ++o) {
            if (px) px[o] = v->s[s].px;
            if (py) py[o] = v->s[s].py;
            if (qx) qx[o] = v->s[s].qx;
            if (qy) qy[o] = v->s[s].qy;
            if (len) len[o] = v->s[s].len;
            if (element) element[o] = v->s[s].element;
        }
    }
}

void orc_seg_stats(const orc_tg *t, int64_t stats[8]) { memcpy(stats, t->stats, sizeof(t->stats)); }

/* src/trackgenerator.jl:371-386: uid order, then segment order; divide by n_azim_2 at the end */
void orc_volumes(const orc_tg *t, double *volumes) {
    int32_t nc = t->mesh->n_cells;
    for (int32_t c = 0; c < nc; ++c) volumes[c] = 0.0;
    if (t->segs)
        for (int64_t u = 0; u < t->n_total; ++u) {
            const orc_segvec *v = &t->segs[u];
            if (!v->done) continue;
            double ds = t->deltas[t->t_azim[u] - 1];
            for (int32_t s = 0; s < v->n; ++s) volumes[v->s[s].element - 1] += ds * v->s[s].len;
        }
    for (int32_t c = 0; c < nc; ++c) volumes[c] /= (double)t->n2;
}
