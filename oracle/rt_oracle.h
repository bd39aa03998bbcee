/*
 * rt_oracle.h -- CPU ORACLE for the trace! -> segmentize! hot path of rvignolo/RayTracing.jl v0.2.3.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under raytracing.jl_b200/ may include, link, load or call this.
 * Allowed users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY PINNING: the reference is pure Julia and Julia is absent from this
 * image, so the oracle is a restatement, not the reference itself.  It is pinned against every
 * golden the reference's own tests hold for this path (test/runtests.jl:15-27 track counts and
 * quadrature, :30-43 entry/exit/length invariants, :52-334 BC/link/direction tables) -- see
 * tests/test_oracle_golden.py.  Segment-level results (counts, element order, p/q values) are NOT
 * pinned by any reference golden: "parity unpinned" at segment level.
 *
 * Third-party arithmetic restated from memory of the pinned-by-compat packages (Project.toml:15-23):
 * StaticArrays 1.9 (3x3 closed-form solve, norm), NearestNeighbors 0.4 (exact nn / knn; ties are
 * broken here by lowest node id -- orc_set_diag_ties counts the queries where that matters: none on
 * the named workloads, profiles/r2_nn_ties.txt), Julia Base isapprox.
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_mesh orc_mesh;
typedef struct orc_tg orc_tg;

/* status codes of a segmentized track */
enum { ORC_OK = 0, ORC_ERR_TRY_K = 1, ORC_ERR_LENGTH = 2, ORC_ERR_RUNAWAY = 3, ORC_ERR_UNDEF = 4 };
/* error codes of orc_tg_create / orc_trace */
enum {
    ORC_E_NAZIM_POS = -1, ORC_E_NAZIM_MULT4 = -2, ORC_E_DELTA_POS = -3, ORC_E_NO_EXIT = -4,
    ORC_E_BC_MISMATCH = -5, ORC_E_NOT_ON_BOUNDARY = -6, ORC_E_NOT_TRACED = -7, ORC_E_BAD_K = -8
};
/* largest `k` of segmentize!(t; k) the restatement's fixed-size candidate list holds (the reference has no limit) */
#define ORC_MAX_K 32

/* Mesh (src/mesh.jl:24-31).  CSR tables are 1-based exactly as Gridap stores them. Arrays are copied. */
orc_mesh *orc_mesh_create(int32_t n_nodes, const double *xy, int32_t n_cells, const int32_t *cell_ptrs,
                          const int32_t *cell_data, const int32_t *node_cell_ptrs,
                          const int32_t *node_cell_data);
void orc_mesh_destroy(orc_mesh *m);
void orc_mesh_bbox(const orc_mesh *m, double bb_min[2], double bb_max[2]);

/* primitives exposed for unit tests */
void orc_general_form(double xi, double yi, double xo, double yo, double abc[3]);
int orc_intersection(const double abc1[3], const double abc2[3], double xy[2]); /* returns are_parallel */
int orc_point_in_segment(double px, double py, double qx, double qy, double x, double y);
int orc_isapprox_scalar(double x, double y, double atol, double rtol);
int orc_isapprox_point(double px, double py, double qx, double qy);
int32_t orc_nn(const orc_mesh *m, double x, double y);                       /* 1-based node id */
int orc_knn(const orc_mesh *m, double x, double y, int k, int32_t skip, int32_t *ids); /* sorted */
int32_t orc_nn_brute(const orc_mesh *m, double x, double y);
int orc_point_in_triangle(const orc_mesh *m, int32_t cell, double x, double y);
/* triangles: the test above; 4-node cells: point_in_quadrangle (src/mesh.jl:184-201) */
int orc_point_in_element(const orc_mesh *m, int32_t cell, double x, double y);
int32_t orc_find_element(const orc_mesh *m, double x, double y, int k);      /* 1-based cell or -1 */
int orc_inboundary(const orc_mesh *m, double x, double y, double atol);
/* returns 0 ok / ORC_ERR_UNDEF; pq = px,py,qx,qy; edges = local edge (0..2) that produced p and q or -1 */
int orc_intersections(const orc_mesh *m, int32_t cell, const double abc[3], double phi, double pq[4],
                      int edges[2], int *n_int);

/* TrackGenerator ctor (src/trackgenerator.jl:80-125); bcs = top,bottom,right,left (0 V, 1 R, 2 P) */
int orc_tg_create(orc_tg **out, const orc_mesh *m, int n_azim, double delta, const int32_t bcs[4],
                  double tiny_step);
void orc_tg_destroy(orc_tg *t);
int orc_tg_nazim2(const orc_tg *t);
int64_t orc_tg_n_total_tracks(const orc_tg *t);
void orc_tg_counts(const orc_tg *t, int64_t *n_tracks_x, int64_t *n_tracks_y, int64_t *n_tracks);

/* trace! (src/trackgenerator.jl:134-348) */
int orc_trace(orc_tg *t);
void orc_tg_quadrature(const orc_tg *t, double *phis, double *deltas, double *weights);
/* per-angle host tables exactly as the walk uses them: sin, cos, tan of phis, dx_eff, dy_eff */
void orc_tg_angle_tables(const orc_tg *t, double *sin_phi, double *cos_phi, double *tan_phi, double *dx_eff,
                         double *dy_eff);
/* SoA over uid (1..n_total) -> arrays indexed uid-1. Any pointer may be NULL. */
void orc_tg_tracks(const orc_tg *t, int64_t *azim_idx, int64_t *track_idx, double *p /*2n*/,
                   double *q /*2n*/, double *phi, double *len, double *abc /*3n*/, int8_t *bc_fwd,
                   int8_t *bc_bwd, int8_t *dir_fwd, int8_t *dir_bwd, int64_t *next_fwd, int64_t *next_bwd);

/* segmentize! (src/trackgenerator.jl:357-369) on uid range [uid_begin, uid_end) (1-based, end exclusive).
 * nthreads<=1: serial in uid order like the reference; >1: OpenMP over tracks (NOT reference behaviour).
 * Returns 0 or the status of the first bad uid (which is stored in *first_bad_uid). Keeps going after a
 * bad track so that per-track statuses can be compared. */
int orc_segmentize(orc_tg *t, int k, double rtol, int64_t uid_begin, int64_t uid_end, int nthreads,
                   int64_t *n_segments, int64_t *first_bad_uid);
/* counts[i] for uid_begin+i ; status likewise */
void orc_seg_counts(const orc_tg *t, int64_t uid_begin, int64_t uid_end, int64_t *counts, int32_t *status);
/* concatenated in uid order over the range */
void orc_seg_copy(const orc_tg *t, int64_t uid_begin, int64_t uid_end, double *px, double *py, double *qx,
                  double *qy, double *len, int32_t *element);
/* walk statistics accumulated by the last orc_segmentize: steps, knn_fallbacks(k=2 branch taken),
 * k_retries, same_element_resteps, boundary_start_steps, vertex_steps, n_int3, nn_ties (only counted after
 * orc_set_diag_ties(1): find_element queries whose two nearest nodes are at exactly the same distance) */
void orc_seg_stats(const orc_tg *t, int64_t stats[8]);
void orc_set_diag_ties(int on);
int orc_nn_is_tied(const orc_mesh *m, double x, double y); /* the two nearest nodes of (x, y) are at exactly the same distance */
/* fill_volumes (src/trackgenerator.jl:371-386) over the tracks segmentized so far, uid order */
void orc_volumes(const orc_tg *t, double *volumes);
void orc_seg_free(orc_tg *t);

#ifdef __cplusplus
}
#endif
#endif
