/* oracle/mcut_oracle.h — CPU restatement (plain C99) of the intersection-detection hot path of
 * cutdigital/mcut, written from SURVEY.md §8 and the reference's behaviour, NOT copied from it.
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; nothing under mcut_b200/ (the product) may call, link or execute
 * it.  Parity status: PINNED — every function below is checked bit-for-bit against the reference
 * itself (oracle/_ref, built from /root/reference by oracle/Makefile) through
 * tests/golden/make_golden.py; the resulting vectors are committed under tests/golden/.
 *
 * Each function cites the reference file:line it restates (paths relative to the reference root).
 * All arithmetic is IEEE-754 binary64, round-to-nearest, no FMA contraction
 * (built with -ffp-contract=off -frounding-math, like the reference's -frounding-math x86-64 build).
 */
#ifndef MCUT_ORACLE_H_
#define MCUT_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCO_NULL 0xFFFFFFFFu

/* status codes of the narrowphase, mirroring status_t of include/mcut/internal/kernel.h */
enum {
    MCO_SUCCESS = 0,
    MCO_GENERAL_POSITION_VIOLATION = 1,
    MCO_INVALID_SRC_MESH = 2, /* degenerate candidate face in the source mesh (kernel.cpp:2301-2312) */
    MCO_INVALID_CUT_MESH = 3
};

/* ---- a1: coordinate re-centring (source/preproc.cpp:2124-2290, :57-185) -------------------------- */
/* is_float != 0: vertices are float32 triples (MC_DISPATCH_VERTEX_ARRAY_FLOAT), else float64.       */
/* out: com[3], shift[3] (= pre_quantization_translation), src_bbox[6], cut_bbox[6] (min xyz, max   */
/* xyz; already translated by shift as the reference does at :2241-2248).                           */
void mco_vertex_parameters(int is_float, const void* src, uint32_t nsv, const void* cut, uint32_t ncv, double com[3],
    double shift[3], double src_bbox[6], double cut_bbox[6]);
/* x' = (x - com) + shift [+ perturbation]; float input subtracts/adds in float (preproc.cpp:124-134) */
void mco_transform_vertices(int is_float, const void* in, uint32_t nv, const double com[3], const double shift[3],
    const double* perturbation /* NULL or [3] */, double* out /* [nv*3] */);
/* eps used to enlarge cut-mesh face boxes: |cut bbox diag| * constant, or constant alone when       */
/* absolute (preproc.cpp:2518, :2667-2675)                                                          */
double mco_cut_bbox_eps(const double cut_bbox[6], double gp_constant, int absolute);

/* ---- a2/a3: face AABBs, mesh AABB, Morton codes (source/bvh.cpp:196-217, :242-433) ---------------- */
void mco_face_bboxes(const double* xyz, const uint32_t* face_off, const uint32_t* face_vtx, uint32_t nf, double eps,
    double* bboxes /* [nf*6] */, double root[6]);
void mco_face_bboxes_prior(const double* xyz, const uint32_t* face_off, const uint32_t* face_vtx, uint32_t nf, double eps,
    const double* prior /* [n_prior*6] or NULL */, uint32_t n_prior, double* bboxes, double root[6]);
uint32_t mco_morton3D(float x, float y, float z);
void mco_morton_codes(const double* bboxes, uint32_t nf, const double root[6], uint32_t* codes);

/* ---- a5/a6: OIBVH build + BFS dual traversal (source/bvh.cpp:71-193, :437-783) -------------------- */
int mco_oibvh_size(int t);
/* Builds both trees exactly like the reference (Morton sort, implicit complete tree, level refit) and  */
/* runs the BFS traversal.  Output: pairs (src_face << 32 | cut_face), sorted ascending.  Returns the   */
/* number of pairs; *pairs is malloc'ed (caller frees with mco_free).  n_tests receives the number of   */
/* node-pair overlap tests the BFS performed.                                                          */
size_t mco_oibvh_pairs(const double* src_bboxes, uint32_t nsf, const double* cut_bboxes, uint32_t ncf,
    uint64_t** pairs, uint64_t* n_tests);
/* all pairs with closed-interval AABB overlap (math.h:931-941) through a uniform grid: the              */
/* tree-independent definition of the same set, used to cross-check mco_oibvh_pairs.                     */
size_t mco_grid_pairs(const double* src_bboxes, uint32_t nsf, const double* cut_bboxes, uint32_t ncf,
    uint64_t** pairs);
void mco_free(void* p);

/* ---- polygon soup ids (source/kernel.cpp:1593-1732, source/hmesh.cpp:406-651,705-733) -------------- */
typedef struct mco_soup {
    uint32_t nv, nf, ne, nh;
    uint32_t src_nv, src_nf;
    double* xyz; /* [nv*3] src vertices then cut vertices */
    uint32_t* face_off; /* [nf+1] */
    uint32_t* face_vtx; /* [nh] vertices in ps.get_vertices_around_face order */
    uint32_t* face_edge; /* [nh] edge of halfedge slot i (halfedge face_vtx[i-1] -> face_vtx[i]) */
    uint32_t* edge_v; /* [ne*2] source(h0), target(h0) */
    uint32_t* edge_f; /* [ne*2] face(h0), face(h1) (MCO_NULL when border) */
} mco_soup_t;

/* returns 0 on success, -1 if a face would use an already-used halfedge (non-manifold / bad winding) */
int mco_soup_build(const double* src_xyz, uint32_t nsv, const uint32_t* src_off, const uint32_t* src_vtx, uint32_t nsf,
    const double* cut_xyz, uint32_t ncv, const uint32_t* cut_off, const uint32_t* cut_vtx, uint32_t ncf, mco_soup_t* out);
void mco_soup_free(mco_soup_t* s);

/* ---- a9..a14: per-face plane, exact predicates, plane point, point-in-polygon ---------------------- */
/* source/math.cpp:130-239; returns max component index; normal is zeroed and 0 returned when degenerate */
int mco_plane_coefficients(const double* verts /* [n*3] */, int n, double normal[3], double* d);
/* source/shewchuk.c:2367-2410 (+ orient3dadapt :1962-2365), :1695-1729 (+ orient2dadapt :1611-1693) */
double mco_orient3d(const double pa[3], const double pb[3], const double pc[3], const double pd[3]);
double mco_orient3d_stageA(const double pa[3], const double pb[3], const double pc[3], const double pd[3], int* certain);
double mco_orient2d(const double pa[2], const double pb[2], const double pc[2]);
/* source/math.cpp:391-427 (+ :289-389 for faces with more than three vertices)                         */
char mco_segment_plane_type(const double q[3], const double r[3], const double* verts, int n, const double normal[3],
    int max_comp, double* q_res, double* r_res);
/* source/math.cpp:249-287 */
char mco_segment_plane_intersection(double p[3], const double normal[3], double d, const double q[3], const double r[3]);
/* source/math.cpp:710-793: 2x3 projection matrix, row-major */
void mco_projection_matrix(const double normal[3], int max_comp, double P[6]);
/* source/math.cpp:851-902 -> :553-704 */
char mco_point_in_polygon(const double p[3], const double* verts, int n, const double normal[3], int max_comp);

/* ---- a7..a15: the narrowphase over a candidate-pair set (source/kernel.cpp:1779-3231) --------------- */
typedef struct mco_test {
    uint32_t edge, face; /* polygon-soup ids */
    char type; /* '0' '1' 'p' 'q' 'r'  (math.cpp:391-427) */
    char pip; /* 'i' 'o' 'e' 'v', or 0 when no point-in-polygon test ran; for 'p' the first decisive one */
    int8_t sign_q, sign_r; /* sign of the two orient3d results */
    uint8_t exact_q, exact_r; /* 1 when the stage-A filter failed and orient3dadapt ran */
    uint8_t pad[2];
    double point[3]; /* segment/plane point (type '1' only) */
} mco_test_t;

typedef struct mco_record {
    uint32_t edge, face;
    double point[3];
} mco_record_t;

typedef struct mco_narrow_out {
    int status;
    uint32_t bad_face; /* face that made the mesh invalid (status 2/3) */
    size_t n_tests, n_records, n_edge_face_before_cull, n_cand_faces;
    mco_test_t* tests; /* sorted by (edge, face); all edge/face tests that survive the AABB cull */
    mco_record_t* records; /* sorted by (edge, face); the intersection registry */
    /* plane data of every candidate face (keys of the reference's map), sorted by face id */
    uint32_t* cand_faces;
    double* cand_normal; /* [n*3] */
    double* cand_d;
    int32_t* cand_maxcomp;
} mco_narrow_out_t;

/* pairs: (src_face << 32 | cut_face) with cut_face the cut-mesh-local id.  src_bboxes / cut_bboxes are  */
/* the face AABBs produced by the BVH build (cut ones enlarged) — kernel.cpp:2086-2105.                  */
/* stop_on_gp != 0 reproduces the reference: on the first general-position violation the narrowphase     */
/* returns status 1 with no records.  With 0 it keeps going (the offending tests are still logged).      */
int mco_narrowphase(const mco_soup_t* ps, const uint64_t* pairs, size_t npairs, const double* src_bboxes,
    const double* cut_bboxes, int stop_on_gp, mco_narrow_out_t* out);
void mco_narrow_free(mco_narrow_out_t* o);


/* ---- input validation passes ------------------------------------------------------------------------ */
/* find_connected_components (source/kernel.cpp:235-364): components of the VERTEX graph (edges = face edges), ids in the
 * order a scan over the vertices discovers them (so a component's id is the rank of its smallest vertex; a vertex that no
 * face uses is a component of its own).  fccmap[f] = component of the face's vertices; returns the number of components.
 * cc_vertex_count / cc_face_count need room for nv entries. */
int mco_connected_components(uint32_t nv, const uint32_t* face_off, const uint32_t* face_vtx, uint32_t nf, int32_t* fccmap,
    int32_t* cc_vertex_count, int32_t* cc_face_count);
/* mesh_is_closed (source/preproc.cpp:1957-1990): every halfedge has a face, i.e. no edge is used by one face only.
 * Returns the number of border edges (0 = watertight). */
uint32_t mco_border_edges(uint32_t nv, const uint32_t* face_off, const uint32_t* face_vtx, uint32_t nf);

/* ---- winding-number inside/outside query --------------------------------------------------------------- */
/* calculate_signed_solid_angle, triangle (source/preproc.cpp:1650-1698) and quad (:1700-1810) forms, already divided by
 * 2*pi as in the reference; mco_winding_number sums them over the faces in face order (getWindingNumber, :1907-1955; faces
 * with more than 4 vertices go through the reference's CDT, which is out of scope: returns NAN for such a mesh). */
double mco_solid_angle_tri(const double a[3], const double b[3], const double c[3], const double q[3]);
double mco_solid_angle_quad(const double a[3], const double b[3], const double c[3], const double d[3], const double q[3]);
double mco_winding_number(const double* xyz, const uint32_t* face_off, const uint32_t* face_vtx, uint32_t nf, const double q[3]);

/* ---- cut-path segment table (the first consumer of the registry) -------------------------------------------- */
/* What "Create edges with intersection points" (source/kernel.cpp:3332-3617) works from: cutpath_edge_creation_info, the
 * registry's intersection points grouped by {source-mesh face, cut-mesh face} (filled at kernel.cpp:2601-2655: the tested
 * face against each face incident to the tested edge; a std::map, so groups come in ascending (sm, cm) order and a group's
 * points in registry order), with groups of more than two points put in order along their common line by
 * linear_projection_sort (:1496-1531: projection of origin - p on normalize(origin - second point), std::sort ascending).
 * A group of two points is one cut-path edge; a group of one point is the reference's late general-position violation
 * (:3366-3440).  rec[i] is registry entry i (m0 vertex ps_vtx_cnt + i). */
typedef struct mco_cutpath {
    size_t n_groups, n_entries, n_single;
    uint64_t* key; /* [n_groups] sm face << 32 | cm face (polygon-soup ids), ascending */
    uint32_t* off; /* [n_groups + 1] */
    uint32_t* vtx; /* [n_entries] registry indices */
} mco_cutpath_t;
int mco_cutpath_segments(const uint32_t* edge_f /* [ne*2] */, uint32_t src_nf, const mco_record_t* rec, size_t n, mco_cutpath_t* out);
void mco_cutpath_free(mco_cutpath_t* o);

#ifdef __cplusplus
}
#endif
#endif /* MCUT_ORACLE_H_ */
