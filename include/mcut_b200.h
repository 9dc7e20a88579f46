/* include/mcut_b200.h — C-ABI of the B200-native intersection-detection stage for MCUT.
 *
 * This is the drop-in boundary: plain C, pointers and sizes only, no C++/torch types, no exceptions.
 * Everything behind it is hand-written CUDA for sm_100a (mcut_b200/csrc/).  There is NO CPU fallback:
 * every entry point that computes returns MCB200_ERR_NO_DEVICE / a CUDA error when no B200 is usable.
 *
 * What each group replaces in the reference (cutdigital/mcut; paths relative to the reference root):
 *
 *   mcb200_vertex_parameters / mcb200_cut_bbox_eps      source/preproc.cpp:2124-2290, :2518, :2667-2675
 *        (host, sequential on purpose: the centre of mass is an order-dependent sum — SURVEY §7 hard part 2)
 *   mcb200_soup_ids                                     source/kernel.cpp:1593-1732 + hmesh.cpp:406-651,705-733
 *        (host: polygon-soup face/edge numbering rules; input to the narrowphase, like `ps` is in the reference)
 *   mcb200_mesh_create / mcb200_mesh_set_frame          source/preproc.cpp:91-185 (x' = (x - com) + shift [+ perturbation])
 *   mcb200_bvh_build                                    build_oibvh(), include/mcut/internal/bvh.h:117-125,
 *                                                       source/bvh.cpp:219-636   (face AABBs, root AABB, Morton, sort, tree, refit)
 *   mcb200_bvh_intersect                                intersectOIBVHs(), bvh.h:127-133, source/bvh.cpp:638-783
 *   mcb200_narrowphase                                  dispatch(), source/kernel.cpp:1779-3231 (edge/face tests,
 *                                                       intersection points, registry), arithmetic of source/math.cpp
 *                                                       :130-287,:391-427,:553-902 and source/shewchuk.c:1611-2410
 *   mcb200_intersect_stage                              the three of them back to back without host round trips
 *   mcb200_intersect_stage_host                         the same from HOST arrays in one call: uploads pipelined with the builds,
 *                                                       polygon soup numbered on the device (mcb200_soup_number)
 *   mcb200_mesh_validate / _read_components             find_connected_components(), source/kernel.cpp:235-364;
 *                                                       mesh_is_closed(), source/preproc.cpp:1957-1990        
 *   mcb200_mesh_winding_number                          getWindingNumber(), source/preproc.cpp:1650-1955       
 *
 * Conventions: every function returns 0 on success, a negative MCB200_ERR_* or a positive cudaError_t otherwise;
 * mcb200_last_error() gives the text.  Host arrays are borrowed for the duration of the call only.  One
 * context = one device + one stream; a context may be used by one host thread at a time, different contexts
 * are independent (the MultipleContextsInParallel pattern maps one MCUT context to one mcb200_ctx).
 * Ids: source faces keep their index, cut faces are reported cut-mesh-local in pairs and as Fs + id
 * ("polygon-soup id", kernel.cpp:153-156) in narrowphase records; same for vertices and edges.
 */
#ifndef MCUT_B200_H_
#define MCUT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCB200_NULL 0xFFFFFFFFu

enum {
    MCB200_OK = 0,
    MCB200_ERR_NO_DEVICE = -1, /* no CUDA device / not an sm_100 part: the product never falls back to the CPU */
    MCB200_ERR_INVALID = -2, /* bad argument (NULL, zero faces, face smaller than a triangle, ...) */
    MCB200_ERR_NON_MANIFOLD = -3, /* mcb200_soup_ids: an edge is used twice in the same direction (hmesh.cpp:612-628) */
    MCB200_ERR_CAPACITY = -4, /* an output array or a device buffer was too small; device capacities are raised for a rerun */
    MCB200_ERR_INTERNAL = -5
};

/* narrowphase status, the values of status_t in include/mcut/internal/kernel.h that this stage can produce */
enum {
    MCB200_STATUS_SUCCESS = 0,
    MCB200_STATUS_GENERAL_POSITION_VIOLATION = 1,
    MCB200_STATUS_INVALID_SRC_MESH = 2,
    MCB200_STATUS_INVALID_CUT_MESH = 3
};

typedef struct mcb200_ctx mcb200_ctx;
typedef struct mcb200_mesh mcb200_mesh; /* device-resident mesh (+ its face AABBs and LBVH once built) */
typedef struct mcb200_soup mcb200_soup; /* device-resident polygon-soup topology of a (source, cut) pair */
typedef struct mcb200_result mcb200_result; /* device-resident outputs of traversal + narrowphase */

/* The order in which the reference registers intersection points (= how it numbers the intersection vertices of m0):
 * the iteration order of its hash map ps_edge_face_intersection_pairs (an unordered_map keyed by edge) (source/kernel.cpp:1779-1852) walked
 * in parallel_for blocks (include/mcut/internal/tpool.h:354-472; kernel.cpp:2415-2868).  cand_faces = the polygon-soup ids of
 * all faces that have a candidate partner, ASCENDING (the keys of ps_face_to_potentially_intersecting_others); face_off /
 * face_edge as in mcb200_soup_ids (face_off == NULL: triangles); helper_threads = the dispatch's thread-pool size.
 * rank[e] (e < ne) receives the position of edge e in that order, MCB200_NULL for edges of no candidate face.  Registry
 * order = records sorted by (rank[edge], face). */
int mcb200_reference_edge_rank(uint32_t n_cand_faces, const uint32_t* cand_faces, const uint32_t* face_off,
    const uint32_t* face_edge, uint32_t ne, uint32_t helper_threads, uint32_t* rank);
/* The same order as a list, from compact arrays: slot_edge[slot_off[i] .. slot_off[i+1]) = the edges around the i-th candidate
 * face (slot_off == NULL: three each).  order[0 .. *n_order) receives the distinct edges in registration order (capacity:
 * the number of slots).  This is what the adapter calls inside a live dispatch: nothing here is sized by the mesh. */
int mcb200_reference_edge_order(uint32_t n_cand_faces, const uint32_t* slot_off, const uint32_t* slot_edge,
    uint32_t helper_threads, uint32_t* order, uint32_t* n_order);

/* ---------------------------------------------------------------- context ---------------------------------- */
int mcb200_device_count(void);
/* stream == NULL: the context creates its own non-blocking stream.  Otherwise `stream` is a cudaStream_t the
 * caller owns (e.g. torch.cuda.current_stream().cuda_stream) and all work is enqueued on it. */
int mcb200_ctx_create(int device, void* stream, mcb200_ctx** ctx);
void mcb200_ctx_destroy(mcb200_ctx* ctx);
const char* mcb200_last_error(const mcb200_ctx* ctx); /* ctx may be NULL: last error of ctx creation */
int mcb200_ctx_sync(mcb200_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t mcb200_ctx_launch_count(const mcb200_ctx* ctx);
/* Per-kernel device timing for bench.py's roofline figures: while on, every launch is bracketed by a CUDA event
 * pair on the context's stream.  mcb200_ctx_profile_read synchronises, writes one line per kernel name
 * ("<name> <launches> <total_ms>\n") into buf, resets the log and returns the number of bytes written
 * (negative error code on failure).  Off by default; never on inside a timed region that is reported as `value`. */
int mcb200_ctx_set_profiling(mcb200_ctx* ctx, int on);
int mcb200_ctx_profile_read(mcb200_ctx* ctx, char* buf, size_t capacity);

/* ---------------------------------------------------------------- host-side logic (no GPU) ----------------- */
void mcb200_vertex_parameters(int is_float, const void* src_xyz, uint32_t nsv, const void* cut_xyz, uint32_t ncv,
    double com[3], double shift[3], double src_bbox[6], double cut_bbox[6]);
/* the same in two steps: stats[9] = min xyz, max xyz, mean xyz of one mesh (the sequential pass), then the combination;
 * lets a caller that cuts one source mesh many times (planar sections) scan it once */
void mcb200_vertex_stats(int is_float, const void* xyz, uint32_t nv, double stats[9]);
void mcb200_vertex_parameters_from_stats(const double src_stats[9], const double cut_stats[9], double com[3], double shift[3],
    double src_bbox[6], double cut_bbox[6]);
double mcb200_cut_bbox_eps(const double cut_bbox[6], double gp_constant, int absolute);
/* Polygon-soup ids.  *_off are face offsets ([nf+1]); *_vtx the user's vertex lists.  Outputs (caller-allocated):
 *   face_vtx[nh], face_edge[nh]  (nh = src_off[nsf] + cut_off[ncf]) in ps.get_vertices_around_face order,
 *   edge_v[2*nh], edge_f[2*nh]   (first *ne rows valid): source(h0), target(h0) / face(h0), face(h1)|MCB200_NULL */
int mcb200_soup_ids(uint32_t nsv, const uint32_t* src_off, const uint32_t* src_vtx, uint32_t nsf, const uint32_t* cut_off,
    const uint32_t* cut_vtx, uint32_t ncf, uint32_t* face_vtx, uint32_t* face_edge, uint32_t* edge_v, uint32_t* edge_f,
    uint32_t* ne);

/* ---------------------------------------------------------------- meshes ----------------------------------- */
/* Uploads the user's arrays as they are (float or double vertices; face_sizes == NULL means triangles,
 * preproc.cpp:206).  No geometry is computed here. */
int mcb200_mesh_create(mcb200_ctx* ctx, int is_float, const void* xyz, uint32_t nv, const uint32_t* face_vtx,
    const uint32_t* face_sizes, uint32_t nf, mcb200_mesh** mesh);
/* Same upload for a caller that has validated its indices already (the adapter inside a live mcDispatch, where
 * client_input_arrays_to_hmesh has range-checked every index, preproc.cpp:271 / :417): no index check and no host copy of the
 * face array.  Such a mesh cannot be handed to mcb200_soup_create, which reads that copy; number the soup on the device
 * (mcb200_intersect_stage_host) or from the caller's own tables. */
int mcb200_mesh_create_trusted(mcb200_ctx* ctx, int is_float, const void* xyz, uint32_t nv, const uint32_t* face_vtx,
    const uint32_t* face_sizes, uint32_t nf, mcb200_mesh** mesh);
/* Same, from arrays that already live on this context's device (no copy is made of xyz/face_vtx; face_off may be
 * NULL for triangles).  Used when inputs are resident in HBM. */
int mcb200_mesh_adopt_device(mcb200_ctx* ctx, int is_float, const void* d_xyz, uint32_t nv, const uint32_t* d_face_vtx,
    const uint32_t* d_face_off, uint32_t nf, uint32_t nh, mcb200_mesh** mesh);
/* Frame of the internal coordinates: x' = (x - com) + shift (+ perturbation).  com == NULL: the vertices already
 * are internal coordinates (double only).  The transform is applied on the fly by every kernel that reads a
 * vertex (bit-identical to materialising it, and a perturbation retry costs nothing). */
int mcb200_mesh_set_frame(mcb200_ctx* ctx, mcb200_mesh* mesh, const double com[3], const double shift[3],
    const double perturbation[3] /* or NULL */);
/* Replace the coordinates of a mesh in place (same vertex count and type); faces, face boxes and the BVH are kept.
 * This is the perturbation retry of the reference (preproc.cpp:2560-2700): the cut mesh gets new coordinates, the cut
 * BVH of the first pass stays (its boxes were enlarged to cover every perturbation, bvh.cpp:242-314). */
int mcb200_mesh_update_xyz(mcb200_ctx* ctx, mcb200_mesh* mesh, const void* xyz, uint32_t nv);
void mcb200_mesh_free(mcb200_ctx* ctx, mcb200_mesh* mesh);

/* ---------------------------------------------------------------- (1) LBVH build --------------------------- */
/* Face AABBs (enlarged by eps when eps > 0), mesh AABB, 30-bit Morton codes of the reference's formula, one-sweep
 * radix sort, Karras tree, atomic bottom-up refit.  Asynchronous on the context's stream. */
int mcb200_bvh_build(mcb200_ctx* ctx, mcb200_mesh* mesh, double eps);
/* build_oibvh()'s `face_bboxes` argument is in/out: the function resizes the vector and EXPANDS the boxes it finds there
 * (bvh.cpp:242-272), and preproc.cpp never clears it between the builds of one mcDispatch (:2453-2461, :2733-2760).  So
 * on the rebuild that follows a floating-polygon repartition, faces [0, n) start from the box they had (already enlarged
 * by eps once) and are enlarged again.  Hand the caller's incoming boxes over with this call and the NEXT build of `mesh`
 * reproduces that: box = union(prior box, box of the face's vertices), then enlarged by eps.  n = 0 clears. */
int mcb200_mesh_set_prior_face_boxes(mcb200_ctx* ctx, mcb200_mesh* mesh, const double* boxes /* [n*6] min xyz, max xyz */,
    uint32_t n);
/* D2H of what build_oibvh() hands back to its caller: face_bboxes [nf*6] (min xyz, max xyz; may be NULL) and the
 * mesh AABB (bvhAABBs[0]).  Synchronises the stream. */
int mcb200_bvh_read(mcb200_ctx* ctx, const mcb200_mesh* mesh, double* face_bboxes, double root_bbox[6]);
/* debugging / parity: Morton codes [nf] by face id and the sorted leaf order [nf] */
int mcb200_bvh_read_morton(mcb200_ctx* ctx, const mcb200_mesh* mesh, uint32_t* codes_by_face, uint32_t* sorted_faces);

/* ---------------------------------------------------------------- results ---------------------------------- */
int mcb200_result_create(mcb200_ctx* ctx, mcb200_result** res);
void mcb200_result_free(mcb200_ctx* ctx, mcb200_result* res);

typedef struct mcb200_counts {
    uint64_t n_pairs; /* candidate face pairs (closed-interval AABB overlap) */
    uint64_t n_node_tests; /* AABB tests the traversal performed (tree nodes + leaf boxes) */
    uint64_t n_tests; /* edge/face tests that survived ownership + AABB cull */
    uint64_t n_exact; /* of those, how many needed the exact-expansion orient3d stage */
    uint64_t n_records; /* intersection points registered */
    uint64_t n_cand_faces; /* faces that appear in at least one pair */
    int32_t status; /* MCB200_STATUS_* */
    uint32_t bad_face; /* polygon-soup id of a degenerate candidate face when status is INVALID_* */
} mcb200_counts;

/* ---------------------------------------------------------------- (2) traversal ---------------------------- */
/* Both meshes must have been built.  Leaves `n_pairs` pairs on the device, sorted ascending by
 * (src_face << 32 | cut_face).  Asynchronous. */
int mcb200_bvh_intersect(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_result* res);
/* Restrict the traversal to a slice of the query leaf range (multi-GPU sharding of one huge dispatch):
 * chunks of `chunk` consecutive Morton-ordered leaves are dealt round-robin, this call handles the chunks with
 * index % nparts == part.  nparts == 1 restores the full range. */
int mcb200_result_set_shard(mcb200_ctx* ctx, mcb200_result* res, uint32_t part, uint32_t nparts, uint32_t chunk);
/* Size the candidate-pair buffer up front (default: 4 x (Fs + Fc), at least 2^20).  The reference's own container is a
 * growing vector (bvh.cpp:638-720); here an overflowing run is reported by mcb200_result_counts with
 * MCB200_ERR_CAPACITY, which also raises the capacity to what the run needed so the caller simply runs the stage again. */
int mcb200_result_set_pair_capacity(mcb200_ctx* ctx, mcb200_result* res, uint64_t max_pairs);

/* ---------------------------------------------------------------- input validation --------- */
/* The two O(V+F) passes the reference runs on every input mesh before the kernel, on the device-resident mesh:
 * find_connected_components (kernel.cpp:235-364, via check_input_mesh, preproc.cpp:505-578: a mesh with more than one
 * component is rejected) and mesh_is_closed (preproc.cpp:1957-1990, feeds kernel_input.*_is_watertight).  Component ids
 * are the reference's: components numbered in the order of their smallest vertex, unused vertices count. */
typedef struct mcb200_validation {
    uint32_t n_components;
    uint32_t n_border_edges; /* edges used by one face only */
    int is_closed; /* n_border_edges == 0 */
} mcb200_validation;
int mcb200_mesh_validate(mcb200_ctx* ctx, mcb200_mesh* mesh, mcb200_validation* out);
/* After mcb200_mesh_validate: fccmap[nf] (component of every face), cc_vertex_count / cc_face_count [n_components].
 * Any output may be NULL; `capacity_components` bounds the two count arrays. */
int mcb200_mesh_read_components(mcb200_ctx* ctx, mcb200_mesh* mesh, int32_t* fccmap, int32_t* cc_vertex_count,
    int32_t* cc_face_count, size_t capacity_components);

/* Winding number of `query` with respect to the mesh: getWindingNumber(), preproc.cpp:1907-1955, the sum of
 * calculate_signed_solid_angle() over the faces (triangles :1650-1698, quads :1700-1810).  `query` is in the internal
 * coordinates the mesh's frame produces (for the reference's use: vertex 0 of the other mesh, transformed).  ~1 = inside,
 * ~0 = outside (check_and_store_input_mesh_intersection_type uses eps 1e-7, :1999-2122).  Agrees with the reference to
 * ~1e-13 (device atan2); MCB200_ERR_INVALID for a mesh with faces of more than four vertices (those need the reference's
 * CDT, which stays on the host). */
int mcb200_mesh_winding_number(mcb200_ctx* ctx, mcb200_mesh* mesh, const double query[3], double* winding_number);

/* The inside/outside verdict mcDispatch stores for MC_DISPATCH_INCLUDE_INTERSECTION_TYPE when the surfaces do NOT cut each
 * other (no candidate pair, preproc.cpp:2865-2901, or no connected component beyond the two inputs, :3693-3725):
 * check_and_store_input_mesh_intersection_type(), preproc.cpp:1999-2122, decision for decision — watertightness of both
 * meshes (mcb200_mesh_validate), closed-interval overlap of the two mesh AABBs (bvhAABBs[0]; both meshes must be built),
 * winding number of one mesh's first vertex with respect to the other (mcb200_mesh_winding_number, eps 1e-7).  Values are
 * those of McDispatchIntersectionType (mcut.h:486-492); STANDARD is the caller's answer when the surfaces do cut. */
#define MCB200_INTERSECTION_TYPE_STANDARD 0u
#define MCB200_INTERSECTION_TYPE_INSIDE_CUTMESH (1u << 1)
#define MCB200_INTERSECTION_TYPE_INSIDE_SOURCEMESH (1u << 2)
#define MCB200_INTERSECTION_TYPE_NONE (1u << 3)
int mcb200_intersection_type_without_cut(mcb200_ctx* ctx, mcb200_mesh* src, mcb200_mesh* cut, uint32_t* type);

/* ---------------------------------------------------------------- (3) narrowphase -------------------------- */
int mcb200_soup_create(mcb200_ctx* ctx, uint32_t nsf, uint32_t ncf, uint32_t nh, uint32_t ne, const uint32_t* face_vtx,
    const uint32_t* face_edge, const uint32_t* edge_f, mcb200_soup** soup);
/* Number the polygon soup of two meshes ON THE DEVICE (no host pass over the faces): same ids as mcb200_soup_ids /
 * the reference's `ps` (first-use edge ids in add_face order, hmesh.cpp:406-651; vertex lists as halfedge targets).  The
 * meshes' face arrays must be the ones the reference built `ps` from (user order).  `res` receives the error flag: a
 * non-manifold edge or inconsistent winding is reported by mcb200_result_counts as MCB200_ERR_NON_MANIFOLD. */
int mcb200_soup_number(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_result* res, mcb200_soup** soup);
/* Same with explicit face sizes (polygon soups; face_sizes == NULL: triangles).  face_sizes[f] for f in [0, nsf + ncf). */
int mcb200_soup_create_sized(mcb200_ctx* ctx, uint32_t nsf, uint32_t ncf, uint32_t nh, uint32_t ne, const uint32_t* face_vtx,
    const uint32_t* face_edge, const uint32_t* edge_f, const uint32_t* face_sizes, mcb200_soup** soup);
void mcb200_soup_free(mcb200_ctx* ctx, mcb200_soup* soup);
/* Builds the soup ids on the host from the two meshes' face arrays and uploads them (mcb200_soup_ids + create). */
int mcb200_soup_from_meshes(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_soup** soup);

#define MCB200_NARROW_LOG_TESTS 1u /* also keep one log entry per edge/face test (parity checks) */
/* Triangle meshes dismiss most tests before the ownership / box culls are even looked at (both endpoints on one certified
 * side of the tested plane: no output either way), and mcb200_counts.n_tests then counts only the tests that went
 * further.  With this flag the dismissed tests are put through the culls as well, so that n_tests is the number of
 * edge/face tests the reference runs (kernel.cpp:2483).  (MCB200_NARROW_LOG_TESTS implies it.) */
#define MCB200_NARROW_COUNT_TESTS 8u
/* mcb200_intersect_stage_host only: the caller vouches that the source (cut) mesh arrays are bit-for-bit the ones passed to
 * the previous call on this context, so their upload is skipped (C3: one 4M-triangle terrain against 256 planes; the BVH
 * is still rebuilt because the internal coordinates depend on both meshes through `com`).  The C API of the reference
 * only lends its arrays for the duration of a dispatch, so the shim never sets these. */
#define MCB200_STAGE_SRC_RESIDENT 2u
#define MCB200_STAGE_CUT_RESIDENT 4u
/* Consumes res's pairs; uses the face AABBs of the meshes' builds (cut ones enlarged) for the edge cull and the
 * meshes' CURRENT frames for coordinates (so a perturbed cut frame is tested against unperturbed boxes, exactly
 * like preproc.cpp:2562-2945).  Asynchronous. */
int mcb200_narrowphase(mcb200_ctx* ctx, const mcb200_soup* soup, const mcb200_mesh* src, const mcb200_mesh* cut,
    mcb200_result* res, uint32_t flags);

/* build(src) + build(cut) + intersect + narrowphase, nothing but kernel launches in between */
int mcb200_intersect_stage(mcb200_ctx* ctx, mcb200_mesh* src, mcb200_mesh* cut, double cut_eps, const mcb200_soup* soup,
    mcb200_result* res, uint32_t flags);

/* The same stage straight from HOST arrays, with the uploads pipelined against the kernels on a copy stream: the source
 * build starts as soon as the source mesh has landed, the cut build overlaps the polygon-soup upload, and the narrowphase
 * waits for the soup only.  Pin the arrays (cudaHostAlloc / torch pin_memory) to get the overlap; pageable memory works
 * but serialises.  The polygon-soup vertex lists are derived on the device from the face arrays (rotation rules of
 * hmesh.cpp:705-733), so only face_edge / edge_f travel.  Meshes and soup live in context-owned staging buffers that
 * are reused from call to call.  soup == NULL: the polygon soup is numbered on the device (same ids as mcb200_soup_ids:
 * an edge's id is the rank of its first halfedge in add_face order, hmesh.cpp:406-651), nothing but the two meshes
 * travels; a non-manifold edge or inconsistent winding is then reported by mcb200_result_counts as
 * MCB200_ERR_NON_MANIFOLD.
 * LIFETIME (the one exception to "borrowed for the duration of the call"): the call returns with the uploads still in
 * flight, so the host arrays must stay valid and unchanged until the next synchronising call on the context
 * (mcb200_result_counts, any mcb200_result_read_*, mcb200_ctx_sync) has returned. */
typedef struct mcb200_host_mesh {
    int is_float; /* MC_DISPATCH_VERTEX_ARRAY_FLOAT */
    const void* xyz; /* [nv*3] */
    uint32_t nv;
    const uint32_t* face_vtx;
    const uint32_t* face_sizes; /* NULL: triangles */
    uint32_t nf;
} mcb200_host_mesh;
typedef struct mcb200_host_soup {
    uint32_t nh, ne;
    const uint32_t* face_edge; /* [nh] */
    const uint32_t* edge_f; /* [ne*2] */
} mcb200_host_soup;
int mcb200_intersect_stage_host(mcb200_ctx* ctx, const mcb200_host_mesh* src, const mcb200_host_mesh* cut, const double com[3],
    const double shift[3], const double perturbation[3] /* or NULL */, double cut_eps, const mcb200_host_soup* soup,
    mcb200_result* res, uint32_t flags);
/* ---------------------------------------------------------------- one dispatch on several GPUs -------- */
/* One process per GPU, one context each; NCCL (loaded at run time, libnccl.so.2) carries the exchange over NVLink.
 * Rank 0 makes an id (mcb200_comm_unique_id), hands it to the other ranks by whatever means the application has (MPI,
 * torch.distributed, a file), every rank calls mcb200_comm_create.
 * mcb200_intersect_stage_sharded: every rank passes the SAME meshes and soup (replicated); rank r walks its slice of the
 * query leaf range (4096-leaf chunks of the Morton order dealt round-robin) and runs the narrowphase on its pairs; pairs
 * and registry records are then all-gathered (counts by ncclAllGather, payloads by grouped ncclBroadcast), merged and put in
 * canonical order on every rank: afterwards `res` on every rank holds the complete result, byte for byte what one GPU
 * produces.  Collective: all ranks must call it together.  MCB200_ERR_CAPACITY (on all ranks alike): run it again. */
#define MCB200_COMM_ID_BYTES 128
typedef struct mcb200_comm mcb200_comm;
int mcb200_comm_unique_id(char id[MCB200_COMM_ID_BYTES]);
int mcb200_comm_create(mcb200_ctx* ctx, int nranks, int rank, const char id[MCB200_COMM_ID_BYTES], mcb200_comm** comm);
void mcb200_comm_destroy(mcb200_comm* comm);
int mcb200_comm_rank(const mcb200_comm* comm);
int mcb200_comm_size(const mcb200_comm* comm);
int mcb200_intersect_stage_sharded(mcb200_ctx* ctx, mcb200_comm* comm, mcb200_mesh* src, mcb200_mesh* cut, double cut_eps,
    const mcb200_soup* soup, mcb200_result* res, uint32_t flags);

/* Many independent dispatches in one call — the MultipleContextsInParallel pattern (tutorials/MultipleContextsInParallel/
 * MultipleContextsInParallel.cpp:129-345: one MCUT context per task, tasks spread over threads) for dispatches that are too
 * small to fill the machine one at a time.  `nctx` contexts of ONE device (each with its own result object) serve as
 * lanes: item i is enqueued on lane i % nctx through mcb200_intersect_stage_host (for small meshes that is a handful of
 * copies plus one replayed CUDA graph), a lane's previous item is collected (counts read, capacity retries done) right
 * before the lane is reused, so up to nctx dispatches are in flight and ONE host thread feeds them.
 * com == NULL in an item: the frame (mcb200_vertex_parameters) and the cut eps (mcb200_cut_bbox_eps with gp_constant,
 * relative) are computed here, as the reference does per dispatch (preproc.cpp:2124-2290, :2518).  counts[i] receives the
 * counts and status of item i.  Host arrays of ALL items must stay valid until the call returns.  Returns 0, or the first
 * error (the remaining lanes are still drained). */
typedef struct mcb200_batch_item {
    mcb200_host_mesh src, cut;
    const double* com; /* [3] or NULL: compute frame and eps here */
    const double* shift; /* [3] */
    const double* perturbation; /* [3] or NULL */
    double cut_eps; /* used when com != NULL */
    double gp_constant; /* used when com == NULL (the reference's default is 1e-4) */
    uint32_t flags; /* MCB200_NARROW_* / MCB200_STAGE_* */
} mcb200_batch_item;
int mcb200_batch_intersect_host(mcb200_ctx** ctxs, mcb200_result** results, uint32_t nctx, const mcb200_batch_item* items, uint32_t n,
    mcb200_counts* counts);

/* The polygon-soup ids the last mcb200_intersect_stage_host call of this context worked with (uploaded or numbered on the
 * device): face_vtx[nh], face_edge[nh], edge_f[2*ne].  Any output may be NULL.  Synchronises. */
int mcb200_staged_soup_read(mcb200_ctx* ctx, uint32_t* face_vtx, uint32_t* face_edge, uint32_t* edge_f, uint32_t capacity_edges,
    uint32_t* nh, uint32_t* ne);

/* ---------------------------------------------------------------- cut-path segment table ------ */
/* The first consumer of the registry, "Create edges with intersection points" (source/kernel.cpp:3332-3617), works from
 * cutpath_edge_creation_info: the intersection points grouped by {source-mesh face, cut-mesh face} (kernel.cpp:2603-2633).
 * mcb200_cutpath_segments makes that table on the device from the last narrowphase's registry (canonical order: vertex i =
 * record i): groups in ascending (sm face, cm face) order like the reference's ordered map, a group's points in registry
 * order, groups of more than two points in order along their line (linear_projection_sort, kernel.cpp:1496-1531).  A
 * group of two points is one cut-path edge; n_single_point_groups > 0 is the reference's late general-position violation
 * (:3366-3440).  mcb200_cutpath_read: keys[g] = sm face << 32 | cm face (polygon-soup ids), offsets[n_groups + 1],
 * vertices[n_entries] = registry indices.  The m0 bookkeeping that follows stays host code. */
typedef struct mcb200_cutpath_counts {
    uint64_t n_groups, n_entries, n_single_point_groups;
} mcb200_cutpath_counts;
int mcb200_cutpath_segments(mcb200_ctx* ctx, const mcb200_soup* soup, mcb200_result* res, mcb200_cutpath_counts* out);
int mcb200_cutpath_read(mcb200_ctx* ctx, mcb200_result* res, uint64_t* keys, uint32_t* offsets, uint32_t* vertices, size_t cap_groups,
    size_t cap_entries);

/* Diagnostics: how many items each narrowphase kernel worked on in the last run.  out[0] pairs that went past the side
 * prefilter (triangle meshes; every pair otherwise), out[1] tests whose stage-A filter failed, out[2] certified plane
 * crossings, out[3] tests that needed Shewchuk's adaptive stages B-D (inexact coordinate differences). */
int mcb200_result_queue_counts(mcb200_ctx* ctx, mcb200_result* res, uint64_t out[4]);

/* ---------------------------------------------------------------- reading results (D2H, synchronising) ----- */
int mcb200_result_counts(mcb200_ctx* ctx, mcb200_result* res, mcb200_counts* out);
int mcb200_result_read_pairs(mcb200_ctx* ctx, mcb200_result* res, uint64_t* pairs, size_t capacity);

typedef struct mcb200_record {
    uint32_t edge, face; /* polygon-soup ids: the tested edge and the face it pierces */
    double point[3]; /* intersection point, the reference's q + t (r - q) */
} mcb200_record;
/* sorted by (edge, face): the canonical registry order */
int mcb200_result_read_records(mcb200_ctx* ctx, mcb200_result* res, mcb200_record* records, size_t capacity);

typedef struct mcb200_test {
    uint32_t edge, face;
    char type; /* '0' '1' 'p' 'q' 'r' (math.cpp:391-427) */
    char pip; /* 'i' 'o' 'e' 'v' or 0 */
    int8_t sign_q, sign_r; /* signs of the two orient3d determinants */
    uint8_t exact; /* bit0/bit1: q / r needed the exact stage */
    uint8_t pad[3];
    double point[3];
} mcb200_test;
/* only when MCB200_NARROW_LOG_TESTS was passed; sorted by (edge, face) */
int mcb200_result_read_tests(mcb200_ctx* ctx, mcb200_result* res, mcb200_test* tests, size_t capacity);
/* plane data of the candidate faces (keys of ps_face_to_potentially_intersecting_others), ascending face id:
 * faces[n], normal[n*3], d[n], max_comp[n] — kernel.cpp:2184-2356, consumed downstream by the host */
int mcb200_result_read_planes(mcb200_ctx* ctx, mcb200_result* res, uint32_t* faces, double* normal, double* d,
    int32_t* max_comp, size_t capacity);

/* device pointers of the result buffers (for collectives that gather them, e.g. NCCL all-gather in bench.py):
 * which: 0 pairs (u64), 1 records (mcb200_record).  *count is the element count (synchronises). */
int mcb200_result_device_ptr(mcb200_ctx* ctx, mcb200_result* res, int which, void** dptr, uint64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* MCUT_B200_H_ */
