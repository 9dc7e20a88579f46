// mcut_b200/csrc/narrowphase.cu — (3) exact edge/face narrowphase.
//
// Replaces the inline narrowphase of dispatch() (source/kernel.cpp:1779-3231):
//   "Prepare edge-to-face pairs"  :1781-1983   edge e is tested against the union of the candidate lists of its
//                                              incident faces.  Here: a candidate pair (f, g) generates the test
//                                              (e, g) for every halfedge slot e of f; the same (e, g) reached from the
//                                              face on the other side of e is dropped by an OWNERSHIP rule instead of a
//                                              sort/unique: the face of halfedge h0 owns the test, the h1 face runs it
//                                              only when (face(h0), g) is not a candidate pair — decidable locally
//                                              because the candidate set IS the closed AABB overlap set.
//   "Build edge bounding boxes" / "Cull"  :1989-2177   edge box vs the BVH-build box of the tested face
//   "Compute intersecting face properties" :2184-2356  k_planes (also the degenerate-face -> INVALID_*_MESH rule)
//   "Calculate intersection points"  :2415-3231
//        k_filter  : Shewchuk stage-A orient3d x2 (error-bound filter), plane point, point-in-polygon, registry
//                    record; tests whose filter fails are compacted into a queue ...
//        k_exact   : ... and re-evaluated with the exact expansion arithmetic (orient3dadapt), then the same tail.
// Registry order: records are sorted by (edge, face) — the canonical order of SURVEY §8-a15; the reference's own order
// depends on unordered_map iteration and on its thread count.
#include "internal.h"
#include "predicates.cuh"
#include "radix_sort.cuh"

namespace {

constexpr int NBLOCK = 128;

struct narrow_args_t {
    // geometry
    const void* src_xyz;
    const void* cut_xyz;
    const frame_t* src_frame; // device slots (d_frames[1] of each mesh); the kernels repoint them at a shared-memory copy
    const frame_t* cut_frame;
    uint32_t src_nv;
    const double* src_bbox; // [nsf][6]
    const double* cut_bbox; // [ncf][6] (enlarged, from the unperturbed build)
    // polygon-soup topology
    const uint32_t* face_off; // [nf+1]
    const uint32_t* face_vtx; // [nh] ps vertex ids
    const uint32_t* face_edge; // [nh]
    const uint32_t* edge_f; // [ne][2]
    uint32_t nsf, nf;
    // work
    const unsigned long long* pairs;
    unsigned long long cap_pairs;
    result_counters_t* counters;
    // outputs
    uint8_t* cand_flag;
    mcb200_record* records;
    unsigned long long cap_records;
    unsigned long long* exact_queue; // polygons: (pair index << 8 | slot); triangles: see k_tri_classify
    unsigned long long cap_exact;
    unsigned long long* mid_queue; // triangles: pairs the side prefilter could not dismiss; reused for what k_tri_resolve hands on
    unsigned long long cap_mid;
    // what the general exact kernel (k_tests<*, true>) reads: (pair index << 8 | slot) entries and their count
    const unsigned long long* exact_in;
    const unsigned long long* exact_in_n;
    unsigned long long exact_in_cap;
    uint32_t tri_mode; // 0: prefilter on; 1: and the dismissed tests are still counted (n_tests as the reference runs them); 2: prefilter off (test log)
    mcb200_test* tests; // optional log
    unsigned long long cap_tests;
};

__device__ __forceinline__ void load_ps_vertex(const narrow_args_t& a, uint32_t v, double* out)
{
    if (v < a.src_nv) load_vertex(a.src_xyz, *a.src_frame, v, out);
    else load_vertex(a.cut_xyz, *a.cut_frame, v - a.src_nv, out);
}

__device__ __forceinline__ void load_box(const double* p, double* b)
{
    const double2* in = reinterpret_cast<const double2*>(p);
    const double2 x = __ldg(in), y = __ldg(in + 1), z = __ldg(in + 2);
    b[0] = x.x;
    b[1] = x.y;
    b[2] = y.x;
    b[3] = y.y;
    b[4] = z.x;
    b[5] = z.y;
}

__device__ __forceinline__ const double* face_box_ptr(const narrow_args_t& a, uint32_t ps_face)
{
    return ps_face < a.nsf ? a.src_bbox + 6 * (size_t)ps_face : a.cut_bbox + 6 * (size_t)(ps_face - a.nsf);
}

// A face whose vertices are fetched on demand (any polygon size).
struct face_view {
    const narrow_args_t* a;
    uint32_t h0, n;
    __device__ __forceinline__ void vert(uint32_t i, double* out) const { load_ps_vertex(*a, __ldg(a->face_vtx + h0 + i), out); }
};

// warp-aggregated slot allocation for threads that happen to be in the same branch
__device__ __forceinline__ unsigned long long alloc_slot(unsigned long long* counter)
{
    const unsigned m = __activemask();
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if ((int)lane_id() == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & lanemask_lt());
}

struct test_out_t {
    char type, pip;
    int8_t sq, sr;
    uint8_t exact;
    double p[3];
};

__device__ __forceinline__ int sgn(double x) { return (x > 0.0) - (x < 0.0); }

// plane of face G (math.cpp:130-239); returns max_comp
template <bool TRI> __device__ __forceinline__ int face_plane(const face_view& G, const double (*gv)[3], double* normal, double& d)
{
    if (TRI) return pred::plane_tri(gv[0], gv[1], gv[2], normal, d);
    normal[0] = normal[1] = normal[2] = 0.0;
    double first[3], cur[3], nxt[3];
    G.vert(0, first);
    for (int k = 0; k < 3; ++k) cur[k] = first[k];
    for (uint32_t i = 0; i < G.n; ++i) {
        if (i + 1 < G.n) G.vert(i + 1, nxt);
        else
            for (int k = 0; k < 3; ++k) nxt[k] = first[k];
        pred::newell_step(normal, cur, nxt);
        for (int k = 0; k < 3; ++k) cur[k] = nxt[k];
    }
    return pred::plane_finish(normal, first, d);
}

// point in polygon (math.cpp:851-902 -> :553-704)
template <bool TRI>
__device__ __forceinline__ char face_pip(const face_view& G, const double (*gv)[3], const double* p, const double* normal, int mc)
{
    if (TRI) return pred::point_in_triangle(p, gv[0], gv[1], gv[2], normal, mc);
    double P[6], pp[2];
    pred::projection_matrix(normal, mc, P);
    pred::project2(P, p, pp);
    int rcross = 0, lcross = 0;
    double v3[3], prev[2], cur[2];
    G.vert(G.n - 1, v3);
    pred::project2(P, v3, prev);
    prev[0] = pred::sub(prev[0], pp[0]);
    prev[1] = pred::sub(prev[1], pp[1]);
    for (uint32_t i = 0; i < G.n; ++i) {
        G.vert(i, v3);
        pred::project2(P, v3, cur);
        cur[0] = pred::sub(cur[0], pp[0]);
        cur[1] = pred::sub(cur[1], pp[1]);
        if (cur[0] == 0.0 && cur[1] == 0.0) return 'v';
        const bool rstrad = (cur[1] > 0.0) != (prev[1] > 0.0);
        const bool lstrad = (cur[1] < 0.0) != (prev[1] < 0.0);
        if (rstrad || lstrad) {
            const double x = pred::sub(pred::mul(cur[0], prev[1]), pred::mul(prev[0], cur[1])) / pred::sub(prev[1], cur[1]);
            if (rstrad && x > 0.0) rcross++;
            if (lstrad && x < 0.0) lcross++;
        }
        prev[0] = cur[0];
        prev[1] = cur[1];
    }
    if ((rcross & 1) != (lcross & 1)) return 'e';
    return (rcross & 1) ? 'i' : 'o';
}

// math.cpp:289-389 for faces with more than three vertices: the triple (i<j<k) with the largest |orient2d| of the
// projected vertices; first maximal one in enumeration order (libstdc++ insertion sort is stable for <= 16 triples).
__device__ __noinline__ bool best_triple(const face_view& G, const double* normal, int mc, int* ijk)
{
    double P[6];
    pred::projection_matrix(normal, mc, P);
    double best = -1.0;
    bool found = false;
    double vi[3], vj[3], vk[3], xi[2], xj[2], xk[2];
    for (uint32_t i = 0; i < G.n; ++i) {
        G.vert(i, vi);
        pred::project2(P, vi, xi);
        for (uint32_t j = i + 1; j < G.n; ++j) {
            G.vert(j, vj);
            pred::project2(P, vj, xj);
            for (uint32_t k = j + 1; k < G.n; ++k) {
                G.vert(k, vk);
                pred::project2(P, vk, xk);
                const double r = pred::orient2d(xi, xj, xk);
                if (r == 0.0) continue;
                if (fabs(r) > best) {
                    best = fabs(r);
                    ijk[0] = (int)i;
                    ijk[1] = (int)j;
                    ijk[2] = (int)k;
                    found = true;
                }
            }
        }
    }
    return found;
}

__device__ __forceinline__ void classify_signs(int sq, int sr, test_out_t& o);
template <bool TRI>
__device__ __forceinline__ void finish_touching(const face_view& G, const double (*gv)[3], const double* q, const double* r,
    test_out_t& o, unsigned& gp_violation, bool have_plane, double* normal, double d, int mc);

// One edge/face test (kernel.cpp:2483-2656).  With EXACT off it returns false unless the test is a CERTIFIED non-crossing
// (both stage-A determinants pass the error-bound filter and have the same sign): the caller queues everything else for
// the second kernel (`needs_exact` tells it whether a stage-A filter failed); with EXACT on it always fills `o`.
template <bool TRI, bool EXACT>
__device__ __forceinline__ bool eval_test(const face_view& G, const double (*gv)[3], const double* q, const double* r,
    test_out_t& o, unsigned& gp_violation, bool& needs_exact)
{
    double normal[3], d = 0.0;
    int mc = 0;
    bool have_plane = false;
    double A[3], B[3], C[3];
    if (TRI) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            A[k] = gv[0][k];
            B[k] = gv[1][k];
            C[k] = gv[2][k];
        }
    } else {
        int ijk[3] = { 0, 1, 2 };
        if (G.n > 3) {
            mc = face_plane<false>(G, gv, normal, d);
            have_plane = true;
            if (!best_triple(G, normal, mc, ijk)) { // all vertices collinear (math.cpp:408-410)
                o.type = '0';
                o.pip = 0;
                o.sq = o.sr = 0;
                o.exact = 0;
                o.p[0] = o.p[1] = o.p[2] = 0.0;
                return true;
            }
        }
        G.vert((uint32_t)ijk[0], A);
        G.vert((uint32_t)ijk[1], B);
        G.vert((uint32_t)ijk[2], C);
    }
    bool cq, cr;
    double permq, permr;
    double detq = pred::orient3d_stageA(A, B, C, q, cq, permq);
    double detr = pred::orient3d_stageA(A, B, C, r, cr, permr);
    o.exact = (uint8_t)((cq ? 0 : 1) | (cr ? 0 : 2));
    if (!cq || !cr) {
        if (!EXACT) {
            needs_exact = true;
            return false;
        }
        if (!cq) detq = pred::orient3d_adapt(A, B, C, q, permq);
        if (!cr) detr = pred::orient3d_adapt(A, B, C, r, permr);
    }
    classify_signs(sgn(detq), sgn(detr), o);
    if (o.type == '0') return true;
    // The filter kernel settles the certified non-crossings only (98 % of the tests of a dense overlap): an edge that does
    // cross the plane — plane point, point-in-polygon, registry record — is left to the second kernel, so that this
    // kernel's register budget is the two stage-A determinants and nothing else.
    if (!EXACT) return false;
    finish_touching<TRI>(G, gv, q, r, o, gp_violation, have_plane, normal, d, mc);
    return true;
}

// segment vs plane from the two orientation signs (math.cpp:416-426)
__device__ __forceinline__ void classify_signs(int sq, int sr, test_out_t& o)
{
    o.sq = (int8_t)sq;
    o.sr = (int8_t)sr;
    o.pip = 0;
    o.p[0] = o.p[1] = o.p[2] = 0.0;
    if (sq == 0 && sr == 0) o.type = 'p';
    else if (sq == 0) o.type = 'q';
    else if (sr == 0) o.type = 'r';
    else if (sq == sr) o.type = '0';
    else o.type = '1';
}

// everything after the signs for a test that crosses or touches the plane (kernel.cpp:2518-2597)
template <bool TRI>
__device__ __forceinline__ void finish_touching(const face_view& G, const double (*gv)[3], const double* q, const double* r,
    test_out_t& o, unsigned& gp_violation, bool have_plane, double* normal, double d, int mc)
{
    if (!have_plane) mc = face_plane<TRI>(G, gv, normal, d);
    if (o.type == '1') {
        pred::segment_plane_point(o.p, normal, d, q, r); // kernel.cpp:2559-2564
        o.pip = face_pip<TRI>(G, gv, o.p, normal, mc); // :2566-2575
        if (o.pip == 'v' || o.pip == 'e') gp_violation = 1u; // :2588-2597
    } else { // 'p' 'q' 'r': an endpoint touches the plane (kernel.cpp:2518-2557)
        const bool test_q = (o.type != 'r'), test_r = (o.type != 'q');
        bool stop = false;
        if (test_q) {
            o.pip = face_pip<TRI>(G, gv, q, normal, mc);
            stop = (o.pip == 'i' || o.pip == 'v' || o.pip == 'e');
        }
        if (!stop && test_r) {
            o.pip = face_pip<TRI>(G, gv, r, normal, mc);
            stop = (o.pip == 'i' || o.pip == 'v' || o.pip == 'e');
        }
        if (stop) gp_violation = 1u;
    }
}

__device__ __forceinline__ void emit(const narrow_args_t& a, uint32_t edge, uint32_t face, const test_out_t& o)
{
    if (o.type == '1' && o.pip == 'i') {
        const unsigned long long slot = alloc_slot(&a.counters->n_records);
        if (slot < a.cap_records) {
            mcb200_record rec;
            rec.edge = edge;
            rec.face = face;
            rec.point[0] = o.p[0];
            rec.point[1] = o.p[1];
            rec.point[2] = o.p[2];
            a.records[slot] = rec;
        }
    }
}

__device__ __forceinline__ void log_test(const narrow_args_t& a, uint32_t edge, uint32_t face, const test_out_t& o)
{
    if (!a.tests) return;
    const unsigned long long slot = alloc_slot(&a.counters->n_log);
    if (slot < a.cap_tests) {
        mcb200_test t;
        t.edge = edge;
        t.face = face;
        t.type = o.type;
        t.pip = o.pip;
        t.sign_q = o.sq;
        t.sign_r = o.sr;
        t.exact = o.exact;
        t.pad[0] = t.pad[1] = t.pad[2] = 0;
        t.point[0] = o.p[0];
        t.point[1] = o.p[1];
        t.point[2] = o.p[2];
        a.tests[slot] = t;
    }
}

// Decode halfedge slot `slot` of pair (s, c): slots [0, ns) are the halfedges of the source face tested against the
// cut face, slots [ns, ns+nc) those of the cut face tested against the source face.  Applies ownership + AABB cull.
// Returns false when the test must not run.  On success q, r are source(h0), target(h0) (kernel.cpp:2466-2474).
template <bool TRI>
__device__ __forceinline__ bool setup_test(const narrow_args_t& a, uint32_t s, uint32_t c, uint32_t slot, uint32_t hs,
    uint32_t ns, uint32_t hc, uint32_t nc, const double (*sv)[3], const double (*cv)[3], const double* sbox,
    const double* cbox, uint32_t& edge, uint32_t& tested_face, bool& edge_from_src, double* q, double* r,
    const uint32_t* pre_edge = nullptr, const uint2* pre_ef = nullptr, bool* h0_out = nullptr)
{
    edge_from_src = slot < ns;
    const uint32_t i = edge_from_src ? slot : slot - ns;
    const uint32_t n = edge_from_src ? ns : nc;
    const uint32_t hbase = edge_from_src ? hs : hc;
    const uint32_t own_face = edge_from_src ? s : a.nsf + c;
    tested_face = edge_from_src ? a.nsf + c : s;
    // the triangle filter fetches the six edge ids and their face pairs up front (two round trips instead of twelve)
    edge = pre_edge ? pre_edge[slot] : __ldg(a.face_edge + hbase + i);
    const uint2 ef = pre_ef ? pre_ef[slot] : __ldg(reinterpret_cast<const uint2*>(a.edge_f) + edge);
    const bool is_h0 = (ef.x == own_face);
    if (h0_out) *h0_out = is_h0;
    const double* tbox = edge_from_src ? cbox : sbox; // box of the tested face
    if (!is_h0 && ef.x != MCB200_NULL) { // (an edge whose h0 has no face — a border after a repartition — is owned by the h1 face)
        // the face of h0 owns this test whenever it is paired with the tested face too
        double ob[6];
        load_box(face_box_ptr(a, ef.x), ob);
        if (overlap6(ob, tbox)) return false;
    }
    // halfedge i of a face runs from its vertex i-1 to its vertex i (hmesh.cpp:705-733: vertices are halfedge targets)
    const uint32_t ip = (i + n - 1) % n;
    double from[3], to[3];
    if (TRI) {
        const double(*fv)[3] = edge_from_src ? sv : cv;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            from[k] = fv[ip][k];
            to[k] = fv[i][k];
        }
    } else {
        load_ps_vertex(a, __ldg(a.face_vtx + hbase + ip), from);
        load_ps_vertex(a, __ldg(a.face_vtx + hbase + i), to);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        q[k] = is_h0 ? from[k] : to[k];
        r[k] = is_h0 ? to[k] : from[k];
    }
    // edge box vs the tested face's box (kernel.cpp:2004-2027, :2086-2115)
    double eb[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        eb[k] = ref_min(q[k], r[k]);
        eb[3 + k] = ref_max(q[k], r[k]);
    }
    return overlap6(eb, tbox);
}

// Work items: the exact kernel takes queue entries (pair, slot); the filter takes pairs and walks the slots of its pair
// in a loop (one thread per (pair, slot) was measured slower: 2.4x the instructions, the per-pair loads dominate).
template <bool TRI, bool EXACT> __global__ void __launch_bounds__(NBLOCK) k_tests(narrow_args_t a_in)
{
    pdl_prologue();
    __shared__ frame_t s_fr[2];
    load_frame_shared(&s_fr[0], a_in.src_frame);
    if (threadIdx.x >= 64) { // second half of the block fetches the cut frame
        constexpr unsigned W = sizeof(frame_t) / 4;
        const unsigned t = threadIdx.x - 64u;
        if (t < W) reinterpret_cast<unsigned*>(&s_fr[1])[t] = __ldg(reinterpret_cast<const unsigned*>(a_in.cut_frame) + t);
    }
    __syncthreads();
    narrow_args_t a = a_in;
    a.src_frame = &s_fr[0];
    a.cut_frame = &s_fr[1];
    unsigned long long n_items;
    if (EXACT) {
        n_items = *a.exact_in_n < a.exact_in_cap ? *a.exact_in_n : a.exact_in_cap;
    } else {
        n_items = a.counters->n_pairs < a.cap_pairs ? a.counters->n_pairs : a.cap_pairs;
    }
    unsigned n_tests_local = 0, n_exact_local = 0, gp = 0;
    for (unsigned long long it = (unsigned long long)blockIdx.x * NBLOCK + threadIdx.x; it < n_items;
         it += (unsigned long long)gridDim.x * NBLOCK) {
        unsigned long long pair_index = it;
        uint32_t only_slot = 0xFFFFFFFFu;
        if (EXACT) {
            const unsigned long long e = a.exact_in[it];
            pair_index = e >> 8;
            only_slot = (uint32_t)(e & 0xFFu);
        }
        const unsigned long long pr = a.pairs[pair_index];
        const uint32_t s = (uint32_t)(pr >> 32), c = (uint32_t)(pr & 0xFFFFFFFFu);
        const uint32_t hs = TRI ? 3u * s : __ldg(a.face_off + s);
        const uint32_t ns = TRI ? 3u : __ldg(a.face_off + s + 1) - hs;
        const uint32_t hc = TRI ? 3u * (a.nsf + c) : __ldg(a.face_off + a.nsf + c);
        const uint32_t nc = TRI ? 3u : __ldg(a.face_off + a.nsf + c + 1) - hc;
        if (!EXACT) {
            a.cand_flag[s] = 1;
            a.cand_flag[a.nsf + c] = 1;
        }
        double sv[3][3], cv[3][3];
        if (TRI) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                load_ps_vertex(a, __ldg(a.face_vtx + hs + i), sv[i]);
                load_ps_vertex(a, __ldg(a.face_vtx + hc + i), cv[i]);
            }
        }
        double sbox[6], cbox[6];
        load_box(a.src_bbox + 6 * (size_t)s, sbox);
        load_box(a.cut_bbox + 6 * (size_t)c, cbox);
        const face_view SF { &a, hs, ns }, CF { &a, hc, nc };
        const uint32_t nslots = ns + nc;
        uint32_t pre_edge[6];
        uint2 pre_ef[6];
        constexpr bool PREFETCH = TRI && !EXACT;
        if (PREFETCH) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                pre_edge[k] = __ldg(a.face_edge + hs + k);
                pre_edge[3 + k] = __ldg(a.face_edge + hc + k);
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) pre_ef[k] = __ldg(reinterpret_cast<const uint2*>(a.edge_f) + pre_edge[k]);
        }
        auto run_slot = [&](uint32_t slot) {
            uint32_t edge, tested_face;
            bool from_src;
            double q[3], r[3];
            if (!setup_test<TRI>(a, s, c, slot, hs, ns, hc, nc, sv, cv, sbox, cbox, edge, tested_face, from_src, q, r,
                    PREFETCH ? pre_edge : nullptr, PREFETCH ? pre_ef : nullptr))
                return;
            test_out_t o;
            bool needs_exact = false;
            const bool done = from_src ? eval_test<TRI, EXACT>(CF, cv, q, r, o, gp, needs_exact)
                                       : eval_test<TRI, EXACT>(SF, sv, q, r, o, gp, needs_exact);
            if (!EXACT) {
                n_tests_local++;
                n_exact_local += needs_exact ? 1u : 0u;
            }
            if (!done) {
                // slot index fits 8 bits only for faces with < 128 vertices each; larger faces are evaluated in place
                if (nslots <= 255u) {
                    const unsigned long long qs = alloc_slot(&a.counters->n_queue);
                    if (qs < a.cap_exact) a.exact_queue[qs] = (pair_index << 8) | slot;
                    return;
                }
                if (from_src) eval_test<TRI, true>(CF, cv, q, r, o, gp, needs_exact);
                else eval_test<TRI, true>(SF, sv, q, r, o, gp, needs_exact);
            }
            emit(a, edge, tested_face, o);
            log_test(a, edge, tested_face, o);
        };
        if (TRI && !EXACT) {
            // six slots, unrolled: every vertex / edge-id selection below is a compile-time index (no local-memory arrays)
#pragma unroll
            for (uint32_t slot = 0; slot < 6u; ++slot) run_slot(slot);
        } else {
            for (uint32_t slot = EXACT ? only_slot : 0u; slot < (EXACT ? only_slot + 1u : nslots); ++slot) run_slot(slot);
        }
    }
    // counters: one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_tests_local += __shfl_xor_sync(0xffffffffu, n_tests_local, o);
        n_exact_local += __shfl_xor_sync(0xffffffffu, n_exact_local, o);
        gp |= __shfl_xor_sync(0xffffffffu, gp, o);
    }
    if (lane_id() == 0) {
        if (n_tests_local) atomicAdd(&a.counters->n_tests, (unsigned long long)n_tests_local);
        if (n_exact_local) atomicAdd(&a.counters->n_exact, (unsigned long long)n_exact_local);
        if (gp) atomicOr(&a.counters->gp_violation, 1u);
    }
}

// ---- triangle meshes: prefilter -> classify -> resolve -> (general exact kernel) ----------------------------------------
// Nearly all candidate pairs of a dense overlap are two triangles that merely lie close to each other: every vertex of one
// is clearly on one side of the other's plane, so none of the six edge/face tests can cross.  Three kernels, each with
// dense warps of one kind of work and the registers that work needs:
//   k_tri_prefilter  one thread per pair: side of each vertex w.r.t. the other triangle's plane with a bound that implies
//                    orient3d's stage-A certificate (pred::orient3d_side_prefilter).  A test whose two endpoints are on the
//                    same certified side is a certified non-crossing — it has no output in the reference either, whether
//                    the ownership / box culls would have let it run or not — and is dropped here.  Pairs with anything
//                    left go to the mid queue with the sides found so far.
//   k_tri_classify   one thread per queued pair, the undecided slots only: ownership + box culls (kernel.cpp:2004-2115),
//                    then the real stage A where a side is still open.  Outcomes: certified non-crossing (done), certified
//                    crossing (second half of the exact queue), stage A failed (first half).
//   k_tri_resolve    one thread per queued test, stage-A failures first: exact sign when the differences are exact
//                    (pred::det3_sign_exact), plane point, point-in-triangle, registry record.  A test with inexact
//                    differences is handed to the general kernel (k_tests<true, true>, Shewchuk's stages B-D).
// Mid-queue entry:   pair index << 24 | sides << 12 | count-only slots << 6 | slots to test
//                    (sides: 2 bits per vertex, source vertices 0-2 vs the cut plane then cut vertices vs the source plane;
//                     0 open, 1 orient3d > 0, 2 orient3d < 0)
// Exact-queue entry: pair index << 8 | sr>0 << 7 | sq>0 << 6 | r open << 5 | q open << 4 | is_h0 << 3 | slot
constexpr int PF_THREADS = 256;

__device__ __forceinline__ void load_tri(const narrow_args_t& a, uint32_t h, double (*v)[3])
{
    const uint32_t i0 = __ldg(a.face_vtx + h), i1 = __ldg(a.face_vtx + h + 1), i2 = __ldg(a.face_vtx + h + 2);
    load_ps_vertex(a, i0, v[0]);
    load_ps_vertex(a, i1, v[1]);
    load_ps_vertex(a, i2, v[2]);
}

__device__ __forceinline__ void stage_frames(const narrow_args_t& a_in, frame_t* s_fr)
{
    load_frame_shared(&s_fr[0], a_in.src_frame);
    if (threadIdx.x >= 64) { // second half of the block fetches the cut frame
        constexpr unsigned W = sizeof(frame_t) / 4;
        const unsigned t = threadIdx.x - 64u;
        if (t < W) reinterpret_cast<unsigned*>(&s_fr[1])[t] = __ldg(reinterpret_cast<const unsigned*>(a_in.cut_frame) + t);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PF_THREADS, 4) k_tri_prefilter(narrow_args_t a_in)
{
    pdl_prologue();
    __shared__ frame_t s_fr[2];
    stage_frames(a_in, s_fr);
    narrow_args_t a = a_in;
    a.src_frame = &s_fr[0];
    a.cut_frame = &s_fr[1];
    const unsigned long long n_items = a.counters->n_pairs < a.cap_pairs ? a.counters->n_pairs : a.cap_pairs;
    for (unsigned long long it = (unsigned long long)blockIdx.x * PF_THREADS + threadIdx.x; it < n_items;
         it += (unsigned long long)gridDim.x * PF_THREADS) {
        const unsigned long long pr = a.pairs[it];
        const uint32_t s = (uint32_t)(pr >> 32), c = (uint32_t)(pr & 0xFFFFFFFFu);
        a.cand_flag[s] = 1;
        a.cand_flag[a.nsf + c] = 1;
        unsigned sides = 0, todo = 63u;
        if (a.tri_mode != 2u) {
            double sv[3][3], cv[3][3], nrm[3];
            load_tri(a, 3u * s, sv);
            load_tri(a, 3u * (a.nsf + c), cv);
            int sd[6];
            double luv = pred::side_prefilter_plane(cv[0], cv[1], cv[2], nrm);
#pragma unroll
            for (int j = 0; j < 3; ++j) sd[j] = pred::orient3d_side_prefilter(nrm, luv, cv[0], sv[j]);
            luv = pred::side_prefilter_plane(sv[0], sv[1], sv[2], nrm);
#pragma unroll
            for (int j = 0; j < 3; ++j) sd[3 + j] = pred::orient3d_side_prefilter(nrm, luv, sv[0], cv[j]);
            todo = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                sides |= (sd[k] > 0 ? 1u : (sd[k] < 0 ? 2u : 0u)) << (2 * k);
                // halfedge k of a face runs from its vertex k-1 to its vertex k
                const int base = k < 3 ? 0 : 3, i = k - base, ip = (i + 2) % 3;
                const bool same_side = sd[base + i] != 0 && sd[base + i] == sd[base + ip];
                todo |= (same_side ? 0u : 1u) << k;
            }
        }
        const unsigned count_only = a.tri_mode == 1u ? (~todo & 63u) : 0u;
        if (todo | count_only) {
            const unsigned long long slot = alloc_slot(&a.counters->n_mid);
            if (slot < a.cap_mid) a.mid_queue[slot] = (it << 24) | ((unsigned long long)sides << 12) | (count_only << 6) | todo;
        }
    }
}

__global__ void __launch_bounds__(NBLOCK) k_tri_classify(narrow_args_t a_in)
{
    pdl_prologue();
    __shared__ frame_t s_fr[2];
    stage_frames(a_in, s_fr);
    narrow_args_t a = a_in;
    a.src_frame = &s_fr[0];
    a.cut_frame = &s_fr[1];
    const unsigned long long n_items = a.counters->n_mid < a.cap_mid ? a.counters->n_mid : a.cap_mid;
    const unsigned long long half = a.cap_exact / 2;
    unsigned n_tests_local = 0, n_exact_local = 0;
    for (unsigned long long it = (unsigned long long)blockIdx.x * NBLOCK + threadIdx.x; it < n_items;
         it += (unsigned long long)gridDim.x * NBLOCK) {
        const unsigned long long me = a.mid_queue[it];
        const unsigned long long pair_index = me >> 24;
        const unsigned sides = (unsigned)(me >> 12) & 0xFFFu, count_only = (unsigned)(me >> 6) & 63u, todo = (unsigned)me & 63u;
        const unsigned long long pr = a.pairs[pair_index];
        const uint32_t s = (uint32_t)(pr >> 32), c = (uint32_t)(pr & 0xFFFFFFFFu);
        const uint32_t hs = 3u * s, hc = 3u * (a.nsf + c);
        double sv[3][3], cv[3][3];
        load_tri(a, hs, sv);
        load_tri(a, hc, cv);
        double sbox[6], cbox[6];
        load_box(a.src_bbox + 6 * (size_t)s, sbox);
        load_box(a.cut_bbox + 6 * (size_t)c, cbox);
        uint32_t pre_edge[6];
        uint2 pre_ef[6];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            pre_edge[k] = __ldg(a.face_edge + hs + k);
            pre_edge[3 + k] = __ldg(a.face_edge + hc + k);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) pre_ef[k] = __ldg(reinterpret_cast<const uint2*>(a.edge_f) + pre_edge[k]);
        // six slots, unrolled: every vertex / edge-id selection below is a compile-time index (no local-memory arrays)
#pragma unroll
        for (uint32_t slot = 0; slot < 6u; ++slot) {
            if (!(((todo | count_only) >> slot) & 1u)) continue;
            uint32_t edge, tested_face;
            bool from_src, is_h0;
            double q[3], r[3];
            if (!setup_test<true>(a, s, c, slot, hs, 3u, hc, 3u, sv, cv, sbox, cbox, edge, tested_face, from_src, q, r, pre_edge,
                    pre_ef, &is_h0))
                continue;
            n_tests_local++;
            if ((count_only >> slot) & 1u) continue;
            // sides of the halfedge's two vertices as the prefilter left them: q = source(h0), r = target(h0)
            const uint32_t i = slot < 3u ? slot : slot - 3u, ip = (i + 2u) % 3u, vb = slot < 3u ? 0u : 3u;
            const unsigned s_from = (sides >> (2u * (vb + ip))) & 3u, s_to = (sides >> (2u * (vb + i))) & 3u;
            unsigned side_q = is_h0 ? s_from : s_to, side_r = is_h0 ? s_to : s_from;
            unsigned open_q = 0, open_r = 0;
            if (side_q == 0u || side_r == 0u) {
                const double(*gv)[3] = slot < 3u ? cv : sv; // (slot is a constant here: no pointer selection at run time)
                bool cq, cr;
                double permq, permr;
                const double detq = pred::orient3d_stageA(gv[0], gv[1], gv[2], q, cq, permq);
                const double detr = pred::orient3d_stageA(gv[0], gv[1], gv[2], r, cr, permr);
                open_q = cq ? 0u : 1u;
                open_r = cr ? 0u : 1u;
                side_q = cq ? (detq > 0.0 ? 1u : 2u) : 0u;
                side_r = cr ? (detr > 0.0 ? 1u : 2u) : 0u;
                n_exact_local += (open_q | open_r);
                if (!(open_q | open_r) && side_q == side_r) { // certified non-crossing
                    if (a.tests) {
                        test_out_t o;
                        classify_signs(side_q == 1u ? 1 : -1, side_r == 1u ? 1 : -1, o);
                        o.exact = 0;
                        log_test(a, edge, tested_face, o);
                    }
                    continue;
                }
            }
            const unsigned long long entry = (pair_index << 8) | ((side_r == 1u ? 1u : 0u) << 7) | ((side_q == 1u ? 1u : 0u) << 6)
                | (open_r << 5) | (open_q << 4) | ((is_h0 ? 1u : 0u) << 3) | slot;
            if (open_q | open_r) {
                const unsigned long long qs = alloc_slot(&a.counters->n_queue);
                if (qs < half) a.exact_queue[qs] = entry;
            } else {
                const unsigned long long qs = alloc_slot(&a.counters->n_cross);
                if (qs < half) a.exact_queue[half + qs] = entry;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_tests_local += __shfl_xor_sync(0xffffffffu, n_tests_local, o);
        n_exact_local += __shfl_xor_sync(0xffffffffu, n_exact_local, o);
    }
    if (lane_id() == 0) {
        if (n_tests_local) atomicAdd(&a.counters->n_tests, (unsigned long long)n_tests_local);
        if (n_exact_local) atomicAdd(&a.counters->n_exact, (unsigned long long)n_exact_local);
    }
}

__global__ void __launch_bounds__(NBLOCK) k_tri_resolve(narrow_args_t a_in)
{
    pdl_prologue();
    __shared__ frame_t s_fr[2];
    stage_frames(a_in, s_fr);
    narrow_args_t a = a_in;
    a.src_frame = &s_fr[0];
    a.cut_frame = &s_fr[1];
    const unsigned long long half = a.cap_exact / 2;
    const unsigned long long n_open = a.counters->n_queue < half ? a.counters->n_queue : half;
    const unsigned long long n_cross = a.counters->n_cross < half ? a.counters->n_cross : half;
    unsigned gp = 0;
    for (unsigned long long it = (unsigned long long)blockIdx.x * NBLOCK + threadIdx.x; it < n_open + n_cross;
         it += (unsigned long long)gridDim.x * NBLOCK) {
        const unsigned long long e = it < n_open ? a.exact_queue[it] : a.exact_queue[half + (it - n_open)];
        const unsigned long long pair_index = e >> 8;
        const uint32_t slot = (uint32_t)e & 7u;
        const bool is_h0 = (e >> 3) & 1u, open_q = (e >> 4) & 1u, open_r = (e >> 5) & 1u;
        const unsigned long long pr = a.pairs[pair_index];
        const uint32_t s = (uint32_t)(pr >> 32), c = (uint32_t)(pr & 0xFFFFFFFFu);
        const bool from_src = slot < 3u;
        const uint32_t i = from_src ? slot : slot - 3u, ip = (i + 2u) % 3u;
        const uint32_t h_own = from_src ? 3u * s : 3u * (a.nsf + c), h_tst = from_src ? 3u * (a.nsf + c) : 3u * s;
        const uint32_t tested_face = from_src ? a.nsf + c : s;
        double gv[3][3], from[3], to[3];
        load_tri(a, h_tst, gv);
        load_ps_vertex(a, __ldg(a.face_vtx + h_own + ip), from);
        load_ps_vertex(a, __ldg(a.face_vtx + h_own + i), to);
        const uint32_t edge = __ldg(a.face_edge + h_own + i);
        double q[3], r[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            q[k] = is_h0 ? from[k] : to[k];
            r[k] = is_h0 ? to[k] : from[k];
        }
        int sq = ((e >> 6) & 1u) ? 1 : -1, sr = ((e >> 7) & 1u) ? 1 : -1;
        bool decided = true;
        if (open_q | open_r) {
            // orient3d(gv0, gv1, gv2, p): the rows of stage A are gv_k - p
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                if (!(which == 0 ? open_q : open_r) || !decided) continue;
                const double* p = which == 0 ? q : r;
                double dx[9];
                bool exact_rows = true;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        dx[3 * k + j] = pred::sub(gv[k][j], p[j]);
                        exact_rows = exact_rows && pred::two_diff_tail(gv[k][j], p[j], dx[3 * k + j]) == 0.0;
                    }
                }
                int sg = 0;
                if (exact_rows && pred::det3_sign_exact(dx[0], dx[1], dx[2], dx[3], dx[4], dx[5], dx[6], dx[7], dx[8], sg)) {
                    if (which == 0) sq = sg;
                    else sr = sg;
                } else {
                    decided = false;
                }
            }
        }
        if (!decided) { // inexact differences: Shewchuk's adaptive stages, in the general kernel
            const unsigned long long qs = alloc_slot(&a.counters->n_full);
            if (qs < a.cap_mid) a.mid_queue[qs] = (pair_index << 8) | slot;
            continue;
        }
        test_out_t o;
        classify_signs(sq, sr, o);
        o.exact = (uint8_t)((open_q ? 1 : 0) | (open_r ? 2 : 0));
        if (o.type != '0') {
            const face_view G { &a, h_tst, 3u };
            double normal[3];
            finish_touching<true>(G, gv, q, r, o, gp, false, normal, 0.0, 0);
            emit(a, edge, tested_face, o);
        }
        log_test(a, edge, tested_face, o);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gp |= __shfl_xor_sync(0xffffffffu, gp, o);
    if (lane_id() == 0 && gp) atomicOr(&a.counters->gp_violation, 1u);
}

// ---- per-candidate-face plane data + degenerate-face rule (kernel.cpp:2184-2356) ----------------------------------------
struct plane_args_t {
    narrow_args_t n;
    double* plane; // [cap][4] normal, d   (compacted, unordered; the face id travels in plane_face)
    int32_t* plane_mc;
    uint32_t* plane_face;
};

template <bool TRI> __global__ void __launch_bounds__(NBLOCK) k_planes(plane_args_t pa)
{
    pdl_prologue();
    __shared__ frame_t s_fr[2];
    load_frame_shared(&s_fr[0], pa.n.src_frame);
    if (threadIdx.x >= 64) {
        constexpr unsigned W = sizeof(frame_t) / 4;
        const unsigned t = threadIdx.x - 64u;
        if (t < W) reinterpret_cast<unsigned*>(&s_fr[1])[t] = __ldg(reinterpret_cast<const unsigned*>(pa.n.cut_frame) + t);
    }
    __syncthreads();
    narrow_args_t a = pa.n;
    a.src_frame = &s_fr[0];
    a.cut_frame = &s_fr[1];
    for (uint32_t f = blockIdx.x * NBLOCK + threadIdx.x; f < a.nf; f += gridDim.x * NBLOCK) {
        if (!a.cand_flag[f]) continue;
        const uint32_t h0 = TRI ? 3u * f : __ldg(a.face_off + f);
        const uint32_t n = TRI ? 3u : __ldg(a.face_off + f + 1) - h0;
        double gv[3][3];
        if (TRI)
#pragma unroll
            for (int i = 0; i < 3; ++i) load_ps_vertex(a, __ldg(a.face_vtx + h0 + i), gv[i]);
        const face_view G { &a, h0, n };
        double normal[3], d;
        const int mc = face_plane<TRI>(G, gv, normal, d);
        // kernel.cpp:2237-2244: squared_length(normal) == 0 or NaN -> the mesh is invalid
        if (pred::dot3(normal, normal) == 0.0 || isnan(normal[0]) || isnan(normal[1]) || isnan(normal[2]))
            atomicMin(&a.counters->bad_face, f);
        const unsigned long long slot = alloc_slot(&a.counters->n_cand_faces);
        pa.plane[4 * slot + 0] = normal[0];
        pa.plane[4 * slot + 1] = normal[1];
        pa.plane[4 * slot + 2] = normal[2];
        pa.plane[4 * slot + 3] = d;
        pa.plane_mc[slot] = mc;
        pa.plane_face[slot] = f;
    }
}

// ---- canonical ordering of records / logged tests -----------------------------------------------------------------------
// Up to SMALL_SORT items (the usual case: an intersection curve crosses thousands of edges, not millions) are ordered by a
// single block (k_rank_sort_small); the radix path's kernels return at once when n <= SMALL_SORT.
constexpr unsigned SMALL_SORT = 16384;

template <typename T> __device__ __forceinline__ unsigned long long item_key(const T* item)
{
    // {uint32 edge, uint32 face} lead every item: one 8-byte load, halves swapped into (edge << 32 | face)
    const unsigned long long v = __ldg(reinterpret_cast<const unsigned long long*>(item));
    return (v << 32) | (v >> 32);
}

// Order by counting: the rank of an item is the number of items with a smaller (edge, face) key — keys are unique — and the
// item is written straight to its place.  64 items per block, four threads share one item's scan of all keys (staged through
// shared memory 4096 at a time, 32 KB static: no shared-memory carve-out change between this kernel and its neighbours,
// which a 128 KB bucket table cost more than it saved).  One launch instead of key extraction + histogram + six radix
// passes + gather; those kernels return at once when n <= SMALL_SORT.
constexpr int RANK_CHUNK = 4096; // keys staged per round (32 KB)
constexpr int RANK_ITEMS = 64; // items per block
constexpr int RANK_SHARE = 256 / RANK_ITEMS; // threads per item

template <typename T> __global__ void __launch_bounds__(256) k_rank_sort_small(const T* __restrict__ items,
    const unsigned long long* d_n, unsigned long long cap, T* __restrict__ out)
{
    pdl_prologue();
    static_assert(sizeof(T) % 8 == 0 && alignof(T) >= 8 && offsetof(T, edge) == 0 && offsetof(T, face) == 4, "item layout");
    __shared__ unsigned long long s_keys[RANK_CHUNK];
    const unsigned long long n64 = *d_n < cap ? *d_n : cap;
    if (n64 > SMALL_SORT || (unsigned long long)blockIdx.x * RANK_ITEMS >= n64) return;
    const unsigned n = (unsigned)n64;
    const unsigned i = blockIdx.x * RANK_ITEMS + (threadIdx.x / RANK_SHARE), sub = threadIdx.x % RANK_SHARE;
    const bool live = i < n;
    const unsigned long long mine = live ? item_key(items + i) : ~0ull;
    unsigned rank = 0;
    for (unsigned base = 0; base < n; base += RANK_CHUNK) {
        const unsigned cn = (n - base < (unsigned)RANK_CHUNK) ? n - base : (unsigned)RANK_CHUNK;
        unsigned long long tmp[RANK_CHUNK / 256];
#pragma unroll
        for (int k = 0; k < RANK_CHUNK / 256; ++k) {
            const unsigned j = k * 256u + threadIdx.x;
            tmp[k] = j < cn ? item_key(items + base + j) : ~0ull;
        }
#pragma unroll
        for (int k = 0; k < RANK_CHUNK / 256; ++k) s_keys[k * 256u + threadIdx.x] = tmp[k];
        __syncthreads();
        // RANK_SHARE consecutive keys per step and item; the 8 items of a warp read the same addresses (broadcast)
        const unsigned steps = (cn + RANK_SHARE - 1u) / RANK_SHARE;
#pragma unroll 8
        for (unsigned st = 0; st < steps; ++st) rank += (s_keys[st * RANK_SHARE + sub] < mine) ? 1u : 0u;
        __syncthreads();
    }
#pragma unroll
    for (int o = 1; o < RANK_SHARE; o <<= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
    if (live) {
        // the item is a few 8-byte words: the threads of the group copy it together
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(items + i);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(out + rank);
        for (unsigned w = sub; w < sizeof(T) / 8; w += RANK_SHARE) dst[w] = src[w];
    }
}

template <typename T> __global__ void __launch_bounds__(256) k_make_keys(const T* items, const unsigned long long* d_n,
    unsigned long long cap, unsigned long long* keys)
{
    pdl_prologue();
    const unsigned long long n = *d_n < cap ? *d_n : cap;
    if (n <= SMALL_SORT) return;
    for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * 256)
        keys[i] = ((unsigned long long)items[i].edge << 32) | items[i].face;
}

template <typename T> __global__ void __launch_bounds__(256) k_gather(const T* items, const uint32_t* idx,
    const unsigned long long* d_n, unsigned long long cap, T* out)
{
    pdl_prologue();
    const unsigned long long n = *d_n < cap ? *d_n : cap;
    if (n <= SMALL_SORT) return;
    for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * 256)
        out[i] = items[idx[i]];
}

static int bits_for(uint32_t n)
{
    int b = 1;
    while (b < 32 && (1ull << b) < (unsigned long long)n) ++b;
    return b;
}

// part: 1 = the small-n kernel, 2 = the radix path (its kernels return at once when the small-n kernel applies), 3 = both.
// The two parts may run on different lanes: they write items_sorted under mutually exclusive conditions.
template <typename T>
int sort_items(mcb200_ctx* ctx, const mcb200_result* res, const T* items, T* items_sorted, dbuf& keys, dbuf& idx,
    const unsigned long long* d_n, size_t cap, uint32_t nfaces, uint32_t nedges_bound, int part)
{
    if (cap == 0) return 0;
    (void)res;
    const rsort::pass_desc pd = rsort::make_passes(0, bits_for(nfaces), 32, 32 + bits_for(nedges_bound));
    MCB_TRY(ctx->reserve(keys, sizeof(unsigned long long) * cap));
    MCB_TRY(ctx->reserve(idx, sizeof(uint32_t) * cap));
    MCB_TRY((rsort::reserve_scratch<unsigned long long>(ctx, cap, pd.npasses, true, true)));
    mcb200_ctx::sort_scratch_t& sc = ctx->sc();
    const unsigned grid = (unsigned)ctx->num_sms * 2u;
    if (part & 1)
        MCB_LAUNCH_NAMED(ctx, "k_rank_sort_small", (k_rank_sort_small<T>), SMALL_SORT / RANK_ITEMS, 256, 0, items, d_n,
            (unsigned long long)cap, items_sorted);
    if (!(part & 2)) return 0;
    MCB_LAUNCH_NAMED(ctx, "k_make_keys", (k_make_keys<T>), grid, 256, 0, items, d_n, (unsigned long long)cap, keys.as<unsigned long long>());
    unsigned long long* kout = nullptr;
    uint32_t* vout = nullptr;
    MCB_TRY((rsort::sort<unsigned long long, uint32_t, true>(ctx, keys.as<unsigned long long>(), sc.keys_alt.as<unsigned long long>(),
        keys.as<unsigned long long>(), nullptr, sc.vals_alt.as<uint32_t>(), idx.as<uint32_t>(), d_n, cap, pd, &kout, &vout, SMALL_SORT)));
    MCB_LAUNCH_NAMED(ctx, "k_gather", (k_gather<T>), grid, 256, 0, items, vout, d_n, (unsigned long long)cap, items_sorted);
    return 0;
}

} // namespace

// All allocations of the narrowphase (and of its record/test sorts, in the CURRENT scratch set).  Capacities are sized
// from the pair CAPACITY, not the pair count, so nothing between traversal and narrowphase needs the host.
int narrowphase_reserve(mcb200_ctx* ctx, const mcb200_soup* soup, mcb200_result* res, uint32_t flags)
{
    const uint32_t nf = soup->nsf + soup->ncf;
    const bool tri = soup->all_tri != 0;
    const bool want_log = (flags & MCB200_NARROW_LOG_TESTS) != 0;
    // a pair yields at most ns + nc tests; records <= tests
    const size_t avg_slots = tri ? 6 : (size_t)((soup->nh + nf - 1) / nf) * 2 + 2;
    size_t cap_rec = res->cap_pairs * 2;
    size_t cap_exact = res->cap_pairs * avg_slots;
    if (cap_exact > (size_t)1 << 28) cap_exact = (size_t)1 << 28;
    MCB_TRY(ctx->reserve(res->records, sizeof(mcb200_record) * cap_rec));
    MCB_TRY(ctx->reserve(res->records_sorted, sizeof(mcb200_record) * cap_rec));
    res->cap_records = cap_rec;
    MCB_TRY(ctx->reserve(res->exact_queue, sizeof(unsigned long long) * cap_exact));
    res->cap_exact = cap_exact;
    if (tri) MCB_TRY(ctx->reserve(res->mid_queue, sizeof(unsigned long long) * res->cap_pairs));
    res->tri_queues = tri;
    MCB_TRY(ctx->reserve(res->cand_flag, (size_t)nf));
    MCB_TRY(ctx->reserve(res->plane, sizeof(double) * 4 * (size_t)nf));
    MCB_TRY(ctx->reserve(res->plane_mc, sizeof(int32_t) * 2 * (size_t)nf));
    size_t cap_tests = 0;
    if (want_log) {
        cap_tests = res->cap_pairs * avg_slots;
        if (cap_tests > (size_t)1 << 26) cap_tests = (size_t)1 << 26;
        MCB_TRY(ctx->reserve(res->tests, sizeof(mcb200_test) * cap_tests));
        MCB_TRY(ctx->reserve(res->tests_sorted, sizeof(mcb200_test) * cap_tests));
        MCB_TRY(ctx->reserve(res->test_keys, sizeof(unsigned long long) * cap_tests));
        MCB_TRY(ctx->reserve(res->test_idx, sizeof(uint32_t) * cap_tests));
    }
    res->cap_tests = cap_tests;
    res->logged_tests = want_log;
    res->ne_ps = soup->ne;
    res->nf_ps = nf;
    const size_t big = cap_tests > cap_rec ? cap_tests : cap_rec;
    MCB_TRY(ctx->reserve(res->rec_keys, sizeof(unsigned long long) * cap_rec));
    MCB_TRY(ctx->reserve(res->rec_idx, sizeof(uint32_t) * cap_rec));
    MCB_TRY((rsort::reserve_scratch<unsigned long long>(ctx, big, rsort::MAX_PASSES, true, true)));
    return 0;
}

static int make_narrow_args(mcb200_ctx* ctx, const mcb200_soup* soup, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_result* res,
    bool want_log, narrow_args_t& a)
{
    a.src_xyz = src->d_xyz;
    a.cut_xyz = cut->d_xyz;
    MCB_TRY(mesh_sync_frames(ctx, const_cast<mcb200_mesh*>(src), src->eps, const_cast<mcb200_mesh*>(cut), cut->eps));
    a.src_frame = src->d_frames.as<frame_t>() + 1;
    a.cut_frame = cut->d_frames.as<frame_t>() + 1;
    a.src_nv = src->nv;
    a.src_bbox = src->face_bbox.as<double>();
    a.cut_bbox = cut->face_bbox.as<double>();
    a.face_off = soup->face_off.as<uint32_t>();
    a.face_vtx = soup->face_vtx.as<uint32_t>();
    a.face_edge = soup->face_edge.as<uint32_t>();
    a.edge_f = soup->edge_f.as<uint32_t>();
    a.nsf = soup->nsf;
    a.nf = soup->nsf + soup->ncf;
    a.pairs = res->pairs.as<unsigned long long>();
    a.cap_pairs = res->cap_pairs;
    a.counters = res->counters.as<result_counters_t>();
    a.cand_flag = res->cand_flag.as<uint8_t>();
    a.records = res->records.as<mcb200_record>();
    a.cap_records = res->cap_records;
    a.exact_queue = res->exact_queue.as<unsigned long long>();
    a.cap_exact = res->cap_exact;
    a.mid_queue = res->mid_queue.as<unsigned long long>();
    a.cap_mid = res->cap_pairs;
    if (soup->all_tri) { // what k_tri_resolve could not settle, parked in the (by then idle) mid queue
        a.exact_in = a.mid_queue;
        a.exact_in_n = &a.counters->n_full;
        a.exact_in_cap = a.cap_mid;
    } else {
        a.exact_in = a.exact_queue;
        a.exact_in_n = &a.counters->n_queue;
        a.exact_in_cap = a.cap_exact;
    }
    a.tri_mode = 0;
    a.tests = want_log ? res->tests.as<mcb200_test>() : nullptr;
    a.cap_tests = res->cap_tests;
    return 0;
}

// The plane data of the candidate faces (an OUTPUT, consumed downstream by the host; the tests compute the planes they need
// themselves): one row per face flagged in res->cand_flag, on ctx->cur.
int narrowphase_planes(mcb200_ctx* ctx, const mcb200_soup* soup, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_result* res)
{
    const uint32_t nf = soup->nsf + soup->ncf;
    plane_args_t pa;
    MCB_TRY(make_narrow_args(ctx, soup, src, cut, res, false, pa.n));
    pa.plane = res->plane.as<double>();
    pa.plane_mc = res->plane_mc.as<int32_t>();
    pa.plane_face = reinterpret_cast<uint32_t*>(res->plane_mc.as<int32_t>() + nf);
    const unsigned grid = (unsigned)ctx->num_sms * 8u;
    const unsigned pgrid = div_up(nf, NBLOCK) < grid ? div_up(nf, NBLOCK) : grid;
    if (soup->all_tri) MCB_LAUNCH(ctx, k_planes<true>, pgrid, NBLOCK, 0, pa);
    else MCB_LAUNCH(ctx, k_planes<false>, pgrid, NBLOCK, 0, pa);
    return 0;
}

int narrowphase_run(mcb200_ctx* ctx, const mcb200_soup* soup, const mcb200_mesh* src, const mcb200_mesh* cut,
    mcb200_result* res, uint32_t flags)
{
    if (!res->have_pairs) {
        ctx->set_error("narrowphase: run mcb200_bvh_intersect first", __FILE__, __LINE__);
        return MCB200_ERR_INVALID;
    }
    if (soup->nsf != src->nf || soup->ncf != cut->nf) {
        ctx->set_error("narrowphase: soup does not match the meshes", __FILE__, __LINE__);
        return MCB200_ERR_INVALID;
    }
    const uint32_t nf = soup->nsf + soup->ncf;
    const bool tri = soup->all_tri != 0;
    const bool want_log = (flags & MCB200_NARROW_LOG_TESTS) != 0;
    // a shard of a multi-GPU dispatch: the candidate-face rows and the canonical orders are made after the exchange (comm.cu)
    const bool partial = (flags & MCB200_NARROW_INTERNAL_PARTIAL) != 0;
    MCB_TRY(narrowphase_reserve(ctx, soup, res, flags));
    if (!res->cand_flag_fresh) MCB_CUDA(ctx, cudaMemsetAsync(res->cand_flag.p, 0, (size_t)nf, ctx->cur));
    res->cand_flag_fresh = false;

    // reset the narrowphase counters only (pairs / node tests stay)
    if (!res->narrow_counters_fresh) { // a rerun on existing pairs (e.g. another perturbation)
        result_counters_t* c = res->counters.as<result_counters_t>();
        fill_list_t fl {};
        fl.add(&c->n_tests, 10, 0u); // n_tests .. n_log
        fl.add(&c->n_queue, 8, 0u); // n_queue, n_mid, n_cross, n_full
        fl.add(&c->gp_violation, 1, 0u);
        fl.add(&c->bad_face, 1, 0xFFFFFFFFu);
        MCB_LAUNCH(ctx, k_fill, 1, 256, 0, fl);
    }
    res->narrow_counters_fresh = false;

    narrow_args_t a;
    MCB_TRY(make_narrow_args(ctx, soup, src, cut, res, want_log, a));

    const unsigned grid = (unsigned)ctx->num_sms * 8u;
    if (tri) {
        a.tri_mode = want_log ? 2u : ((flags & MCB200_NARROW_COUNT_TESTS) ? 1u : 0u);
        MCB_LAUNCH(ctx, k_tri_prefilter, grid, PF_THREADS, 0, a);
        MCB_LAUNCH(ctx, k_tri_classify, grid, NBLOCK, 0, a);
    } else {
        MCB_LAUNCH_NAMED(ctx, "k_tests_filter_poly", (k_tests<false, false>), grid, NBLOCK, 0, a);
    }

    // The plane rows only depend on the candidate flags the filter just wrote, so they are made beside the exact pass and the
    // record sort on the background lane.
    cudaStream_t lane = ctx->cur;
    const int lane_sci = ctx->sci;
    if (!partial) {
        MCB_CUDA(ctx, cudaEventRecord(ctx->ev_np, lane));
        MCB_CUDA(ctx, cudaStreamWaitEvent(ctx->bg, ctx->ev_np, 0));
        ctx->use_bg();
        const int rcp = narrowphase_planes(ctx, soup, src, cut, res);
        ctx->cur = lane;
        ctx->sci = lane_sci;
        if (rcp) return rcp;
    }

    // exact-expansion pass over the compacted filter failures (own kernel: its local-memory footprint and divergence
    // stay out of the filter kernel)
    const unsigned egrid = (unsigned)ctx->num_sms * 4u;
    if (tri) {
        MCB_LAUNCH(ctx, k_tri_resolve, grid, NBLOCK, 0, a);
        MCB_LAUNCH_NAMED(ctx, "k_tests_exact_tri", (k_tests<true, true>), egrid, NBLOCK, 0, a); // returns at once unless n_full > 0
    } else {
        MCB_LAUNCH_NAMED(ctx, "k_tests_exact_poly", (k_tests<false, true>), egrid, NBLOCK, 0, a);
    }

    res->have_narrow = true;
    res->h_valid = false;
    res->records_sorted_valid = false;
    res->tests_sorted_valid = false;
    if (partial) return 0;
    // Canonical order of the registry.  The small-n kernel stays on this lane; the radix path — nine launches that return
    // at once in the usual small case — goes to the background lane behind the plane kernel, so its launches are not paid
    // on the critical path.  (Scratch set 0 is free: this lane's last radix sort was the source mesh's Morton sort.)
    MCB_CUDA(ctx, cudaEventRecord(ctx->ev_np3, lane));
    MCB_CUDA(ctx, cudaStreamWaitEvent(ctx->bg, ctx->ev_np3, 0));
    ctx->use_bg();
    const bool lazy_radix = (flags & MCB200_NARROW_INTERNAL_LAZY_RADIX) != 0 && !want_log;
    res->record_radix_pending = lazy_radix;
    int rc = lazy_radix ? 0 : narrowphase_sort_records(ctx, res, 2);
    if (!rc && want_log) rc = narrowphase_sort_tests(ctx, res, 2);
    cudaEventRecord(ctx->ev_np2, ctx->bg);
    ctx->cur = lane;
    ctx->sci = lane_sci;
    if (rc) return rc;
    MCB_TRY(narrowphase_sort_records(ctx, res, 1));
    if (want_log) MCB_TRY(narrowphase_sort_tests(ctx, res, 1));
    MCB_CUDA(ctx, cudaStreamWaitEvent(ctx->cur, ctx->ev_np2, 0)); // the plane data and the radix path join here
    return 0;
}

// ps.get_vertices_around_face order of both meshes from their user face arrays (hmesh.cpp:705-733 returns halfedge
// targets: the user's list rotated by one; cut faces were handed to add_face already rotated, kernel.cpp:1678): slot i of a
// source face holds user[(i+1) % n], of a cut face user[(i+2) % n] + nsv.
namespace {
__global__ void __launch_bounds__(256) k_soup_face_vtx(const uint32_t* __restrict__ user_vtx, const uint32_t* __restrict__ face_off,
    uint32_t nf, uint32_t rot, uint32_t voff, uint32_t* __restrict__ ps_vtx, uint32_t* __restrict__ ps_off, uint32_t off_base,
    uint32_t face_base)
{
    pdl_prologue();
    for (uint32_t f = blockIdx.x * 256u + threadIdx.x; f < nf; f += gridDim.x * 256u) {
        const uint32_t h0 = face_off ? face_off[f] : 3u * f;
        const uint32_t n = face_off ? face_off[f + 1] - h0 : 3u;
        for (uint32_t i = 0; i < n; ++i) ps_vtx[off_base + h0 + i] = __ldg(user_vtx + h0 + (i + rot) % n) + voff;
        if (ps_off) {
            ps_off[face_base + f] = off_base + h0;
            if (f == nf - 1) ps_off[face_base + nf] = off_base + h0 + n;
        }
    }
}
} // namespace

int soup_face_vtx_device(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_soup* soup)
{
    const bool tri = src->is_tri && cut->is_tri;
    uint32_t* off = tri ? nullptr : soup->face_off.as<uint32_t>();
    const unsigned gs = div_up(src->nf, 256), gc = div_up(cut->nf, 256);
    MCB_LAUNCH(ctx, k_soup_face_vtx, gs, 256, 0, src->d_face_vtx, src->d_face_off, src->nf, 1u, 0u, soup->face_vtx.as<uint32_t>(), off,
        0u, 0u);
    MCB_LAUNCH(ctx, k_soup_face_vtx, gc, 256, 0, cut->d_face_vtx, cut->d_face_off, cut->nf, 2u, src->nv, soup->face_vtx.as<uint32_t>(),
        off, src->nh, src->nf);
    return 0;
}

// The stage calls clear the candidate flags up front (before the lanes fork), off the path between traversal and filter.
int narrowphase_prezero(mcb200_ctx* ctx, const mcb200_soup* soup, mcb200_result* res)
{
    MCB_CUDA(ctx, cudaMemsetAsync(res->cand_flag.p, 0, (size_t)soup->nsf + soup->ncf, ctx->cur));
    res->cand_flag_fresh = true;
    return 0;
}

int narrowphase_sort_records(mcb200_ctx* ctx, mcb200_result* res, int part)
{
    if (res->records_sorted_valid) return 0;
    result_counters_t* c = res->counters.as<result_counters_t>();
    MCB_TRY(sort_items<mcb200_record>(ctx, res, res->records.as<mcb200_record>(), res->records_sorted.as<mcb200_record>(),
        res->rec_keys, res->rec_idx, &c->n_records, res->cap_records, res->nf_ps, res->ne_ps, part));
    if (part & 1) res->records_sorted_valid = true; // narrowphase_run issues part 2 first, part 1 last
    return 0;
}

// fetch_counters has just read the counters of a run whose radix path was deferred (the stream is idle)
int narrowphase_finish_record_order(mcb200_ctx* ctx, mcb200_result* res)
{
    res->record_radix_pending = false;
    if (res->h.n_records <= SMALL_SORT || res->h.n_records > res->cap_records) return 0;
    ctx->use_main();
    result_counters_t* c = res->counters.as<result_counters_t>();
    MCB_TRY(sort_items<mcb200_record>(ctx, res, res->records.as<mcb200_record>(), res->records_sorted.as<mcb200_record>(), res->rec_keys,
        res->rec_idx, &c->n_records, res->cap_records, res->nf_ps, res->ne_ps, 2));
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int narrowphase_sort_tests(mcb200_ctx* ctx, mcb200_result* res, int part)
{
    if (res->tests_sorted_valid || !res->logged_tests) return 0;
    result_counters_t* c = res->counters.as<result_counters_t>();
    MCB_TRY(sort_items<mcb200_test>(ctx, res, res->tests.as<mcb200_test>(), res->tests_sorted.as<mcb200_test>(), res->test_keys,
        res->test_idx, &c->n_log, res->cap_tests, res->nf_ps, res->ne_ps, part));
    if (part & 1) res->tests_sorted_valid = true;
    return 0;
}
