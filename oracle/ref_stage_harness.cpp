// oracle/ref_stage_harness.cpp — observer harness around the UNMODIFIED reference (cutdigital/mcut).
// TEST INFRASTRUCTURE ONLY: it is how the oracle (and therefore the CUDA path) is pinned to what the
// reference itself computes.  Nothing under mcut_b200/ may link or execute this.
//
// How it works.  The executable links oracle/_ref/libmcut_ref.so (the reference compiled from
// /root/reference by oracle/Makefile) and is linked with -rdynamic.  Every hot-path function of the
// reference is an exported symbol that the reference calls through the PLT, so the
// definitions below *interpose*: the reference's mcDispatch runs its normal code path, and each hook
// forwards to the real function through dlsym(RTLD_NEXT) and records inputs/outputs on the way:
//
//   build_oibvh        include/mcut/internal/bvh.h:117-125   source/bvh.cpp:219-636
//   intersectOIBVHs    include/mcut/internal/bvh.h:127-133   source/bvh.cpp:638-783
//   dispatch           include/mcut/internal/kernel.h:227    source/kernel.cpp:1536
//   client_input_arrays_to_hmesh  source/preproc.cpp:57-468 (exposes com, shift and the perturbation of every retry)
//   dump_mesh          source/kernel.cpp:158-186 ("polygon-soup" at :1737 exposes ps, "m0.v" at :3260
//                      exposes ps vertices + intersection points in registry order)
//   compute_polygon_plane_coefficients / compute_segment_plane_intersection_type /
//   compute_segment_plane_intersection / compute_point_in_polygon_test   source/math.cpp:130,391,249,851
//   orient3d / orient2d (C)   source/shewchuk.c:2367,1695
//
// Usage: stage_harness <in.mcb> <out.mcb> [--helpers N] [--no-events] [--abort-after-narrowphase]
//                      [--no-cc] [--repeat R] [--planar nx ny nz offset]
// Input arrays: src_xyz (f64|f32 [V,3]), src_faces (u32), [src_sizes (u32)], cut_xyz, cut_faces,
// [cut_sizes], flags (u32 [1]).

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "mcut/mcut.h"
#include "mcut/internal/bvh.h"
#include "mcut/internal/frontend.h"
#include "mcut/internal/hmesh.h"
#include "mcut/internal/kernel.h"
#include "mcut/internal/math.h"

#include "mcb_io.hpp"

namespace {

typedef bounding_box_t<vec3_<double>> bbox_t;

struct abort_dispatch_t : public std::runtime_error {
    abort_dispatch_t()
        : std::runtime_error("harness: abort after narrowphase")
    {
    }
};

struct state_t {
    std::mutex mtx;
    mcb::file_t out;
    bool log_events = true;
    bool abort_after_narrowphase = false;
    int build_calls = 0;
    int isect_calls = 0;
    int dispatch_calls = 0;
    int c2h_calls = 0;
    std::vector<double> events; // [kind, len, payload...]
    size_t max_events = (size_t)1 << 28;
    std::vector<double> timings_ms; // [kind(0 build,1 isect,2 ps->m0.v), value]
    std::chrono::steady_clock::time_point t_ps;
    int cur_ps_vtx_cnt = 0;
};

state_t g;

template <typename F> F real_fn(const char* mangled)
{
    void* p = dlsym(RTLD_NEXT, mangled);
    if (!p) {
        std::fprintf(stderr, "harness: dlsym(%s) failed: %s\n", mangled, dlerror());
        std::abort();
    }
    return reinterpret_cast<F>(p);
}

std::string idx_name(const char* base, int i, const char* field)
{
    return std::string(base) + std::to_string(i) + "_" + field;
}

void dump_hmesh_arrays(mcb::file_t& f, const std::string& prefix, const hmesh_t& m)
{
    std::vector<double> xyz;
    xyz.reserve((size_t)m.number_of_vertices() * 3);
    for (vertex_array_iterator_t v = m.vertices_begin(); v != m.vertices_end(); ++v) {
        const vec3& p = m.vertex(*v);
        xyz.push_back(p.x());
        xyz.push_back(p.y());
        xyz.push_back(p.z());
    }
    mcb::put(f, prefix + "xyz", xyz, 3);
    std::vector<uint32_t> sizes, idx;
    std::vector<vd_t> tmp;
    for (face_array_iterator_t fi = m.faces_begin(); fi != m.faces_end(); ++fi) {
        m.get_vertices_around_face(tmp, *fi);
        sizes.push_back((uint32_t)tmp.size());
        for (const vd_t& v : tmp) idx.push_back((uint32_t)v);
    }
    mcb::put(f, prefix + "face_sizes", sizes);
    mcb::put(f, prefix + "face_vtx", idx);
}

void push_bbox(std::vector<double>& o, const bbox_t& b)
{
    o.push_back(b.minimum().x());
    o.push_back(b.minimum().y());
    o.push_back(b.minimum().z());
    o.push_back(b.maximum().x());
    o.push_back(b.maximum().y());
    o.push_back(b.maximum().z());
}

struct log_event_t {
    std::vector<double> d;
    explicit log_event_t(double kind)
    {
        d.push_back(kind);
        d.push_back(0.0);
    }
    void v3(const vec3& p)
    {
        d.push_back(p.x());
        d.push_back(p.y());
        d.push_back(p.z());
    }
    void s(double x) { d.push_back(x); }
    void commit()
    {
        d[1] = (double)(d.size() - 2);
        std::lock_guard<std::mutex> lk(g.mtx);
        if (g.events.size() + d.size() <= g.max_events) g.events.insert(g.events.end(), d.begin(), d.end());
    }
};

} // namespace

#ifndef HARNESS_NO_HOOKS
// ------------------------------------------------------------------------------------------------
// interposed reference functions
// ------------------------------------------------------------------------------------------------

void dump_mesh(const hmesh_t& mesh, const char* fbasename, const double /*multiplier*/)
{
    // replaces source/kernel.cpp:158-186 (which writes <name>.off into the CWD on every dispatch)
    const std::string name(fbasename);
    if (name == "polygon-soup") {
        std::lock_guard<std::mutex> lk(g.mtx);
        g.t_ps = std::chrono::steady_clock::now();
        g.cur_ps_vtx_cnt = mesh.number_of_vertices();
        const int k = g.dispatch_calls - 1;
        if (g.log_events) {
            // ps edge table: source(h0), target(h0), face(h0), face(h1)  (kernel.cpp:2466-2480)
            std::vector<uint32_t> et;
            et.reserve((size_t)mesh.number_of_edges() * 4);
            for (edge_array_iterator_t e = mesh.edges_begin(); e != mesh.edges_end(); ++e) {
                const hd_t h0 = mesh.halfedge(*e, 0);
                const hd_t h1 = mesh.halfedge(*e, 1);
                et.push_back((uint32_t)mesh.source(h0));
                et.push_back((uint32_t)mesh.target(h0));
                et.push_back((uint32_t)mesh.face(h0));
                et.push_back((uint32_t)mesh.face(h1));
            }
            mcb::put(g.out, idx_name("dispatch", k, "ps_edges"), et, 4);
            dump_hmesh_arrays(g.out, idx_name("dispatch", k, "ps_"), mesh);
            // per face: edge id of each halfedge around the face
            std::vector<uint32_t> fe;
            for (face_array_iterator_t fi = mesh.faces_begin(); fi != mesh.faces_end(); ++fi) {
                const std::vector<hd_t>& hs = mesh.get_halfedges_around_face(*fi);
                for (const hd_t& h : hs) fe.push_back((uint32_t)mesh.edge(h));
            }
            mcb::put(g.out, idx_name("dispatch", k, "ps_face_edges"), fe);
        }
    } else if (name == "m0.v") {
        bool do_abort = false;
        {
            std::lock_guard<std::mutex> lk(g.mtx);
            const auto t1 = std::chrono::steady_clock::now();
            g.timings_ms.push_back(2.0);
            g.timings_ms.push_back(std::chrono::duration<double, std::milli>(t1 - g.t_ps).count());
            const int k = g.dispatch_calls - 1;
            std::vector<double> ip;
            int idx = 0;
            for (vertex_array_iterator_t v = mesh.vertices_begin(); v != mesh.vertices_end(); ++v, ++idx) {
                if (idx < g.cur_ps_vtx_cnt) continue;
                const vec3& p = mesh.vertex(*v);
                ip.push_back(p.x());
                ip.push_back(p.y());
                ip.push_back(p.z());
            }
            mcb::put(g.out, idx_name("dispatch", k, "ipoints"), ip, 3);
            do_abort = g.abort_after_narrowphase;
        }
        if (do_abort) throw abort_dispatch_t();
    }
}

bool client_input_arrays_to_hmesh(std::shared_ptr<context_t>& context_ptr, McFlags dispatchFlags, hmesh_t& halfedgeMesh,
    const void* pVertices, const McUint32* pFaceIndices, const McUint32* pFaceSizes, const McUint32 numVertices,
    const McUint32 numFaces, const double multiplier, const vec3_<double> srcmesh_cutmesh_com,
    const vec3_<double> pre_quantization_translation, const vec3_<double>* perturbation)
{
    typedef bool (*fn_t)(std::shared_ptr<context_t>&, McFlags, hmesh_t&, const void*, const McUint32*, const McUint32*,
        const McUint32, const McUint32, const double, const vec3_<double>, const vec3_<double>, const vec3_<double>*);
    static fn_t real = real_fn<fn_t>(
        "_Z28client_input_arrays_to_hmeshRSt10shared_ptrI9context_tEjR7hmesh_tPKvPKjS8_jjd5vec3_IdESA_PKSA_");
    {
        std::lock_guard<std::mutex> lk(g.mtx);
        const int k = g.c2h_calls++;
        const double com[3] = { srcmesh_cutmesh_com.x(), srcmesh_cutmesh_com.y(), srcmesh_cutmesh_com.z() };
        const double sh[3] = { pre_quantization_translation.x(), pre_quantization_translation.y(), pre_quantization_translation.z() };
        double pe[3] = { 0, 0, 0 };
        if (perturbation) {
            pe[0] = perturbation->x();
            pe[1] = perturbation->y();
            pe[2] = perturbation->z();
        }
        mcb::put<double>(g.out, idx_name("c2h", k, "com"), com, { 3 });
        mcb::put<double>(g.out, idx_name("c2h", k, "shift"), sh, { 3 });
        mcb::put<double>(g.out, idx_name("c2h", k, "pert"), pe, { 3 });
        mcb::put_scalar<int32_t>(g.out, idx_name("c2h", k, "has_pert"), perturbation ? 1 : 0);
        mcb::put_scalar<uint32_t>(g.out, idx_name("c2h", k, "num_vertices"), numVertices);
    }
    return real(context_ptr, dispatchFlags, halfedgeMesh, pVertices, pFaceIndices, pFaceSizes, numVertices, numFaces, multiplier,
        srcmesh_cutmesh_com, pre_quantization_translation, perturbation);
}

void build_oibvh(thread_pool& pool, const hmesh_t& mesh, std::vector<bbox_t>& bvhAABBs,
    std::vector<fd_t>& bvhLeafNodeFaces, std::vector<bbox_t>& face_bboxes, const double& slightEnlargmentEps,
    const double multiplier)
{
    typedef void (*fn_t)(thread_pool&, const hmesh_t&, std::vector<bbox_t>&, std::vector<fd_t>&,
        std::vector<bbox_t>&, const double&, const double);
    static fn_t real = real_fn<fn_t>(
        "_Z11build_oibvhR11thread_poolRK7hmesh_tRSt6vectorI14bounding_box_tI5vec3_IdEESaIS8_EERS4_I17face_descriptor_tSaISC_EESB_RKdd");
    const auto t0 = std::chrono::steady_clock::now();
    real(pool, mesh, bvhAABBs, bvhLeafNodeFaces, face_bboxes, slightEnlargmentEps, multiplier);
    const auto t1 = std::chrono::steady_clock::now();

    std::lock_guard<std::mutex> lk(g.mtx);
    const int k = g.build_calls++;
    g.timings_ms.push_back(0.0);
    g.timings_ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
    if (g.log_events) dump_hmesh_arrays(g.out, idx_name("build", k, ""), mesh);
    mcb::put_scalar<double>(g.out, idx_name("build", k, "eps"), slightEnlargmentEps);
    std::vector<double> fb;
    fb.reserve(face_bboxes.size() * 6);
    for (const bbox_t& b : face_bboxes) push_bbox(fb, b);
    mcb::put(g.out, idx_name("build", k, "face_bboxes"), fb, 6);
    std::vector<double> root;
    push_bbox(root, bvhAABBs.front());
    mcb::put(g.out, idx_name("build", k, "root_bbox"), root);
    mcb::put_scalar<uint64_t>(g.out, idx_name("build", k, "node_count"), (uint64_t)bvhAABBs.size());
    std::vector<uint32_t> leaves;
    for (const fd_t& f : bvhLeafNodeFaces) leaves.push_back((uint32_t)f);
    mcb::put(g.out, idx_name("build", k, "leaf_faces"), leaves);
}

void intersectOIBVHs(std::map<fd_t, std::vector<fd_t>>& ps_face_to_potentially_intersecting_others,
    const std::vector<bbox_t>& srcMeshBvhAABBs, const std::vector<fd_t>& srcMeshBvhLeafNodeFaces,
    const std::vector<bbox_t>& cutMeshBvhAABBs, const std::vector<fd_t>& cutMeshBvhLeafNodeFaces)
{
    typedef void (*fn_t)(std::map<fd_t, std::vector<fd_t>>&, const std::vector<bbox_t>&, const std::vector<fd_t>&,
        const std::vector<bbox_t>&, const std::vector<fd_t>&);
    static fn_t real = real_fn<fn_t>(
        "_Z15intersectOIBVHsRSt3mapI17face_descriptor_tSt6vectorIS0_SaIS0_EESt4lessIS0_ESaISt4pairIKS0_S3_EEERKS1_I14bounding_box_tI5vec3_IdEESaISF_EERKS3_SJ_SL_");
    const auto t0 = std::chrono::steady_clock::now();
    real(ps_face_to_potentially_intersecting_others, srcMeshBvhAABBs, srcMeshBvhLeafNodeFaces, cutMeshBvhAABBs,
        cutMeshBvhLeafNodeFaces);
    const auto t1 = std::chrono::steady_clock::now();

    std::lock_guard<std::mutex> lk(g.mtx);
    const int k = g.isect_calls++;
    g.timings_ms.push_back(1.0);
    g.timings_ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
    // every (key, value) entry of the map, both directions, as the reference produced them
    std::vector<uint32_t> kv;
    for (const auto& e : ps_face_to_potentially_intersecting_others) {
        std::vector<fd_t> sorted = e.second;
        std::sort(sorted.begin(), sorted.end());
        for (const fd_t& o : sorted) {
            kv.push_back((uint32_t)e.first);
            kv.push_back((uint32_t)o);
        }
    }
    mcb::put(g.out, idx_name("isect", k, "map_entries"), kv, 2);
    mcb::put_scalar<uint32_t>(g.out, idx_name("isect", k, "src_face_count"), (uint32_t)srcMeshBvhLeafNodeFaces.size());
}

void dispatch(output_t& out, const input_t& in)
{
    typedef void (*fn_t)(output_t&, const input_t&);
    static fn_t real = real_fn<fn_t>("_Z8dispatchR8output_tRK7input_t");
    int k;
    {
        std::lock_guard<std::mutex> lk(g.mtx);
        k = g.dispatch_calls++;
        if (g.log_events) {
            dump_hmesh_arrays(g.out, idx_name("dispatch", k, "src_"), *in.src_mesh);
            dump_hmesh_arrays(g.out, idx_name("dispatch", k, "cut_"), *in.cut_mesh);
        }
        mcb::put_scalar<int32_t>(g.out, idx_name("dispatch", k, "gp_count"), in.general_position_enforcement_count);
        // which intersectOIBVHs / build_oibvh results this invocation consumes: the latest ones (a retry after a
        // floating-polygon repartition rebuilds one tree and traverses again, a perturbation retry does neither)
        mcb::put_scalar<int32_t>(g.out, idx_name("dispatch", k, "isect_calls"), g.isect_calls);
        mcb::put_scalar<int32_t>(g.out, idx_name("dispatch", k, "build_calls"), g.build_calls);
        mcb::put_scalar<int32_t>(g.out, idx_name("dispatch", k, "c2h_calls"), g.c2h_calls);
        mcb::put_scalar<uint64_t>(g.out, idx_name("dispatch", k, "event_offset"), (uint64_t)g.events.size());
    }
    bool aborted = false;
    try {
        real(out, in);
    } catch (const abort_dispatch_t&) {
        aborted = true;
    }
    {
        std::lock_guard<std::mutex> lk(g.mtx);
        mcb::put_scalar<int32_t>(g.out, idx_name("dispatch", k, "status"), aborted ? -1000 : (int32_t)out.status.load());
        mcb::put_scalar<uint64_t>(g.out, idx_name("dispatch", k, "event_end"), (uint64_t)g.events.size());
    }
    if (aborted) throw abort_dispatch_t();
}

int compute_polygon_plane_coefficients(vec3& normal, scalar_t& d_coeff, const vec3* polygon_vertices,
    const int polygon_vertex_count, const double multiplier)
{
    typedef int (*fn_t)(vec3&, scalar_t&, const vec3*, const int, const double);
    static fn_t real = real_fn<fn_t>("_Z34compute_polygon_plane_coefficientsR5vec3_IdERdPKS0_id");
    const int r = real(normal, d_coeff, polygon_vertices, polygon_vertex_count, multiplier);
    if (g.log_events) {
        log_event_t e(1);
        e.s(polygon_vertex_count);
        for (int i = 0; i < polygon_vertex_count; ++i) e.v3(polygon_vertices[i]);
        e.v3(normal);
        e.s(d_coeff);
        e.s(r);
        e.commit();
    }
    return r;
}

char compute_segment_plane_intersection_type(const vec3& q, const vec3& r, const std::vector<vec3>& polygon_vertices,
    const vec3& polygon_normal, const int polygon_normal_largest_component, const double multiplier)
{
    typedef char (*fn_t)(const vec3&, const vec3&, const std::vector<vec3>&, const vec3&, const int, const double);
    static fn_t real = real_fn<fn_t>("_Z39compute_segment_plane_intersection_typeRK5vec3_IdES2_RKSt6vectorIS0_SaIS0_EES2_id");
    // the start marker is logged BEFORE the call so nested orient3d/orient2d events follow it
    if (g.log_events) {
        log_event_t e(2);
        e.v3(q);
        e.v3(r);
        e.s((double)polygon_vertices.size());
        for (const vec3& v : polygon_vertices) e.v3(v);
        e.v3(polygon_normal);
        e.s(polygon_normal_largest_component);
        e.commit();
    }
    const char res = real(q, r, polygon_vertices, polygon_normal, polygon_normal_largest_component, multiplier);
    if (g.log_events) {
        log_event_t e(7);
        e.s((double)res);
        e.commit();
    }
    return res;
}

char compute_segment_plane_intersection(vec3& p, const vec3& normal, const scalar_t& d_coeff, const vec3& q, const vec3& r)
{
    typedef char (*fn_t)(vec3&, const vec3&, const scalar_t&, const vec3&, const vec3&);
    static fn_t real = real_fn<fn_t>("_Z34compute_segment_plane_intersectionR5vec3_IdERKS0_RKdS3_S3_");
    const char res = real(p, normal, d_coeff, q, r);
    if (g.log_events) {
        log_event_t e(5);
        e.v3(normal);
        e.s(d_coeff);
        e.v3(q);
        e.v3(r);
        e.v3(p);
        e.s((double)res);
        e.commit();
    }
    return res;
}

char compute_point_in_polygon_test(const vec3& p, const std::vector<vec3>& polygon_vertices, const vec3& polygon_normal,
    const int polygon_normal_largest_component, const double multiplier)
{
    typedef char (*fn_t)(const vec3&, const std::vector<vec3>&, const vec3&, const int, const double);
    static fn_t real = real_fn<fn_t>("_Z29compute_point_in_polygon_testRK5vec3_IdERKSt6vectorIS0_SaIS0_EES2_id");
    const char res = real(p, polygon_vertices, polygon_normal, polygon_normal_largest_component, multiplier);
    if (g.log_events) {
        log_event_t e(6);
        e.v3(p);
        e.s((double)polygon_vertices.size());
        for (const vec3& v : polygon_vertices) e.v3(v);
        e.v3(polygon_normal);
        e.s(polygon_normal_largest_component);
        e.s((double)res);
        e.commit();
    }
    return res;
}

extern "C" double orient3d(const double* pa, const double* pb, const double* pc, const double* pd)
{
    typedef double (*fn_t)(const double*, const double*, const double*, const double*);
    static fn_t real = real_fn<fn_t>("orient3d");
    const double r = real(pa, pb, pc, pd);
    if (g.log_events) {
        log_event_t e(3);
        for (int i = 0; i < 3; ++i) e.s(pa[i]);
        for (int i = 0; i < 3; ++i) e.s(pb[i]);
        for (int i = 0; i < 3; ++i) e.s(pc[i]);
        for (int i = 0; i < 3; ++i) e.s(pd[i]);
        e.s(r);
        e.commit();
    }
    return r;
}

static std::atomic<long> g_adapt_calls(0);
extern "C" double orient3dadapt(const double* pa, const double* pb, const double* pc, const double* pd, double permanent)
{
    typedef double (*fn_t)(const double*, const double*, const double*, const double*, double);
    static fn_t real = real_fn<fn_t>("orient3dadapt");
    g_adapt_calls++;
    return real(pa, pb, pc, pd, permanent);
}

extern "C" double orient2d(const double* pa, const double* pb, const double* pc)
{
    typedef double (*fn_t)(const double*, const double*, const double*);
    static fn_t real = real_fn<fn_t>("orient2d");
    const double r = real(pa, pb, pc);
    if (g.log_events) {
        log_event_t e(4);
        for (int i = 0; i < 2; ++i) e.s(pa[i]);
        for (int i = 0; i < 2; ++i) e.s(pb[i]);
        for (int i = 0; i < 2; ++i) e.s(pc[i]);
        e.s(r);
        e.commit();
    }
    return r;
}

#endif // HARNESS_NO_HOOKS

// ------------------------------------------------------------------------------------------------

static void query_ccs(McContext ctx, mcb::file_t& out)
{
    uint32_t n = 0;
    McResult err = mcGetConnectedComponents(ctx, MC_CONNECTED_COMPONENT_TYPE_ALL, 0, NULL, &n);
    if (err != MC_NO_ERROR) n = 0;
    std::vector<McConnectedComponent> ccs(n);
    if (n) mcGetConnectedComponents(ctx, MC_CONNECTED_COMPONENT_TYPE_ALL, n, ccs.data(), NULL);
    std::vector<uint32_t> types, nvs, nfs, attrs;
    std::vector<double> verts;
    std::vector<uint32_t> faces, sizes;
    for (uint32_t i = 0; i < n; ++i) {
        McConnectedComponentType t = (McConnectedComponentType)0;
        mcGetConnectedComponentData(ctx, ccs[i], MC_CONNECTED_COMPONENT_DATA_TYPE, sizeof(t), &t, NULL);
        types.push_back((uint32_t)t);
        uint32_t a0 = 0, a1 = 0, a2 = 0;
        if (t == MC_CONNECTED_COMPONENT_TYPE_FRAGMENT) {
            mcGetConnectedComponentData(ctx, ccs[i], MC_CONNECTED_COMPONENT_DATA_FRAGMENT_LOCATION, 4, &a0, NULL);
            mcGetConnectedComponentData(ctx, ccs[i], MC_CONNECTED_COMPONENT_DATA_FRAGMENT_SEAL_TYPE, 4, &a1, NULL);
            mcGetConnectedComponentData(ctx, ccs[i], MC_CONNECTED_COMPONENT_DATA_PATCH_LOCATION, 4, &a2, NULL);
        } else if (t == MC_CONNECTED_COMPONENT_TYPE_PATCH) {
            mcGetConnectedComponentData(ctx, ccs[i], MC_CONNECTED_COMPONENT_DATA_PATCH_LOCATION, 4, &a0, NULL);
        } else if (t == MC_CONNECTED_COMPONENT_TYPE_SEAM || t == MC_CONNECTED_COMPONENT_TYPE_INPUT) {
            mcGetConnectedComponentData(ctx, ccs[i], MC_CONNECTED_COMPONENT_DATA_ORIGIN, 4, &a0, NULL);
        }
        attrs.push_back(a0);
        attrs.push_back(a1);
        attrs.push_back(a2);
        McSize nb = 0;
        mcGetConnectedComponentData(ctx, ccs[i], MC_CONNECTED_COMPONENT_DATA_VERTEX_DOUBLE, 0, NULL, &nb);
        std::vector<double> v(nb / sizeof(double));
        if (nb) mcGetConnectedComponentData(ctx, ccs[i], MC_CONNECTED_COMPONENT_DATA_VERTEX_DOUBLE, nb, v.data(), NULL);
        nvs.push_back((uint32_t)(v.size() / 3));
        verts.insert(verts.end(), v.begin(), v.end());
        nb = 0;
        mcGetConnectedComponentData(ctx, ccs[i], MC_CONNECTED_COMPONENT_DATA_FACE, 0, NULL, &nb);
        std::vector<uint32_t> fi(nb / 4);
        if (nb) mcGetConnectedComponentData(ctx, ccs[i], MC_CONNECTED_COMPONENT_DATA_FACE, nb, fi.data(), NULL);
        nb = 0;
        mcGetConnectedComponentData(ctx, ccs[i], MC_CONNECTED_COMPONENT_DATA_FACE_SIZE, 0, NULL, &nb);
        std::vector<uint32_t> fs(nb / 4);
        if (nb) mcGetConnectedComponentData(ctx, ccs[i], MC_CONNECTED_COMPONENT_DATA_FACE_SIZE, nb, fs.data(), NULL);
        nfs.push_back((uint32_t)fs.size());
        faces.insert(faces.end(), fi.begin(), fi.end());
        sizes.insert(sizes.end(), fs.begin(), fs.end());
    }
    mcb::put(out, "cc_type", types);
    mcb::put(out, "cc_attrs", attrs, 3);
    mcb::put(out, "cc_nv", nvs);
    mcb::put(out, "cc_nf", nfs);
    mcb::put(out, "cc_vertices", verts, 3);
    mcb::put(out, "cc_faces", faces);
    mcb::put(out, "cc_face_sizes", sizes);
    if (n) mcReleaseConnectedComponents(ctx, n, ccs.data());
}

int main(int argc, char** argv)
{
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s <in.mcb> <out.mcb> [--helpers N] [--no-events] [--abort-after-narrowphase] [--no-cc] [--repeat R] [--planar nx ny nz off]\n", argv[0]);
        return 2;
    }
    int helpers = 0, repeat = 1, contexts = 1;
    bool want_cc = true, planar = false;
    double pn[3] = { 0, 0, 1 }, poff = 0.5;
    for (int i = 3; i < argc; ++i) {
        const std::string a(argv[i]);
        if (a == "--helpers" && i + 1 < argc) helpers = std::atoi(argv[++i]);
        else if (a == "--no-events") g.log_events = false;
        else if (a == "--abort-after-narrowphase") g.abort_after_narrowphase = true;
        else if (a == "--no-cc") want_cc = false;
        else if (a == "--repeat" && i + 1 < argc) repeat = std::atoi(argv[++i]);
        else if (a == "--contexts" && i + 1 < argc) contexts = std::atoi(argv[++i]);
        else if (a == "--planar" && i + 4 < argc) {
            planar = true;
            pn[0] = std::atof(argv[++i]);
            pn[1] = std::atof(argv[++i]);
            pn[2] = std::atof(argv[++i]);
            poff = std::atof(argv[++i]);
        } else {
            std::fprintf(stderr, "unknown arg %s\n", a.c_str());
            return 2;
        }
    }
    mcb::file_t in = mcb::read(argv[1]);
    const mcb::array_t& sx = in.at("src_xyz");
    const uint32_t flags = in.at("flags").as<uint32_t>()[0];
    const uint32_t nsv = (uint32_t)sx.dims[0];
    const uint32_t* ssz = in.count("src_sizes") ? in.at("src_sizes").as<uint32_t>() : NULL;
    const uint32_t nsf = ssz ? (uint32_t)in.at("src_sizes").count() : (uint32_t)(in.at("src_faces").count() / 3);

    if (contexts > 1) {
        // The MultipleContextsInParallel pattern (tutorials/MultipleContextsInParallel, tests/source/
        // concurrentSynchronizedContexts.cpp): N contexts, each dispatching the same input from its own thread at the same
        // time.  Every context must come back with the same connected components; the first one's are written out.
        const mcb::array_t& cx = in.at("cut_xyz");
        const uint32_t ncv = (uint32_t)cx.dims[0];
        const uint32_t* csz = in.count("cut_sizes") ? in.at("cut_sizes").as<uint32_t>() : NULL;
        const uint32_t ncf = csz ? (uint32_t)in.at("cut_sizes").count() : (uint32_t)(in.at("cut_faces").count() / 3);
        std::vector<mcb::file_t> outs((size_t)contexts);
        std::vector<int32_t> results((size_t)contexts, -1000);
        std::vector<std::thread> th;
        for (int c = 0; c < contexts; ++c)
            th.emplace_back([&, c]() {
                McContext ctx = MC_NULL_HANDLE;
                if (mcCreateContextWithHelpers(&ctx, MC_NULL_HANDLE, (uint32_t)helpers) != MC_NO_ERROR) return;
                const McResult err = mcDispatch(ctx, flags, sx.bytes.data(), in.at("src_faces").as<uint32_t>(), ssz, nsv, nsf,
                    cx.bytes.data(), in.at("cut_faces").as<uint32_t>(), csz, ncv, ncf);
                results[(size_t)c] = (int32_t)err;
                if (err == MC_NO_ERROR) query_ccs(ctx, outs[(size_t)c]);
                mcReleaseContext(ctx);
            });
        for (std::thread& t : th) t.join();
        int identical = 1;
        for (int c = 1; c < contexts; ++c) {
            if (results[(size_t)c] != results[0] || outs[(size_t)c].size() != outs[0].size()) identical = 0;
            for (const auto& kv : outs[0]) {
                const auto it = outs[(size_t)c].find(kv.first);
                if (it == outs[(size_t)c].end() || it->second.bytes != kv.second.bytes) identical = 0;
            }
        }
        for (const auto& kv : outs[0]) g.out[kv.first] = kv.second;
        mcb::put_scalar<int32_t>(g.out, "mcDispatch_result", results[0]);
        mcb::put_scalar<int32_t>(g.out, "contexts", contexts);
        mcb::put_scalar<int32_t>(g.out, "contexts_identical", identical);
        mcb::put(g.out, "contexts_results", results);
        mcb::write(argv[2], g.out);
        std::fprintf(stderr, "harness: %d contexts in parallel, identical=%d, mcDispatch=%d\n", contexts, identical, results[0]);
        return 0;
    }

    int rc = 0;
    for (int rep = 0; rep < repeat; ++rep) {
        McContext ctx = MC_NULL_HANDLE;
        McResult err = mcCreateContextWithHelpers(&ctx, MC_NULL_HANDLE, (uint32_t)helpers);
        if (err != MC_NO_ERROR) {
            std::fprintf(stderr, "mcCreateContextWithHelpers failed %d\n", (int)err);
            return 1;
        }
        const auto t0 = std::chrono::steady_clock::now();
        if (planar) {
            McEvent ev = MC_NULL_HANDLE;
            err = mcEnqueueDispatchPlanarSection(ctx, flags, sx.bytes.data(), in.at("src_faces").as<uint32_t>(), ssz, nsv,
                nsf, pn, poff, 0, NULL, &ev);
            if (err == MC_NO_ERROR) {
                mcWaitForEvents(1, &ev);
                McResult st = MC_NO_ERROR;
                mcGetEventInfo(ev, MC_EVENT_RUNTIME_EXECUTION_STATUS, sizeof(McResult), &st, NULL);
                err = st;
                mcReleaseEvents(1, &ev);
            }
        } else {
            const mcb::array_t& cx = in.at("cut_xyz");
            const uint32_t ncv = (uint32_t)cx.dims[0];
            const uint32_t* csz = in.count("cut_sizes") ? in.at("cut_sizes").as<uint32_t>() : NULL;
            const uint32_t ncf = csz ? (uint32_t)in.at("cut_sizes").count() : (uint32_t)(in.at("cut_faces").count() / 3);
            err = mcDispatch(ctx, flags, sx.bytes.data(), in.at("src_faces").as<uint32_t>(), ssz, nsv, nsf, cx.bytes.data(),
                in.at("cut_faces").as<uint32_t>(), csz, ncv, ncf);
        }
        const auto t1 = std::chrono::steady_clock::now();
        rc = (int)err;
        if (rep == repeat - 1) {
            mcb::put_scalar<int32_t>(g.out, "mcDispatch_result", (int32_t)err);
            mcb::put_scalar<double>(g.out, "mcDispatch_ms", std::chrono::duration<double, std::milli>(t1 - t0).count());
            {
                // MC_DISPATCH_INCLUDE_INTERSECTION_TYPE: what check_and_store_input_mesh_intersection_type() decided
                McDispatchIntersectionType it = MC_DISPATCH_INTERSECTION_TYPE_MAX_ENUM;
                if (mcGetInfo(ctx, MC_CONTEXT_DISPATCH_INTERSECTION_TYPE, sizeof(it), &it, NULL) != MC_NO_ERROR)
                    it = MC_DISPATCH_INTERSECTION_TYPE_MAX_ENUM;
                mcb::put_scalar<uint32_t>(g.out, "intersection_type", (uint32_t)it);
            }
            if (want_cc && !g.abort_after_narrowphase) query_ccs(ctx, g.out);
        }
        mcReleaseContext(ctx);
    }
    mcb::put_scalar<int32_t>(g.out, "build_calls", g.build_calls);
    mcb::put_scalar<int32_t>(g.out, "isect_calls", g.isect_calls);
    mcb::put_scalar<int32_t>(g.out, "dispatch_calls", g.dispatch_calls);
    mcb::put_scalar<int32_t>(g.out, "c2h_calls", g.c2h_calls);
#ifndef HARNESS_NO_HOOKS
    mcb::put_scalar<int64_t>(g.out, "orient3dadapt_calls", (int64_t)g_adapt_calls.load());
#endif
    mcb::put(g.out, "events", g.events);
    mcb::put(g.out, "timings_ms", g.timings_ms, 2);
    mcb::write(argv[2], g.out);
    std::fprintf(stderr, "harness: mcDispatch=%d builds=%d isects=%d dispatches=%d events=%zu\n", rc, g.build_calls, g.isect_calls,
        g.dispatch_calls, g.events.size());
    return 0;
}
