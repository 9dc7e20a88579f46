// mcut_b200/csrc/shim/mcut_shim.cpp — the reference-facing adapter: the ONLY translation unit that knows the
// reference's C++ types.  It defines the two broadphase functions of the reference with their exact signatures
//
//     build_oibvh()      include/mcut/internal/bvh.h:117-125   (source/bvh.cpp:219-636)
//     intersectOIBVHs()  include/mcut/internal/bvh.h:127-133   (source/bvh.cpp:638-783)
//
// and forwards them to the C-ABI of include/mcut_b200.h.  Both are called through the PLT inside libmcut.so
// (SURVEY.md §8-b), so loading this library ahead of libmcut.so (LD_PRELOAD, or link order) makes every unmodified
// mcDispatch / mcEnqueueDispatch / mcEnqueueDispatchPlanarSection run its broadphase on the B200 — the public C API of
// include/mcut/mcut.h is untouched.  There is no CPU fallback: if the device layer fails, a std::runtime_error is thrown,
// which the reference's own CATCH_POSSIBLE_EXCEPTIONS (include/mcut/internal/frontend.h:69-89) turns into
// MC_INVALID_OPERATION.
//
// What crosses back to the host is exactly what build_oibvh's callers consume: face_bboxes (kernel cull step,
// kernel.cpp:2086-2105), bvhAABBs[0] (preproc.cpp:2896-2897) and — from intersectOIBVHs — the candidate map.  The device
// tree itself stays resident and is found again through a side table keyed by the address of the caller's bvhAABBs
// vector.  Built only where the reference's headers exist (see ../Makefile: target shim); it contains no reference code.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include <dlfcn.h>

#include "mcut/mcut.h"
#include "mcut/internal/bvh.h"
#include "mcut/internal/frontend.h"
#include "mcut/internal/hmesh.h"
#include "mcut/internal/math.h"

#include "../../../include/mcut_b200.h"
#include "mcut_hook.h"
#include "hook_fill.h"

namespace {

typedef bounding_box_t<vec3_<double>> bbox_t;

// The user arrays a half-edge mesh was built from (client_input_arrays_to_hmesh, source/preproc.cpp:57-468, interposed
// below).  mcDispatch only lends them for the duration of the dispatch, which is exactly as long as they are needed here.
// With them the adapter never walks a half-edge mesh: the device reads the user's arrays and applies the frame itself.
struct capture_t {
    const void* xyz = nullptr;
    const McUint32* faces = nullptr;
    const McUint32* sizes = nullptr;
    uint32_t nv = 0, nf = 0;
    int is_float = 0;
    double com[3] = { 0, 0, 0 }, shift[3] = { 0, 0, 0 }, pert[3] = { 0, 0, 0 };
    bool has_pert = false;
};
struct device_tree_t {
    mcb200_ctx* ctx = nullptr;
    mcb200_mesh* mesh = nullptr;
    uint32_t nf = 0;
    bool from_arrays = false;
    capture_t cap;
};

// What the last intersectOIBVHs() of this API thread worked with: dispatch() runs on the same thread right after it (and
// again, with new cut coordinates, on every general-position retry) and its narrowphase hook continues from here.
thread_local std::unordered_map<const hmesh_t*, capture_t> t_captures; // by address of the half-edge mesh
thread_local capture_t t_latest_capture; // the most recent conversion on this API thread (a retry's perturbed cut mesh)

struct last_intersect_t {
    mcb200_ctx* ctx = nullptr;
    mcb200_mesh* src = nullptr;
    mcb200_mesh* cut = nullptr;
    mcb200_result* res = nullptr;
    mcb200_soup* soup = nullptr; // device-numbered polygon soup of (src, cut), made on first use
    bool from_arrays = false; // both meshes came from captured user arrays: the hook takes the fast path
    bool src_from_arrays = false, cut_from_arrays = false; // each mesh on its own (a repartition retry rebuilds only one)
    capture_t src_cap, cut_cap;
};
thread_local last_intersect_t t_last;

std::mutex g_mutex;
std::unordered_map<const void*, device_tree_t> g_trees; // key: address of the caller's bvhAABBs vector
std::vector<mcb200_ctx*> g_idle_ctx; // device contexts whose API thread has ended, warm buffers and all

// One device context per API thread of the reference (every MCUT context owns its API thread, frontend.h:548-584), so
// concurrently dispatching contexts never share a stream.  Device: MCB200_DEVICE, else threads are dealt round-robin.
// When the API thread ends (mcReleaseContext) its device objects are released and the device context is parked for the next
// API thread: a program that creates an MCUT context per dispatch keeps its reserved device buffers instead of mapping a few
// hundred MB anew each time.
struct early_mesh_t { // a device mesh made for the input checks, before build_oibvh asked for one (adopted there)
    mcb200_mesh* mesh = nullptr;
    capture_t cap;
};
struct thread_state_t {
    mcb200_ctx* ctx = nullptr;
    std::deque<const void*> recent; // keys of the trees this thread built, oldest first
    std::unordered_map<const hmesh_t*, early_mesh_t> early;
    ~thread_state_t()
    {
        if (!ctx) return;
        for (auto& kv : early) mcb200_mesh_free(ctx, kv.second.mesh);
        early.clear();
        if (t_last.res) mcb200_result_free(t_last.ctx, t_last.res);
        if (t_last.soup) mcb200_soup_free(t_last.ctx, t_last.soup);
        t_last = last_intersect_t();
        std::lock_guard<std::mutex> lk(g_mutex);
        for (const void* key : recent) {
            auto it = g_trees.find(key);
            if (it == g_trees.end() || it->second.ctx != ctx) continue;
            mcb200_mesh_free(ctx, it->second.mesh);
            g_trees.erase(it);
        }
        g_idle_ctx.push_back(ctx);
    }
};
thread_local thread_state_t t_state;

mcb200_ctx* thread_ctx()
{
    if (!t_state.ctx) {
        static int next_device = 0;
        int device = -1;
        if (const char* e = std::getenv("MCB200_DEVICE")) device = std::atoi(e);
        {
            std::lock_guard<std::mutex> lk(g_mutex);
            if (!g_idle_ctx.empty()) { // (with MCB200_DEVICE every context of the process is on that device)
                t_state.ctx = g_idle_ctx.back();
                g_idle_ctx.pop_back();
                return t_state.ctx;
            }
            if (device < 0) {
                const int n = mcb200_device_count();
                device = n > 0 ? next_device++ % n : 0;
            }
        }
        const int rc = mcb200_ctx_create(device, nullptr, &t_state.ctx);
        if (rc != 0) throw std::runtime_error(std::string("mcut_b200: ") + mcb200_last_error(nullptr));
    }
    return t_state.ctx;
}

typedef mcb200_scope_timer scope_timer; // MCB200_SHIM_TIMING=1 (hook_fill.h)

void check(mcb200_ctx* ctx, int rc, const char* what)
{
    if (rc == 0) return;
    const std::string msg = std::string("mcut_b200: ") + what + ": " + mcb200_last_error(ctx);
    // the reference turns the exception into MC_INVALID_OPERATION and drops its text: keep it visible
    std::fprintf(stderr, "%s\n", msg.c_str());
    throw std::runtime_error(msg);
}

// internal coordinate of user vertex v exactly as client_input_arrays_to_hmesh computes it (preproc.cpp:91-185)
void captured_vertex(const capture_t& c, uint32_t v, double out[3])
{
    for (int j = 0; j < 3; ++j) {
        if (c.is_float) {
            const float x = (static_cast<const float*>(c.xyz)[3 * (size_t)v + j] - (float)c.com[j]) + (float)c.shift[j];
            out[j] = double(x) + (c.has_pert ? c.pert[j] : double(0.));
        } else {
            const double x = (static_cast<const double*>(c.xyz)[3 * (size_t)v + j] - c.com[j]) + c.shift[j];
            out[j] = x + (c.has_pert ? c.pert[j] : double(0.));
        }
    }
}

// Is `mesh` still what the capture says it is?  (Same counts, and three of its vertices bit for bit: a half-edge mesh that
// was modified after the conversion, or another mesh living at a recycled address, takes the generic path.)
bool capture_matches(const capture_t& c, const hmesh_t& mesh)
{
    if (!c.xyz || (uint32_t)mesh.number_of_vertices() != c.nv || (uint32_t)mesh.number_of_faces() != c.nf) return false;
    const uint32_t probe[3] = { 0u, c.nv / 2u, c.nv - 1u };
    for (uint32_t v : probe) {
        double want[3];
        captured_vertex(c, v, want);
        const vec3& p = mesh.vertex(vd_t(v));
        if (p.x() != want[0] || p.y() != want[1] || p.z() != want[2]) return false;
    }
    return true;
}

bool same_capture(const capture_t& a, const capture_t& b)
{
    if (a.xyz != b.xyz || a.faces != b.faces || a.sizes != b.sizes || a.nv != b.nv || a.nf != b.nf || a.is_float != b.is_float
        || a.has_pert != b.has_pert)
        return false;
    for (int j = 0; j < 3; ++j)
        if (a.com[j] != b.com[j] || a.shift[j] != b.shift[j] || a.pert[j] != b.pert[j]) return false;
    return true;
}

void drop_early_mesh(const hmesh_t* key)
{
    auto it = t_state.early.find(key);
    if (it == t_state.early.end()) return;
    mcb200_mesh_free(t_state.ctx, it->second.mesh);
    t_state.early.erase(it);
}

// The device copy of a half-edge mesh that is still exactly what the user's arrays say (or nullptr): made on first use,
// handed over to build_oibvh when the reference gets there.
mcb200_mesh* early_device_mesh(const hmesh_t& mesh, const capture_t** cap_out)
{
    auto cap = t_captures.find(&mesh);
    if (cap == t_captures.end() || !capture_matches(cap->second, mesh)) return nullptr;
    const capture_t& c = cap->second;
    if (cap_out) *cap_out = &c;
    auto it = t_state.early.find(&mesh);
    if (it != t_state.early.end()) {
        if (same_capture(it->second.cap, c)) return it->second.mesh;
        drop_early_mesh(&mesh);
    }
    mcb200_ctx* ctx = thread_ctx();
    early_mesh_t e;
    e.cap = c;
    check(ctx, mcb200_mesh_create_trusted(ctx, c.is_float, c.xyz, c.nv, c.faces, c.sizes, c.nf, &e.mesh), "mesh_create");
    check(ctx, mcb200_mesh_set_frame(ctx, e.mesh, c.com, c.shift, c.has_pert ? c.pert : nullptr), "set_frame");
    t_state.early[&mesh] = e;
    return e.mesh;
}

} // namespace

// source/preproc.cpp:57-468: called through the PLT by preproc() for the source mesh (:2338) and for the cut mesh of every
// general-position attempt (:2650).  The reference's own function does the work; the adapter only remembers the arguments.
bool client_input_arrays_to_hmesh(std::shared_ptr<context_t>& context_ptr, McFlags dispatchFlags, hmesh_t& halfedgeMesh,
    const void* pVertices, const McUint32* pFaceIndices, const McUint32* pFaceSizes, const McUint32 numVertices,
    const McUint32 numFaces, const double multiplier, const vec3_<double> srcmesh_cutmesh_com,
    const vec3_<double> pre_quantization_translation, const vec3_<double>* perturbation)
{
    typedef bool (*fn_t)(std::shared_ptr<context_t>&, McFlags, hmesh_t&, const void*, const McUint32*, const McUint32*,
        const McUint32, const McUint32, const double, const vec3_<double>, const vec3_<double>, const vec3_<double>*);
    static fn_t real = reinterpret_cast<fn_t>(
        dlsym(RTLD_NEXT, "_Z28client_input_arrays_to_hmeshRSt10shared_ptrI9context_tEjR7hmesh_tPKvPKjS8_jjd5vec3_IdESA_PKSA_"));
    if (!real) throw std::runtime_error("mcut_b200: the reference's client_input_arrays_to_hmesh was not found");
    const bool ok = real(context_ptr, dispatchFlags, halfedgeMesh, pVertices, pFaceIndices, pFaceSizes, numVertices, numFaces,
        multiplier, srcmesh_cutmesh_com, pre_quantization_translation, perturbation);
    capture_t c;
    if (ok) {
        c.xyz = pVertices;
        c.faces = pFaceIndices;
        c.sizes = pFaceSizes;
        c.nv = numVertices;
        c.nf = numFaces;
        c.is_float = (dispatchFlags & MC_DISPATCH_VERTEX_ARRAY_FLOAT) ? 1 : 0;
        for (int j = 0; j < 3; ++j) {
            c.com[j] = srcmesh_cutmesh_com[j];
            c.shift[j] = pre_quantization_translation[j];
            c.pert[j] = perturbation ? (*perturbation)[j] : 0.0;
        }
        c.has_pert = perturbation != nullptr;
    }
    if (t_state.ctx) drop_early_mesh(&halfedgeMesh); // whatever lived at this address before is gone
    t_captures[&halfedgeMesh] = c; // a failed conversion leaves an empty capture: generic path
    t_latest_capture = c;
    return ok;
}

// ---------------------------------------------------------------------------------------------------------------------
// Input validation, PLT-called by preproc(): answered from the device copy of the user's arrays.
// MCB200_SHIM_HOST_CHECKS=1 leaves all three to the reference.
// ---------------------------------------------------------------------------------------------------------------------
static bool host_checks()
{
    static const bool v = std::getenv("MCB200_SHIM_HOST_CHECKS") != nullptr;
    return v;
}

// source/preproc.cpp:505-578 (called at :2353, :2786, :2797).  Triangle meshes only: the coplanarity warning loop (:552-575)
// has nothing to look at then; the vertex / face count checks and polygon meshes stay with the reference's function.
bool check_input_mesh(std::shared_ptr<context_t>& context_ptr, const hmesh_t& m)
{
    typedef bool (*fn_t)(std::shared_ptr<context_t>&, const hmesh_t&);
    static fn_t real = reinterpret_cast<fn_t>(dlsym(RTLD_NEXT, "_Z16check_input_meshRSt10shared_ptrI9context_tERK7hmesh_t"));
    if (!real) throw std::runtime_error("mcut_b200: the reference's check_input_mesh was not found");
    if (host_checks() || m.number_of_vertices() < 3 || m.number_of_faces() < 1) return real(context_ptr, m);
    const capture_t* c = nullptr;
    mcb200_mesh* dm = early_device_mesh(m, &c);
    if (!dm || c->sizes != nullptr) return real(context_ptr, m);
    scope_timer timer("check_input_mesh (device)");
    mcb200_validation v;
    if (mcb200_mesh_validate(t_state.ctx, dm, &v) != 0) return real(context_ptr, m);
    if (v.n_components != 1) { // :541-550
        context_ptr->dbg_cb(MC_DEBUG_SOURCE_API, MC_DEBUG_TYPE_ERROR, 0, MC_DEBUG_SEVERITY_HIGH,
            "Detected multiple connected components in mesh (N=" + std::to_string(v.n_components) + ")");
        return false;
    }
    return true;
}

// source/preproc.cpp:1957-1990 (called at :2805, :2814): no border edge
bool mesh_is_closed(const hmesh_t& mesh)
{
    typedef bool (*fn_t)(const hmesh_t&);
    static fn_t real = reinterpret_cast<fn_t>(dlsym(RTLD_NEXT, "_Z14mesh_is_closedRK7hmesh_t"));
    if (!real) throw std::runtime_error("mcut_b200: the reference's mesh_is_closed was not found");
    if (host_checks()) return real(mesh);
    // only the copy check_input_mesh just made, or the tree build_oibvh holds, is consulted: no upload for this question alone
    mcb200_mesh* dm = nullptr;
    auto early = t_state.early.find(&mesh);
    auto cap = t_captures.find(&mesh);
    if (early != t_state.early.end() && cap != t_captures.end() && same_capture(early->second.cap, cap->second) && capture_matches(cap->second, mesh))
        dm = early->second.mesh;
    if (!dm && cap != t_captures.end() && capture_matches(cap->second, mesh)) {
        std::lock_guard<std::mutex> lk(g_mutex);
        for (const void* key : t_state.recent) {
            auto it = g_trees.find(key);
            if (it != g_trees.end() && it->second.from_arrays && it->second.ctx == t_state.ctx && same_capture(it->second.cap, cap->second))
                dm = it->second.mesh;
        }
    }
    if (!dm) return real(mesh);
    scope_timer timer("mesh_is_closed (device)");
    mcb200_validation v;
    if (mcb200_mesh_validate(t_state.ctx, dm, &v) != 0) return real(mesh);
    return v.is_closed != 0;
}

// source/preproc.cpp:1999-2122 (called at :2891 when the BVHs do not overlap and at :3716 when the surfaces do not cut each
// other): the inside / outside verdict for MC_DISPATCH_INCLUDE_INTERSECTION_TYPE, from the two device trees of this dispatch.
void check_and_store_input_mesh_intersection_type(std::shared_ptr<context_t>& context_ptr, const std::shared_ptr<hmesh_t>& source_hmesh,
    const std::shared_ptr<hmesh_t>& cut_hmesh, const bool sm_is_watertight, const bool cm_is_watertight,
    const bounding_box_t<vec3_<double>>& sm_aabb, const bounding_box_t<vec3_<double>>& cm_aabb, const double multiplier)
{
    typedef void (*fn_t)(std::shared_ptr<context_t>&, const std::shared_ptr<hmesh_t>&, const std::shared_ptr<hmesh_t>&, const bool,
        const bool, const bounding_box_t<vec3_<double>>&, const bounding_box_t<vec3_<double>>&, const double);
    static fn_t real = reinterpret_cast<fn_t>(dlsym(RTLD_NEXT,
        "_Z44check_and_store_input_mesh_intersection_typeRSt10shared_ptrI9context_tERKS_I7hmesh_tES6_bbRK14bounding_box_tI5vec3_IdEESC_d"));
    if (!real) throw std::runtime_error("mcut_b200: the reference's check_and_store_input_mesh_intersection_type was not found");
    // the two trees the last intersectOIBVHs of this thread worked with are this dispatch's meshes as they are NOW (the cut
    // mesh of the last general-position attempt); both must be the user's own arrays on the device
    const bool usable = !host_checks() && t_last.ctx && t_last.src && t_last.cut && t_last.src_from_arrays && t_last.cut_from_arrays
        && t_last.src_cap.nv == (uint32_t)source_hmesh->number_of_vertices() && t_last.src_cap.nf == (uint32_t)source_hmesh->number_of_faces()
        && t_last.cut_cap.nv == (uint32_t)cut_hmesh->number_of_vertices() && t_last.cut_cap.nf == (uint32_t)cut_hmesh->number_of_faces();
    if (usable) {
        scope_timer timer("intersection type (device)");
        uint32_t type = 0;
        if (mcb200_intersection_type_without_cut(t_last.ctx, t_last.src, t_last.cut, &type) == 0) {
            context_ptr->set_most_recent_dispatch_intersection_type((McDispatchIntersectionType)type);
            return;
        } // (faces with more than four vertices need the reference's CDT: the host function answers)
    }
    real(context_ptr, source_hmesh, cut_hmesh, sm_is_watertight, cm_is_watertight, sm_aabb, cm_aabb, multiplier);
}

void build_oibvh(thread_pool& /*pool*/, const hmesh_t& mesh, std::vector<bbox_t>& bvhAABBs, std::vector<fd_t>& bvhLeafNodeFaces,
    std::vector<bbox_t>& face_bboxes, const double& slightEnlargmentEps, const double /*multiplier*/)
{
    scope_timer timer("build_oibvh");
    mcb200_ctx* ctx = thread_ctx();
    {
        // the caller's vector is being rebuilt (next dispatch / next attempt): release the previous tree FIRST, so that
        // the new one is carved out of the blocks the memory pool just got back instead of freshly mapped memory
        std::lock_guard<std::mutex> lk(g_mutex);
        auto it = g_trees.find(&bvhAABBs);
        if (it != g_trees.end()) {
            if (t_last.src == it->second.mesh || t_last.cut == it->second.mesh) { // the hook state refers to it
                if (t_last.res) mcb200_result_free(t_last.ctx, t_last.res);
                if (t_last.soup) mcb200_soup_free(t_last.ctx, t_last.soup);
                t_last = last_intersect_t();
            }
            mcb200_mesh_free(it->second.ctx, it->second.mesh);
            g_trees.erase(it);
        }
    }
    device_tree_t t;
    t.ctx = ctx;
    const uint32_t nf = (uint32_t)mesh.number_of_faces();
    t.nf = nf;
    auto cap = t_captures.find(&mesh);
    if (cap != t_captures.end() && capture_matches(cap->second, mesh)) {
        // fast path: the device reads the user's own arrays and applies the frame itself — no walk over the half-edge mesh
        // (the conversion itself has range-checked every index, preproc.cpp:271 / :417, and returned true)
        const capture_t& c = cap->second;
        auto early = t_state.early.find(&mesh);
        if (early != t_state.early.end() && same_capture(early->second.cap, c)) {
            t.mesh = early->second.mesh; // the input checks have uploaded it already (check_input_mesh below)
            t_state.early.erase(early);
        } else {
            scope_timer t2("  build_oibvh: upload of the user's arrays");
            check(ctx, mcb200_mesh_create_trusted(ctx, c.is_float, c.xyz, c.nv, c.faces, c.sizes, c.nf, &t.mesh), "mesh_create");
            check(ctx, mcb200_mesh_set_frame(ctx, t.mesh, c.com, c.shift, c.has_pert ? c.pert : nullptr), "set_frame");
        }
        t.from_arrays = true;
        t.cap = c;
    } else {
        // generic path: flatten the half-edge mesh (internal coordinates: the reference has already re-centred them)
        uint32_t nv = 0;
        for (vertex_array_iterator_t v = mesh.vertices_begin(); v != mesh.vertices_end(); ++v)
            if ((uint32_t)*v + 1u > nv) nv = (uint32_t)*v + 1u;
        std::vector<double> xyz(3 * (size_t)nv, 0.0);
        for (vertex_array_iterator_t v = mesh.vertices_begin(); v != mesh.vertices_end(); ++v) {
            const vec3& p = mesh.vertex(*v);
            xyz[3 * (size_t)(uint32_t)*v + 0] = p.x();
            xyz[3 * (size_t)(uint32_t)*v + 1] = p.y();
            xyz[3 * (size_t)(uint32_t)*v + 2] = p.z();
        }
        std::vector<uint32_t> sizes, idx;
        sizes.reserve(nf);
        idx.reserve(3 * (size_t)nf);
        std::vector<vd_t> tmp;
        for (face_array_iterator_t f = mesh.faces_begin(); f != mesh.faces_end(); ++f) {
            mesh.get_vertices_around_face(tmp, *f);
            sizes.push_back((uint32_t)tmp.size());
            for (const vd_t& v : tmp) idx.push_back((uint32_t)v);
        }
        check(ctx, mcb200_mesh_create(ctx, 0, xyz.data(), nv, idx.data(), sizes.data(), nf, &t.mesh), "mesh_create");
        check(ctx, mcb200_mesh_set_frame(ctx, t.mesh, nullptr, nullptr, nullptr), "set_frame");
    }
    if (!face_bboxes.empty()) {
        // `face_bboxes` is in/out in the reference: build_oibvh resizes it and EXPANDS what is there (bvh.cpp:242-272), and
        // preproc.cpp keeps one vector per mesh for a whole mcDispatch, so the rebuild after a floating-polygon repartition
        // (preproc.cpp:2733-2760) starts from the previous boxes.  Same here.  (With the hooked kernel the boxes never leave
        // the device and the vector arrives empty: that rebuild then gets the tight boxes — fewer candidates, same result.)
        const size_t n = face_bboxes.size() < (size_t)nf ? face_bboxes.size() : (size_t)nf;
        std::vector<double> prior(6 * n);
        for (size_t f = 0; f < n; ++f) {
            const vec3_<double>&lo = face_bboxes[f].minimum(), &hi = face_bboxes[f].maximum();
            double* b = prior.data() + 6 * f;
            b[0] = lo.x(), b[1] = lo.y(), b[2] = lo.z(), b[3] = hi.x(), b[4] = hi.y(), b[5] = hi.z();
        }
        check(ctx, mcb200_mesh_set_prior_face_boxes(ctx, t.mesh, prior.data(), (uint32_t)n), "set_prior_face_boxes");
    }
    check(ctx, mcb200_bvh_build(ctx, t.mesh, slightEnlargmentEps), "bvh_build");

    // face_bboxes are read by the kernel's cull step only (kernel.cpp:2086-2160); the hooked kernel has no such step and
    // says so with a marker symbol, in which case the boxes stay on the device
    static const bool kernel_is_hooked = dlsym(RTLD_DEFAULT, "mcb200_kernel_is_hooked") != nullptr;
    double root[6];
    if (kernel_is_hooked) {
        check(ctx, mcb200_bvh_read(ctx, t.mesh, nullptr, root), "bvh_read");
        face_bboxes.clear();
    } else {
        std::vector<double> boxes(6 * (size_t)nf);
        check(ctx, mcb200_bvh_read(ctx, t.mesh, boxes.data(), root), "bvh_read");
        face_bboxes.resize(nf);
        for (uint32_t f = 0; f < nf; ++f) {
            const double* b = boxes.data() + 6 * (size_t)f;
            face_bboxes[f] = bbox_t(vec3_<double>(b[0], b[1], b[2]), vec3_<double>(b[3], b[4], b[5]));
        }
    }
    // callers read bvhAABBs[0] (the mesh AABB) only (preproc.cpp:2896-2897, :3721-3722); bvhLeafNodeFaces is opaque to
    // everyone but intersectOIBVHs, which uses the device tree instead
    bvhAABBs.assign(1, bbox_t(vec3_<double>(root[0], root[1], root[2]), vec3_<double>(root[3], root[4], root[5])));
    bvhLeafNodeFaces.clear();

    std::lock_guard<std::mutex> lk(g_mutex);
    g_trees[&bvhAABBs] = t;
    // A dispatch uses two trees.  The caller's vectors usually live at the same addresses from one dispatch to the next
    // (then the tree was released above); if they do not, trees of this thread beyond the four most recent are released.
    t_state.recent.push_back(&bvhAABBs);
    while (t_state.recent.size() > 4) {
        const void* old = t_state.recent.front();
        t_state.recent.pop_front();
        auto it = g_trees.find(old);
        if (it == g_trees.end() || old == static_cast<const void*>(&bvhAABBs)) continue;
        if (it->second.mesh == t_last.src || it->second.mesh == t_last.cut) continue; // still needed by the hook
        mcb200_mesh_free(it->second.ctx, it->second.mesh);
        g_trees.erase(it);
    }
}

void intersectOIBVHs(std::map<fd_t, std::vector<fd_t>>& ps_face_to_potentially_intersecting_others,
    const std::vector<bbox_t>& srcMeshBvhAABBs, const std::vector<fd_t>& srcMeshBvhLeafNodeFaces,
    const std::vector<bbox_t>& cutMeshBvhAABBs, const std::vector<fd_t>& /*cutMeshBvhLeafNodeFaces*/)
{
    scope_timer timer("intersectOIBVHs");
    device_tree_t s, c;
    {
        std::lock_guard<std::mutex> lk(g_mutex);
        auto is = g_trees.find(&srcMeshBvhAABBs), ic = g_trees.find(&cutMeshBvhAABBs);
        if (is == g_trees.end() || ic == g_trees.end())
            throw std::runtime_error("mcut_b200: intersectOIBVHs called with BVHs that build_oibvh did not produce");
        s = is->second;
        c = ic->second;
    }
    mcb200_ctx* ctx = s.ctx; // both trees were built on this API thread's context
    // the previous dispatch's result and soup go back to the memory pool BEFORE the new ones are carved out of it
    if (t_last.res) mcb200_result_free(t_last.ctx, t_last.res);
    if (t_last.soup) mcb200_soup_free(t_last.ctx, t_last.soup);
    t_last = last_intersect_t();
    mcb200_result* res = nullptr;
    check(ctx, mcb200_result_create(ctx, &res), "result_create");
    check(ctx, mcb200_bvh_intersect(ctx, s.mesh, c.mesh, res), "bvh_intersect");
    mcb200_counts counts;
    check(ctx, mcb200_result_counts(ctx, res, &counts), "result_counts");
    // With the hooked kernel the map's CONTENT has no reader left (its consumers were the replaced region of dispatch());
    // preproc() only asks whether it is empty (preproc.cpp:2863-2865).  One placeholder entry answers that; the real set
    // stays on the device for the hook.
    static const bool kernel_is_hooked = dlsym(RTLD_DEFAULT, "mcb200_kernel_is_hooked") != nullptr;
    std::vector<uint64_t> pairs;
    if (kernel_is_hooked) {
        if (counts.n_pairs) pairs.push_back(0ull);
    } else {
        pairs.resize((size_t)counts.n_pairs);
        check(ctx, mcb200_result_read_pairs(ctx, res, pairs.data(), pairs.size()), "read_pairs");
    }
    // the pairs stay on the device for the narrowphase hook (mcb200_hook_narrowphase below)
    t_last.ctx = ctx;
    t_last.src = s.mesh;
    t_last.cut = c.mesh;
    t_last.res = res;
    t_last.from_arrays = s.from_arrays && c.from_arrays;
    t_last.src_from_arrays = s.from_arrays;
    t_last.cut_from_arrays = c.from_arrays;
    t_last.src_cap = s.cap;
    t_last.cut_cap = c.cap;

    (void)srcMeshBvhLeafNodeFaces;
    const uint32_t nsf = s.nf;
    // pairs are sorted by (src, cut): source keys arrive in ascending order -> amortised O(1) hinted inserts
    auto hint = ps_face_to_potentially_intersecting_others.end();
    uint32_t cur = 0xFFFFFFFFu;
    std::vector<fd_t>* cur_list = nullptr;
    for (uint64_t p : pairs) {
        const uint32_t sf = (uint32_t)(p >> 32), cf = (uint32_t)(p & 0xFFFFFFFFu) + nsf; // cut ids are offset (bvh.cpp:713)
        if (sf != cur) {
            hint = ps_face_to_potentially_intersecting_others.emplace_hint(ps_face_to_potentially_intersecting_others.end(), fd_t(sf),
                std::vector<fd_t>());
            cur_list = &hint->second;
            cur = sf;
        }
        cur_list->push_back(fd_t(cf));
    }
    for (uint64_t p : pairs) {
        const uint32_t sf = (uint32_t)(p >> 32), cf = (uint32_t)(p & 0xFFFFFFFFu) + nsf;
        ps_face_to_potentially_intersecting_others[fd_t(cf)].push_back(fd_t(sf));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Narrowphase hook (mcut_hook.h): replaces kernel.cpp:1779-3206 inside a live dispatch().
// ---------------------------------------------------------------------------------------------------------------------
static thread_local uint32_t t_pool_threads = 0;
void mcb200_hook_set_scheduler_threads(uint32_t helper_threads) { t_pool_threads = helper_threads; }
void mcb200_hook_set_face_boxes(const std::vector<bounding_box_t<vec3_<double>>>*, const std::vector<bounding_box_t<vec3_<double>>>*) {}

int mcb200_hook_narrowphase(const hmesh_t& ps, int sm_vtx_cnt, int sm_face_count,
    const std::map<fd_t, std::vector<fd_t>>& ps_face_to_potentially_intersecting_others, hmesh_t& m0,
    std::unordered_map<fd_t, vec3>& ps_tested_face_to_plane_normal,
    std::unordered_map<fd_t, scalar_t>& ps_tested_face_to_plane_normal_d_param,
    std::unordered_map<fd_t, int>& ps_tested_face_to_plane_normal_max_comp,
    std::unordered_map<fd_t, std::vector<vec3>>& ps_tested_face_to_vertices,
    std::vector<std::pair<ed_t, fd_t>>& m0_ivtx_to_intersection_registry_entry, std::vector<vd_t>& cm_border_reentrant_ivtx_list,
    std::unordered_map<ed_t, std::vector<vd_t>>& ps_intersecting_edges,
    std::map<pair<fd_t>, std::vector<vd_t>>& cutpath_edge_creation_info,
    std::unordered_map<fd_t, std::vector<vd_t>>& ps_iface_to_ivtx_list, bool& partial_cut_detected, int& bad_face)
{
    scope_timer timer(t_last.from_arrays ? "narrowphase hook (arrays)" : "narrowphase hook (generic or mixed)");
    // (the candidate pairs themselves are still on the device, in t_last.res; the map is only replayed for its order)
    if (!t_last.res) throw std::runtime_error("mcut_b200: narrowphase hook reached without a device broadphase on this thread");
    mcb200_ctx* ctx = t_last.ctx;
    const uint32_t nv = (uint32_t)ps.number_of_vertices(), nf = (uint32_t)ps.number_of_faces(), ne = (uint32_t)ps.number_of_edges();
    const uint32_t nsv = (uint32_t)sm_vtx_cnt, nsf = (uint32_t)sm_face_count, ncv = nv - nsv, ncf = nf - nsf;

    mcb200_soup* soup = nullptr;
    bool own_soup = true;
    // ---- fast path: both meshes live on the device as the user's own arrays ----
    // The cut mesh of THIS attempt is the most recent conversion on this thread (same arrays, new perturbation); three of
    // its vertices are compared with `ps` bit for bit before it is trusted.  A device mesh that was made from the user's own arrays keeps them (possibly float) and applies the frame itself; one
    // that was flattened from a half-edge mesh (a repartitioned mesh, or no capture) holds internal coordinates.
    const bool src_arrays = t_last.src_from_arrays && t_last.src_cap.nv == nsv && t_last.src_cap.nf == nsf;
    bool cut_arrays = t_last.cut_from_arrays && t_last.cut_cap.nv == ncv && t_last.cut_cap.nf == ncf;
    capture_t cut_now = t_last.cut_cap;
    if (cut_arrays) {
        if (t_latest_capture.xyz == t_last.cut_cap.xyz && t_latest_capture.nv == ncv && t_latest_capture.nf == ncf) cut_now = t_latest_capture;
        const uint32_t probe[3] = { 0u, ncv / 2u, ncv - 1u };
        for (uint32_t v : probe) {
            double want[3];
            captured_vertex(cut_now, v, want);
            const vec3& p = ps.vertex(vd_t(nsv + v));
            if (p.x() != want[0] || p.y() != want[1] || p.z() != want[2]) cut_arrays = false;
        }
    }
    if (t_last.cut_from_arrays && !cut_arrays) {
        // not the conversion this thread saw last: fall back to the coordinates of `ps` (possible when the device copy
        // holds doubles and the counts agree; the frame becomes the identity)
        if (t_last.cut_cap.is_float || t_last.cut_cap.nv != ncv || t_last.cut_cap.nf != ncf)
            throw std::runtime_error("mcut_b200: the cut mesh of this dispatch() is not the one its device tree was built from");
        check(ctx, mcb200_mesh_set_frame(ctx, t_last.cut, nullptr, nullptr, nullptr), "set_frame(cut, identity)");
    }
    if (t_last.src_from_arrays && !src_arrays)
        throw std::runtime_error("mcut_b200: the source mesh of this dispatch() is not the one its device tree was built from");
    const bool fast = src_arrays && cut_arrays;
    if (cut_arrays)
        check(ctx, mcb200_mesh_set_frame(ctx, t_last.cut, cut_now.com, cut_now.shift, cut_now.has_pert ? cut_now.pert : nullptr), "set_frame(cut)");
    if (fast) {
        scope_timer t2("  hook: soup numbering (device)");
        if (!t_last.soup) check(ctx, mcb200_soup_number(ctx, t_last.src, t_last.cut, t_last.res, &t_last.soup), "soup_number");
        soup = t_last.soup; // the numbering does not depend on coordinates: one per broadphase, reused by every retry
        own_soup = false;
    } else {
        // ---- generic path: coordinates as dispatch() sees them now (the cut mesh moves on every general-position retry) ----
        {
            std::vector<double> xyz(3 * (size_t)(nsv > ncv ? nsv : ncv));
            if (!src_arrays) {
                for (uint32_t v = 0; v < nsv; ++v) {
                    const vec3& p = ps.vertex(vd_t(v));
                    xyz[3 * (size_t)v] = p.x();
                    xyz[3 * (size_t)v + 1] = p.y();
                    xyz[3 * (size_t)v + 2] = p.z();
                }
                check(ctx, mcb200_mesh_update_xyz(ctx, t_last.src, xyz.data(), nsv), "mesh_update_xyz(src)");
            }
            if (!cut_arrays) {
                for (uint32_t v = 0; v < ncv; ++v) {
                    const vec3& p = ps.vertex(vd_t(nsv + v));
                    xyz[3 * (size_t)v] = p.x();
                    xyz[3 * (size_t)v + 1] = p.y();
                    xyz[3 * (size_t)v + 2] = p.z();
                }
                check(ctx, mcb200_mesh_update_xyz(ctx, t_last.cut, xyz.data(), ncv), "mesh_update_xyz(cut)");
            }
        }
        // ---- the ids of `ps` as flat arrays: vertex and edge of every halfedge slot, faces of h0 / h1 of every edge ----
        std::vector<uint32_t> face_vtx, face_edge, edge_f(2 * (size_t)ne);
        face_vtx.reserve(3 * (size_t)nf);
        face_edge.reserve(3 * (size_t)nf);
        std::vector<uint32_t> sizes(nf);
        for (uint32_t f = 0; f < nf; ++f) {
            const std::vector<hd_t>& hs = ps.get_halfedges_around_face(fd_t(f));
            sizes[f] = (uint32_t)hs.size();
            for (const hd_t& h : hs) {
                face_vtx.push_back((uint32_t)ps.target(h));
                face_edge.push_back((uint32_t)ps.edge(h));
            }
        }
        for (uint32_t e = 0; e < ne; ++e) {
            const fd_t f0 = ps.face(ps.halfedge(ed_t(e), 0)), f1 = ps.face(ps.halfedge(ed_t(e), 1));
            edge_f[2 * (size_t)e] = (f0 == hmesh_t::null_face()) ? MCB200_NULL : (uint32_t)f0;
            edge_f[2 * (size_t)e + 1] = (f1 == hmesh_t::null_face()) ? MCB200_NULL : (uint32_t)f1;
        }
        check(ctx, mcb200_soup_create_sized(ctx, nsf, ncf, (uint32_t)face_vtx.size(), ne, face_vtx.data(), face_edge.data(),
                       edge_f.data(), sizes.data(), &soup),
            "soup_create");
    }

    // ---- device narrowphase on the resident trees / pairs ----
    scope_timer* t_dev = new scope_timer("  hook: device narrowphase + counters");
    int rc = mcb200_narrowphase(ctx, soup, t_last.src, t_last.cut, t_last.res, 0);
    mcb200_counts counts;
    int rc2 = rc ? rc : mcb200_result_counts(ctx, t_last.res, &counts);
    for (int attempt = 0; rc2 == MCB200_ERR_CAPACITY && attempt < 3; ++attempt) {
        // a narrowphase buffer was too small for this input: the library has raised the capacities; the pairs are
        // produced again (their buffer moves when it grows) and the narrowphase repeated
        rc = mcb200_bvh_intersect(ctx, t_last.src, t_last.cut, t_last.res);
        if (!rc) rc = mcb200_narrowphase(ctx, soup, t_last.src, t_last.cut, t_last.res, 0);
        rc2 = rc ? rc : mcb200_result_counts(ctx, t_last.res, &counts);
    }
    delete t_dev;
    if (rc2) {
        if (own_soup) mcb200_soup_free(ctx, soup);
        check(ctx, rc2, "narrowphase");
    }
    if (getenv("MCB200_HOOK_DEBUG"))
        std::fprintf(stderr, "[mcut_b200 hook] %s path: ps nv=%u nf=%u ne=%u (src nv=%u nf=%u) pairs=%llu tests=%llu exact=%llu records=%llu cand_faces=%llu status=%d\n",
            fast ? "arrays" : "generic", nv, nf, ne, nsv, nsf, (unsigned long long)counts.n_pairs, (unsigned long long)counts.n_tests,
            (unsigned long long)counts.n_exact, (unsigned long long)counts.n_records, (unsigned long long)counts.n_cand_faces, (int)counts.status);
    if (counts.status == MCB200_STATUS_INVALID_SRC_MESH || counts.status == MCB200_STATUS_INVALID_CUT_MESH) {
        bad_face = (int)counts.bad_face;
        if (own_soup) mcb200_soup_free(ctx, soup);
        return counts.status == MCB200_STATUS_INVALID_CUT_MESH ? MCB200_HOOK_INVALID_CUT_MESH : MCB200_HOOK_INVALID_SRC_MESH;
    }
    if (counts.status == MCB200_STATUS_GENERAL_POSITION_VIOLATION) {
        if (own_soup) mcb200_soup_free(ctx, soup);
        return MCB200_HOOK_GENERAL_POSITION_VIOLATION;
    }

    // ---- plane rows of the candidate faces (kernel.cpp:2184-2356, faces ascending) and the registry records
    // (kernel.cpp:2601-2655, merged form :2673-2868; canonical (edge, face) order), then the host half (hook_fill.h) ----
    const size_t n_cand = (size_t)counts.n_cand_faces;
    std::vector<uint32_t> cand_faces(n_cand);
    std::vector<double> normal(3 * n_cand), d(n_cand);
    std::vector<int32_t> mc(n_cand);
    std::vector<mcb200_record> rec((size_t)counts.n_records);
    {
        mcb200_scope_timer t2("  hook: read planes + records");
        check(ctx, mcb200_result_read_planes(ctx, t_last.res, cand_faces.data(), normal.data(), d.data(), mc.data(), n_cand), "read_planes");
        check(ctx, mcb200_result_read_records(ctx, t_last.res, rec.data(), rec.size()), "read_records");
    }
    if (own_soup) mcb200_soup_free(ctx, soup);
    mcb200_hook_finish(ps, sm_vtx_cnt, sm_face_count, "device", n_cand, cand_faces.data(), normal.data(), d.data(), mc.data(), rec,
        t_pool_threads, m0, ps_tested_face_to_plane_normal, ps_tested_face_to_plane_normal_d_param,
        ps_tested_face_to_plane_normal_max_comp, ps_tested_face_to_vertices, m0_ivtx_to_intersection_registry_entry,
        cm_border_reentrant_ivtx_list, ps_intersecting_edges, cutpath_edge_creation_info, ps_iface_to_ivtx_list, partial_cut_detected);
    return MCB200_HOOK_OK;
}
