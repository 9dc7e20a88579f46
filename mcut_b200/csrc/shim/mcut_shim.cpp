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
#include <cstdlib>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "mcut/internal/bvh.h"
#include "mcut/internal/hmesh.h"
#include "mcut/internal/math.h"

#include "../../../include/mcut_b200.h"

namespace {

typedef bounding_box_t<vec3_<double>> bbox_t;

struct device_tree_t {
    mcb200_ctx* ctx = nullptr;
    mcb200_mesh* mesh = nullptr;
    uint32_t nf = 0;
};

std::mutex g_mutex;
std::unordered_map<const void*, device_tree_t> g_trees; // key: address of the caller's bvhAABBs vector

// One device context per API thread of the reference (every MCUT context owns its API thread, frontend.h:548-584), so
// concurrently dispatching contexts never share a stream.  Device: MCB200_DEVICE, else threads are dealt round-robin.
mcb200_ctx* thread_ctx()
{
    static thread_local mcb200_ctx* ctx = nullptr;
    if (!ctx) {
        static int next_device = 0;
        int device = 0;
        const int n = mcb200_device_count();
        if (const char* e = std::getenv("MCB200_DEVICE")) device = std::atoi(e);
        else if (n > 0) {
            std::lock_guard<std::mutex> lk(g_mutex);
            device = next_device++ % n;
        }
        const int rc = mcb200_ctx_create(device, nullptr, &ctx);
        if (rc != 0) throw std::runtime_error(std::string("mcut_b200: ") + mcb200_last_error(nullptr));
    }
    return ctx;
}

void check(mcb200_ctx* ctx, int rc, const char* what)
{
    if (rc != 0) throw std::runtime_error(std::string("mcut_b200: ") + what + ": " + mcb200_last_error(ctx));
}

} // namespace

void build_oibvh(thread_pool& /*pool*/, const hmesh_t& mesh, std::vector<bbox_t>& bvhAABBs, std::vector<fd_t>& bvhLeafNodeFaces,
    std::vector<bbox_t>& face_bboxes, const double& slightEnlargmentEps, const double /*multiplier*/)
{
    mcb200_ctx* ctx = thread_ctx();
    // flatten the half-edge mesh (internal coordinates: the reference has already re-centred them)
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
    const uint32_t nf = (uint32_t)mesh.number_of_faces();
    std::vector<uint32_t> sizes, idx;
    sizes.reserve(nf);
    idx.reserve(3 * (size_t)nf);
    std::vector<vd_t> tmp;
    for (face_array_iterator_t f = mesh.faces_begin(); f != mesh.faces_end(); ++f) {
        mesh.get_vertices_around_face(tmp, *f);
        sizes.push_back((uint32_t)tmp.size());
        for (const vd_t& v : tmp) idx.push_back((uint32_t)v);
    }

    device_tree_t t;
    t.ctx = ctx;
    t.nf = nf;
    check(ctx, mcb200_mesh_create(ctx, 0, xyz.data(), nv, idx.data(), sizes.data(), nf, &t.mesh), "mesh_create");
    check(ctx, mcb200_mesh_set_frame(ctx, t.mesh, nullptr, nullptr, nullptr), "set_frame");
    check(ctx, mcb200_bvh_build(ctx, t.mesh, slightEnlargmentEps), "bvh_build");

    std::vector<double> boxes(6 * (size_t)nf);
    double root[6];
    check(ctx, mcb200_bvh_read(ctx, t.mesh, boxes.data(), root), "bvh_read");
    face_bboxes.resize(nf);
    for (uint32_t f = 0; f < nf; ++f) {
        const double* b = boxes.data() + 6 * (size_t)f;
        face_bboxes[f] = bbox_t(vec3_<double>(b[0], b[1], b[2]), vec3_<double>(b[3], b[4], b[5]));
    }
    // callers read bvhAABBs[0] (the mesh AABB) only; the node count keeps the reference's size so nothing else changes
    const int np2 = [&]() { int x = (int)nf - 1; x |= x >> 1; x |= x >> 2; x |= x >> 4; x |= x >> 8; x |= x >> 16; return x + 1; }();
    bvhAABBs.assign((size_t)(2 * (int)nf - 1 + __builtin_popcount((unsigned)(np2 - (int)nf))), bbox_t());
    bvhAABBs[0] = bbox_t(vec3_<double>(root[0], root[1], root[2]), vec3_<double>(root[3], root[4], root[5]));
    bvhLeafNodeFaces.assign(nf, fd_t(0)); // opaque to everyone but intersectOIBVHs, which uses the device tree instead

    std::lock_guard<std::mutex> lk(g_mutex);
    auto it = g_trees.find(&bvhAABBs);
    if (it != g_trees.end()) mcb200_mesh_free(it->second.ctx, it->second.mesh); // same caller vector rebuilt
    g_trees[&bvhAABBs] = t;
}

void intersectOIBVHs(std::map<fd_t, std::vector<fd_t>>& ps_face_to_potentially_intersecting_others,
    const std::vector<bbox_t>& srcMeshBvhAABBs, const std::vector<fd_t>& srcMeshBvhLeafNodeFaces,
    const std::vector<bbox_t>& cutMeshBvhAABBs, const std::vector<fd_t>& /*cutMeshBvhLeafNodeFaces*/)
{
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
    mcb200_result* res = nullptr;
    check(ctx, mcb200_result_create(ctx, &res), "result_create");
    check(ctx, mcb200_bvh_intersect(ctx, s.mesh, c.mesh, res), "bvh_intersect");
    mcb200_counts counts;
    check(ctx, mcb200_result_counts(ctx, res, &counts), "result_counts");
    std::vector<uint64_t> pairs((size_t)counts.n_pairs);
    check(ctx, mcb200_result_read_pairs(ctx, res, pairs.data(), pairs.size()), "read_pairs");
    mcb200_result_free(ctx, res);

    const uint32_t nsf = (uint32_t)srcMeshBvhLeafNodeFaces.size();
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
