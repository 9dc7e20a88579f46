// mcut_b200/csrc/shim/hook_fill.h — host half of the narrowphase hook: from flat results (plane rows, registry records) to the
// containers the rest of the reference's dispatch() reads.  Shared by the real hook (mcut_shim.cpp: results from the device)
// and by the CPU stand-in that the `-m "not gpu"` tests drive (oracle/hook_oracle.cpp: results from the oracle), so the
// container logic and the registry order are checked on every round, GPU or not.
#pragma once

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../../include/mcut_b200.h"
#include "mcut/internal/hmesh.h"
#include "mcut/internal/math.h"
#include "mcut/internal/utils.h"

// MCB200_SHIM_TIMING=1: wall time of a scope on stderr (the adapter's entries and the steps of the hook's host half)
struct mcb200_scope_timer {
    const char* what;
    std::chrono::steady_clock::time_point t0;
    explicit mcb200_scope_timer(const char* w) : what(w), t0(std::chrono::steady_clock::now()) {}
    ~mcb200_scope_timer()
    {
        static const bool on = std::getenv("MCB200_SHIM_TIMING") != nullptr;
        if (on)
            std::fprintf(stderr, "[mcut_b200 shim] %s: %.3f ms\n", what,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
};

// MCB200_HOOK_DEBUG=2: what the hook hands over, bit patterns included (to diff the device hook against the CPU stand-in)
static inline void mcb200_hook_dump(const char* who, size_t np, const uint32_t* faces, const double* normal, const double* d,
    const int32_t* mc, const mcb200_record* rec, size_t nr)
{
    const char* e = getenv("MCB200_HOOK_DEBUG");
    if (!e || e[0] != '2') return;
    for (size_t k = 0; k < np; ++k) {
        unsigned long long b[4];
        memcpy(b, normal + 3 * k, 24);
        memcpy(b + 3, d + k, 8);
        fprintf(stderr, "[%s] plane f=%u n=%016llx %016llx %016llx d=%016llx mc=%d\n", who, faces[k], b[0], b[1], b[2], b[3], (int)mc[k]);
    }
    for (size_t i = 0; i < nr; ++i) {
        unsigned long long b[3];
        memcpy(b, rec[i].point, 24);
        fprintf(stderr, "[%s] rec e=%u f=%u p=%016llx %016llx %016llx\n", who, rec[i].edge, rec[i].face, b[0], b[1], b[2]);
    }
}

// kernel.cpp:2184-2356: plane normal, d, largest component and vertex list of the candidate faces.  The reference fills these
// maps for every candidate face, but the rest of dispatch() only ever looks up faces named by a registry entry — the tested
// face and the faces incident to the tested edge (ps_get_ivtx_registry_entry_faces; kernel.cpp:3536-3567, :4187-4191,
// :6637-6664, :6811-6814) — and never iterates them; `needed` restricts the rows to those faces.
static inline void mcb200_hook_fill_planes(const hmesh_t& ps, size_t n, const uint32_t* faces, const double* normal, const double* d,
    const int32_t* max_comp, const unsigned char* needed /* by face id, or NULL: every row */,
    std::unordered_map<fd_t, vec3>& ps_tested_face_to_plane_normal,
    std::unordered_map<fd_t, scalar_t>& ps_tested_face_to_plane_normal_d_param,
    std::unordered_map<fd_t, int>& ps_tested_face_to_plane_normal_max_comp,
    std::unordered_map<fd_t, std::vector<vec3>>& ps_tested_face_to_vertices)
{
    mcb200_scope_timer timer("  hook: plane maps");
    std::vector<vd_t> tmp;
    for (size_t k = 0; k < n; ++k) {
        if (needed && !needed[faces[k]]) continue;
        const fd_t f(faces[k]);
        ps_tested_face_to_plane_normal[f] = vec3(normal[3 * k], normal[3 * k + 1], normal[3 * k + 2]);
        ps_tested_face_to_plane_normal_d_param[f] = d[k];
        ps_tested_face_to_plane_normal_max_comp[f] = (int)max_comp[k];
        std::vector<vec3>& verts = ps_tested_face_to_vertices[f];
        ps.get_vertices_around_face(tmp, f);
        verts.reserve(tmp.size());
        for (const vd_t& v : tmp) verts.push_back(ps.vertex(v));
    }
}

// The reference's registry order (mcb200_reference_edge_order, host_logic.cpp): the edges around the candidate faces as flat
// arrays, the library's replay of the reference's unordered_map, and the records sorted into that order.
static inline void mcb200_hook_reference_order(const hmesh_t& ps, const uint32_t* cand_faces, size_t n_cand, uint32_t pool_threads,
    std::vector<mcb200_record>& rec)
{
    // cand_faces: every face with a candidate partner, ascending = the keys of the reference's
    // ps_face_to_potentially_intersecting_others (the hooked adapter does not fill that map on the host: the pairs stay
    // on the device; the plane rows name the same faces)
    mcb200_scope_timer timer("  hook: registry order replay");
    const uint32_t ne = (uint32_t)ps.number_of_edges();
    std::vector<uint32_t> off(n_cand + 1, 0u), slots;
    slots.reserve(3 * n_cand);
    for (size_t i = 0; i < n_cand; ++i) {
        for (const hd_t& he : ps.get_halfedges_around_face(fd_t(cand_faces[i]))) slots.push_back((uint32_t)ps.edge(he));
        off[i + 1] = (uint32_t)slots.size();
    }
    std::vector<uint32_t> order(slots.size() ? slots.size() : 1u);
    uint32_t n = 0;
    if (mcb200_reference_edge_order((uint32_t)n_cand, off.data(), slots.data(), pool_threads, order.data(), &n))
        throw std::runtime_error("mcut_b200: mcb200_reference_edge_order failed");
    if (const char* e = getenv("MCB200_HOOK_DEBUG")) {
        if (e[0] == '2') {
            fprintf(stderr, "[order] threads=%u:", pool_threads);
            for (uint32_t k = 0; k < n; ++k) fprintf(stderr, " %u", order[k]);
            fprintf(stderr, "\n");
        }
    }
    // rank of the edges that matter (uninitialised elsewhere: every record's edge belongs to a candidate face)
    std::unique_ptr<uint32_t[]> rank(new uint32_t[ne ? ne : 1u]);
    for (const mcb200_record& r : rec) rank[r.edge] = MCB200_NULL;
    for (uint32_t k = 0; k < n; ++k)
        if (order[k] < ne) rank[order[k]] = k;
    for (const mcb200_record& r : rec)
        if (rank[r.edge] == MCB200_NULL) throw std::runtime_error("mcut_b200: a registry record names an edge of no candidate face");
    // the records arrive sorted by (edge, face): a stable sort on the rank leaves each edge's faces ascending
    std::stable_sort(rec.begin(), rec.end(), [&](const mcb200_record& a, const mcb200_record& b) { return rank[a.edge] < rank[b.edge]; });
}

// kernel.cpp:2601-2655 (merged form :2673-2868): one m0 vertex per record, in the order given, and everything keyed by it
static inline void mcb200_hook_fill_registry(const hmesh_t& ps, int sm_vtx_cnt, int sm_face_count, const mcb200_record* rec, size_t n,
    hmesh_t& m0, std::vector<std::pair<ed_t, fd_t>>& m0_ivtx_to_intersection_registry_entry, std::vector<vd_t>& cm_border_reentrant_ivtx_list,
    std::unordered_map<ed_t, std::vector<vd_t>>& ps_intersecting_edges, std::map<pair<fd_t>, std::vector<vd_t>>& cutpath_edge_creation_info,
    std::unordered_map<fd_t, std::vector<vd_t>>& ps_iface_to_ivtx_list, bool& partial_cut_detected)
{
    mcb200_scope_timer timer("  hook: registry containers");
    m0_ivtx_to_intersection_registry_entry.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        const mcb200_record& r = rec[i];
        const ed_t tested_edge(r.edge);
        const fd_t tested_face(r.face);
        const vd_t v = m0.add_vertex(vec3(r.point[0], r.point[1], r.point[2])); // ps_vtx_cnt + index in the registry
        m0_ivtx_to_intersection_registry_entry.push_back(std::make_pair(tested_edge, tested_face));
        ps_intersecting_edges[tested_edge].push_back(v);
        const hd_t h0 = ps.halfedge(tested_edge, 0), h1 = ps.halfedge(tested_edge, 1);
        const fd_t h0_face = ps.face(h0), h1_face = ps.face(h1);
        const fd_t tested_edge_face = h0_face != hmesh_t::null_face() ? h0_face : h1_face;
        const bool tested_edge_belongs_to_cm = ((int)tested_edge_face) >= sm_face_count;
        const fd_t face_pqr = tested_edge_face;
        const fd_t face_pqs = tested_edge_face == h0_face ? h1_face : hmesh_t::null_face();
        if (tested_edge_belongs_to_cm) { // key format: {source-mesh face, cut-mesh face}
            cutpath_edge_creation_info[make_pair(tested_face, face_pqr)].push_back(v);
            if (face_pqs != hmesh_t::null_face()) cutpath_edge_creation_info[make_pair(tested_face, face_pqs)].push_back(v);
        } else {
            cutpath_edge_creation_info[make_pair(tested_edge_face, tested_face)].push_back(v);
            const fd_t other = (tested_edge_face == h0_face) ? h1_face : h0_face;
            if (other != hmesh_t::null_face()) cutpath_edge_creation_info[make_pair(other, tested_face)].push_back(v);
        }
        if (tested_edge_belongs_to_cm && (h0_face == hmesh_t::null_face() || h1_face == hmesh_t::null_face())) // ps.is_border(tested_edge)
            cm_border_reentrant_ivtx_list.push_back(v);
        ps_iface_to_ivtx_list[tested_face].push_back(v);
        if (h0_face != hmesh_t::null_face()) ps_iface_to_ivtx_list[h0_face].push_back(v);
        if (h1_face != hmesh_t::null_face()) ps_iface_to_ivtx_list[h1_face].push_back(v);
        if (!partial_cut_detected) {
            const bool is_cs_edge = ((int)ps.source(h0)) >= sm_vtx_cnt;
            const bool is_border = (h0_face == hmesh_t::null_face() || h1_face == hmesh_t::null_face());
            partial_cut_detected = (is_cs_edge && is_border);
        }
    }
}

// Everything after the narrowphase itself: plane rows -> maps (registry faces only), records -> the reference's own order ->
// containers.  `rec` arrives sorted by (edge, face) and is reordered in place.
static inline void mcb200_hook_finish(const hmesh_t& ps, int sm_vtx_cnt, int sm_face_count, const char* who, size_t n_cand,
    const uint32_t* cand_faces, const double* normal, const double* d, const int32_t* max_comp, std::vector<mcb200_record>& rec,
    uint32_t pool_threads, hmesh_t& m0, std::unordered_map<fd_t, vec3>& ps_tested_face_to_plane_normal,
    std::unordered_map<fd_t, scalar_t>& ps_tested_face_to_plane_normal_d_param,
    std::unordered_map<fd_t, int>& ps_tested_face_to_plane_normal_max_comp,
    std::unordered_map<fd_t, std::vector<vec3>>& ps_tested_face_to_vertices,
    std::vector<std::pair<ed_t, fd_t>>& m0_ivtx_to_intersection_registry_entry, std::vector<vd_t>& cm_border_reentrant_ivtx_list,
    std::unordered_map<ed_t, std::vector<vd_t>>& ps_intersecting_edges, std::map<pair<fd_t>, std::vector<vd_t>>& cutpath_edge_creation_info,
    std::unordered_map<fd_t, std::vector<vd_t>>& ps_iface_to_ivtx_list, bool& partial_cut_detected)
{
    mcb200_hook_dump(who, n_cand, cand_faces, normal, d, max_comp, nullptr, 0);
    {
        std::vector<unsigned char> needed;
        if (!getenv("MCB200_HOOK_ALL_PLANES")) {
            needed.assign((size_t)ps.number_of_faces(), 0);
            for (const mcb200_record& r : rec) {
                needed[r.face] = 1;
                const fd_t f0 = ps.face(ps.halfedge(ed_t(r.edge), 0)), f1 = ps.face(ps.halfedge(ed_t(r.edge), 1));
                if (f0 != hmesh_t::null_face()) needed[(uint32_t)f0] = 1;
                if (f1 != hmesh_t::null_face()) needed[(uint32_t)f1] = 1;
            }
        }
        mcb200_hook_fill_planes(ps, n_cand, cand_faces, normal, d, max_comp, needed.empty() ? nullptr : needed.data(),
            ps_tested_face_to_plane_normal, ps_tested_face_to_plane_normal_d_param, ps_tested_face_to_plane_normal_max_comp,
            ps_tested_face_to_vertices);
    }
    if (!rec.empty() && !getenv("MCB200_CANONICAL_REGISTRY")) {
        // registry in the reference's own order (mcb200_reference_edge_order), an edge's faces ascending
        mcb200_hook_reference_order(ps, cand_faces, n_cand, pool_threads, rec);
    }
    mcb200_hook_dump(who, 0, nullptr, nullptr, nullptr, nullptr, rec.data(), rec.size());
    mcb200_hook_fill_registry(ps, sm_vtx_cnt, sm_face_count, rec.data(), rec.size(), m0, m0_ivtx_to_intersection_registry_entry,
        cm_border_reentrant_ivtx_list, ps_intersecting_edges, cutpath_edge_creation_info, ps_iface_to_ivtx_list, partial_cut_detected);
    if (const char* path = getenv("MCB200_HOOK_DUMP_CUTPATH")) {
        // what "Create edges with intersection points" (kernel.cpp:3332-3617) is about to consume, for tests/test_oracle_cutpath.py:
        // the registry in its final order (+ the faces of every tested edge) and cutpath_edge_creation_info as the live dispatch holds it
        if (FILE* fp = fopen(path, "w")) {
            const uint32_t base = (uint32_t)m0.number_of_vertices() - (uint32_t)rec.size();
            fprintf(fp, "R %zu %d\n", rec.size(), sm_face_count);
            for (const mcb200_record& r : rec) {
                const fd_t f0 = ps.face(ps.halfedge(ed_t(r.edge), 0)), f1 = ps.face(ps.halfedge(ed_t(r.edge), 1));
                unsigned long long b[3];
                memcpy(b, r.point, 24);
                fprintf(fp, "%u %u %u %u %016llx %016llx %016llx\n", r.edge, r.face, f0 == hmesh_t::null_face() ? 0xFFFFFFFFu : (uint32_t)f0,
                    f1 == hmesh_t::null_face() ? 0xFFFFFFFFu : (uint32_t)f1, b[0], b[1], b[2]);
            }
            fprintf(fp, "G %zu\n", cutpath_edge_creation_info.size());
            for (const auto& kv : cutpath_edge_creation_info) {
                fprintf(fp, "%u %u %zu", (uint32_t)kv.first.first, (uint32_t)kv.first.second, kv.second.size());
                for (const vd_t& v : kv.second) fprintf(fp, " %u", (uint32_t)v - base);
                fprintf(fp, "\n");
            }
            fclose(fp);
        }
    }
}
