// mcut_b200/csrc/shim/hook_fill.h — host half of the narrowphase hook: from flat results (plane rows, registry records) to the
// containers the rest of the reference's dispatch() reads.  Shared by the real hook (mcut_shim.cpp: results from the device)
// and by the CPU stand-in that the `-m "not gpu"` tests drive (oracle/hook_oracle.cpp: results from the oracle), so the
// container logic and the registry order are checked on every round, GPU or not.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../../include/mcut_b200.h"
#include "mcut/internal/hmesh.h"
#include "mcut/internal/math.h"
#include "mcut/internal/utils.h"

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

// kernel.cpp:2184-2356: plane normal, d, largest component and vertex list of every candidate face
static inline void mcb200_hook_fill_planes(const hmesh_t& ps, size_t n, const uint32_t* faces, const double* normal, const double* d,
    const int32_t* max_comp, std::unordered_map<fd_t, vec3>& ps_tested_face_to_plane_normal,
    std::unordered_map<fd_t, scalar_t>& ps_tested_face_to_plane_normal_d_param,
    std::unordered_map<fd_t, int>& ps_tested_face_to_plane_normal_max_comp,
    std::unordered_map<fd_t, std::vector<vec3>>& ps_tested_face_to_vertices)
{
    std::vector<vd_t> tmp;
    for (size_t k = 0; k < n; ++k) {
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

// The reference's registry order (mcb200_reference_edge_rank, host_logic.cpp): the candidate faces and the edge of every
// halfedge slot of `ps` as flat arrays, then the library's replay of the reference's unordered_map.
static inline std::vector<uint32_t> mcb200_hook_reference_edge_rank(const hmesh_t& ps, const uint32_t* cand_faces, size_t n_cand,
    uint32_t pool_threads)
{
    // cand_faces: every face with a candidate partner, ascending = the keys of the reference's
    // ps_face_to_potentially_intersecting_others (the hooked adapter does not fill that map on the host: the pairs stay
    // on the device; the plane rows name the same faces)
    const uint32_t nf = (uint32_t)ps.number_of_faces(), ne = (uint32_t)ps.number_of_edges();
    std::vector<uint32_t> faces(cand_faces, cand_faces + n_cand), off((size_t)nf + 1, 0u), fe;
    std::vector<uint32_t> size_of(nf, 0u); // only the candidate faces' slots are read: the other faces get empty ranges
    for (uint32_t f : faces) size_of[f] = (uint32_t)ps.get_halfedges_around_face(fd_t(f)).size();
    for (uint32_t f = 0; f < nf; ++f) off[f + 1] = off[f] + size_of[f];
    fe.resize(off[nf] ? off[nf] : 1u);
    for (uint32_t f : faces) {
        uint32_t h = off[f];
        for (const hd_t& he : ps.get_halfedges_around_face(fd_t(f))) fe[h++] = (uint32_t)ps.edge(he);
    }
    std::vector<uint32_t> rank(ne ? ne : 1u);
    if (mcb200_reference_edge_rank((uint32_t)faces.size(), faces.data(), off.data(), fe.data(), ne, pool_threads, rank.data()))
        throw std::runtime_error("mcut_b200: mcb200_reference_edge_rank failed");
    if (const char* e = getenv("MCB200_HOOK_DEBUG")) {
        if (e[0] == '2') {
            fprintf(stderr, "[rank] threads=%u cand:", pool_threads);
            for (uint32_t f : faces) {
                fprintf(stderr, " %u(", f);
                for (uint32_t h = off[f]; h < off[f + 1]; ++h) fprintf(stderr, "%u ", fe[h]);
                fprintf(stderr, ")");
            }
            fprintf(stderr, " rank:");
            for (uint32_t k = 0; k < ne; ++k) fprintf(stderr, " %d", (int)rank[k]);
            fprintf(stderr, "\n");
        }
    }
    return rank;
}

// kernel.cpp:2601-2655 (merged form :2673-2868): one m0 vertex per record, in the order given, and everything keyed by it
static inline void mcb200_hook_fill_registry(const hmesh_t& ps, int sm_vtx_cnt, int sm_face_count, const mcb200_record* rec, size_t n,
    hmesh_t& m0, std::vector<std::pair<ed_t, fd_t>>& m0_ivtx_to_intersection_registry_entry, std::vector<vd_t>& cm_border_reentrant_ivtx_list,
    std::unordered_map<ed_t, std::vector<vd_t>>& ps_intersecting_edges, std::map<pair<fd_t>, std::vector<vd_t>>& cutpath_edge_creation_info,
    std::unordered_map<fd_t, std::vector<vd_t>>& ps_iface_to_ivtx_list, bool& partial_cut_detected)
{
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
