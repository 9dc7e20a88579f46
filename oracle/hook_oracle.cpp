// oracle/hook_oracle.cpp — TEST INFRASTRUCTURE: a CPU stand-in for the device narrowphase hook.
//
// oracle/_ref/libmcut_hooked.so is the reference with the inline narrowphase of dispatch() replaced by one call to
// mcb200_hook_narrowphase() (mcut_b200/csrc/shim/mcut_hook.h).  In the product that call lands in the adapter
// (mcut_shim.cpp) and runs on the B200.  Here the same entry point is answered by the ORACLE (oracle/mcut_oracle.c), so that
// the host half of the hook — how flat records become the reference's containers, and in which order the registry is handed
// over (mcut_b200/csrc/shim/hook_fill.h, mcb200_reference_edge_rank) — can be checked against the unmodified reference
// without a GPU: tests/test_hook_cpu.py drives oracle/_ref/api_driver_hooked_cpu over the reference's regression corpus.
// Nothing in the product links this file.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "mcut_hook.h"
#include "hook_fill.h"
#include "mcut_oracle.h"

extern "C" { int mcb200_kernel_is_hooked_cpu = 1; }

static thread_local uint32_t t_pool_threads = 0;
static thread_local const std::vector<bounding_box_t<vec3_<double>>>* t_src_boxes = nullptr;
static thread_local const std::vector<bounding_box_t<vec3_<double>>>* t_cut_boxes = nullptr;

void mcb200_hook_set_scheduler_threads(uint32_t helper_threads) { t_pool_threads = helper_threads; }
void mcb200_hook_set_face_boxes(const std::vector<bounding_box_t<vec3_<double>>>* s, const std::vector<bounding_box_t<vec3_<double>>>* c)
{
    t_src_boxes = s;
    t_cut_boxes = c;
}

int mcb200_hook_narrowphase(const hmesh_t& ps, int sm_vtx_cnt, int sm_face_count,
    const std::map<fd_t, std::vector<fd_t>>& cand, hmesh_t& m0, std::unordered_map<fd_t, vec3>& ps_tested_face_to_plane_normal,
    std::unordered_map<fd_t, scalar_t>& ps_tested_face_to_plane_normal_d_param,
    std::unordered_map<fd_t, int>& ps_tested_face_to_plane_normal_max_comp,
    std::unordered_map<fd_t, std::vector<vec3>>& ps_tested_face_to_vertices,
    std::vector<std::pair<ed_t, fd_t>>& m0_ivtx_to_intersection_registry_entry, std::vector<vd_t>& cm_border_reentrant_ivtx_list,
    std::unordered_map<ed_t, std::vector<vd_t>>& ps_intersecting_edges, std::map<pair<fd_t>, std::vector<vd_t>>& cutpath_edge_creation_info,
    std::unordered_map<fd_t, std::vector<vd_t>>& ps_iface_to_ivtx_list, bool& partial_cut_detected, int& bad_face)
{
    if (!t_src_boxes || !t_cut_boxes) throw std::runtime_error("hook_oracle: face boxes were not handed over");
    const uint32_t nv = (uint32_t)ps.number_of_vertices(), nf = (uint32_t)ps.number_of_faces(), ne = (uint32_t)ps.number_of_edges();
    const uint32_t nsf = (uint32_t)sm_face_count, ncf = nf - nsf;
    // ---- `ps` as the oracle's flat polygon soup ----
    std::vector<double> xyz(3 * (size_t)nv);
    for (uint32_t v = 0; v < nv; ++v) {
        const vec3& p = ps.vertex(vd_t(v));
        xyz[3 * (size_t)v] = p.x();
        xyz[3 * (size_t)v + 1] = p.y();
        xyz[3 * (size_t)v + 2] = p.z();
    }
    std::vector<uint32_t> face_off((size_t)nf + 1, 0u), face_vtx, face_edge, edge_v(2 * (size_t)ne), edge_f(2 * (size_t)ne);
    for (uint32_t f = 0; f < nf; ++f) {
        const std::vector<hd_t>& hs = ps.get_halfedges_around_face(fd_t(f));
        face_off[f + 1] = face_off[f] + (uint32_t)hs.size();
        for (const hd_t& h : hs) {
            face_vtx.push_back((uint32_t)ps.target(h));
            face_edge.push_back((uint32_t)ps.edge(h));
        }
    }
    for (uint32_t e = 0; e < ne; ++e) {
        const hd_t h0 = ps.halfedge(ed_t(e), 0), h1 = ps.halfedge(ed_t(e), 1);
        const fd_t f0 = ps.face(h0), f1 = ps.face(h1);
        edge_v[2 * (size_t)e] = (uint32_t)ps.source(h0);
        edge_v[2 * (size_t)e + 1] = (uint32_t)ps.target(h0);
        edge_f[2 * (size_t)e] = (f0 == hmesh_t::null_face()) ? MCO_NULL : (uint32_t)f0;
        edge_f[2 * (size_t)e + 1] = (f1 == hmesh_t::null_face()) ? MCO_NULL : (uint32_t)f1;
    }
    mco_soup_t soup;
    soup.nv = nv;
    soup.nf = nf;
    soup.ne = ne;
    soup.nh = (uint32_t)face_vtx.size();
    soup.src_nv = (uint32_t)sm_vtx_cnt;
    soup.src_nf = nsf;
    soup.xyz = xyz.data();
    soup.face_off = face_off.data();
    soup.face_vtx = face_vtx.data();
    soup.face_edge = face_edge.data();
    soup.edge_v = edge_v.data();
    soup.edge_f = edge_f.data();
    // ---- candidate pairs (source keys of the map) and the BVH-build face boxes ----
    std::vector<uint64_t> pairs;
    for (const auto& kv : cand) {
        if ((uint32_t)kv.first >= nsf) continue;
        for (const fd_t& c : kv.second) pairs.push_back(((uint64_t)(uint32_t)kv.first << 32) | ((uint32_t)c - nsf));
    }
    std::sort(pairs.begin(), pairs.end());
    auto flat = [](const std::vector<bounding_box_t<vec3_<double>>>& b, uint32_t n) {
        std::vector<double> out(6 * (size_t)n);
        for (uint32_t f = 0; f < n && f < b.size(); ++f) {
            for (int k = 0; k < 3; ++k) {
                out[6 * (size_t)f + k] = b[f].minimum()[k];
                out[6 * (size_t)f + 3 + k] = b[f].maximum()[k];
            }
        }
        return out;
    };
    const std::vector<double> sb = flat(*t_src_boxes, nsf), cb = flat(*t_cut_boxes, ncf);
    mco_narrow_out_t out;
    mco_narrowphase(&soup, pairs.data(), pairs.size(), sb.data(), cb.data(), 1, &out); // returns out.status
    if (getenv("MCB200_HOOK_DEBUG"))
        std::fprintf(stderr, "[hook_oracle] ps nv=%u nf=%u ne=%u pairs=%zu tests=%zu records=%zu cand_faces=%zu status=%d\n", nv, nf, ne,
            pairs.size(), out.n_tests, out.n_records, out.n_cand_faces, out.status);
    int rc = MCB200_HOOK_OK;
    if (out.status == MCO_INVALID_SRC_MESH || out.status == MCO_INVALID_CUT_MESH) {
        bad_face = (int)out.bad_face;
        rc = out.status == MCO_INVALID_CUT_MESH ? MCB200_HOOK_INVALID_CUT_MESH : MCB200_HOOK_INVALID_SRC_MESH;
    } else if (out.status == MCO_GENERAL_POSITION_VIOLATION) {
        rc = MCB200_HOOK_GENERAL_POSITION_VIOLATION;
    } else {
        std::vector<mcb200_record> rec(out.n_records);
        for (size_t i = 0; i < out.n_records; ++i) {
            rec[i].edge = out.records[i].edge;
            rec[i].face = out.records[i].face;
            for (int k = 0; k < 3; ++k) rec[i].point[k] = out.records[i].point[k];
        }
        mcb200_hook_finish(ps, sm_vtx_cnt, sm_face_count, "oracle", out.n_cand_faces, out.cand_faces, out.cand_normal, out.cand_d,
            out.cand_maxcomp, rec, t_pool_threads, m0, ps_tested_face_to_plane_normal, ps_tested_face_to_plane_normal_d_param,
            ps_tested_face_to_plane_normal_max_comp, ps_tested_face_to_vertices, m0_ivtx_to_intersection_registry_entry,
            cm_border_reentrant_ivtx_list, ps_intersecting_edges, cutpath_edge_creation_info, ps_iface_to_ivtx_list, partial_cut_detected);
    }
    mco_narrow_free(&out);
    return rc;
}
