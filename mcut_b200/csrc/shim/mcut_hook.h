// mcut_b200/csrc/shim/mcut_hook.h — the narrowphase hook the patched reference kernel calls.
//
// `oracle/make_hooked_kernel.py` builds a copy of the reference's source/kernel.cpp in which the inline narrowphase of
// dispatch() (kernel.cpp:1779-3206: "Prepare edge-to-face pairs" ... "Calculate intersection points") is replaced by ONE
// call to this function; everything before and after is the reference's own code.  The function is defined in
// mcut_shim.cpp next to the interposed build_oibvh()/intersectOIBVHs(), whose device-resident trees and candidate pairs it
// reuses.  It fills exactly the containers the rest of dispatch() reads.
#pragma once

#include <map>
#include <unordered_map>
#include <utility>
#include <vector>

#include "mcut/internal/hmesh.h"
#include "mcut/internal/math.h"
#include "mcut/internal/utils.h"

enum mcb200_hook_status {
    MCB200_HOOK_OK = 0,
    MCB200_HOOK_INVALID_SRC_MESH = 1, // kernel.cpp:2301-2312
    MCB200_HOOK_INVALID_CUT_MESH = 2,
    MCB200_HOOK_GENERAL_POSITION_VIOLATION = 3 // kernel.cpp:2543-2551, :2588-2597
};

// how many helper threads the calling dispatch() has (input.scheduler->get_num_threads()): the reference's registry order
// depends on it (its parallel_for block layout), and the hook reproduces that order
void mcb200_hook_set_scheduler_threads(uint32_t helper_threads);

// the face AABBs of the dispatch (input.*_hmesh_face_aabb_array_ptr).  The device hook ignores them (its boxes never left the
// GPU); the CPU stand-in of the tests (oracle/hook_oracle.cpp) needs them for the edge-box cull.
void mcb200_hook_set_face_boxes(const std::vector<bounding_box_t<vec3_<double>>>* src_boxes,
    const std::vector<bounding_box_t<vec3_<double>>>* cut_boxes);

int mcb200_hook_narrowphase(
    const hmesh_t& ps, // polygon soup: source mesh + cut-mesh faces (kernel.cpp:1593-1732)
    int sm_vtx_cnt, int sm_face_count,
    const std::map<fd_t, std::vector<fd_t>>& ps_face_to_potentially_intersecting_others, // what intersectOIBVHs produced
    hmesh_t& m0, // receives one vertex per intersection point, in registry order (kernel.cpp:2686-2689)
    std::unordered_map<fd_t, vec3>& ps_tested_face_to_plane_normal, // kernel.cpp:2188-2191
    std::unordered_map<fd_t, scalar_t>& ps_tested_face_to_plane_normal_d_param,
    std::unordered_map<fd_t, int>& ps_tested_face_to_plane_normal_max_comp,
    std::unordered_map<fd_t, std::vector<vec3>>& ps_tested_face_to_vertices,
    std::vector<std::pair<ed_t, fd_t>>& m0_ivtx_to_intersection_registry_entry, // kernel.cpp:2377-2412
    std::vector<vd_t>& cm_border_reentrant_ivtx_list,
    std::unordered_map<ed_t, std::vector<vd_t>>& ps_intersecting_edges,
    std::map<pair<fd_t>, std::vector<vd_t>>& cutpath_edge_creation_info,
    std::unordered_map<fd_t, std::vector<vd_t>>& ps_iface_to_ivtx_list,
    bool& partial_cut_detected,
    int& bad_face); // polygon-soup id of the degenerate face when the status is INVALID_*_MESH
