// mcut_b200/csrc/host_logic.cpp — the host-side parts of the intersect stage that must stay sequential.
//
//  * mcb200_vertex_parameters / mcb200_cut_bbox_eps: the frame of the internal coordinates.  The centre of mass is a
//    left-to-right double sum in the reference (source/preproc.cpp:2143-2196); a parallel reduction would change its
//    bits and with them every coordinate downstream (SURVEY §7 hard part 2), so it is computed here, on the host, in
//    the same order.  The per-vertex application of the frame runs on the device (common.cuh: load_vertex).
//  * mcb200_soup_ids: polygon-soup numbering.  The reference's `ps` is a copy of the source half-edge mesh with the
//    cut mesh's faces appended through add_face() (source/kernel.cpp:1593-1732); edge ids are handed out the first time
//    an unordered vertex pair is met while walking faces in order (source/hmesh.cpp:406-651) and the vertex list of a
//    face is the list of its halfedge TARGETS, i.e. rotated by one with respect to what add_face() was given
//    (source/hmesh.cpp:705-733) — once for source faces, twice for cut faces (kernel.cpp:1678 feeds add_face with an
//    already rotated list).  Pure integer work; O(halfedges) with one open-addressing table.
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/mcut_b200.h"

// also declared in internal.h (not included here: this file is plain C++, no CUDA headers)
int host_soup_ids(uint32_t nsv, const uint32_t* src_off, const uint32_t* src_vtx, uint32_t nsf, const uint32_t* cut_off,
    const uint32_t* cut_vtx, uint32_t ncf, uint32_t* face_vtx, uint32_t* face_edge, uint32_t* edge_v, uint32_t* edge_f,
    uint32_t* ne);

namespace {

struct stats_t {
    double mn[3], mx[3], mean[3];
};

template <typename T> stats_t scan_vertices(const T* p, uint32_t n)
{
    stats_t s;
    for (int j = 0; j < 3; ++j) {
        s.mn[j] = DBL_MAX;
        s.mx[j] = -DBL_MAX;
        s.mean[j] = 0.0;
    }
    for (uint32_t v = 0; v < n; ++v) {
        for (int j = 0; j < 3; ++j) {
            const T c = p[3 * (size_t)v + j];
            // the running extrema pass through T (float input: static_cast<float>(bbox), preproc.cpp:2160-2161)
            const T hi = static_cast<T>(s.mx[j]);
            const T lo = static_cast<T>(s.mn[j]);
            s.mx[j] = static_cast<double>(hi < c ? c : hi);
            s.mn[j] = static_cast<double>(c < lo ? c : lo);
            s.mean[j] += static_cast<double>(c);
        }
    }
    for (int j = 0; j < 3; ++j) s.mean[j] = s.mean[j] / static_cast<double>(n);
    return s;
}

inline uint64_t mix(uint64_t x)
{
    x ^= x >> 31;
    x *= 0x7fb5d329728ea185ULL;
    x ^= x >> 27;
    x *= 0x81dadef4bc2dd44dULL;
    x ^= x >> 33;
    return x;
}

} // namespace

extern "C" void mcb200_vertex_parameters(int is_float, const void* src_xyz, uint32_t nsv, const void* cut_xyz, uint32_t ncv,
    double com[3], double shift[3], double src_bbox[6], double cut_bbox[6])
{
    const stats_t s = is_float ? scan_vertices(static_cast<const float*>(src_xyz), nsv)
                               : scan_vertices(static_cast<const double*>(src_xyz), nsv);
    const stats_t c = is_float ? scan_vertices(static_cast<const float*>(cut_xyz), ncv)
                               : scan_vertices(static_cast<const double*>(cut_xyz), ncv);
    double to_positive[3];
    for (int j = 0; j < 3; ++j) {
        com[j] = (s.mean[j] + c.mean[j]) / 2.0; // preproc.cpp:2215
        const double lo = c.mn[j] < s.mn[j] ? c.mn[j] : s.mn[j];
        to_positive[j] = com[j] - lo; // :2221
    }
    double len2 = 0.0; // dot_product accumulates from 0.0 (math.h:634-642)
    for (int j = 0; j < 3; ++j) len2 += to_positive[j] * to_positive[j];
    const double len = std::sqrt(len2);
    for (int j = 0; j < 3; ++j) shift[j] = to_positive[j] + to_positive[j] / len; // :2222-2225
    for (int j = 0; j < 3; ++j) { // :2241-2246
        src_bbox[j] = s.mn[j] + shift[j];
        src_bbox[3 + j] = s.mx[j] + shift[j];
        cut_bbox[j] = c.mn[j] + shift[j];
        cut_bbox[3 + j] = c.mx[j] + shift[j];
    }
}

extern "C" double mcb200_cut_bbox_eps(const double cut_bbox[6], double gp_constant, int absolute)
{
    double s = 0.0;
    for (int j = 0; j < 3; ++j) {
        const double d = cut_bbox[3 + j] - cut_bbox[j];
        s += d * d;
    }
    const double scalar = absolute ? 1.0 : std::sqrt(s); // preproc.cpp:2518, :2667-2672
    return scalar * gp_constant;
}

int host_soup_ids(uint32_t nsv, const uint32_t* src_off, const uint32_t* src_vtx, uint32_t nsf, const uint32_t* cut_off,
    const uint32_t* cut_vtx, uint32_t ncf, uint32_t* face_vtx, uint32_t* face_edge, uint32_t* edge_v, uint32_t* edge_f,
    uint32_t* ne_out)
{
    const uint32_t nh = src_off[nsf] + cut_off[ncf];
    size_t cap = 16;
    while (cap < 2 * (size_t)nh) cap <<= 1;
    std::vector<uint64_t> keys(cap, 0);
    std::vector<uint32_t> vals(cap, 0);
    uint32_t ne = 0, h = 0;
    for (uint32_t f = 0; f < nsf + ncf; ++f) {
        const bool cutf = f >= nsf;
        const uint32_t* list = cutf ? cut_vtx + cut_off[f - nsf] : src_vtx + src_off[f];
        const uint32_t n = cutf ? cut_off[f - nsf + 1] - cut_off[f - nsf] : src_off[f + 1] - src_off[f];
        if (n < 3) return MCB200_ERR_INVALID;
        const uint32_t base = cutf ? nsv : 0u;
        const uint32_t pre_rot = cutf ? 1u : 0u; // cut faces reach add_face already rotated once
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t from = list[(i + pre_rot) % n] + base;
            const uint32_t to = list[(i + pre_rot + 1) % n] + base;
            const uint64_t key = ((static_cast<uint64_t>(from < to ? from : to) << 32) | (from < to ? to : from)) + 1;
            size_t slot = static_cast<size_t>(mix(key)) & (cap - 1);
            while (keys[slot] != 0 && keys[slot] != key) slot = (slot + 1) & (cap - 1);
            uint32_t e;
            if (keys[slot] == 0) { // first use: new edge, its h0 runs from -> to and belongs to f
                e = ne++;
                keys[slot] = key;
                vals[slot] = e;
                edge_v[2 * (size_t)e] = from;
                edge_v[2 * (size_t)e + 1] = to;
                edge_f[2 * (size_t)e] = f;
                edge_f[2 * (size_t)e + 1] = MCB200_NULL;
            } else {
                e = vals[slot];
                // a second face must use the opposite halfedge, and only once (hmesh.cpp:612-628)
                if (edge_v[2 * (size_t)e] == from || edge_f[2 * (size_t)e + 1] != MCB200_NULL) return MCB200_ERR_NON_MANIFOLD;
                edge_f[2 * (size_t)e + 1] = f;
            }
            face_vtx[h] = to; // hmesh.cpp:705-733
            face_edge[h] = e;
            ++h;
        }
    }
    *ne_out = ne;
    return 0;
}

extern "C" int mcb200_soup_ids(uint32_t nsv, const uint32_t* src_off, const uint32_t* src_vtx, uint32_t nsf,
    const uint32_t* cut_off, const uint32_t* cut_vtx, uint32_t ncf, uint32_t* face_vtx, uint32_t* face_edge, uint32_t* edge_v,
    uint32_t* edge_f, uint32_t* ne)
{
    if (!src_off || !src_vtx || !cut_off || !cut_vtx || !face_vtx || !face_edge || !edge_v || !edge_f || !ne)
        return MCB200_ERR_INVALID;
    return host_soup_ids(nsv, src_off, src_vtx, nsf, cut_off, cut_vtx, ncf, face_vtx, face_edge, edge_v, edge_f, ne);
}

// ---- the reference's registry order -------------------------------------------------------------------------------------
// In which order does the reference register intersection points, i.e. how does it number the intersection vertices of
// m0?  It is an accident of its containers, but everything downstream of the narrowphase — including which edges the
// floating-polygon resolution computes its partition segment from (preproc.cpp:1000-1126) — follows that numbering, so a
// drop-in has to reproduce it:
//   * kernel.cpp:1781-1852 fills std::unordered_map<ed_t, ...> ps_edge_face_intersection_pairs: the candidate faces (keys of
//     a std::map: ascending) are cut into blocks by parallel_for (tpool.h:354-392, :420-472: `helpers + 1` threads, at least
//     1024 elements each); a block inserts the edges of its faces in first-seen order into a local map; the MASTER's
//     block — the last one — becomes the map, the other blocks' edges are then inserted in block order;
//   * kernel.cpp:2415-2672 walks that map in ITERATION order, again in blocks, and concatenates the per-block registries
//     master block first (:2673-2868); an edge's faces are visited in ascending id.
// The iteration order of a libstdc++ unordered_map is a function of the insertion sequence (identity hash of the
// descriptor's 32-bit index, hmesh.h:628-634; the mapped type plays no part), so replaying the sequence into a map of
// 32-bit keys built by the same toolchain yields the same order.  rank[e] = position of edge e in the registry order, or
// MCB200_NULL for edges of no candidate face.
#include <algorithm>
#include <unordered_map>

extern "C" int mcb200_reference_edge_rank(uint32_t n_cand_faces, const uint32_t* cand_faces, const uint32_t* face_off,
    const uint32_t* face_edge, uint32_t ne, uint32_t helper_threads, uint32_t* rank)
{
    if ((n_cand_faces && !cand_faces) || !face_edge || (ne && !rank)) return MCB200_ERR_INVALID;
    for (uint32_t e = 0; e < ne; ++e) rank[e] = MCB200_NULL;
    if (n_cand_faces == 0) return 0;
    for (uint32_t i = 1; i < n_cand_faces; ++i)
        if (cand_faces[i] <= cand_faces[i - 1]) return MCB200_ERR_INVALID; // keys of a std::map
    typedef std::unordered_map<uint32_t, char> edge_set_t;
    const uint32_t available = helper_threads + 1u;
    auto schedule = [&](uint32_t length, uint32_t& nthreads, uint32_t& block_size) { // get_scheduling_parameters, tpool.h:354-392
        const uint32_t max_threads = (length + 1023u) / 1024u;
        nthreads = std::min(available, max_threads);
        block_size = length / nthreads;
    };
    auto block = [&](uint32_t a, uint32_t b) {
        edge_set_t local;
        for (uint32_t i = a; i < b; ++i) {
            const uint32_t f = cand_faces[i];
            const uint32_t h0 = face_off ? face_off[f] : 3u * f, h1 = face_off ? face_off[f + 1] : 3u * f + 3u;
            for (uint32_t h = h0; h < h1; ++h) local[face_edge[h]];
        }
        return local;
    };
    uint32_t nthreads = 1, block_size = 0;
    schedule(n_cand_faces, nthreads, block_size);
    std::vector<edge_set_t> futures(nthreads - 1);
    uint32_t block_start = 0;
    for (uint32_t i = 0; i + 1 < nthreads; ++i) {
        futures[i] = block(block_start, block_start + block_size);
        block_start += block_size;
    }
    edge_set_t all = block(block_start, n_cand_faces);
    for (const edge_set_t& f : futures)
        for (edge_set_t::const_iterator i = f.cbegin(); i != f.cend(); ++i)
            if (all.find(i->first) == all.cend()) all[i->first] = i->second;
    std::vector<uint32_t> order;
    order.reserve(all.size());
    for (edge_set_t::const_iterator i = all.cbegin(); i != all.cend(); ++i) {
        if (i->first >= ne) return MCB200_ERR_INVALID;
        order.push_back(i->first);
    }
    schedule((uint32_t)order.size(), nthreads, block_size);
    const size_t master_first = (size_t)(nthreads - 1) * block_size;
    uint32_t next = 0;
    for (size_t k = master_first; k < order.size(); ++k) rank[order[k]] = next++;
    for (size_t k = 0; k < master_first; ++k) rank[order[k]] = next++;
    return 0;
}
