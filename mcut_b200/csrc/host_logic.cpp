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

// The frame in two steps: per-mesh statistics (a sequential pass over the vertices: min, max, mean), then their combination.
// A caller that cuts the SAME source mesh again and again (planar sections: 256 planes through one terrain) keeps the
// source's statistics and only scans the three vertices of each new cut mesh.
extern "C" void mcb200_vertex_stats(int is_float, const void* xyz, uint32_t nv, double stats[9])
{
    const stats_t s = is_float ? scan_vertices(static_cast<const float*>(xyz), nv) : scan_vertices(static_cast<const double*>(xyz), nv);
    for (int j = 0; j < 3; ++j) {
        stats[j] = s.mn[j];
        stats[3 + j] = s.mx[j];
        stats[6 + j] = s.mean[j];
    }
}

extern "C" void mcb200_vertex_parameters_from_stats(const double s[9], const double c[9], double com[3], double shift[3],
    double src_bbox[6], double cut_bbox[6])
{
    double to_positive[3];
    for (int j = 0; j < 3; ++j) {
        com[j] = (s[6 + j] + c[6 + j]) / 2.0; // preproc.cpp:2215
        const double lo = c[j] < s[j] ? c[j] : s[j];
        to_positive[j] = com[j] - lo; // :2221
    }
    double len2 = 0.0; // dot_product accumulates from 0.0 (math.h:634-642)
    for (int j = 0; j < 3; ++j) len2 += to_positive[j] * to_positive[j];
    const double len = std::sqrt(len2);
    for (int j = 0; j < 3; ++j) shift[j] = to_positive[j] + to_positive[j] / len; // :2222-2225
    for (int j = 0; j < 3; ++j) { // :2241-2246
        src_bbox[j] = s[j] + shift[j];
        src_bbox[3 + j] = s[3 + j] + shift[j];
        cut_bbox[j] = c[j] + shift[j];
        cut_bbox[3 + j] = c[3 + j] + shift[j];
    }
}

extern "C" void mcb200_vertex_parameters(int is_float, const void* src_xyz, uint32_t nsv, const void* cut_xyz, uint32_t ncv,
    double com[3], double shift[3], double src_bbox[6], double cut_bbox[6])
{
    double s[9], c[9];
    mcb200_vertex_stats(is_float, src_xyz, nsv, s);
    mcb200_vertex_stats(is_float, cut_xyz, ncv, c);
    mcb200_vertex_parameters_from_stats(s, c, com, shift, src_bbox, cut_bbox);
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

namespace {
// node storage for the replayed maps: one arena, released at once (the iteration order does not depend on the allocator)
struct arena_t {
    std::vector<char*> blocks;
    size_t left = 0;
    char* cur = nullptr;
    ~arena_t()
    {
        for (char* b : blocks) ::operator delete(b);
    }
    void* take(size_t bytes)
    {
        bytes = (bytes + 15u) & ~(size_t)15u;
        if (bytes > left) {
            const size_t block = std::max<size_t>(bytes, (size_t)1 << 20);
            cur = static_cast<char*>(::operator new(block));
            blocks.push_back(cur);
            left = block;
        }
        void* p = cur;
        cur += bytes;
        left -= bytes;
        return p;
    }
};
template <typename T>
struct arena_alloc {
    typedef T value_type;
    arena_t* a;
    explicit arena_alloc(arena_t* a_) : a(a_) {}
    template <typename U>
    arena_alloc(const arena_alloc<U>& o) : a(o.a) {}
    T* allocate(size_t n) { return static_cast<T*>(a->take(n * sizeof(T))); }
    void deallocate(T*, size_t) {}
    template <typename U>
    bool operator==(const arena_alloc<U>& o) const { return a == o.a; }
    template <typename U>
    bool operator!=(const arena_alloc<U>& o) const { return a != o.a; }
};
}

extern "C" int mcb200_reference_edge_order(uint32_t n_cand_faces, const uint32_t* slot_off, const uint32_t* slot_edge,
    uint32_t helper_threads, uint32_t* order, uint32_t* n_order)
{
    if (!n_order || (n_cand_faces && (!slot_edge || !order))) return MCB200_ERR_INVALID;
    *n_order = 0;
    if (n_cand_faces == 0) return 0;
    arena_t arena;
    typedef std::unordered_map<uint32_t, char, std::hash<uint32_t>, std::equal_to<uint32_t>, arena_alloc<std::pair<const uint32_t, char>>> edge_set_t;
    const arena_alloc<std::pair<const uint32_t, char>> alloc(&arena);
    const uint32_t available = helper_threads + 1u;
    auto schedule = [&](uint32_t length, uint32_t& nthreads, uint32_t& block_size) { // get_scheduling_parameters, tpool.h:354-392
        const uint32_t max_threads = (length + 1023u) / 1024u;
        nthreads = std::min(available, max_threads);
        block_size = length / nthreads;
    };
    auto block = [&](edge_set_t& local, uint32_t a, uint32_t b) {
        for (uint32_t i = a; i < b; ++i) {
            const uint32_t h0 = slot_off ? slot_off[i] : 3u * i, h1 = slot_off ? slot_off[i + 1] : 3u * i + 3u;
            for (uint32_t h = h0; h < h1; ++h) local[slot_edge[h]];
        }
    };
    uint32_t nthreads = 1, block_size = 0;
    schedule(n_cand_faces, nthreads, block_size);
    std::vector<edge_set_t> futures;
    futures.reserve(nthreads - 1);
    uint32_t block_start = 0;
    for (uint32_t i = 0; i + 1 < nthreads; ++i) {
        futures.emplace_back(alloc); // as default-constructed: same bucket growth
        block(futures.back(), block_start, block_start + block_size);
        block_start += block_size;
    }
    edge_set_t all(alloc);
    block(all, block_start, n_cand_faces);
    for (const edge_set_t& f : futures)
        for (edge_set_t::const_iterator i = f.cbegin(); i != f.cend(); ++i)
            if (all.find(i->first) == all.cend()) all[i->first] = i->second;
    const uint32_t n = (uint32_t)all.size();
    schedule(n, nthreads, block_size);
    const uint32_t master_first = (nthreads - 1) * block_size;
    // the map's iteration order, rotated: the master's block (the last one) registers first
    uint32_t k = 0;
    for (edge_set_t::const_iterator i = all.cbegin(); i != all.cend(); ++i, ++k)
        order[k >= master_first ? k - master_first : k + (n - master_first)] = i->first;
    *n_order = n;
    return 0;
}

extern "C" int mcb200_reference_edge_rank(uint32_t n_cand_faces, const uint32_t* cand_faces, const uint32_t* face_off,
    const uint32_t* face_edge, uint32_t ne, uint32_t helper_threads, uint32_t* rank)
{
    if ((n_cand_faces && !cand_faces) || !face_edge || (ne && !rank)) return MCB200_ERR_INVALID;
    for (uint32_t e = 0; e < ne; ++e) rank[e] = MCB200_NULL;
    if (n_cand_faces == 0) return 0;
    for (uint32_t i = 1; i < n_cand_faces; ++i)
        if (cand_faces[i] <= cand_faces[i - 1]) return MCB200_ERR_INVALID; // keys of a std::map
    std::vector<uint32_t> off((size_t)n_cand_faces + 1, 0u), slots;
    for (uint32_t i = 0; i < n_cand_faces; ++i) {
        const uint32_t f = cand_faces[i];
        const uint32_t h0 = face_off ? face_off[f] : 3u * f, h1 = face_off ? face_off[f + 1] : 3u * f + 3u;
        for (uint32_t h = h0; h < h1; ++h) {
            if (face_edge[h] >= ne) return MCB200_ERR_INVALID;
            slots.push_back(face_edge[h]);
        }
        off[i + 1] = (uint32_t)slots.size();
    }
    std::vector<uint32_t> order(slots.size() ? slots.size() : 1u);
    uint32_t n = 0;
    const int rc = mcb200_reference_edge_order(n_cand_faces, off.data(), slots.data(), helper_threads, order.data(), &n);
    if (rc) return rc;
    for (uint32_t k = 0; k < n; ++k) rank[order[k]] = k;
    return 0;
}
