// mcut_b200/csrc/lbvh.cu — (1) LBVH construction on the device.
//
// Replaces build_oibvh() (include/mcut/internal/bvh.h:117-125, source/bvh.cpp:219-636):
//   k_face_bbox    face AABBs (+eps enlargement) and the mesh AABB               bvh.cpp:242-368, math.h:866-928
//   k_morton       30-bit Morton codes, the reference's float formula;           bvh.cpp:196-217, :373-433
//                  also the digit histograms of the sort's passes
//   radix passes   one-sweep radix sort of (code, face)  (radix_sort.cuh)        bvh.cpp:437-442 (std::sort there)
//   k_tree         radix tree over the sorted codes (Karras 2012) fused with     replaces the implicit OIBVH layout :444-493
//                  the box refit of every subtree of <= 32 leaves                and the bottom of bvh.cpp:498-635
//   k_refit_climb  atomic bottom-up refit of the nodes above those subtrees      bvh.cpp:498-635 (one parallel_for per level there)
// Only face AABBs, the mesh AABB and the leaf-pair SET escape this stage, and every internal box is the exact
// min/max union of its leaves, so the tree shape is free: an LBVH yields the same pairs.
#include "internal.h"
#include "radix_sort.cuh"

namespace {

constexpr int BLOCK = 256;

__device__ __forceinline__ unsigned spread10(unsigned v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__device__ __forceinline__ unsigned morton3D(float x, float y, float z)
{
    // bvh.cpp:206-217; fminf/fmaxf have std::fmin/std::fmax semantics (NaN -> the other operand)
    x = fminf(fmaxf(__fmul_rn(x, 1024.0f), 0.0f), 1023.0f);
    y = fminf(fmaxf(__fmul_rn(y, 1024.0f), 0.0f), 1023.0f);
    z = fminf(fmaxf(__fmul_rn(z, 1024.0f), 0.0f), 1023.0f);
    return spread10((unsigned)x) * 4u + spread10((unsigned)y) * 2u + spread10((unsigned)z);
}

// ---- K_aabb -----------------------------------------------------------------------------------------------------
// One thread per face, grid-stride; the block's union goes to the mesh AABB through 6 ordered-integer atomics.
template <bool TRI>
__global__ void __launch_bounds__(BLOCK) k_face_bbox(const void* __restrict__ xyz, frame_t fr,
    const uint32_t* __restrict__ face_vtx, const uint32_t* __restrict__ face_off, uint32_t nf, double eps,
    double* __restrict__ face_bbox, unsigned long long* __restrict__ root_ordered, unsigned* __restrict__ arrival_flags,
    const double* __restrict__ prior, uint32_t n_prior)
{
    pdl_prologue();
    double bmin[3] = { DBL_MAX, DBL_MAX, DBL_MAX }, bmax[3] = { -DBL_MAX, -DBL_MAX, -DBL_MAX };
    for (uint32_t f = blockIdx.x * BLOCK + threadIdx.x; f < nf; f += gridDim.x * BLOCK) {
        const uint32_t h0 = TRI ? 3u * f : face_off[f];
        const uint32_t h1 = TRI ? h0 + 3u : face_off[f + 1];
        double mn[3] = { DBL_MAX, DBL_MAX, DBL_MAX }, mx[3] = { -DBL_MAX, -DBL_MAX, -DBL_MAX };
        for (uint32_t h = h0; h < h1; ++h) {
            double p[3];
            load_vertex(xyz, fr, __ldg(face_vtx + h), p);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                mx[j] = ref_max(mx[j], p[j]);
                mn[j] = ref_min(mn[j], p[j]);
            }
        }
        if (f < n_prior) {
            // build_oibvh() only RESIZES the caller's face_bboxes (bvh.cpp:242) and expands what is there: on the rebuild
            // after a floating-polygon repartition (preproc.cpp:2733-2760) a face keeps the (enlarged) box it had
            const double* pb = prior + 6 * (size_t)f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                mn[j] = ref_min(__ldg(pb + j), mn[j]);
                mx[j] = ref_max(__ldg(pb + 3 + j), mx[j]);
            }
        }
        if (eps > 0.0) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                mx[j] = __dadd_rn(mx[j], eps);
                mn[j] = __dsub_rn(mn[j], eps);
            }
        }
        arrival_flags[f] = 0u; // the refit's arrival counter of node f (saves a memset pass over the array)
        double* out = face_bbox + 6 * (size_t)f;
        // 48-byte rows: three 16-byte stores
        reinterpret_cast<double2*>(out)[0] = make_double2(mn[0], mn[1]);
        reinterpret_cast<double2*>(out)[1] = make_double2(mn[2], mx[0]);
        reinterpret_cast<double2*>(out)[2] = make_double2(mx[1], mx[2]);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            bmax[j] = ref_max(bmax[j], mx[j]);
            bmin[j] = ref_min(bmin[j], mn[j]);
        }
    }
    // warp reduce, then one lane per warp hits the 6 global words (persistent grid => few atomics)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bmin[j] = fmin(bmin[j], __shfl_xor_sync(0xffffffffu, bmin[j], o));
            bmax[j] = fmax(bmax[j], __shfl_xor_sync(0xffffffffu, bmax[j], o));
        }
    }
    __shared__ double s_min[BLOCK / 32][3], s_max[BLOCK / 32][3];
    const unsigned w = threadIdx.x >> 5;
    if (lane_id() == 0) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            s_min[w][j] = bmin[j];
            s_max[w][j] = bmax[j];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int j = threadIdx.x % 3;
        if (threadIdx.x < 3) {
            double v = s_min[0][j];
            for (int i = 1; i < BLOCK / 32; ++i) v = fmin(v, s_min[i][j]);
            if (v != DBL_MAX) atomicMin(root_ordered + j, dbl_to_ordered(v));
        } else {
            double v = s_max[0][j];
            for (int i = 1; i < BLOCK / 32; ++i) v = fmax(v, s_max[i][j]);
            if (v != -DBL_MAX) atomicMax(root_ordered + 3 + j, dbl_to_ordered(v));
        }
    }
}

// ---- K_morton ---------------------------------------------------------------------------------------------------
// Codes by face + the same codes as sort keys; the digit histograms of all radix passes of the sort are accumulated here (shared
// memory, one flush per block) and the sort's look-back status words are cleared, so the sort needs no histogram kernel
// and no second read of the keys.
__global__ void __launch_bounds__(BLOCK) k_morton(const double* __restrict__ face_bbox, uint32_t nf,
    const unsigned long long* __restrict__ root_ordered, double* __restrict__ root_decoded, uint32_t* __restrict__ codes,
    uint32_t* __restrict__ sort_keys, unsigned* __restrict__ hist /* [4][256] */, unsigned* __restrict__ status,
    unsigned status_words, unsigned key_shift, int npasses)
{
    pdl_prologue();
    __shared__ unsigned s_hist[4 * 256];
    for (int i = threadIdx.x; i < 4 * 256; i += BLOCK) s_hist[i] = 0;
    for (unsigned i = blockIdx.x * BLOCK + threadIdx.x; i < status_words; i += gridDim.x * BLOCK) status[i] = 0u;
    __syncthreads();
    double rmin[3], dims[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        rmin[j] = ordered_to_dbl(root_ordered[j]);
        dims[j] = __dsub_rn(ordered_to_dbl(root_ordered[3 + j]), rmin[j]);
    }
    if (blockIdx.x == 0 && threadIdx.x < 6) root_decoded[threadIdx.x] = ordered_to_dbl(root_ordered[threadIdx.x]);
    for (uint32_t f = blockIdx.x * BLOCK + threadIdx.x; f < nf; f += gridDim.x * BLOCK) {
        const double2* in = reinterpret_cast<const double2*>(face_bbox + 6 * (size_t)f);
        const double2 a = in[0], b = in[1], c = in[2];
        const double mn[3] = { a.x, a.y, b.x }, mx[3] = { b.y, c.x, c.y };
        float nrm[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double centre = __dadd_rn(mn[j], mx[j]) / 2; // bvh.cpp:275
            const double off = __dsub_rn(centre, rmin[j]); // bvh.cpp:382
            nrm[j] = (float)(off / dims[j]); // bvh.cpp:399-402
        }
        const uint32_t code = morton3D(nrm[0], nrm[1], nrm[2]);
        codes[f] = code;
        const uint32_t key = code >> key_shift; // the leaves are ordered by the top bits of the code (lbvh_build)
        sort_keys[f] = key;
#pragma unroll
        for (int p = 0; p < 4; ++p)
            if (p < npasses) atomicAdd(&s_hist[p * 256 + ((key >> (8 * p)) & 255u)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * 256; i += BLOCK) {
        const unsigned c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// ---- Karras topology (used by k_tree) ---------------------------------------------------------------------------
// One thread per internal node (Karras 2012).  A block owns 256 consecutive nodes and stages the sorted codes of a
// +-512 window in shared memory: the range/split binary searches of almost every node (every range up to 256 leaves,
// the doubling search overshoots by 2x) stay inside that window, so their dependent probes cost shared-memory latency
// instead of L2 latency; the few larger nodes fall back to global loads.
constexpr int KHALO = 512;
constexpr int KWIN = BLOCK + 2 * KHALO;

struct code_window {
    const uint32_t* codes;
    const uint32_t* win; // shared copy of codes[base, base + KWIN)
    int base, n;
    __device__ __forceinline__ uint32_t get(int j) const
    {
        const int r = j - base;
        return ((unsigned)r < (unsigned)KWIN) ? win[r] : __ldg(codes + j);
    }
    // window-only probe: a key outside the shared window reads as "no common prefix" and raises `miss`
    __device__ __forceinline__ int delta_win(int i, uint32_t ci, int j, bool& miss) const
    {
        if (j < 0 || j >= n) return -1;
        const int r = j - base;
        if ((unsigned)r >= (unsigned)KWIN) {
            miss = true;
            return -1;
        }
        const uint32_t cj = win[r];
        if (ci == cj) return 32 + __clz((unsigned)i ^ (unsigned)j);
        return __clz(ci ^ cj);
    }
    __device__ __forceinline__ int delta(int i, uint32_t ci, int j) const
    {
        if (j < 0 || j >= n) return -1;
        const uint32_t cj = get(j);
        if (ci == cj) return 32 + __clz((unsigned)i ^ (unsigned)j); // duplicate codes: tie-break on the index
        return __clz(ci ^ cj);
    }
};

// ---- refit helpers ------------------------------------------------------------------------------------------------
// One thread per leaf carries its box up; the first thread to reach a node parks its box in the node and leaves,
// the second one merges and continues (atomic arrival counter per internal node).
__device__ __forceinline__ void store_box(float* dst, const float* b) // 24 bytes, 8-byte aligned
{
    reinterpret_cast<float2*>(dst)[0] = make_float2(b[0], b[1]);
    reinterpret_cast<float2*>(dst)[1] = make_float2(b[2], b[3]);
    reinterpret_cast<float2*>(dst)[2] = make_float2(b[4], b[5]);
}
__device__ __forceinline__ void load_box_cg(const float* src, float* b)
{
    // written by another SM moments ago: read through L2
    const float2 a = __ldcg(reinterpret_cast<const float2*>(src));
    const float2 c = __ldcg(reinterpret_cast<const float2*>(src) + 1);
    const float2 e = __ldcg(reinterpret_cast<const float2*>(src) + 2);
    b[0] = a.x;
    b[1] = a.y;
    b[2] = c.x;
    b[3] = c.y;
    b[4] = e.x;
    b[5] = e.y;
}

__device__ __forceinline__ void load_face_box(const double* __restrict__ face_bbox, uint32_t face, double* b)
{
    const double2* in = reinterpret_cast<const double2*>(face_bbox + 6 * (size_t)face);
    const double2 a = __ldg(in), c = __ldg(in + 1), e = __ldg(in + 2);
    b[0] = a.x;
    b[1] = a.y;
    b[2] = c.x;
    b[3] = c.y;
    b[4] = e.x;
    b[5] = e.y;
}

// Range, split and direction of internal node i (Karras 2012, Fig. 4).  WINDOW_ONLY: probes never leave the shared code
// window; a search that would is abandoned with `miss` set (such a node covers more than 512 leaves) and is redone later
// with global probes, off the block's common path.
struct node_topology {
    int lo, hi, gamma, d, dmin;
};
template <bool WINDOW_ONLY> __device__ __forceinline__ node_topology karras_node(const code_window& cw, int ii, bool& miss)
{
    auto dl = [&](int j) -> int { return WINDOW_ONLY ? cw.delta_win(ii, cw.get(ii), j, miss) : cw.delta(ii, cw.get(ii), j); };
    node_topology t;
    t.d = (dl(ii + 1) - dl(ii - 1)) >= 0 ? 1 : -1;
    t.dmin = dl(ii - t.d); // = length of the parent's prefix
    int lmax = 2;
    while (dl(ii + lmax * t.d) > t.dmin) lmax <<= 1;
    int l = 0;
    for (int s = lmax >> 1; s >= 1; s >>= 1)
        if (dl(ii + (l + s) * t.d) > t.dmin) l += s;
    const int j = ii + l * t.d;
    const int dnode = dl(j);
    int sp = 0;
    int s = l;
    do {
        s = (s + 1) >> 1;
        if (dl(ii + (sp + s) * t.d) > dnode) sp += s;
    } while (s > 1);
    t.gamma = ii + sp * t.d + (t.d < 0 ? -1 : 0);
    t.lo = ii < j ? ii : j;
    t.hi = ii < j ? j : ii;
    return t;
}

// `face_of(j)` = face id of sorted leaf j.  A leaf child is recorded by its FACE id (the traversal needs nothing else of
// it); the parent word of a leaf still lives at its sorted position.
template <typename FaceOf>
__device__ __forceinline__ void write_topology(bvh_node_t* nodes, uint32_t* parent, uint32_t nf, uint32_t i, const node_topology& t,
    FaceOf face_of)
{
    const bool lleaf = (t.lo == t.gamma), rleaf = (t.hi == t.gamma + 1);
    const uint32_t left = lleaf ? (MCB_LEAF_BIT | face_of(t.gamma)) : (uint32_t)t.gamma;
    const uint32_t right = rleaf ? (MCB_LEAF_BIT | face_of(t.gamma + 1)) : (uint32_t)(t.gamma + 1);
    // parent word of a child: (parent index << 2) | (child is the right one); read by the climb
    const uint32_t pw = i << 2;
    parent[lleaf ? (nf - 1 + (uint32_t)t.gamma) : (uint32_t)t.gamma] = pw;
    parent[rleaf ? (nf - 1 + (uint32_t)(t.gamma + 1)) : (uint32_t)(t.gamma + 1)] = pw | 1u;
    if (i == 0) parent[0] = MCB200_NULL;
    *reinterpret_cast<uint4*>(&nodes[i].left) = make_uint4(left, right, (uint32_t)t.lo, (uint32_t)t.hi);
}

// Slots in the group list for a whole block with ONE global atomic (tens of thousands of same-address atomics from
// individual warps serialise in L2 and were the most expensive part of this kernel).  Every thread of the block must
// call this; `want` is 0, 1 or 2.  Returns the first slot of the calling thread.
__device__ __forceinline__ unsigned alloc_groups_block(unsigned* n_groups, unsigned want, unsigned* s_warp /*[BLOCK/32]*/,
    unsigned* s_base)
{
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
    unsigned inc = want;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int i = 0; i < BLOCK / 32; ++i) {
            const unsigned c = s_warp[i];
            s_warp[i] = tot;
            tot += c;
        }
        *s_base = tot ? atomicAdd(n_groups, tot) : 0u;
    }
    __syncthreads();
    return *s_base + s_warp[w] + inc - want;
}

// Tree and refit in two regimes, two kernels:
//  * k_tree: one thread per internal node finds its range and split (Karras 2012) from the block's shared window of
//    sorted codes.  A node whose leaf range has at most 32 leaves then gets both child boxes straight from the leaf boxes
//    of its range (it knows its range and its split, so no other node's result is needed): no atomics, no fences, and the
//    whole 64-byte record (boxes + topology) is written at once.  Node i lies inside its own range, so every such range
//    falls in the block's leaf window [i0-32, i0+288), gathered into shared memory while the code window loads.  The
//    maximal treelets ("group roots") are listed: they are the traversal's query groups and the starting points of ...
//    Whether a node's PARENT covers more than 32 leaves is decided locally: the parent's range is the set of keys sharing
//    the parent's prefix (length delta(i, i - d), the `dmin` of the range search), so it has more than 32 leaves exactly
//    when the key 32 positions beyond the node's far end still shares that prefix.
//  * k_refit_climb: ... the classic atomic bottom-up pass for the ~nf/16 nodes above the treelets: one thread per group
//    root carries its box upwards; the first thread to reach a node parks its box there and leaves, the second one
//    merges and continues.  All climbers are resident at once, so the pass costs (levels above the treelets) x (one
//    store / fence / atomic / load round trip), not a block-scheduling queue.
constexpr int RHALO = 32;
constexpr int RWIN = BLOCK + 2 * RHALO;

// NODES = false: the mesh will only be the QUERY side of a traversal — it needs its groups (maximal treelets + union boxes)
// but nobody will walk its tree, so no node record, no parent word is written and the climb kernel is not run.
template <bool NODES>
__global__ void __launch_bounds__(BLOCK) k_tree(const uint32_t* __restrict__ codes, const double* __restrict__ face_bbox,
    const uint32_t* __restrict__ sorted_faces, uint32_t nf, bvh_node_t* nodes, uint32_t* __restrict__ parent,
    uint2* __restrict__ groups, group_up_t* __restrict__ group_up, unsigned* __restrict__ n_groups)
{
    pdl_prologue();
    __shared__ uint32_t s_win[KWIN];
    __shared__ float s_box[RWIN][6]; // conservative single-precision leaf boxes of the block's leaf window
    __shared__ uint32_t s_face[RWIN]; // their face ids
    __shared__ unsigned s_warp[BLOCK / 32], s_base;
    if (nf == 1) {
        // a single leaf (e.g. the planar-section triangle): pseudo-root whose right child can never be hit
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            double bd[6];
            load_face_box(face_bbox, sorted_faces[0], bd);
            float b[6];
            box_to_float(bd, b);
            store_box(nodes[0].lbox, b);
            const float e[6] = { FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX };
            store_box(nodes[0].rbox, e);
            nodes[0].left = MCB_LEAF_BIT | sorted_faces[0];
            nodes[0].right = MCB200_NULL;
            nodes[0].first = 0;
            nodes[0].last = 0;
            groups[0] = make_uint2(0u, 1u);
            store_box(group_up[0].box, b);
            group_up[0].pw = MCB200_NULL;
            *n_groups = 1u;
        }
        return;
    }
    const int n = (int)nf;
    const uint32_t i0 = blockIdx.x * BLOCK;
    const int cbase = (int)i0 - KHALO;
    {
        // both windows with every load in flight before the first store (a load->store loop would serialise L2 round trips)
        static_assert(KWIN % BLOCK == 0 && RWIN <= 2 * BLOCK, "window shapes");
        uint32_t cw_reg[KWIN / BLOCK];
#pragma unroll
        for (int k = 0; k < KWIN / BLOCK; ++k) {
            const int j = cbase + k * BLOCK + (int)threadIdx.x;
            cw_reg[k] = (j >= 0 && j < n) ? __ldg(codes + j) : 0u;
        }
        const long long j0 = (long long)i0 - RHALO + threadIdx.x, j1 = j0 + BLOCK;
        const bool have0 = j0 >= 0 && j0 < (long long)nf, have1 = threadIdx.x < RWIN - BLOCK && j1 < (long long)nf;
        const uint32_t f0 = have0 ? __ldg(sorted_faces + j0) : 0u, f1 = have1 ? __ldg(sorted_faces + j1) : 0u;
        double b0[6], b1[6];
        if (have0) load_face_box(face_bbox, f0, b0);
        if (have1) load_face_box(face_bbox, f1, b1);
#pragma unroll
        for (int k = 0; k < KWIN / BLOCK; ++k) s_win[k * BLOCK + threadIdx.x] = cw_reg[k];
        if (have0) {
            box_to_float(b0, s_box[threadIdx.x]);
            s_face[threadIdx.x] = f0;
        }
        if (have1) {
            box_to_float(b1, s_box[BLOCK + threadIdx.x]);
            s_face[BLOCK + threadIdx.x] = f1;
        }
    }
    __syncthreads();
    const uint32_t i = i0 + threadIdx.x;
    const int wbase = (int)i0 - RHALO; // leaf j sits in s_box[j - wbase]
    const code_window cw { codes, s_win, cbase, n };

    // face id of sorted leaf j: from the block's window when it is there (every leaf child of a small node is)
    auto face_of = [&](int j) -> uint32_t {
        const int r = j - wbase;
        return ((unsigned)r < (unsigned)RWIN) ? s_face[r] : __ldg(sorted_faces + j);
    };
    // ---- topology first: node i's range, split and children; is it (or leaf i) the root of a maximal treelet? ----
    bool root0 = false, root1 = false, small = false, deferred = false;
    int lo = 0, hi = 0, gamma = 0;
    if (i < nf - 1u) {
        const int ii = (int)i;
        const node_topology t = karras_node<true>(cw, ii, deferred);
        if (!deferred) {
            if (NODES) write_topology(nodes, parent, nf, i, t, face_of);
            lo = t.lo;
            hi = t.hi;
            gamma = t.gamma;
            small = hi - lo + 1 <= 32;
            if (small) {
                // a maximal treelet: the whole tree, or the parent's range reaches past 32 leaves
                const int probe = (t.d > 0) ? hi - 32 : lo + 32;
                root0 = (i == 0u) || (probe >= 0 && probe < n && cw.delta(ii, cw.get(ii), probe) >= t.dmin);
            }
        }
    }
    // leaf i hangs directly under a node that covers more than 32 leaves?  A leaf joins the neighbour it shares the longer
    // prefix with; that prefix is its parent's.
    if (i < nf) {
        const int ii = (int)i;
        const uint32_t ci = cw.get(ii);
        const int dl = cw.delta(ii, ci, ii - 1), dr = cw.delta(ii, ci, ii + 1);
        const bool is_left = dr > dl;
        const int dp = is_left ? dr : dl;
        const int probe = is_left ? ii + 32 : ii - 32;
        root1 = probe >= 0 && probe < n && cw.delta(ii, ci, probe) >= dp;
    }
    // the block-wide slot allocation (a barrier) sits here, before the box loops whose length differs from thread to thread
    unsigned g = alloc_groups_block(n_groups, (root0 ? 1u : 0u) + (root1 ? 1u : 0u), s_warp, &s_base);

    // ---- boxes of the small nodes straight from the leaf window ----
    if (small && (NODES || root0)) {
        float box[6], rb[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            box[k] = s_box[lo - wbase][k];
            rb[k] = s_box[gamma + 1 - wbase][k];
        }
        for (int q = lo + 1; q <= gamma; ++q)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                box[k] = fminf(box[k], s_box[q - wbase][k]);
                box[3 + k] = fmaxf(box[3 + k], s_box[q - wbase][3 + k]);
            }
        for (int q = gamma + 2; q <= hi; ++q)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                rb[k] = fminf(rb[k], s_box[q - wbase][k]);
                rb[3 + k] = fmaxf(rb[3 + k], s_box[q - wbase][3 + k]);
            }
        if (NODES) {
            store_box(nodes[i].lbox, box);
            store_box(nodes[i].rbox, rb);
        }
        if (root0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                box[k] = fminf(box[k], rb[k]);
                box[3 + k] = fmaxf(box[3 + k], rb[3 + k]);
            }
            groups[g] = make_uint2((uint32_t)lo, (uint32_t)(hi - lo + 1));
            store_box(group_up[g].box, box);
            group_up[g].pw = (i == 0u) ? MCB200_NULL : i; // where the climb finds this root's parent word
            ++g;
        }
    }
    if (root1) {
        groups[g] = make_uint2(i, 1u);
        float lb[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) lb[k] = s_box[(int)i - wbase][k];
        store_box(group_up[g].box, lb);
        group_up[g].pw = nf - 1u + i;
    }
    // the few nodes whose range leaves the window (> 512 leaves): dependent global probes, now that nobody waits for them
    if (NODES && deferred) {
        bool unused = false;
        const node_topology t = karras_node<false>(cw, (int)i, unused);
        write_topology(nodes, parent, nf, i, t, face_of);
    }
}

__global__ void __launch_bounds__(BLOCK) k_refit_climb(bvh_node_t* nodes, const uint32_t* __restrict__ parent, unsigned* flags,
    const group_up_t* __restrict__ group_up, const unsigned* __restrict__ n_groups)
{
    pdl_prologue();
    const unsigned ng = *n_groups;
    for (unsigned g = blockIdx.x * BLOCK + threadIdx.x; g < ng; g += gridDim.x * BLOCK) {
        const uint32_t slot = group_up[g].pw;
        if (slot == MCB200_NULL) continue; // the whole tree was one treelet
        uint32_t pw = __ldg(parent + slot);
        float box[6];
        {
            const float2* in = reinterpret_cast<const float2*>(group_up[g].box);
            const float2 a = in[0], b = in[1], c = in[2];
            box[0] = a.x; box[1] = a.y; box[2] = b.x; box[3] = b.y; box[4] = c.x; box[5] = c.y;
        }
        for (;;) {
            const uint32_t p = pw >> 2;
            bvh_node_t* nd = nodes + p;
            const bool is_left = !(pw & 1u);
            const uint32_t next_pw = (p == 0u) ? MCB200_NULL : __ldg(parent + p); // in flight while the fence drains
            store_box(is_left ? nd->lbox : nd->rbox, box);
            __threadfence();
            const unsigned arrived = atomicAdd(flags + p, 1u);
            if (arrived == 0) break; // sibling subtree not finished yet; its thread will continue from here
            float sib[6];
            load_box_cg(is_left ? nd->rbox : nd->lbox, sib);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                box[k] = fminf(box[k], sib[k]);
                box[3 + k] = fmaxf(box[3 + k], sib[3 + k]);
            }
            if (p == 0u) break; // root merged; the mesh AABB itself comes from K_aabb's reduction
            pw = next_pw;
        }
    }
}

} // namespace

int lbvh_reserve(mcb200_ctx* ctx, mcb200_mesh* m)
{
    if (m->nf == 0) {
        ctx->set_error("bvh_build: mesh has no faces", __FILE__, __LINE__);
        return MCB200_ERR_INVALID;
    }
    const uint32_t nf = m->nf;
    MCB_TRY(ctx->reserve(m->face_bbox, sizeof(double) * 6 * (size_t)nf));
    MCB_TRY(ctx->reserve(m->root, sizeof(unsigned long long) * 6 + sizeof(double) * 6));
    MCB_TRY(ctx->reserve(m->codes, sizeof(uint32_t) * (size_t)nf));
    MCB_TRY(ctx->reserve(m->sorted_codes, sizeof(uint32_t) * (size_t)nf));
    MCB_TRY(ctx->reserve(m->sorted_faces, sizeof(uint32_t) * (size_t)nf));
    MCB_TRY(ctx->reserve(m->nodes, sizeof(bvh_node_t) * (size_t)(nf > 1 ? nf - 1 : 1)));
    MCB_TRY(ctx->reserve(m->parent, sizeof(uint32_t) * (2 * (size_t)nf)));
    MCB_TRY(ctx->reserve(m->flags, sizeof(unsigned) * (size_t)nf));
    MCB_TRY(ctx->reserve(m->groups, sizeof(uint2) * (size_t)nf + sizeof(unsigned) * 4));
    MCB_TRY(ctx->reserve(m->group_up, sizeof(group_up_t) * (size_t)nf));
    MCB_TRY((rsort::reserve_scratch<uint32_t>(ctx, nf, 4, true, true)));
    return 0;
}

// Everything is enqueued on ctx->cur (the caller picks the lane); all allocations happen in lbvh_reserve.
int lbvh_build(mcb200_ctx* ctx, mcb200_mesh* m, double eps, bool query_only)
{
    MCB_TRY(lbvh_reserve(ctx, m));
    const uint32_t nf = m->nf;
    mcb200_ctx::sort_scratch_t& sc = ctx->sc();

    unsigned long long* root_ord = m->root.as<unsigned long long>();
    double* root_dec = reinterpret_cast<double*>(root_ord + 6);
    unsigned* n_groups = reinterpret_cast<unsigned*>(m->groups.as<uint2>() + nf);
    {
        // every small reset of the build in one launch: mesh AABB accumulators (min side all-ones, max side zero in the
        // ordered encoding), group counter, radix histograms and tile tickets.  The arrival flags (one word per face) are
        // cleared by k_face_bbox on its way through the faces.
        fill_list_t fl {};
        fl.add(root_ord, 6, 0xFFFFFFFFu);
        fl.add(root_ord + 3, 6, 0u);
        fl.add(n_groups, 4, 0u);
        fl.add(sc.hist.p, (size_t)rsort::MAX_PASSES * rsort::RADIX, 0u);
        fl.add(sc.tilectr.p, rsort::MAX_PASSES, 0u);
        MCB_LAUNCH(ctx, k_fill, 8, 256, 0, fl);
    }

    const unsigned max_grid = (unsigned)ctx->num_sms * 8u;
    const unsigned grid = div_up(nf, BLOCK) < max_grid ? div_up(nf, BLOCK) : max_grid;
    const double* prior = m->n_prior ? m->prior_bbox.as<double>() : nullptr;
    const uint32_t n_prior = m->n_prior < nf ? m->n_prior : nf;
    if (m->is_tri)
        MCB_LAUNCH(ctx, k_face_bbox<true>, grid, BLOCK, 0, m->d_xyz, m->frame, m->d_face_vtx, m->d_face_off, nf, eps,
            m->face_bbox.as<double>(), root_ord, m->flags.as<unsigned>(), prior, n_prior);
    else
        MCB_LAUNCH(ctx, k_face_bbox<false>, grid, BLOCK, 0, m->d_xyz, m->frame, m->d_face_vtx, m->d_face_off, nf, eps,
            m->face_bbox.as<double>(), root_ord, m->flags.as<unsigned>(), prior, n_prior);
    m->n_prior = 0; // consumed: the boxes are part of face_bbox now
    // (key, face) ascending by key (values implicit 0..nf-1), ping-pong scratch <-> mesh arrays; the histograms come out
    // of k_morton.  The leaves are ordered by the top `morton_sort_bits` bits of their code.  Nothing that leaves this stage depends on the
    // order (the pair SET is tree-independent and the treelets hold up to 32 leaves anyway), so the default sorts 24 bits in
    // three passes; codes that tie are told apart by their position, as equal codes always were.  With an odd number of
    // passes the keys start in the scratch buffer so that the last pass lands in the mesh's own arrays.
    const int sort_bits = ctx->morton_sort_bits >= 30 ? 32 : 24;
    const unsigned key_shift = sort_bits == 32 ? 0u : 6u;
    const rsort::pass_desc pd = rsort::make_passes(0, sort_bits);
    constexpr int SORT_TILE = rsort::THREADS * rsort::items_for<uint32_t>::value;
    const unsigned status_words = (unsigned)rsort::status_rows(((size_t)nf + SORT_TILE - 1) / SORT_TILE) * rsort::RADIX * (unsigned)pd.npasses;
    const bool odd = (pd.npasses & 1) != 0;
    uint32_t* keys_in = odd ? sc.keys_alt.as<uint32_t>() : m->sorted_codes.as<uint32_t>();
    uint32_t* keys_a = odd ? m->sorted_codes.as<uint32_t>() : sc.keys_alt.as<uint32_t>();
    uint32_t* keys_b = odd ? sc.keys_alt.as<uint32_t>() : m->sorted_codes.as<uint32_t>();
    uint32_t* vals_a = odd ? m->sorted_faces.as<uint32_t>() : sc.vals_alt.as<uint32_t>();
    uint32_t* vals_b = odd ? sc.vals_alt.as<uint32_t>() : m->sorted_faces.as<uint32_t>();
    MCB_LAUNCH(ctx, k_morton, grid, BLOCK, 0, m->face_bbox.as<double>(), nf, root_ord, root_dec, m->codes.as<uint32_t>(), keys_in,
        sc.hist.as<unsigned>(), sc.status.as<unsigned>(), status_words, key_shift, pd.npasses);
    uint32_t *kout = nullptr, *vout = nullptr;
    MCB_TRY((rsort::sort_passes<uint32_t, uint32_t, true>(ctx, keys_in, keys_a, keys_b, nullptr, vals_a, vals_b, nullptr, nf, pd, &kout,
        &vout)));
    if (kout != m->sorted_codes.as<uint32_t>() || vout != m->sorted_faces.as<uint32_t>()) {
        ctx->set_error("internal: the Morton sort did not end in the mesh's arrays", __FILE__, __LINE__);
        return MCB200_ERR_INTERNAL;
    }
    if (nf <= 1) query_only = false; // the one-leaf pseudo tree is free
    if (query_only)
        MCB_LAUNCH(ctx, k_tree<false>, div_up(nf, BLOCK), BLOCK, 0, m->sorted_codes.as<uint32_t>(), m->face_bbox.as<double>(),
            m->sorted_faces.as<uint32_t>(), nf, m->nodes.as<bvh_node_t>(), m->parent.as<uint32_t>(), m->groups.as<uint2>(),
            m->group_up.as<group_up_t>(), n_groups);
    else
        MCB_LAUNCH(ctx, k_tree<true>, div_up(nf, BLOCK), BLOCK, 0, m->sorted_codes.as<uint32_t>(), m->face_bbox.as<double>(),
            m->sorted_faces.as<uint32_t>(), nf, m->nodes.as<bvh_node_t>(), m->parent.as<uint32_t>(), m->groups.as<uint2>(),
            m->group_up.as<group_up_t>(), n_groups);
    m->has_nodes = !query_only;
    if (nf > 1 && !query_only) {
        // enough threads for every group root to climb concurrently (about nf/16 of them; nf/4 is a safe bound for the grid,
        // the kernel strides over the device-side count anyway)
        const unsigned want = div_up((size_t)nf / 4u + 1u, BLOCK);
        const unsigned gc = want < max_grid ? want : max_grid;
        MCB_LAUNCH(ctx, k_refit_climb, gc, BLOCK, 0, m->nodes.as<bvh_node_t>(), m->parent.as<uint32_t>(), m->flags.as<unsigned>(),
            m->group_up.as<group_up_t>(), n_groups);
    }
    m->built = true;
    m->groups_valid = true;
    m->eps = eps;
    return 0;
}
