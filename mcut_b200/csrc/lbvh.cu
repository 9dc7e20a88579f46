// mcut_b200/csrc/lbvh.cu — (1) LBVH construction on the device.
//
// Replaces build_oibvh() (include/mcut/internal/bvh.h:117-125, source/bvh.cpp:219-636):
//   k_face_bbox    face AABBs (+eps enlargement) and the mesh AABB               bvh.cpp:242-368, math.h:866-928
//   k_morton       30-bit Morton codes, the reference's float formula;           bvh.cpp:196-217, :373-433
//                  also the digit histograms of the sort's passes
//   radix passes   one-sweep radix sort of (code, face)  (radix_sort.cuh)        bvh.cpp:437-442 (std::sort there)
//   k_leaves       everything above the sorted leaves in ONE kernel:             the implicit OIBVH layout bvh.cpp:444-493 and the
//                  * the leaf boxes in sorted order (single precision, outwards)  level-by-level refit bvh.cpp:498-635
//                  * an implicit 32-wide tree over them (level l node i = nodes
//                    [32i, 32i+32) of level l-1), every level's boxes, the upper
//                    levels by the last block to finish (atomic ticket)
//                  * the QUERY GROUPS: the maximal subtrees of <= 32 leaves of the
//                    binary radix tree over the sorted codes (Karras 2012), found
//                    from the tree's delta function alone, without building it
// Only face AABBs, the mesh AABB and the leaf-pair SET escape this stage, and every internal box is a (conservative) union
// of its leaves' boxes, so the tree shape is free: any hierarchy over the same leaves yields the same pairs.
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
__global__ void __launch_bounds__(BLOCK) k_face_bbox(const void* __restrict__ xyz, const frame_t* __restrict__ frp,
    const uint32_t* __restrict__ face_vtx, const uint32_t* __restrict__ face_off, uint32_t nf,
    double* __restrict__ face_bbox, unsigned long long* __restrict__ root_ordered, const double* __restrict__ prior, uint32_t n_prior)
{
    pdl_prologue();
    __shared__ frame_t s_fr;
    load_frame_shared(&s_fr, frp);
    __syncthreads();
    const frame_t& fr = s_fr;
    const double eps = s_fr.eps;
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

// ---- K_vertex_bbox + K_face_codes: the first two kernels in one ---------------------------------------------------------
// The Morton code of a face needs the mesh AABB, which is the union of all face boxes — a grid-wide dependency that used to
// cost a second kernel re-reading all 48-byte boxes.  But the union of the face boxes is the box of the VERTICES, enlarged
// by eps (every vertex of a mesh the reference accepts belongs to a face — check_input_mesh, preproc.cpp:505-578, rejects
// stray vertices — and x -> x -/+ eps is monotone, so min/max and the enlargement commute bit for bit).  A pass over the
// vertices (a quarter of the bytes) yields it up front; one kernel then gathers each face once and writes box, code and sort
// key together.  The mesh AABB that is RETURNED is still reduced from the face boxes themselves, so it is exact whatever
// the input.  Meshes that carry prior boxes (the in/out vector after a repartition) take the two-kernel path.
__global__ void __launch_bounds__(BLOCK) k_vertex_bbox(const void* __restrict__ xyz, const frame_t* __restrict__ frp, uint32_t nv,
    unsigned long long* __restrict__ vroot_ordered)
{
    pdl_prologue();
    __shared__ frame_t s_fr;
    load_frame_shared(&s_fr, frp);
    __syncthreads();
    double bmin[3] = { DBL_MAX, DBL_MAX, DBL_MAX }, bmax[3] = { -DBL_MAX, -DBL_MAX, -DBL_MAX };
    for (uint32_t v = blockIdx.x * BLOCK + threadIdx.x; v < nv; v += gridDim.x * BLOCK) {
        double p[3];
        load_vertex(xyz, s_fr, v, p);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            bmax[j] = ref_max(bmax[j], p[j]);
            bmin[j] = ref_min(bmin[j], p[j]);
        }
    }
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
            if (v != DBL_MAX) atomicMin(vroot_ordered + j, dbl_to_ordered(v));
        } else {
            double v = s_max[0][j];
            for (int i = 1; i < BLOCK / 32; ++i) v = fmax(v, s_max[i][j]);
            if (v != -DBL_MAX) atomicMax(vroot_ordered + 3 + j, dbl_to_ordered(v));
        }
    }
}

template <bool TRI>
__global__ void __launch_bounds__(BLOCK) k_face_codes(const void* __restrict__ xyz, const frame_t* __restrict__ frp,
    const uint32_t* __restrict__ face_vtx, const uint32_t* __restrict__ face_off, uint32_t nf, double* __restrict__ face_bbox,
    const unsigned long long* __restrict__ vroot_ordered, unsigned long long* __restrict__ root_ordered, uint32_t* __restrict__ codes,
    uint32_t* __restrict__ sort_keys, unsigned* __restrict__ hist /* [4][256] */, unsigned* __restrict__ status, unsigned status_words,
    unsigned key_shift, int npasses)
{
    pdl_prologue();
    __shared__ frame_t s_fr;
    __shared__ unsigned s_hist[4 * 256];
    load_frame_shared(&s_fr, frp);
    for (int i = threadIdx.x; i < 4 * 256; i += BLOCK) s_hist[i] = 0;
    for (unsigned i = blockIdx.x * BLOCK + threadIdx.x; i < status_words; i += gridDim.x * BLOCK) status[i] = 0u;
    __syncthreads();
    const frame_t& fr = s_fr;
    const double eps = s_fr.eps;
    // the mesh AABB the codes are normalised with: vertex box enlarged like every face box is (bvh.cpp:262-272, :330-368)
    double rmin[3], dims[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double lo = ordered_to_dbl(vroot_ordered[j]), hi = ordered_to_dbl(vroot_ordered[3 + j]);
        if (eps > 0.0) {
            hi = __dadd_rn(hi, eps);
            lo = __dsub_rn(lo, eps);
        }
        rmin[j] = lo;
        dims[j] = __dsub_rn(hi, lo);
    }
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
        if (eps > 0.0) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                mx[j] = __dadd_rn(mx[j], eps);
                mn[j] = __dsub_rn(mn[j], eps);
            }
        }
        double* out = face_bbox + 6 * (size_t)f;
        reinterpret_cast<double2*>(out)[0] = make_double2(mn[0], mn[1]);
        reinterpret_cast<double2*>(out)[1] = make_double2(mn[2], mx[0]);
        reinterpret_cast<double2*>(out)[2] = make_double2(mx[1], mx[2]);
        float nrm[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            bmax[j] = ref_max(bmax[j], mx[j]);
            bmin[j] = ref_min(bmin[j], mn[j]);
            const double centre = __dadd_rn(mn[j], mx[j]) / 2; // bvh.cpp:275
            const double off = __dsub_rn(centre, rmin[j]); // bvh.cpp:382
            nrm[j] = (float)(off / dims[j]); // bvh.cpp:399-402
        }
        const uint32_t code = morton3D(nrm[0], nrm[1], nrm[2]);
        codes[f] = code;
        const uint32_t key = code >> key_shift;
        sort_keys[f] = key;
#pragma unroll
        for (int p = 0; p < 4; ++p)
            if (p < npasses) atomicAdd(&s_hist[p * 256 + ((key >> (8 * p)) & 255u)], 1u);
    }
    // the returned mesh AABB: reduced from the face boxes themselves
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
    for (int i = threadIdx.x; i < 4 * 256; i += BLOCK) {
        const unsigned c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// 24 bytes, 8-byte aligned
__device__ __forceinline__ void store_box(float* dst, const float* b) // 24 bytes, 8-byte aligned
{
    reinterpret_cast<float2*>(dst)[0] = make_float2(b[0], b[1]);
    reinterpret_cast<float2*>(dst)[1] = make_float2(b[2], b[3]);
    reinterpret_cast<float2*>(dst)[2] = make_float2(b[4], b[5]);
}

// Slots in the group list for a whole block with ONE global atomic (tens of thousands of same-address atomics from
// individual warps serialise in L2 and were the most expensive part of this kernel).  Every thread of the block must
// call this; `want` is small.  Returns the first slot of the calling thread.
__device__ __forceinline__ unsigned alloc_groups_block(unsigned* n_groups, unsigned want, unsigned* s_warp /*[blockDim.x/32]*/,
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
        for (unsigned i = 0; i < blockDim.x / 32u; ++i) {
            const unsigned c = s_warp[i];
            s_warp[i] = tot;
            tot += c;
        }
        *s_base = tot ? atomicAdd(n_groups, tot) : 0u;
    }
    __syncthreads();
    return *s_base + s_warp[w] + inc - want;
}

// ---- K_leaves -----------------------------------------------------------------------------------------------------
// One block = 1024 consecutive sorted leaves = 32 level-1 nodes = ONE level-2 node.
//
// Wide tree.  Leaf j's exact box is gathered through the sorted order, rounded outwards to single precision and stored in the
// level-0 array; a warp-wide min/max of 32 consecutive leaves is a level-1 box, the block's 32 level-1 boxes give its
// level-2 box.  Levels 3.. (F/32768 nodes and fewer) are finished by whichever block takes the last ticket.  Fixed runs of
// 32 leaves may straddle a coarse cell boundary of the Morton curve and get a fat box; on the TREE side of a traversal that
// costs one extra step now and then (the children are tight again) and nothing else.
//
// Query groups.  On the QUERY side a fat box would be ruinous (the group's box steers its whole walk), so the leaves are
// grouped the way a binary radix tree over the sorted codes (Karras 2012) would group them: a group is a maximal subtree
// with at most 32 leaves.  The tree is never built.  With delta(j) = the length of the common prefix of codes j-1 and j
// (ties broken by the index, as in the paper), the radix-tree node that splits between leaves j-1 and j covers the maximal
// run of leaves around that boundary whose inner boundaries all have a LARGER delta; leaves j-1 and j lie in different
// groups exactly when that run holds more than 32 leaves.  So each boundary counts larger deltas to its left and right
// (it stops at 32; most boundaries stop after one or two steps) and is a group boundary or not — independent, no atomics.
constexpr int LB_THREADS = 256;
constexpr int LB_LEAVES = 1024;
constexpr int LB_HL = 34, LB_HR = 66; // halo of the code window: boundaries up to 32 + 31 past the block are classified
constexpr int LB_WIN = LB_LEAVES + LB_HL + LB_HR;
constexpr int LB_BOXES = LB_LEAVES + 32; // leaf boxes kept in shared memory: a group may run 31 leaves into the next block
constexpr int LB_ITEMS = LB_LEAVES / LB_THREADS; // 4 leaf blocks (chunks of 32 leaves) per warp
constexpr int LB_LOG = 5; // range-minimum tables over 1, 2, 4, 8, 16 boundaries

__device__ __forceinline__ void warp_union(float* b)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            b[k] = fminf(b[k], __shfl_xor_sync(0xffffffffu, b[k], o));
            b[3 + k] = fmaxf(b[3 + k], __shfl_xor_sync(0xffffffffu, b[3 + k], o));
        }
    }
}

__device__ __forceinline__ void empty_box(float* b)
{
    b[0] = b[1] = b[2] = FLT_MAX;
    b[3] = b[4] = b[5] = -FLT_MAX;
}

// Segmented SUFFIX union of the 32 boxes of one chunk (lane = leaf; a segment starts at every lane whose bit is set in
// `heads`, the lanes below the first head form a segment of their own): suf = union of the boxes from the lane to the end of
// its segment.  At a head lane that is its group's box (as far as this chunk goes); at lane 0, when it is no head, it is
// the part of the previous chunk's last group that lies in this chunk.  (redux.sync per segment mask was tried: fewer
// instructions, but slower than these five shuffle rounds.)
__device__ __forceinline__ void segmented_suffix_union(const float* box, unsigned heads, float* suf)
{
    const unsigned lane = lane_id();
#pragma unroll
    for (int c = 0; c < 6; ++c) suf[c] = box[c];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        // lanes (lane, lane + d] hold no segment start: the value d lanes above belongs to the same segment
        const bool down = lane + d < 32u && ((heads >> (lane + 1u)) & ((1u << d) - 1u)) == 0u;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float sd = __shfl_down_sync(0xffffffffu, suf[c], d), sD = __shfl_down_sync(0xffffffffu, suf[3 + c], d);
            if (down) {
                suf[c] = fminf(suf[c], sd);
                suf[3 + c] = fmaxf(suf[3 + c], sD);
            }
        }
    }
}

__global__ void __launch_bounds__(LB_THREADS, 3) k_leaves(const uint32_t* __restrict__ codes, const double* __restrict__ face_bbox,
    const uint32_t* __restrict__ sorted_faces, uint32_t nf, float* __restrict__ wide, double* __restrict__ sorted_bbox, wide_levels_t lv,
    uint2* __restrict__ groups, group_box_t* __restrict__ group_box, unsigned* __restrict__ n_groups, unsigned* __restrict__ done_ticket,
    const unsigned long long* __restrict__ root_ordered, double* __restrict__ root_decoded)
{
    pdl_prologue();
    if (blockIdx.x == 0 && threadIdx.x < 6) root_decoded[threadIdx.x] = ordered_to_dbl(root_ordered[threadIdx.x]); // the mesh AABB callers read
    __shared__ uint32_t s_code[LB_WIN];
    // s_lam[0]: delta + 1 of the boundary before window leaf k (0 at the ends of the array); s_lam[t]: minimum over 2^t boundaries
    __shared__ uint8_t s_lam[LB_LOG][LB_WIN];
    __shared__ float s_pre[6][LB_BOXES / 32]; // per chunk of 32 leaves: union of the leaves BEFORE its first group start (the tail of a group of the previous chunk)
    __shared__ uint32_t s_flag[LB_BOXES / 32 + 2];
    __shared__ float s_l1[6][32];
    __shared__ unsigned s_warp[LB_THREADS / 32], s_base, s_ticket;
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * LB_LEAVES; // first leaf of the block
    const long long wbase = base - LB_HL; // leaf index of window slot 0

    // ---- loads: the code window and the face ids (coalesced), then the exact boxes through the face ids ----
    {
        constexpr int R = (LB_WIN + LB_THREADS - 1) / LB_THREADS;
        uint32_t creg[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int k = r * LB_THREADS + (int)threadIdx.x;
            const long long j = wbase + k;
            creg[r] = (k < LB_WIN && j >= 0 && j < (long long)nf) ? __ldg(codes + j) : 0u;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int k = r * LB_THREADS + (int)threadIdx.x;
            if (k < LB_WIN) s_code[k] = creg[r];
        }
    }
    // chunks q = 4 w + i (i < 4) belong to warp w (their face ids are requested now: one hop less when the boxes are gathered); warp 0 also fetches the 32 leaves after the block (i == 4, q = 32)
    constexpr int NI = LB_ITEMS + 1;
    uint32_t face[NI];
    bool have[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const unsigned q = (i < LB_ITEMS) ? w * LB_ITEMS + i : 32u;
        const long long j = base + 32ll * q + lane;
        have[i] = (i < LB_ITEMS || w == 0) && j < (long long)nf;
        face[i] = have[i] ? __ldg(sorted_faces + j) : 0u;
    }
    __syncthreads();
    // ---- delta of every boundary of the window, then its range minima ----
    for (int k = threadIdx.x; k < LB_WIN; k += LB_THREADS) {
        const long long j = wbase + k;
        unsigned lam = 0;
        if (k >= 1 && j >= 1 && j < (long long)nf) {
            const uint32_t ca = s_code[k - 1], cb = s_code[k];
            lam = 1u + (unsigned)((ca == cb) ? 32 + __clz((unsigned)(j - 1) ^ (unsigned)j) : __clz(ca ^ cb));
        }
        s_lam[0][k] = (uint8_t)lam;
    }
#pragma unroll
    for (int t = 1; t < LB_LOG; ++t) {
        __syncthreads();
        const int half = 1 << (t - 1);
        for (int k = threadIdx.x; k < LB_WIN; k += LB_THREADS) {
            const unsigned x = s_lam[t - 1][k], y = (k + half < LB_WIN) ? s_lam[t - 1][k + half] : 0u;
            s_lam[t][k] = (uint8_t)(x < y ? x : y);
        }
    }
    __syncthreads();
    // ---- group boundaries among the boundaries [base, base + 1056): how many consecutive boundaries on either side have
    //      a larger delta (binary lifting over the range minima, at most 31 each way) ----
    for (int r = 0; r < (LB_BOXES + LB_THREADS - 1) / LB_THREADS; ++r) {
        const int b = r * LB_THREADS + (int)threadIdx.x; // boundary before leaf base + b
        bool flag = false;
        if (b < LB_BOXES) {
            const long long j = base + b;
            if (j == 0 || j >= (long long)nf) {
                flag = true;
            } else {
                const int k = b + LB_HL;
                const unsigned lam = s_lam[0][k];
                int lo = k, hi = k + 1; // boundaries (lo, k) and [k + 1, hi) are known to be larger
#pragma unroll
                for (int t = LB_LOG - 1; t >= 0; --t) {
                    const int span = 1 << t;
                    if (s_lam[t][lo - span] > lam) lo -= span;
                    if (s_lam[t][hi] > lam) hi += span;
                }
                flag = (k - lo) + (hi - k - 1) + 2 > 32;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, flag);
        const unsigned word = (unsigned)(r * LB_THREADS + (int)threadIdx.x) >> 5;
        if (lane == 0 && word < LB_BOXES / 32 + 2) s_flag[word] = m;
    }
    __syncthreads();
    double2 bx[NI][3];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        if (have[i]) {
            const double2* in = reinterpret_cast<const double2*>(face_bbox + 6 * (size_t)face[i]);
            bx[i][0] = __ldg(in);
            bx[i][1] = __ldg(in + 1);
            bx[i][2] = __ldg(in + 2);
        }
    }
    // ---- leaf boxes: exact copy in sorted order, single precision outwards -> level 0 ([6][32] blocks), level-1 boxes,
    //      segmented unions for the groups ----
    const uint32_t nb0 = (nf + 31u) / 32u;
    float suf[LB_ITEMS][6];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        if (i == LB_ITEMS && w != 0) break;
        const unsigned q = (i < LB_ITEMS) ? w * LB_ITEMS + i : 32u;
        float fb[6];
        if (have[i]) {
            const double d[6] = { bx[i][0].x, bx[i][0].y, bx[i][1].x, bx[i][1].y, bx[i][2].x, bx[i][2].y };
            box_to_float(d, fb);
            if (i < LB_ITEMS) {
                double2* out = reinterpret_cast<double2*>(sorted_bbox + 6 * (size_t)(base + 32ll * q + lane));
                out[0] = bx[i][0];
                out[1] = bx[i][1];
                out[2] = bx[i][2];
            }
        } else {
            empty_box(fb);
        }
        // union over the lane's own segment: for a head lane that is its group (as far as this chunk goes), for lane 0 — when
        // it is no head — the part of the previous chunk's last group that lies in this chunk
        float sf[6];
        segmented_suffix_union(fb, s_flag[q], sf);
        if (lane == 0)
#pragma unroll
            for (int c = 0; c < 6; ++c) s_pre[c][q] = sf[c];
        if (i < LB_ITEMS) {
#pragma unroll
            for (int c = 0; c < 6; ++c) suf[i][c] = sf[c];
            const unsigned long long blk = (unsigned long long)blockIdx.x * 32u + q;
            if (blk < nb0) {
                float* dst = wide + ((size_t)lv.off[0] + blk) * MCB_WBLOCK_FLOATS + lane;
#pragma unroll
                for (int c = 0; c < 6; ++c) dst[c * 32] = fb[c];
            }
            warp_union(fb);
            if (lane == 0)
#pragma unroll
                for (int c = 0; c < 6; ++c) s_l1[c][q] = fb[c];
        }
    }
    __syncthreads();
    // ---- groups that start in this block: a group is a segment, possibly continued in the next chunk ----
    {
        unsigned want = 0;
        unsigned cnt_of[LB_ITEMS];
#pragma unroll
        for (int i = 0; i < LB_ITEMS; ++i) {
            const unsigned q = w * LB_ITEMS + i, p = q * 32u + lane; // leaf base + p
            cnt_of[i] = 0;
            if (base + p < (long long)nf && ((s_flag[q] >> lane) & 1u)) {
                // the next group boundary is at most 32 leaves away
                const unsigned rest = lane < 31u ? (s_flag[q] >> (lane + 1u)) : 0u;
                if (rest) {
                    cnt_of[i] = (unsigned)__ffs((int)rest);
                } else {
                    const unsigned t = (unsigned)__ffs((int)s_flag[q + 1u]) - 1u; // leaves of the next chunk that still belong
                    cnt_of[i] = 32u - lane + t;
                    if (t > 0u) { // lane 0 of the next chunk is no head: its segment is the rest of this group
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            suf[i][c] = fminf(suf[i][c], s_pre[c][q + 1u]);
                            suf[i][3 + c] = fmaxf(suf[i][3 + c], s_pre[3 + c][q + 1u]);
                        }
                    }
                }
                ++want;
            }
        }
        unsigned g = alloc_groups_block(n_groups, want, s_warp, &s_base);
#pragma unroll
        for (int i = 0; i < LB_ITEMS; ++i) {
            if (!cnt_of[i]) continue;
            const unsigned p = (w * LB_ITEMS + i) * 32u + lane;
            groups[g] = make_uint2((uint32_t)(base + p), cnt_of[i]);
            store_box(group_box[g].box, suf[i]);
            ++g;
        }
    }
    // ---- level 1 block of this leaf range, its level-2 node; the levels above by the last block ----
    if (w == 0) {
        float b1[6];
        const unsigned long long node1 = (unsigned long long)blockIdx.x * 32u + lane;
        if (node1 < lv.n[1]) {
#pragma unroll
            for (int c = 0; c < 6; ++c) b1[c] = s_l1[c][lane];
        } else {
            empty_box(b1);
        }
        float* dst = wide + ((size_t)lv.off[1] + blockIdx.x) * MCB_WBLOCK_FLOATS + lane;
#pragma unroll
        for (int c = 0; c < 6; ++c) dst[c * 32] = b1[c];
        if (lv.top >= 2) {
            warp_union(b1);
            if (lane == 0) {
                float* d2 = wide + ((size_t)lv.off[2] + (blockIdx.x >> 5)) * MCB_WBLOCK_FLOATS + (blockIdx.x & 31u);
#pragma unroll
                for (int c = 0; c < 6; ++c) d2[c * 32] = b1[c];
                __threadfence();
                s_ticket = atomicAdd(done_ticket, 1u);
            }
        }
    }
    if (lv.top < 2) return;
    __syncthreads();
    if (s_ticket != gridDim.x - 1u) return;
    __threadfence();
    for (int l = 2; l <= lv.top; ++l) {
        // pad level l up to a whole block with empty boxes
        const uint32_t nl = lv.n[l], padded = ((nl + 31u) / 32u) * 32u;
        for (uint32_t i = nl + threadIdx.x; i < padded; i += LB_THREADS) {
            float* d = wide + ((size_t)lv.off[l] + (i >> 5)) * MCB_WBLOCK_FLOATS + (i & 31u);
            d[0] = d[32] = d[64] = FLT_MAX;
            d[96] = d[128] = d[160] = -FLT_MAX;
        }
        if (l < lv.top) {
            const uint32_t nup = lv.n[l + 1];
            for (uint32_t i = w; i < nup; i += LB_THREADS / 32) {
                const uint32_t child = 32u * i + lane;
                float b[6];
                if (child < nl) {
                    const float* src = wide + ((size_t)lv.off[l] + i) * MCB_WBLOCK_FLOATS + lane;
#pragma unroll
                    for (int c = 0; c < 6; ++c) b[c] = __ldcg(src + c * 32);
                } else {
                    empty_box(b);
                }
                warp_union(b);
                if (lane == 0) {
                    float* d = wide + ((size_t)lv.off[l + 1] + (i >> 5)) * MCB_WBLOCK_FLOATS + (i & 31u);
#pragma unroll
                    for (int c = 0; c < 6; ++c) d[c * 32] = b[c];
                }
            }
        }
        __syncthreads();
    }
}

} // namespace

// level table of the implicit 32-wide tree over nf leaves
static void make_levels(uint32_t nf, wide_levels_t& lv)
{
    std::memset(&lv, 0, sizeof(lv));
    uint32_t acc = 0;
    lv.n[0] = nf;
    for (int l = 0; l < MCB_MAX_LEVELS; ++l) {
        const uint32_t nb = (lv.n[l] + 31u) / 32u;
        lv.off[l] = acc;
        acc += nb;
        if (l >= 1 && lv.n[l] <= 32u) {
            lv.top = l;
            break;
        }
        if (l + 1 < MCB_MAX_LEVELS) lv.n[l + 1] = nb;
    }
}

static size_t wide_blocks(const wide_levels_t& lv)
{
    return (size_t)lv.off[lv.top] + 1u;
}

int lbvh_reserve(mcb200_ctx* ctx, mcb200_mesh* m)
{
    if (m->nf == 0) {
        ctx->set_error("bvh_build: mesh has no faces", __FILE__, __LINE__);
        return MCB200_ERR_INVALID;
    }
    const uint32_t nf = m->nf;
    if (!m->lv) m->lv = new wide_levels_t();
    make_levels(nf, *m->lv);
    MCB_TRY(ctx->reserve(m->face_bbox, sizeof(double) * 6 * (size_t)nf));
    MCB_TRY(ctx->reserve(m->root, sizeof(unsigned long long) * 6 + sizeof(double) * 6 + sizeof(unsigned long long) * 6));
    MCB_TRY(ctx->reserve(m->codes, sizeof(uint32_t) * (size_t)nf));
    MCB_TRY(ctx->reserve(m->sorted_codes, sizeof(uint32_t) * (size_t)nf));
    MCB_TRY(ctx->reserve(m->sorted_faces, sizeof(uint32_t) * (size_t)nf));
    MCB_TRY(ctx->reserve(m->wide, sizeof(float) * MCB_WBLOCK_FLOATS * wide_blocks(*m->lv)));
    MCB_TRY(ctx->reserve(m->sorted_bbox, sizeof(double) * 6 * (size_t)nf));
    MCB_TRY(ctx->reserve(m->flags, sizeof(unsigned) * 4));
    MCB_TRY(ctx->reserve(m->d_frames, sizeof(frame_t) * 2));
    MCB_TRY(ctx->reserve(m->groups, sizeof(uint2) * (size_t)nf + sizeof(unsigned) * 4));
    MCB_TRY(ctx->reserve(m->group_box, sizeof(group_box_t) * (size_t)nf));
    MCB_TRY((rsort::reserve_scratch<uint32_t>(ctx, nf, 4, true, true)));
    m->lv->boxes = m->wide.as<float>();
    return 0;
}

// Everything is enqueued on ctx->cur (the caller picks the lane); all allocations happen in lbvh_reserve.
int mesh_sync_frames(mcb200_ctx* ctx, mcb200_mesh* a, double eps_a, mcb200_mesh* b, double eps_b)
{
    frame_pack_t pk;
    std::memset(&pk, 0, sizeof(pk));
    mcb200_mesh* ms[2] = { a, b };
    const double eps[2] = { eps_a, eps_b };
    for (int k = 0; k < 2; ++k) {
        mcb200_mesh* m = ms[k];
        if (!m) continue;
        MCB_TRY(ctx->reserve(m->d_frames, sizeof(frame_t) * 2));
        frame_t f[2];
        f[0] = m->frame;
        f[0].has_pert = 0;
        f[0].pert[0] = f[0].pert[1] = f[0].pert[2] = 0.0;
        f[0].eps = eps[k];
        f[1] = m->frame;
        f[1].eps = 0.0;
        if (m->dev_frames_valid && m->dev_frames_ptr == m->d_frames.p && std::memcmp(f, m->dev_frames, sizeof(f)) == 0) continue;
        std::memcpy(m->dev_frames, f, sizeof(f));
        m->dev_frames_valid = true;
        m->dev_frames_ptr = m->d_frames.p;
        pk.dst[pk.n] = m->d_frames.as<frame_t>();
        pk.f[pk.n][0] = f[0];
        pk.f[pk.n][1] = f[1];
        ++pk.n;
    }
    if (pk.n) MCB_LAUNCH(ctx, k_set_frames, 1, 128, 0, pk);
    return 0;
}

// The build reads the mesh's frame and eps from its device slot: the caller has called mesh_sync_frames(ctx, m, eps).
int lbvh_build(mcb200_ctx* ctx, mcb200_mesh* m, double eps)
{
    MCB_TRY(lbvh_reserve(ctx, m));
    const uint32_t nf = m->nf;
    mcb200_ctx::sort_scratch_t& sc = ctx->sc();

    unsigned long long* root_ord = m->root.as<unsigned long long>();
    double* root_dec = reinterpret_cast<double*>(root_ord + 6);
    unsigned* n_groups = reinterpret_cast<unsigned*>(m->groups.as<uint2>() + nf);
    {
        // every small reset of the build in one launch: mesh AABB accumulators (min side all-ones, max side zero in the
        // ordered encoding), group counter, the last-block ticket, radix histograms and tile tickets
        fill_list_t fl {};
        fl.add(root_ord, 6, 0xFFFFFFFFu);
        fl.add(root_ord + 3, 6, 0u);
        fl.add(root_ord + 12, 6, 0xFFFFFFFFu); // the same for the vertex box
        fl.add(root_ord + 15, 6, 0u);
        fl.add(n_groups, 4, 0u);
        fl.add(m->flags.p, 4, 0u);
        fl.add(sc.hist.p, (size_t)rsort::MAX_PASSES * rsort::RADIX, 0u);
        fl.add(sc.tilectr.p, rsort::MAX_PASSES, 0u);
        MCB_LAUNCH(ctx, k_fill, 8, 256, 0, fl);
    }

    const unsigned max_grid = (unsigned)ctx->num_sms * 8u;
    const unsigned grid = div_up(nf, BLOCK) < max_grid ? div_up(nf, BLOCK) : max_grid;
    const double* prior = m->n_prior ? m->prior_bbox.as<double>() : nullptr;
    const uint32_t n_prior = m->n_prior < nf ? m->n_prior : nf;
    // (key, face) ascending by key (values implicit 0..nf-1), ping-pong scratch <-> mesh arrays; the histograms come out
    // of the kernel that makes the keys.  The leaves are ordered by the top `morton_sort_bits` bits of their code.  Nothing
    // that leaves this stage depends on the order (the pair SET is tree-independent and the groups hold up to 32 leaves
    // anyway), so the default sorts 16 bits in two passes; codes that tie are told apart by their position, as equal codes
    // always were.  Measured against 24 bits / three passes (one B200, ms per step): C2 0.301 vs 0.320, C3 1.253 vs 1.409 (a flat
    // terrain wastes the z bits of the interleaved code: 8.5 M node tests instead of 14.4 M), C5 3.428 vs 3.383.  With an odd number of passes the keys start in the scratch buffer so that the last pass lands in the
    // mesh's own arrays.
    const int sort_bits = ctx->morton_sort_bits >= 30 ? 32 : (ctx->morton_sort_bits <= 16 ? 16 : 24);
    const unsigned key_shift = sort_bits == 32 ? 0u : (sort_bits == 16 ? 14u : 6u);
    const rsort::pass_desc pd = rsort::make_passes(0, sort_bits);
    const int SORT_TILE = rsort::THREADS * rsort::items_rt<uint32_t>(nf);
    const unsigned status_words = (unsigned)rsort::status_rows(((size_t)nf + SORT_TILE - 1) / SORT_TILE) * rsort::RADIX * (unsigned)pd.npasses;
    const bool odd = (pd.npasses & 1) != 0;
    uint32_t* keys_in = odd ? sc.keys_alt.as<uint32_t>() : m->sorted_codes.as<uint32_t>();
    uint32_t* keys_a = odd ? m->sorted_codes.as<uint32_t>() : sc.keys_alt.as<uint32_t>();
    uint32_t* keys_b = odd ? sc.keys_alt.as<uint32_t>() : m->sorted_codes.as<uint32_t>();
    uint32_t* vals_a = odd ? m->sorted_faces.as<uint32_t>() : sc.vals_alt.as<uint32_t>();
    uint32_t* vals_b = odd ? sc.vals_alt.as<uint32_t>() : m->sorted_faces.as<uint32_t>();
    if (n_prior == 0 && !ctx->two_kernel_boxes) {
        unsigned long long* vroot = root_ord + 12; // vertex-box accumulators (cleared by the fill above)
        const unsigned vgrid = div_up(m->nv, BLOCK) < max_grid ? div_up(m->nv, BLOCK) : max_grid;
        MCB_LAUNCH(ctx, k_vertex_bbox, vgrid, BLOCK, 0, m->d_xyz, m->d_frames.as<frame_t>(), m->nv, vroot);
        if (m->is_tri)
            MCB_LAUNCH(ctx, k_face_codes<true>, grid, BLOCK, 0, m->d_xyz, m->d_frames.as<frame_t>(), m->d_face_vtx, m->d_face_off, nf,
                m->face_bbox.as<double>(), vroot, root_ord, m->codes.as<uint32_t>(), keys_in, sc.hist.as<unsigned>(), sc.status.as<unsigned>(),
                status_words, key_shift, pd.npasses);
        else
            MCB_LAUNCH(ctx, k_face_codes<false>, grid, BLOCK, 0, m->d_xyz, m->d_frames.as<frame_t>(), m->d_face_vtx, m->d_face_off, nf,
                m->face_bbox.as<double>(), vroot, root_ord, m->codes.as<uint32_t>(), keys_in, sc.hist.as<unsigned>(), sc.status.as<unsigned>(),
                status_words, key_shift, pd.npasses);
    } else {
        if (m->is_tri)
            MCB_LAUNCH(ctx, k_face_bbox<true>, grid, BLOCK, 0, m->d_xyz, m->d_frames.as<frame_t>(), m->d_face_vtx, m->d_face_off, nf,
                m->face_bbox.as<double>(), root_ord, prior, n_prior);
        else
            MCB_LAUNCH(ctx, k_face_bbox<false>, grid, BLOCK, 0, m->d_xyz, m->d_frames.as<frame_t>(), m->d_face_vtx, m->d_face_off, nf,
                m->face_bbox.as<double>(), root_ord, prior, n_prior);
        MCB_LAUNCH(ctx, k_morton, grid, BLOCK, 0, m->face_bbox.as<double>(), nf, root_ord, root_dec, m->codes.as<uint32_t>(), keys_in,
            sc.hist.as<unsigned>(), sc.status.as<unsigned>(), status_words, key_shift, pd.npasses);
    }
    m->n_prior = 0; // consumed: the boxes are part of face_bbox now
    uint32_t *kout = nullptr, *vout = nullptr;
    MCB_TRY((rsort::sort_passes<uint32_t, uint32_t, true>(ctx, keys_in, keys_a, keys_b, nullptr, vals_a, vals_b, nullptr, nf, pd, &kout,
        &vout)));
    if (kout != m->sorted_codes.as<uint32_t>() || vout != m->sorted_faces.as<uint32_t>()) {
        ctx->set_error("internal: the Morton sort did not end in the mesh's arrays", __FILE__, __LINE__);
        return MCB200_ERR_INTERNAL;
    }
    MCB_LAUNCH(ctx, k_leaves, div_up(nf, LB_LEAVES), LB_THREADS, 0, m->sorted_codes.as<uint32_t>(), m->face_bbox.as<double>(),
        m->sorted_faces.as<uint32_t>(), nf, m->wide.as<float>(), m->sorted_bbox.as<double>(), *m->lv, m->groups.as<uint2>(), m->group_box.as<group_box_t>(), n_groups,
        m->flags.as<unsigned>(), root_ord, root_dec);
    m->built = true;
    m->groups_valid = true;
    m->eps = eps;
    return 0;
}
