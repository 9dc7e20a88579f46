// mcut_b200/csrc/traverse.cu — (2) BVH x BVH overlap traversal.
//
// Replaces intersectOIBVHs() (include/mcut/internal/bvh.h:127-133, source/bvh.cpp:638-783), a serial BFS over
// node pairs feeding a std::map.  The emitted set is {(s, c) : face_bbox_src[s] overlaps face_bbox_cut[c]} with
// closed intervals (math.h:931-941) whatever the trees look like, so the device walks its own structures (lbvh.cu):
//
//   * the mesh with more faces is the QUERY side: its leaves come in groups of at most 32 (maximal subtrees of the radix
//     tree over its sorted codes: spatially compact by construction), each with a union box;
//   * the other mesh is the TREE side: an implicit 32-wide tree, every level a dense array of boxes in [6][32] blocks;
//   * k_group_top — ONE THREAD per query group — tests the group box against the top of the tree (the top level and the
//     level below it, at most 32 + 1024 boxes, staged in shared memory once per block) and lists the (group, node) pairs
//     that overlap.  Most groups are nowhere near the other mesh and end here without a single global load beyond their
//     own box;
//   * k_traverse — ONE WARP per listed (group, node) item — walks the subtree of that node: a step pops one or two nodes and
//     tests the 32 children of each, one per lane, with six coalesced 128-byte loads; children that overlap the group box
//     are pushed (__ballot_sync/__popc slots) or, at level 0, listed as candidate leaves.  Candidates are then tested
//     exactly — double boxes — against the 32 lane-resident leaf boxes of the group, two candidates per step, hits
//     compacted with __ballot_sync/__popc into a per-warp buffer that is flushed with ONE atomicAdd per ~100 pairs.
// Pairs come out as (src_face << 32 | cut_face) and are then put in ascending order by the one-sweep sort, which
// makes the output independent of scheduling (and of how many GPUs produced it).
#include "internal.h"
#include "radix_sort.cuh"

namespace {

constexpr int WARPS_PER_BLOCK = 4;
constexpr int TBLOCK = WARPS_PER_BLOCK * 32;
constexpr int STACK_CAP = 256;
constexpr int CAND_CAP = 128;
constexpr int OUT_CAP = 128;

struct warp_scratch_t {
    uint32_t stack[STACK_CAP];
    double cand_box[CAND_CAP][6];
    uint32_t cand_face[CAND_CAP];
    unsigned long long out[OUT_CAP];
};

// A work item: a query group (its leaf range and union box travel with it: the warp that takes the item needs no other
// lookup) and the node of the other tree's level S whose box the group's box overlaps.
struct __align__(8) trav_item_t {
    uint32_t first, count, node, pad;
    float box[6];
};
static_assert(sizeof(trav_item_t) == 40, "five 8-byte words");

struct traverse_args_t {
    // query side
    const double* q_sorted_bbox; // exact face boxes in leaf order
    const uint32_t* q_sorted_faces;
    uint32_t q_nf;
    const uint2* groups; // (first sorted leaf, leaf count <= 32)
    const group_box_t* group_box; // union box of each group
    const unsigned* n_groups;
    // tree side
    wide_levels_t t;
    const uint32_t* t_sorted_faces;
    const double* t_sorted_bbox; // exact face boxes of the tree side in leaf order: the decisive test of a candidate leaf
    uint32_t t_nf;
    int start_level; // S: the level whose nodes k_group_top lists (max(1, top - 1))
    int query_is_cut; // emit (tree_face << 32 | query_face) instead
    // sharding of the query leaf range
    uint32_t shard_part, shard_nparts, shard_chunk;
    // work items
    trav_item_t* items;
    unsigned cap_items;
    // output
    unsigned long long* pairs;
    unsigned long long cap_pairs;
    result_counters_t* counters;
};

__device__ __forceinline__ void flush_out(warp_scratch_t& ws, unsigned& nout, const traverse_args_t& a)
{
    if (nout == 0) return;
    unsigned long long base = 0;
    if (lane_id() == 0) base = atomicAdd(&a.counters->n_pairs, (unsigned long long)nout);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (unsigned i = lane_id(); i < nout; i += 32) {
        const unsigned long long dst = base + i;
        if (dst < a.cap_pairs) a.pairs[dst] = ws.out[i];
        else a.counters->pair_overflow = 1u;
    }
    __syncwarp();
    nout = 0;
}

// test candidates [0, count) of the shared list (sorted-leaf indices of the tree side) against the 32 lane-resident query boxes
__device__ __forceinline__ void drain_candidates(warp_scratch_t& ws, unsigned count, const double* mybox, bool valid, uint32_t myface,
    unsigned& nout, unsigned long long& ntests, const traverse_args_t& a)
{
    const unsigned lt = lanemask_lt();
    // The tree's boxes are conservative single-precision hulls; the decisive test uses the exact face boxes, fetched here for
    // the whole batch with independent loads (boxes and face ids both live in leaf order: one round trip, and candidates
    // that are neighbours in the tree are neighbours in memory).
    __syncwarp();
    for (unsigned k = lane_id(); k < count; k += 32) {
        const uint32_t leaf = ws.cand_face[k];
        const uint32_t f = __ldg(a.t_sorted_faces + leaf);
        const double2* in = reinterpret_cast<const double2*>(a.t_sorted_bbox + 6 * (size_t)leaf);
        const double2 x = __ldg(in), y = __ldg(in + 1), z = __ldg(in + 2);
        double* cb = ws.cand_box[k];
        cb[0] = x.x; cb[1] = x.y; cb[2] = y.x; cb[3] = y.y; cb[4] = z.x; cb[5] = z.y;
        ws.cand_face[k] = f;
    }
    __syncwarp();
    auto emit_hits = [&](unsigned k, bool hit, unsigned mask) {
        if (hit) {
            const uint32_t tf = ws.cand_face[k];
            const unsigned long long pair = a.query_is_cut ? (((unsigned long long)tf << 32) | myface)
                                                           : (((unsigned long long)myface << 32) | tf);
            ws.out[nout + __popc(mask & lt)] = pair;
        }
        nout += __popc(mask);
        __syncwarp();
        if (nout > OUT_CAP - 32) flush_out(ws, nout, a);
    };
    // two candidates per step: their box loads and compares are independent
    unsigned k = 0;
    for (; k + 1 < count; k += 2) {
        const bool hit0 = valid && overlap6(mybox, ws.cand_box[k]);
        const bool hit1 = valid && overlap6(mybox, ws.cand_box[k + 1]);
        const unsigned m0 = __ballot_sync(0xffffffffu, hit0);
        const unsigned m1 = __ballot_sync(0xffffffffu, hit1);
        if (m0) emit_hits(k, hit0, m0);
        if (m1) emit_hits(k + 1, hit1, m1);
    }
    if (k < count) {
        const bool hit = valid && overlap6(mybox, ws.cand_box[k]);
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) emit_hits(k, hit, m);
    }
    ntests += count;
    __syncwarp();
}

__device__ __forceinline__ void load_group_box(const group_box_t* g, float* gbox)
{
    const float2* in = reinterpret_cast<const float2*>(g->box);
    const float2 x = __ldg(in), y = __ldg(in + 1), z = __ldg(in + 2);
    gbox[0] = x.x; gbox[1] = x.y; gbox[2] = y.x; gbox[3] = y.y; gbox[4] = z.x; gbox[5] = z.y;
}

// warp-aggregated slot allocation for threads that happen to be in the same branch
__device__ __forceinline__ unsigned alloc_item(unsigned* counter)
{
    const unsigned m = __activemask();
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if ((int)lane_id() == leader) base = atomicAdd(counter, (unsigned)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & lanemask_lt());
}

constexpr int FBLOCK = 256;

__device__ __forceinline__ void store_item(trav_item_t* it, uint2 grp, uint32_t node, const float* gbox)
{
    uint2* w = reinterpret_cast<uint2*>(it);
    w[0] = grp;
    w[1] = make_uint2(node, 0u);
    w[2] = make_uint2(__float_as_uint(gbox[0]), __float_as_uint(gbox[1]));
    w[3] = make_uint2(__float_as_uint(gbox[2]), __float_as_uint(gbox[3]));
    w[4] = make_uint2(__float_as_uint(gbox[4]), __float_as_uint(gbox[5]));
}

__global__ void __launch_bounds__(FBLOCK) k_group_top(traverse_args_t a)
{
    pdl_prologue();
    __shared__ __align__(16) float s_top[MCB_WBLOCK_FLOATS]; // the top level: one block of at most 32 boxes
    __shared__ __align__(16) float s_mid[32 * MCB_WBLOCK_FLOATS]; // the level below it: at most 32 blocks (24 KB)
    const int top = a.t.top, S = a.start_level;
    const uint32_t n_top = a.t.n[top];
    {
        // both levels with every load in flight before the first store (16-byte loads: at most 7 per thread)
        const float4* gtop = reinterpret_cast<const float4*>(a.t.boxes + (size_t)a.t.off[top] * MCB_WBLOCK_FLOATS);
        const float4* gmid = reinterpret_cast<const float4*>(a.t.boxes + (size_t)a.t.off[S] * MCB_WBLOCK_FLOATS);
        const uint32_t nmid = (S < top) ? n_top * (MCB_WBLOCK_FLOATS / 4) : 0u; // block c of level S = children of top-level node c
        constexpr int R = 32 * (MCB_WBLOCK_FLOATS / 4) / FBLOCK; // 6
        float4 rt = make_float4(0.f, 0.f, 0.f, 0.f), rm[R];
        if (threadIdx.x < MCB_WBLOCK_FLOATS / 4) rt = __ldg(gtop + threadIdx.x);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t i = r * FBLOCK + threadIdx.x;
            rm[r] = i < nmid ? __ldg(gmid + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (threadIdx.x < MCB_WBLOCK_FLOATS / 4) reinterpret_cast<float4*>(s_top)[threadIdx.x] = rt;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t i = r * FBLOCK + threadIdx.x;
            if (i < nmid) reinterpret_cast<float4*>(s_mid)[i] = rm[r];
        }
    }
    __syncthreads();
    const uint32_t ngroups = *a.n_groups;
    unsigned long long ntests = 0;
    for (uint32_t g = blockIdx.x * FBLOCK + threadIdx.x; g < ngroups; g += gridDim.x * FBLOCK) {
        const uint2 grp = __ldg(a.groups + g);
        if (a.shard_nparts > 1 && (grp.x / a.shard_chunk) % a.shard_nparts != a.shard_part) continue;
        float gbox[6];
        load_group_box(a.group_box + g, gbox);
        for (uint32_t c = 0; c < n_top; ++c) {
            float cb[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) cb[k] = s_top[k * 32 + c];
            ntests += 1ull;
            if (!overlap6f(gbox, cb)) continue;
            if (S == top) {
                const unsigned slot = alloc_item(&a.counters->work_counter);
                if (slot < a.cap_items) store_item(a.items + slot, grp, c, gbox);
                continue;
            }
            const float* blk = s_mid + c * MCB_WBLOCK_FLOATS;
            ntests += 32ull;
            for (uint32_t c2 = 0; c2 < 32u; ++c2) {
                float db[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) db[k] = blk[k * 32 + c2];
                if (!overlap6f(gbox, db)) continue; // unused slots hold an empty box
                const unsigned slot = alloc_item(&a.counters->work_counter);
                if (slot < a.cap_items) store_item(a.items + slot, grp, 32u * c + c2, gbox);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ntests += __shfl_xor_sync(0xffffffffu, ntests, o);
    if (lane_id() == 0 && ntests) atomicAdd(&a.counters->n_node_tests, ntests);
}

// the 32 children of node `i` of level `l` (l >= 1): lane k tests child k
__device__ __forceinline__ bool child_hit(const traverse_args_t& a, unsigned l, uint32_t i, const float* gbox)
{
    const float* p = a.t.boxes + ((size_t)a.t.off[l - 1] + i) * MCB_WBLOCK_FLOATS + lane_id();
    float cb[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) cb[k] = __ldg(p + k * 32);
    return overlap6f(gbox, cb);
}

__global__ void __launch_bounds__(TBLOCK) k_traverse(traverse_args_t a)
{
    pdl_prologue();
    __shared__ warp_scratch_t s_ws[WARPS_PER_BLOCK];
    warp_scratch_t& ws = s_ws[threadIdx.x >> 5];
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_lt();
    const unsigned listed = a.counters->work_counter; // written by k_group_top (it keeps counting past the capacity)
    const unsigned nitems = listed < a.cap_items ? listed : a.cap_items;
    if (listed > a.cap_items && blockIdx.x == 0 && threadIdx.x == 0) a.counters->pair_overflow = 1u;
    unsigned nout = 0;
    unsigned long long ntests = 0;
    const uint32_t warp_global = (blockIdx.x * TBLOCK + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * TBLOCK) >> 5;
    for (uint32_t it = warp_global; it < nitems; it += nwarps) {
        const uint2* iw = reinterpret_cast<const uint2*>(a.items + it);
        const uint2 grp = __ldg(iw), nd = __ldg(iw + 1), b01 = __ldg(iw + 2), b23 = __ldg(iw + 3), b45 = __ldg(iw + 4);
        const float gbox[6] = { __uint_as_float(b01.x), __uint_as_float(b01.y), __uint_as_float(b23.x), __uint_as_float(b23.y),
            __uint_as_float(b45.x), __uint_as_float(b45.y) };
        const bool valid = lane < grp.y;
        // the group's exact leaf boxes (one per lane, contiguous in leaf order): requested now, needed when the first
        // candidates are drained
        uint32_t myface = 0;
        double mybox[6] = { DBL_MAX, DBL_MAX, DBL_MAX, -DBL_MAX, -DBL_MAX, -DBL_MAX };
        if (valid) {
            myface = __ldg(a.q_sorted_faces + grp.x + lane);
            const double2* in = reinterpret_cast<const double2*>(a.q_sorted_bbox + 6 * (size_t)(grp.x + lane));
            const double2 x = __ldg(in), y = __ldg(in + 1), z = __ldg(in + 2);
            mybox[0] = x.x; mybox[1] = x.y; mybox[2] = y.x; mybox[3] = y.y; mybox[4] = z.x; mybox[5] = z.y;
        }
        auto fetch_leaf = [&]() {};
        // stack entries: level << 28 | node index (levels >= 1 have fewer than 2^28 nodes for any 32-bit face count)
        unsigned size = 1, ncand = 0;
        if (lane == 0) ws.stack[0] = ((unsigned)a.start_level << 28) | nd.x;
        __syncwarp();
        while (size > 0) {
            if (ncand > CAND_CAP - 64) { // room for the up-to-64 leaves one step can add
                fetch_leaf();
                drain_candidates(ws, ncand, mybox, valid, myface, nout, ntests, a);
                ncand = 0;
            }
            // two nodes per step when there are two (independent loads); a step adds at most 64 entries and the walk is
            // depth-first, so the stack holds at most 63 entries per level below the start
            const unsigned take = (size >= 2 && size + 64 <= STACK_CAP) ? 2u : 1u;
            const uint32_t e0 = ws.stack[size - 1];
            const uint32_t e1 = take == 2 ? ws.stack[size - 2] : 0u;
            __syncwarp();
            size -= take;
            const unsigned l0 = e0 >> 28, l1 = e1 >> 28;
            const uint32_t i0 = e0 & 0x0FFFFFFFu, i1 = e1 & 0x0FFFFFFFu;
            const bool h0 = child_hit(a, l0, i0, gbox);
            const bool h1 = take == 2 ? child_hit(a, l1, i1, gbox) : false;
            ntests += 32ull * take;
            const unsigned m0 = __ballot_sync(0xffffffffu, h0);
            const unsigned m1 = __ballot_sync(0xffffffffu, h1);
            // the second entry first: it was deeper in the stack, so its children go below the first one's
            if (m1) {
                if (l1 == 1u) {
                    if (h1) ws.cand_face[ncand + __popc(m1 & lt)] = 32u * i1 + lane;
                    ncand += __popc(m1);
                } else {
                    if (h1) ws.stack[size + __popc(m1 & lt)] = ((l1 - 1u) << 28) | (32u * i1 + lane);
                    size += __popc(m1);
                }
            }
            if (m0) {
                if (l0 == 1u) {
                    if (h0) ws.cand_face[ncand + __popc(m0 & lt)] = 32u * i0 + lane;
                    ncand += __popc(m0);
                } else {
                    if (h0) ws.stack[size + __popc(m0 & lt)] = ((l0 - 1u) << 28) | (32u * i0 + lane);
                    size += __popc(m0);
                }
            }
            __syncwarp();
        }
        if (ncand) {
            fetch_leaf();
            drain_candidates(ws, ncand, mybox, valid, myface, nout, ntests, a);
        }
    }
    flush_out(ws, nout, a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ntests += __shfl_xor_sync(0xffffffffu, ntests, o);
    if (lane == 0 && ntests) atomicAdd(&a.counters->n_node_tests, ntests / 32ull);
}

static int bits_for(uint32_t n)
{
    int b = 1;
    while (b < 32 && (1ull << b) < (unsigned long long)n) ++b;
    return b;
}

} // namespace

static rsort::pass_desc pair_passes(uint32_t nsf, uint32_t ncf)
{
    // ascending (src << 32 | cut): only the bits that can be set take part in the sort
    return rsort::make_passes(0, bits_for(ncf), 32, 32 + bits_for(nsf));
}

int traverse_reserve(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_result* res)
{
    if (res->cap_pairs == 0) {
        size_t want = 4ull * ((size_t)src->nf + cut->nf);
        if (want < (1u << 20)) want = 1u << 20;
        res->cap_pairs = want;
    }
    // work items (group, node): a few per group that comes near the other mesh; the list grows with the pair capacity when
    // an input needs more (k_group_top keeps counting, mcb200_result_counts raises the capacity)
    const size_t qf = src->nf > cut->nf ? src->nf : cut->nf;
    size_t want_items = qf > (1u << 18) ? qf : (1u << 18);
    if (want_items < res->cap_pairs / 8) want_items = res->cap_pairs / 8;
    if (want_items > 0x7FFFFFFFull) want_items = 0x7FFFFFFFull;
    res->cap_items = want_items;
    MCB_TRY(ctx->reserve(res->pairs, sizeof(unsigned long long) * res->cap_pairs));
    MCB_TRY(ctx->reserve(res->pairs_a, sizeof(unsigned long long) * res->cap_pairs));
    MCB_TRY(ctx->reserve(res->pairs_b, sizeof(unsigned long long) * res->cap_pairs));
    MCB_TRY(ctx->reserve(res->counters, sizeof(result_counters_t)));
    MCB_TRY(ctx->reserve(res->items, sizeof(trav_item_t) * res->cap_items));
    return 0;
}

int result_reset_counters(mcb200_ctx* ctx, mcb200_result* res)
{
    // all zero, bad_face starts at "none"
    unsigned* w = res->counters.as<unsigned>();
    const size_t bf = offsetof(result_counters_t, bad_face) / 4, nw = sizeof(result_counters_t) / 4;
    fill_list_t fl {}; // three disjoint ranges: entries of one launch are not ordered against each other
    fl.add(w, bf, 0u);
    fl.add(w + bf, 1, 0xFFFFFFFFu);
    fl.add(w + bf + 1, nw - bf - 1, 0u);
    MCB_LAUNCH(ctx, k_fill, 1, 256, 0, fl);
    res->narrow_counters_fresh = true;
    return 0;
}

// the traversal kernels alone, on ctx->cur: pairs land in res->pairs in emission order
int traverse_pairs(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_result* res)
{
    if (!src->built || !cut->built) {
        ctx->set_error("bvh_intersect: both meshes need mcb200_bvh_build first", __FILE__, __LINE__);
        return MCB200_ERR_INVALID;
    }
    const bool query_is_cut = cut->nf > src->nf;
    const mcb200_mesh* q = query_is_cut ? cut : src;
    const mcb200_mesh* t = query_is_cut ? src : cut;
    MCB_TRY(traverse_reserve(ctx, src, cut, res));
    if (!res->counters_zeroed) MCB_TRY(result_reset_counters(ctx, res));
    res->counters_zeroed = false;
    res->nsf = src->nf;
    res->nf_ps = src->nf + cut->nf;
    res->h_valid = false;
    res->have_narrow = false;
    res->records_sorted_valid = false;
    res->tests_sorted_valid = false;
    res->pairs_sorted = nullptr;

    traverse_args_t a;
    a.q_sorted_bbox = q->sorted_bbox.as<double>();
    a.q_sorted_faces = q->sorted_faces.as<uint32_t>();
    a.q_nf = q->nf;
    a.t = *t->lv;
    a.t.boxes = t->wide.as<float>();
    a.start_level = a.t.top >= 2 ? a.t.top - 1 : 1;
    a.t_sorted_faces = t->sorted_faces.as<uint32_t>();
    a.t_sorted_bbox = t->sorted_bbox.as<double>();
    a.t_nf = t->nf;
    a.query_is_cut = query_is_cut ? 1 : 0;
    a.shard_part = res->shard_part;
    a.shard_nparts = res->shard_nparts;
    a.shard_chunk = res->shard_chunk ? res->shard_chunk : 4096u;
    a.items = res->items.as<trav_item_t>();
    a.cap_items = (unsigned)res->cap_items;
    a.pairs = res->pairs.as<unsigned long long>();
    a.cap_pairs = res->cap_pairs;
    a.counters = res->counters.as<result_counters_t>();
    a.groups = q->groups.as<uint2>();
    a.n_groups = reinterpret_cast<const unsigned*>(q->groups.as<uint2>() + q->nf);
    a.group_box = q->group_box.as<group_box_t>();

    // one thread per group: there are about nf/16 of them (the kernel strides over the device-side count)
    {
        const unsigned fb = div_up((size_t)q->nf / 8u + 1u, FBLOCK);
        const unsigned fmax = (unsigned)ctx->num_sms * 8u;
        MCB_LAUNCH(ctx, k_group_top, fb < fmax ? fb : fmax, FBLOCK, 0, a);
    }
    // one warp per item, items dealt round-robin to a grid sized for the machine (usually fewer items than warps)
    const unsigned max_blocks = (unsigned)ctx->num_sms * 12u;
    const unsigned want_blocks = div_up((size_t)q->nf / 16u + 1u, WARPS_PER_BLOCK);
    const unsigned grid = want_blocks < max_blocks ? (want_blocks ? want_blocks : 1u) : max_blocks;
    MCB_LAUNCH(ctx, k_traverse, grid, TBLOCK, 0, a);
    res->have_pairs = true;
    return 0;
}

int sort_pairs_reserve(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_result* res)
{
    const rsort::pass_desc pd = pair_passes(src->nf, cut->nf);
    return rsort::reserve_scratch<unsigned long long>(ctx, res->cap_pairs, pd.npasses, false, false);
}

// canonical order of the pair set, on ctx->cur with ctx's current scratch set; res->pairs itself is left untouched
// (the narrowphase may be reading it on the other lane)
int sort_pairs(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_result* res)
{
    const rsort::pass_desc pd = pair_passes(src->nf, cut->nf);
    unsigned long long* out = nullptr;
    MCB_TRY((rsort::sort<unsigned long long, uint32_t, false>(ctx, res->pairs.as<unsigned long long>(),
        res->pairs_a.as<unsigned long long>(), res->pairs_b.as<unsigned long long>(), nullptr, nullptr, nullptr,
        &res->counters.as<result_counters_t>()->n_pairs, res->cap_pairs, pd, &out, nullptr)));
    res->pairs_sorted = out;
    return 0;
}
