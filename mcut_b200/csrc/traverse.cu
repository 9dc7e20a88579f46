// mcut_b200/csrc/traverse.cu — (2) BVH x BVH overlap traversal.
//
// Replaces intersectOIBVHs() (include/mcut/internal/bvh.h:127-133, source/bvh.cpp:638-783), a serial BFS over
// node pairs feeding a std::map.  The emitted set is {(s, c) : face_bbox_src[s] overlaps face_bbox_cut[c]} with
// closed intervals (math.h:931-941) whatever the trees look like, so the device walks its own LBVHs:
//
//   * the mesh with more faces is the QUERY side; its leaves are grouped by the query mesh's OWN tree: a group is a
//     maximal subtree with at most 32 leaves (listed by lbvh.cu's k_tree).  Such treelets are spatially compact — cutting the Morton
//     order into fixed runs of 32 is not: a run that straddles an octant boundary has a union box spanning the mesh;
//   * most groups are nowhere near the other mesh.  k_group_filter settles those with ONE THREAD per group: a depth-first
//     walk of the other tree with the group's union box that stops at the first leaf it reaches ("live") or when the stack
//     runs empty ("dead").  Only the live groups — a few percent, the ones along the intersection curve — get a warp;
//     before this split every warp had ~13 groups dealt to it and the few warps holding two or three live ones set the time;
//   * one warp owns a live group: lane l keeps leaf l's box in registers, the warp keeps the group's union box;
//   * the warp walks the other mesh's tree with a stack in shared memory, up to 32 nodes per step — one node per
//     lane, one 128-byte line per node carrying both children's boxes; surviving internal children are pushed with
//     __ballot_sync/__popc slots, surviving leaves go to a shared candidate list;
//   * candidates are then tested against the 32 lane-resident leaf boxes (box broadcast from shared memory) and hits
//     are compacted with __ballot_sync/__popc into a per-warp buffer that is flushed to global memory with ONE
//     atomicAdd per ~200 pairs.
// Pairs come out as (src_face << 32 | cut_face) and are then put in ascending order by the one-sweep sort, which
// makes the output independent of scheduling (and of how many GPUs produced it).
#include "internal.h"
#include "radix_sort.cuh"

namespace {

constexpr int WARPS_PER_BLOCK = 4;
constexpr int TBLOCK = WARPS_PER_BLOCK * 32;
constexpr int STACK_CAP = 512;
constexpr int CAND_CAP = 96;
constexpr int OUT_CAP = 128;
constexpr int GROUP_BATCH = 2;

struct warp_scratch_t {
    uint32_t stack[STACK_CAP];
    double cand_box[CAND_CAP][6];
    uint32_t cand_face[CAND_CAP];
    unsigned long long out[OUT_CAP];
};

// A live group and where its warp resumes the walk: the filter thread's unexplored frontier (the node whose leaf child it
// hit + its stack).  A frontier that does not fit restarts at the root.
constexpr int LIVE_NODES = 14;
struct __align__(64) live_group_t {
    uint32_t group, count;
    uint32_t node[LIVE_NODES];
};
static_assert(sizeof(live_group_t) == 64, "one 64-byte row per live group");

struct traverse_args_t {
    // query side
    const double* q_face_bbox;
    const uint32_t* q_sorted_faces;
    uint32_t q_nf;
    const uint2* groups; // (first sorted leaf, leaf count <= 32)
    const group_up_t* group_box; // union box of each group (written by the refit)
    const unsigned* n_groups;
    live_group_t* live; // groups that reach a leaf of the other tree (k_group_filter); count in counters->work_counter
    const double* t_root; // mesh AABB of the tree side (6 doubles)
    // tree side
    const bvh_node_t* t_nodes;
    const uint32_t* t_sorted_faces;
    const double* t_face_bbox; // exact face boxes of the tree side: the decisive test of a candidate leaf
    uint32_t t_nf;
    int query_is_cut; // emit (tree_face << 32 | query_face) instead
    // sharding of the query leaf range
    uint32_t shard_part, shard_nparts, shard_chunk;
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

// test candidates [first, first+count) of the shared list against the 32 lane-resident query boxes
__device__ __forceinline__ void drain_candidates(warp_scratch_t& ws, unsigned first, unsigned count, const double* mybox,
    bool valid, uint32_t myface, unsigned& nout, unsigned long long& ntests, const traverse_args_t& a)
{
    const unsigned lt = lanemask_lt();
    // The tree's node boxes are conservative single-precision hulls; the decisive test uses the exact face boxes, fetched
    // here for the whole batch with independent loads (one round trip).
    __syncwarp();
    for (unsigned k = lane_id(); k < count; k += 32) {
        const double2* in = reinterpret_cast<const double2*>(a.t_face_bbox + 6 * (size_t)ws.cand_face[first + k]);
        const double2 x = __ldg(in), y = __ldg(in + 1), z = __ldg(in + 2);
        double* cb = ws.cand_box[first + k];
        cb[0] = x.x; cb[1] = x.y; cb[2] = y.x; cb[3] = y.y; cb[4] = z.x; cb[5] = z.y;
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
    unsigned k = first;
    for (; k + 1 < first + count; k += 2) {
        const bool hit0 = valid && overlap6(mybox, ws.cand_box[k]);
        const bool hit1 = valid && overlap6(mybox, ws.cand_box[k + 1]);
        const unsigned m0 = __ballot_sync(0xffffffffu, hit0);
        const unsigned m1 = __ballot_sync(0xffffffffu, hit1);
        if (m0) emit_hits(k, hit0, m0);
        if (m1) emit_hits(k + 1, hit1, m1);
    }
    if (k < first + count) {
        const bool hit = valid && overlap6(mybox, ws.cand_box[k]);
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) emit_hits(k, hit, m);
    }
    ntests += count;
}

// one 64-byte node record: four 16-byte loads
__device__ __forceinline__ void fetch_node(const bvh_node_t* nd, float* lb, float* rb, uint32_t& left, uint32_t& right)
{
    const uint4* p = reinterpret_cast<const uint4*>(nd);
    const uint4 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2), q3 = __ldg(p + 3);
    lb[0] = __uint_as_float(q0.x); lb[1] = __uint_as_float(q0.y); lb[2] = __uint_as_float(q0.z); lb[3] = __uint_as_float(q0.w);
    lb[4] = __uint_as_float(q1.x); lb[5] = __uint_as_float(q1.y);
    rb[0] = __uint_as_float(q1.z); rb[1] = __uint_as_float(q1.w);
    rb[2] = __uint_as_float(q2.x); rb[3] = __uint_as_float(q2.y); rb[4] = __uint_as_float(q2.z); rb[5] = __uint_as_float(q2.w);
    left = q3.x;
    right = q3.y;
}

__device__ __forceinline__ void load_group_box(const group_up_t* g, float* gbox)
{
    const float2* in = reinterpret_cast<const float2*>(g->box);
    const float2 x = __ldg(in), y = __ldg(in + 1), z = __ldg(in + 2);
    gbox[0] = x.x; gbox[1] = x.y; gbox[2] = y.x; gbox[3] = y.y; gbox[4] = z.x; gbox[5] = z.y;
}

constexpr int FBLOCK = 128;
constexpr int FSTACK = 48; // private depth-first stack; a walk that would outgrow it declares the group live (conservative)
constexpr int FVISITS = 12; // so does a walk that has not settled after this many nodes: its warp finishes it, 32 nodes a step

__global__ void __launch_bounds__(FBLOCK) k_group_filter(traverse_args_t a)
{
    pdl_prologue();
    const uint32_t ngroups = *a.n_groups;
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_lt();
    unsigned long long ntests = 0;
    for (uint32_t g0 = (blockIdx.x * FBLOCK + threadIdx.x) & ~31u; g0 < ngroups; g0 += gridDim.x * FBLOCK) {
        const uint32_t g = g0 + lane;
        bool live = false;
        uint32_t stack[FSTACK];
        int size = 0;
        if (g < ngroups) {
            const uint2 grp = __ldg(a.groups + g);
            const bool mine = !(a.shard_nparts > 1 && (grp.x / a.shard_chunk) % a.shard_nparts != a.shard_part);
            if (mine) {
                float gbox[6];
                load_group_box(a.group_box + g, gbox);
                float troot[6];
                {
                    double td[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) td[k] = __ldg(a.t_root + k);
                    box_to_float(td, troot);
                }
                ntests += 1ull;
                if (overlap6f(gbox, troot)) {
                    size = 1;
                    stack[0] = 0u;
                    int visits = 0;
                    while (size > 0 && !live) {
                        const uint32_t node = stack[size - 1];
                        float lb[6], rb[6];
                        uint2 ch;
                        fetch_node(a.t_nodes + node, lb, rb, ch.x, ch.y);
                        const bool hitL = overlap6f(gbox, lb);
                        const bool hitR = (ch.y != MCB200_NULL) && overlap6f(gbox, rb);
                        ntests += 2ull;
                        // live: a leaf is reached (the node stays on the stack: its warp collects the leaves), or the walk is
                        // taking long, or the stack is about to overflow
                        if ((hitL && (ch.x & MCB_LEAF_BIT)) || (hitR && (ch.y & MCB_LEAF_BIT)) || ++visits >= FVISITS || size + 1 > FSTACK) {
                            live = true;
                        } else {
                            --size;
                            if (hitR) stack[size++] = ch.y;
                            if (hitL) stack[size++] = ch.x; // left first out: depth-first, left to right
                        }
                    }
                }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, live);
        if (m) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&a.counters->work_counter, (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (live) {
                live_group_t rec;
                rec.group = g;
                const bool fits = size <= LIVE_NODES;
                rec.count = fits ? (uint32_t)size : 1u;
#pragma unroll
                for (int k = 0; k < LIVE_NODES; ++k) rec.node[k] = (fits && k < size) ? stack[k] : 0u;
                a.live[base + __popc(m & lt)] = rec;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ntests += __shfl_xor_sync(0xffffffffu, ntests, o);
    if (lane == 0 && ntests) atomicAdd(&a.counters->n_node_tests, ntests);
}

__global__ void __launch_bounds__(TBLOCK) k_traverse(traverse_args_t a)
{
    pdl_prologue();
    __shared__ warp_scratch_t s_ws[WARPS_PER_BLOCK];
    warp_scratch_t& ws = s_ws[threadIdx.x >> 5];
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_lt();
    const uint32_t nlive = a.counters->work_counter; // written by k_group_filter
    unsigned nout = 0;
    unsigned long long ntests = 0;

    // Live groups are dealt to warps round-robin (static); there are usually fewer of them than warps.
    const uint32_t warp_global = (blockIdx.x * TBLOCK + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * TBLOCK) >> 5;
    {
        {
            for (uint32_t li = warp_global; li < nlive; li += nwarps) {
            // ---- descriptor + union box (from the refit); the 32 leaf boxes are fetched when the first candidates are drained
            const live_group_t* lg = a.live + li;
            const uint32_t g = __ldg(&lg->group);
            const uint32_t nstart = __ldg(&lg->count);
            const uint32_t start_node = lane < nstart ? __ldg(&lg->node[lane]) : 0u;
            const uint2 grp = __ldg(a.groups + g);
            float gbox[6];
            load_group_box(a.group_box + g, gbox);
            const uint32_t q = grp.x + lane;
            const bool valid = lane < grp.y;
            bool have_leaf = false;
            uint32_t myface = 0;
            double mybox[6] = { DBL_MAX, DBL_MAX, DBL_MAX, -DBL_MAX, -DBL_MAX, -DBL_MAX };
            auto fetch_leaf = [&]() {
                if (have_leaf) return;
                have_leaf = true;
                if (valid) {
                    myface = __ldg(a.q_sorted_faces + q);
                    const double2* in = reinterpret_cast<const double2*>(a.q_face_bbox + 6 * (size_t)myface);
                    const double2 x = __ldg(in), y = __ldg(in + 1), z = __ldg(in + 2);
                    mybox[0] = x.x; mybox[1] = x.y; mybox[2] = y.x; mybox[3] = y.y; mybox[4] = z.x; mybox[5] = z.y;
                }
            };

            // ---- walk the tree ----
            unsigned size = nstart, ncand = 0;
            if (lane < nstart) ws.stack[lane] = start_node;
            __syncwarp();
            while (size > 0) {
                while (ncand > 32) { // keep room for the up-to-64 leaves one step can add (CAND_CAP = 32 + 64)
                    fetch_leaf();
                    drain_candidates(ws, ncand - 32, 32, mybox, valid, myface, nout, ntests, a);
                    ncand -= 32;
                }
                // wide steps while the stack has room; near the cap fall back to one node per step, whose growth is
                // bounded by the tree depth (LIFO order keeps it a depth-first walk)
                const unsigned width = (size <= STACK_CAP - 160) ? 32u : 1u; // 160 = 32 (wide growth) + 128 (depth bound)
                const unsigned take = size < width ? size : width;
                const bool active = lane < take;
                bool hitL = false, hitR = false;
                uint32_t left = 0, right = 0;
                if (active) {
                    const uint32_t node = ws.stack[size - 1 - lane];
                    float lb[6], rb[6];
                    fetch_node(a.t_nodes + node, lb, rb, left, right);
                    hitL = overlap6f(gbox, lb);
                    hitR = (right != MCB200_NULL) && overlap6f(gbox, rb);
                }
                __syncwarp();
                size -= take;
                ntests += 2ull * take;
                // internal children -> stack
                {
                    const bool pl = hitL && !(left & MCB_LEAF_BIT);
                    const unsigned ml = __ballot_sync(0xffffffffu, pl);
                    if (pl) ws.stack[size + __popc(ml & lt)] = left;
                    size += __popc(ml);
                    const bool pr = hitR && !(right & MCB_LEAF_BIT);
                    const unsigned mr = __ballot_sync(0xffffffffu, pr);
                    if (pr) ws.stack[size + __popc(mr & lt)] = right;
                    size += __popc(mr);
                }
                // leaf children -> candidate list (the child id IS the face id; its exact box is fetched when the list is drained)
                {
                    const bool cl = hitL && (left & MCB_LEAF_BIT);
                    const unsigned ml = __ballot_sync(0xffffffffu, cl);
                    if (cl) ws.cand_face[ncand + __popc(ml & lt)] = left & ~MCB_LEAF_BIT;
                    ncand += __popc(ml);
                    const bool cr = hitR && (right & MCB_LEAF_BIT);
                    const unsigned mr = __ballot_sync(0xffffffffu, cr);
                    if (cr) ws.cand_face[ncand + __popc(mr & lt)] = right & ~MCB_LEAF_BIT;
                    ncand += __popc(mr);
                }
                __syncwarp();
            }
            if (ncand) {
                fetch_leaf();
                drain_candidates(ws, 0, ncand, mybox, valid, myface, nout, ntests, a);
            }
            }
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
    MCB_TRY(ctx->reserve(res->pairs, sizeof(unsigned long long) * res->cap_pairs));
    MCB_TRY(ctx->reserve(res->pairs_a, sizeof(unsigned long long) * res->cap_pairs));
    MCB_TRY(ctx->reserve(res->pairs_b, sizeof(unsigned long long) * res->cap_pairs));
    MCB_TRY(ctx->reserve(res->counters, sizeof(result_counters_t)));
    MCB_TRY(ctx->reserve(res->live_groups, sizeof(live_group_t) * (size_t)(src->nf > cut->nf ? src->nf : cut->nf)));
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

// the traversal kernel alone, on ctx->cur: pairs land in res->pairs in emission order
int traverse_pairs(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_result* res)
{
    if (!src->built || !cut->built) {
        ctx->set_error("bvh_intersect: both meshes need mcb200_bvh_build first", __FILE__, __LINE__);
        return MCB200_ERR_INVALID;
    }
    const bool query_is_cut = cut->nf > src->nf;
    const mcb200_mesh* q = query_is_cut ? cut : src;
    const mcb200_mesh* t = query_is_cut ? src : cut;
    if (!t->has_nodes) {
        ctx->set_error("bvh_intersect: the tree-side mesh was built query-only (internal: rebuild it fully first)", __FILE__, __LINE__);
        return MCB200_ERR_INTERNAL;
    }
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
    a.q_face_bbox = q->face_bbox.as<double>();
    a.q_sorted_faces = q->sorted_faces.as<uint32_t>();
    a.q_nf = q->nf;
    a.t_nodes = t->nodes.as<bvh_node_t>();
    a.t_sorted_faces = t->sorted_faces.as<uint32_t>();
    a.t_face_bbox = t->face_bbox.as<double>();
    a.t_nf = t->nf;
    a.query_is_cut = query_is_cut ? 1 : 0;
    a.shard_part = res->shard_part;
    a.shard_nparts = res->shard_nparts;
    a.shard_chunk = res->shard_chunk ? res->shard_chunk : 4096u;
    a.pairs = res->pairs.as<unsigned long long>();
    a.cap_pairs = res->cap_pairs;
    a.counters = res->counters.as<result_counters_t>();
    // query groups = the maximal <=32-leaf treelets of the query mesh's own tree, listed by its refit kernel
    a.groups = q->groups.as<uint2>();
    a.n_groups = reinterpret_cast<const unsigned*>(q->groups.as<uint2>() + q->nf);
    a.group_box = q->group_up.as<group_up_t>();
    a.t_root = reinterpret_cast<const double*>(t->root.as<unsigned long long>() + 6);

    a.live = res->live_groups.as<live_group_t>();
    {
        const unsigned fb = div_up((size_t)q->nf / 8u + 1u, FBLOCK); // about two thirds of the leaves' groups per pass of the grid
        const unsigned fmax = (unsigned)ctx->num_sms * 16u;
        MCB_LAUNCH(ctx, k_group_filter, fb < fmax ? fb : fmax, FBLOCK, 0, a);
    }
    // a persistent grid sized for the machine
    const unsigned max_blocks = (unsigned)ctx->num_sms * 8u;
    const unsigned want_blocks = div_up(div_up((size_t)q->nf / 8u + 1u, GROUP_BATCH), WARPS_PER_BLOCK);
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
