// mcut_b200/csrc/traverse.cu — (2) BVH x BVH overlap traversal.
//
// Replaces intersectOIBVHs() (include/mcut/internal/bvh.h:127-133, source/bvh.cpp:638-783), a serial BFS over
// node pairs feeding a std::map.  The emitted set is {(s, c) : face_bbox_src[s] overlaps face_bbox_cut[c]} with
// closed intervals (math.h:931-941) whatever the trees look like, so the device walks its own structures (lbvh.cu):
//
//   * the mesh with more faces is the QUERY side: its leaves come in groups of at most 32 (maximal subtrees of the radix
//     tree over its sorted codes: spatially compact by construction), each with a union box;
//   * the other mesh is the TREE side: an implicit 32-wide tree, every level a dense array of boxes in [6][32] blocks;
//   * ONE kernel.  A block stages the top of the tree (the top level and the level below it: at most 32 + 1024 boxes) in
//     shared memory.  A warp takes eight groups at a time off a global ticket; eight lanes test their group's box against
//     the top level — shared memory only: most groups are nowhere near the other mesh and end here, a dozen
//     instructions after their one coalesced load.  For each group that survives, the whole warp walks the tree below:
//     a step pops one or two nodes and tests the 32 children of each, one per lane, with six coalesced 128-byte loads;
//     children that overlap the group box are pushed (__ballot_sync/__popc slots) or, at level 0, listed as candidate
//     leaves.  Candidates are then tested exactly — double boxes, read in leaf order — against the 32 lane-resident leaf
//     boxes of the group, two candidates per step, hits compacted with __ballot_sync/__popc into a per-warp buffer that
//     is flushed with ONE atomicAdd per ~100 pairs.
// Pairs come out as (src_face << 32 | cut_face) and are then put in ascending order by the one-sweep sort, which
// makes the output independent of scheduling (and of how many GPUs produced it).
#include "internal.h"
#include "radix_sort.cuh"

namespace {

constexpr int WARPS_PER_BLOCK = 8;
constexpr int TBLOCK = WARPS_PER_BLOCK * 32;
constexpr int STACK_CAP = 192;
constexpr int CAND_CAP = 128; // candidate leaves listed before a drain
constexpr int BOX_BATCH = 32; // exact boxes staged per drain round
constexpr int OUT_CAP = 128;
constexpr int DYN_BATCH = 8; // groups per ticket afterwards

struct warp_scratch_t {
    uint32_t stack[STACK_CAP];
    double cand_box[BOX_BATCH][6];
    uint32_t cand_face[CAND_CAP];
    uint32_t cand_tf[BOX_BATCH];
    unsigned long long out[OUT_CAP];
};

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
    unsigned static_batch; // groups every warp starts with, taken with stride nwarps (no ticket); <= 32
    int start_level; // S: the lowest level staged in shared memory (max(1, top - 1)); the walk in global memory starts below it
    int query_is_cut; // emit (tree_face << 32 | query_face) instead
    // sharding of the query leaf range
    uint32_t shard_part, shard_nparts, shard_chunk;
    // output
    unsigned long long* pairs;
    unsigned long long cap_pairs;
    unsigned* src_count; // per source face: number of pairs emitted (input of the counting order, sort_pairs)
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
    unsigned& nout, unsigned long long& ntests, unsigned& mycnt, const traverse_args_t& a)
{
    const unsigned lt = lanemask_lt();
    const unsigned lane = lane_id();
    auto emit_hits = [&](unsigned k, bool hit, unsigned mask) {
        if (hit) {
            const uint32_t tf = ws.cand_tf[k];
            const unsigned long long pair = a.query_is_cut ? (((unsigned long long)tf << 32) | myface)
                                                           : (((unsigned long long)myface << 32) | tf);
            ws.out[nout + __popc(mask & lt)] = pair;
            ++mycnt; // pairs of this lane's query face
        }
        // pairs per SOURCE face: the query faces' counts are stored when their group is done (a face belongs to one group);
        // a tree-side source face collects its count from every group that meets it
        if (a.query_is_cut && lane == 0) atomicAdd(a.src_count + ws.cand_tf[k], (unsigned)__popc(mask));
        nout += __popc(mask);
        __syncwarp();
        if (nout > OUT_CAP - 32) flush_out(ws, nout, a);
    };
    // The tree's boxes are conservative single-precision hulls; the decisive test uses the exact face boxes: 32 candidates per
    // round, one per lane, boxes and face ids read in leaf order (candidates that are neighbours in the tree are neighbours in
    // memory), staged in shared memory and then broadcast against the 32 lane-resident query boxes, two per step.
    __syncwarp();
    for (unsigned base = 0; base < count; base += BOX_BATCH) {
        const unsigned n = count - base < (unsigned)BOX_BATCH ? count - base : (unsigned)BOX_BATCH;
        if (lane < n) {
            const uint32_t leaf = ws.cand_face[base + lane];
            const uint32_t f = __ldg(a.t_sorted_faces + leaf);
            const double2* in = reinterpret_cast<const double2*>(a.t_sorted_bbox + 6 * (size_t)leaf);
            const double2 x = __ldg(in), y = __ldg(in + 1), z = __ldg(in + 2);
            double* cb = ws.cand_box[lane];
            cb[0] = x.x; cb[1] = x.y; cb[2] = y.x; cb[3] = y.y; cb[4] = z.x; cb[5] = z.y;
            ws.cand_tf[lane] = f;
        }
        __syncwarp();
        unsigned k = 0;
        for (; k + 1 < n; k += 2) {
            const bool hit0 = valid && overlap6(mybox, ws.cand_box[k]);
            const bool hit1 = valid && overlap6(mybox, ws.cand_box[k + 1]);
            const unsigned m0 = __ballot_sync(0xffffffffu, hit0);
            const unsigned m1 = __ballot_sync(0xffffffffu, hit1);
            if (m0) emit_hits(k, hit0, m0);
            if (m1) emit_hits(k + 1, hit1, m1);
        }
        if (k < n) {
            const bool hit = valid && overlap6(mybox, ws.cand_box[k]);
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) emit_hits(k, hit, m);
        }
        __syncwarp();
    }
    ntests += count;
}

__device__ __forceinline__ void load_group_box(const group_box_t* g, float* gbox)
{
    const float2* in = reinterpret_cast<const float2*>(g->box);
    const float2 x = __ldg(in), y = __ldg(in + 1), z = __ldg(in + 2);
    gbox[0] = x.x; gbox[1] = x.y; gbox[2] = y.x; gbox[3] = y.y; gbox[4] = z.x; gbox[5] = z.y;
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

__global__ void __launch_bounds__(TBLOCK, 3) k_traverse(traverse_args_t a)
{
    pdl_prologue();
    __shared__ warp_scratch_t s_ws[WARPS_PER_BLOCK];
    warp_scratch_t& ws = s_ws[threadIdx.x >> 5];
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_lt();
    const int top = a.t.top, S = a.start_level;
    const float* const mid = a.t.boxes + (size_t)a.t.off[S] * MCB_WBLOCK_FLOATS; // level S: block c = children of top-level node c
    const uint32_t ngroups = *a.n_groups;
    unsigned nout = 0;
    unsigned long long ntests = 0, ntop = 0;
    // lanes = top-level nodes: their boxes sit in registers for the life of the warp (unused slots hold an empty box); the
    // level below (at most 1024 boxes, 24 KB) is read through L1 when a group gets that far
    float topbox[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) topbox[k] = __ldg(a.t.boxes + (size_t)a.t.off[top] * MCB_WBLOCK_FLOATS + k * 32 + lane);
    const uint32_t warp_global = (blockIdx.x * TBLOCK + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * TBLOCK) >> 5;
    const bool dynamic_rounds = (unsigned long long)nwarps * a.static_batch < ngroups; // more groups than the static round covers
    bool first = true;
    for (;;) {
        // ---- the next groups of this warp, one per lane: a static share first, then tickets.  (One ticket per batch from
        //      every warp from the start put ~10^4 atomics on one address: they serialise in L2 and cost more than the walk.) ----
        // The static share is STRIDED (group warp + k * nwarps): the groups along the intersection curve are neighbours in
        // the list, and a warp that owned sixteen consecutive ones would work alone long after the others have finished.
        uint32_t g0 = 0, gstride = 1;
        unsigned want = 0;
        if (first) {
            g0 = warp_global;
            gstride = nwarps;
            want = a.static_batch;
            first = false;
        } else {
            if (!dynamic_rounds) break;
            if (lane == 0) g0 = atomicAdd(&a.counters->work_counter, (unsigned)DYN_BATCH);
            g0 = nwarps * a.static_batch + __shfl_sync(0xffffffffu, g0, 0);
            want = DYN_BATCH;
        }
        if (g0 >= ngroups) break;
        uint2 my_grp = make_uint2(0u, 0u);
        float my_box[6] = { FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX }; // an empty box overlaps nothing
        const unsigned long long mine = (unsigned long long)g0 + (unsigned long long)lane * gstride;
        if (lane < want && mine < ngroups) {
            my_grp = __ldg(a.groups + mine);
            if (!(a.shard_nparts > 1 && (my_grp.x / a.shard_chunk) % a.shard_nparts != a.shard_part)) load_group_box(a.group_box + mine, my_box);
        }
        const unsigned batch = want; // lanes past the end of the list hold an empty box
        // ---- lanes = top-level nodes (their boxes sit in registers for the life of the warp); one group at a time ----
        for (unsigned j = 0; j < batch; ++j) {
            float gbox[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) gbox[k] = __shfl_sync(0xffffffffu, my_box[k], j);
            unsigned hits = __ballot_sync(0xffffffffu, overlap6f(gbox, topbox)); // unused top slots hold an empty box
            ntop += 1ull;
            if (!hits) continue;
            const uint2 grp = make_uint2(__shfl_sync(0xffffffffu, my_grp.x, j), __shfl_sync(0xffffffffu, my_grp.y, j));
            // stack entries: level << 28 | node index (levels >= 1 have fewer than 2^28 nodes for any 32-bit face count)
            unsigned size = 0, ncand = 0;
            if (S == top) {
                if ((hits >> lane) & 1u) ws.stack[__popc(hits & lt)] = ((unsigned)S << 28) | lane;
                size = __popc(hits);
                hits = 0u;
            } else {
                while (hits) {
                    const uint32_t c = (uint32_t)__ffs((int)hits) - 1u;
                    hits &= hits - 1u;
                    const float* blk = mid + c * MCB_WBLOCK_FLOATS + lane;
                    float cb[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) cb[k] = __ldg(blk + k * 32);
                    const bool h = overlap6f(gbox, cb); // unused slots hold an empty box
                    const unsigned m = __ballot_sync(0xffffffffu, h);
                    if (h) ws.stack[size + __popc(m & lt)] = ((unsigned)S << 28) | (32u * c + lane);
                    size += __popc(m);
                    ntests += 32ull;
                    if (size > STACK_CAP - 96) break; // (cannot happen with <= 1024 level-S nodes and a 256-entry stack unless hits > 5)
                }
                // a group overlapping more level-S nodes than the stack holds: finish the remaining top nodes after the walk
            }
            __syncwarp();
            if (size == 0 && hits == 0u) continue;
            const bool valid = lane < grp.y;
            // the group's exact leaf boxes (one per lane, contiguous in leaf order): requested now, needed when the first
            // candidates are drained
            uint32_t myface = 0;
            unsigned mycnt = 0;
            double mybox[6] = { DBL_MAX, DBL_MAX, DBL_MAX, -DBL_MAX, -DBL_MAX, -DBL_MAX };
            if (valid) {
                myface = __ldg(a.q_sorted_faces + grp.x + lane);
                const double2* in = reinterpret_cast<const double2*>(a.q_sorted_bbox + 6 * (size_t)(grp.x + lane));
                const double2 x = __ldg(in), y = __ldg(in + 1), z = __ldg(in + 2);
                mybox[0] = x.x; mybox[1] = x.y; mybox[2] = y.x; mybox[3] = y.y; mybox[4] = z.x; mybox[5] = z.y;
            }
            for (;;) {
                while (size > 0) {
                    if (ncand > CAND_CAP - 64) { // room for the up-to-64 leaves one step can add
                        drain_candidates(ws, ncand, mybox, valid, myface, nout, ntests, mycnt, a);
                        ncand = 0;
                    }
                    // two nodes per step when there are two (independent loads); a step adds at most 64 entries and the walk
                    // is depth-first, so the stack holds at most 63 entries per level below the start
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
                if (hits == 0u) break;
                // top-level nodes that did not fit the stack in one go
                while (hits && size <= STACK_CAP - 96) {
                    const uint32_t c = (uint32_t)__ffs((int)hits) - 1u;
                    hits &= hits - 1u;
                    const float* blk = mid + c * MCB_WBLOCK_FLOATS + lane;
                    float cb[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) cb[k] = __ldg(blk + k * 32);
                    const bool h = overlap6f(gbox, cb);
                    const unsigned m = __ballot_sync(0xffffffffu, h);
                    if (h) ws.stack[size + __popc(m & lt)] = ((unsigned)S << 28) | (32u * c + lane);
                    size += __popc(m);
                    ntests += 32ull;
                }
                __syncwarp();
            }
            if (ncand) drain_candidates(ws, ncand, mybox, valid, myface, nout, ntests, mycnt, a);
            if (!a.query_is_cut && mycnt) a.src_count[myface] = mycnt;
        }
    }
    flush_out(ws, nout, a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ntests += __shfl_xor_sync(0xffffffffu, ntests, o);
        ntop += __shfl_xor_sync(0xffffffffu, ntop, o);
    }
    if (lane == 0 && (ntests | ntop)) atomicAdd(&a.counters->n_node_tests, ntests / 32ull + ntop);
}

// ---- canonical order of the pair list by COUNTING ----------------------------------------------------------------------
// Ascending (src << 32 | cut).  The traversal already counted the pairs of every source face, so the order costs two
// passes over the counts and two over the pairs instead of six radix passes over the pairs:
//   k_pair_tilesum + k_pair_offsets   exclusive scan of the per-face counts = first slot of each source face
//   k_pair_scatter                    every pair into the range of its source face (cursor by atomicAdd)
//   k_pair_segsort                    one thread per source face orders its handful of pairs by cut face
// A face with more than SEG_LIMIT pairs (one huge face over a fine mesh) would make its thread crawl: the radix sort takes
// such inputs instead.  Which case it is only the device knows (pair_seg_max), and enqueueing seven radix launches that
// return at once in every ordinary dispatch costs more than the counting order itself; so the counting kernels simply stand
// aside above the limit, and the host runs the radix sort when it next reads the counters and finds pair_seg_max above it
// (sort_pairs_fallback, called from fetch_counters) — before anybody can see the ordered list.
constexpr unsigned SEG_LIMIT = 256;
constexpr int PS_THREADS = 256;
constexpr int PS_TILE = PS_THREADS * 8;

__device__ __forceinline__ unsigned block_sum(unsigned v, unsigned* s_warp)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane_id() == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned t = 0;
#pragma unroll
    for (int i = 0; i < PS_THREADS / 32; ++i) t += s_warp[i];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(PS_THREADS) k_pair_tilesum(const unsigned* __restrict__ cnt, uint32_t nsf, unsigned* __restrict__ tile_sum)
{
    pdl_prologue();
    __shared__ unsigned s_warp[PS_THREADS / 32];
    const size_t base = (size_t)blockIdx.x * PS_TILE + threadIdx.x * 8u;
    unsigned v = 0;
    if (base + 8 <= nsf) {
        const uint4 a = *reinterpret_cast<const uint4*>(cnt + base), b = *reinterpret_cast<const uint4*>(cnt + base + 4);
        v = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    } else {
        for (size_t i = base; i < nsf; ++i) v += cnt[i];
    }
    const unsigned t = block_sum(v, s_warp);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = t;
}

__global__ void __launch_bounds__(PS_THREADS) k_pair_offsets(unsigned* __restrict__ cnt, uint32_t nsf, const unsigned* __restrict__ tile_sum,
    unsigned* __restrict__ off, result_counters_t* counters)
{
    pdl_prologue();
    __shared__ unsigned s_warp[PS_THREADS / 32];
    // everything before this tile
    unsigned before = 0;
    for (unsigned i = threadIdx.x; i < blockIdx.x; i += PS_THREADS) before += __ldg(tile_sum + i);
    before = block_sum(before, s_warp);
    // (the arrays are padded to a multiple of 8 entries: whole 16-byte accesses everywhere; entries past nsf hold zeros)
    const size_t base = (size_t)blockIdx.x * PS_TILE + threadIdx.x * 8u;
    unsigned c[8] = { 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u };
    const bool in_range = base < nsf;
    if (in_range) {
        const uint4 a = *reinterpret_cast<const uint4*>(cnt + base), b = *reinterpret_cast<const uint4*>(cnt + base + 4);
        c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
    }
    unsigned mine = 0, mx = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        mine += c[k];
        mx = c[k] > mx ? c[k] : mx;
    }
    // exclusive scan of `mine` over the block
    unsigned inc = mine;
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    unsigned wbase = 0;
#pragma unroll
    for (int i = 0; i < PS_THREADS / 32; ++i)
        if ((unsigned)i < w) wbase += s_warp[i];
    unsigned run = before + wbase + inc - mine;
    if (in_range) {
        unsigned o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            o[k] = run;
            run += c[k];
        }
        *reinterpret_cast<uint4*>(off + base) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(off + base + 4) = make_uint4(o[4], o[5], o[6], o[7]);
        if (mine) { // becomes the scatter's cursor
            *reinterpret_cast<uint4*>(cnt + base) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(cnt + base + 4) = make_uint4(0u, 0u, 0u, 0u);
        }
        if (base + 8 >= nsf) off[nsf] = run; // (entries past nsf are zero, so `run` is the total here)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned t = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = t > mx ? t : mx;
    }
    if (lane == 0 && mx) atomicMax(&counters->pair_seg_max, mx);
}

__global__ void __launch_bounds__(256) k_pair_scatter(const unsigned long long* __restrict__ pairs, unsigned long long cap,
    const unsigned* __restrict__ off, unsigned* __restrict__ cursor, unsigned long long* __restrict__ out, const result_counters_t* counters)
{
    pdl_prologue();
    if (counters->pair_seg_max > SEG_LIMIT) return;
    const unsigned long long n = counters->n_pairs < cap ? counters->n_pairs : cap;
    for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * 256) {
        const unsigned long long p = pairs[i];
        const uint32_t s = (uint32_t)(p >> 32);
        const unsigned long long pos = (unsigned long long)__ldg(off + s) + atomicAdd(cursor + s, 1u);
        if (pos < cap) out[pos] = p; // (beyond the capacity only after a pair overflow: that run is repeated anyway)
    }
}

// One block orders the segments of SS_FACES consecutive source faces.  Their pairs are contiguous in `out`: the span is
// staged in shared memory with coalesced loads, every thread orders its own face's handful of entries there (insertion
// sort: the same source face, so the cut face decides), and the span goes back coalesced.  In global memory the same sort is
// a chain of dependent round trips — tens of microseconds for a 15-entry segment, whatever the size of the input.
constexpr int SS_FACES = 128;
constexpr unsigned SS_SPAN = 4096; // pairs staged per tile (32 KB); a denser tile is ordered in place

__global__ void __launch_bounds__(SS_FACES) k_pair_segsort(unsigned long long* __restrict__ out, const unsigned* __restrict__ off,
    unsigned* __restrict__ cursor, uint32_t nsf, unsigned long long cap, const result_counters_t* counters)
{
    pdl_prologue();
    __shared__ unsigned long long sh[SS_SPAN];
    const bool counting = counters->pair_seg_max <= SEG_LIMIT && !counters->pair_overflow;
    const uint32_t tiles = (nsf + SS_FACES - 1) / SS_FACES;
    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint32_t f0 = tile * SS_FACES, f = f0 + threadIdx.x;
        const uint32_t f1 = f0 + SS_FACES < nsf ? f0 + SS_FACES : nsf;
        const unsigned lo = __ldg(off + f0), hi = __ldg(off + f1); // (written by k_pair_offsets, an earlier launch)
        if (hi == lo) continue; // no pair in this tile: its cursors are zero already
        unsigned c = 0, first = 0;
        if (f < nsf) {
            c = cursor[f];
            if (c) cursor[f] = 0u; // all zero again for the next run
            first = __ldg(off + f) - lo;
        }
        if (!counting || hi > cap) continue;
        const unsigned span = hi - lo;
        unsigned long long* seg = span <= SS_SPAN ? sh + first : out + lo + first;
        if (span <= SS_SPAN) {
            for (unsigned i = threadIdx.x; i < span; i += SS_FACES) sh[i] = out[lo + i];
            __syncthreads();
        }
        for (unsigned i = 1; i < c; ++i) {
            const unsigned long long x = seg[i];
            unsigned j = i;
            while (j > 0 && seg[j - 1] > x) {
                seg[j] = seg[j - 1];
                --j;
            }
            seg[j] = x;
        }
        if (span <= SS_SPAN) {
            __syncthreads();
            for (unsigned i = threadIdx.x; i < span; i += SS_FACES) out[lo + i] = sh[i];
            __syncthreads();
        }
    }
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
    a.pairs = res->pairs.as<unsigned long long>();
    a.cap_pairs = res->cap_pairs;
    MCB_TRY(ctx->reserve(res->pair_cnt, sizeof(unsigned) * ((size_t)src->nf + 8)));
    if (res->pair_cnt_zeroed != res->pair_cnt.p) {
        // first use of this buffer; afterwards every run leaves it all zero again (k_pair_offsets / k_pair_segsort)
        MCB_CUDA(ctx, cudaMemsetAsync(res->pair_cnt.p, 0, res->pair_cnt.cap, ctx->cur));
        res->pair_cnt_zeroed = res->pair_cnt.p;
    }
    a.src_count = res->pair_cnt.as<unsigned>();
    a.counters = res->counters.as<result_counters_t>();
    a.groups = q->groups.as<uint2>();
    a.n_groups = reinterpret_cast<const unsigned*>(q->groups.as<uint2>() + q->nf);
    a.group_box = q->group_box.as<group_box_t>();

    // a grid sized for the machine; every warp starts with a strided static share of the expected group count (about
    // nf/16: no tickets at all in the usual case), what is left continues on a ticket counter (counters->work_counter)
    const unsigned max_blocks = (unsigned)ctx->num_sms * 3u; // 3 resident blocks per SM (registers)
    const size_t est_groups = (size_t)q->nf / 10u + 1u;
    const unsigned want_blocks = div_up(div_up(est_groups, 8), WARPS_PER_BLOCK); // at least ~8 groups per warp
    const unsigned grid = want_blocks < max_blocks ? (want_blocks ? want_blocks : 1u) : max_blocks;
    const size_t per_warp = div_up(est_groups, (size_t)grid * WARPS_PER_BLOCK);
    a.static_batch = (unsigned)(per_warp < 1 ? 1 : (per_warp > 32 ? 32 : per_warp));
    MCB_LAUNCH(ctx, k_traverse, grid, TBLOCK, 0, a);
    res->have_pairs = true;
    return 0;
}

int exclusive_scan_u32(mcb200_ctx* ctx, unsigned* cnt, uint32_t n, unsigned* tile, unsigned* off, result_counters_t* counters)
{
    const unsigned tiles = div_up(n, PS_TILE);
    MCB_LAUNCH(ctx, k_pair_tilesum, tiles, PS_THREADS, 0, cnt, n, tile);
    MCB_LAUNCH(ctx, k_pair_offsets, tiles, PS_THREADS, 0, cnt, n, tile, off, counters);
    return 0;
}

int sort_pairs_reserve(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_result* res)
{
    const rsort::pass_desc pd = pair_passes(src->nf, cut->nf);
    const size_t tiles = ((size_t)src->nf + PS_TILE - 1) / PS_TILE;
    MCB_TRY(ctx->reserve(res->pair_cnt, sizeof(unsigned) * ((size_t)src->nf + 8)));
    MCB_TRY(ctx->reserve(res->pair_off, sizeof(unsigned) * ((size_t)src->nf + 8)));
    MCB_TRY(ctx->reserve(res->pair_tile, sizeof(unsigned) * (tiles + 1)));
    return rsort::reserve_scratch<unsigned long long>(ctx, res->cap_pairs, pd.npasses, false, false);
}

// canonical order of the pair set, on ctx->cur with ctx's current scratch set; res->pairs itself is left untouched
// (the narrowphase may be reading it on the other lane)
int sort_pairs(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_result* res)
{
    const rsort::pass_desc pd = pair_passes(src->nf, cut->nf);
    MCB_TRY(sort_pairs_reserve(ctx, src, cut, res));
    result_counters_t* c = res->counters.as<result_counters_t>();
    // the radix sort ends in pairs_a after an odd number of passes, in pairs_b after an even one: the counting order
    // writes to the same buffer, so the host knows where the ordered list is whichever path the device takes
    unsigned long long* dst = (pd.npasses & 1) ? res->pairs_a.as<unsigned long long>() : res->pairs_b.as<unsigned long long>();
    const uint32_t nsf = src->nf;
    const unsigned tiles = div_up(nsf, PS_TILE);
    unsigned* cnt = res->pair_cnt.as<unsigned>();
    unsigned* off = res->pair_off.as<unsigned>();
    MCB_LAUNCH(ctx, k_pair_tilesum, tiles, PS_THREADS, 0, cnt, nsf, res->pair_tile.as<unsigned>());
    MCB_LAUNCH(ctx, k_pair_offsets, tiles, PS_THREADS, 0, cnt, nsf, res->pair_tile.as<unsigned>(), off, c);
    const unsigned grid = (unsigned)ctx->num_sms * 8u;
    MCB_LAUNCH(ctx, k_pair_scatter, grid, 256, 0, res->pairs.as<unsigned long long>(), (unsigned long long)res->cap_pairs, off, cnt, dst, c);
    const unsigned sgrid = div_up(nsf, SS_FACES) < grid * 2u ? div_up(nsf, SS_FACES) : grid * 2u;
    MCB_LAUNCH(ctx, k_pair_segsort, sgrid, SS_FACES, 0, dst, off, cnt, nsf, (unsigned long long)res->cap_pairs, c);
    res->pairs_order_input = res->pairs.p;
    res->pairs_order_unchecked = true; // fetch_counters looks at pair_seg_max once
    res->pairs_sorted = dst;
    return 0;
}

// The rare case: some source face has more pairs than the counting order handles per thread.  Radix sort of the same
// input into the same buffer; called by fetch_counters right after it has read the counters (the stream is idle).
int sort_pairs_fallback(mcb200_ctx* ctx, mcb200_result* res)
{
    res->pairs_order_unchecked = false;
    if (res->h.pair_seg_max <= SEG_LIMIT || res->h.pair_overflow) return 0;
    const uint32_t nsf = res->nsf, ncf = res->nf_ps - res->nsf;
    const rsort::pass_desc pd = pair_passes(nsf, ncf);
    ctx->use_main();
    unsigned long long* out = nullptr;
    MCB_TRY((rsort::sort<unsigned long long, uint32_t, false>(ctx, static_cast<const unsigned long long*>(res->pairs_order_input),
        res->pairs_a.as<unsigned long long>(), res->pairs_b.as<unsigned long long>(), nullptr, nullptr, nullptr,
        &res->counters.as<result_counters_t>()->n_pairs, res->cap_pairs, pd, &out, nullptr)));
    if (out != res->pairs_sorted) {
        ctx->set_error("internal: the two pair orders do not end in the same buffer", __FILE__, __LINE__);
        return MCB200_ERR_INTERNAL;
    }
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
