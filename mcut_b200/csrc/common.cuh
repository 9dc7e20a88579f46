// mcut_b200/csrc/common.cuh — context, device buffers and small device helpers shared by every kernel file.
// sm_100a only; compiled with -fmad=false so no a*b+c is ever contracted (the reference build has no FMA,
// SURVEY §8-c) — the only fused operations in the product are the explicit ones in predicates.cuh.
#pragma once

#include <cuda_runtime.h>

#include <cfloat>
#include <cstddef>
#include <cmath>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "../../include/mcut_b200.h"

#define MCB_NUM_SMS_DEFAULT 148

#define MCB_CUDA(ctx, expr)                                                                              \
    do {                                                                                                 \
        cudaError_t err__ = (expr);                                                                      \
        if (err__ != cudaSuccess) {                                                                      \
            (ctx)->set_error(std::string(#expr) + ": " + cudaGetErrorString(err__), __FILE__, __LINE__); \
            return (int)err__;                                                                           \
        }                                                                                                \
    } while (0)

#define MCB_TRY(expr)            \
    do {                         \
        int rc__ = (expr);       \
        if (rc__ != 0) return rc__; \
    } while (0)

// A grow-only device allocation.
struct dbuf {
    void* p = nullptr;
    size_t cap = 0;
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct mcb200_ctx {
    int device = 0;
    int num_sms = MCB_NUM_SMS_DEFAULT;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    std::string error;
    uint64_t launches = 0;
    // optional per-kernel timing (mcb200_ctx_set_profiling): an event pair around every launch
    bool profiling = false;
    struct prof_rec {
        const char* name;
        cudaEvent_t a, b;
    };
    std::vector<prof_rec> prof;
    std::vector<cudaEvent_t> prof_pool;
    cudaEvent_t prof_event()
    {
        if (!prof_pool.empty()) {
            cudaEvent_t e = prof_pool.back();
            prof_pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    // pinned staging for small D2H reads (counters)
    void* h_pinned = nullptr;
    size_t h_pinned_cap = 0;
    // Two lanes of execution: `stream` (main) and `aux`.  Independent pieces of one intersect stage — the two meshes'
    // LBVH builds, the pair sort next to the narrowphase — run side by side; `cur` is where launches currently go and
    // `sci` selects the sort scratch set that belongs to that lane.
    cudaStream_t aux = nullptr;
    cudaStream_t cur = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_fork2 = nullptr, ev_join2 = nullptr;
    int sci = 0;
    struct sort_scratch_t {
        dbuf keys_alt, vals_alt, hist, status, tilectr;
    } scratch[2];
    sort_scratch_t& sc() { return scratch[sci]; }
    // staging for mcb200_intersect_stage_host: copy stream, upload-done events, reusable device copies of the inputs
    cudaStream_t copy = nullptr;
    cudaStream_t bg = nullptr; // lowest-priority lane: work that must not take SM slots from the builds (soup numbering)
    cudaEvent_t ev_bg = nullptr, ev_np = nullptr, ev_np2 = nullptr, ev_np3 = nullptr; // narrowphase: fork to / join from the background lane
    cudaEvent_t ev_up[4] = { nullptr, nullptr, nullptr };
    struct mcb200_mesh* st_mesh[2] = { nullptr, nullptr };
    struct mcb200_soup* st_soup = nullptr;
    dbuf st_xyz[2], st_fv[2], st_fo[2];
    dbuf st_tab_keys, st_hfirst, st_bsum; // device-side polygon-soup numbering (soup_ids.cu)
    size_t st_tab_cap = 0;
    std::vector<std::function<int()>>* recording = nullptr; // MCB_LAUNCH queues here instead of launching (see the macro)
    int morton_sort_bits = 16; // the build orders the leaves by the top 16 of the 30 Morton bits: 2 radix passes (MCB200_MORTON_SORT_BITS=24 / 30: three / four)
    bool pdl = true; // programmatic dependent launch between in-stream kernels (MCB200_PDL=0 turns it off)
    bool sort_smem_opt_in[8] = { false, false, false, false, false, false, false, false }; // radix_sort.cuh: dynamic shared memory opt-in done
    bool traverse_smem_opt_in = false;
    bool two_kernel_boxes = false; // MCB200_TWO_KERNEL_BOXES=1: face boxes and Morton codes in two kernels (lbvh.cu)
    // CUDA graphs of stage bodies (api.cu)
    bool use_graphs = true;
    size_t graph_max_faces = 400000; // mcb200_intersect_stage_host: above this the pipelined-upload path is used instead
    std::vector<struct mcb200_graph*> graphs;
    void use_main() { cur = stream; sci = 0; }
    void use_aux() { cur = aux; sci = 1; }
    void use_bg() { cur = bg; sci = 0; } // kernels on this lane bring their own buffers, never the sort scratch

    void set_error(const std::string& msg, const char* file, int line)
    {
        error = msg + " (" + file + ":" + std::to_string(line) + ")";
    }
    uint64_t alloc_epoch = 0; // bumped whenever a device buffer moves: captured graphs hold raw pointers
    int reserve(dbuf& b, size_t bytes)
    {
        if (bytes <= b.cap) return 0;
        ++alloc_epoch;
        // grow geometrically so repeated dispatches of growing size do not reallocate every time
        size_t want = bytes + bytes / 4 + 256;
        if (b.p) {
            cudaError_t e = cudaFreeAsync(b.p, stream);
            if (e != cudaSuccess) {
                set_error(std::string("cudaFreeAsync: ") + cudaGetErrorString(e), __FILE__, __LINE__);
                return (int)e;
            }
            b.p = nullptr;
            b.cap = 0;
        }
        cudaError_t e = cudaMallocAsync(&b.p, want, stream);
        if (e != cudaSuccess) {
            set_error(std::string("cudaMallocAsync(") + std::to_string(want) + "): " + cudaGetErrorString(e), __FILE__, __LINE__);
            return (int)e;
        }
        b.cap = want;
        return 0;
    }
    void release(dbuf& b)
    {
        if (b.p) ++alloc_epoch;
        if (b.p) cudaFreeAsync(b.p, stream);
        b.p = nullptr;
        b.cap = 0;
    }
    int pinned(size_t bytes)
    {
        if (bytes <= h_pinned_cap) return 0;
        if (h_pinned) cudaFreeHost(h_pinned);
        h_pinned = nullptr;
        h_pinned_cap = 0;
        cudaError_t e = cudaMallocHost(&h_pinned, bytes);
        if (e != cudaSuccess) {
            set_error(std::string("cudaMallocHost: ") + cudaGetErrorString(e), __FILE__, __LINE__);
            return (int)e;
        }
        h_pinned_cap = bytes;
        return 0;
    }
};

// The frame of the internal coordinates (preproc.cpp:91-185), applied on the fly.
struct frame_t {
    double com[3];
    double shift[3];
    double pert[3];
    float fcom[3];
    float fshift[3];
    int has_frame; // 0: vertices already are internal coordinates
    int has_pert;
    int is_float;
    int pad_;
    double eps; // enlargement of the face boxes (the build's frame only)
};
static_assert(sizeof(frame_t) % 8 == 0, "frame_t is copied word by word");

struct mcb200_mesh {
    uint32_t nv = 0, nf = 0, nh = 0;
    int is_float = 0;
    int is_tri = 1;
    bool owns_arrays = true;
    const void* d_xyz = nullptr; // user-frame vertices, float3 or double3
    const uint32_t* d_face_vtx = nullptr; // [nh]
    const uint32_t* d_face_off = nullptr; // [nf+1] or nullptr for triangles
    frame_t frame;
    // The kernels read the frame from DEVICE memory (so that a captured launch sequence can be replayed for another frame):
    // d_frames[0] = the build's frame (perturbation stripped: the reference builds the cut BVH from the unperturbed mesh,
    // preproc.cpp:2676-2698; eps = the face-box enlargement), d_frames[1] = the narrowphase's frame.  dev_frames = what is there.
    dbuf d_frames;
    frame_t dev_frames[2];
    bool dev_frames_valid = false;
    void* dev_frames_ptr = nullptr;
    // host copies of the face arrays (needed by mcb200_soup_from_meshes)
    std::vector<uint32_t> h_face_vtx, h_face_off;
    // build products
    bool built = false;
    double eps = 0.0;
    dbuf face_bbox; // [nf][6] double
    dbuf prior_bbox; // boxes the next build starts from (mcb200_mesh_set_prior_face_boxes), [n_prior][6] double
    uint32_t n_prior = 0;
    dbuf root; // 6 x u64 (order-preserving encoding) + 6 double (decoded)
    dbuf codes; // [nf] u32 Morton code by face (kept for parity reads)
    dbuf sorted_codes; // [nf] u32
    dbuf sorted_faces; // [nf] u32 (leaf -> face)
    dbuf wide; // float [blocks][6][32]: the boxes of every level of the wide tree, level 0 (sorted leaves) first
    dbuf sorted_bbox; // [nf][6] double: the exact face boxes in leaf order (what the traversal's decisive test reads)
    dbuf flags; // u32 [4]: block-done ticket of k_leaves (+ spare words)
    dbuf groups; // [nf] uint2 query groups (first leaf, count) + u32 counter after them
    dbuf group_box; // [<= nf] union box of every group
    struct wide_levels_t* lv = nullptr; // host copy of the level table (set by lbvh_build)
    // input validation products (validate.cu)
    dbuf cc_label, cc_id, cc_vcount, cc_fcount, cc_fmap, cc_info, cc_wn;
    bool validated = false;
    uint32_t n_components = 0;
    bool groups_valid = false;
};

// The tree over the Morton-sorted leaves is IMPLICIT and 32-wide (the reference's OIBVH is implicit and binary,
// bvh.cpp:444-493): level 0 = the sorted leaves, node i of level l covers nodes [32 i, 32 i + 32) of level l - 1, the top level
// has at most 32 nodes.  Only boxes are stored — single precision, rounded OUTWARDS (inner levels only prune; the decisive
// test uses the exact double face boxes) — in blocks of 32 boxes laid out [6][32], so that a warp testing the 32 children of
// a node issues six fully coalesced 128-byte loads.  Unused slots hold an empty box (min = +FLT_MAX, max = -FLT_MAX).
constexpr int MCB_MAX_LEVELS = 8;
struct wide_levels_t {
    const float* boxes; // all levels, level 0 first
    uint32_t n[MCB_MAX_LEVELS]; // nodes per level; n[0] = faces
    uint32_t off[MCB_MAX_LEVELS]; // first block of level l in `boxes` (in blocks of 32 boxes = 192 floats)
    int top; // highest level (>= 1); n[top] <= 32
};
#define MCB_WBLOCK_FLOATS 192

// Union box (conservative, single precision) of a query group: a maximal subtree of at most 32 leaves of the radix tree over the
// sorted codes (lbvh.cu: k_leaves).
struct __align__(32) group_box_t {
    float box[6];
    uint32_t pad[2];
};
static_assert(sizeof(group_box_t) == 32, "group_box_t is one 32-byte sector");

// conservative single-precision copy of a double box
__device__ __forceinline__ void box_to_float(const double* b, float* f)
{
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        f[k] = __double2float_rd(b[k]);
        f[3 + k] = __double2float_ru(b[3 + k]);
    }
}
__device__ __forceinline__ bool overlap6f(const float* a, const float* b)
{
    return ((a[0] <= b[3]) & (a[3] >= b[0]) & (a[1] <= b[4]) & (a[4] >= b[1]) & (a[2] <= b[5]) & (a[5] >= b[2])) != 0;
}

struct mcb200_soup {
    uint32_t nsf = 0, ncf = 0, nh = 0, ne = 0;
    dbuf face_vtx; // [nh] ps vertex ids in ps.get_vertices_around_face order
    dbuf face_edge; // [nh]
    dbuf face_off; // [nf+1] (always materialised for the soup)
    dbuf edge_f; // [ne][2]
    int all_tri = 1;
};

// device-side counters of one result (one 128-byte block, zeroed per run)
struct result_counters_t {
    unsigned long long n_pairs;
    unsigned long long n_node_tests;
    unsigned long long n_tests;
    unsigned long long n_exact;
    unsigned long long n_records;
    unsigned long long n_cand_faces;
    unsigned long long n_log;
    unsigned int gp_violation;
    unsigned int bad_face; // min polygon-soup id of a degenerate candidate face, 0xFFFFFFFF if none
    unsigned int pair_overflow;
    unsigned int work_counter; // ticket counter of the traversal: next query group to hand out
    unsigned int soup_error; // device-side soup numbering: an edge with three faces or two faces wound the same way
    unsigned int soup_ne; // number of polygon-soup edges it found
    unsigned long long n_queue; // polygons: tests the filter kernel handed on; triangles: tests whose stage A failed
    unsigned long long n_mid; // triangles: pairs the side prefilter could not dismiss
    unsigned long long n_cross; // triangles: certified plane crossings (second half of the exact queue)
    unsigned long long n_full; // triangles: tests with inexact differences, for the general exact kernel
    unsigned int pair_seg_max; // most candidate pairs of one source face (decides how the pair list is put in order)
    unsigned int pad[1];
};
static_assert(offsetof(result_counters_t, n_queue) % 8 == 0, "n_queue is atomically incremented as a 64-bit word");

struct mcb200_result {
    dbuf counters; // result_counters_t
    dbuf pairs; // u64 [cap_pairs], in the order the traversal emitted them (what the narrowphase consumes)
    dbuf pairs_a, pairs_b; // ping-pong buffers of the pair sort
    dbuf pair_cnt, pair_off, pair_tile; // per source face: pair count / cursor, first slot in the ordered list; tile sums of the scan
    const void* pairs_order_input = nullptr; // the unordered list the last sort_pairs took (fallback: sort_pairs_fallback)
    bool pairs_order_unchecked = false; // pair_seg_max has not been looked at since the last sort_pairs
    void* pair_cnt_zeroed = nullptr; // the count array is all zero between runs (each run clears what it touched)
    unsigned long long* pairs_sorted = nullptr; // ascending (src << 32 | cut): points into pairs_a or pairs_b
    size_t cap_pairs = 0;
    bool cand_flag_fresh = false; // candidate flags already cleared for the coming narrowphase
    bool narrow_counters_fresh = false; // the narrowphase counters are still as result_reset_counters left them
    bool counters_zeroed = false; // the caller already reset the counters for this run (mcb200_intersect_stage_host)
    dbuf cand_flag; // u8 [nf_ps]
    dbuf plane; // per ps face: normal[3], d  (4 doubles) ; maxcomp in separate int array
    dbuf plane_mc; // i32 per ps face
    dbuf exact_queue; // u64 test keys needing the exact stage
    dbuf mid_queue; // u64 [cap_pairs], triangle meshes only (narrowphase.cu: k_tri_prefilter)
    // cut-path segment table (cutpath.cu)
    dbuf cp_keys, cp_idx, cp_head, cp_rank, cp_tile, cp_seg_key, cp_seg_off, cp_seg_vtx, cp_info;
    size_t cp_groups = 0, cp_entries = 0;
    bool cp_valid = false;
    size_t cap_exact = 0;
    bool record_radix_pending = false; // the registry's radix order was deferred to the next read of the counters
    bool tri_queues = false; // the last narrowphase used the triangle pipeline (split exact queue + mid queue)
    dbuf records; // mcb200_record [cap_records]
    dbuf rec_keys; // u64
    dbuf rec_idx; // u32
    dbuf records_sorted;
    size_t cap_records = 0;
    dbuf tests; // mcb200_test log
    dbuf tests_sorted;
    dbuf test_keys;
    dbuf test_idx;
    size_t cap_tests = 0;
    uint32_t shard_part = 0, shard_nparts = 1, shard_chunk = 4096;
    uint32_t nf_ps = 0, nsf = 0, ne_ps = 0;
    bool have_pairs = false, have_narrow = false, logged_tests = false;
    bool records_sorted_valid = false, tests_sorted_valid = false;
    result_counters_t h; // last host copy
    bool h_valid = false;
};

// a captured stage body (api.cu: stage_run)
struct mcb200_graph {
    std::vector<uint64_t> sig;
    uint64_t epoch = 0;
    cudaGraphExec_t exec = nullptr; // nullptr: seen once, not captured yet
    bool refused = false; // capture failed once: do not try again
    uint64_t launches = 0;
    mcb200_result res_state; // host-side bookkeeping of the result after a run of the body
    double src_eps = 0.0, cut_eps = 0.0;
};

// ------------------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// order-preserving map double <-> u64 so mesh-AABB min/max can use integer atomics
__device__ __forceinline__ unsigned long long dbl_to_ordered(double d)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ordered_to_dbl(unsigned long long u)
{
    u = (u & 0x8000000000000000ull) ? (u & 0x7FFFFFFFFFFFFFFFull) : ~u;
    return __longlong_as_double((long long)u);
}

// the reference's min/max (math.h:590-600): min(a,b) = (b < a) ? b : a ; max(a,b) = (a < b) ? b : a
__device__ __forceinline__ double ref_min(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double ref_max(double a, double b) { return (a < b) ? b : a; }

// closed-interval AABB overlap (math.h:931-941); boxes as min xyz, max xyz
__device__ __forceinline__ bool overlap6(const double* a, const double* b)
{
    // six independent compares combined without short-circuit branches (no NaNs reach here: boxes are min/max of finite input)
    return ((a[0] <= b[3]) & (a[3] >= b[0]) & (a[1] <= b[4]) & (a[4] >= b[1]) & (a[2] <= b[5]) & (a[5] >= b[2])) != 0;
}

// internal coordinate of vertex v: x' = (x - com) + shift (+ perturbation); float input does the first two
// operations in float with (float)com, (float)shift and widens afterwards (preproc.cpp:124-134, :166-176)
__device__ __forceinline__ void load_vertex(const void* __restrict__ xyz, const frame_t& fr, uint32_t v, double out[3])
{
    if (fr.is_float) {
        const float* p = reinterpret_cast<const float*>(xyz) + 3 * (size_t)v;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float x = __ldg(p + j);
            if (fr.has_frame) {
                x = __fsub_rn(x, fr.fcom[j]);
                x = __fadd_rn(x, fr.fshift[j]);
            }
            double d = (double)x;
            if (fr.has_frame) d = __dadd_rn(d, fr.has_pert ? fr.pert[j] : 0.0);
            out[j] = d;
        }
    } else {
        const double* p = reinterpret_cast<const double*>(xyz) + 3 * (size_t)v;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double x = __ldg(p + j);
            if (fr.has_frame) {
                x = __dadd_rn(__dsub_rn(x, fr.com[j]), fr.shift[j]);
                x = __dadd_rn(x, fr.has_pert ? fr.pert[j] : 0.0);
            }
            out[j] = x;
        }
    }
}

// Programmatic dependent launch: every kernel is launched with programmatic stream serialisation allowed and waits here
// until its in-stream predecessor has completed and flushed before touching memory, so the launch itself (block
// scheduling, parameter setup) overlaps the predecessor's last blocks instead of following them.  First statement of
// every kernel; kernels after a memset or an event wait simply see a dependency that is already satisfied.
// Measured on C2: step 0.629 -> 0.592 ms.  Triggering the successor EARLY (griddepcontrol.launch_dependents at the top,
// MCB_PDL_EARLY_TRIGGER) was worse (0.694 ms): the waiting blocks take SM slots from the other lane's kernels.
__device__ __forceinline__ void pdl_prologue()
{
#ifdef MCB_PDL_EARLY_TRIGGER
    asm volatile("griddepcontrol.launch_dependents;");
#endif
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Several small fills in ONE launch (a cudaMemsetAsync per counter block costs a stream operation each, and every one
// of them breaks the programmatic launch chain between the kernels around it).
struct fill_list_t {
    unsigned* p[10];
    unsigned words[10];
    unsigned value[10];
    int n;
    void add(void* ptr, size_t nwords, unsigned v)
    {
        p[n] = static_cast<unsigned*>(ptr);
        words[n] = (unsigned)nwords;
        value[n] = v;
        ++n;
    }
};
static __global__ void __launch_bounds__(256) k_fill(fill_list_t L)
{
    pdl_prologue();
    for (int e = 0; e < L.n; ++e)
        for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < L.words[e]; i += gridDim.x * 256u) L.p[e][i] = L.value[e];
}

// the frames of up to two meshes into their device slots (kernel parameters: no staging buffer, no synchronisation)
struct frame_pack_t {
    frame_t* dst[2];
    frame_t f[2][2];
    int n;
};
static __global__ void __launch_bounds__(128) k_set_frames(frame_pack_t p)
{
    pdl_prologue();
    constexpr unsigned W = sizeof(frame_t) / 4;
    for (int m = 0; m < p.n; ++m)
        if (threadIdx.x < 2 * W) reinterpret_cast<unsigned*>(p.dst[m])[threadIdx.x] = reinterpret_cast<const unsigned*>(&p.f[m][0])[threadIdx.x];
}

// block-wide copy of a frame from device memory into shared memory (first statement group of the kernels that use frames)
__device__ __forceinline__ void load_frame_shared(frame_t* s_dst, const frame_t* __restrict__ g_src)
{
    constexpr unsigned W = sizeof(frame_t) / 4;
    if (threadIdx.x < W) reinterpret_cast<unsigned*>(s_dst)[threadIdx.x] = __ldg(reinterpret_cast<const unsigned*>(g_src) + threadIdx.x);
}

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// launch accounting: every kernel launch of the product goes through this macro
#define MCB_LAUNCH(ctx, kernel, grid, block, smem, ...) MCB_LAUNCH_NAMED(ctx, #kernel, kernel, grid, block, smem, __VA_ARGS__)

// The launch is a closure: normally it runs at once; while ctx->recording is set it is queued instead, so that the
// launches of two lanes can be issued alternately (the host needs ~5 us per launch: issuing one lane's ten kernels before
// the other lane's first one would start that lane ~50 us late).
#define MCB_LAUNCH_NAMED(ctx, name, kernel, grid, block, smem, ...)              \
    do {                                                                         \
        mcb200_ctx* c__ = (ctx);                                                 \
        cudaStream_t st__ = c__->cur;                                            \
        auto fn__ = [=]() -> int {                                               \
            mcb200_ctx::prof_rec pr__ { name, nullptr, nullptr };                \
            if (c__->profiling) {                                                \
                pr__.a = c__->prof_event();                                      \
                pr__.b = c__->prof_event();                                      \
                cudaEventRecord(pr__.a, st__);                                   \
            }                                                                    \
            cudaLaunchConfig_t cfg__ = {};                                       \
            cfg__.gridDim = dim3(grid);                                          \
            cfg__.blockDim = dim3(block);                                        \
            cfg__.dynamicSmemBytes = (smem);                                     \
            cfg__.stream = st__;                                                 \
            cudaLaunchAttribute at__[1];                                         \
            at__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;     \
            at__[0].val.programmaticStreamSerializationAllowed = 1;              \
            cfg__.attrs = at__;                                                  \
            cfg__.numAttrs = c__->pdl ? 1u : 0u;                                 \
            cudaLaunchKernelEx(&cfg__, kernel, __VA_ARGS__);                     \
            if (c__->profiling) {                                                \
                cudaEventRecord(pr__.b, st__);                                   \
                c__->prof.push_back(pr__);                                       \
            }                                                                    \
            c__->launches++;                                                     \
            cudaError_t le__ = cudaPeekAtLastError();                            \
            if (le__ != cudaSuccess) {                                           \
                c__->set_error(std::string(name) + ": " + cudaGetErrorString(le__), __FILE__, __LINE__); \
                return (int)le__;                                                \
            }                                                                    \
            return 0;                                                            \
        };                                                                       \
        if (c__->recording) {                                                    \
            c__->recording->push_back(fn__);                                     \
        } else {                                                                 \
            const int rc__ = fn__();                                             \
            if (rc__) return rc__;                                               \
        }                                                                        \
    } while (0)
