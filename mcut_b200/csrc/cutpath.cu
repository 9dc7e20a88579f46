// mcut_b200/csrc/cutpath.cu — SURVEY.md §8-f4, first half: the cut-path segment table, the first consumer of the registry.
//
// "Create edges with intersection points" (source/kernel.cpp:3332-3617) walks cutpath_edge_creation_info, a
// std::map<pair<fd_t>, std::vector<vd_t>> filled next to the registry (kernel.cpp:2603-2633): every intersection point is
// filed under {source-mesh face, cut-mesh face} for the tested face against each face incident to the tested edge.  A group
// of two points is one cut-path edge, a group of more is put in order along the common line of the two faces
// (linear_projection_sort, kernel.cpp:1496-1531) and yields an edge per consecutive couple, a group of one point is the
// reference's late general-position violation (:3366-3440).  Here the table is a sort: two keys per registry record
// (sm face << 32 | cm face), one stable radix sort (the record index is the value, so a group's points stay in registry
// order like the map's vectors), group heads by a scan, and one thread per group for the rare groups of more than two.
// The m0 half-edge bookkeeping that follows (m0.add_edge, ps_iface_to_m0_edge_list, the mid-point test of :3518-3600)
// stays the reference's host code.
#include "internal.h"
#include "radix_sort.cuh"

namespace {

constexpr unsigned long long CP_NONE = ~0ull;
constexpr unsigned CP_MAX_GROUP = 64; // points of one face pair ordered per thread; more is reported, not guessed

__global__ void __launch_bounds__(256) k_cp_emit(const mcb200_record* __restrict__ rec, uint32_t n, const uint32_t* __restrict__ edge_f,
    uint32_t nsf, unsigned long long* __restrict__ keys, unsigned long long* d_m, unsigned* info)
{
    pdl_prologue();
    if (blockIdx.x == 0 && threadIdx.x == 0) *d_m = 2ull * n;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n; i += gridDim.x * 256u) {
        const uint32_t edge = rec[i].edge, tested = rec[i].face;
        const uint2 ef = __ldg(reinterpret_cast<const uint2*>(edge_f) + edge);
        // kernel.cpp:2603-2605: the face of h0 unless h0 is a border halfedge; the second face only in the first case
        const uint32_t own = ef.x != MCB200_NULL ? ef.x : ef.y;
        const uint32_t other = own == ef.x ? ef.y : MCB200_NULL;
        unsigned long long k0 = CP_NONE, k1 = CP_NONE;
        if (own == MCB200_NULL) {
            atomicOr(info + 3, 1u); // an edge without faces cannot be in the registry
        } else {
            const bool edge_is_cut = own >= nsf; // key = {source-mesh face, cut-mesh face} (:2618-2633)
            k0 = edge_is_cut ? ((unsigned long long)tested << 32 | own) : ((unsigned long long)own << 32 | tested);
            if (other != MCB200_NULL) k1 = edge_is_cut ? ((unsigned long long)tested << 32 | other) : ((unsigned long long)other << 32 | tested);
        }
        keys[2 * (size_t)i] = k0;
        keys[2 * (size_t)i + 1] = k1;
    }
}

__device__ __forceinline__ bool cp_is_head(const unsigned long long* keys, uint32_t i)
{
    const unsigned long long k = keys[i];
    return k != CP_NONE && (i == 0 || keys[i - 1] != k);
}

__global__ void __launch_bounds__(256) k_cp_heads(const unsigned long long* __restrict__ keys, uint32_t m, unsigned* __restrict__ head)
{
    pdl_prologue();
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < m; i += gridDim.x * 256u) head[i] = cp_is_head(keys, i) ? 1u : 0u;
}

// group g starts at the g-th head; the entries behind the last valid key are the absent second faces (CP_NONE sorts last)
__global__ void __launch_bounds__(256) k_cp_groups(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ idx, uint32_t m,
    const unsigned* __restrict__ head_rank /* exclusive scan of the head flags; [m] = number of groups */,
    unsigned long long* __restrict__ seg_key, uint32_t* __restrict__ seg_off, uint32_t* __restrict__ seg_vtx, unsigned* info)
{
    pdl_prologue();
    const unsigned n_groups = head_rank[m];
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < m; i += gridDim.x * 256u) {
        const unsigned long long k = keys[i];
        seg_vtx[i] = idx[i] >> 1; // entry 2v / 2v+1 belongs to registry record v
        if (cp_is_head(keys, i)) {
            seg_key[head_rank[i]] = k;
            seg_off[head_rank[i]] = i;
        }
        if (k != CP_NONE && (i + 1 == m || keys[i + 1] == CP_NONE)) { // last valid entry
            seg_off[n_groups] = i + 1;
            info[1] = i + 1;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        info[0] = n_groups;
        if (n_groups == 0) {
            seg_off[0] = 0;
            info[1] = 0;
        }
    }
}

// linear_projection_sort (kernel.cpp:1496-1531) for the groups of more than two points; arithmetic of math.h:635-642
// (dot product accumulated from 0.0 in x, y, z order) and :718-721 (normalize = v / sqrt(dot(v, v)))
__global__ void __launch_bounds__(128) k_cp_order(const mcb200_record* __restrict__ rec, const uint32_t* __restrict__ seg_off,
    uint32_t* __restrict__ seg_vtx, unsigned* info)
{
    pdl_prologue();
    const unsigned n_groups = info[0];
    unsigned single = 0;
    for (uint32_t g = blockIdx.x * 128u + threadIdx.x; g < n_groups; g += gridDim.x * 128u) {
        const uint32_t lo = seg_off[g], c = seg_off[g + 1] - lo;
        if (c == 1) single++;
        if (c <= 2) continue;
        if (c > CP_MAX_GROUP) {
            atomicOr(info + 3, 2u);
            continue;
        }
        uint32_t* v = seg_vtx + lo;
        const double* o = rec[v[0]].point;
        const double* d = rec[v[1]].point;
        double dir[3], len2 = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) dir[k] = __dsub_rn(o[k], d[k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) len2 = __dadd_rn(len2, __dmul_rn(dir[k], dir[k]));
        const double len = __dsqrt_rn(len2);
#pragma unroll
        for (int k = 0; k < 3; ++k) dir[k] = __ddiv_rn(dir[k], len);
        double proj[CP_MAX_GROUP];
        for (uint32_t i = 0; i < c; ++i) {
            const double* p = rec[v[i]].point;
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) acc = __dadd_rn(acc, __dmul_rn(__dsub_rn(o[k], p[k]), dir[k]));
            proj[i] = acc;
        }
        for (uint32_t i = 1; i < c; ++i) { // std::sort of a handful of elements: a stable insertion sort, ascending
            const double x = proj[i];
            const uint32_t xv = v[i];
            uint32_t j = i;
            while (j > 0 && x < proj[j - 1]) {
                proj[j] = proj[j - 1];
                v[j] = v[j - 1];
                --j;
            }
            proj[j] = x;
            v[j] = xv;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) single += __shfl_xor_sync(0xffffffffu, single, o);
    if (lane_id() == 0 && single) atomicAdd(info + 2, single);
}

#define MCB_FAIL(ctx, code, msg)                        \
    do {                                                \
        (ctx)->set_error((msg), __FILE__, __LINE__);    \
        return (code);                                  \
    } while (0)

int bits_for_cp(uint32_t n)
{
    int b = 1;
    while (b < 32 && (1ull << b) < (unsigned long long)n) ++b;
    return b;
}

} // namespace

extern "C" {

int mcb200_cutpath_segments(mcb200_ctx* ctx, const mcb200_soup* soup, mcb200_result* res, mcb200_cutpath_counts* out)
{
    if (!ctx || !soup || !res || !out) return MCB200_ERR_INVALID;
    MCB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!res->have_narrow) MCB_FAIL(ctx, MCB200_ERR_INVALID, "cutpath_segments: run the narrowphase first");
    MCB_TRY(fetch_counters(ctx, res));
    ctx->use_main();
    MCB_TRY(narrowphase_sort_records(ctx, res)); // (no-op when the stage already ordered the registry)
    const uint32_t n = (uint32_t)(res->h.n_records < res->cap_records ? res->h.n_records : res->cap_records);
    const uint32_t m = 2u * n;
    std::memset(out, 0, sizeof(*out));
    res->cp_groups = res->cp_entries = 0;
    res->cp_valid = true;
    if (n == 0) return 0;
    const uint32_t nf = soup->nsf + soup->ncf;
    const size_t pad = ((size_t)m + 8 + 7) & ~(size_t)7;
    MCB_TRY(ctx->reserve(res->cp_keys, sizeof(unsigned long long) * m));
    MCB_TRY(ctx->reserve(res->cp_idx, sizeof(uint32_t) * m));
    MCB_TRY(ctx->reserve(res->cp_head, sizeof(unsigned) * pad));
    MCB_TRY(ctx->reserve(res->cp_rank, sizeof(unsigned) * pad));
    MCB_TRY(ctx->reserve(res->cp_tile, sizeof(unsigned) * (pad / 8 + 8)));
    MCB_TRY(ctx->reserve(res->cp_seg_key, sizeof(unsigned long long) * m));
    MCB_TRY(ctx->reserve(res->cp_seg_off, sizeof(uint32_t) * ((size_t)m + 1)));
    MCB_TRY(ctx->reserve(res->cp_seg_vtx, sizeof(uint32_t) * m));
    MCB_TRY(ctx->reserve(res->cp_info, sizeof(unsigned long long) * 4));
    // keys: bits(nf) + 1 low bits and bits(nsf) + 1 high bits, so that CP_NONE (all ones) sorts behind every real key
    const rsort::pass_desc pd = rsort::make_passes(0, bits_for_cp(nf) + 1, 32, 32 + bits_for_cp(soup->nsf) + 1);
    MCB_TRY((rsort::reserve_scratch<unsigned long long>(ctx, m, pd.npasses, true, true)));
    mcb200_ctx::sort_scratch_t& sc = ctx->sc();
    unsigned* info = res->cp_info.as<unsigned>() + 2; // [0..1] = the entry count as a 64-bit word, then four 32-bit words
    unsigned long long* d_m = res->cp_info.as<unsigned long long>();
    MCB_CUDA(ctx, cudaMemsetAsync(res->cp_info.p, 0, sizeof(unsigned long long) * 4, ctx->cur));
    MCB_CUDA(ctx, cudaMemsetAsync(res->cp_head.p, 0, sizeof(unsigned) * pad, ctx->cur));
    const mcb200_record* rec = res->records_sorted.as<mcb200_record>();
    const unsigned grid = (unsigned)ctx->num_sms * 4u;
    const unsigned g_n = div_up(n, 256) < grid ? div_up(n, 256) : grid, g_m = div_up(m, 256) < grid ? div_up(m, 256) : grid;
    MCB_LAUNCH(ctx, k_cp_emit, g_n, 256, 0, rec, n, soup->edge_f.as<uint32_t>(), soup->nsf, res->cp_keys.as<unsigned long long>(), d_m, info);
    unsigned long long* kout = nullptr;
    uint32_t* vout = nullptr;
    MCB_TRY((rsort::sort<unsigned long long, uint32_t, true>(ctx, res->cp_keys.as<unsigned long long>(), sc.keys_alt.as<unsigned long long>(),
        res->cp_keys.as<unsigned long long>(), nullptr, sc.vals_alt.as<uint32_t>(), res->cp_idx.as<uint32_t>(), d_m, m, pd, &kout, &vout)));
    MCB_LAUNCH(ctx, k_cp_heads, g_m, 256, 0, kout, m, res->cp_head.as<unsigned>());
    MCB_TRY(exclusive_scan_u32(ctx, res->cp_head.as<unsigned>(), m, res->cp_tile.as<unsigned>(), res->cp_rank.as<unsigned>(),
        res->counters.as<result_counters_t>()));
    MCB_LAUNCH(ctx, k_cp_groups, g_m, 256, 0, kout, vout, m, res->cp_rank.as<unsigned>(), res->cp_seg_key.as<unsigned long long>(),
        res->cp_seg_off.as<uint32_t>(), res->cp_seg_vtx.as<uint32_t>(), info);
    MCB_LAUNCH(ctx, k_cp_order, g_m, 128, 0, rec, res->cp_seg_off.as<uint32_t>(), res->cp_seg_vtx.as<uint32_t>(), info);
    unsigned h[4];
    MCB_CUDA(ctx, cudaMemcpyAsync(h, info, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h[3] & 1u) MCB_FAIL(ctx, MCB200_ERR_INVALID, "cutpath_segments: a registry record names an edge without faces");
    if (h[3] & 2u) MCB_FAIL(ctx, MCB200_ERR_CAPACITY, "cutpath_segments: more than 64 intersection points on one face pair");
    res->cp_groups = h[0];
    res->cp_entries = h[1];
    out->n_groups = h[0];
    out->n_entries = h[1];
    out->n_single_point_groups = h[2];
    return 0;
}

int mcb200_cutpath_read(mcb200_ctx* ctx, mcb200_result* res, uint64_t* keys, uint32_t* offsets, uint32_t* vertices, size_t cap_groups,
    size_t cap_entries)
{
    if (!ctx || !res) return MCB200_ERR_INVALID;
    MCB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!res->cp_valid) MCB_FAIL(ctx, MCB200_ERR_INVALID, "cutpath_read: run mcb200_cutpath_segments first");
    if (cap_groups < res->cp_groups || cap_entries < res->cp_entries) MCB_FAIL(ctx, MCB200_ERR_CAPACITY, "cutpath_read: caller's arrays are too small");
    if (offsets && res->cp_groups == 0) offsets[0] = 0;
    if (res->cp_groups == 0) return 0;
    if (keys) MCB_CUDA(ctx, cudaMemcpyAsync(keys, res->cp_seg_key.p, sizeof(uint64_t) * res->cp_groups, cudaMemcpyDeviceToHost, ctx->stream));
    if (offsets) MCB_CUDA(ctx, cudaMemcpyAsync(offsets, res->cp_seg_off.p, sizeof(uint32_t) * (res->cp_groups + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if (vertices) MCB_CUDA(ctx, cudaMemcpyAsync(vertices, res->cp_seg_vtx.p, sizeof(uint32_t) * res->cp_entries, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

} // extern "C"
