// mcut_b200/csrc/validate.cu — input validation passes on the device-resident mesh.
//
// Replaces, for one mesh:
//   find_connected_components()  source/kernel.cpp:235-364 (called by check_input_mesh, source/preproc.cpp:505-578):
//       components of the VERTEX graph whose edges are the face edges; ids in the order a scan over the vertices
//       discovers them, i.e. the id of a component is the rank of its smallest vertex; a vertex no face uses is a
//       component of its own; fccmap[f] = component of the face, plus per-component vertex and face counts.
//   mesh_is_closed()             source/preproc.cpp:1957-1990: no edge is used by one face only.
// The reference floods the graph breadth-first on one thread.  Here: lock-free union-find (roots hooked to the SMALLER
// index, so the root of a component is its smallest vertex), a flatten pass, a prefix count over "I am a root" for the
// ids, and atomics for the counts; border edges are the slots of an edge hash table that saw one halfedge.
// Pure integer work: results equal the reference's exactly (tests/test_gpu_validate.py against oracle/ref_unit.cpp).
#include "internal.h"

namespace {

constexpr int VBLOCK = 256;

struct validate_args_t {
    const uint32_t* face_vtx;
    const uint32_t* face_off; // nullptr: triangles
    uint32_t nv, nf;
    uint32_t* label; // [nv] union-find parent, after the flatten pass: smallest vertex of the component
    uint32_t* ccid; // [nv] component id of a root
    uint32_t* bsum; // [blocks + 1]
    int32_t* fccmap; // [nf]
    int32_t* cc_vertex_count; // [nv]
    int32_t* cc_face_count; // [nv]
    unsigned long long* ekeys; // [cap] edge table: (lo << 32 | hi) + 1
    uint32_t* ecount; // [cap] halfedges seen
    uint32_t emask;
    uint32_t* info; // [0] components, [1] border edges
};

__device__ __forceinline__ uint32_t uf_find(uint32_t* label, uint32_t x)
{
    for (;;) {
        const uint32_t p = *reinterpret_cast<volatile uint32_t*>(label + x);
        if (p == x) return x;
        const uint32_t gp = *reinterpret_cast<volatile uint32_t*>(label + p);
        if (gp != p) label[x] = gp; // path halving; a stale write only lengthens a path, labels never increase
        x = p;
    }
}

__device__ __forceinline__ void uf_union(uint32_t* label, uint32_t a, uint32_t b)
{
    for (;;) {
        a = uf_find(label, a);
        b = uf_find(label, b);
        if (a == b) return;
        if (a > b) {
            const uint32_t t = a;
            a = b;
            b = t;
        }
        const uint32_t old = atomicCAS(label + b, b, a); // hook the larger root under the smaller one
        if (old == b) return;
        b = old;
    }
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x ^= x >> 31;
    x *= 0x7fb5d329728ea185ULL;
    x ^= x >> 27;
    x *= 0x81dadef4bc2dd44dULL;
    x ^= x >> 33;
    return x;
}

__global__ void __launch_bounds__(VBLOCK) k_cc_init(validate_args_t a)
{
    pdl_prologue();
    for (uint32_t v = blockIdx.x * VBLOCK + threadIdx.x; v < a.nv; v += gridDim.x * VBLOCK) {
        a.label[v] = v;
        a.cc_vertex_count[v] = 0;
        a.cc_face_count[v] = 0;
    }
}

// unions along the face boundaries + one edge-table insert per halfedge
__global__ void __launch_bounds__(VBLOCK) k_cc_union(validate_args_t a)
{
    pdl_prologue();
    for (uint32_t f = blockIdx.x * VBLOCK + threadIdx.x; f < a.nf; f += gridDim.x * VBLOCK) {
        const uint32_t h0 = a.face_off ? a.face_off[f] : 3u * f;
        const uint32_t n = a.face_off ? a.face_off[f + 1] - h0 : 3u;
        uint32_t prev = __ldg(a.face_vtx + h0 + n - 1);
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t cur = __ldg(a.face_vtx + h0 + i);
            if (i > 0) uf_union(a.label, prev, cur); // n - 1 unions connect the n vertices of the face
            const uint32_t lo = prev < cur ? prev : cur, hi = prev < cur ? cur : prev;
            const unsigned long long key = (((unsigned long long)lo << 32) | hi) + 1ull;
            uint32_t slot = (uint32_t)mix64(key) & a.emask;
            for (;;) {
                const unsigned long long seen = *reinterpret_cast<volatile unsigned long long*>(a.ekeys + slot);
                if (seen == key) break;
                if (seen == 0ull) {
                    const unsigned long long was = atomicCAS(a.ekeys + slot, 0ull, key);
                    if (was == 0ull || was == key) break;
                }
                slot = (slot + 1u) & a.emask;
            }
            atomicAdd(a.ecount + slot, 1u);
            prev = cur;
        }
    }
}

// label[v] := root; per-block count of roots
__global__ void __launch_bounds__(VBLOCK) k_cc_flatten(validate_args_t a)
{
    pdl_prologue();
    __shared__ unsigned wsum[VBLOCK / 32];
    const uint32_t v = blockIdx.x * VBLOCK + threadIdx.x;
    unsigned is_root = 0;
    if (v < a.nv) {
        const uint32_t r = uf_find(a.label, v);
        is_root = (r == v) ? 1u : 0u;
        if (!is_root) a.label[v] = r; // roots keep label[v] == v; nobody hooks any more, so this is final
    }
    unsigned s = is_root;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
#pragma unroll
        for (int w = 0; w < VBLOCK / 32; ++w) t += wsum[w];
        a.bsum[blockIdx.x] = t;
    }
}

// exclusive scan of the block counts (one block); total -> info[0]
__global__ void __launch_bounds__(1024) k_cc_scan(uint32_t* bsum, uint32_t nb, uint32_t* info)
{
    pdl_prologue();
    __shared__ uint32_t wtot[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024u) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nb ? bsum[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = wtot[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            wtot[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t wpre = (threadIdx.x >> 5) ? wtot[(threadIdx.x >> 5) - 1] : 0u;
        const uint32_t c = carry;
        if (i < nb) bsum[i] = c + wpre + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + wpre + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) info[0] = carry;
}

// roots take their ids: rank in vertex order (the reference discovers components in that order)
__global__ void __launch_bounds__(VBLOCK) k_cc_ids(validate_args_t a)
{
    pdl_prologue();
    __shared__ unsigned wtot[VBLOCK / 32];
    const uint32_t v = blockIdx.x * VBLOCK + threadIdx.x;
    const unsigned is_root = (v < a.nv && a.label[v] == v) ? 1u : 0u;
    unsigned x = is_root;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= (unsigned)o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = x;
    __syncthreads();
    unsigned id = a.bsum[blockIdx.x] + x - is_root;
    for (unsigned w = 0; w < (threadIdx.x >> 5); ++w) id += wtot[w];
    if (is_root) a.ccid[v] = id;
}

// per-component counts, face -> component map, border edges
__global__ void __launch_bounds__(VBLOCK) k_cc_counts(validate_args_t a, uint32_t ecap)
{
    pdl_prologue();
    const uint32_t stride = gridDim.x * VBLOCK;
    for (uint32_t v = blockIdx.x * VBLOCK + threadIdx.x; v < a.nv; v += stride)
        atomicAdd(a.cc_vertex_count + a.ccid[a.label[v]], 1);
    for (uint32_t f = blockIdx.x * VBLOCK + threadIdx.x; f < a.nf; f += stride) {
        const uint32_t h0 = a.face_off ? a.face_off[f] : 3u * f;
        const uint32_t c = a.ccid[a.label[__ldg(a.face_vtx + h0)]];
        a.fccmap[f] = (int32_t)c;
        atomicAdd(a.cc_face_count + c, 1);
    }
    unsigned border = 0;
    for (uint32_t s = blockIdx.x * VBLOCK + threadIdx.x; s < ecap; s += stride) border += (a.ecount[s] == 1u) ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) border += __shfl_xor_sync(0xffffffffu, border, o);
    if ((threadIdx.x & 31) == 0 && border) atomicAdd(a.info + 1, border);
}

} // namespace

int mesh_validate_run(mcb200_ctx* ctx, mcb200_mesh* m)
{
    const uint32_t nv = m->nv, nf = m->nf;
    size_t ecap = 1024;
    while (ecap < (size_t)m->nh + m->nh / 4) ecap <<= 1;
    MCB_TRY(ctx->reserve(m->cc_label, sizeof(uint32_t) * (size_t)nv));
    MCB_TRY(ctx->reserve(m->cc_id, sizeof(uint32_t) * (size_t)nv));
    MCB_TRY(ctx->reserve(m->cc_vcount, sizeof(int32_t) * (size_t)nv));
    MCB_TRY(ctx->reserve(m->cc_fcount, sizeof(int32_t) * (size_t)nv));
    MCB_TRY(ctx->reserve(m->cc_fmap, sizeof(int32_t) * (size_t)nf));
    MCB_TRY(ctx->reserve(m->cc_info, sizeof(uint32_t) * 4));
    const unsigned nb = div_up(nv, VBLOCK);
    MCB_TRY(ctx->reserve(ctx->st_bsum, sizeof(uint32_t) * ((size_t)nb + 1)));
    MCB_TRY(ctx->reserve(ctx->st_tab_keys, sizeof(unsigned long long) * ecap));
    MCB_TRY(ctx->reserve(ctx->st_hfirst, sizeof(uint32_t) * ecap));
    validate_args_t a;
    a.face_vtx = m->d_face_vtx;
    a.face_off = m->d_face_off;
    a.nv = nv;
    a.nf = nf;
    a.label = m->cc_label.as<uint32_t>();
    a.ccid = m->cc_id.as<uint32_t>();
    a.bsum = ctx->st_bsum.as<uint32_t>();
    a.fccmap = m->cc_fmap.as<int32_t>();
    a.cc_vertex_count = m->cc_vcount.as<int32_t>();
    a.cc_face_count = m->cc_fcount.as<int32_t>();
    a.ekeys = ctx->st_tab_keys.as<unsigned long long>();
    a.ecount = ctx->st_hfirst.as<uint32_t>();
    a.emask = (uint32_t)(ecap - 1);
    a.info = m->cc_info.as<uint32_t>();
    MCB_CUDA(ctx, cudaMemsetAsync(a.ekeys, 0, sizeof(unsigned long long) * ecap, ctx->cur));
    MCB_CUDA(ctx, cudaMemsetAsync(a.ecount, 0, sizeof(uint32_t) * ecap, ctx->cur));
    MCB_CUDA(ctx, cudaMemsetAsync(a.info, 0, sizeof(uint32_t) * 4, ctx->cur));
    const unsigned max_grid = (unsigned)ctx->num_sms * 8u;
    const unsigned gv = nb < max_grid ? nb : max_grid, gf = div_up(nf, VBLOCK) < max_grid ? div_up(nf, VBLOCK) : max_grid;
    MCB_LAUNCH(ctx, k_cc_init, gv, VBLOCK, 0, a);
    MCB_LAUNCH(ctx, k_cc_union, gf, VBLOCK, 0, a);
    MCB_LAUNCH(ctx, k_cc_flatten, nb, VBLOCK, 0, a);
    MCB_LAUNCH(ctx, k_cc_scan, 1, 1024, 0, a.bsum, nb, a.info);
    MCB_LAUNCH(ctx, k_cc_ids, nb, VBLOCK, 0, a);
    MCB_LAUNCH(ctx, k_cc_counts, max_grid, VBLOCK, 0, a, (uint32_t)ecap);
    m->validated = true;
    return 0;
}
