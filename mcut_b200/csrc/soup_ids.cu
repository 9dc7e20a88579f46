// mcut_b200/csrc/soup_ids.cu — polygon-soup numbering on the device.
//
// The reference's `ps` is the source half-edge mesh with the cut mesh's faces appended through add_face()
// (source/kernel.cpp:1593-1732).  Its ids are sequential by construction: an edge gets the next id the first time its
// unordered vertex pair is met while faces are walked in order (source/hmesh.cpp:406-651), h0 of the edge is the halfedge
// of that first face and h1 belongs to the second face, and get_vertices_around_face returns halfedge targets
// (hmesh.cpp:705-733).  "First time in face order" has a parallel form: number the halfedges in the order add_face meets
// them (h = face offset + slot); an edge's id is the RANK of its smallest halfedge among all edges' smallest halfedges.
// So: (1) every halfedge inserts its vertex pair into an open-addressing table and registers itself with the slot,
// (2) each halfedge learns the slot's smallest user, (3) a prefix count over "I am the smallest user" in halfedge order
// hands out the ids, (4) the other user copies the id.
// The table is built for locality, not for uniformity: a slot is 16 bytes {key, ~min user, max user} so one 32-byte sector
// serves the probe, the claim and both registrations, and an edge hashes into the small region of its LOWER vertex, so
// faces that are neighbours in the index arrays (every mesh that came out of a mesher) touch neighbouring lines.
// Pure integer work, bit-identical to mcb200_soup_ids (host_logic.cpp) and checked against it in tests/test_gpu_parity.py.
#include "internal.h"

namespace {

constexpr int SBLOCK = 256;

struct __align__(16) edge_slot_t {
    unsigned long long key; // (lo vertex << 32 | hi vertex) + 1, 0 = empty
    uint32_t nfirst; // ~(smallest halfedge registered); 0 = none yet
    uint32_t last; // largest halfedge registered
};

struct soup_args_t {
    const uint32_t* src_vtx;
    const uint32_t* src_off; // nullptr: triangles
    const uint32_t* cut_vtx;
    const uint32_t* cut_off;
    uint32_t nsf, ncf, nsv, src_nh, nh;
    edge_slot_t* tab; // [cap], zeroed
    uint32_t mask, region; // slots per vertex region
    uint32_t* hfirst; // [nh] pass 1: slot of the halfedge; pass 2 on: smallest halfedge of its edge
    uint32_t* bsum; // [nblocks + 1]
    uint32_t* face_vtx;
    uint32_t* face_edge;
    uint32_t* face_off; // nullptr when both meshes are triangle meshes
    uint32_t* edge_f;
    result_counters_t* counters;
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x ^= x >> 31;
    x *= 0x7fb5d329728ea185ULL;
    x ^= x >> 27;
    x *= 0x81dadef4bc2dd44dULL;
    x ^= x >> 33;
    return x;
}

// where face f of the soup starts in halfedge order, and how long it is
__device__ __forceinline__ void face_span(const soup_args_t& a, uint32_t f, uint32_t& h0, uint32_t& n)
{
    if (f < a.nsf) {
        const uint32_t b = a.src_off ? a.src_off[f] : 3u * f;
        n = a.src_off ? a.src_off[f + 1] - b : 3u;
        h0 = b;
    } else {
        const uint32_t fl = f - a.nsf;
        const uint32_t b = a.cut_off ? a.cut_off[fl] : 3u * fl;
        n = a.cut_off ? a.cut_off[fl + 1] - b : 3u;
        h0 = a.src_nh + b;
    }
}

// (1) insert + register.  Slot i of a face is the halfedge from user[(i + pre) % n] to user[(i + pre + 1) % n], pre = 1 for
// cut faces (kernel.cpp:1678 hands add_face a list that is already rotated once).
__global__ void __launch_bounds__(SBLOCK) k_soup_insert(soup_args_t a)
{
    pdl_prologue();
    const uint32_t nf = a.nsf + a.ncf;
    for (uint32_t f = blockIdx.x * SBLOCK + threadIdx.x; f < nf; f += gridDim.x * SBLOCK) {
        uint32_t h0, n;
        face_span(a, f, h0, n);
        const bool cutf = f >= a.nsf;
        const uint32_t* list = cutf ? a.cut_vtx + (h0 - a.src_nh) : a.src_vtx + h0;
        const uint32_t vbase = cutf ? a.nsv : 0u;
        const uint32_t pre = cutf ? 1u : 0u;
        uint32_t from = __ldg(list + pre % n) + vbase;
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t to = __ldg(list + (i + pre + 1) % n) + vbase;
            const uint32_t lo = from < to ? from : to, hi = from < to ? to : from;
            const unsigned long long key = (((unsigned long long)lo << 32) | hi) + 1ull;
            uint32_t slot = (lo * a.region + (uint32_t)(mix64(hi) % a.region)) & a.mask;
            for (;;) {
                const unsigned long long seen = *reinterpret_cast<volatile unsigned long long*>(&a.tab[slot].key);
                if (seen == key) break;
                if (seen == 0ull) {
                    const unsigned long long prev = atomicCAS(&a.tab[slot].key, 0ull, key);
                    if (prev == 0ull || prev == key) break;
                }
                slot = (slot + 1u) & a.mask;
            }
            atomicMax(&a.tab[slot].nfirst, ~(h0 + i));
            atomicMax(&a.tab[slot].last, h0 + i);
            a.hfirst[h0 + i] = slot;
            a.face_vtx[h0 + i] = to;
            from = to;
        }
        if (a.face_off) {
            a.face_off[f] = h0;
            if (f == nf - 1) a.face_off[nf] = h0 + n;
        }
    }
}

// (2) slot -> smallest user, per-block count of edge owners.  Two users running the same way = inconsistent winding.
__global__ void __launch_bounds__(SBLOCK) k_soup_first(soup_args_t a)
{
    pdl_prologue();
    __shared__ uint32_t wsum[SBLOCK / 32];
    const uint32_t nf = a.nsf + a.ncf;
    const uint32_t f = blockIdx.x * SBLOCK + threadIdx.x;
    uint32_t owned = 0;
    if (f < nf) {
        uint32_t h0, n;
        face_span(a, f, h0, n);
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t h = h0 + i;
            const uint32_t slot = a.hfirst[h];
            const uint2 u = *reinterpret_cast<const uint2*>(&a.tab[slot].nfirst);
            const uint32_t first = ~u.x, last = u.y;
            if (first != last) {
                const uint32_t other = (first == h) ? last : first;
                // a third face on one edge, or two faces running along it the same way (hmesh.cpp:612-628)
                if ((h != first && h != last) || a.face_vtx[other] == a.face_vtx[h]) a.counters->soup_error = 1u;
            }
            a.hfirst[h] = first;
            owned += (first == h) ? 1u : 0u;
        }
    }
    uint32_t s = owned;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < SBLOCK / 32; ++w) t += wsum[w];
        a.bsum[blockIdx.x] = t;
    }
}

// (3a) exclusive scan of the block counts, one block
__global__ void __launch_bounds__(1024) k_soup_scan(uint32_t* bsum, uint32_t nb, result_counters_t* counters)
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
            wtot[threadIdx.x] = w; // inclusive warp totals
        }
        __syncthreads();
        const uint32_t wpre = (threadIdx.x >> 5) ? wtot[(threadIdx.x >> 5) - 1] : 0u;
        const uint32_t c = carry;
        if (i < nb) bsum[i] = c + wpre + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + wpre + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        bsum[nb] = carry;
        counters->soup_ne = carry;
    }
}

// (3b) owners take their ids: rank in halfedge order
__global__ void __launch_bounds__(SBLOCK) k_soup_assign(soup_args_t a)
{
    pdl_prologue();
    __shared__ uint32_t wtot[SBLOCK / 32];
    const uint32_t nf = a.nsf + a.ncf;
    const uint32_t f = blockIdx.x * SBLOCK + threadIdx.x;
    uint32_t h0 = 0, n = 0, owned = 0;
    if (f < nf) {
        face_span(a, f, h0, n);
        for (uint32_t i = 0; i < n; ++i) owned += (a.hfirst[h0 + i] == h0 + i) ? 1u : 0u;
    }
    uint32_t x = owned;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t e = a.bsum[blockIdx.x] + x - owned;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) e += wtot[w];
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t h = h0 + i;
        if (a.hfirst[h] == h) {
            a.face_edge[h] = e;
            a.edge_f[2u * e] = f;
            a.edge_f[2u * e + 1u] = MCB200_NULL; // overwritten below when a second face uses the edge
            ++e;
        }
    }
}

// (4) the second user copies the id and signs in as the face of h1
__global__ void __launch_bounds__(SBLOCK) k_soup_twin(soup_args_t a)
{
    pdl_prologue();
    const uint32_t nf = a.nsf + a.ncf;
    const uint32_t f = blockIdx.x * SBLOCK + threadIdx.x;
    if (f >= nf) return;
    uint32_t h0, n;
    face_span(a, f, h0, n);
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t h = h0 + i;
        const uint32_t first = a.hfirst[h];
        if (first != h) {
            const uint32_t e = a.face_edge[first];
            a.face_edge[h] = e;
            a.edge_f[2u * e + 1u] = f;
        }
    }
}

} // namespace

int soup_number_reserve(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_soup* soup)
{
    const uint32_t nh = src->nh + cut->nh;
    const uint32_t nf = src->nf + cut->nf;
    size_t cap = 1024;
    while (cap < (size_t)nh) cap <<= 1; // distinct edges <= nh, = nh / 2 for closed meshes: load factor <= 1/2 .. 1
    if (cap < (size_t)nh + nh / 4) cap <<= 1;
    MCB_TRY(ctx->reserve(ctx->st_tab_keys, sizeof(edge_slot_t) * cap));
    MCB_TRY(ctx->reserve(ctx->st_hfirst, sizeof(uint32_t) * (size_t)nh));
    MCB_TRY(ctx->reserve(ctx->st_bsum, sizeof(uint32_t) * ((size_t)div_up(nf, SBLOCK) + 1)));
    ctx->st_tab_cap = cap;
    soup->nsf = src->nf;
    soup->ncf = cut->nf;
    soup->nh = nh;
    soup->ne = nh; // bound; the count lands in the result counters (soup_ne)
    soup->all_tri = (src->is_tri && cut->is_tri) ? 1 : 0;
    MCB_TRY(ctx->reserve(soup->face_vtx, sizeof(uint32_t) * (size_t)nh));
    MCB_TRY(ctx->reserve(soup->face_edge, sizeof(uint32_t) * (size_t)nh));
    MCB_TRY(ctx->reserve(soup->edge_f, sizeof(uint32_t) * 2 * (size_t)nh));
    if (!soup->all_tri) MCB_TRY(ctx->reserve(soup->face_off, sizeof(uint32_t) * ((size_t)nf + 1)));
    return 0;
}

// All five launches on ctx->cur.  `counters` must already be zeroed for this run (soup_error / soup_ne are written here).
int soup_number_device(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_soup* soup, result_counters_t* counters)
{
    soup_args_t a;
    a.src_vtx = src->d_face_vtx;
    a.src_off = src->d_face_off;
    a.cut_vtx = cut->d_face_vtx;
    a.cut_off = cut->d_face_off;
    a.nsf = src->nf;
    a.ncf = cut->nf;
    a.nsv = src->nv;
    a.src_nh = src->nh;
    a.nh = src->nh + cut->nh;
    a.tab = ctx->st_tab_keys.as<edge_slot_t>();
    a.mask = (uint32_t)(ctx->st_tab_cap - 1);
    const size_t nvt = (size_t)src->nv + cut->nv;
    a.region = (uint32_t)(ctx->st_tab_cap / nvt);
    if (a.region == 0) a.region = 1;
    a.hfirst = ctx->st_hfirst.as<uint32_t>();
    a.bsum = ctx->st_bsum.as<uint32_t>();
    a.face_vtx = soup->face_vtx.as<uint32_t>();
    a.face_edge = soup->face_edge.as<uint32_t>();
    a.face_off = soup->all_tri ? nullptr : soup->face_off.as<uint32_t>();
    a.edge_f = soup->edge_f.as<uint32_t>();
    a.counters = counters;
    const uint32_t nf = a.nsf + a.ncf;
    const unsigned nb = div_up(nf, SBLOCK);
    MCB_CUDA(ctx, cudaMemsetAsync(a.tab, 0, sizeof(edge_slot_t) * ctx->st_tab_cap, ctx->cur));
    MCB_LAUNCH(ctx, k_soup_insert, nb, SBLOCK, 0, a);
    MCB_LAUNCH(ctx, k_soup_first, nb, SBLOCK, 0, a);
    MCB_LAUNCH(ctx, k_soup_scan, 1, 1024, 0, a.bsum, nb, counters);
    MCB_LAUNCH(ctx, k_soup_assign, nb, SBLOCK, 0, a);
    MCB_LAUNCH(ctx, k_soup_twin, nb, SBLOCK, 0, a);
    return 0;
}
