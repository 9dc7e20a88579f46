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

// read-only walk to the root (no path halving): used once hooking is over, when other threads finalise labels concurrently
__device__ __forceinline__ uint32_t uf_root(const uint32_t* label, uint32_t x)
{
    for (;;) {
        const uint32_t p = *reinterpret_cast<const volatile uint32_t*>(label + x);
        if (p == x) return x;
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
        // Read-only find: a halving write from a concurrent walker could otherwise land AFTER another thread's final store
        // and put a non-root ancestor back into its label.  The only stores of this kernel write a vertex's own root into
        // its own slot, which keeps every path valid (and only shortens it) for the walkers still under way.
        const uint32_t r = uf_root(a.label, v);
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

// ---------------------------------------------------------------------------------------------------------------------
// Winding number of a query point with respect to a mesh.
// Replaces getWindingNumber() / computeWindingNumberOnFace() / calculate_signed_solid_angle() (source/preproc.cpp:1650-1955),
// which check_and_store_input_mesh_intersection_type() (:1999-2122) uses to tell INSIDE from OUTSIDE when two watertight
// meshes do not intersect: winding number ~1 = inside, ~0 = outside (eps 1e-7).  One thread per face evaluates the
// reference's triangle or quad formula in the same operation order (no FMA); block sums are written to an array and added
// up by one block in a fixed order, so the result does not depend on scheduling.  atan2 is the device's (<= 2 ulp from
// libm's), so values agree with the reference to ~1e-13, classifications exactly.  Faces with more than four vertices
// go through the reference's constrained Delaunay triangulation, which stays on the host: such a mesh is refused.
// ---------------------------------------------------------------------------------------------------------------------
namespace {

struct wvec {
    double x, y, z;
};
__device__ __forceinline__ wvec w_sub(const wvec& a, const wvec& b) { return { __dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y), __dsub_rn(a.z, b.z) }; }
__device__ __forceinline__ double w_dot(const wvec& a, const wvec& b)
{
    double r = 0.0; // dot_product accumulates from 0.0 (math.h:634-642)
    r = __dadd_rn(r, __dmul_rn(a.x, b.x));
    r = __dadd_rn(r, __dmul_rn(a.y, b.y));
    r = __dadd_rn(r, __dmul_rn(a.z, b.z));
    return r;
}
__device__ __forceinline__ wvec w_cross(const wvec& a, const wvec& b)
{
    return { __dsub_rn(__dmul_rn(a.y, b.z), __dmul_rn(a.z, b.y)), __dsub_rn(__dmul_rn(a.z, b.x), __dmul_rn(a.x, b.z)),
        __dsub_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)) };
}
__device__ __forceinline__ wvec w_div(const wvec& a, double s) { return { a.x / s, a.y / s, a.z / s }; }
__device__ __forceinline__ double w_len(const wvec& a) { return sqrt(w_dot(a, a)); }

constexpr double W_PI = 3.14159265358979323846;

__device__ double solid_angle_tri(const wvec& a, const wvec& b, const wvec& c, const wvec& q) // preproc.cpp:1650-1698
{
    const wvec qa = w_sub(a, q), qb = w_sub(b, q), qc = w_sub(c, q);
    const double al = w_len(qa), bl = w_len(qb), cl = w_len(qc);
    if (al == 0.0 || bl == 0.0 || cl == 0.0) return 0.0;
    const wvec na = w_div(qa, al), nb = w_div(qb, bl), nc = w_div(qc, cl);
    const double numerator = w_dot(na, w_cross(w_sub(nb, na), w_sub(nc, na)));
    if (numerator == 0.0) return 0.0;
    const double denominator = __dadd_rn(__dadd_rn(__dadd_rn(1.0, w_dot(na, nb)), w_dot(na, nc)), w_dot(nb, nc));
    return atan2(numerator, denominator) / (2. * W_PI);
}

__device__ double solid_angle_quad(const wvec& a, const wvec& b, const wvec& c, const wvec& d, const wvec& q) // :1700-1810
{
    wvec v[4] = { w_sub(a, q), w_sub(b, q), w_sub(c, q), w_sub(d, q) };
    double len[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) len[i] = w_len(v[i]);
    if (len[0] == 0.0 || len[1] == 0.0 || len[2] == 0.0 || len[3] == 0.0) return 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = w_div(v[i], len[i]);
    const wvec diag02 = w_sub(v[2], v[0]), diag13 = w_sub(v[3], v[1]), v01 = w_sub(v[1], v[0]), v23 = w_sub(v[3], v[2]);
    const double bary0 = w_dot(v[3], w_cross(v23, diag13));
    const double bary1 = -w_dot(v[2], w_cross(v23, diag02));
    const double bary2 = -w_dot(v[1], w_cross(v01, diag13));
    const double bary3 = w_dot(v[0], w_cross(v01, diag02));
    const double dot01 = w_dot(v[0], v[1]), dot12 = w_dot(v[1], v[2]), dot23 = w_dot(v[2], v[3]), dot30 = w_dot(v[3], v[0]);
    double omega = 0.0;
    if (__dmul_rn(bary0, bary2) < __dmul_rn(bary1, bary3)) { // split 0-2
        const double dot02 = w_dot(v[0], v[2]);
        if (bary3 != 0.0) omega = atan2(bary3, __dadd_rn(__dadd_rn(__dadd_rn(1.0, dot01), dot12), dot02));
        if (bary1 != 0.0) omega = __dadd_rn(omega, atan2(bary1, __dadd_rn(__dadd_rn(__dadd_rn(1.0, dot02), dot23), dot30)));
    } else { // split 1-3
        const double dot13 = w_dot(v[1], v[3]);
        if (-bary2 != 0.0) omega = atan2(-bary2, __dadd_rn(__dadd_rn(__dadd_rn(1.0, dot01), dot13), dot30));
        if (-bary0 != 0.0) omega = __dadd_rn(omega, atan2(-bary0, __dadd_rn(__dadd_rn(__dadd_rn(1.0, dot12), dot23), dot13)));
    }
    return omega / (2. * W_PI);
}

struct winding_args_t {
    const void* xyz;
    frame_t frame;
    const uint32_t* face_vtx;
    const uint32_t* face_off;
    uint32_t nf;
    double q[3];
    double* partial; // [gridDim.x]
    unsigned* unsupported; // set when a face has more than four vertices
};

__global__ void __launch_bounds__(VBLOCK) k_winding_partial(winding_args_t a)
{
    pdl_prologue();
    __shared__ double s_sum[VBLOCK];
    const wvec q = { a.q[0], a.q[1], a.q[2] };
    double acc = 0.0;
    for (uint32_t f = blockIdx.x * VBLOCK + threadIdx.x; f < a.nf; f += gridDim.x * VBLOCK) {
        const uint32_t h0 = a.face_off ? a.face_off[f] : 3u * f;
        const uint32_t n = a.face_off ? a.face_off[f + 1] - h0 : 3u;
        if (n > 4u) {
            *a.unsupported = 1u;
            continue;
        }
        double p[4][3];
        for (uint32_t i = 0; i < n; ++i) load_vertex(a.xyz, a.frame, __ldg(a.face_vtx + h0 + i), p[i]);
        const wvec A = { p[0][0], p[0][1], p[0][2] }, B = { p[1][0], p[1][1], p[1][2] }, C = { p[2][0], p[2][1], p[2][2] };
        if (n == 3u) acc += solid_angle_tri(A, B, C, q);
        else acc += solid_angle_quad(A, B, C, wvec { p[3][0], p[3][1], p[3][2] }, q);
    }
    s_sum[threadIdx.x] = acc;
    __syncthreads();
    for (int o = VBLOCK / 2; o > 0; o >>= 1) { // fixed tree: the block's sum does not depend on scheduling
        if ((int)threadIdx.x < o) s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) a.partial[blockIdx.x] = s_sum[0];
}

__global__ void __launch_bounds__(VBLOCK) k_winding_final(const double* partial, unsigned n, double* out)
{
    pdl_prologue();
    __shared__ double s_sum[VBLOCK];
    double acc = 0.0;
    for (unsigned i = threadIdx.x; i < n; i += VBLOCK) acc += partial[i];
    s_sum[threadIdx.x] = acc;
    __syncthreads();
    for (int o = VBLOCK / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = s_sum[0];
}

__global__ void k_vertex_position(const void* xyz, frame_t fr, uint32_t v, double* out)
{
    pdl_prologue();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double p[3];
        load_vertex(xyz, fr, v, p);
        out[0] = p[0], out[1] = p[1], out[2] = p[2];
    }
}

} // namespace

// internal coordinates of vertex v (what hmesh_t::vertex(v) holds in the reference after client_input_arrays_to_hmesh)
int mesh_vertex_position(mcb200_ctx* ctx, mcb200_mesh* m, uint32_t v, double out[3])
{
    if (v >= m->nv) return MCB200_ERR_INVALID;
    const unsigned grid = (unsigned)ctx->num_sms * 4u;
    MCB_TRY(ctx->reserve(m->cc_wn, sizeof(double) * ((size_t)grid + 2 + 3)));
    double* d = m->cc_wn.as<double>() + (size_t)grid + 2;
    MCB_LAUNCH(ctx, k_vertex_position, 1, 32, 0, m->d_xyz, m->frame, v, d);
    MCB_CUDA(ctx, cudaMemcpyAsync(out, d, sizeof(double) * 3, cudaMemcpyDeviceToHost, ctx->cur));
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->cur));
    return 0;
}

// winding number of `query` (internal coordinates, like the mesh's frame produces) into mesh->cc_wn[0]; [1] = unsupported flag
int mesh_winding_run(mcb200_ctx* ctx, mcb200_mesh* m, const double query[3])
{
    const unsigned grid = (unsigned)ctx->num_sms * 4u;
    MCB_TRY(ctx->reserve(m->cc_wn, sizeof(double) * ((size_t)grid + 2 + 3)));
    winding_args_t a;
    a.xyz = m->d_xyz;
    a.frame = m->frame;
    a.face_vtx = m->d_face_vtx;
    a.face_off = m->d_face_off;
    a.nf = m->nf;
    for (int j = 0; j < 3; ++j) a.q[j] = query[j];
    a.partial = m->cc_wn.as<double>() + 2;
    a.unsupported = reinterpret_cast<unsigned*>(m->cc_wn.as<double>() + 1);
    MCB_CUDA(ctx, cudaMemsetAsync(m->cc_wn.p, 0, sizeof(double) * 2, ctx->cur));
    MCB_LAUNCH(ctx, k_winding_partial, grid, VBLOCK, 0, a);
    MCB_LAUNCH(ctx, k_winding_final, 1, VBLOCK, 0, a.partial, grid, m->cc_wn.as<double>());
    return 0;
}
