// mcut_b200/csrc/api.cu — the C-ABI of include/mcut_b200.h: handle management, H2D/D2H, stage launchers.
// No computation happens on the host here except argument checking and the final ordering of the (few) candidate
// faces' plane rows; there is no CPU fallback for any device stage.
#include <algorithm>
#include <cstddef>
#include <mutex>
#include <numeric>

#include "internal.h"

namespace {
std::string g_create_error;
std::mutex g_create_mutex;

int fail(mcb200_ctx* ctx, int code, const char* msg, const char* file, int line)
{
    if (ctx) ctx->set_error(msg, file, line);
    return code;
}
#define MCB_FAIL(ctx, code, msg) return fail((ctx), (code), (msg), __FILE__, __LINE__)

} // namespace
// did a queue between the narrowphase kernels run out of room?  (triangle meshes: the exact queue is split in halves, stage-A
// failures / certified crossings; what neither settles is parked in the mid queue, one entry per pair of capacity)
bool narrow_queue_overflow(const mcb200_result* res, const result_counters_t& h)
{
    if (!res->tri_queues) return h.n_queue > res->cap_exact;
    return h.n_queue > res->cap_exact / 2 || h.n_cross > res->cap_exact / 2 || h.n_full > res->cap_pairs;
}

int fetch_counters(mcb200_ctx* ctx, mcb200_result* res)
{
    if (res->h_valid) return 0;
    if (!res->counters.p) MCB_FAIL(ctx, MCB200_ERR_INVALID, "result has not been produced yet");
    MCB_TRY(ctx->pinned(sizeof(result_counters_t)));
    MCB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, res->counters.p, sizeof(result_counters_t), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::memcpy(&res->h, ctx->h_pinned, sizeof(result_counters_t));
    res->h_valid = true;
    {
        static const bool dbg = std::getenv("MCB200_DEBUG_COUNTERS") != nullptr; // how the narrowphase queues filled up
        if (dbg)
            std::fprintf(stderr, "[mcut_b200] pairs=%llu tests=%llu exact=%llu records=%llu | mid=%llu open=%llu cross=%llu full=%llu seg_max=%u\n",
                res->h.n_pairs, res->h.n_tests, res->h.n_exact, res->h.n_records, res->h.n_mid, res->h.n_queue, res->h.n_cross, res->h.n_full,
                res->h.pair_seg_max);
    }
    if (res->pairs_order_unchecked) MCB_TRY(sort_pairs_fallback(ctx, res)); // a face with very many pairs: see traverse.cu
    if (res->record_radix_pending) MCB_TRY(narrowphase_finish_record_order(ctx, res));
    return 0;
}
namespace {

int upload(mcb200_ctx* ctx, dbuf& dst, const void* src, size_t bytes)
{
    MCB_TRY(ctx->reserve(dst, bytes ? bytes : 16));
    if (bytes) MCB_CUDA(ctx, cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
} // namespace

extern "C" {

int mcb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int mcb200_ctx_create(int device, void* stream, mcb200_ctx** out)
{
    std::lock_guard<std::mutex> lk(g_create_mutex);
    if (!out) return MCB200_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0")
            + " — mcut_b200 has no CPU fallback";
        return MCB200_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        g_create_error = "device index out of range";
        return MCB200_ERR_INVALID;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        g_create_error = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
        return (int)e;
    }
    if (prop.major != 10) {
        g_create_error = std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor)
            + "; this library carries sm_100a code only";
        return MCB200_ERR_NO_DEVICE;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return (int)e;
    }
    mcb200_ctx* ctx = new mcb200_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    if (stream) {
        ctx->stream = reinterpret_cast<cudaStream_t>(stream);
        ctx->owns_stream = false;
    } else {
        if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
            g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
            delete ctx;
            return (int)e;
        }
        ctx->owns_stream = true;
    }
    int prio_lo = 0, prio_hi = 0; // numerically: lo = least urgent, hi = most urgent
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if ((e = cudaStreamCreateWithPriority(&ctx->aux, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) {
        g_create_error = std::string("cudaStreamCreate(aux): ") + cudaGetErrorString(e);
        delete ctx;
        return (int)e;
    }
    if (const char* e = std::getenv("MCB200_PDL")) ctx->pdl = (e[0] != '0');
    if (const char* e = std::getenv("MCB200_GRAPHS")) ctx->use_graphs = (e[0] != '0');
    if (const char* e = std::getenv("MCB200_TWO_KERNEL_BOXES")) ctx->two_kernel_boxes = (e[0] != '0');
    if (const char* e = std::getenv("MCB200_GRAPH_MAX_FACES")) ctx->graph_max_faces = (size_t)std::atoll(e);
    if (const char* e = std::getenv("MCB200_MORTON_SORT_BITS")) ctx->morton_sort_bits = (std::atoi(e) >= 30) ? 30 : (std::atoi(e) <= 16 ? 16 : 24);
    {
        // keep freed blocks in the stream-ordered pool instead of handing them back to the driver at every synchronisation:
        // a dispatch allocates a few hundred MB of build products, and mapping that memory anew costs far more than the stage
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking);
    cudaStreamCreateWithPriority(&ctx->bg, cudaStreamNonBlocking, prio_lo);
    cudaEventCreateWithFlags(&ctx->ev_bg, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_np, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_np2, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_np3, cudaEventDisableTiming);
    for (auto& ev : ctx->ev_up) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_fork2, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_join2, cudaEventDisableTiming);
    ctx->use_main();
    // keep freed blocks in the pool: repeated dispatches reuse them without going back to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = ctx;
    return 0;
}

void mcb200_ctx_destroy(mcb200_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->aux);
    cudaStreamSynchronize(ctx->copy);
    cudaStreamSynchronize(ctx->bg);
    cudaStreamDestroy(ctx->bg);
    cudaEventDestroy(ctx->ev_bg);
    cudaEventDestroy(ctx->ev_np);
    cudaEventDestroy(ctx->ev_np2);
    cudaEventDestroy(ctx->ev_np3);
    for (int k = 0; k < 2; ++k) {
        if (ctx->st_mesh[k]) mcb200_mesh_free(ctx, ctx->st_mesh[k]);
        ctx->release(ctx->st_xyz[k]);
        ctx->release(ctx->st_fv[k]);
        ctx->release(ctx->st_fo[k]);
    }
    if (ctx->st_soup) mcb200_soup_free(ctx, ctx->st_soup);
    for (dbuf* b : { &ctx->st_tab_keys, &ctx->st_hfirst, &ctx->st_bsum }) ctx->release(*b);
    cudaStreamDestroy(ctx->copy);
    for (auto& ev : ctx->ev_up) cudaEventDestroy(ev);
    for (auto& sc : ctx->scratch) {
        ctx->release(sc.keys_alt);
        ctx->release(sc.vals_alt);
        ctx->release(sc.hist);
        ctx->release(sc.status);
        ctx->release(sc.tilectr);
    }
    cudaStreamSynchronize(ctx->stream);
    cudaStreamDestroy(ctx->aux);
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_join);
    cudaEventDestroy(ctx->ev_fork2);
    cudaEventDestroy(ctx->ev_join2);
    for (mcb200_graph* g : ctx->graphs) {
        if (g->exec) cudaGraphExecDestroy(g->exec);
        delete g;
    }
    ctx->graphs.clear();
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    for (auto& r : ctx->prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* mcb200_last_error(const mcb200_ctx* ctx)
{
    if (!ctx) return g_create_error.c_str();
    return ctx->error.c_str();
}

int mcb200_ctx_sync(mcb200_ctx* ctx)
{
    if (!ctx) return MCB200_ERR_INVALID;
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

uint64_t mcb200_ctx_launch_count(const mcb200_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mcb200_ctx_set_profiling(mcb200_ctx* ctx, int on)
{
    if (!ctx) return MCB200_ERR_INVALID;
    ctx->profiling = on != 0;
    return 0;
}

int mcb200_ctx_profile_read(mcb200_ctx* ctx, char* buf, size_t capacity)
{
    if (!ctx || !buf || capacity == 0) return MCB200_ERR_INVALID;
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<std::string> names;
    std::vector<double> total;
    std::vector<unsigned> count;
    for (const auto& r : ctx->prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        size_t k = 0;
        for (; k < names.size(); ++k)
            if (names[k] == r.name) break;
        if (k == names.size()) {
            names.push_back(r.name);
            total.push_back(0.0);
            count.push_back(0);
        }
        total[k] += ms;
        count[k] += 1;
        ctx->prof_pool.push_back(r.a);
        ctx->prof_pool.push_back(r.b);
    }
    ctx->prof.clear();
    std::string out;
    for (size_t k = 0; k < names.size(); ++k) {
        std::string nm = names[k];
        for (char& c : nm)
            if (c == ' ') c = '_';
        char line[512];
        std::snprintf(line, sizeof(line), "%s %u %.6f\n", nm.c_str(), count[k], total[k]);
        out += line;
    }
    if (out.size() + 1 > capacity) MCB_FAIL(ctx, MCB200_ERR_CAPACITY, "profile_read: buffer too small");
    std::memcpy(buf, out.c_str(), out.size() + 1);
    return (int)out.size();
}

// ---------------------------------------------------------------------------------------------------------- meshes

static void identity_frame(frame_t& fr, int is_float)
{
    std::memset(&fr, 0, sizeof(fr));
    fr.is_float = is_float;
}

static int mesh_create_impl(mcb200_ctx* ctx, int is_float, const void* xyz, uint32_t nv, const uint32_t* face_vtx,
    const uint32_t* face_sizes, uint32_t nf, mcb200_mesh** out, bool trusted);

int mcb200_mesh_create(mcb200_ctx* ctx, int is_float, const void* xyz, uint32_t nv, const uint32_t* face_vtx,
    const uint32_t* face_sizes, uint32_t nf, mcb200_mesh** out)
{
    return mesh_create_impl(ctx, is_float, xyz, nv, face_vtx, face_sizes, nf, out, false);
}

int mcb200_mesh_create_trusted(mcb200_ctx* ctx, int is_float, const void* xyz, uint32_t nv, const uint32_t* face_vtx,
    const uint32_t* face_sizes, uint32_t nf, mcb200_mesh** out)
{
    return mesh_create_impl(ctx, is_float, xyz, nv, face_vtx, face_sizes, nf, out, true);
}

static int mesh_create_impl(mcb200_ctx* ctx, int is_float, const void* xyz, uint32_t nv, const uint32_t* face_vtx,
    const uint32_t* face_sizes, uint32_t nf, mcb200_mesh** out, bool trusted)
{
    if (!ctx || !out) return MCB200_ERR_INVALID;
    *out = nullptr;
    if (!xyz || !face_vtx || nv == 0 || nf == 0) MCB_FAIL(ctx, MCB200_ERR_INVALID, "mesh_create: empty mesh or NULL array");
    MCB_CUDA(ctx, cudaSetDevice(ctx->device));
    mcb200_mesh* m = new mcb200_mesh();
    m->nv = nv;
    m->nf = nf;
    m->is_float = is_float ? 1 : 0;
    m->is_tri = 1;
    if (face_sizes || !trusted) m->h_face_off.resize((size_t)nf + 1);
    if (face_sizes) {
        uint32_t acc = 0;
        for (uint32_t f = 0; f < nf; ++f) {
            if (face_sizes[f] < 3) {
                delete m;
                MCB_FAIL(ctx, MCB200_ERR_INVALID, "mesh_create: a face has fewer than 3 vertices");
            }
            if (face_sizes[f] != 3) m->is_tri = 0;
            m->h_face_off[f] = acc;
            acc += face_sizes[f];
        }
        m->h_face_off[nf] = acc;
    } else if (!trusted) {
        for (uint32_t f = 0; f <= nf; ++f) m->h_face_off[f] = 3u * f;
    }
    m->nh = face_sizes ? m->h_face_off[nf] : 3u * nf;
    if (!trusted) {
        // the host copy of the faces serves mcb200_soup_from_meshes; a caller that has validated its indices already and numbers
        // the soup on the device (the adapter inside a live mcDispatch) skips both passes over the face array
        m->h_face_vtx.assign(face_vtx, face_vtx + m->nh);
        for (uint32_t h = 0; h < m->nh; ++h)
            if (face_vtx[h] >= nv) {
                delete m;
                MCB_FAIL(ctx, MCB200_ERR_INVALID, "mesh_create: face index out of range");
            }
    } else if (m->is_tri) {
        m->h_face_off.clear();
        m->h_face_off.shrink_to_fit();
    }
    identity_frame(m->frame, m->is_float);
    const size_t vbytes = (size_t)nv * 3 * (is_float ? sizeof(float) : sizeof(double));
    void* d_xyz = nullptr;
    uint32_t* d_fv = nullptr;
    uint32_t* d_fo = nullptr;
    cudaError_t e = cudaMallocAsync(&d_xyz, vbytes, ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&d_fv, sizeof(uint32_t) * (size_t)m->nh, ctx->stream);
    if (e == cudaSuccess && !m->is_tri) e = cudaMallocAsync((void**)&d_fo, sizeof(uint32_t) * ((size_t)nf + 1), ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_xyz, xyz, vbytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_fv, face_vtx, sizeof(uint32_t) * (size_t)m->nh, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && d_fo)
        e = cudaMemcpyAsync(d_fo, m->h_face_off.data(), sizeof(uint32_t) * ((size_t)nf + 1), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) {
        if (d_xyz) cudaFreeAsync(d_xyz, ctx->stream);
        if (d_fv) cudaFreeAsync(d_fv, ctx->stream);
        if (d_fo) cudaFreeAsync(d_fo, ctx->stream);
        delete m;
        ctx->set_error(std::string("mesh_create: ") + cudaGetErrorString(e), __FILE__, __LINE__);
        return (int)e;
    }
    // the caller's arrays are only borrowed for the call
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    m->d_xyz = d_xyz;
    m->d_face_vtx = d_fv;
    m->d_face_off = d_fo;
    m->owns_arrays = true;
    *out = m;
    return 0;
}

int mcb200_mesh_update_xyz(mcb200_ctx* ctx, mcb200_mesh* mesh, const void* xyz, uint32_t nv)
{
    if (!ctx || !mesh || !xyz) return MCB200_ERR_INVALID;
    if (nv != mesh->nv) MCB_FAIL(ctx, MCB200_ERR_INVALID, "mesh_update_xyz: vertex count differs from the mesh");
    if (!mesh->owns_arrays) MCB_FAIL(ctx, MCB200_ERR_INVALID, "mesh_update_xyz: the mesh borrows its arrays (adopted device memory)");
    MCB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t vbytes = (size_t)nv * 3 * (mesh->is_float ? sizeof(float) : sizeof(double));
    MCB_CUDA(ctx, cudaMemcpyAsync(const_cast<void*>(mesh->d_xyz), xyz, vbytes, cudaMemcpyHostToDevice, ctx->stream));
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // the caller's array is only borrowed for the call
    return 0;
}

int mcb200_mesh_validate(mcb200_ctx* ctx, mcb200_mesh* mesh, mcb200_validation* out)
{
    if (!ctx || !mesh || !out) return MCB200_ERR_INVALID;
    MCB_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->use_main();
    if (!mesh->validated) MCB_TRY(mesh_validate_run(ctx, mesh)); // (a mesh's faces never change; asking twice costs one read-back)
    uint32_t info[4];
    MCB_CUDA(ctx, cudaMemcpyAsync(info, mesh->cc_info.p, sizeof(info), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    out->n_components = info[0];
    out->n_border_edges = info[1];
    out->is_closed = info[1] == 0u ? 1 : 0;
    mesh->n_components = info[0];
    return 0;
}

int mcb200_mesh_winding_number(mcb200_ctx* ctx, mcb200_mesh* mesh, const double query[3], double* winding_number)
{
    if (!ctx || !mesh || !query || !winding_number) return MCB200_ERR_INVALID;
    MCB_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->use_main();
    MCB_TRY(mesh_winding_run(ctx, mesh, query));
    double out[2];
    MCB_CUDA(ctx, cudaMemcpyAsync(out, mesh->cc_wn.p, sizeof(out), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned unsupported;
    std::memcpy(&unsupported, &out[1], sizeof(unsigned));
    if (unsupported) MCB_FAIL(ctx, MCB200_ERR_INVALID, "winding_number: the mesh has faces with more than four vertices (host CDT path)");
    *winding_number = out[0];
    return 0;
}

int mcb200_intersection_type_without_cut(mcb200_ctx* ctx, mcb200_mesh* src, mcb200_mesh* cut, uint32_t* type)
{
    // check_and_store_input_mesh_intersection_type(), preproc.cpp:1999-2122, decision for decision
    if (!ctx || !src || !cut || !type) return MCB200_ERR_INVALID;
    if (!src->built || !cut->built) MCB_FAIL(ctx, MCB200_ERR_INVALID, "intersection_type: both meshes must be built (their mesh AABBs are an input)");
    mcb200_validation vs, vc;
    MCB_TRY(mcb200_mesh_validate(ctx, src, &vs)); // mesh_is_closed(): no border edge (preproc.cpp:1957-1990)
    MCB_TRY(mcb200_mesh_validate(ctx, cut, &vc));
    const bool sm_closed = vs.is_closed != 0, cm_closed = vc.is_closed != 0;
    double sb[6], cb[6];
    MCB_TRY(mcb200_bvh_read(ctx, src, nullptr, sb)); // bvhAABBs[0] of each mesh
    MCB_TRY(mcb200_bvh_read(ctx, cut, nullptr, cb));
    bool boxes_meet = true; // intersect_bounding_boxes(): closed intervals (math.h:931-941)
    for (int j = 0; j < 3; ++j)
        if (sb[j] > cb[3 + j] || cb[j] > sb[3 + j]) boxes_meet = false;
    const double eps = 1e-7; // windingNumberEps
    auto inside = [&](mcb200_mesh* point_of, mcb200_mesh* mesh, bool& in) -> int {
        double q[3], wn = 0.0;
        ctx->use_main();
        MCB_TRY(mesh_vertex_position(ctx, point_of, 0u, q)); // "pick any point (we chose the 1st)"
        MCB_TRY(mcb200_mesh_winding_number(ctx, mesh, q, &wn));
        in = std::fabs(1.0 - wn) < eps;
        return 0;
    };
    *type = MCB200_INTERSECTION_TYPE_NONE;
    if ((!sm_closed && !cm_closed) || !boxes_meet) return 0;
    bool in = false;
    if (sm_closed && cm_closed) {
        // the mesh with the larger AABB is tested first
        auto diag2 = [](const double* b) {
            const double x = b[3] - b[0], y = b[4] - b[1], z = b[5] - b[2];
            return 0.0 + x * x + y * y + z * z; // squared_length = dot_product, accumulated left to right (math.h:634-642)
        };
        const bool sm_larger = diag2(sb) > diag2(cb);
        mcb200_mesh* a = sm_larger ? src : cut;
        mcb200_mesh* b = sm_larger ? cut : src;
        MCB_TRY(inside(b, a, in));
        if (in) {
            *type = sm_larger ? MCB200_INTERSECTION_TYPE_INSIDE_SOURCEMESH : MCB200_INTERSECTION_TYPE_INSIDE_CUTMESH;
            return 0;
        }
        MCB_TRY(inside(src, b, in)); // the reference takes the SOURCE mesh's first vertex here, whichever mesh is "A" (:2053)
        if (in) *type = sm_larger ? MCB200_INTERSECTION_TYPE_INSIDE_CUTMESH : MCB200_INTERSECTION_TYPE_INSIDE_SOURCEMESH;
        return 0;
    }
    if (sm_closed) {
        MCB_TRY(inside(cut, src, in));
        if (in) *type = MCB200_INTERSECTION_TYPE_INSIDE_SOURCEMESH;
        return 0;
    }
    MCB_TRY(inside(src, cut, in));
    if (in) *type = MCB200_INTERSECTION_TYPE_INSIDE_CUTMESH;
    return 0;
}

int mcb200_mesh_read_components(mcb200_ctx* ctx, mcb200_mesh* mesh, int32_t* fccmap, int32_t* cc_vertex_count,
    int32_t* cc_face_count, size_t capacity_components)
{
    if (!ctx || !mesh) return MCB200_ERR_INVALID;
    if (!mesh->validated) MCB_FAIL(ctx, MCB200_ERR_INVALID, "read_components: run mcb200_mesh_validate first");
    MCB_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((cc_vertex_count || cc_face_count) && capacity_components < mesh->n_components)
        MCB_FAIL(ctx, MCB200_ERR_CAPACITY, "read_components: capacity smaller than the number of components");
    if (fccmap) MCB_CUDA(ctx, cudaMemcpyAsync(fccmap, mesh->cc_fmap.p, sizeof(int32_t) * mesh->nf, cudaMemcpyDeviceToHost, ctx->stream));
    if (cc_vertex_count)
        MCB_CUDA(ctx, cudaMemcpyAsync(cc_vertex_count, mesh->cc_vcount.p, sizeof(int32_t) * mesh->n_components, cudaMemcpyDeviceToHost, ctx->stream));
    if (cc_face_count)
        MCB_CUDA(ctx, cudaMemcpyAsync(cc_face_count, mesh->cc_fcount.p, sizeof(int32_t) * mesh->n_components, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mcb200_mesh_adopt_device(mcb200_ctx* ctx, int is_float, const void* d_xyz, uint32_t nv, const uint32_t* d_face_vtx,
    const uint32_t* d_face_off, uint32_t nf, uint32_t nh, mcb200_mesh** out)
{
    if (!ctx || !out) return MCB200_ERR_INVALID;
    *out = nullptr;
    if (!d_xyz || !d_face_vtx || nv == 0 || nf == 0) MCB_FAIL(ctx, MCB200_ERR_INVALID, "mesh_adopt_device: empty mesh or NULL array");
    if (!d_face_off && nh != 3u * nf) MCB_FAIL(ctx, MCB200_ERR_INVALID, "mesh_adopt_device: triangle mesh needs nh == 3 nf");
    mcb200_mesh* m = new mcb200_mesh();
    m->nv = nv;
    m->nf = nf;
    m->nh = nh;
    m->is_float = is_float ? 1 : 0;
    m->is_tri = d_face_off ? 0 : 1;
    m->d_xyz = d_xyz;
    m->d_face_vtx = d_face_vtx;
    m->d_face_off = d_face_off;
    m->owns_arrays = false;
    identity_frame(m->frame, m->is_float);
    *out = m;
    return 0;
}

int mcb200_mesh_set_frame(mcb200_ctx* ctx, mcb200_mesh* m, const double com[3], const double shift[3], const double pert[3])
{
    if (!ctx || !m) return MCB200_ERR_INVALID;
    identity_frame(m->frame, m->is_float);
    if (!com) {
        if (m->is_float) MCB_FAIL(ctx, MCB200_ERR_INVALID, "set_frame: internal-coordinate meshes must be double");
        if (pert) MCB_FAIL(ctx, MCB200_ERR_INVALID, "set_frame: a perturbation needs a frame");
        return 0;
    }
    if (!shift) MCB_FAIL(ctx, MCB200_ERR_INVALID, "set_frame: shift is NULL");
    m->frame.has_frame = 1;
    for (int j = 0; j < 3; ++j) {
        m->frame.com[j] = com[j];
        m->frame.shift[j] = shift[j];
        m->frame.fcom[j] = (float)com[j]; // preproc.cpp:124-129
        m->frame.fshift[j] = (float)shift[j];
        m->frame.pert[j] = pert ? pert[j] : 0.0;
    }
    m->frame.has_pert = pert ? 1 : 0;
    return 0;
}

void mcb200_mesh_free(mcb200_ctx* ctx, mcb200_mesh* m)
{
    if (!ctx || !m) return;
    cudaSetDevice(ctx->device);
    if (m->owns_arrays) {
        if (m->d_xyz) cudaFreeAsync(const_cast<void*>(m->d_xyz), ctx->stream);
        if (m->d_face_vtx) cudaFreeAsync(const_cast<uint32_t*>(m->d_face_vtx), ctx->stream);
        if (m->d_face_off) cudaFreeAsync(const_cast<uint32_t*>(m->d_face_off), ctx->stream);
    }
    ctx->release(m->face_bbox);
    ctx->release(m->prior_bbox);
    ctx->release(m->root);
    ctx->release(m->codes);
    ctx->release(m->sorted_codes);
    ctx->release(m->sorted_faces);
    ctx->release(m->wide);
    ctx->release(m->sorted_bbox);
    ctx->release(m->flags);
    ctx->release(m->groups);
    ctx->release(m->group_box);
    delete m->lv;
    for (dbuf* b : { &m->cc_label, &m->cc_id, &m->cc_vcount, &m->cc_fcount, &m->cc_fmap, &m->cc_info, &m->cc_wn }) ctx->release(*b);
    delete m;
}

// ---------------------------------------------------------------------------------------------------------- (1) build

int mcb200_bvh_build(mcb200_ctx* ctx, mcb200_mesh* m, double eps)
{
    if (!ctx || !m) return MCB200_ERR_INVALID;
    MCB_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->use_main();
    MCB_TRY(mesh_sync_frames(ctx, m, eps));
    return lbvh_build(ctx, m, eps);
}

int mcb200_mesh_set_prior_face_boxes(mcb200_ctx* ctx, mcb200_mesh* m, const double* boxes, uint32_t n)
{
    if (!ctx || !m || (n && !boxes)) return MCB200_ERR_INVALID;
    MCB_CUDA(ctx, cudaSetDevice(ctx->device));
    m->n_prior = 0;
    if (n == 0) return 0;
    if (n > m->nf) n = m->nf; // build_oibvh resizes the vector to the face count first
    ctx->use_main();
    MCB_TRY(ctx->reserve(m->prior_bbox, sizeof(double) * 6 * (size_t)n));
    // a rare path (one call per repartition retry): a plain blocking copy, visible to whichever lane builds next
    MCB_CUDA(ctx, cudaMemcpyAsync(m->prior_bbox.p, boxes, sizeof(double) * 6 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    m->n_prior = n;
    return 0;
}

int mcb200_bvh_read(mcb200_ctx* ctx, const mcb200_mesh* m, double* face_bboxes, double root_bbox[6])
{
    if (!ctx || !m) return MCB200_ERR_INVALID;
    if (!m->built) MCB_FAIL(ctx, MCB200_ERR_INVALID, "bvh_read: mesh has not been built");
    if (face_bboxes)
        MCB_CUDA(ctx, cudaMemcpyAsync(face_bboxes, m->face_bbox.p, sizeof(double) * 6 * (size_t)m->nf, cudaMemcpyDeviceToHost, ctx->stream));
    if (root_bbox) {
        const double* dec = reinterpret_cast<const double*>(m->root.as<unsigned long long>() + 6);
        MCB_CUDA(ctx, cudaMemcpyAsync(root_bbox, dec, sizeof(double) * 6, cudaMemcpyDeviceToHost, ctx->stream));
    }
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mcb200_bvh_read_morton(mcb200_ctx* ctx, const mcb200_mesh* m, uint32_t* codes_by_face, uint32_t* sorted_faces)
{
    if (!ctx || !m) return MCB200_ERR_INVALID;
    if (!m->built) MCB_FAIL(ctx, MCB200_ERR_INVALID, "bvh_read_morton: mesh has not been built");
    if (codes_by_face)
        MCB_CUDA(ctx, cudaMemcpyAsync(codes_by_face, m->codes.p, sizeof(uint32_t) * (size_t)m->nf, cudaMemcpyDeviceToHost, ctx->stream));
    if (sorted_faces)
        MCB_CUDA(ctx, cudaMemcpyAsync(sorted_faces, m->sorted_faces.p, sizeof(uint32_t) * (size_t)m->nf, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------- results

int mcb200_result_create(mcb200_ctx* ctx, mcb200_result** out)
{
    if (!ctx || !out) return MCB200_ERR_INVALID;
    *out = new mcb200_result();
    return 0;
}

void mcb200_result_free(mcb200_ctx* ctx, mcb200_result* r)
{
    if (!ctx || !r) return;
    cudaSetDevice(ctx->device);
    dbuf* all[] = { &r->counters, &r->pairs, &r->pairs_a, &r->pairs_b, &r->pair_cnt, &r->pair_off, &r->pair_tile, &r->cand_flag, &r->plane, &r->plane_mc, &r->exact_queue, &r->mid_queue, &r->cp_keys, &r->cp_idx, &r->cp_head, &r->cp_rank, &r->cp_tile, &r->cp_seg_key, &r->cp_seg_off, &r->cp_seg_vtx, &r->cp_info, &r->records, &r->rec_keys,
        &r->rec_idx, &r->records_sorted, &r->tests, &r->tests_sorted, &r->test_keys, &r->test_idx };
    for (dbuf* b : all) ctx->release(*b);
    delete r;
}

int mcb200_result_set_shard(mcb200_ctx* ctx, mcb200_result* res, uint32_t part, uint32_t nparts, uint32_t chunk)
{
    if (!ctx || !res) return MCB200_ERR_INVALID;
    if (nparts == 0 || part >= nparts || chunk == 0 || (chunk % 32u) != 0u)
        MCB_FAIL(ctx, MCB200_ERR_INVALID, "set_shard: need part < nparts and a chunk that is a positive multiple of 32");
    res->shard_part = part;
    res->shard_nparts = nparts;
    res->shard_chunk = chunk;
    return 0;
}

int mcb200_result_set_pair_capacity(mcb200_ctx* ctx, mcb200_result* res, uint64_t max_pairs)
{
    if (!ctx || !res) return MCB200_ERR_INVALID;
    if (max_pairs == 0) MCB_FAIL(ctx, MCB200_ERR_INVALID, "set_pair_capacity: capacity must be positive");
    res->cap_pairs = (size_t)max_pairs;
    res->h_valid = false;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------- (2) traversal

int mcb200_bvh_intersect(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_result* res)
{
    if (!ctx || !src || !cut || !res) return MCB200_ERR_INVALID;
    MCB_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->use_main();
    for (int attempt = 0; attempt < 2; ++attempt) {
        MCB_TRY(traverse_pairs(ctx, src, cut, res));
        MCB_TRY(sort_pairs(ctx, src, cut, res));
        if (attempt == 1) break;
        // The only host round trip of the stage: did the pair buffer hold everything?  (One 128-byte read; skipped by
        // mcb200_intersect_stage, which checks after the fact.)
        MCB_TRY(fetch_counters(ctx, res));
        if (!res->h.pair_overflow) break;
        res->cap_pairs = (size_t)res->h.n_pairs + (size_t)res->h.n_pairs / 8 + 1024;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------- (3) narrowphase

int mcb200_soup_create(mcb200_ctx* ctx, uint32_t nsf, uint32_t ncf, uint32_t nh, uint32_t ne, const uint32_t* face_vtx,
    const uint32_t* face_edge, const uint32_t* edge_f, mcb200_soup** out)
{
    if (!ctx || !out) return MCB200_ERR_INVALID;
    *out = nullptr;
    if (!face_vtx || !face_edge || !edge_f || nsf == 0 || ncf == 0) MCB_FAIL(ctx, MCB200_ERR_INVALID, "soup_create: NULL array or empty mesh");
    {
        // the narrowphase indexes face boxes by edge_f and edge_f by face_edge: an id out of range would be an illegal
        // address on the device (and poison the context), so it is refused here.  edge_f[2e] (the face of h0) may be
        // MCB200_NULL for a border edge of a repartitioned mesh: the h1 face then owns the edge's tests.
        const uint32_t nf = nsf + ncf;
        uint32_t bad = 0;
        for (size_t i = 0; i < (size_t)nh; ++i) bad |= (face_edge[i] >= ne) ? 1u : 0u;
        for (size_t i = 0; i < 2 * (size_t)ne; ++i) bad |= (edge_f[i] >= nf && edge_f[i] != MCB200_NULL) ? 1u : 0u;
        if (bad) MCB_FAIL(ctx, MCB200_ERR_INVALID, "soup_create: an edge id in face_edge or a face id in edge_f is out of range");
    }
    MCB_CUDA(ctx, cudaSetDevice(ctx->device));
    mcb200_soup* s = new mcb200_soup();
    s->nsf = nsf;
    s->ncf = ncf;
    s->nh = nh;
    s->ne = ne;
    s->all_tri = (nh == 3u * (nsf + ncf)) ? 1 : 0; // callers with polygons go through mcb200_soup_from_meshes
    int rc = upload(ctx, s->face_vtx, face_vtx, sizeof(uint32_t) * (size_t)nh);
    if (!rc) rc = upload(ctx, s->face_edge, face_edge, sizeof(uint32_t) * (size_t)nh);
    if (!rc) rc = upload(ctx, s->edge_f, edge_f, sizeof(uint32_t) * 2 * (size_t)ne);
    if (!rc) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = (int)e;
    }
    if (rc) {
        mcb200_soup_free(ctx, s);
        return rc;
    }
    *out = s;
    return 0;
}

int mcb200_soup_number(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_result* res, mcb200_soup** out)
{
    if (!ctx || !src || !cut || !res || !out) return MCB200_ERR_INVALID;
    *out = nullptr;
    MCB_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->use_main();
    MCB_TRY(ctx->reserve(res->counters, sizeof(result_counters_t)));
    mcb200_soup* s = new mcb200_soup();
    int rc = soup_number_reserve(ctx, src, cut, s);
    if (!rc) {
        result_counters_t* c = res->counters.as<result_counters_t>();
        fill_list_t fl {};
        fl.add(&c->soup_error, 2, 0u); // soup_error, soup_ne
        MCB_LAUNCH(ctx, k_fill, 1, 256, 0, fl);
        rc = soup_number_device(ctx, src, cut, s, c);
    }
    if (rc) {
        mcb200_soup_free(ctx, s);
        return rc;
    }
    res->h_valid = false;
    *out = s;
    return 0;
}

int mcb200_soup_create_sized(mcb200_ctx* ctx, uint32_t nsf, uint32_t ncf, uint32_t nh, uint32_t ne, const uint32_t* face_vtx,
    const uint32_t* face_edge, const uint32_t* edge_f, const uint32_t* face_sizes, mcb200_soup** out)
{
    MCB_TRY(mcb200_soup_create(ctx, nsf, ncf, nh, ne, face_vtx, face_edge, edge_f, out));
    if (!face_sizes) return 0;
    mcb200_soup* s = *out;
    const size_t nf = (size_t)nsf + ncf;
    std::vector<uint32_t> off(nf + 1);
    bool tri = true;
    uint32_t acc = 0;
    for (size_t f = 0; f < nf; ++f) {
        off[f] = acc;
        acc += face_sizes[f];
        if (face_sizes[f] != 3u) tri = false;
        if (face_sizes[f] < 3u) {
            mcb200_soup_free(ctx, s);
            *out = nullptr;
            MCB_FAIL(ctx, MCB200_ERR_INVALID, "soup_create_sized: a face has fewer than 3 vertices");
        }
    }
    off[nf] = acc;
    if (acc != nh) {
        mcb200_soup_free(ctx, s);
        *out = nullptr;
        MCB_FAIL(ctx, MCB200_ERR_INVALID, "soup_create_sized: face sizes do not add up to the halfedge count");
    }
    s->all_tri = tri ? 1 : 0;
    if (!tri) {
        MCB_TRY(upload(ctx, s->face_off, off.data(), sizeof(uint32_t) * off.size()));
        MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

void mcb200_soup_free(mcb200_ctx* ctx, mcb200_soup* s)
{
    if (!ctx || !s) return;
    cudaSetDevice(ctx->device);
    ctx->release(s->face_vtx);
    ctx->release(s->face_edge);
    ctx->release(s->face_off);
    ctx->release(s->edge_f);
    delete s;
}

int mcb200_soup_from_meshes(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_soup** out)
{
    if (!ctx || !src || !cut || !out) return MCB200_ERR_INVALID;
    *out = nullptr;
    if (src->h_face_vtx.empty() || cut->h_face_vtx.empty())
        MCB_FAIL(ctx, MCB200_ERR_INVALID, "soup_from_meshes: meshes adopted from device memory carry no host face arrays");
    const uint32_t nh = src->nh + cut->nh;
    std::vector<uint32_t> fv(nh), fe(nh), ev(2 * (size_t)nh), ef(2 * (size_t)nh);
    uint32_t ne = 0;
    const int rc = host_soup_ids(src->nv, src->h_face_off.data(), src->h_face_vtx.data(), src->nf, cut->h_face_off.data(),
        cut->h_face_vtx.data(), cut->nf, fv.data(), fe.data(), ev.data(), ef.data(), &ne);
    if (rc == MCB200_ERR_NON_MANIFOLD) MCB_FAIL(ctx, rc, "soup_from_meshes: non-manifold edge or inconsistent winding");
    if (rc) MCB_FAIL(ctx, rc, "soup_from_meshes: invalid face");
    MCB_TRY(mcb200_soup_create(ctx, src->nf, cut->nf, nh, ne, fv.data(), fe.data(), ef.data(), out));
    mcb200_soup* s = *out;
    s->all_tri = (src->is_tri && cut->is_tri) ? 1 : 0;
    if (!s->all_tri) {
        std::vector<uint32_t> off((size_t)src->nf + cut->nf + 1);
        for (uint32_t f = 0; f <= src->nf; ++f) off[f] = src->h_face_off[f];
        for (uint32_t f = 0; f <= cut->nf; ++f) off[src->nf + f] = src->nh + cut->h_face_off[f];
        MCB_TRY(upload(ctx, s->face_off, off.data(), sizeof(uint32_t) * off.size()));
        MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

int mcb200_narrowphase(mcb200_ctx* ctx, const mcb200_soup* soup, const mcb200_mesh* src, const mcb200_mesh* cut,
    mcb200_result* res, uint32_t flags)
{
    if (!ctx || !soup || !src || !cut || !res) return MCB200_ERR_INVALID;
    MCB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!soup->all_tri && !soup->face_off.p) MCB_FAIL(ctx, MCB200_ERR_INVALID, "narrowphase: polygon soup without face offsets");
    ctx->use_main();
    return narrowphase_run(ctx, soup, src, cut, res, flags);
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------------- the stage
// One body for both stage entry points.  Everything is enqueued from ctx->stream outwards (aux / background lanes fork from
// it by event and join it again before the body returns), so the body can be CAPTURED into a CUDA graph and replayed:
//   * all allocations happen before it (stage_reserve), the frames live in device memory (mesh_sync_frames),
//   * no host decision inside it depends on device data, every launch's arguments are functions of the signature below.
// wait_uploads: the inputs are still travelling on ctx->copy (mcb200_intersect_stage_host without a graph): each lane waits
// for the upload event of what it reads.  number_soup: 0 = the caller's soup as is, 1 = number the soup on the device,
// 2 = the caller's edge ids, vertex lists derived on the device.
int stage_body(mcb200_ctx* ctx, mcb200_mesh* src, mcb200_mesh* cut, double cut_eps, mcb200_soup* soup, mcb200_result* res,
    uint32_t flags, bool wait_uploads, int number_soup, bool interleave, bool order)
{
    ctx->use_main();
    // resets that nothing before the traversal depends on: up front, not between the kernels of the critical path
    MCB_TRY(result_reset_counters(ctx, res));
    res->counters_zeroed = true;
    MCB_TRY(narrowphase_prezero(ctx, soup, res));
    MCB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    MCB_CUDA(ctx, cudaStreamWaitEvent(ctx->aux, ctx->ev_fork, 0));
    int rc = 0;
    if (number_soup) {
        // the polygon soup needs the face arrays only: it is numbered on the lowest-priority lane while the builds run (and,
        // with uploads in flight, while the last coordinates still travel), and gives way to them whenever they have blocks to place
        ctx->use_bg();
        MCB_CUDA(ctx, cudaStreamWaitEvent(ctx->bg, ctx->ev_fork, 0));
        if (wait_uploads) MCB_CUDA(ctx, cudaStreamWaitEvent(ctx->bg, ctx->ev_up[2], 0));
        rc = number_soup == 1 ? soup_number_device(ctx, src, cut, soup, res->counters.as<result_counters_t>())
                              : soup_face_vtx_device(ctx, src, cut, soup);
        cudaEventRecord(ctx->ev_bg, ctx->bg);
        ctx->use_main();
        if (rc) return rc;
    }
    // ---- the two LBVH builds side by side: main lane <- src, aux lane <- cut ----
    if (wait_uploads) {
        MCB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_up[0], 0));
        MCB_CUDA(ctx, cudaStreamWaitEvent(ctx->aux, ctx->ev_up[1], 0));
    }
    if (interleave) {
        // issue the two builds' launches alternately (the host needs ~5 us per launch: one lane's ten kernels before the
        // other lane's first one would start that lane ~50 us late)
        std::vector<std::function<int()>> qa, qb;
        ctx->use_main();
        ctx->recording = &qa;
        rc = lbvh_build(ctx, src, 0.0);
        ctx->use_aux();
        ctx->recording = &qb;
        if (!rc) rc = lbvh_build(ctx, cut, cut_eps);
        ctx->recording = nullptr;
        ctx->use_main();
        for (size_t i = 0; !rc && (i < qa.size() || i < qb.size()); ++i) {
            if (i < qa.size()) rc = qa[i]();
            if (!rc && i < qb.size()) rc = qb[i]();
        }
    } else {
        ctx->use_main();
        rc = lbvh_build(ctx, src, 0.0);
        ctx->use_aux();
        if (!rc) rc = lbvh_build(ctx, cut, cut_eps);
        ctx->use_main();
    }
    cudaEventRecord(ctx->ev_join, ctx->aux);
    cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
    if (rc) return rc;
    // ---- traversal, then the pair sort (aux lane) next to the narrowphase (main lane, reads the unsorted pairs) ----
    MCB_TRY(traverse_pairs(ctx, src, cut, res));
    if (order) {
        MCB_CUDA(ctx, cudaEventRecord(ctx->ev_fork2, ctx->stream));
        MCB_CUDA(ctx, cudaStreamWaitEvent(ctx->aux, ctx->ev_fork2, 0));
        ctx->use_aux();
        rc = sort_pairs(ctx, src, cut, res);
        cudaEventRecord(ctx->ev_join2, ctx->aux);
        ctx->use_main();
    }
    if (number_soup) cudaStreamWaitEvent(ctx->stream, ctx->ev_bg, 0);
    if (wait_uploads && number_soup == 2) cudaStreamWaitEvent(ctx->stream, ctx->ev_up[3], 0);
    if (!rc) rc = narrowphase_run(ctx, soup, src, cut, res, order ? flags : (flags | MCB200_NARROW_INTERNAL_PARTIAL));
    if (order) cudaStreamWaitEvent(ctx->stream, ctx->ev_join2, 0);
    return rc;
}

// every allocation of a stage, on the main lane, before anything forks (the other lanes only ever see memory that exists)
int stage_reserve(mcb200_ctx* ctx, mcb200_mesh* src, mcb200_mesh* cut, const mcb200_soup* soup, mcb200_result* res, uint32_t flags)
{
    ctx->use_main();
    int rc = lbvh_reserve(ctx, src);
    if (!rc) rc = traverse_reserve(ctx, src, cut, res);
    if (!rc) rc = narrowphase_reserve(ctx, soup, res, flags);
    ctx->use_aux();
    if (!rc) rc = lbvh_reserve(ctx, cut);
    if (!rc) rc = sort_pairs_reserve(ctx, src, cut, res);
    ctx->use_main();
    return rc;
}

// ---- CUDA graphs ---------------------------------------------------------------------------------------------------------
// A stage is ~40 launches on three lanes; for small dispatches (the MultipleContextsInParallel pattern: thousands of
// 5k-triangle pairs) issuing them costs more than running them.  The body of a stage call is therefore captured the
// SECOND time a signature is seen (the first run makes every allocation and per-function attribute call) and replayed from
// then on: one cudaGraphLaunch per dispatch.  The signature names everything the captured launches depend on: array
// addresses and sizes, capacities, flags, shard; the frames are not part of it (device memory).  Any buffer (re)allocation in
// the context invalidates all graphs (alloc_epoch).  MCB200_GRAPHS=0 turns the mechanism off.
static std::vector<uint64_t> stage_signature(const mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, const mcb200_soup* soup,
    const mcb200_result* res, uint32_t flags, int number_soup, double cut_eps_is_param)
{
    (void)cut_eps_is_param;
    std::vector<uint64_t> g;
    auto mesh = [&](const mcb200_mesh* m) {
        g.push_back((uint64_t)(uintptr_t)m);
        g.push_back((uint64_t)(uintptr_t)m->d_xyz);
        g.push_back((uint64_t)(uintptr_t)m->d_face_vtx);
        g.push_back((uint64_t)(uintptr_t)m->d_face_off);
        g.push_back(((uint64_t)m->nv << 32) | m->nf);
        g.push_back(((uint64_t)m->nh << 8) | (uint64_t)(m->is_float ? 1 : 0) | (uint64_t)(m->is_tri ? 2 : 0));
    };
    mesh(src);
    mesh(cut);
    g.push_back((uint64_t)(uintptr_t)soup);
    g.push_back((uint64_t)(uintptr_t)soup->face_vtx.p);
    g.push_back((uint64_t)(uintptr_t)soup->face_edge.p);
    g.push_back((uint64_t)(uintptr_t)soup->edge_f.p);
    g.push_back((uint64_t)(uintptr_t)soup->face_off.p);
    g.push_back(((uint64_t)soup->nh << 32) | soup->ne);
    g.push_back(((uint64_t)soup->nsf << 32) | soup->ncf);
    g.push_back((uint64_t)(uintptr_t)res);
    g.push_back((uint64_t)res->cap_pairs);
    g.push_back(((uint64_t)res->shard_part << 40) | ((uint64_t)res->shard_nparts << 20) | res->shard_chunk);
    g.push_back(((uint64_t)flags << 8) | (uint64_t)number_soup | ((uint64_t)ctx->morton_sort_bits << 40) | ((uint64_t)(ctx->pdl ? 1 : 0) << 48));
    return g;
}

static void drop_stale_graphs(mcb200_ctx* ctx)
{
    for (size_t i = 0; i < ctx->graphs.size();) {
        if (ctx->graphs[i]->epoch != ctx->alloc_epoch) {
            if (ctx->graphs[i]->exec) cudaGraphExecDestroy(ctx->graphs[i]->exec);
            delete ctx->graphs[i];
            ctx->graphs.erase(ctx->graphs.begin() + (long)i);
        } else {
            ++i;
        }
    }
}

// runs the body of a stage: replayed from a graph when one exists, captured when the signature is seen for the second time
static int stage_run(mcb200_ctx* ctx, mcb200_mesh* src, mcb200_mesh* cut, double cut_eps, mcb200_soup* soup, mcb200_result* res,
    uint32_t flags, bool wait_uploads, int number_soup, bool interleave)
{
    const bool eligible = ctx->use_graphs && !ctx->profiling && !wait_uploads && src->n_prior == 0 && cut->n_prior == 0;
    if (!eligible) return stage_body(ctx, src, cut, cut_eps, soup, res, flags, wait_uploads, number_soup, interleave);
    drop_stale_graphs(ctx);
    const std::vector<uint64_t> sig = stage_signature(ctx, src, cut, soup, res, flags, number_soup, cut_eps);
    mcb200_graph* g = nullptr;
    for (mcb200_graph* e : ctx->graphs)
        if (e->sig == sig) g = e;
    auto finish = [&](mcb200_graph* e) { // what a run of the body leaves behind on the host
        const size_t cap_keep = res->cap_pairs;
        *res = e->res_state;
        res->cap_pairs = cap_keep;
        res->h_valid = false;
        src->built = cut->built = true;
        src->groups_valid = cut->groups_valid = true;
        src->eps = 0.0;
        cut->eps = cut_eps;
    };
    if (g && g->exec) {
        MCB_CUDA(ctx, cudaGraphLaunch(g->exec, ctx->stream));
        ctx->launches += g->launches;
        finish(g);
        return 0;
    }
    if (!g) { // first sighting: a plain run (allocations, attribute calls), remember the signature
        const int rc = stage_body(ctx, src, cut, cut_eps, soup, res, flags, false, number_soup, interleave);
        if (rc) return rc;
        if (ctx->graphs.size() < 64) {
            g = new mcb200_graph();
            // the run may have allocated: take the signature and the epoch as they are NOW
            g->sig = stage_signature(ctx, src, cut, soup, res, flags, number_soup, cut_eps);
            g->epoch = ctx->alloc_epoch;
            ctx->graphs.push_back(g);
        }
        return 0;
    }
    if (g->refused) return stage_body(ctx, src, cut, cut_eps, soup, res, flags, false, number_soup, interleave);
    // second sighting: capture
    const uint64_t launches0 = ctx->launches;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
    int rc = 0;
    if (e == cudaSuccess) {
        rc = stage_body(ctx, src, cut, cut_eps, soup, res, flags, false, number_soup, false);
        e = cudaStreamEndCapture(ctx->stream, &graph);
    }
    if (e == cudaSuccess && !rc && graph && g->epoch == ctx->alloc_epoch) e = cudaGraphInstantiate(&g->exec, graph, 0);
    else if (e == cudaSuccess) e = cudaErrorUnknown;
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess || !g->exec) {
        cudaGetLastError(); // a failed capture leaves a sticky-looking error behind; the stream itself is fine
        g->exec = nullptr;
        g->refused = true;
        if (rc) return rc;
        return stage_body(ctx, src, cut, cut_eps, soup, res, flags, false, number_soup, interleave);
    }
    g->launches = ctx->launches - launches0;
    g->res_state = *res;
    ctx->launches = launches0;
    MCB_CUDA(ctx, cudaGraphLaunch(g->exec, ctx->stream));
    ctx->launches += g->launches;
    finish(g);
    return 0;
}

extern "C" {

int mcb200_intersect_stage(mcb200_ctx* ctx, mcb200_mesh* src, mcb200_mesh* cut, double cut_eps, const mcb200_soup* soup,
    mcb200_result* res, uint32_t flags)
{
    if (!ctx || !src || !cut || !soup || !res) return MCB200_ERR_INVALID;
    MCB_CUDA(ctx, cudaSetDevice(ctx->device));
    MCB_TRY(stage_reserve(ctx, src, cut, soup, res, flags));
    // the frames into their device slots (the cut BVH is built from the UNPERTURBED cut frame: the reference builds it on the
    // first pass only, preproc.cpp:2676-2698; a perturbation only moves the narrowphase coordinates)
    MCB_TRY(mesh_sync_frames(ctx, src, 0.0, cut, cut_eps));
    return stage_run(ctx, src, cut, cut_eps, const_cast<mcb200_soup*>(soup), res, flags, false, 0, true);
}

// Describe the host arrays of one mesh in the context-owned staging mesh `k` (0 source, 1 cut) and size its buffers.
static int stage_mesh_describe(mcb200_ctx* ctx, int k, const mcb200_host_mesh* hm, bool resident, std::vector<uint32_t>& off)
{
    if (!hm || !hm->xyz || !hm->face_vtx || hm->nv == 0 || hm->nf == 0) MCB_FAIL(ctx, MCB200_ERR_INVALID, "stage_host: empty mesh or NULL array");
    if (!ctx->st_mesh[k]) {
        ctx->st_mesh[k] = new mcb200_mesh();
        ctx->st_mesh[k]->owns_arrays = false;
    }
    mcb200_mesh* m = ctx->st_mesh[k];
    if (resident) {
        // the caller vouches that this mesh's arrays are the ones of the previous call on this context: nothing travels
        if (!m->d_xyz || m->nv != hm->nv || m->nf != hm->nf || m->is_float != (hm->is_float ? 1 : 0))
            MCB_FAIL(ctx, MCB200_ERR_INVALID, "stage_host: *_RESIDENT flag, but the staged mesh has different counts (or none was staged)");
        m->built = false;
        return 0;
    }
    m->nv = hm->nv;
    m->nf = hm->nf;
    m->is_float = hm->is_float ? 1 : 0;
    m->is_tri = 1;
    m->h_face_vtx.clear(); // host copies are only needed by mcb200_soup_from_meshes; the staged path never keeps them
    m->h_face_off.clear();
    uint32_t nh = 3u * hm->nf;
    off.clear();
    if (hm->face_sizes) {
        off.resize((size_t)hm->nf + 1);
        uint32_t acc = 0;
        for (uint32_t f = 0; f < hm->nf; ++f) {
            if (hm->face_sizes[f] < 3) MCB_FAIL(ctx, MCB200_ERR_INVALID, "stage_host: a face has fewer than 3 vertices");
            if (hm->face_sizes[f] != 3) m->is_tri = 0;
            off[f] = acc;
            acc += hm->face_sizes[f];
        }
        off[hm->nf] = acc;
        nh = acc;
    }
    m->nh = nh;
    const size_t vbytes = (size_t)hm->nv * 3 * (hm->is_float ? sizeof(float) : sizeof(double));
    MCB_TRY(ctx->reserve(ctx->st_xyz[k], vbytes));
    MCB_TRY(ctx->reserve(ctx->st_fv[k], sizeof(uint32_t) * (size_t)nh));
    if (!m->is_tri) MCB_TRY(ctx->reserve(ctx->st_fo[k], sizeof(uint32_t) * ((size_t)hm->nf + 1)));
    m->d_xyz = ctx->st_xyz[k].p;
    m->d_face_vtx = ctx->st_fv[k].as<uint32_t>();
    m->d_face_off = m->is_tri ? nullptr : ctx->st_fo[k].as<uint32_t>();
    m->built = false;
    return 0;
}

// The uploads of one mesh on stream `st`.  `last`: the mesh that travels last sends its faces first — with them the polygon
// soup can be numbered while the coordinates are still on their way.  Records ev_up[k] (mesh landed) and, for the last mesh,
// ev_up[2] (all face arrays landed).
static int stage_mesh_upload(mcb200_ctx* ctx, cudaStream_t st, int k, const mcb200_host_mesh* hm, bool last, bool resident,
    const std::vector<uint32_t>& off, bool events)
{
    mcb200_mesh* m = ctx->st_mesh[k];
    if (!resident) {
        const size_t vbytes = (size_t)hm->nv * 3 * (hm->is_float ? sizeof(float) : sizeof(double));
        if (!last) MCB_CUDA(ctx, cudaMemcpyAsync(ctx->st_xyz[k].p, hm->xyz, vbytes, cudaMemcpyHostToDevice, st));
        MCB_CUDA(ctx, cudaMemcpyAsync(ctx->st_fv[k].p, hm->face_vtx, sizeof(uint32_t) * (size_t)m->nh, cudaMemcpyHostToDevice, st));
        if (!m->is_tri) {
            // `off` is a local of the caller: this (small, polygon-only) copy must complete before it goes out of scope
            MCB_CUDA(ctx, cudaMemcpyAsync(ctx->st_fo[k].p, off.data(), sizeof(uint32_t) * off.size(), cudaMemcpyHostToDevice, st));
            MCB_CUDA(ctx, cudaStreamSynchronize(st));
        }
        if (last) {
            if (events) MCB_CUDA(ctx, cudaEventRecord(ctx->ev_up[2], st)); // all face arrays are on the device
            MCB_CUDA(ctx, cudaMemcpyAsync(ctx->st_xyz[k].p, hm->xyz, vbytes, cudaMemcpyHostToDevice, st));
        }
    } else if (last && events) {
        MCB_CUDA(ctx, cudaEventRecord(ctx->ev_up[2], st));
    }
    if (events) MCB_CUDA(ctx, cudaEventRecord(ctx->ev_up[k], st));
    return 0;
}

int mcb200_intersect_stage_host(mcb200_ctx* ctx, const mcb200_host_mesh* hsrc, const mcb200_host_mesh* hcut, const double com[3],
    const double shift[3], const double perturbation[3], double cut_eps, const mcb200_host_soup* hsoup, mcb200_result* res,
    uint32_t flags)
{
    if (!ctx || !hsrc || !hcut || !res) return MCB200_ERR_INVALID;
    MCB_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->use_main();
    flags |= MCB200_NARROW_INTERNAL_LAZY_RADIX; // (results of this call are only reachable through mcb200_result_counts)
    const bool src_res = (flags & MCB200_STAGE_SRC_RESIDENT) != 0, cut_res = (flags & MCB200_STAGE_CUT_RESIDENT) != 0;
    std::vector<uint32_t> off_s, off_c;
    MCB_TRY(stage_mesh_describe(ctx, 0, hsrc, src_res, off_s));
    MCB_TRY(stage_mesh_describe(ctx, 1, hcut, cut_res, off_c));
    mcb200_mesh* src = ctx->st_mesh[0];
    mcb200_mesh* cut = ctx->st_mesh[1];
    MCB_TRY(mcb200_mesh_set_frame(ctx, src, com, shift, nullptr));
    MCB_TRY(mcb200_mesh_set_frame(ctx, cut, com, shift, perturbation));
    if (!ctx->st_soup) ctx->st_soup = new mcb200_soup();
    mcb200_soup* soup = ctx->st_soup;
    const bool number_on_device = (hsoup == nullptr);
    if (number_on_device) {
        MCB_TRY(soup_number_reserve(ctx, src, cut, soup));
    } else {
        if (hsoup->nh != src->nh + cut->nh || !hsoup->face_edge || !hsoup->edge_f)
            MCB_FAIL(ctx, MCB200_ERR_INVALID, "stage_host: soup halfedge count does not match the meshes");
        soup->nsf = src->nf;
        soup->ncf = cut->nf;
        soup->nh = hsoup->nh;
        soup->ne = hsoup->ne;
        soup->all_tri = (src->is_tri && cut->is_tri) ? 1 : 0;
        MCB_TRY(ctx->reserve(soup->face_vtx, sizeof(uint32_t) * (size_t)soup->nh));
        MCB_TRY(ctx->reserve(soup->face_edge, sizeof(uint32_t) * (size_t)soup->nh));
        MCB_TRY(ctx->reserve(soup->edge_f, sizeof(uint32_t) * 2 * (size_t)(soup->ne ? soup->ne : 1)));
        if (!soup->all_tri) MCB_TRY(ctx->reserve(soup->face_off, sizeof(uint32_t) * ((size_t)soup->nsf + soup->ncf + 1)));
    }
    // every other allocation of the stage, still before any lane forks
    MCB_TRY(stage_reserve(ctx, src, cut, soup, res, flags));
    const int number_soup = number_on_device ? 1 : 2;
    // Small dispatches are bound by the cost of issuing ~45 launches, not by the uploads (a few hundred KB): everything
    // travels on the main stream and the body is replayed from a graph.  Large ones keep the pipelined uploads on the copy
    // stream (the builds start as each mesh lands), where the launches do not matter.
    const bool small = ctx->use_graphs && !ctx->profiling && ((size_t)src->nf + cut->nf) <= ctx->graph_max_faces;
    if (small) {
        // uploads in the order of use on the main stream itself; the previous call's kernels are ahead of them in the stream
        MCB_TRY(stage_mesh_upload(ctx, ctx->stream, 0, hsrc, false, src_res, off_s, false));
        MCB_TRY(stage_mesh_upload(ctx, ctx->stream, 1, hcut, false, cut_res, off_c, false));
        if (!number_on_device) {
            MCB_CUDA(ctx, cudaMemcpyAsync(soup->face_edge.p, hsoup->face_edge, sizeof(uint32_t) * (size_t)soup->nh, cudaMemcpyHostToDevice, ctx->stream));
            MCB_CUDA(ctx, cudaMemcpyAsync(soup->edge_f.p, hsoup->edge_f, sizeof(uint32_t) * 2 * (size_t)soup->ne, cudaMemcpyHostToDevice, ctx->stream));
        }
        MCB_TRY(mesh_sync_frames(ctx, src, 0.0, cut, cut_eps));
        return stage_run(ctx, src, cut, cut_eps, soup, res, flags, false, number_soup, false);
    }
    // the previous call's kernels may still be reading the staging buffers
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // ---- uploads on the copy stream, in the order the stage consumes them: the mesh with more faces travels last ----
    MCB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    MCB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy, ctx->ev_fork, 0));
    if (hcut->nf > hsrc->nf) {
        MCB_TRY(stage_mesh_upload(ctx, ctx->copy, 0, hsrc, false, src_res, off_s, true));
        MCB_TRY(stage_mesh_upload(ctx, ctx->copy, 1, hcut, true, cut_res, off_c, true));
    } else {
        MCB_TRY(stage_mesh_upload(ctx, ctx->copy, 1, hcut, false, cut_res, off_c, true));
        MCB_TRY(stage_mesh_upload(ctx, ctx->copy, 0, hsrc, true, src_res, off_s, true));
    }
    if (!number_on_device) {
        MCB_CUDA(ctx, cudaMemcpyAsync(soup->face_edge.p, hsoup->face_edge, sizeof(uint32_t) * (size_t)soup->nh, cudaMemcpyHostToDevice, ctx->copy));
        MCB_CUDA(ctx, cudaMemcpyAsync(soup->edge_f.p, hsoup->edge_f, sizeof(uint32_t) * 2 * (size_t)soup->ne, cudaMemcpyHostToDevice, ctx->copy));
        MCB_CUDA(ctx, cudaEventRecord(ctx->ev_up[3], ctx->copy));
    }
    MCB_TRY(mesh_sync_frames(ctx, src, 0.0, cut, cut_eps));
    return stage_body(ctx, src, cut, cut_eps, soup, res, flags, true, number_soup, false);
}

int mcb200_batch_intersect_host(mcb200_ctx** ctxs, mcb200_result** results, uint32_t nctx, const mcb200_batch_item* items, uint32_t n,
    mcb200_counts* counts)
{
    if (!ctxs || !results || nctx == 0 || (n && (!items || !counts))) return MCB200_ERR_INVALID;
    for (uint32_t k = 0; k < nctx; ++k)
        if (!ctxs[k] || !results[k]) return MCB200_ERR_INVALID;
    struct frame_of_item {
        double com[3], shift[3], eps;
    };
    std::vector<frame_of_item> fr(nctx); // the frame of the item a lane is working on
    std::vector<uint32_t> in_flight(nctx, MCB200_NULL);
    int first_error = 0;
    double src_stats[9];
    const void* src_stats_of = nullptr;
    uint32_t src_stats_nv = 0;
    auto enqueue = [&](uint32_t lane, uint32_t i) -> int {
        const mcb200_batch_item& it = items[i];
        frame_of_item& f = fr[lane];
        const double *com = it.com, *shift = it.shift;
        double eps = it.cut_eps;
        if (!com) {
            double sb[6], cb[6], cs[9];
            if (it.src.is_float != it.cut.is_float) return MCB200_ERR_INVALID;
            // the source statistics are kept while the caller vouches (MCB200_STAGE_SRC_RESIDENT) that the arrays are unchanged
            if (!((it.flags & MCB200_STAGE_SRC_RESIDENT) && src_stats_of == it.src.xyz && src_stats_nv == it.src.nv)) {
                mcb200_vertex_stats(it.src.is_float, it.src.xyz, it.src.nv, src_stats);
                src_stats_of = it.src.xyz;
                src_stats_nv = it.src.nv;
            }
            mcb200_vertex_stats(it.cut.is_float, it.cut.xyz, it.cut.nv, cs);
            mcb200_vertex_parameters_from_stats(src_stats, cs, f.com, f.shift, sb, cb);
            f.eps = mcb200_cut_bbox_eps(cb, it.gp_constant > 0.0 ? it.gp_constant : 1e-4, 0);
            com = f.com;
            shift = f.shift;
            eps = f.eps;
        }
        return mcb200_intersect_stage_host(ctxs[lane], &it.src, &it.cut, com, shift, it.perturbation, eps, nullptr, results[lane], it.flags);
    };
    auto collect = [&](uint32_t lane) -> int {
        const uint32_t i = in_flight[lane];
        if (i == MCB200_NULL) return 0;
        in_flight[lane] = MCB200_NULL;
        int rc = mcb200_result_counts(ctxs[lane], results[lane], &counts[i]);
        for (int attempt = 0; rc == MCB200_ERR_CAPACITY && attempt < 4; ++attempt) { // a buffer was too small: it has grown, run again
            rc = enqueue(lane, i);
            if (!rc) rc = mcb200_result_counts(ctxs[lane], results[lane], &counts[i]);
        }
        return rc;
    };
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t lane = i % nctx;
        int rc = collect(lane);
        if (!rc) rc = enqueue(lane, i);
        if (rc) {
            if (!first_error) first_error = rc;
            std::memset(&counts[i], 0, sizeof(mcb200_counts));
            counts[i].status = -1;
            continue;
        }
        in_flight[lane] = i;
    }
    for (uint32_t lane = 0; lane < nctx; ++lane) {
        const int rc = collect(lane);
        if (rc && !first_error) first_error = rc;
    }
    return first_error;
}

// Reads back the polygon-soup ids the last mcb200_intersect_stage_host call used (tests: device numbering == mcb200_soup_ids).
int mcb200_staged_soup_read(mcb200_ctx* ctx, uint32_t* face_vtx, uint32_t* face_edge, uint32_t* edge_f, uint32_t capacity_edges,
    uint32_t* nh_out, uint32_t* ne_out)
{
    if (!ctx || !ctx->st_soup) return MCB200_ERR_INVALID;
    MCB_CUDA(ctx, cudaSetDevice(ctx->device));
    mcb200_soup* soup = ctx->st_soup;
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // device numbering leaves the edge count in the ids themselves: the largest id + 1
    std::vector<uint32_t> fe(soup->nh);
    MCB_CUDA(ctx, cudaMemcpy(fe.data(), soup->face_edge.p, sizeof(uint32_t) * fe.size(), cudaMemcpyDeviceToHost));
    uint32_t ne = 0;
    for (uint32_t e : fe) ne = e + 1 > ne ? e + 1 : ne;
    if (nh_out) *nh_out = soup->nh;
    if (ne_out) *ne_out = ne;
    if (face_edge) std::memcpy(face_edge, fe.data(), sizeof(uint32_t) * fe.size());
    if (face_vtx) MCB_CUDA(ctx, cudaMemcpy(face_vtx, soup->face_vtx.p, sizeof(uint32_t) * soup->nh, cudaMemcpyDeviceToHost));
    if (edge_f) {
        if (capacity_edges < ne) MCB_FAIL(ctx, MCB200_ERR_CAPACITY, "staged_soup_read: edge_f capacity too small");
        MCB_CUDA(ctx, cudaMemcpy(edge_f, soup->edge_f.p, sizeof(uint32_t) * 2 * (size_t)ne, cudaMemcpyDeviceToHost));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------- reads


int mcb200_result_counts(mcb200_ctx* ctx, mcb200_result* res, mcb200_counts* out)
{
    if (!ctx || !res || !out) return MCB200_ERR_INVALID;
    MCB_TRY(fetch_counters(ctx, res));
    const result_counters_t& h = res->h;
    out->n_pairs = h.n_pairs;
    out->n_node_tests = h.n_node_tests;
    out->n_tests = h.n_tests;
    out->n_exact = h.n_exact;
    out->n_records = h.n_records;
    out->n_cand_faces = h.n_cand_faces;
    out->bad_face = h.bad_face;
    // status precedence follows dispatch(): the degenerate-face check returns before any edge/face test runs
    // (kernel.cpp:2301-2312); note the reference's `>` when naming the mesh (kernel.cpp:2304)
    if (res->have_narrow && h.bad_face != MCB200_NULL)
        out->status = (h.bad_face > res->nsf) ? MCB200_STATUS_INVALID_CUT_MESH : MCB200_STATUS_INVALID_SRC_MESH;
    else if (res->have_narrow && h.gp_violation)
        out->status = MCB200_STATUS_GENERAL_POSITION_VIOLATION;
    else
        out->status = MCB200_STATUS_SUCCESS;
    if (h.soup_error)
        MCB_FAIL(ctx, MCB200_ERR_NON_MANIFOLD, "polygon-soup numbering: an edge is shared by three faces or by two faces wound the same way");
    if (res->have_narrow && (h.n_records > res->cap_records || narrow_queue_overflow(res, h) || (res->logged_tests && h.n_log > res->cap_tests))) {
        // The narrowphase buffers are sized from the pair capacity (2 records, 6 queue entries per pair); an input that
        // needs more gets a pair capacity that provides it, and the caller runs the stage again — nothing is dropped silently.
        size_t need = res->cap_pairs;
        if (h.n_records > res->cap_records) need = std::max(need, (size_t)h.n_records / 2 + 1024);
        if (narrow_queue_overflow(res, h) || (res->logged_tests && h.n_log > res->cap_tests)) need = std::max(need, res->cap_pairs * 2);
        res->cap_pairs = need;
        res->h_valid = false;
        MCB_FAIL(ctx, MCB200_ERR_CAPACITY, "narrowphase buffer overflow: capacity has been raised, run the stage again");
    }
    if (h.pair_overflow) {
        // regrow for the caller's retry: the counter kept counting past the capacity, so the needed size is known
        res->cap_pairs = (size_t)h.n_pairs + (size_t)h.n_pairs / 8 + 1024;
        res->h_valid = false;
        MCB_FAIL(ctx, MCB200_ERR_CAPACITY, "pair buffer overflow: capacity has been raised, run the stage again");
    }
    return 0;
}

int mcb200_result_queue_counts(mcb200_ctx* ctx, mcb200_result* res, uint64_t out[4])
{
    if (!ctx || !res || !out) return MCB200_ERR_INVALID;
    MCB_TRY(fetch_counters(ctx, res));
    out[0] = res->tri_queues ? res->h.n_mid : res->h.n_pairs;
    out[1] = res->h.n_queue;
    out[2] = res->h.n_cross;
    out[3] = res->h.n_full;
    return 0;
}

int mcb200_result_read_pairs(mcb200_ctx* ctx, mcb200_result* res, uint64_t* pairs, size_t capacity)
{
    if (!ctx || !res) return MCB200_ERR_INVALID;
    MCB_TRY(fetch_counters(ctx, res));
    const size_t n = (size_t)res->h.n_pairs;
    if (n > res->cap_pairs) MCB_FAIL(ctx, MCB200_ERR_CAPACITY, "pair buffer overflowed on the device");
    if (n > capacity || (n && !pairs)) MCB_FAIL(ctx, MCB200_ERR_CAPACITY, "read_pairs: output array too small");
    if (n && !res->pairs_sorted) MCB_FAIL(ctx, MCB200_ERR_INTERNAL, "read_pairs: pairs were not sorted");
    if (n) MCB_CUDA(ctx, cudaMemcpyAsync(pairs, res->pairs_sorted, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mcb200_result_read_records(mcb200_ctx* ctx, mcb200_result* res, mcb200_record* records, size_t capacity)
{
    if (!ctx || !res) return MCB200_ERR_INVALID;
    if (!res->have_narrow) MCB_FAIL(ctx, MCB200_ERR_INVALID, "read_records: narrowphase has not run");
    MCB_TRY(fetch_counters(ctx, res));
    const size_t n = (size_t)res->h.n_records;
    if (n > res->cap_records) MCB_FAIL(ctx, MCB200_ERR_CAPACITY, "record buffer overflowed on the device");
    if (n > capacity || (n && !records)) MCB_FAIL(ctx, MCB200_ERR_CAPACITY, "read_records: output array too small");
    if (n) MCB_CUDA(ctx, cudaMemcpyAsync(records, res->records_sorted.p, sizeof(mcb200_record) * n, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mcb200_result_read_tests(mcb200_ctx* ctx, mcb200_result* res, mcb200_test* tests, size_t capacity)
{
    if (!ctx || !res) return MCB200_ERR_INVALID;
    if (!res->have_narrow || !res->logged_tests) MCB_FAIL(ctx, MCB200_ERR_INVALID, "read_tests: run the narrowphase with MCB200_NARROW_LOG_TESTS");
    MCB_TRY(fetch_counters(ctx, res));
    const size_t n = (size_t)res->h.n_log;
    if (n > res->cap_tests) MCB_FAIL(ctx, MCB200_ERR_CAPACITY, "test log overflowed on the device");
    if (n > capacity || (n && !tests)) MCB_FAIL(ctx, MCB200_ERR_CAPACITY, "read_tests: output array too small");
    if (n) MCB_CUDA(ctx, cudaMemcpyAsync(tests, res->tests_sorted.p, sizeof(mcb200_test) * n, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mcb200_result_read_planes(mcb200_ctx* ctx, mcb200_result* res, uint32_t* faces, double* normal, double* d, int32_t* max_comp,
    size_t capacity)
{
    if (!ctx || !res) return MCB200_ERR_INVALID;
    if (!res->have_narrow) MCB_FAIL(ctx, MCB200_ERR_INVALID, "read_planes: narrowphase has not run");
    MCB_TRY(fetch_counters(ctx, res));
    const size_t n = (size_t)res->h.n_cand_faces;
    if (n > capacity) MCB_FAIL(ctx, MCB200_ERR_CAPACITY, "read_planes: output arrays too small");
    if (n == 0) return 0;
    std::vector<double> pl(4 * n);
    std::vector<int32_t> mc(n);
    std::vector<uint32_t> fc(n);
    MCB_CUDA(ctx, cudaMemcpyAsync(pl.data(), res->plane.p, sizeof(double) * 4 * n, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(ctx, cudaMemcpyAsync(mc.data(), res->plane_mc.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(ctx, cudaMemcpyAsync(fc.data(), res->plane_mc.as<int32_t>() + res->nf_ps, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost,
        ctx->stream));
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // rows were appended in scheduling order; hand them back by ascending face id
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return fc[x] < fc[y]; });
    for (size_t i = 0; i < n; ++i) {
        const uint32_t k = order[i];
        if (faces) faces[i] = fc[k];
        if (normal) {
            normal[3 * i] = pl[4 * (size_t)k];
            normal[3 * i + 1] = pl[4 * (size_t)k + 1];
            normal[3 * i + 2] = pl[4 * (size_t)k + 2];
        }
        if (d) d[i] = pl[4 * (size_t)k + 3];
        if (max_comp) max_comp[i] = mc[k];
    }
    return 0;
}

int mcb200_result_device_ptr(mcb200_ctx* ctx, mcb200_result* res, int which, void** dptr, uint64_t* count)
{
    if (!ctx || !res || !dptr || !count) return MCB200_ERR_INVALID;
    MCB_TRY(fetch_counters(ctx, res));
    if (which == 0) {
        *dptr = res->pairs_sorted;
        *count = res->h.n_pairs;
    } else if (which == 1) {
        if (!res->have_narrow) MCB_FAIL(ctx, MCB200_ERR_INVALID, "device_ptr: narrowphase has not run");
        *dptr = res->records_sorted.p;
        *count = res->h.n_records;
    } else {
        MCB_FAIL(ctx, MCB200_ERR_INVALID, "device_ptr: unknown buffer");
    }
    return 0;
}

} // extern "C"
