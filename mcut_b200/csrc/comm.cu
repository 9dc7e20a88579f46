// mcut_b200/csrc/comm.cu — one huge dispatch on several GPUs (BASELINE north star: "every GPU holds replicated
// meshes and BVHs and traverses its own slice of the source-mesh leaf range, with pair and intersection buffers gathered by
// NCCL allgatherv over NVLink").
//
// One process per GPU, one mcb200_ctx each.  Every rank holds both meshes and builds both trees (the build is a fraction of
// a dense dispatch and cheaper than shipping a tree); rank r walks the query groups of its 4096-leaf chunks of the Morton
// order (dealt round-robin: the intersection curve is never owned by one rank) and runs the narrowphase on ITS pairs.
// Then ONE exchange, enqueued on the context's stream:
//     ncclAllGather   the ranks' counter blocks (104 bytes each)            -> one host read: every rank knows all counts
//     ncclBroadcast x N (grouped)  pairs and registry records of every rank -> each lands at its offset in the merged lists
//     ncclAllReduce   per-source-face pair counts (sum), candidate-face flags (max)
// after which every rank puts the merged lists in canonical order (counting order of the pairs, record sort) and makes the
// plane rows of all candidate faces: the result on every rank is byte-for-byte the single-GPU result.
// NCCL is loaded at run time (dlopen "libnccl.so.2": the copy the process already has — e.g. PyTorch's — is reused), so the
// library itself has no NCCL dependency and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>

#include "internal.h"

namespace {

struct nccl_api_t {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

nccl_api_t g_nccl;
std::string g_nccl_error;

bool load_nccl()
{
    if (g_nccl.handle) return true;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    void* h = nullptr;
    for (const char* n : names)
        if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
    if (!h) {
        g_nccl_error = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "?");
        return false;
    }
    nccl_api_t a;
    a.handle = h;
#define MCB_SYM(field, name)                                                     \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name));               \
    if (!a.field) {                                                              \
        g_nccl_error = std::string("NCCL symbol missing: ") + name;              \
        return false;                                                            \
    }
    MCB_SYM(GetUniqueId, "ncclGetUniqueId")
    MCB_SYM(CommInitRank, "ncclCommInitRank")
    MCB_SYM(CommDestroy, "ncclCommDestroy")
    MCB_SYM(AllGather, "ncclAllGather")
    MCB_SYM(AllReduce, "ncclAllReduce")
    MCB_SYM(Broadcast, "ncclBroadcast")
    MCB_SYM(GroupStart, "ncclGroupStart")
    MCB_SYM(GroupEnd, "ncclGroupEnd")
    MCB_SYM(GetErrorString, "ncclGetErrorString")
#undef MCB_SYM
    g_nccl = a;
    return true;
}

#define MCB_NCCL(ctx, expr)                                                                                       \
    do {                                                                                                          \
        ncclResult_t r__ = (expr);                                                                                \
        if (r__ != ncclSuccess) {                                                                                 \
            (ctx)->set_error(std::string(#expr) + ": " + g_nccl.GetErrorString(r__), __FILE__, __LINE__);         \
            return MCB200_ERR_INTERNAL;                                                                           \
        }                                                                                                         \
    } while (0)

__global__ void k_set_counters(result_counters_t* dst, result_counters_t v)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) *dst = v;
}

} // namespace

struct mcb200_comm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    mcb200_ctx* ctx = nullptr;
    dbuf all_counters; // result_counters_t [nranks]
    dbuf pairs_merged, records_merged;
    void* h_counters = nullptr; // pinned, result_counters_t [nranks]
};

static_assert(sizeof(ncclUniqueId) == MCB200_COMM_ID_BYTES, "mcb200_comm id size");

extern "C" {

int mcb200_comm_unique_id(char id[MCB200_COMM_ID_BYTES])
{
    if (!id) return MCB200_ERR_INVALID;
    if (!load_nccl()) return MCB200_ERR_NO_DEVICE;
    ncclUniqueId u;
    if (g_nccl.GetUniqueId(&u) != ncclSuccess) return MCB200_ERR_INTERNAL;
    std::memcpy(id, &u, sizeof(u));
    return 0;
}

int mcb200_comm_create(mcb200_ctx* ctx, int nranks, int rank, const char id[MCB200_COMM_ID_BYTES], mcb200_comm** out)
{
    if (!ctx || !out || !id || nranks < 1 || rank < 0 || rank >= nranks) return MCB200_ERR_INVALID;
    *out = nullptr;
    if (!load_nccl()) {
        ctx->set_error(g_nccl_error, __FILE__, __LINE__);
        return MCB200_ERR_NO_DEVICE;
    }
    MCB_CUDA(ctx, cudaSetDevice(ctx->device));
    mcb200_comm* c = new mcb200_comm();
    c->nranks = nranks;
    c->rank = rank;
    c->ctx = ctx;
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    const ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, u, rank);
    if (r != ncclSuccess) {
        ctx->set_error(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r), __FILE__, __LINE__);
        delete c;
        return MCB200_ERR_INTERNAL;
    }
    if (cudaMallocHost(&c->h_counters, sizeof(result_counters_t) * (size_t)nranks) != cudaSuccess) {
        g_nccl.CommDestroy(c->comm);
        delete c;
        return MCB200_ERR_INTERNAL;
    }
    *out = c;
    return 0;
}

void mcb200_comm_destroy(mcb200_comm* c)
{
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    c->ctx->release(c->all_counters);
    c->ctx->release(c->pairs_merged);
    c->ctx->release(c->records_merged);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    delete c;
}

int mcb200_comm_rank(const mcb200_comm* c) { return c ? c->rank : -1; }
int mcb200_comm_size(const mcb200_comm* c) { return c ? c->nranks : 0; }

int mcb200_intersect_stage_sharded(mcb200_ctx* ctx, mcb200_comm* comm, mcb200_mesh* src, mcb200_mesh* cut, double cut_eps,
    const mcb200_soup* soup_in, mcb200_result* res, uint32_t flags)
{
    if (!ctx || !comm || !src || !cut || !soup_in || !res || comm->ctx != ctx) return MCB200_ERR_INVALID;
    if (flags & MCB200_NARROW_LOG_TESTS) {
        ctx->set_error("intersect_stage_sharded: the per-test log is a single-GPU debugging aid", __FILE__, __LINE__);
        return MCB200_ERR_INVALID;
    }
    mcb200_soup* soup = const_cast<mcb200_soup*>(soup_in);
    MCB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int N = comm->nranks;
    // ---- this rank's shard: build both, walk its chunks, narrowphase on its pairs; nothing ordered yet ----
    res->shard_part = (uint32_t)comm->rank;
    res->shard_nparts = (uint32_t)N;
    if (!res->shard_chunk) res->shard_chunk = 4096u;
    MCB_TRY(stage_reserve(ctx, src, cut, soup, res, flags));
    MCB_TRY(mesh_sync_frames(ctx, src, 0.0, cut, cut_eps));
    MCB_TRY(stage_body(ctx, src, cut, cut_eps, soup, res, flags, false, 0, true, false));
    ctx->use_main();
    // ---- everybody's counters to everybody, then ONE host read ----
    MCB_TRY(ctx->reserve(comm->all_counters, sizeof(result_counters_t) * (size_t)N));
    MCB_NCCL(ctx, g_nccl.AllGather(res->counters.p, comm->all_counters.p, sizeof(result_counters_t), ncclChar, comm->comm, ctx->stream));
    MCB_CUDA(ctx, cudaMemcpyAsync(comm->h_counters, comm->all_counters.p, sizeof(result_counters_t) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const result_counters_t* hc = static_cast<const result_counters_t*>(comm->h_counters);
    result_counters_t tot;
    std::memset(&tot, 0, sizeof(tot));
    tot.bad_face = MCB200_NULL;
    std::vector<unsigned long long> npairs((size_t)N), nrec((size_t)N);
    bool overflow = false;
    for (int r = 0; r < N; ++r) {
        overflow = overflow || hc[r].pair_overflow || hc[r].n_pairs > res->cap_pairs || hc[r].n_records > res->cap_records
            || narrow_queue_overflow(res, hc[r]);
        npairs[(size_t)r] = std::min<unsigned long long>(hc[r].n_pairs, res->cap_pairs);
        nrec[(size_t)r] = std::min<unsigned long long>(hc[r].n_records, res->cap_records);
        tot.n_pairs += hc[r].n_pairs;
        tot.n_node_tests += hc[r].n_node_tests;
        tot.n_tests += hc[r].n_tests;
        tot.n_exact += hc[r].n_exact;
        tot.n_records += hc[r].n_records;
        tot.n_queue = std::max(tot.n_queue, hc[r].n_queue);
        tot.n_cross = std::max(tot.n_cross, hc[r].n_cross);
        tot.n_full = std::max(tot.n_full, hc[r].n_full);
        tot.gp_violation |= hc[r].gp_violation;
        tot.bad_face = std::min(tot.bad_face, hc[r].bad_face);
        tot.soup_error |= hc[r].soup_error;
    }
    if (overflow) {
        // some rank's buffers were too small: every rank raises its capacity the same way and reports it; the caller repeats
        // the call (all ranks take this branch together: they all see the same counters)
        size_t need = res->cap_pairs;
        for (int r = 0; r < N; ++r) {
            need = std::max(need, (size_t)hc[r].n_pairs + (size_t)hc[r].n_pairs / 8 + 1024);
            need = std::max(need, (size_t)hc[r].n_records / 2 + 1024);
            if (narrow_queue_overflow(res, hc[r])) need = std::max(need, res->cap_pairs * 2);
        }
        res->cap_pairs = need;
        res->h_valid = false;
        ctx->set_error("sharded stage: a buffer overflowed on some rank; capacities have been raised, run the stage again", __FILE__, __LINE__);
        return MCB200_ERR_CAPACITY;
    }
    // ---- room for the merged lists: the totals may exceed what one shard needed.  Growing the result's buffers now would
    //      drop the shard's own lists (grow-only buffers do not copy), so the capacity is raised and the call repeated —
    //      like any other overflow, and on all ranks alike ----
    const size_t total_pairs = (size_t)tot.n_pairs, total_rec = (size_t)tot.n_records;
    if (total_pairs > res->cap_pairs || total_rec > res->cap_records) {
        res->cap_pairs = std::max(std::max(res->cap_pairs, total_pairs + total_pairs / 8 + 1024), total_rec / 2 + 1024);
        res->h_valid = false;
        ctx->set_error("sharded stage: the merged lists need more room; capacities have been raised, run the stage again", __FILE__, __LINE__);
        return MCB200_ERR_CAPACITY;
    }
    MCB_TRY(ctx->reserve(comm->pairs_merged, sizeof(unsigned long long) * res->cap_pairs));
    MCB_TRY(ctx->reserve(comm->records_merged, sizeof(mcb200_record) * res->cap_records));
    // ---- the exchange: every rank's lists land at their offsets in everybody's merged lists ----
    MCB_NCCL(ctx, g_nccl.GroupStart());
    {
        size_t po = 0, ro = 0;
        for (int r = 0; r < N; ++r) {
            if (npairs[(size_t)r])
                MCB_NCCL(ctx, g_nccl.Broadcast(res->pairs.p, comm->pairs_merged.as<unsigned long long>() + po, (size_t)npairs[(size_t)r] * 8u, ncclChar, r,
                                  comm->comm, ctx->stream));
            if (nrec[(size_t)r])
                MCB_NCCL(ctx, g_nccl.Broadcast(res->records.p, comm->records_merged.as<mcb200_record>() + ro,
                                  (size_t)nrec[(size_t)r] * sizeof(mcb200_record), ncclChar, r, comm->comm, ctx->stream));
            po += (size_t)npairs[(size_t)r];
            ro += (size_t)nrec[(size_t)r];
        }
    }
    MCB_NCCL(ctx, g_nccl.GroupEnd());
    // per-source-face pair counts of all shards (input of the counting order), candidate faces of all shards
    MCB_NCCL(ctx, g_nccl.AllReduce(res->pair_cnt.p, res->pair_cnt.p, (size_t)src->nf, ncclUint32, ncclSum, comm->comm, ctx->stream));
    MCB_NCCL(ctx, g_nccl.AllReduce(res->cand_flag.p, res->cand_flag.p, (size_t)src->nf + cut->nf, ncclUint8, ncclMax, comm->comm, ctx->stream));
    // ---- the merged lists take the place of the shard's lists; counters become the totals; then the usual orders ----
    std::swap(res->pairs, comm->pairs_merged);
    std::swap(res->records, comm->records_merged);
    tot.work_counter = hc[comm->rank].work_counter;
    k_set_counters<<<1, 32, 0, ctx->stream>>>(res->counters.as<result_counters_t>(), tot);
    ctx->launches++;
    MCB_TRY(sort_pairs(ctx, src, cut, res));
    MCB_TRY(narrowphase_planes(ctx, soup, src, cut, res));
    res->records_sorted_valid = false;
    MCB_TRY(narrowphase_sort_records(ctx, res, 3));
    // the shard's own buffers come back (same sizes: they only swapped roles); the ordered lists live in their own buffers
    std::swap(res->pairs, comm->pairs_merged);
    std::swap(res->records, comm->records_merged);
    res->h_valid = false;
    res->shard_part = 0;
    res->shard_nparts = 1;
    return 0;
}

} // extern "C"
