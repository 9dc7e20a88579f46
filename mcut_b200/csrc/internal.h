// mcut_b200/csrc/internal.h — launchers shared between the kernel files and the C-ABI (api.cu).
#pragma once

#include "common.cuh"

// lbvh.cu
// puts the frames of the given meshes (build frame with `eps`, narrowphase frame) into their device slots when they changed;
// enqueued on ctx->cur — call it before forking lanes
int mesh_sync_frames(mcb200_ctx* ctx, mcb200_mesh* a, double eps_a, mcb200_mesh* b = nullptr, double eps_b = 0.0);
int lbvh_reserve(mcb200_ctx* ctx, mcb200_mesh* mesh);
int lbvh_build(mcb200_ctx* ctx, mcb200_mesh* mesh, double eps);
// traverse.cu
int traverse_reserve(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_result* res);
int result_reset_counters(mcb200_ctx* ctx, mcb200_result* res);
int traverse_pairs(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_result* res);
int sort_pairs_reserve(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_result* res);
int sort_pairs(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_result* res);
int sort_pairs_fallback(mcb200_ctx* ctx, mcb200_result* res);
// narrowphase.cu
int narrowphase_reserve(mcb200_ctx* ctx, const mcb200_soup* soup, mcb200_result* res, uint32_t flags);
int narrowphase_run(mcb200_ctx* ctx, const mcb200_soup* soup, const mcb200_mesh* src, const mcb200_mesh* cut,
    mcb200_result* res, uint32_t flags);
// internal flag (never set by callers of the C-ABI): this run is one shard of a multi-GPU dispatch — no plane rows, no orders
#define MCB200_NARROW_INTERNAL_PARTIAL 0x40000000u
// The radix path of the registry order (ten launches that return at once unless there are more than 16384 records) is not
// enqueued; fetch_counters runs it when the count calls for it.  Set by the host-array entry points, whose callers cannot
// look at a record before they have asked for the counts.
#define MCB200_NARROW_INTERNAL_LAZY_RADIX 0x20000000u
int narrowphase_planes(mcb200_ctx* ctx, const mcb200_soup* soup, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_result* res);
int narrowphase_prezero(mcb200_ctx* ctx, const mcb200_soup* soup, mcb200_result* res);
int soup_face_vtx_device(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_soup* soup);
int narrowphase_sort_records(mcb200_ctx* ctx, mcb200_result* res, int part = 3);
int narrowphase_sort_tests(mcb200_ctx* ctx, mcb200_result* res, int part = 3);
int narrowphase_finish_record_order(mcb200_ctx* ctx, mcb200_result* res); // the deferred radix path, if the count needs it
// api.cu: the stage body shared by the stage entry points (order = false: a shard — pairs and records stay unordered)
int stage_reserve(mcb200_ctx* ctx, mcb200_mesh* src, mcb200_mesh* cut, const mcb200_soup* soup, mcb200_result* res, uint32_t flags);
int stage_body(mcb200_ctx* ctx, mcb200_mesh* src, mcb200_mesh* cut, double cut_eps, mcb200_soup* soup, mcb200_result* res, uint32_t flags,
    bool wait_uploads, int number_soup, bool interleave, bool order = true);
int fetch_counters(mcb200_ctx* ctx, mcb200_result* res);
bool narrow_queue_overflow(const mcb200_result* res, const result_counters_t& h);
// traverse.cu: off[i] = cnt[0] + ... + cnt[i-1], off[n] = total; both arrays padded to a multiple of 8 entries (+8) with zeros behind
// n; `tile` needs n / 2048 + 2 words.  cnt is zeroed on the way (it is the pair order's scatter cursor).
int exclusive_scan_u32(mcb200_ctx* ctx, unsigned* cnt, uint32_t n, unsigned* tile, unsigned* off, result_counters_t* counters);
// soup_ids.cu
int soup_number_reserve(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_soup* soup);
int soup_number_device(mcb200_ctx* ctx, const mcb200_mesh* src, const mcb200_mesh* cut, mcb200_soup* soup, result_counters_t* counters);
// validate.cu
int mesh_validate_run(mcb200_ctx* ctx, mcb200_mesh* mesh);
int mesh_winding_run(mcb200_ctx* ctx, mcb200_mesh* mesh, const double query[3]);
int mesh_vertex_position(mcb200_ctx* ctx, mcb200_mesh* mesh, uint32_t v, double out[3]);
// host_logic.cpp
int host_soup_ids(uint32_t nsv, const uint32_t* src_off, const uint32_t* src_vtx, uint32_t nsf, const uint32_t* cut_off,
    const uint32_t* cut_vtx, uint32_t ncf, uint32_t* face_vtx, uint32_t* face_edge, uint32_t* edge_v, uint32_t* edge_f,
    uint32_t* ne);
