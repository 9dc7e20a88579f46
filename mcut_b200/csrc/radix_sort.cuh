// mcut_b200/csrc/radix_sort.cuh — hand-written one-sweep LSD radix sort (no CUB/Thrust).
//
// Replaces the reference's single-threaded std::sort of (face, Morton code) pairs (source/bvh.cpp:437-442)
// and provides the canonical ordering of candidate pairs and registry records.
//
// Structure ("onesweep"): ONE histogram kernel reads the keys once and produces the digit histograms of
// every pass; each pass is then a single sweep: a persistent grid takes 256-thread tiles by ticket, ranks
// the tile's keys per digit with per-lane byte counters in shared memory, publishes the tile's digit counts and resolves its
// global offsets by decoupled look-back over the previous tiles (no second read of the keys, no separate
// scan kernel), reorders the tile through shared memory and writes coalesced runs.
// The sort is stable, so LSD passes compose.  Element counts may live on the device (d_n), so a sort can
// follow the kernel that produced its input without a host round trip.
#pragma once

#include "common.cuh"

namespace rsort {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int MAX_PASSES = 8;

constexpr unsigned FLAG_MASK = 0xC0000000u;
constexpr unsigned FLAG_AGG = 0x40000000u;
constexpr unsigned FLAG_INCL = 0x80000000u;
constexpr unsigned VALUE_MASK = 0x3FFFFFFFu;

// Status rows of one pass: one row of RADIX words per tile (the tile's digit counts) followed by one row per GROUP of
// LOOK_GROUP consecutive tiles (the group's digit counts) — see the look-back in k_onesweep_pass.
constexpr int LOOK_GROUP = 16;
__host__ __device__ __forceinline__ size_t status_rows(size_t tiles) { return tiles + (tiles + LOOK_GROUP - 1) / LOOK_GROUP; }

// One radix digit: `bits` bits of the key starting at `shift`, continued by `bits2` bits starting at `shift2` (bits2 = 0
// for an ordinary digit).  A digit may straddle the two bit ranges of a packed key such as (src face << 32 | cut face):
// the digits then tile the CONCATENATION of the ranges, so 20 + 20 significant bits cost five 8-bit passes, not six.
struct digit_desc {
    int shift, bits, shift2, bits2;
};

struct pass_desc {
    digit_desc d[MAX_PASSES];
    int npasses;
};

// 8-bit chunks over the concatenation of the bit ranges [lo0, hi0) and [lo1, hi1) (second range optional: hi1 <= lo1)
static inline pass_desc make_passes(int lo0, int hi0, int lo1 = 0, int hi1 = 0)
{
    pass_desc p;
    p.npasses = 0;
    const int n0 = hi0 > lo0 ? hi0 - lo0 : 0, n1 = hi1 > lo1 ? hi1 - lo1 : 0;
    for (int pos = 0; pos < n0 + n1 && p.npasses < MAX_PASSES; pos += RADIX_BITS) {
        digit_desc& d = p.d[p.npasses++];
        const int len = (n0 + n1 - pos < RADIX_BITS) ? n0 + n1 - pos : RADIX_BITS;
        if (pos >= n0) { // entirely in the second range
            d.shift = lo1 + (pos - n0);
            d.bits = len;
            d.shift2 = 0;
            d.bits2 = 0;
        } else {
            d.shift = lo0 + pos;
            d.bits = (n0 - pos < len) ? n0 - pos : len;
            d.shift2 = lo1;
            d.bits2 = len - d.bits;
        }
    }
    for (int i = p.npasses; i < MAX_PASSES; ++i) p.d[i] = digit_desc { 0, 0, 0, 0 };
    return p;
}

template <typename KeyT> __device__ __forceinline__ unsigned digit_of(KeyT k, const digit_desc& d)
{
    const unsigned lo = (unsigned)(k >> d.shift) & ((1u << d.bits) - 1u);
    const unsigned hi = d.bits2 ? ((unsigned)(k >> d.shift2) & ((1u << d.bits2) - 1u)) << d.bits : 0u;
    return lo | hi;
}

// The same digit with the loop-invariant parts hoisted (masks in place, second shift already reduced by `bits`):
// two shifts and one (a & m1) | (b & m2).
struct digit_fast {
    int shift, shift2;
    unsigned m1, m2;
    __device__ __forceinline__ explicit digit_fast(const digit_desc& d)
        : shift(d.shift), shift2(d.bits2 ? d.shift2 - d.bits : 0), m1((1u << d.bits) - 1u), m2(d.bits2 ? ((1u << d.bits2) - 1u) << d.bits : 0u)
    {
    }
    template <typename KeyT> __device__ __forceinline__ unsigned operator()(KeyT k) const
    {
        return ((unsigned)(k >> shift) & m1) | ((unsigned)(k >> shift2) & m2);
    }
};

__device__ __forceinline__ size_t resolve_n(const unsigned long long* d_n, size_t n_max)
{
    if (!d_n) return n_max;
    const unsigned long long v = *d_n;
    return v < n_max ? (size_t)v : n_max;
}

// ---- histogram of every pass in one read of the keys -------------------------------------------------------------
// Also clears the look-back status words of the tiles that will exist (a function of the LIVE count, which may
// only be known on the device), so no capacity-sized memset is needed.
template <typename KeyT>
__global__ void __launch_bounds__(THREADS) k_histogram(const KeyT* __restrict__ keys, const unsigned long long* d_n,
    size_t n_max, pass_desc pd, int tile_items, unsigned* __restrict__ hist /* [npasses][RADIX] */,
    unsigned* __restrict__ status, size_t skip_le, const unsigned* d_flag, unsigned flag_le)
{
    pdl_prologue();
    __shared__ unsigned s_hist[MAX_PASSES * RADIX];
    const size_t n = resolve_n(d_n, n_max);
    if (n <= skip_le) return; // a small-n kernel already produced the sorted output
    if (d_flag && *d_flag <= flag_le) return; // another method took this input (decided on the device)
    for (int i = threadIdx.x; i < pd.npasses * RADIX; i += THREADS) s_hist[i] = 0;
    __syncthreads();
    {
        const size_t tiles = (n + tile_items - 1) / tile_items;
        const size_t words = status_rows(tiles) * RADIX * (size_t)pd.npasses;
        for (size_t i = (size_t)blockIdx.x * THREADS + threadIdx.x; i < words; i += (size_t)gridDim.x * THREADS) status[i] = 0u;
    }
    for (size_t i = (size_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * THREADS) {
        const KeyT k = keys[i];
        for (int p = 0; p < pd.npasses; ++p) atomicAdd(&s_hist[p * RADIX + digit_of(k, pd.d[p])], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < pd.npasses * RADIX; i += THREADS) {
        const unsigned c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// block-wide exclusive scan of one value per thread (256 threads)
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* s_warp /* [WARPS] */, unsigned* total)
{
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    unsigned base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < WARPS; ++i) {
        const unsigned c = s_warp[i];
        if ((unsigned)i < w) base += c;
        tot += c;
    }
    __syncthreads();
    if (total) *total = tot;
    return base + inc - v;
}

// Per-phase clocks of a pass (thread 0 of every block), compiled in only by tools/sort_phase.cu (-DMCB_SORT_PROFILE).
#ifdef MCB_SORT_PROFILE
__device__ unsigned long long g_sort_phase[8];
#define MCB_PHASE(i)                                                                                                            \
    do {                                                                                                                        \
        if (threadIdx.x == 0) {                                                                                                 \
            const long long now_ = clock64();                                                                                   \
            atomicAdd(&g_sort_phase[i], (unsigned long long)(now_ - phase_t_));                                                 \
            phase_t_ = now_;                                                                                                    \
        }                                                                                                                       \
    } while (0)
#else
#define MCB_PHASE(i)
#endif

// ---- one pass ------------------------------------------------------------------------------------------------
// Ranking.  A warp owns 32 * ITEMS consecutive keys of the tile, striped: item j of a lane is key j * 32 + lane of the
// warp's chunk, so "input order" inside the chunk is (j, lane).  For item j the lanes find their peers (same digit) with
// one __ballot_sync per digit bit; the highest peer adds the group's size to the warp's running count of that digit
// (shared-memory atomicAdd returning the old value) and hands the old value to its peers by shuffle:
//     rank = keys with this digit in earlier items of the warp (old) + peers in lower lanes.
// After the last item the warp's counters ARE its digit histogram.  Measured on sm_100 (tools/mb_match.cu, 32 warps on
// an SM): this step costs ~31 SM cycles per warp-item, __match_any_sync-based ranking ~135, and the previous scheme here
// (per-lane byte counters in shared memory + a multiply-based prefix over every digit row) ~76 at the IPC it reached -
// with 64 KB of counters per block (2 blocks per SM) against 8 KB now.
// vals_in == nullptr with HAS_VALS: the value of element i is i (saves materialising an iota array).
template <typename KeyT, typename ValT, bool HAS_VALS, int ITEMS> constexpr size_t pass_smem_bytes()
{
    return (size_t)THREADS * ITEMS * (sizeof(KeyT) + (HAS_VALS ? sizeof(ValT) : 0)) + (size_t)WARPS * RADIX * sizeof(unsigned);
}

template <typename KeyT, typename ValT, bool HAS_VALS, int ITEMS>
__global__ void __launch_bounds__(THREADS, 3) k_onesweep_pass(const KeyT* __restrict__ keys_in, KeyT* __restrict__ keys_out,
    const ValT* __restrict__ vals_in, ValT* __restrict__ vals_out, const unsigned long long* d_n, size_t n_max,
    digit_desc dd, int pass_index, const unsigned* __restrict__ hist /* [RADIX] of this pass */,
    unsigned* status_all /* [npasses][tiles(n)][RADIX] */, unsigned* tile_counter, size_t skip_le, const unsigned* d_flag,
    unsigned flag_le)
{
    pdl_prologue();
    static_assert(THREADS == RADIX, "one thread per digit in the per-digit steps");
    constexpr int TILE = THREADS * ITEMS;
    constexpr size_t STAGING = (size_t)TILE * (sizeof(KeyT) + (HAS_VALS ? sizeof(ValT) : 0));
    extern __shared__ __align__(16) unsigned char s_dyn[];
    // region 0: key/value staging of the reorder; then the warps' digit counters
    KeyT* s_keys = reinterpret_cast<KeyT*>(s_dyn);
    ValT* s_vals = reinterpret_cast<ValT*>(s_dyn + (size_t)TILE * sizeof(KeyT));
    unsigned* s_wcnt = reinterpret_cast<unsigned*>(s_dyn + STAGING); // [WARPS][RADIX]
    __shared__ unsigned s_tile_excl[RADIX];
    __shared__ unsigned s_global_off[RADIX];
    __shared__ unsigned s_tile_hist[RADIX];
    __shared__ unsigned s_scan[WARPS];
    __shared__ unsigned s_tile;

    const size_t n = resolve_n(d_n, n_max);
    if (n <= skip_le) return;
    if (d_flag && *d_flag <= flag_le) return;
    const unsigned num_tiles = (unsigned)((n + TILE - 1) / TILE);
    unsigned* status = status_all + (size_t)pass_index * status_rows(num_tiles) * RADIX;
    unsigned* gstatus = status + (size_t)num_tiles * RADIX; // group rows
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    unsigned* my_cnt = s_wcnt + w * RADIX; // this warp's counters
    const digit_fast digit(dd);
    const int nbits = dd.bits + dd.bits2;
    const unsigned lanes_below = (1u << lane) - 1u;

    // exclusive scan of the pass histogram: where each digit's bucket starts in the output
    unsigned my_bucket_base = block_exclusive_scan(hist[threadIdx.x], s_scan, nullptr);

    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
        __syncthreads();
        const unsigned tile = s_tile;
        if (tile >= num_tiles) break;
        const size_t tile_base = (size_t)tile * TILE;
        const unsigned tile_n = (unsigned)((n - tile_base < (size_t)TILE) ? (n - tile_base) : (size_t)TILE);
        const bool full = tile_n == (unsigned)TILE;
        const unsigned local0 = w * (32 * ITEMS) + lane; // item j: local0 + 32 j
#ifdef MCB_SORT_PROFILE
        long long phase_t_ = clock64();
#endif

        // ---- load (coalesced: a warp reads 32 consecutive keys per item) ----
        KeyT key[ITEMS];
        ValT val[HAS_VALS ? ITEMS : 1];
        unsigned short rank[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const unsigned local = local0 + 32u * j;
            const bool valid = full || local < tile_n;
            key[j] = valid ? keys_in[tile_base + local] : (KeyT)0;
            if (HAS_VALS) val[j] = valid ? (vals_in ? vals_in[tile_base + local] : (ValT)(tile_base + local)) : (ValT)0;
        }

#ifdef MCB_SORT_PROFILE
        {
            unsigned acc_ = 0;
            for (int j = 0; j < ITEMS; ++j) acc_ += (unsigned)key[j];
            if (acc_ == 0x12345u) s_scan[0] = acc_; // forces the loads to land
            __syncthreads();
        }
        MCB_PHASE(0);
#endif
        // ---- the tile's digit counts first: published before the ranking starts, so that the tiles that wait for them in
        //      their look-back wait for a few hundred cycles of shared-memory atomics instead of a whole ranking phase ----
        s_tile_hist[threadIdx.x] = 0u;
#pragma unroll
        for (int q = 0; q < RADIX / 32; ++q) my_cnt[q * 32 + lane] = 0u;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ITEMS; ++j)
            if (full || local0 + 32u * j < tile_n) atomicAdd(&s_tile_hist[digit(key[j])], 1u);
        __syncthreads();
        const unsigned total = s_tile_hist[threadIdx.x];
        {
            volatile unsigned* st = status + (size_t)tile * RADIX + threadIdx.x;
            *st = FLAG_AGG | total;
        }
        MCB_PHASE(7);

        const unsigned excl_in_tile = block_exclusive_scan(total, s_scan, nullptr);
        s_tile_excl[threadIdx.x] = excl_in_tile;

        // ---- two-level look-back: sum of this digit's counts over all previous tiles.  It runs BEFORE the ranking: what it
        //      waits for is other tiles' histograms, which are a few hundred cycles old by now, and the tiles that wait for
        //      THIS tile's group total do not have to sit through its ranking phase ----
        // At these sizes every tile of a pass is resident at once and publishes at about the same moment, so a chained
        // look-back would walk all the way back (tiles/16 dependent L2 round trips).  Instead tiles form groups of 16:
        // a tile sums the counts of the earlier tiles of its own group (one round of independent loads); the LAST tile
        // of a group also publishes the group total; then the totals of all earlier groups are summed, 16 per round.
        // Two or three dependent round trips whatever the tile count.  Every awaited tile holds an earlier ticket, so
        // it is running (or done) and never waits for a later one: no deadlock under any block schedule.
        {
            const unsigned d = threadIdx.x;
            const unsigned g = tile / LOOK_GROUP, j = tile % LOOK_GROUP;
            unsigned prefix = 0;
            {
                unsigned sv[LOOK_GROUP - 1];
                bool all;
                do {
                    all = true;
#pragma unroll
                    for (int k = 0; k < LOOK_GROUP - 1; ++k) {
                        unsigned cur = FLAG_AGG;
                        if ((unsigned)k < j) cur = *reinterpret_cast<volatile unsigned*>(status + (size_t)(g * LOOK_GROUP + k) * RADIX + d);
                        sv[k] = cur;
                        all = all && ((cur & FLAG_MASK) != 0u);
                    }
                } while (!all);
#pragma unroll
                for (int k = 0; k < LOOK_GROUP - 1; ++k) prefix += sv[k] & VALUE_MASK;
            }
            if (j == LOOK_GROUP - 1) {
                volatile unsigned* gp = gstatus + (size_t)g * RADIX + d;
                *gp = FLAG_AGG | (prefix + total);
            }
            // earlier groups: decoupled look-back over GROUP rows, 16 polled per step, nearest first.  A group row starts as
            // the group's total (FLAG_AGG) and is upgraded by its last tile to the inclusive prefix (FLAG_INCL) once that tile
            // knows its own prefix — with many waves of tiles the walk stops at the first inclusive row it meets.
            unsigned before = 0;
            if (g > 0) {
                int t = (int)g - 1;
                bool done = false;
                while (!done) {
                    unsigned sv[LOOK_GROUP];
#pragma unroll
                    for (int k = 0; k < LOOK_GROUP; ++k) {
                        const int gg = t - k;
                        unsigned cur = FLAG_INCL + 0u; // before group 0: an inclusive prefix of zero
                        if (gg >= 0) cur = *reinterpret_cast<volatile unsigned*>(gstatus + (size_t)gg * RADIX + d);
                        sv[k] = cur;
                    }
                    int used = 0;
#pragma unroll
                    for (int k = 0; k < LOOK_GROUP; ++k) {
                        if (done || used != k) continue;
                        const unsigned sk = sv[k];
                        if ((sk & FLAG_MASK) == 0) continue; // not published yet: poll again from here
                        before += sk & VALUE_MASK;
                        used = k + 1;
                        if (sk & FLAG_INCL) done = true;
                    }
                    t -= used;
                }
            }
            if (j == LOOK_GROUP - 1) {
                volatile unsigned* gp = gstatus + (size_t)g * RADIX + d;
                *gp = FLAG_INCL | (before + prefix + total);
            }
            prefix += before;
            s_global_off[d] = my_bucket_base + prefix;
        }
        MCB_PHASE(3);

        // ---- rank ----
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const bool valid = full || local0 + 32u * j < tile_n;
            const unsigned d = digit(key[j]);
            unsigned peers = full ? 0xffffffffu : __ballot_sync(0xffffffffu, valid);
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                if (b < nbits) { // uniform
                    const bool bit = (d >> b) & 1u;
                    const unsigned bal = __ballot_sync(0xffffffffu, bit);
                    peers &= bit ? bal : ~bal;
                }
            }
            // a valid lane is its own peer, so peers != 0 there; lanes past the end take no part
            const int leader = 31 - __clz(peers);
            unsigned before = 0;
            if (valid && (int)lane == leader) before = atomicAdd(&my_cnt[d], (unsigned)__popc(peers));
            before = __shfl_sync(0xffffffffu, before, leader & 31);
            rank[j] = (unsigned short)(before + __popc(peers & lanes_below));
        }
        __syncthreads();
        MCB_PHASE(1);

        // ---- per-digit: warp counts -> exclusive warp offsets ----
        {
            const unsigned d = threadIdx.x;
            unsigned run = 0;
#pragma unroll
            for (int i = 0; i < WARPS; ++i) {
                const unsigned c = s_wcnt[i * RADIX + d];
                s_wcnt[i * RADIX + d] = run;
                run += c;
            }
        }
        MCB_PHASE(2);
        __syncthreads(); // warp offsets and tile offsets are in place
        MCB_PHASE(4);

        // ---- reorder the tile through shared memory ----
        // (all the offset loads first: the compiler cannot prove that the staging stores leave the offset tables alone)
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const unsigned d = digit(key[j]);
            rank[j] = (unsigned short)(s_tile_excl[d] + my_cnt[d] + rank[j]); // position in the tile, < TILE <= 4096
        }
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            if (full || local0 + 32u * j < tile_n) {
                s_keys[rank[j]] = key[j];
                if (HAS_VALS) s_vals[rank[j]] = val[j];
            }
        }
        __syncthreads();
        MCB_PHASE(5);
        // ---- coalesced runs out ----
        for (unsigned i = threadIdx.x; i < tile_n; i += THREADS) {
            const KeyT k = s_keys[i];
            const unsigned d = digit(k);
            const size_t dst = (size_t)s_global_off[d] + (i - s_tile_excl[d]);
            keys_out[dst] = k;
            if (HAS_VALS) vals_out[dst] = s_vals[i];
        }
        __syncthreads();
        MCB_PHASE(6);
    }
}

template <typename KeyT> struct items_for {
    static constexpr int value = sizeof(KeyT) == 4 ? 16 : 8;
};
// 32-bit keys: inputs of up to about a million keys are sorted in tiles of half the size.  A pass over so few keys is bound
// by the latency of one tile's load -> rank -> look-back -> scatter chain, not by bandwidth: twice the blocks per SM hide it.
inline size_t& small_tile_max()
{
    static size_t v = [] {
        const char* e = std::getenv("MCB200_SORT_SMALL_TILE_MAX");
        return e ? (size_t)std::atoll(e) : (size_t)0; // measured on C2 (1M keys): 0.331 ms with half tiles, 0.322 ms without - off by default
    }();
    return v;
}
template <typename KeyT> inline int items_rt(size_t n_max)
{
    return (sizeof(KeyT) == 4 && n_max <= small_tile_max()) ? 8 : items_for<KeyT>::value;
}

// scratch sizes a sort of n_max keys needs in the CURRENT scratch set (so callers can reserve before forking lanes)
template <typename KeyT> int reserve_scratch(mcb200_ctx* ctx, size_t n_max, int npasses, bool need_alt_keys, bool need_alt_vals)
{
    const int TILE = THREADS * items_rt<KeyT>(n_max);
    const size_t tiles = (n_max + TILE - 1) / TILE;
    mcb200_ctx::sort_scratch_t& sc = ctx->sc();
    MCB_TRY(ctx->reserve(sc.hist, sizeof(unsigned) * MAX_PASSES * RADIX));
    MCB_TRY(ctx->reserve(sc.status, sizeof(unsigned) * (size_t)(npasses ? npasses : 1) * status_rows(tiles ? tiles : 1) * RADIX));
    MCB_TRY(ctx->reserve(sc.tilectr, sizeof(unsigned) * MAX_PASSES));
    if (need_alt_keys) MCB_TRY(ctx->reserve(sc.keys_alt, sizeof(KeyT) * (n_max ? n_max : 1)));
    if (need_alt_vals) MCB_TRY(ctx->reserve(sc.vals_alt, sizeof(uint32_t) * (n_max ? n_max : 1)));
    return 0;
}

// A sort is: prepare (clear histogram + tile tickets), histogram of every pass (k_histogram, or fused into the kernel
// that produces the keys — see lbvh.cu: k_morton), then one sweep per pass.
inline int sort_prepare(mcb200_ctx* ctx)
{
    mcb200_ctx::sort_scratch_t& sc = ctx->sc();
    fill_list_t fl {};
    fl.add(sc.hist.p, (size_t)MAX_PASSES * RADIX, 0u);
    fl.add(sc.tilectr.p, MAX_PASSES, 0u);
    MCB_LAUNCH(ctx, k_fill, 8, 256, 0, fl);
    return 0;
}

// Pass 0 reads keys_in/vals_in, the passes then ping-pong between the (a) and (b) buffers: in -> a -> b -> a ...
// (b may alias the input when it may be overwritten).  *keys_out / *vals_out receive the buffer the sorted data ended up
// in.  vals_in == nullptr with HAS_VALS: the value of element i is i.
template <typename KeyT, typename ValT, bool HAS_VALS>
int sort_passes(mcb200_ctx* ctx, const KeyT* keys_in, KeyT* keys_a, KeyT* keys_b, const ValT* vals_in, ValT* vals_a, ValT* vals_b,
    const unsigned long long* d_n, size_t n_max, const pass_desc& pd, KeyT** keys_out, ValT** vals_out, size_t skip_le = 0,
    const unsigned* d_flag = nullptr, unsigned flag_le = 0);

template <typename KeyT, typename ValT, bool HAS_VALS, int ITEMS>
int sort_passes_impl(mcb200_ctx* ctx, const KeyT* keys_in, KeyT* keys_a, KeyT* keys_b, const ValT* vals_in, ValT* vals_a, ValT* vals_b,
    const unsigned long long* d_n, size_t n_max, const pass_desc& pd, KeyT** keys_out, ValT** vals_out, size_t skip_le,
    const unsigned* d_flag, unsigned flag_le)
{
    constexpr int TILE = THREADS * ITEMS;
    const size_t tiles = (n_max + TILE - 1) / TILE;
    mcb200_ctx::sort_scratch_t& sc = ctx->sc();
    const char* pname = sizeof(KeyT) == 4 ? "onesweep_pass_u32_kv" : (HAS_VALS ? "onesweep_pass_u64_kv" : "onesweep_pass_u64_k");
    // persistent grid: enough resident blocks to fill the machine, never more than tiles
    const unsigned pgrid = (unsigned)((tiles < (size_t)ctx->num_sms * 4) ? tiles : (size_t)ctx->num_sms * 4);
    const KeyT* kin = keys_in;
    const ValT* vin = vals_in;
    constexpr size_t smem = pass_smem_bytes<KeyT, ValT, HAS_VALS, ITEMS>();
    {
        // the opt-in is per function and per device; a context is bound to one device
        const int which = (sizeof(KeyT) == 8 ? 2 : 0) + (HAS_VALS ? 1 : 0) + (ITEMS != items_for<KeyT>::value ? 4 : 0);
        if (!ctx->sort_smem_opt_in[which]) {
            MCB_CUDA(ctx, cudaFuncSetAttribute(k_onesweep_pass<KeyT, ValT, HAS_VALS, ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                (int)smem));
            ctx->sort_smem_opt_in[which] = true;
        }
    }
    for (int p = 0; p < pd.npasses; ++p) {
        KeyT* kout = (p & 1) ? keys_b : keys_a;
        ValT* vout = (p & 1) ? vals_b : vals_a;
        MCB_LAUNCH_NAMED(ctx, pname, (k_onesweep_pass<KeyT, ValT, HAS_VALS, ITEMS>), pgrid, THREADS, smem, kin, kout, vin, vout, d_n, n_max,
            pd.d[p], p, sc.hist.as<unsigned>() + p * RADIX, sc.status.as<unsigned>(), sc.tilectr.as<unsigned>() + p, skip_le, d_flag, flag_le);
        kin = kout;
        vin = vout;
    }
    if (keys_out) *keys_out = const_cast<KeyT*>(kin);
    if (vals_out) *vals_out = const_cast<ValT*>(vin);
    return 0;
}

template <typename KeyT, typename ValT, bool HAS_VALS>
int sort_passes(mcb200_ctx* ctx, const KeyT* keys_in, KeyT* keys_a, KeyT* keys_b, const ValT* vals_in, ValT* vals_a, ValT* vals_b,
    const unsigned long long* d_n, size_t n_max, const pass_desc& pd, KeyT** keys_out, ValT** vals_out, size_t skip_le,
    const unsigned* d_flag, unsigned flag_le)
{
    if (sizeof(KeyT) == 4 && items_rt<KeyT>(n_max) == 8)
        return sort_passes_impl<KeyT, ValT, HAS_VALS, 8>(ctx, keys_in, keys_a, keys_b, vals_in, vals_a, vals_b, d_n, n_max, pd, keys_out,
            vals_out, skip_le, d_flag, flag_le);
    return sort_passes_impl<KeyT, ValT, HAS_VALS, items_for<KeyT>::value>(ctx, keys_in, keys_a, keys_b, vals_in, vals_a, vals_b, d_n, n_max,
        pd, keys_out, vals_out, skip_le, d_flag, flag_le);
}

template <typename KeyT, typename ValT, bool HAS_VALS>
int sort(mcb200_ctx* ctx, const KeyT* keys_in, KeyT* keys_a, KeyT* keys_b, const ValT* vals_in, ValT* vals_a, ValT* vals_b,
    const unsigned long long* d_n, size_t n_max, const pass_desc& pd, KeyT** keys_out, ValT** vals_out, size_t skip_le = 0,
    const unsigned* d_flag = nullptr, unsigned flag_le = 0)
{
    const int TILE = THREADS * items_rt<KeyT>(n_max);
    if (keys_out) *keys_out = const_cast<KeyT*>(keys_in);
    if (vals_out) *vals_out = const_cast<ValT*>(vals_in);
    if (n_max == 0 || pd.npasses == 0) return 0;
    const size_t tiles = (n_max + TILE - 1) / TILE;
    MCB_TRY((reserve_scratch<KeyT>(ctx, n_max, pd.npasses, false, false)));
    MCB_TRY(sort_prepare(ctx));
    mcb200_ctx::sort_scratch_t& sc = ctx->sc();
    const unsigned hgrid = (unsigned)((tiles < (size_t)ctx->num_sms * 4) ? tiles : (size_t)ctx->num_sms * 4);
    const char* hname = sizeof(KeyT) == 4 ? "sort_histogram_u32" : "sort_histogram_u64";
    MCB_LAUNCH_NAMED(ctx, hname, (k_histogram<KeyT>), hgrid, THREADS, 0, keys_in, d_n, n_max, pd, TILE, sc.hist.as<unsigned>(),
        sc.status.as<unsigned>(), skip_le, d_flag, flag_le);
    return sort_passes<KeyT, ValT, HAS_VALS>(ctx, keys_in, keys_a, keys_b, vals_in, vals_a, vals_b, d_n, n_max, pd, keys_out, vals_out,
        skip_le, d_flag, flag_le);
}

} // namespace rsort
