// EXPERIMENT (not part of the product, not linked into libmcut_b200): an MSD-first sort of (24-bit Morton key, face id)
// as planned in DESIGN.md §9 (a).  One unstable bucket scatter on the top 12 bits (no look-back, no ranking votes), then
// one block per bucket sorts its keys in shared memory.  Stand-alone: generates cube-sphere-ordered keys like C2's,
// checks the result against std::sort, prints the event time of every kernel next to a three-pass reference figure the
// caller supplies by running tools/sort_phase.cu.  Measured once on a B200 (1,003,686 keys): histogram 8-10 us, scan 7-8 us,
// scatter 12.3 us, bitonic bucket sort 182 us (result correct) against 3 x 25 us for three one-sweep passes: the scatter
// is worth having, the bitonic network is not - a counting sort per bucket is the next thing to try.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/msd_sort.bin tools/msd_sort.cu
// run:   tools/msd_sort.bin [k = 409] [1 = counting bucket sort (k_local_count, not measured yet) instead of the bitonic one]
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int MSD_BITS = 12, NB = 1 << MSD_BITS, LOW_BITS = 12; // key = [bucket:12][low:12]
constexpr int THREADS = 256, ITEMS = 16, TILE = THREADS * ITEMS;
constexpr int CAP = 4096; // keys a bucket may hold for the shared-memory sort (8 B each: 32 KB)

// ---- K1: bucket histogram (shared-memory aggregation, one global atomic per block and non-empty bucket) ----
__global__ void __launch_bounds__(THREADS) k_hist(const uint32_t* __restrict__ keys, uint32_t n, unsigned* __restrict__ hist)
{
    __shared__ unsigned s_h[NB];
    for (int i = threadIdx.x; i < NB; i += THREADS) s_h[i] = 0u;
    __syncthreads();
    for (uint32_t i = blockIdx.x * THREADS + threadIdx.x; i < n; i += gridDim.x * THREADS) atomicAdd(&s_h[keys[i] >> LOW_BITS], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < NB; i += THREADS)
        if (s_h[i]) atomicAdd(&hist[i], s_h[i]);
}

// ---- K2: exclusive scan of the 4096 counts (one block), cursors for the scatter, size of the largest bucket ----
__global__ void __launch_bounds__(1024) k_scan(const unsigned* __restrict__ hist, unsigned* __restrict__ start /* [NB+1] */,
    unsigned* __restrict__ cursor /* [NB] */, unsigned* __restrict__ largest)
{
    __shared__ unsigned s_w[32];
    const unsigned t = threadIdx.x, lane = t & 31u, w = t >> 5;
    unsigned c[4], sum = 0, mx = 0;
    for (int k = 0; k < 4; ++k) {
        c[k] = hist[t * 4 + k];
        sum += c[k];
        mx = c[k] > mx ? c[k] : mx;
    }
    unsigned inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += v;
    }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    unsigned base = 0;
    for (unsigned i = 0; i < w; ++i) base += s_w[i];
    unsigned run = base + inc - sum;
    for (int k = 0; k < 4; ++k) {
        start[t * 4 + k] = run;
        cursor[t * 4 + k] = run;
        run += c[k];
    }
    if (t == 1023) start[NB] = run;
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned v = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = v > mx ? v : mx;
    }
    if (lane == 0) atomicMax(largest, mx);
}

// ---- K3: unstable scatter into the buckets.  A block counts its tile's keys per bucket in shared memory (the atomicAdd's
// return value is the key's slot among the block's keys of that bucket), reserves one range per non-empty bucket with a
// single global atomicAdd, and writes.  Input in face order is spatially coherent: a tile touches few buckets. ----
__global__ void __launch_bounds__(THREADS) k_scatter(const uint32_t* __restrict__ keys, uint32_t n, unsigned* __restrict__ cursor,
    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out)
{
    __shared__ unsigned s_cnt[NB]; // count, then the reserved base
    const uint32_t tile_base = blockIdx.x * TILE;
    for (int i = threadIdx.x; i < NB; i += THREADS) s_cnt[i] = 0u;
    __syncthreads();
    uint32_t key[ITEMS];
    unsigned short slot[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t i = tile_base + j * THREADS + threadIdx.x;
        key[j] = i < n ? keys[i] : 0xFFFFFFFFu;
        slot[j] = i < n ? (unsigned short)atomicAdd(&s_cnt[key[j] >> LOW_BITS], 1u) : (unsigned short)0;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < NB; b += THREADS) {
        const unsigned c = s_cnt[b];
        if (c) s_cnt[b] = atomicAdd(&cursor[b], c);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t i = tile_base + j * THREADS + threadIdx.x;
        if (i < n) {
            const unsigned dst = s_cnt[key[j] >> LOW_BITS] + slot[j];
            keys_out[dst] = key[j];
            vals_out[dst] = i; // the value of element i is i (face id)
        }
    }
}

// ---- K4: one block per bucket: bitonic sort of (key, value) in shared memory, written to the final arrays ----
__global__ void __launch_bounds__(THREADS) k_local(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
    const unsigned* __restrict__ start, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, unsigned* __restrict__ overflow)
{
    __shared__ unsigned long long s_e[CAP];
    for (unsigned b = blockIdx.x; b < (unsigned)NB; b += gridDim.x) {
        const unsigned lo = start[b], cnt = start[b + 1] - lo;
        if (cnt == 0) continue;
        if (cnt > (unsigned)CAP) { // the fallback (three LSD passes over the whole array) is the caller's job
            if (threadIdx.x == 0) atomicAdd(overflow, 1u);
            continue;
        }
        unsigned p = 1;
        while (p < cnt) p <<= 1;
        for (unsigned i = threadIdx.x; i < p; i += THREADS)
            s_e[i] = i < cnt ? ((unsigned long long)keys_in[lo + i] << 32) | vals_in[lo + i] : ~0ull;
        __syncthreads();
        for (unsigned k = 2; k <= p; k <<= 1)
            for (unsigned j = k >> 1; j > 0; j >>= 1) {
                for (unsigned i = threadIdx.x; i < p; i += THREADS) {
                    const unsigned x = i ^ j;
                    if (x > i) {
                        const unsigned long long a = s_e[i], c = s_e[x];
                        const bool up = (i & k) == 0;
                        if ((a > c) == up) {
                            s_e[i] = c;
                            s_e[x] = a;
                        }
                    }
                }
                __syncthreads();
            }
        for (unsigned i = threadIdx.x; i < cnt; i += THREADS) {
            keys_out[lo + i] = (uint32_t)(s_e[i] >> 32);
            vals_out[lo + i] = (uint32_t)s_e[i];
        }
        __syncthreads();
    }
}

// ---- K4': one block per bucket, counting sort on the 12 low bits.  After the bucket scatter nothing needs to be stable any
// more (keys that agree in all 24 bits may come out in any order), so the slot inside a bin is just the return value of a
// shared-memory atomicAdd: no ballots, no network.  NOT MEASURED YET (written after the GPU budget of round 1 was spent). ----
__global__ void __launch_bounds__(THREADS) k_local_count(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
    const unsigned* __restrict__ start, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out)
{
    constexpr int LB = 1 << LOW_BITS, PER = LB / THREADS; // 4096 bins, 16 per thread
    __shared__ unsigned s_bin[LB];
    __shared__ unsigned s_warp[THREADS / 32];
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    for (unsigned b = blockIdx.x; b < (unsigned)NB; b += gridDim.x) {
        const unsigned lo = start[b], cnt = start[b + 1] - lo;
        if (cnt == 0) continue;
        for (int i = threadIdx.x; i < LB; i += THREADS) s_bin[i] = 0u;
        __syncthreads();
        for (unsigned i = threadIdx.x; i < cnt; i += THREADS) atomicAdd(&s_bin[keys_in[lo + i] & (LB - 1)], 1u);
        __syncthreads();
        // exclusive scan over the bins: thread t owns bins [16 t, 16 t + 16)
        unsigned c[PER], sum = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            c[k] = s_bin[threadIdx.x * PER + k];
            sum += c[k];
        }
        unsigned inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += v;
        }
        if (lane == 31) s_warp[w] = inc;
        __syncthreads();
        unsigned run = inc - sum;
        for (unsigned i = 0; i < w; ++i) run += s_warp[i];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            s_bin[threadIdx.x * PER + k] = run; // the bin's cursor
            run += c[k];
        }
        __syncthreads();
        for (unsigned i = threadIdx.x; i < cnt; i += THREADS) {
            const uint32_t key = keys_in[lo + i];
            const unsigned dst = lo + atomicAdd(&s_bin[key & (LB - 1)], 1u);
            keys_out[dst] = key;
            vals_out[dst] = vals_in[lo + i];
        }
        __syncthreads();
    }
}

static uint32_t spread8(uint32_t v)
{
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

int main(int argc, char** argv)
{
    const int k = argc > 1 ? atoi(argv[1]) : 409; // 6 k^2 keys, cube-sphere order (C2 has 1,002,252 triangles)
    const bool counting = argc > 2 && atoi(argv[2]) != 0; // second argument 1: the counting bucket sort instead of the bitonic one
    std::vector<uint32_t> h;
    h.reserve(6 * (size_t)k * k);
    for (int f = 0; f < 6; ++f)
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) {
                const double a = 2.0 * (i + 0.5) / k - 1.0, b = 2.0 * (j + 0.5) / k - 1.0;
                double p[3];
                const int ax = f >> 1;
                p[ax] = (f & 1) ? 1.0 : -1.0;
                p[(ax + 1) % 3] = a;
                p[(ax + 2) % 3] = b;
                const double len = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
                uint32_t q[3];
                for (int c = 0; c < 3; ++c) {
                    const double u = (p[c] / len + 1.0) * 0.5 * 256.0;
                    q[c] = (uint32_t)std::min(255.0, std::max(0.0, u));
                }
                h.push_back(spread8(q[0]) * 4u + spread8(q[1]) * 2u + spread8(q[2]));
            }
    const uint32_t n = (uint32_t)h.size();
    uint32_t *kin, *kmid, *vmid, *kout, *vout;
    unsigned *hist, *start, *cursor, *misc;
    cudaMalloc(&kin, 4ull * n); cudaMalloc(&kmid, 4ull * n); cudaMalloc(&vmid, 4ull * n); cudaMalloc(&kout, 4ull * n); cudaMalloc(&vout, 4ull * n);
    cudaMalloc(&hist, 4 * NB); cudaMalloc(&start, 4 * (NB + 1)); cudaMalloc(&cursor, 4 * NB); cudaMalloc(&misc, 8);
    cudaMemcpy(kin, h.data(), 4ull * n, cudaMemcpyHostToDevice);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t ev[5];
    for (auto& e : ev) cudaEventCreate(&e);
    const unsigned tiles = (n + TILE - 1) / TILE;
    for (int rep = 0; rep < 5; ++rep) {
        cudaMemset(hist, 0, 4 * NB);
        cudaMemset(misc, 0, 8);
        cudaDeviceSynchronize();
        cudaEventRecord(ev[0]);
        k_hist<<<sms * 4, THREADS>>>(kin, n, hist);
        cudaEventRecord(ev[1]);
        k_scan<<<1, 1024>>>(hist, start, cursor, misc);
        cudaEventRecord(ev[2]);
        k_scatter<<<tiles, THREADS>>>(kin, n, cursor, kmid, vmid);
        cudaEventRecord(ev[3]);
        if (counting)
            k_local_count<<<sms * 8, THREADS>>>(kmid, vmid, start, kout, vout);
        else
            k_local<<<sms * 8, THREADS>>>(kmid, vmid, start, kout, vout, misc + 1);
        cudaEventRecord(ev[4]);
        cudaDeviceSynchronize();
        float t[4];
        for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], ev[i], ev[i + 1]);
        unsigned m[2];
        cudaMemcpy(m, misc, 8, cudaMemcpyDeviceToHost);
        printf("n=%u: hist %.1f us, scan %.1f us, scatter %.1f us, local sort %.1f us, total %.1f us; largest bucket %u, buckets over %d keys: %u (%s)\n",
            n, t[0] * 1e3f, t[1] * 1e3f, t[2] * 1e3f, t[3] * 1e3f, (t[0] + t[1] + t[2] + t[3]) * 1e3f, m[0], CAP, m[1],
            cudaGetErrorString(cudaGetLastError()));
    }
    std::vector<uint32_t> ok(n), ov(n);
    cudaMemcpy(ok.data(), kout, 4ull * n, cudaMemcpyDeviceToHost);
    cudaMemcpy(ov.data(), vout, 4ull * n, cudaMemcpyDeviceToHost);
    std::vector<uint32_t> want = h;
    std::sort(want.begin(), want.end());
    size_t bad_keys = 0, bad_vals = 0;
    std::vector<unsigned char> seen(n, 0);
    for (uint32_t i = 0; i < n; ++i) {
        bad_keys += ok[i] != want[i];
        if (ov[i] >= n || seen[ov[i]] || h[ov[i]] != ok[i]) ++bad_vals;
        else seen[ov[i]] = 1;
    }
    printf("check against std::sort: %zu wrong keys, %zu wrong values\n", bad_keys, bad_vals);
    return (bad_keys || bad_vals) ? 1 : 0;
}
