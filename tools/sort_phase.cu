// Where a one-sweep pass spends its time: per-phase clock64 sums of thread 0 of every block (MCB_SORT_PROFILE), and the
// event time of the same launch without instrumentation is what bench.py reports.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DMCB_SORT_PROFILE -I mcut_b200/csrc -I include -o tools/sort_phase.bin tools/sort_phase.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "internal.h"
#include "radix_sort.cuh"

int main(int argc, char** argv)
{
    const size_t n = argc > 1 ? (size_t)atoll(argv[1]) : 1002252;
    constexpr int ITEMS = rsort::items_for<uint32_t>::value;
    constexpr int TILE = rsort::THREADS * ITEMS;
    const size_t tiles = (n + TILE - 1) / TILE;
    std::vector<uint32_t> h(n);
    uint32_t x = 12345u;
    for (size_t i = 0; i < n; ++i) { x = x * 1664525u + 1013904223u; h[i] = x >> 2; }
    uint32_t *kin, *kout, *vout;
    unsigned *hist, *status, *ctr;
    cudaMalloc(&kin, 4 * n); cudaMalloc(&kout, 4 * n); cudaMalloc(&vout, 4 * n);
    cudaMalloc(&hist, 4 * 256); cudaMalloc(&status, 4 * rsort::status_rows(tiles) * 256); cudaMalloc(&ctr, 4);
    cudaMemcpy(kin, h.data(), 4 * n, cudaMemcpyHostToDevice);
    std::vector<unsigned> hh(256, 0);
    for (size_t i = 0; i < n; ++i) hh[h[i] & 255u]++;
    cudaMemcpy(hist, hh.data(), 4 * 256, cudaMemcpyHostToDevice);
    const rsort::digit_desc dd { 0, 8, 0, 0 };
    constexpr size_t smem = rsort::pass_smem_bytes<uint32_t, uint32_t, true, ITEMS>();
    auto kern = rsort::k_onesweep_pass<uint32_t, uint32_t, true, ITEMS>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const unsigned grid = (unsigned)(tiles < (size_t)sms * 4 ? tiles : (size_t)sms * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 4; ++rep) {
        cudaMemset(status, 0, 4 * rsort::status_rows(tiles) * 256); cudaMemset(ctr, 0, 4);
#ifdef MCB_SORT_PROFILE
        unsigned long long z[8] = {};
        cudaMemcpyToSymbol(rsort::g_sort_phase, z, sizeof(z));
#endif
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        kern<<<grid, rsort::THREADS, smem>>>(kin, kout, nullptr, vout, nullptr, n, dd, 0, hist, status, ctr, 0);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long ph[8] = {};
#ifdef MCB_SORT_PROFILE
        cudaMemcpyFromSymbol(ph, rsort::g_sort_phase, sizeof(ph));
#endif
        printf("n=%zu tiles=%zu grid=%u  %.1f us  (%s)\n", n, tiles, grid, ms * 1000.f, cudaGetErrorString(cudaGetLastError()));
        const char* name[8] = { "load", "rank", "warp offsets+scan", "look-back", "sync after look-back", "reorder", "output", "tile histogram+publish" };
#ifdef MCB_SORT_PROFILE
        for (int i = 0; i < 8; ++i) printf("   %-28s %8.0f cycles per tile\n", name[i], (double)ph[i] / (double)tiles);
#endif
    }
    std::vector<uint32_t> o(n);
    cudaMemcpy(o.data(), kout, 4 * n, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (size_t i = 1; i < n; ++i) bad += (o[i - 1] & 255u) > (o[i] & 255u);
    printf("out-of-order digits: %zu\n", bad);
    return 0;
}
