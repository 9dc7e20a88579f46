// micro-benchmark: cost of MATCH.ANY vs an 8-ballot emulation vs smem atomics, per warp, dependent and independent
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned* out, long long* cyc, int mode)
{
    unsigned lane = threadIdx.x & 31;
    unsigned x = (lane * 2654435761u + blockIdx.x) & 255u;
    __shared__ unsigned sh[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    unsigned acc = 0;
    long long t0 = clock64();
    if (mode == 0) { // dependent match chain
        for (int i = 0; i < 256; ++i) { unsigned p = __match_any_sync(0xffffffffu, x); x = (x + __popc(p) + i) & 255u; acc += p; }
    } else if (mode == 1) { // 4 independent matches per iteration
        for (int i = 0; i < 64; ++i) {
            unsigned p0 = __match_any_sync(0xffffffffu, x), p1 = __match_any_sync(0xffffffffu, x ^ 1u), p2 = __match_any_sync(0xffffffffu, x ^ 2u), p3 = __match_any_sync(0xffffffffu, x ^ 7u);
            acc += p0 + p1 + p2 + p3; x = (x + (acc & 3) + i) & 255u;
        }
    } else if (mode == 2) { // ballot emulation, dependent
        for (int i = 0; i < 256; ++i) {
            unsigned p = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 8; ++b) { unsigned bal = __ballot_sync(0xffffffffu, (x >> b) & 1u); p &= ((x >> b) & 1u) ? bal : ~bal; }
            x = (x + __popc(p) + i) & 255u; acc += p;
        }
    } else if (mode == 3) { // smem atomicAdd returning old, dependent
        for (int i = 0; i < 256; ++i) { unsigned o = atomicAdd(&sh[x], 1u); x = (x + o + i) & 255u; acc += o; }
    } else if (mode == 4) { // full rank step as in the sort
        unsigned lt = (1u << lane) - 1u;
        for (int i = 0; i < 256; ++i) {
            unsigned p = __match_any_sync(0xffffffffu, x); int leader = __ffs(p) - 1; unsigned old = 0;
            if ((int)lane == leader) { old = sh[x]; sh[x] = old + __popc(p); }
            old = __shfl_sync(p, old, leader); unsigned r = old + __popc(p & lt); __syncwarp();
            acc += r; x = (x * 5u + 1u + i) & 255u;
        }
    } else if (mode == 5) { // ballot ranking of 16 register-resident keys per lane, leader atomicAdd + shuffle (throughput form)
        unsigned lt = (1u << lane) - 1u;
        unsigned keys[16];
        for (int j = 0; j < 16; ++j) keys[j] = (x * (2 * j + 1) + j * 37u) & 255u;
        for (int i = 0; i < 16; ++i) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const unsigned d = keys[j];
                unsigned p = 0xffffffffu;
#pragma unroll
                for (int b = 0; b < 8; ++b) { unsigned bal = __ballot_sync(0xffffffffu, (d >> b) & 1u); p &= ((d >> b) & 1u) ? bal : ~bal; }
                const int leader = 31 - __clz(p);
                unsigned old = 0;
                if ((int)lane == leader) old = atomicAdd(&sh[d], __popc(p));
                old = __shfl_sync(0xffffffffu, old, leader);
                acc += old + __popc(p & lt);
            }
            for (int j = 0; j < 16; ++j) keys[j] = (keys[j] * 5u + 1u + i) & 255u;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[mode] = t1 - t0;
}
int main()
{
    unsigned* out; long long* cyc; cudaMalloc(&out, 4 * 1024 * 1024); cudaMallocManaged(&cyc, 64);
    for (int warps = 1; warps <= 32; warps *= (warps == 8 ? 4 : 8))
        for (int mode = 0; mode < 6; ++mode) {
            k<<<1, 32 * warps>>>(out, cyc, mode); cudaDeviceSynchronize();
            k<<<1, 32 * warps>>>(out, cyc, mode); cudaDeviceSynchronize();
            printf("warps=%d mode=%d cycles/op=%.1f\n", warps, mode, (double)cyc[mode] / 256.0);
        }
    return 0;
}
