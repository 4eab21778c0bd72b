// Development probe: dependent-load latency (pointer chase) for a few footprints / load flavours on one warp.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

template <int MODE>
__global__ void chase(const uint32_t* __restrict__ next, uint32_t start, int iters, long long* out, uint32_t* sink) {
    uint32_t i = start;
    long long t0 = clock64();
    for (int k = 0; k < iters; ++k) {
        uint32_t v;
        if (MODE == 0) v = __ldg(next + i);
        else if (MODE == 1) asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(next + i) : "memory");
        else if (MODE == 2) v = __ldcg(next + i);
        else asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(next + i));
        i = v;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; sink[0] = i; }
}

int main() {
    const size_t sizes_mb[] = {1, 16, 64, 256, 1024};
    for (size_t mb : sizes_mb) {
        size_t n = mb * 1024 * 1024 / 4;
        // random cycle with stride >= 32 words (one sector per hop)
        size_t hops = n / 32;
        std::vector<uint32_t> perm(hops);
        for (size_t i = 0; i < hops; ++i) perm[i] = (uint32_t)i;
        std::mt19937 rng(1);
        std::shuffle(perm.begin(), perm.end(), rng);
        std::vector<uint32_t> h(n, 0);
        for (size_t i = 0; i < hops; ++i) h[(size_t)perm[i] * 32] = perm[(i + 1) % hops] * 32;
        uint32_t* d; long long* dout; uint32_t* sink;
        cudaMalloc(&d, n * 4); cudaMalloc(&dout, 8); cudaMalloc(&sink, 4);
        cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
        int iters = 20000;
        for (int mode = 0; mode < 4; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) chase<0><<<1, 32>>>(d, perm[0] * 32, iters, dout, sink);
                if (mode == 1) chase<1><<<1, 32>>>(d, perm[0] * 32, iters, dout, sink);
                if (mode == 2) chase<2><<<1, 32>>>(d, perm[0] * 32, iters, dout, sink);
                if (mode == 3) chase<3><<<1, 32>>>(d, perm[0] * 32, iters, dout, sink);
                cudaDeviceSynchronize();
            }
            long long c; cudaMemcpy(&c, dout, 8, cudaMemcpyDeviceToHost);
            printf("footprint %4zu MB mode %d (0 ldg,1 ca,2 cg,3 nc.noalloc): %.1f cycles/hop\n", mb, mode, (double)c / iters);
        }
        cudaFree(d); cudaFree(dout); cudaFree(sink);
    }
    return 0;
}
