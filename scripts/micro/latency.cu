// Development probe: dependent-chain latencies (cycles) of the primitives the single-CTA tail engine is built from.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/latency scripts/micro/latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N_IT 64

__device__ __forceinline__ uint32_t ld_ca_u32(const uint32_t* p) {
    uint32_t r;
    asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {
    uint32_t r;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}

__global__ void __launch_bounds__(1024, 1) probe(uint32_t* chase, uint32_t n_chase, long long* out, uint32_t* sink, double* dsink) {
    __shared__ uint32_t s_chase[1024];
    __shared__ uint32_t s_atom;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 1024; i += blockDim.x) s_chase[i] = (i * 37 + 11) & 1023;
    if (tid == 0) s_atom = 0;
    __syncthreads();
    long long t0, t1;
    uint32_t x = lane + 1;
    int slot = 0;
#define REPORT(v) do { if (tid == 0) out[slot] = (v); slot++; } while (0)

    // 0: clock overhead
    t0 = clock64(); t1 = clock64(); REPORT(t1 - t0);

    // 1: REDUX max chain
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_IT; ++i) x = __reduce_max_sync(0xffffffffu, x + lane) ^ (uint32_t)i;
    t1 = clock64(); REPORT((t1 - t0) / N_IT);

    // 2: match.any unique values
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_IT; ++i) x = __match_any_sync(0xffffffffu, (x & 0xffff0000u) + lane) + i;
    t1 = clock64(); REPORT((t1 - t0) / N_IT);

    // 3: match.any all same
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_IT; ++i) x = __match_any_sync(0xffffffffu, x | 0xffffffffu) + i;
    t1 = clock64(); REPORT((t1 - t0) / N_IT);

    // 4: ld.ca pointer chase (L1 hits after first pass? footprint n_chase*4)
    __syncthreads();
    uint32_t pidx = (uint32_t)tid % n_chase;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_IT; ++i) pidx = ld_ca_u32(chase + pidx);
    t1 = clock64(); REPORT((t1 - t0) / N_IT);
    x += pidx;

    // 5: ld.cg pointer chase (L2)
    __syncthreads();
    pidx = (uint32_t)(tid * 977) % n_chase;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_IT; ++i) pidx = ld_cg_u32(chase + pidx);
    t1 = clock64(); REPORT((t1 - t0) / N_IT);
    x += pidx;

    // 6: LDS chase
    __syncthreads();
    pidx = tid & 1023;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_IT; ++i) pidx = s_chase[pidx];
    t1 = clock64(); REPORT((t1 - t0) / N_IT);
    x += pidx;

    // 7: __syncthreads chain, all 1024 threads
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_IT; ++i) __syncthreads();
    t1 = clock64(); REPORT((t1 - t0) / N_IT);

    // 8: DADD chain
    double d = (double)x;
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_IT; ++i) d = d + 1.5;
    t1 = clock64(); REPORT((t1 - t0) / N_IT);

    // 9: shfl chain
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_IT; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31) + 1;
    t1 = clock64(); REPORT((t1 - t0) / N_IT);

    // 10: ballot chain
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_IT; ++i) x = __ballot_sync(0xffffffffu, (x >> lane) & 1) + i;
    t1 = clock64(); REPORT((t1 - t0) / N_IT);

    // 11: smem atomicOr by lane 0 of each warp then barrier
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_IT; ++i) { if (lane == 0) atomicOr(&s_atom, 1u << warp); __syncthreads(); x += s_atom; __syncthreads(); }
    t1 = clock64(); REPORT((t1 - t0) / N_IT);

    // 12: only warp 0 active doing the REDUX chain while 31 warps wait at a barrier (does the waiting crowd slow it down?)
    __syncthreads();
    t0 = clock64();
    if (warp == 0) {
#pragma unroll 1
        for (int i = 0; i < N_IT; ++i) x = __reduce_max_sync(0xffffffffu, x + lane) ^ (uint32_t)i;
    }
    t1 = clock64(); REPORT((t1 - t0) / N_IT);
    __syncthreads();

    // 13: ld.ca chase by warp 0 only, others at barrier
    pidx = (uint32_t)(tid * 131) % n_chase;
    t0 = clock64();
    if (warp == 0) {
#pragma unroll 1
        for (int i = 0; i < N_IT; ++i) pidx = ld_ca_u32(chase + pidx);
    }
    t1 = clock64(); REPORT((t1 - t0) / N_IT);
    x += pidx;
    __syncthreads();

    // 14: global store then barrier chain (does BAR wait for STG acks?)
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_IT; ++i) { if (lane == 0) sink[1024 + warp * 32 + (i & 31)] = x; __syncthreads(); }
    t1 = clock64(); REPORT((t1 - t0) / N_IT);

    // 15: store to global then ld.ca of the same address by same thread (store->load forwarding latency)
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_IT; ++i) { sink[2048 + tid] = x; x += ld_ca_u32(sink + 2048 + tid); }
    t1 = clock64(); REPORT((t1 - t0) / N_IT);

    // 16: 64-bit REDUX emulation (two dependent 32-bit)
    unsigned long long k = ((unsigned long long)x << 20) + lane;
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_IT; ++i) {
        const uint32_t hi = __reduce_max_sync(0xffffffffu, (uint32_t)(k >> 32));
        const uint32_t lo = __reduce_max_sync(0xffffffffu, ((uint32_t)(k >> 32) == hi) ? (uint32_t)k : 0u);
        k = (((unsigned long long)hi << 32) | lo) + lane + i;
    }
    t1 = clock64(); REPORT((t1 - t0) / N_IT);
    x += (uint32_t)k;

    sink[tid] = x;
    dsink[tid] = d;
}

int main() {
    const uint32_t n = 1u << 22;   // 16 MB chase array: L2-resident, far beyond L1
    uint32_t* h = (uint32_t*)malloc(n * 4);
    for (uint32_t i = 0; i < n; ++i) h[i] = (uint32_t)(((unsigned long long)i * 2654435761ull + 12345ull) % n);
    uint32_t *d, *sink; double* dsink; long long* out;
    cudaMalloc(&d, n * 4); cudaMalloc(&sink, 1 << 20); cudaMalloc(&dsink, 1 << 20); cudaMalloc(&out, 64 * 8);
    cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
    cudaMemset(out, 0, 64 * 8);
    for (int rep = 0; rep < 2; ++rep) probe<<<1, 1024>>>(d, n, out, sink, dsink);
    cudaDeviceSynchronize();
    long long r[64];
    cudaMemcpy(r, out, 64 * 8, cudaMemcpyDeviceToHost);
    const char* names[] = {"clock64 overhead", "REDUX.max chain", "match.any unique", "match.any same", "ld.ca chase 16MB",
                           "ld.cg chase 16MB", "LDS chase", "__syncthreads x1024thr", "DADD chain", "shfl chain", "ballot chain",
                           "atomicOr+2 barriers+LDS", "REDUX warp0 only (others wait)", "ld.ca chase warp0 only", "STG + barrier",
                           "STG then ld.ca same addr", "64-bit max via 2 REDUX"};
    for (int i = 0; i < 17; ++i) printf("%-36s %lld\n", names[i], r[i]);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
