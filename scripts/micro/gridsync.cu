// Development probe: cost of a cooperative grid-wide barrier on B200 for the grid shapes the wide engine could use.
#include <cstdio>
#include <cooperative_groups.h>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void k(int n, long long* out) {
    cg::grid_group g = cg::this_grid();
    g.sync();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) g.sync();
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (t1 - t0) / n;
}
__global__ void empty() {}

int main() {
    long long* out; cudaMalloc(&out, 8);
    int n = 200;
    int shapes[][2] = {{148, 1024}, {148, 256}, {296, 256}, {592, 256}, {888, 256}, {148 * 8, 128}};
    for (auto& s : shapes) {
        void* args[] = {&n, &out};
        cudaError_t e = cudaLaunchCooperativeKernel((void*)k, dim3(s[0]), dim3(s[1]), args, 0, 0);
        cudaDeviceSynchronize();
        long long r = 0; cudaMemcpy(&r, out, 8, cudaMemcpyDeviceToHost);
        printf("grid %4d x %4d: %lld cycles per grid.sync (%s)\n", s[0], s[1], r, cudaGetErrorString(e));
    }
    // launch-to-launch gap of empty kernels inside a graph
    cudaStream_t st; cudaStreamCreate(&st);
    cudaGraph_t gr; cudaGraphExec_t ex;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < 100; ++i) empty<<<592, 256, 0, st>>>();
    cudaStreamEndCapture(st, &gr);
    cudaGraphInstantiate(&ex, gr, 0);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a, st); cudaGraphLaunch(ex, st); cudaEventRecord(b, st); cudaStreamSynchronize(st);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("graph of 100 empty 592x256 kernels: %.2f us per kernel\n", ms * 10.f);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
