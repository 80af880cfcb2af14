// Micro-benchmark: cost of a chain of dependent kernels inside a CUDA graph on B200 (how much of the HAM iteration's
// ~13-kernel critical path is launch / drain overhead), and of a software grid barrier inside one persistent kernel.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_touch(float* p, int n, int iters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float v = 0.f;
    for (int k = 0; k < iters; k++) v += p[(i + k * 4099) % n];   // dependent-ish gathers
    if (v == 123.456f) p[i % n] = v;
}
__global__ void k_coop(float* p, int n, int iters, int phases) {
    cg::grid_group g = cg::this_grid();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (int ph = 0; ph < phases; ph++) {
        float v = 0.f;
        for (int k = 0; k < iters; k++) v += p[(i + k * 4099 + ph) % n];
        if (v == 123.456f) p[i % n] = v;
        g.sync();
    }
}

int main() {
    const int n = 1 << 22;
    float* p;
    CK(cudaMalloc(&p, n * sizeof(float)));
    CK(cudaMemset(p, 0, n * sizeof(float)));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int iters : {0, 4, 16}) {
        for (int grid : {148, 592}) {
            for (int K : {1, 8, 16}) {
                cudaGraph_t gr; cudaGraphExec_t ge;
                CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
                for (int k = 0; k < K; k++) k_touch<<<grid, 256, 0, st>>>(p, n, iters);
                CK(cudaStreamEndCapture(st, &gr));
                CK(cudaGraphInstantiate(&ge, gr, 0));
                for (int r = 0; r < 20; r++) CK(cudaGraphLaunch(ge, st));
                CK(cudaEventRecord(e0, st));
                const int reps = 200;
                for (int r = 0; r < reps; r++) CK(cudaGraphLaunch(ge, st));
                CK(cudaEventRecord(e1, st));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                printf("graph: iters %2d grid %3d  K %2d kernels: %7.2f us / graph  (%.2f us / kernel)\n", iters, grid, K,
                       ms * 1000.f / reps, ms * 1000.f / reps / K);
                CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(gr));
            }
        }
    }
    // cooperative persistent kernel with grid barriers
    for (int iters : {0, 4, 16}) {
        for (int phases : {1, 8, 16}) {
            int grid = 592, nn = n;
            void* args[] = {&p, &nn, &iters, &phases};
            for (int r = 0; r < 5; r++) CK(cudaLaunchCooperativeKernel((void*)k_coop, dim3(grid), dim3(256), args, 0, st));
            CK(cudaEventRecord(e0, st));
            const int reps = 100;
            for (int r = 0; r < reps; r++) CK(cudaLaunchCooperativeKernel((void*)k_coop, dim3(grid), dim3(256), args, 0, st));
            CK(cudaEventRecord(e1, st));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("coop:  iters %2d grid 592  %2d phases: %7.2f us / launch (%.2f us / phase)\n", iters, phases,
                   ms * 1000.f / reps, ms * 1000.f / reps / phases);
        }
    }
    return 0;
}
