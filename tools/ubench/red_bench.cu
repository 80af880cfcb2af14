// Micro-benchmark: throughput of red.global.add.v4.f32 scatter patterns on B200 (design input for the HAM pixel backward).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o red_bench red_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// A: 48-byte records, 3 x v4 per ref, one ref per lane
__global__ void k_a(const int* __restrict__ idx, int n, float4* G) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int i = idx[e];
        const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
        atomicAdd(G + 3 * (size_t)i, v); atomicAdd(G + 3 * (size_t)i + 1, v); atomicAdd(G + 3 * (size_t)i + 2, v);
    }
}
// B: 32-byte records, 2 x v4 per ref, one ref per lane
__global__ void k_b(const int* __restrict__ idx, int n, float4* G) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int i = idx[e];
        const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
        atomicAdd(G + 2 * (size_t)i, v); atomicAdd(G + 2 * (size_t)i + 1, v);
    }
}
// C: 32-byte records, lane pairs complete a sector per instruction (same number of instructions as B)
__global__ void k_c(const int* __restrict__ idx, int n, float4* G) {
    const int lane = threadIdx.x & 31;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int i = idx[e];
        const int ie = __shfl_sync(0xffffffffu, i, lane & ~1), io = __shfl_sync(0xffffffffu, i, lane | 1);
        const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
        atomicAdd(G + 2 * (size_t)ie + (lane & 1), v);
        atomicAdd(G + 2 * (size_t)io + (lane & 1), v);
    }
}
// D: 32-byte records, 1 x v4 per ref
__global__ void k_d(const int* __restrict__ idx, int n, float4* G) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int i = idx[e];
        atomicAdd(G + 2 * (size_t)i, make_float4(1.f, 2.f, 3.f, 4.f));
    }
}
// E: scalar reds, 8 per ref
__global__ void k_e(const int* __restrict__ idx, int n, float* G) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int i = idx[e];
#pragma unroll
        for (int k = 0; k < 8; k++) atomicAdd(G + 8 * (size_t)i + k, 1.0f);
    }
}
// F: v2 reds, 4 per ref
__global__ void k_f(const int* __restrict__ idx, int n, float2* G) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int i = idx[e];
#pragma unroll
        for (int k = 0; k < 4; k++) atomicAdd(G + 4 * (size_t)i + k, make_float2(1.f, 2.f));
    }
}
// G: shared-memory pre-aggregation is not modelled; H: plain stores as the no-atomic floor (2 x v4)
__global__ void k_h(const int* __restrict__ idx, int n, float4* G) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int i = idx[e];
        const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
        G[2 * (size_t)i] = v; G[2 * (size_t)i + 1] = v;
    }
}

int main(int argc, char** argv) {
    const int V = 49281, n = 660000 * 3;
    const int mode = argc > 1 ? atoi(argv[1]) : 0;  // 0 random, 1 locally coherent (window of 512 vertices)
    std::vector<int> h(n);
    uint32_t s = 12345u;
    for (int e = 0; e < n; e++) {
        s = s * 1664525u + 1013904223u;
        if (mode == 0) h[e] = (int)((s >> 8) % V);
        else h[e] = (int)((((uint64_t)e * V) / n + ((s >> 8) % 512)) % V);
    }
    int* idx; float4* G;
    CK(cudaMalloc(&idx, n * sizeof(int)));
    CK(cudaMalloc(&G, (size_t)V * 64));
    CK(cudaMemcpy(idx, h.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemset(G, 0, (size_t)V * 64));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grid = 148 * 4, reps = 50;
    const char* names[] = {"A 48B rec, 3 x v4/ref", "B 32B rec, 2 x v4/ref", "C 32B rec, lane-pair sectors, 2 x v4/ref",
                           "D 32B rec, 1 x v4/ref", "E 8 scalar red/ref", "F 4 x v2/ref", "H 2 x v4 plain stores/ref"};
    for (int k = 0; k < 7; k++) {
        for (int r = -3; r < reps; r++) {
            if (r == 0) CK(cudaEventRecord(e0));
            switch (k) {
                case 0: k_a<<<grid, 256>>>(idx, n, G); break;
                case 1: k_b<<<grid, 256>>>(idx, n, G); break;
                case 2: k_c<<<grid, 256>>>(idx, n, G); break;
                case 3: k_d<<<grid, 256>>>(idx, n, G); break;
                case 4: k_e<<<grid, 256>>>(idx, n, (float*)G); break;
                case 5: k_f<<<grid, 256>>>(idx, n, (float2*)G); break;
                case 6: k_h<<<grid, 256>>>(idx, n, G); break;
            }
        }
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("mode %d  %-44s %8.2f us / launch  (%d refs)\n", mode, names[k], ms * 1000.f / reps, n);
    }
    return 0;
}
