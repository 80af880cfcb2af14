// Micro-benchmark: how fast can SMs read PINNED HOST memory directly (zero-copy) with 16-byte loads?
//   (a) one contiguous 32 MiB block, (b) row segments of `seg` bytes at a pitch of 1002 bytes (a 334-pixel BGR row),
// against cudaMemcpyAsync of the same bytes.  Context for fmhr_ham_host_u8_submit_boxes (profiles/README.md).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void pull(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v;
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + i));
        dst[i] = v;
    }
}
__global__ void pull_cv(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v;
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + i) : "memory");
        dst[i] = v;
    }
}
// one warp per row segment: lane l reads 16-byte chunk l, l + 32, ...
__global__ void pull_rows(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, int rows, int pitch, int seg) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < rows; r += nw) {
        const size_t base = ((size_t)r * pitch) & ~(size_t)15;
        for (int c = lane * 16; c < seg + 16; c += 512) {
            uint4 v;
            asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + base + c));
            *(uint4*)(dst + base + c) = v;
        }
    }
}
int main() {
    const size_t bytes = 32u << 20;
    unsigned char *h, *d;
    cudaHostAlloc(&h, bytes, cudaHostAllocMapped);
    cudaMalloc(&d, bytes);
    for (size_t i = 0; i < bytes; i++) h[i] = (unsigned char)i;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        for (int k = 0; k < 10; k++) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("cudaMemcpyAsync 32 MiB contiguous      : %6.1f GB/s\n", 10 * bytes / (ms * 1e-3) / 1e9);
    for (int blocks : {148, 592, 2368}) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            for (int k = 0; k < 10; k++) pull<<<blocks, 256>>>((const uint4*)h, (uint4*)d, bytes / 16);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("kernel pull, contiguous, %4d blocks    : %6.1f GB/s\n", blocks, 10 * bytes / (ms * 1e-3) / 1e9);
    }
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        for (int k = 0; k < 10; k++) pull_cv<<<592, 256>>>((const uint4*)h, (uint4*)d, bytes / 16);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("kernel pull, ld.relaxed.sys, 592 blocks : %6.1f GB/s\n", 10 * bytes / (ms * 1e-3) / 1e9);
    {   // does a second launch see bytes the host changed in place?  (.nc and .relaxed.sys)
        unsigned char* back = (unsigned char*)malloc(1 << 20);
        for (int mode = 0; mode < 2; mode++) {
            int stale = 0;
            for (int round = 0; round < 4; round++) {
                for (size_t i = 0; i < (1 << 20); i++) h[i] = (unsigned char)(i * 7 + round * 13 + mode);
                if (mode == 0) pull<<<592, 256>>>((const uint4*)h, (uint4*)d, (1 << 20) / 16);
                else pull_cv<<<592, 256>>>((const uint4*)h, (uint4*)d, (1 << 20) / 16);
                cudaMemcpy(back, d, 1 << 20, cudaMemcpyDeviceToHost);
                for (size_t i = 0; i < (1 << 20); i++) stale += back[i] != (unsigned char)(i * 7 + round * 13 + mode);
            }
            printf("in-place host updates seen by %s loads: %s (%d stale bytes)\n", mode ? "ld.relaxed.sys" : "ld.global.nc",
                   stale ? "NO" : "yes", stale);
        }
        free(back);
    }
    const int pitch = 1002, rows = (int)(bytes / pitch) - 1;
    for (int seg : {150, 450, 1002}) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            for (int k = 0; k < 10; k++) pull_rows<<<592, 256>>>(h, d, rows, pitch, seg);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("kernel pull, %5d rows of %4d B / 1002  : %6.1f GB/s useful, %.3f ms per 22k rows\n", rows, seg,
               10.0 * rows * seg / (ms * 1e-3) / 1e9, ms / 10 * 22000.0 / rows);
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            for (int k = 0; k < 10; k++) cudaMemcpy2DAsync(d, pitch, h, pitch, seg, rows, cudaMemcpyHostToDevice);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("cudaMemcpy2DAsync, same rows            : %6.1f GB/s useful, %.3f ms per 22k rows\n",
               10.0 * rows * seg / (ms * 1e-3) / 1e9, ms / 10 * 22000.0 / rows);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
