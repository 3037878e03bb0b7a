// Calibration for the detector's leader warp: what does a lone warp pay per instruction on this part?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench_single_warp microbench_single_warp.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_dep_alu(unsigned *out, int iters, long long *cyc) {
    unsigned a = threadIdx.x, b = 0x9e3779b9u;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 32; k++) a = (a & b) | (a >> 1);        // 32 dependent LOP3/SHF pairs
    }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_indep_alu(unsigned *out, int iters, long long *cyc) {
    unsigned a[8];
    for (int k = 0; k < 8; k++) a[k] = threadIdx.x + k;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) a[k] = (a[k] * 3u) ^ 0x5bd1e995u;   // 8 independent chains
    }
    long long t1 = clock64();
    unsigned s = 0;
    for (int k = 0; k < 8; k++) s ^= a[k];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_lds_vote(unsigned *out, int iters, long long *cyc) {
    __shared__ unsigned sm[32 * 256];
    for (int i = threadIdx.x; i < 32 * 256; i += 32) sm[i] = 0;
    __syncwarp();
    unsigned fv0 = 1, fv1 = 2, fv2 = 4, fv3 = 8, acc_all = 0;
    int row = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        const uint4 x = *reinterpret_cast<const uint4 *>(&sm[row * 256 + threadIdx.x * 8]);
        const uint4 y = *reinterpret_cast<const uint4 *>(&sm[row * 256 + threadIdx.x * 8 + 4]);
        unsigned acc = (x.x & fv0) | (x.y & fv1) | (x.z & fv2) | (x.w & fv3) | (y.x & fv0) | (y.y & fv1) | (y.z & fv2) | (y.w & fv3);
        if (__any_sync(0xffffffffu, acc != 0)) acc_all += 1;
        row = (row + 1) & 31;
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc_all;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_sysmem_store(unsigned long long *host_mapped, unsigned long long *dev, int iters, long long *cyc) {
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (threadIdx.x == (i & 31)) {
            unsigned long long *g = host_mapped + (size_t)(i & 1023) * 7;
#pragma unroll
            for (int k = 0; k < 7; k++) g[k] = (unsigned long long)i + k;
        }
        __syncwarp();
    }
    long long t1 = clock64();
    for (int i = 0; i < iters; i++) {
        if (threadIdx.x == (i & 31)) {
            unsigned long long *g = dev + (size_t)(i & 1023) * 7;
#pragma unroll
            for (int k = 0; k < 7; k++) g[k] = (unsigned long long)i + k;
        }
        __syncwarp();
    }
    long long t2 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
}
int main() {
    unsigned *out; long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    const int iters = 100000;
    for (int rep = 0; rep < 2; rep++) {
        k_dep_alu<<<1, 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("dependent ALU chain: %.2f cycles per instruction (64 per iteration)\n", (double)h / iters / 64);
        k_indep_alu<<<1, 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("8 independent chains: %.2f cycles per instruction (64 per iteration)\n", (double)h / iters / 64);
        k_lds_vote<<<1, 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("2x LDS.128 + 8 LOP3 + vote + branch loop: %.1f cycles per iteration\n", (double)h / iters);
    }
    {
        unsigned long long *hm, *dm, *dev; long long *c2, hc[2];
        cudaHostAlloc((void **)&hm, 1024 * 7 * 8, cudaHostAllocMapped);
        cudaHostGetDevicePointer((void **)&dm, hm, 0);
        cudaMalloc(&dev, 1024 * 7 * 8); cudaMalloc(&c2, 16);
        for (int rep = 0; rep < 2; rep++) { k_sysmem_store<<<1, 32>>>(dm, dev, 20000, c2); cudaMemcpy(hc, c2, 16, cudaMemcpyDeviceToHost); }
        printf("one lane, 7 x 8-byte stores per iteration: mapped host memory %.1f cycles, device memory %.1f cycles\n",
               (double)hc[0] / 20000, (double)hc[1] / 20000);
    }
    return 0;
}
