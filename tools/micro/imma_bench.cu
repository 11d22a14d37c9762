// Micro-benchmark: legacy-path integer tensor-core throughput on B200, mma.sync.m16n8k16 (u8 x u8 -> s32)
// and m16n8k32, as cycles per instruction per SM sub-partition for 1..8 warps per sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_bench imma_bench.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int K32, int CHAINS>
__global__ void bench(int iters, long long* cycles, int* sink) {
    uint32_t a0 = threadIdx.x * 2654435761u, a1 = a0 ^ 0x9e3779b9u, a2 = a0 + 77u, a3 = a1 + 99u, b0 = a0 * 31u + 7u, b1 = b0 ^ a1;
    int c[CHAINS][4];
#pragma unroll
    for (int j = 0; j < CHAINS; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[j][i] = j + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) {
            if (K32)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3]) : "r"(a0), "r"(a1), "r"(b0));
        }
    }
    const long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    long long* d_cycles;
    int* d_sink;
    cudaMalloc(&d_cycles, 148 * sizeof(long long));
    cudaMalloc(&d_sink, 148 * 1024 * sizeof(int));
    const int iters = 2000;
    for (int k32 = 0; k32 < 2; ++k32)
        for (int warps = 4; warps <= 32; warps *= 2) {
            if (k32) bench<1, 4><<<148, warps * 32>>>(iters, d_cycles, d_sink);
            else bench<0, 4><<<148, warps * 32>>>(iters, d_cycles, d_sink);
            long long h[148];
            cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
            cudaError_t e = cudaDeviceSynchronize();
            // per sub-partition: warps/4 warps x 4 chains x iters instructions
            const double per = (double)h[0] / ((double)iters * 4 * (warps / 4));
            printf("{\"instr\": \"IMMA m16n8k%d u8\", \"warps_per_sm\": %d, \"cycles_per_instr_per_subpartition\": %.2f, \"cuda\": \"%s\"}\n",
                   k32 ? 32 : 16, warps, per, cudaGetErrorString(e));
        }
    return 0;
}
