// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bench tmem_bench.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define LD32(v, addr) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
    : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),"=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(addr) : "memory")
#define ST32(v, addr) asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
    :: "r"(addr), "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]),"r"(v[16]),"r"(v[17]),"r"(v[18]),"r"(v[19]),"r"(v[20]),"r"(v[21]),"r"(v[22]),"r"(v[23]),"r"(v[24]),"r"(v[25]),"r"(v[26]),"r"(v[27]),"r"(v[28]),"r"(v[29]),"r"(v[30]),"r"(v[31]) : "memory")

template <int MODE>   // 0 = ld, 1 = st
__global__ void bench(int iters, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 32 * ((warp >> 2) & 7);
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
    ST32(v, addr);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
            LD32(v, addr);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) acc ^= v[i];
        } else if (MODE == 2) {      // two loads in flight per wait
            uint32_t w[32];
            LD32(v, addr);
            LD32(w, addr + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) acc ^= v[i] ^ w[i];
        } else if (MODE == 3) {      // no wait between loads (wait every 8)
            LD32(v, addr);
            if ((it & 7) == 7) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
            v[5] ^= acc + it;
            ST32(v, addr);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    if (acc == 0x12345678u) sink[threadIdx.x] = acc + v[3];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long* d_cycles;
    uint32_t* sink;
    cudaMalloc(&d_cycles, 8);
    cudaMalloc(&sink, 4096);
    const int iters = 2000;
    for (int mode = 0; mode < 4; ++mode)
        for (int warps : {1, 4, 8, 16}) {
            long long c = 0;
            if (mode == 0) bench<0><<<148, warps * 32>>>(iters, d_cycles, sink);
            else if (mode == 1) bench<1><<<148, warps * 32>>>(iters, d_cycles, sink);
            else if (mode == 2) bench<2><<<148, warps * 32>>>(iters, d_cycles, sink);
            else bench<3><<<148, warps * 32>>>(iters, d_cycles, sink);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(&c, d_cycles, 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)iters * warps * 32 * 32 * 4 * (mode == 2 ? 2 : 1);
            printf("mode %d %s warps=%2d cycles=%lld  %.1f B/clk/SM  %.1f clk per x32 op per warp  (%s)\n", mode, mode == 1 ? "st" : "ld", warps, c, bytes / c,
                   (double)c / iters, cudaGetErrorString(e));
        }
    return 0;
}
