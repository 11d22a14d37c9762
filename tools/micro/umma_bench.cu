// Micro-benchmark: tcgen05.mma issue/execute rate for small N, A operand in tensor memory (TS) or
// shared memory (SS).  One CTA per SM, one issuing thread; cycles per MMA from clock64 around a
// batch that ends with tcgen05.commit + mbarrier wait.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t desc128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int N, bool TS, bool CHAIN>
__global__ void bench(int batches, int per_batch, long long* cycles) {
    extern __shared__ uint8_t raw[];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar;
    const uint32_t base = ((uint32_t)__cvta_generic_to_shared(raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_addr), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw + (base - (uint32_t)__cvta_generic_to_shared(raw)))[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    if (threadIdx.x == 0) {
        const uint64_t a_desc = desc128(base), b_desc = desc128(base + 32768);
        uint32_t parity = 0;
        long long total = 0;
        for (int bt = 0; bt < batches; ++bt) {
            const long long t0 = clock64();
            for (int i = 0; i < per_batch; ++i) {
                const uint32_t d = tmem + 256 + (CHAIN ? 0 : (i & 1) * N);   // D at columns 256.., A (TS) at columns 0..
                const uint32_t k = i & 3;
                if (TS) mma_ts(d, tmem + 8 * k, b_desc + 2 * k, idesc, CHAIN ? (i > 0) : 0);
                else mma_ss(d, a_desc + 2 * k, b_desc + 2 * k, idesc, CHAIN ? (i > 0) : 0);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar_addr), "r"(parity) : "memory");
            parity ^= 1;
            total += clock64() - t0;
        }
        if (blockIdx.x == 0) *cycles = total;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, bool TS, bool CHAIN>
void run(long long* d_cycles, int per_batch) {
    const int batches = 200;
    cudaFuncSetAttribute(bench<N, TS, CHAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    bench<N, TS, CHAIN><<<148, 128, 70000>>>(batches, per_batch, d_cycles);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_cycles, 8, cudaMemcpyDeviceToHost);
    printf("N=%3d %s %s per_batch=%2d: %.1f clk per batch, %.1f clk per MMA (%s)\n", N, TS ? "TS" : "SS", CHAIN ? "chain" : "indep", per_batch,
           (double)c / batches, (double)c / batches / per_batch, cudaGetErrorString(e));
}

int main() {
    long long* d_cycles;
    cudaMalloc(&d_cycles, 8);
    for (int pb : {8, 64}) {
        run<64, true, true>(d_cycles, pb);
        run<64, true, false>(d_cycles, pb);
        run<64, false, true>(d_cycles, pb);
        run<128, true, true>(d_cycles, pb);
        run<128, false, true>(d_cycles, pb);
        run<256, true, true>(d_cycles, pb);
        run<32, true, true>(d_cycles, pb);
    }
    return 0;
}
