// Micro-benchmark behind DESIGN.md's "shrink reads more than it keeps": what DRAM fetches when a kernel gathers 16-byte
// pieces out of larger aligned chunks.  Every lane reads ONE 16-byte piece; `stride` bytes separate the pieces
// (16 = dense, 32 = every other piece, 64, 128); a 2 GiB buffer, so nothing is served from L2.  Load forms:
//   cs      ld.global.cs (streaming, what the shrink kernel uses)        plain   ld.global
//   l2_64   ld.global.L2::64B (smallest L2 prefetch-size hint PTX has)   cs_64   ld.global.cs.L2::64B
//   nc      ld.global.nc.L1::no_allocate                                 lim32   ld.global.cs after
//                                                                                cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32)
// Under `ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,...` the rows give DRAM bytes per
// useful byte (launch order = print order); stand-alone the program prints useful GB/s.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_granularity gather_granularity.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE> __device__ __forceinline__ uint4 load16(const uint4* p) {
    uint4 v;
    if (MODE == 0) asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 1) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 2) asm volatile("ld.global.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 3) asm volatile("ld.global.cs.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 4) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int MODE>
__global__ void gather(const uint4* __restrict__ src, long long pieces, int stride16, unsigned* sink) {
    unsigned acc = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pieces; i += (long long)gridDim.x * blockDim.x) {
        const uint4 v = load16<MODE>(src + i * stride16);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int MODE> void run(const char* name, const uint4* buf, size_t bytes, unsigned* sink) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int stride = 16; stride <= 128; stride *= 2) {
        const long long pieces = (long long)(bytes / stride);
        cudaEventRecord(a);
        gather<MODE><<<148 * 16, 256>>>(buf, pieces, stride / 16, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        printf("{\"load\": \"%s\", \"piece_bytes\": 16, \"stride_bytes\": %d, \"useful_bytes\": %lld, \"ms\": %.3f, \"useful_gbs\": %.1f, \"span_gbs\": %.1f}\n",
               name, stride, pieces * 16, ms, pieces * 16 / ms / 1e6, (double)bytes / ms / 1e6);
    }
}

int main() {
    const size_t bytes = 2ull << 30;
    uint4* buf;
    unsigned* sink;
    cudaMalloc(&buf, bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bytes);
    gather<0><<<148 * 16, 256>>>(buf, (long long)(bytes / 16), 1, sink);      // warm-up
    cudaDeviceSynchronize();
    run<0>("cs", buf, bytes, sink);
    run<1>("plain", buf, bytes, sink);
    run<2>("l2_64", buf, bytes, sink);
    run<3>("cs_64", buf, bytes, sink);
    run<4>("nc", buf, bytes, sink);
    size_t g = 0;
    cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    const cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    size_t g2 = 0;
    cudaDeviceGetLimit(&g2, cudaLimitMaxL2FetchGranularity);
    printf("{\"cudaLimitMaxL2FetchGranularity\": {\"before\": %zu, \"set_32\": \"%s\", \"after\": %zu}}\n", g, cudaGetErrorString(e), g2);
    run<0>("lim32_cs", buf, bytes, sink);
    run<1>("lim32_plain", buf, bytes, sink);
    run<3>("lim32_cs_64", buf, bytes, sink);
    return 0;
}
