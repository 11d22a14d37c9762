// Micro-benchmark behind DESIGN.md's "shrink reads more than it keeps": what DRAM fetches when a kernel gathers 16-byte
// pieces out of larger aligned chunks.  Every warp lane reads ONE 16-byte piece; `stride` bytes separate the pieces
// (16 = dense, 32 = every other piece, 64, 128), streaming loads (ld.global.cs, as the shrink kernel uses), a buffer
// of 2 GiB so nothing is served from L2.  Under `ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum`
// the rows give DRAM bytes per useful byte; stand-alone the kernel prints useful GB/s.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_granularity gather_granularity.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void gather(const uint4* __restrict__ src, long long pieces, int stride16, unsigned* sink) {
    unsigned acc = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pieces; i += (long long)gridDim.x * blockDim.x) {
        const uint4 v = __ldcs(src + i * stride16);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

int main() {
    const size_t bytes = 2ull << 30;
    uint4* buf;
    unsigned* sink;
    cudaMalloc(&buf, bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int stride = 16; stride <= 128; stride *= 2) {
        const long long pieces = (long long)(bytes / stride);
        gather<<<148 * 16, 256>>>(buf, pieces, stride / 16, sink);
        cudaEventRecord(a);
        gather<<<148 * 16, 256>>>(buf, pieces, stride / 16, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        printf("{\"piece_bytes\": 16, \"stride_bytes\": %d, \"useful_bytes\": %lld, \"ms\": %.3f, \"useful_gbs\": %.1f, \"span_gbs\": %.1f}\n", stride,
               pieces * 16, ms, pieces * 16 / ms / 1e6, (double)bytes / ms / 1e6);
    }
    return 0;
}
