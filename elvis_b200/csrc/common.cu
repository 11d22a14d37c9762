#include "common.cuh"

namespace elvis {
thread_local int g_last_cuda_error = 0;

int num_sms() {
    static std::atomic<int> cache[128];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 128) {
        const int v = cache[dev].load(std::memory_order_relaxed);
        if (v > 0) return v;
    }
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (dev >= 0 && dev < 128) cache[dev].store(n, std::memory_order_relaxed);
    return n;
}
}

extern "C" int elvis_abi_version(void) { return ELVIS_B200_ABI_VERSION; }

extern "C" int elvis_last_cuda_error(void) { return elvis::g_last_cuda_error; }

extern "C" const char* elvis_error_string(int code) {
    switch (code) {
        case ELVIS_OK: return "ok";
        case ELVIS_ERR_INVALID_ARG: return "invalid argument";
        case ELVIS_ERR_UNSUPPORTED: return "unsupported configuration";
        case ELVIS_ERR_CUDA: return "CUDA runtime error";
        case ELVIS_ERR_SHAPE: return "image dimensions must be divisible by block_size";
        default: return "unknown error";
    }
}
