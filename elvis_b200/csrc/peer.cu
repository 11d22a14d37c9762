// Multi-GPU plumbing of the frame-sharded path over peer memory (NVLink / NVSwitch), SURVEY.md 8e:
//   * the one-frame luma halo goes to the neighbour with a COPY-ENGINE peer copy straight into its halo
//     slot (no SM-resident communication kernel competes with the three compute kernels of the stream
//     pipeline), followed by a release store of a sequence number into the neighbour's flag;
//   * the global min / max normalisations (elvis.py:1175-1178, 1216) are one-shot all-reduces through
//     per-rank mailboxes: every rank stores its few values into its slot of every peer's mailbox, raises
//     its flag there, waits for the N flags of its own mailbox and reduces locally -- one tiny kernel,
//     about one NVLink store latency, instead of an NCCL launch per reduction.
// Memory that peers touch is allocated here with cudaMalloc and shared through CUDA IPC handles (the
// handles travel over torch.distributed, elvis_b200/peer.py).  Every wait has a time limit: a peer that
// never arrives sets an error word instead of hanging the GPU.  NCCL send/recv + all_reduce
// (elvis_b200/sharding.py) stays the default transport and the fallback.
#include "common.cuh"
#include <cstring>

namespace elvis {
namespace {

constexpr int kMaxRanks = 16;
constexpr int kMaxValues = 8;
constexpr unsigned long long kWaitLimitNs = 4000000000ull;   // 4 s

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// sequence numbers wrap: "a has reached b" in modular arithmetic
__device__ __forceinline__ bool reached(uint32_t a, uint32_t b) { return (int32_t)(a - b) >= 0; }

__device__ bool wait_flag(const uint32_t* flag, uint32_t seq) {
    const unsigned long long t0 = global_ns();
    while (!reached(ld_acquire_sys(flag), seq)) {
        if (global_ns() - t0 > kWaitLimitNs) return false;
        __nanosleep(100);
    }
    return true;
}

__global__ void flag_write_kernel(uint32_t* flag, uint32_t seq) {
    __threadfence_system();
    st_release_sys(flag, seq);
}

__global__ void flag_wait_kernel(const uint32_t* flag, uint32_t seq, int32_t* error) {
    if (!wait_flag(flag, seq) && error) atomicExch(error, 1);
}

// Mailbox of one rank: values[slot][rank][kMaxValues] doubles, then flags[slot][rank] words.
struct Mailboxes {
    double* values[kMaxRanks];     // peers' value areas (own included), as mapped in THIS process
    uint32_t* flags[kMaxRanks];
};

template <typename F>
__global__ void __launch_bounds__(32) allreduce_minmax_kernel(F* inout, int n, int rank, int world, Mailboxes mb, int slot,
                                                              uint32_t seq, int32_t* error) {
    const int lane = threadIdx.x;
    // publish: lane r stores my values into my slot of rank r's mailbox, then raises my flag there
    if (lane < world) {
        double* dst = mb.values[lane] + ((int64_t)slot * kMaxRanks + rank) * kMaxValues;
        for (int j = 0; j < n; ++j) dst[j] = (double)inout[j];
        __threadfence_system();
        st_release_sys(mb.flags[lane] + slot * kMaxRanks + rank, seq);
    }
    // gather + reduce: lane j owns value j (even index = a minimum, odd = a maximum).  It waits for every rank's
    // flag in MY mailbox itself (acquire), so that the values it then reads are the ones released with the flag.
    bool ok = true;
    double acc = 0.0;
    if (lane < n) {
        const double* src = mb.values[rank] + (int64_t)slot * kMaxRanks * kMaxValues;
        const uint32_t* fl = mb.flags[rank] + slot * kMaxRanks;
        for (int r = 0; r < world && ok; ++r) {
            ok = wait_flag(fl + r, seq);
            if (!ok) break;
            const double v = src[r * kMaxValues + lane];
            acc = r == 0 ? v : ((lane & 1) ? fmax(acc, v) : fmin(acc, v));
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    if (!ok) {
        if (lane == 0 && error) atomicExch(error, 1);
        return;
    }
    if (lane < n) inout[lane] = (F)acc;
}

}  // namespace
}  // namespace elvis

using namespace elvis;

extern "C" int64_t elvis_peer_mailbox_bytes(void) {
    return (int64_t)4 * kMaxRanks * kMaxValues * sizeof(double) + (int64_t)4 * kMaxRanks * sizeof(uint32_t);   // 4 slots
}

extern "C" int elvis_peer_alloc(int64_t bytes, void** device_ptr, void* host_handle) {
    if (bytes <= 0 || !device_ptr || !host_handle) return ELVIS_ERR_INVALID_ARG;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e != cudaSuccess) return cuda_fail(e);
    e = cudaMemset(p, 0, (size_t)bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return cuda_fail(e);
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(host_handle, &h, sizeof(h));
    *device_ptr = p;
    return ELVIS_OK;
}

extern "C" int elvis_peer_open(const void* host_handle, void** device_ptr) {
    if (!host_handle || !device_ptr) return ELVIS_ERR_INVALID_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, host_handle, sizeof(h));
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return cuda_fail(e);
    *device_ptr = p;
    return ELVIS_OK;
}

extern "C" int elvis_peer_close(void* device_ptr) {
    if (!device_ptr) return ELVIS_ERR_INVALID_ARG;
    const cudaError_t e = cudaIpcCloseMemHandle(device_ptr);
    return e == cudaSuccess ? ELVIS_OK : cuda_fail(e);
}

extern "C" int elvis_peer_free(void* device_ptr) {
    if (!device_ptr) return ELVIS_ERR_INVALID_ARG;
    const cudaError_t e = cudaFree(device_ptr);
    return e == cudaSuccess ? ELVIS_OK : cuda_fail(e);
}

extern "C" int elvis_peer_put(void* peer_dst, const void* src, int64_t bytes, uint32_t* peer_flag, uint32_t seq, elvis_stream_t stream) {
    if (!peer_dst || !src || bytes <= 0 || !peer_flag) return ELVIS_ERR_INVALID_ARG;
    cudaStream_t st = as_stream(stream);
    const cudaError_t e = cudaMemcpyAsync(peer_dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, st);   // copy engine over NVLink
    if (e != cudaSuccess) return cuda_fail(e);
    flag_write_kernel<<<1, 1, 0, st>>>(peer_flag, seq);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_peer_signal(uint32_t* peer_flag, uint32_t seq, elvis_stream_t stream) {
    if (!peer_flag) return ELVIS_ERR_INVALID_ARG;
    flag_write_kernel<<<1, 1, 0, as_stream(stream)>>>(peer_flag, seq);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_peer_wait(const uint32_t* local_flag, uint32_t seq, int32_t* error_word, elvis_stream_t stream) {
    if (!local_flag) return ELVIS_ERR_INVALID_ARG;
    flag_wait_kernel<<<1, 1, 0, as_stream(stream)>>>(local_flag, seq, error_word);
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}

extern "C" int elvis_peer_allreduce_minmax(void* inout, int32_t dtype, int32_t n_values, int32_t rank, int32_t world,
                                           void* const* host_mailboxes, int32_t slot, uint32_t seq, int32_t* error_word,
                                           elvis_stream_t stream) {
    if (!inout || !host_mailboxes || n_values <= 0 || n_values > kMaxValues || world <= 0 || world > kMaxRanks || rank < 0 || rank >= world ||
        slot < 0 || slot >= 4)
        return ELVIS_ERR_INVALID_ARG;
    Mailboxes mb;
    for (int r = 0; r < kMaxRanks; ++r) {
        uint8_t* base = static_cast<uint8_t*>(host_mailboxes[r < world ? r : rank]);
        if (!base) return ELVIS_ERR_INVALID_ARG;
        mb.values[r] = reinterpret_cast<double*>(base);
        mb.flags[r] = reinterpret_cast<uint32_t*>(base + (size_t)4 * kMaxRanks * kMaxValues * sizeof(double));
    }
    cudaStream_t st = as_stream(stream);
    if (dtype == ELVIS_F32)
        allreduce_minmax_kernel<float><<<1, 32, 0, st>>>(static_cast<float*>(inout), n_values, rank, world, mb, slot, seq, error_word);
    else if (dtype == ELVIS_F64)
        allreduce_minmax_kernel<double><<<1, 32, 0, st>>>(static_cast<double*>(inout), n_values, rank, world, mb, slot, seq, error_word);
    else
        return ELVIS_ERR_INVALID_ARG;
    ELVIS_CHECK_LAUNCH();
    return ELVIS_OK;
}
