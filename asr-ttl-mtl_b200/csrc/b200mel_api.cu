// C ABI of the B200 log-mel front-end (include/b200mel.h): plan construction, argument
// validation, L2-chunked launch sequencing and the host-buffer streaming pipeline.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/b200mel.h"
#include "kernels.h"

namespace b200mel {

static std::atomic<uint64_t> g_launches{0};
uint64_t launches_so_far() { return g_launches.load(std::memory_order_relaxed); }
void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static thread_local char tl_cuda_error[256] = "";

// ---- optional per-launch event timing (bench.py roofline) ----
struct ProfileRecord { cudaEvent_t start, stop; int kind; };
static std::atomic<int> g_profile_on{0};
static std::mutex g_profile_mutex;
static std::vector<ProfileRecord> g_profile_records;

ProfileScope::ProfileScope(int kind, cudaStream_t stream) : stream_(stream), kind_(kind) {
    if (!g_profile_on.load(std::memory_order_relaxed)) return;
    if (cudaEventCreate(&start_) != cudaSuccess || cudaEventCreate(&stop_) != cudaSuccess) { start_ = stop_ = nullptr; return; }
    cudaEventRecord(start_, stream_);
}
ProfileScope::~ProfileScope() {
    if (start_ == nullptr) return;
    cudaEventRecord(stop_, stream_);
    std::lock_guard<std::mutex> lock(g_profile_mutex);
    g_profile_records.push_back({start_, stop_, kind_});
}

static int cuda_fail(cudaError_t e, const char* where) {
    std::snprintf(tl_cuda_error, sizeof(tl_cuda_error), "%s: %s", where, cudaGetErrorString(e));
    return B200MEL_ERR_CUDA;
}
#define B200_CUDA(call)                                       \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

// One reusable slot of the host-buffer pipeline: device staging + its own stream.
struct HostSlot {
    cudaStream_t stream = nullptr;
    void* d_in = nullptr;
    float* d_out = nullptr;
    int32_t* d_len = nullptr;
    void* d_ws = nullptr;
    size_t in_bytes = 0, out_bytes = 0, len_bytes = 0, ws_bytes = 0;
};

constexpr int kHostSlots = 3;

}  // namespace b200mel

struct b200mel_plan {
    int device = -1;
    int n_mels = 0;
    int n_rows = 0;  // rows of the mel partial-sum tile (DeviceTables::n_rows)
    b200mel::DeviceTables* d_tables = nullptr;
    b200mel::TcTables* d_tc_tables = nullptr;  // operands of the tcgen05 variant (nullptr: the filter matrix is not the
                                               // Whisper bank that variant's epilogue is compiled for)
    std::mutex host_mutex;  // the host pipeline's staging buffers are per plan
    b200mel::HostSlot slots[b200mel::kHostSlots];
};

namespace b200mel {

static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// An utterance longer than this many 32-frame tiles (~5 min of audio) is normalised by the
// stand-alone pass-2 kernel instead of by the single CTA that finishes it.
constexpr int64_t kMaxFusedNormTiles = 1024;

static void free_slot(HostSlot& s) {
    if (s.d_in) cudaFree(s.d_in);
    if (s.d_out) cudaFree(s.d_out);
    if (s.d_len) cudaFree(s.d_len);
    if (s.d_ws) cudaFree(s.d_ws);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = HostSlot();
}

static cudaError_t ensure(void** p, size_t* have, size_t need) {
    if (*have >= need) return cudaSuccess;
    if (*p) { cudaError_t e = cudaFree(*p); *p = nullptr; *have = 0; if (e != cudaSuccess) return e; }
    cudaError_t e = cudaMalloc(p, need);
    if (e == cudaSuccess) *have = need;
    return e;
}

}  // namespace b200mel

using namespace b200mel;

extern "C" {

int b200mel_abi_version(void) { return B200MEL_ABI_VERSION; }

const char* b200mel_status_string(int status) {
    switch (status) {
        case B200MEL_OK: return "ok";
        case B200MEL_ERR_NULL_POINTER: return "null pointer argument";
        case B200MEL_ERR_BAD_N_MELS: return "Unsupported n_mels";
        case B200MEL_ERR_TOO_SHORT: return "audio too short for reflect padding (need more than 200 samples)";
        case B200MEL_ERR_BAD_ARGUMENT: return "bad argument";
        case B200MEL_ERR_BAD_FILTERS: return "mel filterbank is not a set of contiguous bands below bin 200";
        case B200MEL_ERR_CUDA: return "CUDA error";
        case B200MEL_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
        default: return "unknown status";
    }
}

const char* b200mel_last_cuda_error(void) { return tl_cuda_error; }

uint64_t b200mel_launch_count(void) { return launches_so_far(); }

unsigned b200mel_kernel_fault(unsigned* cta) { return tc_kernel_fault(cta); }

int b200mel_profile_enable(int on) {
    g_profile_on.store(on ? 1 : 0, std::memory_order_relaxed);
    return B200MEL_OK;
}

int b200mel_profile_collect(double* ms_by_kind, uint64_t* launches_by_kind) {
    if (ms_by_kind == nullptr || launches_by_kind == nullptr) return B200MEL_ERR_NULL_POINTER;
    for (int k = 0; k < B200MEL_PROFILE_KINDS; ++k) { ms_by_kind[k] = 0.0; launches_by_kind[k] = 0; }
    std::vector<ProfileRecord> records;
    {
        std::lock_guard<std::mutex> lock(g_profile_mutex);
        records.swap(g_profile_records);
    }
    int result = B200MEL_OK;
    for (const ProfileRecord& r : records) {
        float ms = 0.f;
        cudaError_t e = cudaEventSynchronize(r.stop);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.start, r.stop);
        if (e == cudaSuccess) {
            const int k = (r.kind >= 0 && r.kind < B200MEL_PROFILE_KINDS) ? r.kind : B200MEL_PROFILE_KINDS - 1;
            ms_by_kind[k] += ms;
            launches_by_kind[k] += 1;
        } else if (result == B200MEL_OK) {
            result = cuda_fail(e, "b200mel_profile_collect");
        }
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
    }
    return result;
}

int b200mel_frames(int64_t n_samples, int64_t right_zero_pad, int64_t* n_frames) {
    if (n_frames == nullptr) return B200MEL_ERR_NULL_POINTER;
    if (n_samples < 0) return B200MEL_ERR_BAD_ARGUMENT;
    const int64_t total = n_samples + (right_zero_pad > 0 ? right_zero_pad : 0);
    if (total <= kHalfWin) return B200MEL_ERR_TOO_SHORT;
    *n_frames = total / kHop;
    return B200MEL_OK;
}

int b200mel_plan_create(int n_mels, const float* filters_host, b200mel_plan** plan_out) {
    if (filters_host == nullptr || plan_out == nullptr) return B200MEL_ERR_NULL_POINTER;
    *plan_out = nullptr;
    if (n_mels != 80 && n_mels != 128) return B200MEL_ERR_BAD_N_MELS;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return B200MEL_ERR_NO_DEVICE;
    }
    std::vector<DeviceTables> host(1);
    const int st = build_tables(n_mels, filters_host, host.data());
    if (st != B200MEL_OK) return st;
    std::vector<TcTables> tc_host(1);
    const bool tc_ok = build_tc_tables(n_mels, filters_host, tc_host.data()) == B200MEL_OK;
    b200mel_plan* plan = new (std::nothrow) b200mel_plan();
    if (plan == nullptr) return B200MEL_ERR_BAD_ARGUMENT;
    plan->n_mels = n_mels;
    plan->n_rows = host[0].n_rows;
    cudaError_t e = cudaGetDevice(&plan->device);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&plan->d_tables), sizeof(DeviceTables));
    if (e == cudaSuccess) e = cudaMemcpy(plan->d_tables, host.data(), sizeof(DeviceTables), cudaMemcpyHostToDevice);
    if (tc_ok) {
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&plan->d_tc_tables), sizeof(TcTables));
        if (e == cudaSuccess) e = cudaMemcpy(plan->d_tc_tables, tc_host.data(), sizeof(TcTables), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        if (plan->d_tc_tables) cudaFree(plan->d_tc_tables);
        if (plan->d_tables) cudaFree(plan->d_tables);
        delete plan;
        return cuda_fail(e, "b200mel_plan_create");
    }
    *plan_out = plan;
    return B200MEL_OK;
}

int b200mel_plan_destroy(b200mel_plan* plan) {
    if (plan == nullptr) return B200MEL_OK;
    for (auto& s : plan->slots) free_slot(s);
    if (plan->d_tc_tables) cudaFree(plan->d_tc_tables);
    if (plan->d_tables) cudaFree(plan->d_tables);
    delete plan;
    return B200MEL_OK;
}

int b200mel_plan_n_mels(const b200mel_plan* plan) { return plan ? plan->n_mels : 0; }

// workspace = [max keys: batch u32][completion counters: batch u32][tile queue head: 1 u32][min keys: batch u32]
static size_t workspace_words(int64_t batch) { return 3 * static_cast<size_t>(batch) + 1; }
size_t b200mel_workspace_bytes(int64_t batch) {
    if (batch < 1) batch = 1;
    return round_up(workspace_words(batch) * sizeof(uint32_t), 256);
}

// ... followed, with B200MEL_FLAG_TILE_KEYS, by [tile keys: batch * ceil(T / 128) * 2 u32] at the next 256-byte boundary
static_assert(kTcTileFrames == 128, "tile-key layout");
static int64_t tc_tiles_per_clip(int64_t n_frames) { return (n_frames + kTcTileFrames - 1) / kTcTileFrames; }
size_t b200mel_workspace_bytes_tiles(int64_t batch, int64_t n_frames) {
    if (batch < 1) batch = 1;
    if (n_frames < 1) n_frames = 1;
    return b200mel_workspace_bytes(batch) + round_up(static_cast<size_t>(batch) * tc_tiles_per_clip(n_frames) * 2 * sizeof(uint32_t), 256);
}

int b200mel_normalise_device(float* out, const void* workspace, int64_t batch, int64_t elems_per_clip,
                             unsigned flags, void* stream) {
    if (out == nullptr || workspace == nullptr) return B200MEL_ERR_NULL_POINTER;
    if (batch < 0 || elems_per_clip < 0) return B200MEL_ERR_BAD_ARGUMENT;
    B200_CUDA(launch_normalise(out, static_cast<const uint32_t*>(workspace), batch, elems_per_clip,
                               (flags & B200MEL_FLAG_GLOBAL_MAX) ? 1 : 0, static_cast<cudaStream_t>(stream)));
    return B200MEL_OK;
}

int b200mel_logmel_device(const b200mel_plan* plan, const void* audio, int dtype, int64_t batch,
                          int64_t n_samples, int64_t stride_b, const int32_t* lengths,
                          int64_t right_zero_pad, float* out, void* workspace, unsigned flags,
                          int variant, void* stream_v) {
    if (plan == nullptr || out == nullptr || workspace == nullptr) return B200MEL_ERR_NULL_POINTER;
    if (dtype != B200MEL_F32 && dtype != B200MEL_S16) return B200MEL_ERR_BAD_ARGUMENT;
    if (batch < 0 || n_samples < 0 || stride_b < 0) return B200MEL_ERR_BAD_ARGUMENT;
    // the variant ncu picked: the tcgen05 kernel, unless this plan's filterbank could not be turned into its
    // compile-time band tables - then the shared-memory FFT kernel, which takes any banded filterbank
    if (variant == B200MEL_VARIANT_AUTO) variant = plan->d_tc_tables != nullptr ? B200MEL_VARIANT_TCGEN05 : B200MEL_VARIANT_FFT;
    if (variant != B200MEL_VARIANT_FFT && variant != B200MEL_VARIANT_TCGEN05) return B200MEL_ERR_BAD_ARGUMENT;
    if (variant == B200MEL_VARIANT_TCGEN05 && plan->d_tc_tables == nullptr) return B200MEL_ERR_BAD_FILTERS;
    int64_t n_frames = 0;
    const int st = b200mel_frames(n_samples, right_zero_pad, &n_frames);
    if (st != B200MEL_OK) return st;
    if (n_frames > 0x7fffffff) return B200MEL_ERR_BAD_ARGUMENT;
    if (batch == 0) return B200MEL_OK;
    if (audio == nullptr && n_samples > 0) return B200MEL_ERR_NULL_POINTER;
    if (batch > 1 && stride_b < n_samples) return B200MEL_ERR_BAD_ARGUMENT;

    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    const int global_max = (flags & B200MEL_FLAG_GLOBAL_MAX) ? 1 : 0;
    uint32_t* keys = static_cast<uint32_t*>(workspace);
    const bool tile_keys = (flags & B200MEL_FLAG_TILE_KEYS) && variant == B200MEL_VARIANT_TCGEN05;
    // one memset for everything the kernels count in: the per-utterance words and, right behind them, the per-tile keys
    B200_CUDA(cudaMemsetAsync(keys, 0, tile_keys ? b200mel_workspace_bytes_tiles(batch, n_frames) : workspace_words(batch) * sizeof(uint32_t), stream));

    const int64_t elems_per_clip = static_cast<int64_t>(plan->n_mels) * n_frames;
    const int64_t tiles_per_clip = (n_frames + kTileFrames - 1) / kTileFrames;

    LogmelArgs a;
    a.audio = audio;
    a.stride_b = stride_b;
    a.n_samples = n_samples;
    a.total = n_samples + (right_zero_pad > 0 ? right_zero_pad : 0);
    a.lengths = lengths;
    a.batch = batch;
    a.n_frames = static_cast<int>(n_frames);
    a.n_mels = plan->n_mels;
    a.out = out;
    a.out_f16 = (flags & B200MEL_FLAG_OUT_F16) ? 1 : 0;
    a.max_keys = keys;
    a.done_counters = keys + batch;
    a.tile_counter = keys + 2 * batch;
    a.min_keys = keys + 2 * batch + 1;
    a.tile_keys = nullptr;
    if (tile_keys) a.tile_keys = reinterpret_cast<uint32_t*>(static_cast<char*>(workspace) + b200mel_workspace_bytes(batch));
    a.global_max = global_max;
    // FFT variant: one max per utterance (or a single utterance, where the call's max is the utterance's) is normalised
    // inside the kernel by the CTA that finishes the utterance, unless the utterance is very long; the tcgen05 variant
    // always runs its finish kernel right behind
    a.fused_norm = (variant == B200MEL_VARIANT_FFT && (!global_max || batch == 1) && tiles_per_clip <= kMaxFusedNormTiles) ? 1 : 0;
    a.n_rows = plan->n_rows;
    a.tables = plan->d_tables;
    if (a.out_f16 && variant != B200MEL_VARIANT_TCGEN05) return B200MEL_ERR_BAD_ARGUMENT;
    if ((flags & B200MEL_FLAG_DEFER_CLAMP) && (variant != B200MEL_VARIANT_TCGEN05 || a.out_f16)) return B200MEL_ERR_BAD_ARGUMENT;
    if (variant == B200MEL_VARIANT_TCGEN05)
        B200_CUDA(launch_tc_pass1(a, plan->d_tc_tables, dtype, stream));
    else
        B200_CUDA(launch_fft_fused(a, dtype, stream));
    if (variant == B200MEL_VARIANT_TCGEN05) {
        if (!(flags & B200MEL_FLAG_DEFER_CLAMP)) B200_CUDA(launch_tc_finish(a, stream));
    } else if (!a.fused_norm) B200_CUDA(launch_normalise(out, keys, batch, elems_per_clip, global_max, stream));
    return B200MEL_OK;
}

namespace {
int stem_conv1(const float* mel, const void* workspace, unsigned flags, int64_t batch, int n_mels, int64_t n_frames, const float* weight,
               const float* bias, int n_state, float* out, void* out_fm16, void* stream) {
    if (mel == nullptr || weight == nullptr || bias == nullptr || (out == nullptr && out_fm16 == nullptr)) return B200MEL_ERR_NULL_POINTER;
    if (n_mels != 80) return B200MEL_ERR_BAD_N_MELS;
    if (batch < 0 || n_frames < 0 || n_frames > 0x7ffffff0 || n_state <= 0 || n_state % 128 != 0 || n_state > 128 * 148)
        return B200MEL_ERR_BAD_ARGUMENT;
    const uint32_t* keys = static_cast<const uint32_t*>(workspace);
    const uint32_t* tile_keys = nullptr;
    if (keys != nullptr && (flags & B200MEL_FLAG_TILE_KEYS))
        tile_keys = reinterpret_cast<const uint32_t*>(static_cast<const char*>(workspace) + b200mel_workspace_bytes(batch));
    B200_CUDA(launch_stem_conv1_gelu(mel, keys, tile_keys, (flags & B200MEL_FLAG_GLOBAL_MAX) ? 1 : 0, batch, static_cast<int>(n_frames),
                                     weight, bias, n_state, out, out_fm16, static_cast<cudaStream_t>(stream)));
    return B200MEL_OK;
}
}  // namespace

int b200mel_stem_conv1_gelu_device(const float* mel, const void* workspace, unsigned flags, int64_t batch, int n_mels,
                                   int64_t n_frames, const float* weight, const float* bias, int n_state, float* out,
                                   void* stream) {
    if (out == nullptr) return B200MEL_ERR_NULL_POINTER;
    return stem_conv1(mel, workspace, flags, batch, n_mels, n_frames, weight, bias, n_state, out, nullptr, stream);
}

int b200mel_stem_conv1_gelu_fm16_device(const float* mel, const void* workspace, unsigned flags, int64_t batch, int n_mels,
                                        int64_t n_frames, const float* weight, const float* bias, int n_state, void* out_fm16,
                                        void* stream) {
    if (out_fm16 == nullptr) return B200MEL_ERR_NULL_POINTER;
    return stem_conv1(mel, workspace, flags, batch, n_mels, n_frames, weight, bias, n_state, nullptr, out_fm16, stream);
}

int b200mel_stem_conv2_gelu_device(const void* h_fm16, int64_t batch, int64_t frames_padded, const void* weight_f16, const float* bias,
                                   const float* positional_embedding, int n_state, void* out, unsigned flags, void* stream) {
    if (flags & ~B200MEL_FLAG_OUT_F16) return B200MEL_ERR_BAD_ARGUMENT;
    if (h_fm16 == nullptr || weight_f16 == nullptr || bias == nullptr || out == nullptr) return B200MEL_ERR_NULL_POINTER;
    if (batch < 0 || frames_padded < 0 || frames_padded % 2 != 0 || frames_padded > 0x7ffffff0 || n_state <= 0 || n_state % 128 != 0 ||
        n_state > 128 * 148 || batch > 0x7fffffff)
        return B200MEL_ERR_BAD_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(h_fm16) | reinterpret_cast<uintptr_t>(weight_f16)) % 16 != 0) return B200MEL_ERR_BAD_ARGUMENT;
    B200_CUDA(launch_stem_conv2_gelu(h_fm16, batch, static_cast<int>(frames_padded), weight_f16, bias, positional_embedding, n_state, out,
                                     (flags & B200MEL_FLAG_OUT_F16) ? 1 : 0, static_cast<cudaStream_t>(stream)));
    return B200MEL_OK;
}

int b200mel_mel_windows_device(const float* mel, int n_mels, int64_t n_frames, const int32_t* seeks, const int32_t* sizes,
                               int n_windows, int window_frames, void* out, unsigned flags, void* stream) {
    if (n_windows < 0 || n_mels < 0 || n_frames < 0 || window_frames < 0 || n_mels > 65535 || n_windows > 65535) return B200MEL_ERR_BAD_ARGUMENT;
    if (flags & ~B200MEL_FLAG_OUT_F16) return B200MEL_ERR_BAD_ARGUMENT;
    if (n_windows == 0 || n_mels == 0 || window_frames == 0) return B200MEL_OK;
    if (mel == nullptr || seeks == nullptr || out == nullptr) return B200MEL_ERR_NULL_POINTER;
    B200_CUDA(launch_mel_windows(mel, n_mels, n_frames, seeks, sizes, n_windows, window_frames, out, (flags & B200MEL_FLAG_OUT_F16) ? 1 : 0,
                                 static_cast<cudaStream_t>(stream)));
    return B200MEL_OK;
}

int b200mel_logmel_host(const b200mel_plan* plan_c, const void* audio_host, int dtype, int64_t batch,
                        int64_t n_samples, int64_t stride_b, const int32_t* lengths_host,
                        int64_t right_zero_pad, float* out_host, unsigned flags, int variant) {
    b200mel_plan* plan = const_cast<b200mel_plan*>(plan_c);
    if (plan == nullptr || out_host == nullptr) return B200MEL_ERR_NULL_POINTER;
    if (dtype != B200MEL_F32 && dtype != B200MEL_S16) return B200MEL_ERR_BAD_ARGUMENT;
    if (batch < 0 || n_samples < 0 || stride_b < 0) return B200MEL_ERR_BAD_ARGUMENT;
    int64_t n_frames = 0;
    const int st = b200mel_frames(n_samples, right_zero_pad, &n_frames);
    if (st != B200MEL_OK) return st;
    if (batch == 0) return B200MEL_OK;
    if (audio_host == nullptr && n_samples > 0) return B200MEL_ERR_NULL_POINTER;
    if (batch > 1 && stride_b < n_samples) return B200MEL_ERR_BAD_ARGUMENT;

    std::lock_guard<std::mutex> lock(plan->host_mutex);
    // the caller's current device is restored on every way out
    int caller_device = -1;
    B200_CUDA(cudaGetDevice(&caller_device));
    B200_CUDA(cudaSetDevice(plan->device));
    const size_t in_elem = dtype == B200MEL_F32 ? sizeof(float) : sizeof(int16_t);
    const int64_t elems_per_clip = static_cast<int64_t>(plan->n_mels) * n_frames;
    const bool global_max = (flags & B200MEL_FLAG_GLOBAL_MAX) != 0;
    const size_t out_elem = (flags & B200MEL_FLAG_OUT_F16) ? 2 : 4;
    // chunk so that copy-in, compute and copy-out of neighbouring chunks overlap; a global max
    // needs every un-normalised value on the device at once, so it runs as a single chunk
    int64_t chunk = global_max ? batch : 16;
    if (!global_max && n_samples < 480000) {
        const int64_t scale = 480000 / (n_samples > 0 ? n_samples : 1);
        chunk = chunk * (scale < 64 ? scale : 64);
    }
    if (chunk > batch) chunk = batch;
    const int slots = global_max ? 1 : kHostSlots;

    // one chunk: copy in, compute, copy out - all on the slot's stream.  A failure returns its status; the caller of this
    // lambda stops issuing and still drains every slot, so no copy is in flight into the caller's buffers on return.
    auto run_chunk = [&](HostSlot& s, int64_t c0, int64_t n) -> int {
        if (s.stream == nullptr) B200_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        // the slot's previous chunk must have drained before its buffers are reused / regrown
        B200_CUDA(cudaStreamSynchronize(s.stream));
        B200_CUDA(ensure(&s.d_in, &s.in_bytes, static_cast<size_t>(n) * n_samples * in_elem + 16));
        B200_CUDA(ensure(reinterpret_cast<void**>(&s.d_out), &s.out_bytes, static_cast<size_t>(n) * elems_per_clip * out_elem));
        B200_CUDA(ensure(&s.d_ws, &s.ws_bytes, b200mel_workspace_bytes_tiles(n, n_frames)));
        const char* src = static_cast<const char*>(audio_host) + static_cast<size_t>(c0) * stride_b * in_elem;
        if (stride_b == n_samples || n == 1) {
            B200_CUDA(cudaMemcpyAsync(s.d_in, src, static_cast<size_t>(n) * n_samples * in_elem, cudaMemcpyHostToDevice, s.stream));
        } else {
            B200_CUDA(cudaMemcpy2DAsync(s.d_in, n_samples * in_elem, src, stride_b * in_elem, n_samples * in_elem,
                                        static_cast<size_t>(n), cudaMemcpyHostToDevice, s.stream));
        }
        const int32_t* d_len = nullptr;
        if (lengths_host != nullptr) {
            B200_CUDA(ensure(reinterpret_cast<void**>(&s.d_len), &s.len_bytes, static_cast<size_t>(n) * sizeof(int32_t)));
            B200_CUDA(cudaMemcpyAsync(s.d_len, lengths_host + c0, static_cast<size_t>(n) * sizeof(int32_t),
                                      cudaMemcpyHostToDevice, s.stream));
            d_len = s.d_len;
        }
        const int st = b200mel_logmel_device(plan, s.d_in, dtype, n, n_samples, n_samples, d_len, right_zero_pad, s.d_out,
                                             s.d_ws, flags | B200MEL_FLAG_TILE_KEYS, variant, s.stream);
        if (st != B200MEL_OK) return st;
        B200_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(out_host) + static_cast<size_t>(c0) * elems_per_clip * out_elem, s.d_out,
                                  static_cast<size_t>(n) * elems_per_clip * out_elem,
                                  cudaMemcpyDeviceToHost, s.stream));
        return B200MEL_OK;
    };

    int result = B200MEL_OK;
    int64_t issued = 0;
    for (int64_t c0 = 0; c0 < batch && result == B200MEL_OK; c0 += chunk, ++issued)
        result = run_chunk(plan->slots[issued % slots], c0, (batch - c0 < chunk) ? batch - c0 : chunk);
    for (int i = 0; i < kHostSlots; ++i)
        if (plan->slots[i].stream) {
            cudaError_t e = cudaStreamSynchronize(plan->slots[i].stream);
            if (e != cudaSuccess && result == B200MEL_OK) result = cuda_fail(e, "cudaStreamSynchronize");
        }
    if (caller_device >= 0 && caller_device != plan->device) {
        cudaError_t e = cudaSetDevice(caller_device);
        if (e != cudaSuccess && result == B200MEL_OK) result = cuda_fail(e, "cudaSetDevice");
    }
    return result;
}

}  // extern "C"
