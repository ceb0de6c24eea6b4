/* b200mel — C ABI of the B200-native fused log-mel front-end.
 *
 * This is the drop-in boundary for the hot path of muhkemallgp/asr-ttl-mtl,
 * `whisper/audio.py` (reference file:line cited per entry point).  The reference
 * has no FFI of its own (pure Python calling torch operators); the module
 * namespace of whisper/audio.py is its operator API, so these entry points are
 * what a Python binding for that module needs: plain pointers, sizes and a
 * CUDA stream handle — no torch types.  The ctypes stub that binds them lives
 * in asr-ttl-mtl_b200/_native.py and is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns a b200mel_status (0 = ok); nothing throws.
 *   - device pointers are raw CUDA device addresses on the CURRENT device;
 *     ownership stays with the caller (torch allocates; see audio.py mirror).
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no host
 *     synchronisation happens inside the *_device entry points.
 *   - thread-safe: plans are immutable after creation; calls on different
 *     streams may run concurrently as long as they use different workspaces.
 */
#ifndef B200MEL_H_
#define B200MEL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MEL_ABI_VERSION 2

/* audio constants of whisper/audio.py:13-22 */
#define B200MEL_SAMPLE_RATE 16000
#define B200MEL_N_FFT 400
#define B200MEL_HOP_LENGTH 160
#define B200MEL_N_BINS 201

typedef enum b200mel_status {
    B200MEL_OK = 0,
    B200MEL_ERR_NULL_POINTER = 1,
    B200MEL_ERR_BAD_N_MELS = 2,   /* audio.py:103  assert n_mels in {80, 128}            */
    B200MEL_ERR_TOO_SHORT = 3,    /* torch.stft reflect pad needs n_samples + pad > 200  */
    B200MEL_ERR_BAD_ARGUMENT = 4, /* negative sizes, unknown dtype / variant, ...        */
    B200MEL_ERR_BAD_FILTERS = 5,  /* filterbank rows are not single contiguous bands     */
    B200MEL_ERR_CUDA = 6,         /* a CUDA runtime call or launch failed (see last_cuda_error) */
    B200MEL_ERR_NO_DEVICE = 7     /* no sm_100 device visible: there is NO CPU fallback  */
} b200mel_status;

typedef enum b200mel_dtype {
    B200MEL_F32 = 0, /* float32 waveform, any finite range                (audio.py:141) */
    B200MEL_S16 = 1  /* int16 PCM; scaled by 1/32768 in-register          (audio.py:62)  */
} b200mel_dtype;

typedef enum b200mel_variant {
    B200MEL_VARIANT_AUTO = 0,   /* the variant ncu picked: tcgen05        */
    B200MEL_VARIANT_FFT = 1,    /* shared-memory mixed-radix (20x20) real FFT, fp32 CUDA cores */
    B200MEL_VARIANT_TCGEN05 = 2 /* folded DFT as GEMMs on tcgen05 tensor cores: fp16 hi/lo operands, the three-product
                                   compensation of "3xTF32", fp32 accumulation, per-32-frame power-of-two pre-scale */
} b200mel_variant;

/* flags for b200mel_logmel_device / _host */
#define B200MEL_FLAG_GLOBAL_MAX 1u /* one max over the whole call: the reference's literal
                                      behaviour for a 2-D input (audio.py:155).  Default is one
                                      max per utterance (stack of per-clip calls).             */

#define B200MEL_FLAG_TILE_KEYS 2u  /* `workspace` was sized with b200mel_workspace_bytes_tiles: the tcgen05 kernel
                                      also keeps the extremes of every 128-frame tile there, so that the clamp at
                                      max - 8 (audio.py:155) re-touches only the tiles it changes: silent (zero-padded)
                                      tiles are filled, tiles wholly above the clamp are left alone.           */

#define B200MEL_FLAG_OUT_F16 4u    /* `out` holds IEEE half instead of float32: the values of the float32 result rounded
                                      to nearest (what transcribe.py:286 / decoding.py feed the fp16 model after their
                                      .to(dtype)); half the write bytes.  tcgen05 variant only - otherwise
                                      B200MEL_ERR_BAD_ARGUMENT.                                                   */

#define B200MEL_FLAG_DEFER_CLAMP 8u /* tcgen05 variant, float32 output: leave `out` BEFORE the clamp at max - 8 (audio.py:155) -
                                      (log10 + 4) / 4 un-clamped, all-zero tiles not written at all with B200MEL_FLAG_TILE_KEYS -
                                      for a consumer that applies the clamp on load from `workspace`:
                                      b200mel_stem_conv1_gelu_device.  Saves the clamp's second touch of the output.   */

typedef struct b200mel_plan b200mel_plan; /* opaque: filterbank bands + FFT tables on one device */

int b200mel_abi_version(void);
const char* b200mel_status_string(int status);
/* text of the last CUDA error seen by this thread (empty string if none) */
const char* b200mel_last_cuda_error(void);

/* Frame count rule of audio.py:145-149: T = (n_samples + max(right_zero_pad, 0)) / 160,
 * B200MEL_ERR_TOO_SHORT if n_samples + pad <= 200. */
int b200mel_frames(int64_t n_samples, int64_t right_zero_pad, int64_t* n_frames);

/* Replaces mel_filters(device, n_mels) + torch.hann_window (audio.py:91-107, :147) as the
 * kernel's constant operands.  filters_host: float32 [n_mels, 201] row-major in HOST memory.
 * The plan lives on the current CUDA device. */
int b200mel_plan_create(int n_mels, const float* filters_host, b200mel_plan** plan_out);
int b200mel_plan_destroy(b200mel_plan* plan);
int b200mel_plan_n_mels(const b200mel_plan* plan);

/* Bytes of device scratch b200mel_logmel_device needs for `batch` utterances. */
size_t b200mel_workspace_bytes(int64_t batch);
/* The same plus room for per-tile extremes (pass B200MEL_FLAG_TILE_KEYS with a workspace of this size);
 * n_frames from b200mel_frames. */
size_t b200mel_workspace_bytes_tiles(int64_t batch, int64_t n_frames);

/* Replaces log_mel_spectrogram's compute, audio.py:145-156, for a batch of utterances.
 *   audio      device, [batch, n_samples] of `dtype`, row pitch `stride_b` ELEMENTS
 *   lengths    device int32 [batch] or NULL: samples of each row that are real; the rest of the
 *              row counts as zeros whatever it holds (pad_or_trim semantics, audio.py:83-86) and is not
 *              read beyond the 128-frame tile the utterance ends in (the memory of the row must exist)
 *   right_zero_pad  the `padding` argument (audio.py:145-146); <= 0 is ignored
 *   out        device float32 [batch, n_mels, T] contiguous, T from b200mel_frames (IEEE half with
 *              B200MEL_FLAG_OUT_F16; the pointer is passed through the same parameter)
 *   workspace  device scratch of b200mel_workspace_bytes(batch), or of b200mel_workspace_bytes_tiles(batch, T)
 *              together with B200MEL_FLAG_TILE_KEYS
 * tcgen05 variant: the front-end kernel plus a finish kernel right behind it on the same stream (a few words per tile; it
 * re-touches only the tiles the clamp at max - 8 changes).  FFT variant: one launch when every utterance has its own max
 * and at most 1024 x 32 frames, otherwise a second, clamp-only pass follows.
 */
int b200mel_logmel_device(const b200mel_plan* plan, const void* audio, int dtype, int64_t batch,
                          int64_t n_samples, int64_t stride_b, const int32_t* lengths,
                          int64_t right_zero_pad, float* out, void* workspace, unsigned flags,
                          int variant, void* stream);

/* Second pass only: out = (max(out, g - 8) + 4) / 4 with g decoded from `workspace`
 * (audio.py:155-156).  Exposed for tests; b200mel_logmel_device already runs it. */
int b200mel_normalise_device(float* out, const void* workspace, int64_t batch, int64_t elems_per_clip,
                             unsigned flags, void* stream);

/* The consumer right behind the front-end: the encoder stem's first layer, model.py:179 + :193
 *   out = F.gelu(conv1(x)),  conv1 = Conv1d(n_mels, n_state, kernel_size=3, padding=1)
 * as an implicit GEMM on the tcgen05 tensor cores (TF32 operands, float32 accumulation - what torch's own conv does on
 * this GPU with allow_tf32, cudnn's default), exact (erf) GELU.
 *   mel        device float32 [batch, n_mels, n_frames]
 *   workspace  NULL: `mel` is a finished log-mel spectrogram.  Otherwise the workspace of the b200mel_logmel_device call
 *              that produced `mel` with B200MEL_FLAG_DEFER_CLAMP; `flags` = that call's (GLOBAL_MAX, TILE_KEYS): the
 *              clamp is applied while loading, never-written all-zero tiles are not read
 *   weight     device float32 [n_state, n_mels, 3] (torch Conv1d.weight), bias [n_state]
 *   out        device float32 [batch, n_state, n_frames]
 * n_mels must be 80 (else B200MEL_ERR_BAD_N_MELS), n_state a multiple of 128 (Whisper's: 384 tiny, 512 base, 768 small, 1024 medium, 1280 large). */
int b200mel_stem_conv1_gelu_device(const float* mel, const void* workspace, unsigned flags, int64_t batch, int n_mels,
                                   int64_t n_frames, const float* weight, const float* bias, int n_state, float* out,
                                   void* stream);

/* The whole encoder stem, model.py:193-197:
 *   x = F.gelu(conv1(mel)); x = F.gelu(conv2(x)); x = x.permute(0, 2, 1); x = x + positional_embedding
 * conv2 = Conv1d(n_state, n_state, kernel_size=3, stride=2, padding=1) (model.py:180), in two launches.
 *
 * b200mel_stem_conv1_gelu_fm16_device: b200mel_stem_conv1_gelu_device with its result left for conv2 -
 *   out_fm16   device IEEE half [batch, frames_padded, n_state], frames_padded = n_frames rounded up to even: frames
 *              major, channels contiguous.  For an odd n_frames the caller zeroes the last frame of every clip (the
 *              convolution's padding); the kernel writes frames < n_frames only.
 * b200mel_stem_conv2_gelu_device: the second layer as a GEMM on the tcgen05 tensor cores, IEEE-half operands (TF32's
 *   11-bit significand: the operand precision of torch's conv with allow_tf32 and of the reference's fp16 inference,
 *   transcribe.py:127), float32 accumulation, exact (erf) GELU -
 *   h_fm16     what b200mel_stem_conv1_gelu_fm16_device left (16-byte aligned)
 *   weight_f16 device IEEE half [3, n_state, n_state] = conv2.weight.permute(2, 0, 1) (tap, out channel, in channel)
 *   bias       device float32 [n_state]
 *   positional_embedding  device float32 [frames_padded / 2, n_state] or NULL (model.py:197; the reference asserts
 *              frames_padded / 2 == n_ctx = 1500)
 *   out        device float32 [batch, frames_padded / 2, n_state] - the layout behind the permute of model.py:195 - or, with
 *              flags = B200MEL_FLAG_OUT_F16, IEEE half: the float32 result rounded once (what `.to(x.dtype)` of model.py:197
 *              hands the blocks of a half-precision model, transcribe.py:127).  Other flags: B200MEL_ERR_BAD_ARGUMENT */
int b200mel_stem_conv1_gelu_fm16_device(const float* mel, const void* workspace, unsigned flags, int64_t batch, int n_mels,
                                        int64_t n_frames, const float* weight, const float* bias, int n_state, void* out_fm16,
                                        void* stream);
int b200mel_stem_conv2_gelu_device(const void* h_fm16, int64_t batch, int64_t frames_padded, const void* weight_f16, const float* bias,
                                   const float* positional_embedding, int n_state, void* out, unsigned flags, void* stream);

/* The window cut of the decoding loop, transcribe.py:282-286 (and :150 with seek 0):
 *   mel_segment = pad_or_trim(mel[:, seek : seek + segment_size], N_FRAMES).to(device).to(dtype)
 * for n_windows windows of one long utterance in ONE launch, written straight into zero-padded windows.
 *   mel        device float32 [n_mels, n_frames] (the utterance's finished log-mel spectrogram)
 *   seeks      device int32 [n_windows]: first frame of each window, >= 0
 *   sizes      device int32 [n_windows] or NULL: frames kept of each window (segment_size; NULL = window_frames);
 *              clipped to window_frames and to the end of `mel` like the Python slice
 *   out        device [n_windows, n_mels, window_frames], float32, or IEEE half with B200MEL_FLAG_OUT_F16 (the rounding of
 *              .to(torch.float16)); the frames behind a window's size are zeros
 * n_windows, n_mels <= 65535. */
int b200mel_mel_windows_device(const float* mel, int n_mels, int64_t n_frames, const int32_t* seeks, const int32_t* sizes,
                               int n_windows, int window_frames, void* out, unsigned flags, void* stream);

/* Host-buffer entry point (what a CPU-tensor caller of log_mel_spectrogram hits, audio.py:138-144):
 * audio_host / out_host are HOST pointers (pinned for full speed).  Copies in, computes and copies
 * out in chunks on internal streams so H2D, compute and D2H overlap; returns after the result
 * is complete in out_host. */
int b200mel_logmel_host(const b200mel_plan* plan, const void* audio_host, int dtype, int64_t batch,
                        int64_t n_samples, int64_t stride_b, const int32_t* lengths_host,
                        int64_t right_zero_pad, float* out_host, unsigned flags, int variant);

/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t b200mel_launch_count(void);

/* Diagnostic: non-zero if a hand-over between the warp roles of the tcgen05 kernel ever timed out on the current device
 * (a protocol bug, never a property of the data).  The kernel then ran to its end instead of hanging or trapping - the
 * caller's CUDA context stays usable - but that launch's output is garbage.  *cta (optional) receives the CTA.
 * Synchronises the device. */
unsigned b200mel_kernel_fault(unsigned* cta);

/* Per-kernel device timing for bench.py's roofline: while enabled, every kernel launch is
 * bracketed by CUDA events on its own stream.  b200mel_profile_collect synchronises those
 * events, adds their durations per kernel kind and clears the list.
 *   kinds: 0 = FFT-variant fused pass, 1 = normalise pass, 2 = tcgen05-variant fused pass, 3 = other
 *   ms_by_kind / launches_by_kind: arrays of B200MEL_PROFILE_KINDS entries */
#define B200MEL_PROFILE_KINDS 4
int b200mel_profile_enable(int on);
int b200mel_profile_collect(double* ms_by_kind, uint64_t* launches_by_kind);

#ifdef __cplusplus
}
#endif
#endif /* B200MEL_H_ */
