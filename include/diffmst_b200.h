/*
 * diffmst_b200 — C ABI of the B200-native Diff-MST hot path (libdiffmst_b200.so).
 *
 * The reference (sai-soum/Diff-MST) is pure Python; the "operator API" this library
 * stands behind is the set of Python call signatures cited per function below.  The
 * host side (diffmst_b200/*.py) binds these symbols with ctypes and wraps them in
 * torch.autograd.Function; INTEGRATION.md shows the binding a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - the caller owns every buffer, including the workspace; nothing is allocated,
 *     retained or freed across the ABI;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no
 *     function synchronises the device;
 *   - return value: 0 on success, a positive cudaError_t value if a CUDA call
 *     failed, DMST_EINVAL (-22) for an invalid argument;
 *   - audio rows are float32 with unit stride along time; `*_row_stride` is the
 *     distance in elements between consecutive rows (tracks or channels).
 */
#ifndef DIFFMST_B200_H
#define DIFFMST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMST_EINVAL (-22)

/* flag bits of the console (mst/modules.py:192-198 use_* arguments) */
#define DMST_USE_TRACK_INPUT_FADER 1u
#define DMST_USE_TRACK_EQ 2u
#define DMST_USE_TRACK_COMPRESSOR 4u
#define DMST_USE_TRACK_PANNER 8u
#define DMST_USE_MASTER_BUS 16u
#define DMST_USE_FX_BUS 32u /* accepted only when clear: fx bus is out of scope */
#define DMST_USE_OUTPUT_FADER 64u
#define DMST_WANT_MIXED_TRACKS 128u /* materialise (B,2,N,T) as modules.py:314 returns it */
#define DMST_WANT_GRAD_TRACKS 256u  /* backward also writes dL/dtracks (forward then keeps per-section checkpoints) */
#define DMST_BASIC_CONSOLE 512u     /* BasicMixConsole layout: 2 track params [gain_db, pan] */
#define DMST_FORWARD_ONLY 1024u     /* no backward will follow (torch.no_grad(), mst/mixing.py:72): no checkpoints */
/* DMST_WANT_MIXED_TRACKS, DMST_WANT_GRAD_TRACKS and DMST_FORWARD_ONLY shape the workspace: pass the same flags
 * to dmst_console_workspace_bytes, dmst_console_forward and the matching dmst_console_backward. */

/* number of control parameters (mst/modules.py:182-184) */
#define DMST_NUM_TRACK_PARAMS 27
#define DMST_NUM_FX_PARAMS 25
#define DMST_NUM_MASTER_PARAMS 26

/* Parameter ranges, same order as the normalised parameter vectors
 * (mst/modules.py:121-181 and the index map at :353-460).  lo/hi per entry. */
typedef struct dmst_ranges {
    float track_lo[DMST_NUM_TRACK_PARAMS], track_hi[DMST_NUM_TRACK_PARAMS];
    float master_lo[DMST_NUM_MASTER_PARAMS], master_hi[DMST_NUM_MASTER_PARAMS];
} dmst_ranges;

int dmst_version(void);
/* 1 if the library was built for the CUDA device target (always, for the shipped .so) */
int dmst_is_device_build(void);

/* ---- mix console: replaces AdvancedMixConsole.forward / forward_mix_console
 *      (mst/modules.py:186-314, 316-487) and the dasp_pytorch calls inside it ---- */

/* Bytes of workspace needed by dmst_console_forward/backward for this shape and these flags. */
size_t dmst_console_workspace_bytes(int B, int N, int T, unsigned flags);

/*
 * Forward.  tracks (B,N,T) -> mix (B,2,T) [+ mixed_tracks (B,2,N,T) when
 * DMST_WANT_MIXED_TRACKS].  track_params (B,N,27) and master_params (B,26) are the
 * NORMALISED (0..1) controller outputs, contiguous float32 (for DMST_BASIC_CONSOLE:
 * track_params is (B,N,2), master_params may be NULL).
 * status (int32[4], device): status[0] receives 0, or 1 + the flat index of the first
 * out-of-range parameter (track index space first, then 1000 + master index), which the
 * host turns into the reference's ValueError (mst/modules.py:86-89).
 * track_lookahead / master_lookahead: compressor look-ahead in samples (2048 / 1024 upstream,
 * mst/modules.py:250,304); multiples of 32, at most 4096 samples each.
 * The workspace keeps what backward needs; it must stay untouched until
 * dmst_console_backward for the same call has run.
 */
int dmst_console_forward(const float* tracks, long long tracks_batch_stride,
                         long long tracks_row_stride, const float* track_params,
                         const float* master_params, const dmst_ranges* ranges_host,
                         float sample_rate, int B, int N, int T, unsigned flags,
                         int track_lookahead, int master_lookahead, float* mix,
                         float* mixed_tracks, int* status, void* workspace,
                         size_t workspace_bytes, void* stream);

/*
 * Backward.  grad_mix (B,2,T) and optional grad_mixed_tracks (B,2,N,T) ->
 * grad_track_params (B,N,27 or B,N,2), grad_master_params (B,26) and, with
 * DMST_WANT_GRAD_TRACKS, grad_tracks (B,N,T contiguous).  `tracks` etc. must be the
 * arguments of the matching forward call.
 */
int dmst_console_backward(const float* tracks, long long tracks_batch_stride,
                          long long tracks_row_stride, const float* track_params,
                          const float* master_params, const dmst_ranges* ranges_host,
                          float sample_rate, int B, int N, int T, unsigned flags,
                          int track_lookahead, int master_lookahead, const float* grad_mix,
                          const float* grad_mixed_tracks, float* grad_track_params,
                          float* grad_master_params, float* grad_tracks, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Range check of any other normalised parameter block on the device (the fx-bus parameters of
 * mst/modules.py:462-466, which the kernels do not consume): status[0] = min(status[0], base + 1 + column) over
 * the entries of params (rows x np) outside [0, 1].  `status` must have been initialised by dmst_console_forward
 * of the same call (or hold 0x7f7f7f7f).  Convention: fx-bus block base = 500 (track 0, master 1000), so that the
 * smallest code is the reference's first offender in its traversal order (track, fx bus, master bus). */
int dmst_console_check_ranges(const float* params, int rows, int np, int base, int* status, void* stream);
/* The same test over a strided (rows, np) view (params may be NULL: no test), after which the verdict status[0] is
 * stored to host_status, one int of PINNED host memory (device-visible under unified addressing): the asynchronous
 * form of the reference's range test (mst/modules.py:86-89) costs one small kernel and no copy node. */
int dmst_console_report_ranges(const float* params, long long row_stride, int rows, int np, int base, int* status,
                               int* host_status, void* stream);

/* Sliding-window inference (mst/utils.py:121-166, run_diffmst): Hann-weighted overlap-add of one console window
 * into the running mix on the device.  out[r, t] += window_mix[r, t] * w(t) for t < n <= window_length, rows = bs * 2;
 * w = torch.hann_window(window_length) (periodic), its first half forced to 1 when first_window != 0
 * (mst/utils.py:151-157).  `out` points at the window's first output sample. */
int dmst_ola_hann_add(const float* window_mix, long long window_row_stride, float* out, long long out_row_stride,
                      int rows, int n, int window_length, int first_window, void* stream);

/* ---- measurement hooks (used by bench.py only) ----
 * dmst_profile_enable(n > 0): from now on every console chain-kernel launch is bracketed by a
 * pair of CUDA events on its stream (at most n per kernel kind); n <= 0 disables.
 * dmst_profile_read(kind, ms_host, capacity): synchronises on the recorded events of `kind`
 * (0 track forward, 1 master forward, 2 master backward, 3 track backward) and writes their
 * durations in milliseconds; returns the number written, or a negative value on error. */
int dmst_profile_enable(int max_records);
int dmst_profile_read(int kind, float* ms_host, int capacity);

/* ---- multi-resolution STFT loss: replaces auraloss.freq.MultiResolutionSTFTLoss.forward
 *      (instantiated at configs/models/naive.yaml:54-68 and mst/system.py:61-69, called at
 *      mst/system.py:332; algorithm in SURVEY.md Appendix B) ---- */

#define DMST_MRSTFT_MAX_RES 8
typedef struct dmst_mrstft_cfg {
    int n_res;
    int fft_size[DMST_MRSTFT_MAX_RES], hop_size[DMST_MRSTFT_MAX_RES],
        win_length[DMST_MRSTFT_MAX_RES];
    float w_sc, w_log_mag, w_lin_mag;
    float eps; /* clamp inside the sqrt, auraloss default 1e-8 */
} dmst_mrstft_cfg;

/* Workspace bytes (0 on invalid configuration).  Creates and caches the cuFFT plans. */
size_t dmst_mrstft_workspace_bytes(const dmst_mrstft_cfg* cfg_host, int rows, int T);
/*
 * x, y: (rows, T) float32 with row strides (rows = batch * channels, as auraloss views its
 * (B, C, T) arguments).  windows: the per-resolution windows concatenated (win_length[r]
 * floats each, device).  loss: float32[1 + 3*n_res] (device): [0] the loss, then
 * (L_sc, L_log, L_lin) per resolution.  When grad_x != NULL it receives d loss / d x,
 * (rows, T) contiguous; the target y gets no gradient.
 */
int dmst_mrstft_forward(const float* x, long long x_row_stride, const float* y,
                        long long y_row_stride, const float* windows,
                        const dmst_mrstft_cfg* cfg_host, int rows, int T, float* loss,
                        float* grad_x, void* workspace, size_t workspace_bytes, void* stream);
/*
 * Two-call form for an autograd node (the loss's upstream gradient only exists at backward time, which is how
 * torch differentiates auraloss's forward at mst/system.py:332 + Lightning's loss.backward()):
 * dmst_mrstft_forward_keep computes the loss (loss: float32[1], terms: float32[3*n_res] or NULL) and leaves the
 * per-frame gradients of every resolution in the workspace; dmst_mrstft_backward overlap-adds them into
 * grad_x = grad_loss[0] * d loss / d x ((rows, T) contiguous; grad_loss: device scalar, NULL means 1).  The
 * workspace must be kept untouched between the two calls.
 */
int dmst_mrstft_forward_keep(const float* x, long long x_row_stride, const float* y,
                             long long y_row_stride, const float* windows,
                             const dmst_mrstft_cfg* cfg_host, int rows, int T, float* loss,
                             float* terms, void* workspace, size_t workspace_bytes, void* stream);
int dmst_mrstft_backward(const float* windows, const dmst_mrstft_cfg* cfg_host, int rows, int T,
                         const float* grad_loss, float* grad_x, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ---- audio feature loss: replaces mst.loss.AudioFeatureLoss.forward (mst/loss.py:238-260)
 *      and the compute_* transforms (mst/loss.py:62-195) ---- */

size_t dmst_afl_workspace_bytes(int B, int T, int fft_size, int n_bands);
/*
 * input/target: (B,2,T) float32 with batch/channel strides.  bark_fb: (n_bands,
 * fft_size/2+1) row-major (the transpose of mst/filter.py's (n_freqs, n_barks)); window:
 * float32[fft_size]; weights_host: 5 floats.  losses: float32[5] (device) = weight_i * mse_i
 * in the reference's key order (rms, crest_factor, stereo_width, stereo_imbalance,
 * barkspectrum).  The workspace keeps what dmst_afl_backward needs.
 */
int dmst_afl_forward(const float* input, const float* target, long long batch_stride,
                     long long ch_stride, const float* bark_fb, const float* window,
                     const float* weights_host, int B, int T, int fft_size, int n_bands,
                     float* losses, void* workspace, size_t workspace_bytes, void* stream);
/* grad_input (B,2,T contiguous) = d(sum_i grad_w[i] * losses[i]) / d input; grad_w: float32[5]
 * (device) upstream gradients, NULL means ones. */
int dmst_afl_backward(const float* input, long long batch_stride, long long ch_stride,
                      const float* bark_fb, const float* window, const float* weights_host,
                      const float* grad_w, int B, int T, int fft_size, int n_bands,
                      float* grad_input, void* workspace, size_t workspace_bytes, void* stream);

/* ---- batch_stereo_peak_normalize (mst/utils.py:14-29): y = x / max(peak over (2,T), 1e-8) ---- */
int dmst_peak_normalize(const float* x, long long batch_stride, long long ch_stride, float* y,
                        int B, int T, void* stream);

/* ---- Cnn14 ConvBlock on the tensor cores (mst/panns.py:27-85: conv3x3 -> BatchNorm -> ReLU, twice,
 *      then average pooling).  Activations between these calls are NHWC float32 with a one-pixel zero
 *      border: (B, H+2, W+2, C).
 *
 *      TF32 operand contract.  The tensor core reads float32 bits and ignores the low 13 mantissa bits
 *      (truncation), so operands of the tensor-core entry points (dmst_conv3x3_forward[_ws] with Cin % 32 == 0
 *      and Cout % 64 == 0, dmst_conv3x3_wgrad) should already be rounded to nearest TF32 - as cuDNN's TF32
 *      convolutions (the reference's numerics, torch's default) do internally.  Every entry point below that
 *      PRODUCES such an operand rounds it (cvt.rna.tf32.f32): the weight repacks, dmst_conv_nchw_to_padded_nhwc,
 *      dmst_conv_affine_relu_to, dmst_conv_bn_relu_avgpool / dmst_conv_avgpool with padded NHWC output, the dz of
 *      the BatchNorm backward (all when the channel count is a multiple of 32), and the convolution / affine
 *      epilogues when bit 1 of `relu` is set.  dmst_conv_round_tf32 is for tensors that come from elsewhere. ---- */
/* dst = src rounded to nearest TF32 (n elements; in place allowed).  With B > 0 the tensor is a zero-bordered NHWC
 * (B, H+2, W+2, C) and its border is cleared as well; B == 0: plain element-wise. */
int dmst_conv_round_tf32(const float* src, float* dst, long long n, int B, int H, int W, int C, void* stream);
int dmst_conv_nchw_to_padded_nhwc(const float* x, float* y, int B, int C, int H, int W, void* stream);
/* nn.Conv2d weight (Cout, Cin, 3, 3) -> (9, Cout, Cin) */
int dmst_conv_repack_weights(const float* w, float* w9, int Cout, int Cin, void* stream);
/* the same weight as the operand of the input-gradient convolution: (9, Cin, Cout), taps flipped, so that
 * dL/dx = dmst_conv3x3_forward(dL/dz, w9t) with the channel counts swapped */
int dmst_conv_repack_weights_dgrad(const float* w, float* w9t, int Cout, int Cin, void* stream);
/* y = [relu](conv3x3(x) * scale[c] + shift[c]); scale/shift may be NULL (1 / 0).  TF32 tensor-core
 * path (tcgen05 + TMA) when Cin % 32 == 0 and Cout % 64 == 0, CUDA-core path otherwise (first layer).
 * `relu` is a mask: bit 0 = apply ReLU, bit 1 = round the output to TF32 (it feeds another convolution). */
int dmst_conv3x3_forward(const float* x_padded, const float* w9, const float* scale, const float* shift,
                         float* y_padded, int B, int H, int W, int Cin, int Cout, int relu, void* stream);
/* training-mode BatchNorm: batch mean and biased variance per channel of a raw conv output */
/* Same, with a caller-provided workspace (dmst_conv3x3_workspace_bytes, may be 0): layers with fewer
 * output tiles than SMs (the deep 8x8 / 2x4 layers) split their K loop over several CTAs and reduce the
 * partial sums in a fixed order. */
size_t dmst_conv3x3_workspace_bytes(int B, int H, int W, int Cin, int Cout);
int dmst_conv3x3_forward_ws(const float* x_padded, const float* w9, const float* scale, const float* shift,
                            float* y_padded, int B, int H, int W, int Cin, int Cout, int relu, void* workspace,
                            size_t workspace_bytes, void* stream);
size_t dmst_conv_stats_workspace_bytes(int B, int H, int W, int C);
int dmst_conv_channel_stats(const float* y_padded, int B, int H, int W, int C, float* mean, float* var_biased,
                            void* workspace, size_t workspace_bytes, void* stream);
int dmst_conv_affine_relu(float* y_padded, const float* scale, const float* shift, int B, int H, int W, int C,
                          int relu, void* stream);
/* F.avg_pool2d(x, (kh, kw)) of a padded NHWC tensor -> NCHW (B, C, H/kh, W/kw), or padded NHWC again */
int dmst_conv_avgpool(const float* x_padded, float* y, int B, int C, int H, int W, int kh, int kw,
                      int out_padded_nhwc, void* stream);

/* Weight gradient of the 3x3 convolution (what autograd computes for the nn.Conv2d layers of mst/panns.py:33-47,
 * i.e. torch.nn.grad.conv2d_weight) on the tensor cores (tcgen05, TF32, MN-major operands straight from the
 * NHWC tensors): dw[co][ci][tap] = sum over pixels of dz[p][co] * x[p + tap offset][ci], i.e. the gradient in
 * nn.Conv2d's own (Cout, Cin, 3, 3) layout; dz_padded must be zero on its border.  Needs Cin % 32 == 0 and Cout % 32 == 0, or Cin == 1 (the first layer: a streaming kernel, the layer is
 * memory bound); workspace_bytes returns 0 for other shapes. */
size_t dmst_conv3x3_wgrad_workspace_bytes(int B, int H, int W, int Cin, int Cout);
int dmst_conv3x3_wgrad(const float* x_padded, const float* dz_padded, float* dw, int B, int H, int W, int Cin,
                       int Cout, void* workspace, size_t workspace_bytes, void* stream);

/* Differentiable path of the same units (autograd of mst/panns.py:79-85).  y = relu(z * scale + shift) out of place
 * (the raw convolution output z is kept for backward); backward of y = relu(BatchNorm(z)): given dL/dy it writes
 * dL/dz (zero border), dL/dgamma and dL/dbeta, with scale = gamma * rstd, shift = beta - mean * scale;
 * batch_stats = 1 for training-mode statistics (the mean / variance terms of the BatchNorm gradient), 0 for running
 * statistics.  Workspace: dmst_conv_stats_workspace_bytes.  C % 4 == 0.  Backward of dmst_conv_avgpool. */
int dmst_conv_affine_relu_to(const float* z_padded, float* y_padded, const float* scale, const float* shift, int B,
                             int H, int W, int C, void* stream);
int dmst_conv_bn_relu_backward(const float* z_padded, const float* dy_padded, const float* scale, const float* shift,
                               const float* mean, const float* rstd, int batch_stats, int B, int H, int W, int C,
                               float* dz_padded, float* dgamma, float* dbeta, void* workspace,
                               size_t workspace_bytes, void* stream);
int dmst_conv_avgpool_backward(const float* dy, float* dx_padded, int B, int C, int H, int W, int kh, int kw,
                               int dy_padded_nhwc, void* stream);
/* The second unit of a ConvBlock with its pooling fused (mst/panns.py:80-85): y = avg_pool2d(relu(z * scale + shift))
 * without materialising the BatchNorm+ReLU output, and its backward from the gradient of the pooled tensor
 * (gathered on the fly; kh and kw powers of two).  Same statistics / workspace conventions as the unfused pair. */
int dmst_conv_bn_relu_avgpool(const float* z_padded, const float* scale, const float* shift, float* y, int B, int C,
                              int H, int W, int kh, int kw, int out_padded_nhwc, void* stream);
int dmst_conv_bn_relu_avgpool_backward(const float* z_padded, const float* dpooled, int kh, int kw,
                                       int dpooled_padded_nhwc, const float* scale, const float* shift,
                                       const float* mean, const float* rstd, int batch_stats, int B, int H, int W,
                                       int C, float* dz_padded, float* dgamma, float* dbeta, void* workspace,
                                       size_t workspace_bytes, void* stream);

/* Spectrogram front-end of the encoder, replaces the torch.stft / abs / pow lines of
 * SpectrogramEncoder.forward (mst/modules.py:787-800): x holds B*C waveforms of T samples (row r = b*C + c at
 * x + r*row_stride); window = n_fft Hann coefficients; the result (|STFT| + eps)^power is written as the
 * zero-bordered NHWC tensor (B, n_fft/2+1 + 2, 1 + T/hop + 2, C) that dmst_conv3x3_forward consumes (bins are
 * rows, frames are columns).  STFT semantics are torch.stft's defaults (centred, reflect padding, onesided,
 * unnormalised); requires T > n_fft/2 and n_fft a power of two. */
size_t dmst_spectrogram_workspace_bytes(int B, int C, int T, int n_fft, int hop);
int dmst_spectrogram_frontend(const float* x, long long row_stride, const float* window, int B, int C, int T,
                              int n_fft, int hop, float eps, float power, float* out_padded, void* workspace,
                              size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFMST_B200_H */
