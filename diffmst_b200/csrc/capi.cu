// extern "C" entry points of libdiffmst_b200.so (include/diffmst_b200.h).
#include "console_host.cuh"
#include "afl.cuh"
#include "conv_tc.cuh"
#include "mrstft.cuh"
#include "peaknorm.cuh"
#include "specfront.cuh"

extern "C" {

int dmst_version(void) { return 1; }
int dmst_is_device_build(void) { return DMST_DEVICE_BUILD; }

size_t dmst_console_workspace_bytes(int B, int N, int T, unsigned flags) {
    if (B <= 0 || N <= 0 || T <= 0) return 0;
    // sized for the largest look-ahead the kernels accept (one tile)
    return dmst::carve_console(nullptr, B, N, T, dmst::kTrackTile, dmst::kMasterTile, flags).total;
}

int dmst_console_forward(const float* tracks, long long tracks_batch_stride, long long tracks_row_stride,
                         const float* track_params, const float* master_params,
                         const dmst_ranges* ranges_host, float sample_rate, int B, int N, int T,
                         unsigned flags, int track_lookahead, int master_lookahead, float* mix,
                         float* mixed_tracks, int* status, void* workspace, size_t workspace_bytes,
                         void* stream) {
    dmst::ConsoleCall k{tracks, tracks_batch_stride, tracks_row_stride, track_params, master_params,
                        ranges_host, sample_rate, B, N, T, flags, track_lookahead, master_lookahead};
    return dmst::console_forward(k, mix, mixed_tracks, status, workspace, workspace_bytes,
                                 reinterpret_cast<cudaStream_t>(stream));
}

int dmst_console_backward(const float* tracks, long long tracks_batch_stride, long long tracks_row_stride,
                          const float* track_params, const float* master_params,
                          const dmst_ranges* ranges_host, float sample_rate, int B, int N, int T,
                          unsigned flags, int track_lookahead, int master_lookahead,
                          const float* grad_mix, const float* grad_mixed_tracks, float* grad_track_params,
                          float* grad_master_params, float* grad_tracks, void* workspace,
                          size_t workspace_bytes, void* stream) {
    dmst::ConsoleCall k{tracks, tracks_batch_stride, tracks_row_stride, track_params, master_params,
                        ranges_host, sample_rate, B, N, T, flags, track_lookahead, master_lookahead};
    return dmst::console_backward(k, grad_mix, grad_mixed_tracks, grad_track_params, grad_master_params,
                                  grad_tracks, workspace, workspace_bytes,
                                  reinterpret_cast<cudaStream_t>(stream));
}

// status[0] = min(status[0], base + 1 + column) over the entries of params (rows x np) outside [0, 1]
__global__ void range_check_kernel(const float* params, long long n, int np, int base, int* status) {
    int bad = 0x7fffffff;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = params[i];
        if (v < 0.0f || v > 1.0f) bad = min(bad, base + 1 + (int)(i % np));
    }
    if (bad != 0x7fffffff) atomicMin(status, bad);
}
int dmst_console_check_ranges(const float* params, int rows, int np, int base, int* status, void* stream) {
    if (!params || !status || rows <= 0 || np <= 0) return DMST_EINVAL;
    const long long n = (long long)rows * np;
    const int blocks = (int)((n + 255) / 256 > 64 ? 64 : (n + 255) / 256);
    DMST_LAUNCH(range_check_kernel, dim3(blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), params, n, np, base, status);
    return DMST_LAST_ERROR();
}

// Same test over a strided (rows x np) view (or nothing when params is NULL), then the verdict status[0] (which the
// forward call's own test has already updated, earlier on the stream) is stored to host_status: a word of pinned
// host memory, device-visible under unified addressing.  One kernel instead of a copy kernel, the test and a
// device-to-host copy node.
#ifndef DMST_EMULATE
__global__ void range_report_kernel(const float* params, long long row_stride, int rows, int np, int base, int* status,
                                    int* host_status) {
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0x7fffffff;
    __syncthreads();
    int bad = 0x7fffffff;
    if (params)
        for (int i = threadIdx.x; i < rows * np; i += blockDim.x) {
            const int r = i / np, c = i - r * np;
            const float v = params[(long long)r * row_stride + c];
            if (v < 0.0f || v > 1.0f) bad = min(bad, base + 1 + c);
        }
    if (bad != 0x7fffffff) atomicMin(&s_bad, bad);
    __syncthreads();
    if (threadIdx.x == 0) {
        int v = *reinterpret_cast<volatile int*>(status);
        if (s_bad < v) { v = s_bad; *status = v; }
        *reinterpret_cast<volatile int*>(host_status) = v;
    }
}
#endif
int dmst_console_report_ranges(const float* params, long long row_stride, int rows, int np, int base, int* status,
                               int* host_status, void* stream) {
    if (!status || !host_status || (params && (rows <= 0 || np <= 0))) return DMST_EINVAL;
#ifndef DMST_EMULATE
    range_report_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(params, row_stride, rows, np, base, status, host_status);
    return DMST_LAST_ERROR();
#else
    return DMST_EINVAL;
#endif
}

// Hann-weighted overlap-add of one console window into the running mix (mst/utils.py:151-163): out[r, off + t] +=
// win[r, t] * w(t), w = periodic Hann of length W (torch.hann_window), with the first half forced to 1 for the first window.
__global__ void ola_hann_add_kernel(const float* win, long long win_stride, float* out, long long out_stride, int rows,
                                    int n, int W, int first) {
    const int r = blockIdx.y;
    const float k = 6.283185307179586f / (float)W;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        float w = 0.5f - 0.5f * cosf(k * (float)t);
        if (first && t < W / 2) w = 1.0f;
        out[(long long)r * out_stride + t] += win[(long long)r * win_stride + t] * w;
    }
}
int dmst_ola_hann_add(const float* window_mix, long long window_row_stride, float* out, long long out_row_stride,
                      int rows, int n, int window_length, int first_window, void* stream) {
    if (!window_mix || !out || rows <= 0 || n < 0 || window_length <= 0 || n > window_length) return DMST_EINVAL;
    if (n == 0) return 0;
    const int bx = (n + 255) / 256 > 1024 ? 1024 : (n + 255) / 256;
    DMST_LAUNCH(ola_hann_add_kernel, dim3(bx, rows), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), window_mix,
                window_row_stride, out, out_row_stride, rows, n, window_length, first_window);
    return DMST_LAST_ERROR();
}

int dmst_profile_enable(int max_records) {
#ifndef DMST_EMULATE
    dmst::Profiler& p = dmst::profiler();
    if (max_records <= 0) { p.enabled = false; return 0; }
    if (p.cap < max_records) {
        for (int k = 0; k < dmst::Profiler::kKinds; ++k) {
            cudaEvent_t* n = new cudaEvent_t[2 * max_records];
            for (int i = 0; i < 2 * max_records; ++i) {
                if (p.ev[k] && i < 2 * p.cap) n[i] = p.ev[k][i];
                else if (cudaEventCreate(&n[i]) != cudaSuccess) return (int)cudaGetLastError();
            }
            delete[] p.ev[k];
            p.ev[k] = n;
        }
        p.cap = max_records;
    }
    for (int k = 0; k < dmst::Profiler::kKinds; ++k) p.count[k] = 0;
    p.enabled = true;
    return 0;
#else
    (void)max_records;
    return DMST_EINVAL;
#endif
}

int dmst_profile_read(int kind, float* ms_host, int capacity) {
#ifndef DMST_EMULATE
    dmst::Profiler& p = dmst::profiler();
    if (kind < 0 || kind >= dmst::Profiler::kKinds || !ms_host) return DMST_EINVAL;
    const int n = p.count[kind] < capacity ? p.count[kind] : capacity;
    for (int i = 0; i < n; ++i) {
        if (cudaEventSynchronize(p.ev[kind][2 * i + 1]) != cudaSuccess) return -1;
        if (cudaEventElapsedTime(&ms_host[i], p.ev[kind][2 * i], p.ev[kind][2 * i + 1]) != cudaSuccess) return -1;
    }
    return n;
#else
    (void)kind; (void)ms_host; (void)capacity;
    return DMST_EINVAL;
#endif
}

size_t dmst_mrstft_workspace_bytes(const dmst_mrstft_cfg* cfg, int rows, int T) {
#ifndef DMST_EMULATE
    if (!cfg || cfg->n_res <= 0 || cfg->n_res > DMST_MRSTFT_MAX_RES || rows <= 0 || T <= 0) return 0;
    dmst::MrWs w;
    if (dmst::mr_carve(nullptr, cfg, rows, T, &w) != 0) return 0;
    return w.total;
#else
    return 0;
#endif
}

int dmst_mrstft_forward(const float* x, long long x_row_stride, const float* y, long long y_row_stride,
                        const float* windows, const dmst_mrstft_cfg* cfg, int rows, int T, float* loss,
                        float* grad_x, void* workspace, size_t workspace_bytes, void* stream) {
#ifndef DMST_EMULATE
    return dmst::mrstft_run(x, x_row_stride, y, y_row_stride, windows, cfg, rows, T, loss, loss ? loss + 1 : nullptr,
                            grad_x, false, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
#else
    return DMST_EINVAL;
#endif
}

int dmst_mrstft_forward_keep(const float* x, long long x_row_stride, const float* y, long long y_row_stride,
                             const float* windows, const dmst_mrstft_cfg* cfg, int rows, int T, float* loss,
                             float* terms, void* workspace, size_t workspace_bytes, void* stream) {
#ifndef DMST_EMULATE
    return dmst::mrstft_run(x, x_row_stride, y, y_row_stride, windows, cfg, rows, T, loss, terms, nullptr, true,
                            workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
#else
    return DMST_EINVAL;
#endif
}

int dmst_mrstft_backward(const float* windows, const dmst_mrstft_cfg* cfg, int rows, int T, const float* grad_loss,
                         float* grad_x, void* workspace, size_t workspace_bytes, void* stream) {
#ifndef DMST_EMULATE
    return dmst::mrstft_backward_run(windows, cfg, rows, T, grad_loss, grad_x, workspace, workspace_bytes,
                                     reinterpret_cast<cudaStream_t>(stream));
#else
    return DMST_EINVAL;
#endif
}

size_t dmst_afl_workspace_bytes(int B, int T, int fft_size, int n_bands) {
#ifndef DMST_EMULATE
    dmst::AflWs w;
    if (dmst::afl_carve(nullptr, B, T, fft_size, n_bands, &w) != 0) return 0;
    return w.total;
#else
    return 0;
#endif
}

int dmst_afl_forward(const float* input, const float* target, long long batch_stride, long long ch_stride,
                     const float* bark_fb, const float* window, const float* weights_host, int B, int T,
                     int fft_size, int n_bands, float* losses, void* workspace, size_t workspace_bytes,
                     void* stream) {
#ifndef DMST_EMULATE
    return dmst::afl_forward(input, target, batch_stride, ch_stride, bark_fb, window, weights_host, B, T, fft_size,
                             n_bands, losses, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
#else
    return DMST_EINVAL;
#endif
}

int dmst_afl_backward(const float* input, long long batch_stride, long long ch_stride, const float* bark_fb,
                      const float* window, const float* weights_host, const float* grad_w, int B, int T,
                      int fft_size, int n_bands, float* grad_input, void* workspace, size_t workspace_bytes,
                      void* stream) {
#ifndef DMST_EMULATE
    return dmst::afl_backward(input, batch_stride, ch_stride, bark_fb, window, weights_host, grad_w, B, T, fft_size,
                              n_bands, grad_input, workspace, workspace_bytes,
                              reinterpret_cast<cudaStream_t>(stream));
#else
    return DMST_EINVAL;
#endif
}

int dmst_peak_normalize(const float* x, long long batch_stride, long long ch_stride, float* y, int B, int T,
                        void* stream) {
    return dmst::peak_normalize(x, batch_stride, ch_stride, y, B, T, reinterpret_cast<cudaStream_t>(stream));
}

// ---- Cnn14 ConvBlock pieces (tensor-core row) ----
#ifndef DMST_EMULATE
int dmst_conv_nchw_to_padded_nhwc(const float* x, float* y, int B, int C, int H, int W, void* stream) {
    if (!x || !y || B <= 0 || C <= 0 || H <= 0 || W <= 0) return DMST_EINVAL;
    const long long total = (long long)B * (H + 2) * (W + 2) * C;
    dmst::nchw_to_padded_nhwc_kernel<<<dmst::grid_for(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, B, C, H, W);
    return (int)cudaGetLastError();
}
int dmst_conv_repack_weights(const float* w, float* w9, int Cout, int Cin, void* stream) {
    if (!w || !w9 || Cout <= 0 || Cin <= 0) return DMST_EINVAL;
    dmst::repack_weights_kernel<<<dmst::grid_for(9LL * Cout * Cin), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(w, w9, Cout, Cin);
    return (int)cudaGetLastError();
}
int dmst_conv_round_tf32(const float* src, float* dst, long long n, int B, int H, int W, int C, void* stream) {
    if (!src || !dst || n <= 0) return DMST_EINVAL;
    const bool bordered = B > 0;
    if (bordered && (H <= 0 || W <= 0 || C <= 0 || n != (long long)B * (H + 2) * (W + 2) * C)) return DMST_EINVAL;
    dmst::round_tf32_kernel<<<dmst::grid_for(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        src, dst, n, bordered ? H + 2 : 0, bordered ? W + 2 : 0, C);
    return (int)cudaGetLastError();
}
int dmst_conv_repack_weights_dgrad(const float* w, float* w9t, int Cout, int Cin, void* stream) {
    if (!w || !w9t || Cout <= 0 || Cin <= 0) return DMST_EINVAL;
    dmst::repack_weights_dgrad_kernel<<<dim3((Cin + 31) / 32, (Cout + 31) / 32), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        w, w9t, Cout, Cin);
    return (int)cudaGetLastError();
}
int dmst_conv3x3_forward(const float* x_padded, const float* w9, const float* scale, const float* shift, float* y_padded,
                         int B, int H, int W, int Cin, int Cout, int relu, void* stream) {
    return dmst::conv3x3_forward(x_padded, w9, scale, shift, y_padded, B, H, W, Cin, Cout, relu,
                                 reinterpret_cast<cudaStream_t>(stream));
}
size_t dmst_conv3x3_workspace_bytes(int B, int H, int W, int Cin, int Cout) {
    if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return 0;
    return dmst::conv3x3_workspace_bytes(B, H, W, Cin, Cout);
}
int dmst_conv3x3_forward_ws(const float* x_padded, const float* w9, const float* scale, const float* shift, float* y_padded,
                            int B, int H, int W, int Cin, int Cout, int relu, void* workspace, size_t workspace_bytes,
                            void* stream) {
    return dmst::conv3x3_forward(x_padded, w9, scale, shift, y_padded, B, H, W, Cin, Cout, relu,
                                 reinterpret_cast<cudaStream_t>(stream), workspace, workspace_bytes);
}
size_t dmst_conv_stats_workspace_bytes(int B, int H, int W, int C) {
    const long long P = (long long)B * (H + 2) * (W + 2);
    const int rows = dmst::stat_rows(P);
    return (size_t)((P + rows - 1) / rows) * 2 * C * sizeof(float);
}
int dmst_conv_channel_stats(const float* y_padded, int B, int H, int W, int C, float* mean, float* var_biased,
                            void* workspace, size_t workspace_bytes, void* stream) {
    if (!y_padded || !mean || !var_biased || !workspace || (C & 3)) return DMST_EINVAL;
    if (workspace_bytes < dmst_conv_stats_workspace_bytes(B, H, W, C)) return DMST_EINVAL;
    const int P = B * (H + 2) * (W + 2);
    const int chunks = (P + dmst::stat_rows(P) - 1) / dmst::stat_rows(P);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    dmst::channel_partial_kernel<<<chunks, 256, 0, s>>>(y_padded, P, C, reinterpret_cast<float*>(workspace));
    dmst::channel_final_kernel<<<(C + 7) / 8, 256, 0, s>>>(reinterpret_cast<float*>(workspace), chunks, C,
                                                               (double)B * H * W, mean, var_biased);
    return (int)cudaGetLastError();
}
int dmst_conv_affine_relu(float* y_padded, const float* scale, const float* shift, int B, int H, int W, int C, int relu,
                          void* stream) {
    if (!y_padded || !scale || !shift) return DMST_EINVAL;
    const int P = B * (H + 2) * (W + 2);
    dmst::affine_relu_kernel<<<dmst::grid_for((long long)P * C), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        y_padded, P, H + 2, W + 2, C, scale, shift, relu);
    return (int)cudaGetLastError();
}
int dmst_conv_avgpool(const float* x_padded, float* y, int B, int C, int H, int W, int kh, int kw, int out_padded_nhwc,
                      void* stream) {
    if (!x_padded || !y || kh <= 0 || kw <= 0 || H / kh <= 0 || W / kw <= 0) return DMST_EINVAL;
    const long long total = (long long)B * (H / kh) * (W / kw) * C;
    dmst::avgpool_kernel<<<dmst::grid_for(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x_padded, y, B, C, H, W, kh, kw,
                                                                                                   out_padded_nhwc);
    return (int)cudaGetLastError();
}
size_t dmst_conv3x3_wgrad_workspace_bytes(int B, int H, int W, int Cin, int Cout) {
    if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return 0;
    return dmst::conv3x3_wgrad_workspace_bytes(B, H, W, Cin, Cout);
}
int dmst_conv3x3_wgrad(const float* x_padded, const float* dz_padded, float* dw, int B, int H, int W, int Cin, int Cout,
                       void* workspace, size_t workspace_bytes, void* stream) {
    return dmst::conv3x3_wgrad(x_padded, dz_padded, dw, B, H, W, Cin, Cout, workspace, workspace_bytes,
                               reinterpret_cast<cudaStream_t>(stream));
}
int dmst_conv_affine_relu_to(const float* z_padded, float* y_padded, const float* scale, const float* shift, int B, int H,
                             int W, int C, void* stream) {
    if (!z_padded || !y_padded || !scale || !shift || (C & 3)) return DMST_EINVAL;
    const int P = B * (H + 2) * (W + 2);
    dmst::affine_relu_to_kernel<<<dmst::grid_for((long long)P * (C / 4)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        z_padded, y_padded, P, H + 2, W + 2, C, scale, shift);
    return (int)cudaGetLastError();
}
static int bn_relu_backward_impl(const float* z_padded, const float* dy_padded, dmst::PoolGrad pg, const float* scale,
                                 const float* shift, const float* mean, const float* rstd, int batch_stats, int B, int H, int W,
                                 int C, float* dz_padded, float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    if (!z_padded || !scale || !shift || !mean || !rstd || !dz_padded || !dgamma || !dbeta || !workspace || (C & 3)) return DMST_EINVAL;
    if (workspace_bytes < dmst_conv_stats_workspace_bytes(B, H, W, C)) return DMST_EINVAL;
    const int P = B * (H + 2) * (W + 2);
    const int chunks = (P + dmst::stat_rows(P) - 1) / dmst::stat_rows(P);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    float* partial = reinterpret_cast<float*>(workspace);
    dmst::bn_relu_bwd_partial_kernel<<<chunks, 256, 0, s>>>(z_padded, dy_padded, pg, P, H + 2, W + 2, C, scale, shift, mean, rstd, partial);
    dmst::bn_relu_bwd_final_kernel<<<(C + 7) / 8, 256, 0, s>>>(partial, chunks, C, dgamma, dbeta);
    const float inv_count = batch_stats ? (float)(1.0 / ((double)B * H * W)) : 0.0f;
    dmst::bn_relu_bwd_apply_kernel<<<dmst::grid_for((long long)P * (C / 4)), 256, 0, s>>>(
        z_padded, dy_padded, pg, dz_padded, P, H + 2, W + 2, C, scale, shift, mean, rstd, dgamma, dbeta, inv_count);
    return (int)cudaGetLastError();
}
int dmst_conv_bn_relu_backward(const float* z_padded, const float* dy_padded, const float* scale, const float* shift,
                               const float* mean, const float* rstd, int batch_stats, int B, int H, int W, int C,
                               float* dz_padded, float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                               void* stream) {
    if (!dy_padded) return DMST_EINVAL;
    dmst::PoolGrad pg{nullptr, 0, 0, 0, 0, 0, 0.0f};
    return bn_relu_backward_impl(z_padded, dy_padded, pg, scale, shift, mean, rstd, batch_stats, B, H, W, C, dz_padded, dgamma,
                                 dbeta, workspace, workspace_bytes, stream);
}
static int log2_exact(int v) { int s = 0; while ((1 << s) < v) ++s; return (1 << s) == v ? s : -1; }
int dmst_conv_bn_relu_avgpool(const float* z_padded, const float* scale, const float* shift, float* y, int B, int C, int H,
                              int W, int kh, int kw, int out_padded_nhwc, void* stream) {
    if (!z_padded || !scale || !shift || !y || kh <= 0 || kw <= 0 || H / kh <= 0 || W / kw <= 0 || (C & 3)) return DMST_EINVAL;
    const long long total = (long long)B * (H / kh) * (W / kw) * (C / 4);
    dmst::bn_relu_avgpool_kernel<<<dmst::grid_for(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        z_padded, scale, shift, y, B, C, H, W, kh, kw, out_padded_nhwc);
    return (int)cudaGetLastError();
}
int dmst_conv_bn_relu_avgpool_backward(const float* z_padded, const float* dpooled, int kh, int kw, int dpooled_padded_nhwc,
                                       const float* scale, const float* shift, const float* mean, const float* rstd,
                                       int batch_stats, int B, int H, int W, int C, float* dz_padded, float* dgamma,
                                       float* dbeta, void* workspace, size_t workspace_bytes, void* stream) {
    const int sh = kh > 0 ? log2_exact(kh) : -1, sw = kw > 0 ? log2_exact(kw) : -1;
    if (!dpooled || sh < 0 || sw < 0 || H / kh <= 0 || W / kw <= 0) return DMST_EINVAL;
    dmst::PoolGrad pg{dpooled, sh, sw, H / kh, W / kw, dpooled_padded_nhwc, 1.0f / (float)(kh * kw)};
    return bn_relu_backward_impl(z_padded, nullptr, pg, scale, shift, mean, rstd, batch_stats, B, H, W, C, dz_padded, dgamma,
                                 dbeta, workspace, workspace_bytes, stream);
}
int dmst_conv_avgpool_backward(const float* dy, float* dx_padded, int B, int C, int H, int W, int kh, int kw,
                               int dy_padded_nhwc, void* stream) {
    if (!dy || !dx_padded || kh <= 0 || kw <= 0 || H / kh <= 0 || W / kw <= 0 || (C & 3)) return DMST_EINVAL;
    const long long total = (long long)B * (H + 2) * (W + 2) * (C / 4);
    dmst::avgpool_bwd_kernel<<<dmst::grid_for(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dy, dx_padded, B, C, H, W, kh, kw,
                                                                                                       dy_padded_nhwc);
    return (int)cudaGetLastError();
}
size_t dmst_spectrogram_workspace_bytes(int B, int C, int T, int n_fft, int hop) {
    dmst::SpecWs w;
    if (B <= 0 || C <= 0 || dmst::spec_carve(nullptr, B * C, T, n_fft, hop, &w) != 0) return 0;
    return w.total;
}
int dmst_spectrogram_frontend(const float* x, long long row_stride, const float* window, int B, int C, int T, int n_fft,
                              int hop, float eps, float power, float* out_padded, void* workspace,
                              size_t workspace_bytes, void* stream) {
    return dmst::spectrogram_frontend(x, row_stride, window, B, C, T, n_fft, hop, eps, power, out_padded, workspace,
                                      workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}
#else
size_t dmst_conv3x3_wgrad_workspace_bytes(int, int, int, int, int) { return 0; }
int dmst_conv3x3_wgrad(const float*, const float*, float*, int, int, int, int, int, void*, size_t, void*) { return DMST_EINVAL; }
int dmst_conv_affine_relu_to(const float*, float*, const float*, const float*, int, int, int, int, void*) { return DMST_EINVAL; }
int dmst_conv_bn_relu_backward(const float*, const float*, const float*, const float*, const float*, const float*, int, int, int,
                               int, int, float*, float*, float*, void*, size_t, void*) { return DMST_EINVAL; }
int dmst_conv_avgpool_backward(const float*, float*, int, int, int, int, int, int, int, void*) { return DMST_EINVAL; }
int dmst_conv_bn_relu_avgpool(const float*, const float*, const float*, float*, int, int, int, int, int, int, int, void*) { return DMST_EINVAL; }
int dmst_conv_bn_relu_avgpool_backward(const float*, const float*, int, int, int, const float*, const float*, const float*, const float*,
                                       int, int, int, int, int, float*, float*, float*, void*, size_t, void*) { return DMST_EINVAL; }
size_t dmst_spectrogram_workspace_bytes(int, int, int, int, int) { return 0; }
int dmst_spectrogram_frontend(const float*, long long, const float*, int, int, int, int, int, float, float, float*, void*,
                              size_t, void*) { return DMST_EINVAL; }
int dmst_conv_nchw_to_padded_nhwc(const float*, float*, int, int, int, int, void*) { return DMST_EINVAL; }
int dmst_conv_repack_weights(const float*, float*, int, int, void*) { return DMST_EINVAL; }
int dmst_conv_repack_weights_dgrad(const float*, float*, int, int, void*) { return DMST_EINVAL; }
int dmst_conv3x3_forward(const float*, const float*, const float*, const float*, float*, int, int, int, int, int, int, void*) { return DMST_EINVAL; }
size_t dmst_conv3x3_workspace_bytes(int, int, int, int, int) { return 0; }
int dmst_conv3x3_forward_ws(const float*, const float*, const float*, const float*, float*, int, int, int, int, int, int, void*, size_t, void*) { return DMST_EINVAL; }
size_t dmst_conv_stats_workspace_bytes(int, int, int, int) { return 0; }
int dmst_conv_channel_stats(const float*, int, int, int, int, float*, float*, void*, size_t, void*) { return DMST_EINVAL; }
int dmst_conv_affine_relu(float*, const float*, const float*, int, int, int, int, int, void*) { return DMST_EINVAL; }
int dmst_conv_round_tf32(const float*, float*, long long, int, int, int, int, void*) { return DMST_EINVAL; }
int dmst_conv_avgpool(const float*, float*, int, int, int, int, int, int, int, void*) { return DMST_EINVAL; }
#endif


#ifdef DMST_EMULATE
// TEST INFRASTRUCTURE (host-emulated build only, not declared in include/): the fused STFT + loss-sum kernel on
// host buffers; twiddle tables supplied by the caller
int dmst_emul_stft_loss(const float* x, const float* y, int rows, int T, int n, int hop, int win, const float* window,
                        const float* tw_m, const float* tw_n, float* X, float* PY, float eps, float* partial) {
    dmst::SfArgs a;
    memset(&a, 0, sizeof(a));
    a.x[0] = x; a.x[1] = y; a.row_stride[0] = T; a.row_stride[1] = T;
    a.vec_ok[0] = a.vec_ok[1] = (T % 2 == 0) && (hop % 2 == 0);
    a.rows = rows; a.T = T; a.n = n; a.hop = hop; a.win = win; a.frames = 1 + T / hop;
    a.window = window; a.win_vec_ok = (win == n);
    a.tw_m = reinterpret_cast<const float2*>(tw_m); a.tw_n = reinterpret_cast<const float2*>(tw_n);
    a.X = reinterpret_cast<float2*>(X); a.PY = PY; a.eps = eps; a.partial = partial; a.done = nullptr;
    return dmst::sf_launch(a, nullptr) ? 0 : DMST_EINVAL;
}
int dmst_emul_istft_grad(const float* X, const float* PY, float* dframes, int rows, int n, int frames, const float* tw_m,
                         const float* tw_n, float eps, const float* row_coef, const float* scal, int use_log, int use_lin) {
    dmst::SfGradArgs a;
    memset(&a, 0, sizeof(a));
    a.X = reinterpret_cast<const float2*>(X); a.PY = PY; a.dframes = dframes; a.rows = rows; a.n = n; a.frames = frames;
    a.tw_m = reinterpret_cast<const float2*>(tw_m); a.tw_n = reinterpret_cast<const float2*>(tw_n);
    a.eps = eps; a.row_coef = row_coef; a.scal = scal; a.use_log = use_log; a.use_lin = use_lin;
    return dmst::sf_grad_launch(a, nullptr) ? 0 : DMST_EINVAL;
}
#endif
}
