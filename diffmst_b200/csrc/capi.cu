// extern "C" entry points of libdiffmst_b200.so (include/diffmst_b200.h).
#include "console_host.cuh"
#include "afl.cuh"
#include "mrstft.cuh"
#include "peaknorm.cuh"

extern "C" {

int dmst_version(void) { return 1; }
int dmst_is_device_build(void) { return DMST_DEVICE_BUILD; }

size_t dmst_console_workspace_bytes(int B, int N, int T, unsigned flags) {
    (void)flags;
    if (B <= 0 || N <= 0 || T <= 0) return 0;
    // sized for the largest look-ahead the kernels accept (one tile)
    return dmst::carve_console(nullptr, B, N, T, dmst::kTrackTile, dmst::kMasterTile).total;
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
    return dmst::mrstft_run(x, x_row_stride, y, y_row_stride, windows, cfg, rows, T, loss, grad_x, workspace,
                            workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
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

}  // extern "C"
