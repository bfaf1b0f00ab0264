// extern "C" entry points of libdiffmst_b200.so (include/diffmst_b200.h).
#include "console_host.cuh"

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

}  // extern "C"
