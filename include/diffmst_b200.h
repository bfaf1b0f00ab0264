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
#define DMST_WANT_GRAD_TRACKS 256u  /* backward also writes dL/dtracks */
#define DMST_BASIC_CONSOLE 512u     /* BasicMixConsole layout: 2 track params [gain_db, pan] */

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

/* Bytes of workspace needed by dmst_console_forward/backward for this shape. */
size_t dmst_console_workspace_bytes(int B, int N, int T, unsigned flags);

/*
 * Forward.  tracks (B,N,T) -> mix (B,2,T) [+ mixed_tracks (B,2,N,T) when
 * DMST_WANT_MIXED_TRACKS].  track_params (B,N,27) and master_params (B,26) are the
 * NORMALISED (0..1) controller outputs, contiguous float32 (for DMST_BASIC_CONSOLE:
 * track_params is (B,N,2), master_params may be NULL).
 * status (int32[4], device): status[0] receives 0, or 1 + the flat index of the first
 * out-of-range parameter (track index space first, then 1000 + master index), which the
 * host turns into the reference's ValueError (mst/modules.py:86-89).
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

#ifdef __cplusplus
}
#endif
#endif /* DIFFMST_B200_H */
