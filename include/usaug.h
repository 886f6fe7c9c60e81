/*
 * usaug.h -- C ABI of libusaug.so: the B200 (sm_100a) hot path of the physics-inspired
 * ultrasound augmentation chain (probe-pressure deformation + TGC + remap, random-walk
 * confidence map, SNR change + reverberation).
 *
 * Reference interface replaced: NONE IS VISIBLE.  The mounted reference is
 * /root/reference/README.md:1-2 (title + "Realistic and Physics based Ultrasound
 * Augmentation"); it ships no code, so there is no reference FFI to cite line by line.
 * The boundary below is the one SURVEY.md section 8(b) defines for this path (API UNPINNED by
 * the reference); each entry point names the SURVEY.md Appendix-A clause it implements.
 *
 * Contract for every function:
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - returns 0 (USAUG_OK) or a negative USAUG_E_* code; never throws.
 *   - never allocates, frees or synchronises; all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = the legacy default stream).
 *   - every pointer is a DEVICE pointer owned by the caller and must stay valid until the
 *     stream has executed the work; scratch memory comes from the caller (`ws`, sized by the
 *     matching *_workspace_bytes query, 256-byte aligned).
 *   - re-entrant; one call may be in flight per stream.
 *   - layouts: image float32 [B,H,W] contiguous (NCHW with C=1), label uint8 [B,H,W],
 *     bone = non-zero.  Rows = depth (probe at row 0), columns = scanlines.
 *   - there is NO CPU fallback: without a CUDA device the calls return USAUG_E_CUDA.
 */
#ifndef USAUG_H_
#define USAUG_H_

#include <stddef.h>
#include <stdint.h>

/* libusaug.so is built with -fvisibility=hidden: only the entry points below are exported. */
#if defined(__GNUC__)
#define USAUG_API __attribute__((visibility("default")))
#else
#define USAUG_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define USAUG_VERSION 10200 /* 1.2.0 */

enum {
  USAUG_OK = 0,
  USAUG_E_INVALID_ARG = -1,
  USAUG_E_CUDA = -2,
  USAUG_E_WORKSPACE_TOO_SMALL = -3,
  USAUG_E_BOUNDS = -4 /* bounds-checked debug build only (lib/libusaug_dbg.so): a kernel wrote past one of its workspace planes */
};

/* preconditioner of the confidence-map PCG */
enum { USAUG_PRECOND_JACOBI = 0, USAUG_PRECOND_LINE = 1 };

/* flags of usaug_confmap_pcg */
enum {
  USAUG_CM_FP64_VECTORS = 1, /* keep the CG vectors in fp64 (default: fp32 vectors + fp64 refinement) */
  USAUG_CM_ZERO_GUESS = 2,   /* start from x0 = 0 instead of the linear depth ramp */
  USAUG_CM_DEBUG_TIMING = 4, /* development: the kernel prints per-slot wall times of a few steps */
  USAUG_CM_LEGACY_KERNEL = 8, /* A/B testing: use the two-slot kernel even where the fused fp32 kernel applies */
  USAUG_CM_NO_COARSE = 32,    /* A/B testing: line preconditioner only (no aggregation coarse-grid correction) */
  /* aggregate height of the coarse grid; default: 2-row aggregates for deeply bandwidth-bound batches
     (>= 12 strip tasks per SM), 4 rows (or more, for tall images) otherwise */
  USAUG_CM_COARSE_FINE = 64,      /* always start from 2-row aggregates */
  USAUG_CM_COARSE_STANDARD = 128, /* always start from 4-row aggregates */
  /* Twisted column-line factorisation (two half sweeps per strip; halves the dependent chain of one image):
     default on unless the launch is deeply bandwidth-bound (>= 12 strip tasks per SM) */
  USAUG_CM_TWIST_ON = 256,
  USAUG_CM_TWIST_OFF = 512,
  /* F_UP phase of the fused kernel: recompute the edge weights from the attenuation plane (default for >= 6 strip
     tasks per SM) or read the weight planes.  Pinning the variant makes results independent of the batch size. */
  USAUG_CM_FORCE_RECOMPUTE = 1024,
  USAUG_CM_FORCE_PLANES = 2048,
  USAUG_CM_NO_LEVEL = 4096, /* A/B testing: without the level preconditioner (line solve on 1 x 4 aggregates) */
  USAUG_CM_SINKS = 8192     /* workspace query: room for the sink plane of usaug_confmap_pcg_sinks (sink_mode != bottom row) */
};

/* where the random walks of the confidence map are absorbed (potential 0); the top row is always the source */
enum {
  USAUG_SINK_BOTTOM_ROW = 0,       /* A.3: the bottom row */
  USAUG_SINK_BOTTOM_AND_SIDES = 1, /* ... and the first and the last column (walks may not leave the sector sideways) */
  USAUG_SINK_MASK = 2              /* ... and every pixel of rows 1..H-2 where the caller's mask is non-zero */
};

USAUG_API int usaug_version(void);
/* thread-local message of the last failing call on this thread ("" if none) */
USAUG_API const char* usaug_last_error(void);

/* One stream-ordered host-to-device copy (cudaMemcpyAsync) of `bytes` bytes: the per-call parameter block.
 * `host` should be page-locked memory, so that the copy is asynchronous and can be captured as a memcpy node of a
 * CUDA graph (the node re-reads the host block on every replay). */
USAUG_API int usaug_upload(const void* host, void* dev, size_t bytes, void* stream);

/* ---- K1: bone-surface scan + fused TGC + deformation + remap ---------
 * d[B]      axial probe displacement in pixels (clamped on device to |d| <= 0.5*min s_hat)
 * tgc[B,K]  K >= 1 control gains at depths k(H-1)/(K-1)
 * taps[2R+1] smoothing taps of the bone-surface profile (host-computed, identical to the oracle's)
 * out_img/out_lab must not alias the inputs.  Requires 1 <= W <= 4096.
 */
USAUG_API size_t usaug_warp_workspace_bytes(int B, int H, int W);
USAUG_API int usaug_fused_tgc_deform(const float* img, const uint8_t* lab, float* out_img, uint8_t* out_lab,
                           int B, int H, int W, const float* d, const float* tgc, int K,
                           const float* taps, int R, void* ws, size_t ws_bytes, void* stream);
/* 8-bit ingest: img holds 8-bit B-mode pixels, read as v / 255 (correctly rounded
 * fp32); everything else as above.  2 B/px less HBM and 3 B/px less host-to-device traffic. */
USAUG_API int usaug_fused_tgc_deform_u8(const uint8_t* img, const uint8_t* lab, float* out_img, uint8_t* out_lab,
                              int B, int H, int W, const float* d, const float* tgc, int K,
                              const float* taps, int R, void* ws, size_t ws_bytes, void* stream);

/* ---- K2: confidence map, matrix-free PCG on the 8-neighbour random-walk Laplacian (A.3) ----
 * cm[B,H,W] out (row 0 = 1, row H-1 = 0).  Stops per image when the TRUE residual satisfies
 * ||b - A x||_2 <= rtol*||b||_2 or after maxiter inner iterations.
 * iters[B] (nullable) inner iterations used; relres[B] (nullable) true relative residual reached.
 * An image the scheduler watchdog gave up on reports iters = -1, relres = NaN and keeps cm = 0.
 * img and cm must be 16-byte aligned.  Requires H >= 3.
 */
USAUG_API size_t usaug_confmap_workspace_bytes(int B, int H, int W, int precond, int flags);
USAUG_API int usaug_confmap_pcg(const float* img, float* cm, int B, int H, int W, float alpha, float beta,
                      float gamma, double rtol, int maxiter, int precond, int flags, int* iters,
                      double* relres, void* ws, size_t ws_bytes, void* stream);

/* The same solve with further sink pixels.  sink_mask: uint8 [B,H,W], read only for
 * USAUG_SINK_MASK (NULL otherwise).  A sink pixel is held at potential 0: cm = 0 there.  For sink_mode != bottom row
 * the workspace must be sized with USAUG_CM_SINKS set in the flags of the workspace query. */
USAUG_API int usaug_confmap_pcg_sinks(const float* img, float* cm, int B, int H, int W, float alpha, float beta,
                      float gamma, double rtol, int maxiter, int precond, int flags, int sink_mode,
                      const uint8_t* sink_mask, int* iters, double* relres, void* ws, size_t ws_bytes, void* stream);

/* Measurement hook (not part of the data path): when set (non-NULL cudaEvent_t handles), the
 * next usaug_confmap_pcg calls made by THIS thread record `start`/`stop` on the launching stream
 * immediately around the persistent PCG kernel.  Pass NULL, NULL to clear. */
USAUG_API void usaug_confmap_set_timing_events(void* start_event, void* stop_event);

/* ---- K3: SNR change + reverberation, fused (A.4, A.5) --------------------------------------
 * cm nullable (NULL = no SNR change).  a_snr[B], rho[B], nghost[B] per-sample parameters;
 * feather_taps[2R+1] Gaussian taps of the bone-mask feathering.  out must not alias img.
 */
USAUG_API size_t usaug_snr_reverb_workspace_bytes(int B, int H, int W);
USAUG_API int usaug_snr_reverb(const float* img, const float* cm, const uint8_t* lab, float* out, int B,
                     int H, int W, const float* a_snr, const float* rho, const int* nghost,
                     const float* feather_taps, int R, void* ws, size_t ws_bytes, void* stream);
/* 8-bit egress: out = floor(255 * v + 0.5) of the fp32 result v (two individually rounded fp32 operations). */
USAUG_API int usaug_snr_reverb_u8(const float* img, const float* cm, const uint8_t* lab, uint8_t* out, int B,
                        int H, int W, const float* a_snr, const float* rho, const int* nghost,
                        const float* feather_taps, int R, void* ws, size_t ws_bytes, void* stream);

/* General form of A.4: I1 = clip(I + (a-1) * C * S) with the signal map
 * S = sig * sig_scale[b] instead of S = I (sig, sig_scale nullable together: NULL = the form above).
 * out is float* or, with out_is_u8 != 0, uint8_t*. */
USAUG_API int usaug_snr_reverb_signal(const float* img, const float* cm, const float* sig, const float* sig_scale,
                            const uint8_t* lab, void* out, int out_is_u8, int B, int H, int W,
                            const float* a_snr, const float* rho, const int* nghost,
                            const float* feather_taps, int R, void* ws, size_t ws_bytes, void* stream);

/* ---- monogenic-signal local energy ------------------------------------
 * energy[B,H,W] = sqrt(even^2 + odd1^2 + odd2^2) of the image filtered with the sum of `scales`
 * log-Gabor band-pass filters (wavelengths lambda_min * mult^s pixels, bandwidth sigma_f) and their
 * Riesz transforms; scale[B] = 1 / max(energy) per image (0 if the image has no band-pass content),
 * so energy * scale is the normalised map in [0,1] that usaug_snr_reverb_signal takes.
 * Hand-written batched 2-D FFT, 4 <= H, W <= 2048: sizes that are not powers of two are mirror-extended
 * (bottom / right) to the next power of two, and the energy is cropped back.  1 <= scales <= 4. */
USAUG_API size_t usaug_local_energy_workspace_bytes(int B, int H, int W);
USAUG_API int usaug_local_energy(const float* img, float* energy, float* scale, int B, int H, int W,
                       float lambda_min, float mult, int scales, float sigma_f, void* ws,
                       size_t ws_bytes, void* stream);

/* ---- curvilinear (convex-probe) geometry: scan conversion ----------------
 * direction 0: Cartesian img/lab [B,H,W] -> polar out [B,n_samples,n_rays], sampled at
 *              (x, y) = (apex_x, apex_y) + r_i * (ray_sin[j], ray_cos[j]),  r_i = r_min + i * dr;
 * direction 1: polar img/lab [B,n_samples,n_rays] -> Cartesian out [B,H,W], sampled at
 *              i* = (r - r_min) * inv_dr,  j* = (theta + theta_max) * inv_dtheta  (0 outside the sector).
 * Image 4-tap bilinear (zeros outside), label nearest.  ray_sin / ray_cos [n_rays] are device tables
 * (needed for direction 0 only).  All coordinates are individually rounded fp32 operations. */
USAUG_API int usaug_scan_convert(const float* img, const uint8_t* lab, float* out_img, uint8_t* out_lab, int B,
                       int H, int W, int n_samples, int n_rays, int direction, float apex_x, float apex_y,
                       float theta_max, float r_min, float dr, float inv_dr, float inv_dtheta,
                       const float* ray_sin, const float* ray_cos, void* stream);

/* ---- bone bounding box of a label batch (A.1 step 2 / A.5 "bone-surface row") --------------
 * box[B,4] = {y_top, y_bottom, x_left, x_right}; y_top = -1 when the image has no bone.
 */
USAUG_API size_t usaug_bone_box_workspace_bytes(int B, int H, int W);
USAUG_API int usaug_bone_box(const uint8_t* lab, int B, int H, int W, int* box, void* ws, size_t ws_bytes,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* USAUG_H_ */
