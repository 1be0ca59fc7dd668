/*
 * centerpose_b200 — C ABI of the B200-native (sm_100a) centerpose inference hot path.
 *
 * The reference (tensorboy/centerpose) has no FFI for this path: its seams are Python
 * callables.  Every entry point below names the reference interface it
 * replaces; INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions (mirroring the reference's only native seam, the `_ext` DCNv2 module,
 * lib/models/backbones/DCNv2/src/dcn_v2.h:9-39 / src/cuda/dcn_v2_cuda.cu:42-172):
 *   - all tensor arguments are DEVICE pointers, fp32 unless said otherwise, contiguous;
 *   - work is enqueued on the caller's `stream` (a cudaStream_t passed as void*), no
 *     internal synchronisation, no allocation: scratch comes from caller-owned workspaces
 *     whose size is returned by the matching *_workspace_bytes() query;
 *   - return value 0 = ok; non-zero = error, text via cpb200_last_error() (thread-local);
 *     the Python shim raises RuntimeError, matching AT_ERROR/AT_ASSERTM -> RuntimeError.
 *   - no torch types anywhere in the signatures.
 */
#ifndef CENTERPOSE_B200_H_
#define CENTERPOSE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPB200_OK 0
#define CPB200_ERR_ARG 1      /* bad argument (shape, null pointer, unsupported size) */
#define CPB200_ERR_CUDA 2     /* CUDA runtime / launch failure */
#define CPB200_ERR_STATE 3    /* library not usable (no sm_100 device, driver entry point missing) */

#define CPB200_DECODE_MAX_K 128
#define CPB200_DECODE_MAX_J 32

/* Version / diagnostics. */
int cpb200_version(void);
const char *cpb200_last_error(void);
/* Number of kernels this library has launched since load (used by bench.py "gpu_launches"). */
unsigned long long cpb200_launch_count(void);

/* ------------------------------------------------------------------------------------
 * Fused decode.  Replaces lib/models/decode.py:235-308 `multi_pose_decode` together with
 * its helpers `_nms` (:10-16), `_topk` (:99-115), `_topk_channel` (:87-96) and
 * lib/models/utils.py:11-25 `_gather_feat` / `_transpose_and_gather_feat`; with
 * apply_sigmoid != 0 it also absorbs the two `sigmoid_()` calls of
 * lib/detectors/multi_pose.py:35-37 (the heads hand over logits).
 *
 *   heat      (B,1,H,W)   centre heat-map            (required)
 *   wh        (B,2,H,W)   box size                   (required)
 *   kps       (B,2J,H,W)  keypoint offsets           (required)
 *   reg       (B,2,H,W)   centre sub-pixel offset    (NULL -> +0.5, decode.py:253-255)
 *   hm_hp     (B,J,H,W)   keypoint heat-maps         (required: decode.py:307 needs hm_score)
 *   hp_offset (B,2,H,W)   keypoint sub-pixel offset  (NULL -> +0.5, decode.py:278-280)
 *   out       (B,K,5+3J)  [x1,y1,x2,y2, score, J*(x,y), J*kp_score]
 *
 * 1 <= K <= min(H*W, CPB200_DECODE_MAX_K); 1 <= J <= CPB200_DECODE_MAX_J.
 * Tie order (implementation-defined in the reference) is fixed: value descending, then
 * flat index ascending; nearest-candidate ties -> first (best-scored) candidate.
 * `workspace` must hold cpb200_decode_workspace_bytes(B,J,K) bytes, be 16-byte aligned and
 * ZERO-FILLED ONCE before its first use (the kernel restores the zero state itself);
 * it must not be shared by decodes running concurrently on different streams.
 * ---------------------------------------------------------------------------------- */
size_t cpb200_decode_workspace_bytes(int B, int J, int K);

int cpb200_multi_pose_decode(const float *heat, const float *wh, const float *kps,
                             const float *reg, const float *hm_hp, const float *hp_offset,
                             float *out, int B, int H, int W, int J, int K,
                             int apply_sigmoid, void *workspace, size_t workspace_bytes,
                             void *stream);

/* In-place logistic on n floats — lib/detectors/multi_pose.py:35-37 `hm.sigmoid_()`. */
int cpb200_sigmoid_inplace(float *x, size_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CENTERPOSE_B200_H_ */
