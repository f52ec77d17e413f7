/*
 * pwr.h — C ABI of libpwr_b200.so: the B200 (sm_100a) hot path of
 * PixelwiseRegression (SFR target builder, differentiable decoder, decoder
 * backward fused with the stage loss).
 *
 * The reference (/root/reference, pure Python) has no native interface; each
 * entry point below names the reference lines it replaces.  A reference-side
 * binding is a ctypes stub.
 *
 * Conventions (all entry points)
 *   - return 0 on success; < 0 argument error (PWR_E_*); > 0 a cudaError_t
 *     reported by the launch (cudaGetLastError right after enqueueing).
 *   - every pointer is a DEVICE pointer on the current device, contiguous,
 *     16-byte aligned; the caller owns all buffers; the library never
 *     allocates, frees, synchronises or keeps state.
 *   - `stream` is a cudaStream_t (NULL = legacy default stream); kernels are
 *     only enqueued, never waited for.
 *   - maps are [B, J, 64, 64] float32 row-major ("map" = one 64x64 plane);
 *     label/mask planes are [B, 1, 64, 64]; images [B, 1, 128, 128].
 *   - re-entrant and thread-safe (no globals).
 */
#ifndef PWR_B200_H
#define PWR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PWR_VERSION 201          /* 0.2.0: bumped on every ABI change; the binding asserts it */

#define PWR_LABEL_SIZE 64
#define PWR_IMAGE_SIZE 128
#define PWR_MAP_ELEMS  4096
#define PWR_MAX_JOINTS 64

/* argument errors */
#define PWR_E_NULL      (-1)     /* required pointer is NULL            */
#define PWR_E_SHAPE     (-2)     /* B/J/frame size out of range         */
#define PWR_E_ALIGN     (-3)     /* pointer not 16-byte aligned         */
#define PWR_E_METHOD    (-4)     /* unknown heat-map normalisation      */

/* heat-map normalisation, model.py:81-90 */
#define PWR_METHOD_SOFTMAX 0     /* softmax(w * z)              :83-85  */
#define PWR_METHOD_SUM     1     /* (relu(z)+1e-14) / sum       :88-90  */
#define PWR_METHOD_GIVEN   2     /* z IS the normalised heat map: used when
                                    DepthRegression.forward is called on its
                                    own with caller-supplied heat maps :116 */

/* element type of the conv outputs z, D, of gD_up and of the gradients gz, gD
 * (`map_dtype`): float32, or the float16 / bfloat16 tensors autocast hands over
 * under --mixed_precision (train.py:170-172).  All arithmetic is float32, as
 * autocast runs softmax / sum / mul-with-a-float32-operand; every other map
 * (L, m, H, targets, gH_up) is float32. */
#define PWR_DTYPE_F32  0
#define PWR_DTYPE_F16  1
#define PWR_DTYPE_BF16 2

int pwr_version(void);
const char* pwr_error_string(int rc);

/* Dispatch overrides for A/B measurements and for tests that compare two kernels of
 * the same entry point (process-wide, atomic; 0 = default heuristics).  Nothing on
 * the launch path reads the environment.  Returns the previous value, or
 * PWR_E_METHOD for an unknown option. */
#define PWR_OPT_BWD_DIRECT  0   /* 1: one-CTA-per-item backward instead of the pipelined ones   */
#define PWR_OPT_FWD_DIRECT  1   /* 1: one-CTA-per-item forward even without the heat-map store  */
#define PWR_OPT_FWD_PIPE    2   /* 1: pipelined forward even with the heat-map store            */
#define PWR_OPT_BWD_NO_LEAN 3   /* 1: 1-CTA/SM pipelined backward where the lean one would run  */
#define PWR_OPT_SFR_STAGED  4   /* 1: SFR build stages the source rows of a band in shared memory
                                      (bulk-TMA row copies) instead of gathering its taps from HBM;
                                      measured slower, kept for A/B runs     */
#define PWR_OPT_FUSED_NO_LEAN 5 /* 1: two-CTA/SM one-pass kernel where the lean three-CTA/SM one would run */
#define PWR_OPT_COUNT       6
int pwr_set_option(int option, int value);

/* ------------------------------------------------------------------------ *
 * SFR target builder
 * ------------------------------------------------------------------------ */

/* Centre of mass fallback, datasets.py:208-211 (MSRA: load_from_text returns
 * com=None).  com[b] = (mean col, mean row, mean depth) over frame pixels > 0,
 * float64.  frames [B,Hf,Wf] float32.  A frame without positive pixels yields
 * NaNs (the reference raises; the builder then flags valid=0). */
int pwr_sfr_com(const float* frames, int Hf, int Wf, double* com /*[B,3]*/,
                int B, void* stream);

/* HAND17 bounding-box loader, datasets.py:974-996 (`process_mode='bb'`: test frames that come with
 * a box instead of joints): raw [B,Hf,Wf] uint16 sensor counts, boxes [B,4] f64 = (ustart, vstart,
 * du, dv) -> frames_out [B,Hf,Wf] float32: zero outside MM[int(vstart):int(vstart+dv),
 * int(ustart):int(ustart+du)], and zero where depth > mean + 100 with the reference's two-pass mean
 * (exact integer sums, float64 divisions: bit-identical to NumPy; the values are integers, so the
 * reference's float64 frame is represented exactly).  Follow with pwr_sfr_com and pwr_sfr_crop
 * (frame_f64 = 1), as process_single_data does (datasets.py:203-214). */
int pwr_sfr_bb_filter(const uint16_t* raw, int Hf, int Wf, const double* boxes,
                      float* frames_out, int B, void* stream);

/* Frame formats (raw sensor frames can be fed directly).  The
 * float32 value the reference would hold after plt.imread + its decode line is
 * reproduced bit for bit. */
#define PWR_FRAME_F32   0   /* decoded float32 depth in mm (what
                               process_single_data receives)                */
#define PWR_FRAME_GB16  1   /* NYU PNG: uint16 = G << 8 | B; depth =
                               ((G/255)*256 + B/255)*255, datasets.py:810   */
#define PWR_FRAME_U16   2   /* 16-bit grey PNG (ICVL :632, HAND17 :940):
                               (v/65535)*65535                              */

/* Compact ("sparse") form of one joint's heat map + depth map target: a heat map is the 7x7
 * sigma-1.5 Gaussian (BORDER_REFLECT_101) of four taps (utils.py:37-65), so 64 bytes describe
 * what the dense 2 x 16 KiB planes hold.  The loss kernels can evaluate both targets on the fly
 * from this (see `taps` of pwr_decoder_fwd / pwr_decoder_bwd_loss), which removes the write and
 * the re-read of 2J dense maps per sample. */
typedef struct {
    double tap[4];              /* a, b, c, d at (ty0,tx0) (ty0,tx1) (ty1,tx0) (ty1,tx1)      */
    double cd;                  /* centred joint depth uvd_z - com_z (x scale if augmented)   */
    double cd_norm;             /* cd / cube: Dmap = (cd_norm - label_img) * [heat>0] * mask  */
    int16_t tx0, tx1, ty0, ty1; /* wrapped heat-map indices of the taps                       */
    int32_t ok;                 /* 0: the joint was rejected (maps are all zero)              */
    int32_t pad;
} pwr_joint_taps;

/* Bytes of caller-allocated device scratch pwr_sfr_crop (J = 0) / pwr_sfr_build
 * need: per-sample crop geometry and per-joint splat taps prepared once per
 * sample, plus the counter of the per-sample reject gate.  Contents need no
 * initialisation and are dead after the call.  0 for invalid B / J. */
size_t pwr_sfr_workspace_bytes(int B, int J);

/* Test-only SFR (the 6-tuple of datasets.py:334-348): crop box :306-309,
 * center_crop utils.py:167-173, depth window + centring :312-315, int CoM
 * :317-319, bilinear resize to 128x128 :323, 2x2 mean to 64x64 + mask
 * :330-332, /cube normalisation :337-338.
 *   frames [B,Hf,Wf] in `frame_format`; com [B,3] f64 (u,v,z); cube [B] f64;
 *   prefilter_margin >= 0 applies the hand rectangle of load_from_text
 *   (datasets.py:841-853 NYU / :956-968 HAND17 margin 40, :666-678 ICVL margin
 *   30): pixels outside MM[top:buttom, left:right] read as 0, with
 *   prefilter_umax = 2*halfu, prefilter_vmax = 2*halfv; < 0 = off.  (The depth
 *   window of :855-857 is the same predicate as :312 and idempotent.)
 *   frame_f64 != 0: the reference held the frame as float64 (MSRA), so the
 *   image path is evaluated in float64 and rounded once at the end.
 * outputs: img [B,1,128,128], label_img [B,1,64,64], mask [B,1,64,64],
 *   box_size [B], cube_size [B], com_out [B,3] (int(u), int(v), z) all f32;
 *   valid [B] u8 = 0 where the reference raises (empty crop).
 *   workspace: >= pwr_sfr_workspace_bytes(B, 0) bytes, 16-byte aligned. */
int pwr_sfr_crop(const void* frames, int frame_format, int Hf, int Wf,
                 const double* com, const double* cube, double fx, double fy,
                 int frame_f64,
                 double prefilter_margin, double prefilter_umax,
                 double prefilter_vmax,
                 float* img, float* label_img, float* mask,
                 float* box_size, float* cube_size, float* com_out,
                 uint8_t* valid, void* workspace, size_t workspace_size,
                 const int* win_extent, int win_h, int win_w,
                 int B, void* stream);

/* Train-mode SFR (the 9-tuple of datasets.py:403): everything pwr_sfr_crop
 * does plus uvd normalisation :350-353,381-383, heat-map coordinates
 * :356-358, 4-tap splat utils.py:37-62, 7x7 sigma 1.5 Gaussian
 * (cv2.GaussianBlur, BORDER_REFLECT_101) utils.py:64-65, Dmap :369-380 and
 * the reject gate :385-390 / :362-365.
 *   uvd [B,J,3] f64 joint pixel coordinates + depth.
 *   aug [B,8] f64 or NULL: the draws of the augmented branch, datasets.py:216-299
 *   with train.py's default flags (rotation, scale, shift; no flip): (scale,
 *   shift_u, shift_v, cos(a*(pi/180)), sin(a*(pi/180)), cos(a/180*pi),
 *   sin(a/180*pi), unused) for rotation angle a in degrees (utils.py:72-80; the
 *   two spellings of the radian conversion are the reference's: OpenCV's
 *   getRotationMatrix2D and the joint rotation).  The image is rotated and
 *   scaled with cv2.warpAffine's fixed-point bilinear recipe; a sample whose
 *   augmented branch would raise falls back to the plain branch, as the
 *   reference's try/except does.
 * extra outputs: uvd_norm [B,J,3], heatmaps [B,J,64,64], dmap [B,J,64,64] f32
 *   (both NULL = do not render the dense maps), joint_taps [B,J] (NULL = not
 *   wanted; at least one of the two target forms must be requested);
 *   valid [B] u8 = 0 where the reference raises (empty crop, heat-map index
 *   out of range, NaN, sum(mask) < 10).
 *   workspace: >= pwr_sfr_workspace_bytes(B, J) bytes, 16-byte aligned. */
int pwr_sfr_build(const void* frames, int frame_format, int Hf, int Wf,
                  const double* com, const double* cube, const double* uvd,
                  const double* aug,
                  double fx, double fy, int frame_f64,
                  double prefilter_margin, double prefilter_umax,
                  double prefilter_vmax,
                  float* img, float* label_img, float* mask,
                  float* box_size, float* cube_size, float* com_out,
                  float* uvd_norm, float* heatmaps, float* dmap,
                  pwr_joint_taps* joint_taps,
                  uint8_t* valid, void* workspace, size_t workspace_size,
                  const int* win_extent, int win_h, int win_w,
                  int B, int J, void* stream);

/* Host -> device feed of the depth frames (replaces the `.to(device)` of the whole frame,
 * train.py:161-166 / test.py:96-100).  `frames` [B,Hf,Wf] in `frame_format` may live in PINNED host
 * memory (cudaHostAlloc / torch pin_memory: device-addressable under UVA) - the one exception to
 * "every pointer is a device pointer" - or in device memory.  For every sample the kernel copies
 * only the region the builder can read, rows x cols of the crop box (datasets.py:306-311,
 * utils.py:167-173) intersected with the frame and, with the prefilter, with the hand rectangle of
 * load_from_text, columns rounded outwards to 16 bytes, into windows[b] ([win_h, win_w] elements of
 * the same format) and records it in win_extent[b] = (row0, col0, rows, cols).  With `aug`
 * (the [B,8] block of pwr_sfr_build) the union with the shifted-centre box is fetched.
 * pwr_sfr_crop / pwr_sfr_build then take (windows, win_extent, win_h, win_w) in place of the frames
 * ("window mode": frames = windows, Hf / Wf still the true frame size) and produce bit-identical
 * results; win_extent == NULL there means whole frames.
 *   fetched_bytes: optional device counter, incremented by the bytes read from `frames`.
 *   status: optional device int, OR-ed with 1 if some region did not fit [win_h, win_w] (the caller
 *   sized the windows too small; that sample is then built from a truncated window).
 * Requires Wf * elem % 16 == 0 and win_w * elem % 16 == 0 (PWR_E_SHAPE otherwise). */
int pwr_sfr_fetch(const void* frames, int frame_format, int Hf, int Wf,
                  const double* com, const double* cube, const double* aug,
                  double fx, double fy,
                  double prefilter_margin, double prefilter_umax,
                  double prefilter_vmax,
                  void* windows, int win_h, int win_w, int* win_extent,
                  unsigned long long* fetched_bytes, int* status,
                  int B, void* stream);

/* ------------------------------------------------------------------------ *
 * Differentiable decoder
 * ------------------------------------------------------------------------ */

/* Forward: PlaneRegression.forward post-conv model.py:79-97 +
 * DepthRegression.forward post-conv :123-132 + cat :151.
 *   z, D [B,J,64,64] conv outputs (logits / depth maps); w [J] (softmax
 *   temperature, model.py:74; only read for PWR_METHOD_SOFTMAX);
 *   L, m [B,1,64,64] label image and mask.  D == NULL (then L, m are not
 *   read) evaluates the plane branch alone and returns d = 0.
 *   heat_gt, dmap_gt [B,J,64,64], uvd_gt [B,J,3]: only read when
 *   loss_partial != NULL (an inner stage whose loss VALUE is needed at forward
 *   time, train.py:197-199).  taps [B,J] != NULL replaces the dense heat_gt /
 *   dmap_gt (which may then be NULL): the targets are evaluated on the fly.
 * outputs: H [B,J,64,64] normalised heat maps (NULL = do not store, e.g. the
 *   last stage at inference); uvd [B,J,3]; stats [B,J,4] = (extremum of z
 *   in the direction of sign(w), 1/sum, masked-heat sum + 1e-14, d) saved for the backward
 *   (NULL = do not store); loss_partial [B,J,3] per-(b,j) sums of squares
 *   (heat, dmap, uvd) before lambda/mean scaling (NULL = no loss). */
int pwr_decoder_fwd(const void* z, const float* w, const void* D,
                    const float* L, const float* m,
                    const float* heat_gt, const float* dmap_gt,
                    const float* uvd_gt, const pwr_joint_taps* taps,
                    float* H, float* uvd, float* stats, float* loss_partial,
                    int B, int J, int method, int map_dtype, void* stream);

/* Backward of pwr_decoder_fwd (what autograd derives from model.py:79-132).
 *   g_uvd [B,J,3] upstream gradient on uvd (NULL = zeros);
 *   gH_up, gD_up [B,J,64,64] dense upstream gradients on the heat maps and on
 *   the depth maps (next stage's conv, model.py:208; loss terms computed
 *   outside); either may be NULL.
 * outputs: gz, gD [B,J,64,64] (either may be NULL = not wanted); gw_partial
 *   [B,J] (sum over pixels of dL/d(w z) * z; reduce over B with
 *   pwr_reduce_partials; NULL unless PWR_METHOD_SOFTMAX). */
int pwr_decoder_bwd(const void* z, const float* w, const void* D,
                    const float* L, const float* m, const float* stats,
                    const float* uvd, const float* g_uvd,
                    const float* gH_up, const void* gD_up,
                    void* gz, void* gD, float* gw_partial,
                    int B, int J, int method, int map_dtype, void* stream);

/* Backward fused with the stage loss of train.py:197-205:
 *   loss = alpha*mean(sum (uvd-uvd_gt)^2) + (1-alpha)*(lambda_h*mean(sum
 *   (H-heat_gt)^2) + lambda_d*mean(sum (D-dmap_gt)^2)), means over B*J.
 * Adds d(scale*loss)/d(.), scale = loss_scale * (loss_scale_dev ?
 * *loss_scale_dev : 1) (a device scalar: the upstream gradient on the loss,
 * e.g. a GradScaler factor, without a host sync), to the upstream gradients
 * of pwr_decoder_bwd
 * and emits loss_partial [B,J,3] = per-(b,j) sums of squares (heat, dmap,
 * uvd) BEFORE lambda/mean scaling (NULL = not wanted; the target maps are
 * then only read if their weight (1-alpha)*lambda is non-zero).  `n_mean` is
 * the B*J of the mean (the global batch under data parallelism; 0 = B*J). */
int pwr_decoder_bwd_loss(const void* z, const float* w, const void* D,
                         const float* L, const float* m, const float* stats,
                         const float* uvd, const float* g_uvd,
                         const float* gH_up, const void* gD_up,
                         const float* heat_gt, const float* dmap_gt,
                         const float* uvd_gt, const pwr_joint_taps* taps,
                         float alpha, float lambda_h, float lambda_d,
                         float loss_scale, const float* loss_scale_dev,
                         int n_mean,
                         void* gz, void* gD, float* gw_partial,
                         float* loss_partial,
                         int B, int J, int method, int map_dtype, void* stream);

/* Last stage in one pass: pwr_decoder_fwd + pwr_decoder_bwd_loss without upstream gradients
 * (train.py:192-207 when nothing but the loss consumes the stage's outputs).  z, D, heat_gt, dmap_gt
 * are read once; `stats` never exists.  Arguments as in the two calls it replaces:
 *   outputs H [B,J,64,64] (NULL = not wanted), uvd [B,J,3], gz, gD (either NULL = not wanted),
 *   gw_partial [B,J] (PWR_METHOD_SOFTMAX), loss_partial [B,J,3] (NULL = not wanted).
 * method is PWR_METHOD_SOFTMAX or PWR_METHOD_SUM; the depth branch is mandatory. */
int pwr_decoder_fwd_bwd_loss(const void* z, const float* w, const void* D,
                             const float* L, const float* m,
                             const float* heat_gt, const float* dmap_gt,
                             const float* uvd_gt, const pwr_joint_taps* taps,
                             float alpha, float lambda_h, float lambda_d,
                             float loss_scale, const float* loss_scale_dev,
                             int n_mean,
                             float* H, float* uvd, void* gz, void* gD,
                             float* gw_partial, float* loss_partial,
                             int B, int J, int method, int map_dtype, void* stream);

/* Deterministic sum over the batch axis: out[j*C + c] = sum_b in[(b*J+j)*C+c]
 * (gw_partial -> gw: C=1; loss_partial -> per-joint sums: C=3). */
int pwr_reduce_partials(const float* in, float* out, int B, int J, int C,
                        void* stream);

/* Stage loss values of train.py:197-205 from loss_partial [B,J,3]:
 *   out4 = (lambda_h*mean_h, lambda_d*mean_d, mean_u, alpha*mean_u +
 *   (1-alpha)*(lambda_h*mean_h + lambda_d*mean_d)), means over n_mean (0 =
 *   B*J) of the per-(b,j) sums of squares.  One launch, fixed summation order.
 *   gw_partial [B,J] / gw_out [J] (both or neither): the same launch also sums
 *   dL/dw over the batch (what pwr_reduce_partials with C = 1 would do). */
int pwr_stage_loss(const float* loss_partial, int B, int J,
                   float lambda_h, float lambda_d, float alpha, int n_mean,
                   float* out4, const float* gw_partial, float* gw_out,
                   void* stream);

/* In-place multiply by *scale_dev (device scalar; an early exit when it is 1): n elements
 * (n % 4 == 0, element type map_dtype) of x and, if not NULL, of x2, plus n_small <= 256 float32 of
 * `small`.  Applies a non-unit upstream gradient (GradScaler) to gz, gD and dL/dw that were produced
 * eagerly by the one-pass last stage - one launch for all three. */
int pwr_scale_inplace(void* x, void* x2, float* small, int n_small,
                      const float* scale_dev, long long n,
                      int map_dtype, void* stream);

/* utils.py:332-337 recover_uvd fused with uvd2xyz datasets.py:100-111:
 *   uvd_px = uvd_norm * (box-1 | box-1 | cube) + com ; xyz = ((u-halfu)/fx*d,
 *   (v-halfv)/fy*d, d).  uvd_px / xyz may be NULL. */
int pwr_recover_uvd(const float* uvd_norm, const float* box_size,
                    const float* cube_size, const float* com,
                    double fx, double fy, double halfu, double halfv,
                    float* uvd_px, float* xyz, int B, int J, void* stream);

/* Evaluation metric of train.py:254-276 / test.py:106-113 without the .cpu()
 * round trip: err[b] = mean_j | xyz(recover_uvd(uvd_pred)) -
 * xyz(recover_uvd(uvd_true)) |_2 in mm, float32 in the reference's order. */
int pwr_joint_error(const float* uvd_pred, const float* uvd_true,
                    const float* box_size, const float* cube_size,
                    const float* com, double fx, double fy, double halfu,
                    double halfv, float* err, int B, int J, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PWR_B200_H */
