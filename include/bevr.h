/* bevr.h -- C ABI of libbevr.so, the sm_100a implementation of BEVRender's
 * deformable-attention hot path (rpl-cmu/bevrender; citations are file:line in
 * the reference checkout).
 *
 * Conventions
 *  - The reference is pure Python/PyTorch; what it "binds" for this path are
 *    torch library calls.  Each entry point below names the reference call
 *    sites it replaces.  The Python host side (bevrender_b200/) loads this
 *    library with ctypes and passes raw device pointers of torch tensors.
 *  - Ownership: the caller allocates every input, output and workspace buffer.
 *    The library never allocates, frees or retains device memory.
 *  - Every function only ENQUEUES work on `stream` (a cudaStream_t; NULL is the
 *    legacy default stream) on the caller's current device.  No implicit
 *    synchronisation; re-entrant.  The only process-wide state is a cache of
 *    the device's SM count (a constant) and per-kernel shared-memory opt-in
 *    attributes; kernel selection and tuning are per-call `flags`, never
 *    setters or environment variables.
 *  - Return value: 0 success; <0 argument validation error; >0 a cudaError_t
 *    raised at launch.  bevr_last_error() returns a thread-local message.
 *  - dtype: 0 = float32, 1 = bfloat16 (storage of value/out/weights/grads;
 *    coordinates and accumulators are always float32).
 *  - "host" pointers are read during the call (copied into kernel params).
 */
#ifndef BEVR_H_
#define BEVR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BEVR_VERSION 210

enum { BEVR_F32 = 0, BEVR_BF16 = 1 };

/* coordinate / behaviour flags */
enum {
  BEVR_ALIGN_CORNERS   = 1,  /* pixel = u*(size-1) instead of u*size-0.5 (F.grid_sample align_corners) */
  BEVR_LOC_PM1         = 2,  /* locations are grid_sample grids in [-1,1]; else in [0,1]               */
  BEVR_OFFSET_PIXELS   = 4,  /* with ref_points: loc = ref + off / (W_l, H_l); else loc = ref + off    */
  BEVR_ACCUMULATE      = 8,  /* bwd: add into grad_value instead of overwriting it (fp32: the REDs land on
                                the caller's contents; bf16: the final cast adds them in fp32)            */
  BEVR_BWD_DIRECT      = 16, /* msda_bwd: one RED per tap (round-1 kernel) instead of the tile-sorted
                                backward; the default picks the sorted kernel whenever Dh <= 128         */
  BEVR_GRAD_VALUE_F32  = 32  /* msda_bwd, dtype = bf16: grad_value is a FLOAT32 (B, S, h, Dh) buffer (fp32
                                main gradients of mixed-precision training): the reductions land in it
                                directly, no workspace, no zero-filled copy, no cast pass                */
};

#define BEVR_MAX_LEVELS 8

int bevr_version(void);
const char* bevr_last_error(void);

/* ---------------------------------------------------------------------------
 * K1  multi-scale deformable-attention gather, forward.
 * Replaces F.grid_sample at TSA_deform_attn.py:210-217 and
 * SCA_deform_attn.py:290-301 (L=1, P=1, attn_weight=NULL, BEVR_LOC_PM1 |
 * BEVR_ALIGN_CORNERS) and implements the general operator of the north star
 * (query/value/reference_points/spatial_shapes/level_start_index/
 * sampling_offsets/attention_weights).
 *   value        dev  dtype  (B, S, h, Dh)          channels-last, S = sum H_l*W_l
 *   spatial_hw   host int32  (L, 2)  rows (H_l, W_l)
 *   level_start  host int32  (L)
 *   sampling_loc dev  f32    (B, Nq, h, L, P, 2)    (x, y); offsets if ref_points != NULL
 *   ref_points   dev  f32    (B, Nq, L, 2) or NULL
 *   attn_weight  dev  dtype  (B, Nq, h, L, P) or NULL (all ones)
 *   out          dev  dtype  (B, Nq, h*Dh)
 *   query_w      width of the 2-D query grid (queries are tiled 2-D for cache
 *                locality) or 0 for a flat query list.
 * Requires Dh * sizeof(dtype) to be a multiple of 16 bytes.
 */
int bevr_msda_fwd(const void* value, const int32_t* spatial_hw, const int32_t* level_start,
                  const float* sampling_loc, const float* ref_points, const void* attn_weight,
                  void* out, int B, int S, int h, int Dh, int Nq, int L, int P, int query_w,
                  int dtype, int flags, void* stream);

/* K2  backward of K1 (replaces the autograd of the grid_sample calls above).
 *   grad_out     dev  dtype  (B, Nq, h*Dh)
 *   grad_value   dev  dtype  (B, S, h, Dh) or NULL (skip: TSA trains with a
 *                            detached prev_bev, bevrender.py:125-133)
 *   grad_loc     dev  f32    (B, Nq, h, L, P, 2) or NULL   (d/d sampling_loc)
 *   grad_weight  dev  dtype  (B, Nq, h, L, P) or NULL
 *   workspace    dev  f32    bevr_msda_bwd_workspace_bytes() bytes; only used
 *                            for dtype=bf16 with grad_value (fp32 accumulation)
 */
size_t bevr_msda_bwd_workspace_bytes(int B, int S, int h, int Dh, int dtype);
int bevr_msda_bwd(const void* value, const int32_t* spatial_hw, const int32_t* level_start,
                  const float* sampling_loc, const float* ref_points, const void* attn_weight,
                  const void* grad_out, void* grad_value, float* grad_loc, void* grad_weight,
                  void* workspace, int B, int S, int h, int Dh, int Nq, int L, int P, int query_w,
                  int dtype, int flags, void* stream);

/* ---------------------------------------------------------------------------
 * K4  camera -> BEV reference-point projection.
 * Replaces BEV2CameraProjector.bev_grid_to_camera + get_in_bound_mask
 * (bev_cmr_proj.py:61-124, remove_ref_in_gray=False), bit-exact with the
 * reference's torch-CPU result.
 *   points   dev f32 (4, n)  homogeneous IMU-frame points (SCA.py:112-162)
 *   cam_T    host f32 (16)   row-major inverse(imu_to_cam)  (inverse stays on the host)
 *   K        host f32 (9)    row-major intrinsics K[:, :3], already scaled to the image size
 *   uv       dev f32 (2, n)  normalised (x, y) in [-1,1]; masked points -> (-1,-1)
 *   pix      dev i32 (2, n)  or NULL: truncated pixel indices (x86 cvtt semantics)
 *   mask     dev u8  (n)     or NULL: 1 = inside [0,W-1) x [0,H-1)
 */
int bevr_project_points(const float* points, const float* cam_T, const float* K, float* uv,
                        int32_t* pix, uint8_t* mask, int n, int img_w, int img_h, void* stream);

/* ---------------------------------------------------------------------------
 * K3  fused attention with sampled relative-position bias (flash style; the
 * (Nq, Ns) matrix is never materialised).
 * Replaces einsum/mul/grid_sample/softmax/einsum at TSA_deform_attn.py:245-330
 * and SCA_deform_attn.py:331-411.
 *   q        dev f32 (BH, Dh, Nq)     BH = B*h, channel-major as the reference reshapes NCHW
 *   k, v     dev f32 (BH, Dh, Ns)
 *   rpe      dev f32 (h, Rh, Rw)      rpe_table; Rh = 2*Hq-1, Rw arbitrary (2*Wq-1 or 2*Wq*D-1)
 *   pos      dev f32 (BH/hpg, Ns, 2)  key positions (y, x) in grid_sample units
 *   out      dev f32 (BH, Dh, Nq)
 *   lse      dev f32 (BH, Nq)         log-sum-exp per row (saved for backward)
 *   hpg      heads per group (head bh uses pos[bh / hpg], table rpe[bh % h])
 *   workspace dev, bevr_rpe_attn_workspace_bytes(BH, h, Dh, Nq, Ns, Rh, Rw, backward) bytes: the zero-padded
 *            copy of the table the kernels sample, the forward's key-range partials (small grids are split
 *            over key ranges), the tf32 hi/lo copies of K, V (and Q, dO for backward) that the mma.sync
 *            kernels stage from, and, for backward, the privatised padded table gradients
 * The contractions run as mma.sync.m16n8k8 TF32 with a hi/lo operand split (3 products, fp32 accumulate:
 * fp32-grade results); grids below ~40x40 are served by an FFMA2 kernel family.  `flags`: 0 picks the family by
 * problem size, BEVR_K3_FFMA / BEVR_K3_MMA force one (tests and benchmarks run every case on both).
 * Requires Dh == 32 (the reference uses 32 channels per head everywhere, encoder.py:23-39) or Dh == 16.
 */
size_t bevr_rpe_attn_workspace_bytes(int BH, int h, int Dh, int Nq, int Ns, int Rh, int Rw, int backward);
enum { BEVR_K3_FFMA = 1, BEVR_K3_MMA = 2,
       BEVR_K3_RUN_MERGE = 4   /* backward (mma family): keys whose bias-table cells coincide arrive in runs (the caller
                                  ordered them by table row and x, as the SCA module does for masked reference points):
                                  the table-gradient scatter sums each run in the warp and issues one reduction per run */ };
/* (the workspace size covers both families: it does not depend on `flags`) */
int bevr_rpe_attn_fwd(const float* q, const float* k, const float* v, const float* rpe,
                      const float* pos, float* out, float* lse, void* workspace, int BH, int h,
                      int hpg, int Dh, int Hq, int Wq, int Ns, int Rh, int Rw, float scale,
                      int flags, void* stream);

/* Backward of K3.  grad_rpe (h,Rh,Rw) and grad_pos (BH/hpg,Ns,2) are ACCUMULATED into
 * (caller zero-fills); grad_q/grad_k/grad_v are overwritten.  Any grad pointer may be NULL.
 *   delta_ws dev f32 (BH, Nq) workspace (row-wise sum(dO*O)).
 */
int bevr_rpe_attn_bwd(const float* q, const float* k, const float* v, const float* rpe,
                      const float* pos, const float* out, const float* lse, const float* grad_out,
                      float* grad_q, float* grad_k, float* grad_v, float* grad_rpe, float* grad_pos,
                      float* delta_ws, void* workspace, int BH, int h, int hpg, int Dh, int Hq,
                      int Wq, int Ns, int Rh, int Rw, float scale, int flags, void* stream);

/* K3 with attention dropout (nn.Dropout on the softmax matrix, TSA_deform_attn.py:313-323, SCA_deform_attn.py:402-410):
 * every probability is kept with probability 1 - drop_p and scaled by 1 / (1 - drop_p) before it multiplies v; the
 * softmax denominator (and `lse`) are those of the undropped attention.  The keep mask is a counter-based hash of
 * (seed, b*h, query, key) -- never stored; the backward regenerates it from the same `drop_p` and `seed` the forward got.
 * Same distribution as torch's dropout, a different random stream (the reference draws from torch's Philox generator over
 * a matrix that does not exist here); oracle/deform_attn.py::dropout_keep restates the hash.  drop_p in [0, 1); 0 is the
 * plain entry point above.
 */
int bevr_rpe_attn_fwd_dropout(const float* q, const float* k, const float* v, const float* rpe,
                              const float* pos, float* out, float* lse, void* workspace, int BH, int h,
                              int hpg, int Dh, int Hq, int Wq, int Ns, int Rh, int Rw, float scale,
                              int flags, float drop_p, unsigned long long seed, void* stream);
int bevr_rpe_attn_bwd_dropout(const float* q, const float* k, const float* v, const float* rpe,
                              const float* pos, const float* out, const float* lse, const float* grad_out,
                              float* grad_q, float* grad_k, float* grad_v, float* grad_rpe, float* grad_pos,
                              float* delta_ws, void* workspace, int BH, int h, int hpg, int Dh, int Hq,
                              int Wq, int Ns, int Rh, int Rw, float scale, int flags, float drop_p,
                              unsigned long long seed, void* stream);

/* ---------------------------------------------------------------------------
 * K5  rotation-sweep correlation of a BEV template against an aerial-map tile (tcgen05 tensor cores).
 * The reference has NO registration / correlation code: this operator is defined by
 * oracle/correlate.py ("parity unpinned vs reference").
 *   out[r, y, x] = sum_{c,u,v} map[c, y+u, x+v] * rot_r(tmpl)[c, u, v]      valid mode, fp32 out
 *   map        dev  f32 (C, Hm, Wm)
 *   tmpl       dev  f32 (C, Ht, Wt)     Wt % 16 == 0
 *   angles_rad host f32 (R)             rot_r samples tmpl bilinearly (zeros outside) at
 *                                       centre + R(angle_r) (v - cx, u - cy), centre = ((Wt-1)/2, (Ht-1)/2)
 *   out        dev  f32 (R, Hm-Ht+1, Wm-Wt+1)
 *   workspace  dev  bevr_correlate_workspace_bytes() bytes (bf16 operands: 8 shifted map copies + the
 *                   rotated templates in UMMA core-matrix order)
 * Operands are rounded to bf16, products accumulate in fp32 (TMEM).  R <= 128.
 */
size_t bevr_correlate_workspace_bytes(int C, int Hm, int Wm, int Ht, int Wt, int R);
int bevr_correlate(const float* map, const float* tmpl, const float* angles_rad, float* out,
                   void* workspace, int C, int Hm, int Wm, int Ht, int Wt, int R, void* stream);

/* ---------------------------------------------------------------------------
 * History-BEV ego-motion warp ("next" row f-1): one batched pass of torchvision's tensor F.affine
 * (bilinear, fill = 0) as called twice per sample by EncoderLayer.project_history_bev_feat
 * (encoder.py:436-453).
 *   in, out  dev f32 (B, C, H, W), out != in
 *   theta    dev f32 (B, 2, 3)   inverse affine matrices in centred pixel units
 *                                (torchvision _get_inverse_affine_matrix with center = 0)
 */
int bevr_affine_warp(const float* in, const float* theta, float* out, int B, int C, int H, int W,
                     void* stream);

/* ---------------------------------------------------------------------------
 * Offset head ("next" row f-2): LayerNorm over channels -> GELU -> 1x1 conv (no bias) -> tanh * range * factor
 * (or clamp) -> + reference, i.e. conv_offset[1:] and the position arithmetic of
 * TSA_deform_attn.py:54-68,158-196 / SCA_deform_attn.py:56-77,219-277 (LayerNormProxy: model_utils.py:51-59),
 * fused into one kernel per direction.  Input is the output of the offset network's depthwise convolution.
 *   y          dev f32 (Bg, Cm, H, W)       NCHW
 *   ln_weight, ln_bias dev f32 (Cm); w1 dev f32 (O, Cm)
 *   ref        dev f32 (Bg or 1, Hk, Wk, 2) in (y, x); ref_batch_stride = Hk*Wk*2, or 0 to broadcast one grid
 *   pos        dev f32 (Bg, Hk, Wk, 2)      sampling positions in (y, x)
 *   stats      dev f32 (Bg*H*W, 2)          per-pixel mean / rstd, written by fwd, read by bwd
 *   interleave 0: O == 2, (Hk, Wk) = (H, W), pos[b, y, x, o]                                  (TSA)
 *              1: (Hk, Wk) = (H/2, W*O), pos[b, y>>1, x*O + o, y&1]  "(h n) w -> n h (w d)"   (SCA)
 *   scaled     1: tanh(raw) * (1/(Hk-1), 1/(Wk-1)) * factor + ref;  0: clamp(raw + ref, -1, 1)
 * Backward: grad_y is overwritten; grad_ln_weight / grad_ln_bias / grad_w1 are ACCUMULATED into (caller
 * zero-fills).  Any of the four may be NULL.  O <= 8.
 */
int bevr_offset_head_fwd(const float* y, const float* ln_weight, const float* ln_bias, const float* w1,
                         const float* ref, long long ref_batch_stride, float* pos, float* stats, int Bg,
                         int Cm, int H, int W, int O, int interleave, int scaled, float factor, float eps,
                         void* stream);
int bevr_offset_head_bwd(const float* y, const float* ln_weight, const float* ln_bias, const float* w1,
                         const float* ref, long long ref_batch_stride, const float* stats,
                         const float* grad_pos, float* grad_y, float* grad_ln_weight, float* grad_ln_bias,
                         float* grad_w1, int Bg, int Cm, int H, int W, int O, int interleave, int scaled,
                         float factor, float eps, void* stream);

/* ---------------------------------------------------------------------------
 * Retrieval-recall ranks ("next" row f-4): the counting loop of Trainer.get_recall (train.py:551-572).
 *   dist  dev f32 (n, n)  row-major distances, 2 - 2 X Y^T (the caller forms it with a library GEMM)
 *   rank  dev i32 (n)     OVERWRITTEN: rank[k] = #{ r : dist[r, k] < dist[k, k] }   (strict, as the reference)
 * recall@i = 100 * mean(rank < i).
 */
int bevr_recall_ranks(const float* dist, int32_t* rank, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BEVR_H_ */
