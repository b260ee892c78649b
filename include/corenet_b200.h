/* corenet_b200 — C-ABI of the B200 (sm_100a) CoReNet hot path.
 *
 * Every entry point is `extern "C"`, takes plain device pointers + sizes + a
 * `cudaStream_t` (passed as void*), launches asynchronously on that stream and
 * returns 0 on success or a negative `crn_status`.  No entry point allocates
 * device memory, synchronises the device, or owns any of its arguments.
 *
 * Activation layout inside the path is channels-last ("rows x channels"):
 *   2-D maps  [N, H, W, Cs]      3-D grids [N, D, H, W, Cs]
 * with an explicit channel stride Cs (floats) and channel offset so that a
 * producer can write straight into a slice of a wider (concat) buffer.
 * Weights are consumed in a packed, tap-major layout produced by
 * crn_pack_weights (see there).
 *
 * Each group cites the reference interface it replaces (paths relative to
 * /root/reference/src/corenet).
 */
#ifndef CORENET_B200_H_
#define CORENET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  CRN_OK = 0,
  CRN_ERR_BAD_ARG = -1,     /* shape / alignment / null pointer */
  CRN_ERR_UNSUPPORTED = -2, /* valid request this build has no kernel for */
  CRN_ERR_LAUNCH = -3       /* cudaGetLastError() after the launch was not cudaSuccess */
} crn_status;

/* Version + build info (arch string the kernels were compiled for). */
int crn_version(void);
const char* crn_build_arch(void);
/* Text of the last error on this thread ("" if none). */
const char* crn_last_error(void);
/* Number of CUDA kernels this library has launched in this process (bench.py gpu_launches). */
int64_t crn_launch_count(void);
/* Debug / A-B switches: bit0 = disable the row-direct conv kernels, bit1 = disable the tap-row wgrad
 * kernel (both fall back to the generic implicit-GEMM kernels), bit4 = disable split-K in the generic and the
 * tcgen05 implicit-GEMM kernels, bit5 = disable the stem wgrad kernel, bit6 = 3 narrow MMAs instead of N doubling in
 * conv_tc5, bits 8-11 = gemm_tc_kernel debug (timeline stamps / skip stores / skip gathers / skip MMAs),
 * bit12 = crn_fill_inside uses the global-memory line-sweep kernels instead of the shared-memory cluster kernel,
 * bit13 = single-pass TF32: every tcgen05 kernel issues only the hi x hi product of its 3xTF32 operand split (an
 *         opt-in precision mode, engine.set_precision("tf32"); the default three-term split is fp32-class). */
void crn_set_flags(int32_t flags);

/* ------------------------------------------------------------------------
 * Convolutions.  Replaces the ATen/cuDNN calls behind nn.Conv2d / nn.Conv3d /
 * nn.ConvTranspose3d / nn.Linear at model/resnet50.py:61-70,94-108,122-131,
 * model/reconstruction_decoder.py:47-95 and
 * model/ray_traced_skip_connection.py:38,85.
 *
 * One descriptor covers 2-D (D = 1, kD = 1) and 3-D, plain and transposed.
 * ---------------------------------------------------------------------- */
typedef struct {
  int32_t N;               /* batch */
  int32_t Cin, Cout;       /* logical channels of the conv (x has Cin, y has Cout) */
  int32_t iD, iH, iW;      /* spatial dims of x */
  int32_t oD, oH, oW;      /* spatial dims of y */
  int32_t kD, kH, kW;      /* kernel */
  int32_t stride;          /* same on every spatial axis with extent > 1 */
  int32_t pad;             /* same on every axis with k > 1 */
  int32_t transposed;      /* 0: y = conv(x)   1: y = conv_transpose(x) */
  int32_t x_cs, x_co;      /* channel stride / offset of x rows (floats) */
  int32_t y_cs, y_co;      /* channel stride / offset of y rows (floats) */
  int32_t CinP, CoutP;     /* padded channel counts of the packed weights (multiples of 4) */
  int32_t y_planar;        /* 1: y is NC(D)HW planar (only for fwd; used for the logits) */
  int32_t bias_n_stride;   /* 0: bias[Cout]; else bias[n*bias_n_stride + co] (per-scene bias) */
} crn_conv_desc;

/* Packed weight layouts (floats):
 *   fwd  : Wf[tap][CinP][CoutP]   tap = (kz*kH + ky)*kW + kx
 *   dgrad: Wd[tap][CoutP][CinP]
 * Source is the PyTorch parameter: conv [Cout][Cin][taps], transposed conv
 * [Cin][Cout][taps], linear [Cout][Cin] (taps = 1). Padding entries are zero. */
typedef struct {
  const float* src;        /* parameter tensor */
  float* dst_fwd;          /* may be NULL */
  float* dst_dgrad;        /* may be NULL */
  int32_t Cin, Cout, taps, CinP, CoutP;
  int32_t src_is_transposed; /* 1: src is [Cin][Cout][taps] */
} crn_pack_item;
/* `items` is a DEVICE array of n items; total = sum(taps*CinP*CoutP) elements and
 * `offsets` a DEVICE int64 array [n+1] of exclusive prefix sums of those sizes. */
int crn_pack_weights(const crn_pack_item* items, const int64_t* offsets, int32_t n,
                     int64_t total, void* stream);
/* Inverse for gradients: grad_packed is Wf layout [tap][CinP][CoutP] (what
 * crn_conv_wgrad writes); dst is the PyTorch layout of the parameter's .grad.
 * Here offsets/total count DESTINATION elements (taps*Cin*Cout per item). */
typedef struct {
  const float* src_packed;
  float* dst;
  int32_t Cin, Cout, taps, CinP, CoutP;
  int32_t dst_is_transposed;
} crn_unpack_item;
int crn_unpack_wgrads(const crn_unpack_item* items, const int64_t* offsets, int32_t n,
                      int64_t total, void* stream);

/* y = conv(x, Wf) + bias.           accumulate != 0: y += ... (bias ignored). */
int crn_conv_fwd(const crn_conv_desc* d, const float* x, const float* w_fwd, const float* bias,
                 float* y, int32_t accumulate, void* stream);
/* dx = conv^T(dy, Wd) (the gradient wrt x of crn_conv_fwd with the same desc). */
int crn_conv_dgrad(const crn_conv_desc* d, const float* dy, const float* w_dgrad, float* dx,
                   int32_t accumulate, void* stream);
/* dWf[tap][ci][co] += sum_rows x * dy.  dw must be zeroed by the caller
 * before the first call (split-K partial sums are added atomically). */
int crn_conv_wgrad(const crn_conv_desc* d, const float* x, const float* dy, float* dw_packed,
                   void* stream);

/* ------------------------------------------------------------------------
 * Batch renormalisation.  Replaces model/batch_renorm.py:33-62.
 * x is rows x C channels-last with channel stride x_cs.  `relu_in` applies
 * ReLU to x before the statistics / normalisation (decoder order
 * ReLU -> BRN -> conv, reconstruction_decoder.py:51-60).
 * ---------------------------------------------------------------------- */
/* acc: double[3*C]; [0,2C) zeroed by caller, receives sum(x-k), sum((x-k)^2); [2C,3C) receives the shift k
 * (the channel's value in row 0). */
int crn_brn_stats(const float* x, int64_t rows, int32_t C, int32_t x_cs, int32_t x_co,
                  int32_t relu_in, double* acc, void* stream);
/* Turns the sums into per-channel coefficients and updates the running
 * statistics exactly like the reference (including the channel-count Bessel
 * quirk and num_batches_tracked += 1).  coef: float[6*C] =
 *   a (scale), b (shift), mean, invstd, r, d with y = a*(x - mean) + b.   training==0: running stats only. */
int crn_brn_finalize(const double* acc, int64_t rows, int32_t C, const float* weight,
                     const float* bias, float* running_mean, float* running_var,
                     int64_t* num_batches_tracked, float eps, float momentum, int32_t training,
                     float* coef, void* stream);
/* y = act( a*(f(x) - mean) + b [+ res] ),  f = relu if relu_in.  If y_pre != NULL the
 * pre-activation value is stored there too (encoder skip taps,
 * resnet50.py:76-80).  relu_out: apply ReLU to y. */
int crn_brn_apply(const float* x, int64_t rows, int32_t C, int32_t x_cs, int32_t x_co,
                  const float* coef, const float* res, int32_t relu_in, int32_t relu_out,
                  float* y, int32_t y_cs, int32_t y_co, float* y_pre, void* stream);
/* Backward, pass 1: g = dy * (relu_out ? y>0 : 1) [+ g_extra]; accumulates
 * sum(g) and sum(g*xhat) per channel into acc (double[2*C], zeroed by caller).
 * If g_out != NULL the masked/combined g is stored (it is the residual
 * branch's gradient). */
int crn_brn_bwd_reduce(const float* dy, int32_t dy_cs, int32_t dy_co, const float* y_act,
                       const float* g_extra, const float* x, int32_t x_cs, int32_t x_co,
                       int64_t rows, int32_t C, const float* coef, int32_t relu_in,
                       int32_t relu_out, float* g_out, double* acc, void* stream);
/* Backward, pass 2: dx = (relu_in ? x>0 : 1) * a' * (g - S1/R - xhat*S2/R)  (training)
 *                   dx = (relu_in ? x>0 : 1) * a * g                       (eval)
 * Also writes dweight, dbias (float[C]) and accumulates column sums of dx into
 * dxsum (double[C], zeroed by caller; it is the preceding conv's bias gradient)
 * when dxsum != NULL.  g is the masked gradient (g_out of pass 1, or dy when no mask). */
int crn_brn_bwd_dx(const float* g, int32_t g_cs, int32_t g_co, const float* x, int32_t x_cs,
                   int32_t x_co, int64_t rows, int32_t C, const float* coef, const double* acc,
                   const float* weight, int32_t relu_in, int32_t training, float* dx,
                   int32_t dx_cs, int32_t dx_co, int32_t dx_accumulate, float* dweight,
                   float* dbias, double* dxsum, void* stream);

/* Single-launch forms for SMALL activations (rows * C <= 5 Mi elements, C % 8 == 0: the 53 encoder instances and
 * the coarse decoder stages at the training batch size): a thread-block cluster owns 8 channels, reduces in a fixed
 * order through distributed shared memory (no atomics: bit-reproducible statistics), derives the coefficients and
 * applies them in the same launch; backward = reduce + dx in one launch.  Same arithmetic and the same coef[6*C] as
 * crn_brn_finalize.  Because blocks of a launch may start after others finished, num_batches_tracked is not touched
 * here: crn_brn_nbt_snapshot(counters[n] (device array of pointers), n, snapshot[n]) copies the counters of all fused
 * instances of a forward pass and increments the originals in ONE launch; crn_brn_fwd_fused reads its snapshot entry.
 * dxsum (double[C]) is WRITTEN (not accumulated). */
int crn_brn_fused_supported(int64_t rows, int32_t C);
int crn_brn_nbt_snapshot(int64_t* const* counters, int32_t n, int64_t* snapshot, void* stream);
int crn_brn_fwd_fused(const float* x, int64_t rows, int32_t C, int32_t x_cs, int32_t x_co, int32_t relu_in,
                      const float* weight, const float* bias, float* running_mean, float* running_var,
                      const int64_t* nt_snapshot, float eps, float momentum, int32_t training, const float* res,
                      int32_t relu_out, float* y, int32_t y_cs, int32_t y_co, float* y_pre, float* coef, void* stream);
int crn_brn_bwd_fused(const float* dy, int32_t dy_cs, int32_t dy_co, const float* y_act, const float* g_extra,
                      const float* x, int32_t x_cs, int32_t x_co, int64_t rows, int32_t C, const float* coef,
                      int32_t relu_in, int32_t relu_out, int32_t training, float* g_out, float* dx, int32_t dx_cs,
                      int32_t dx_co, int32_t dx_accumulate, float* dweight, float* dbias, double* dxsum, void* stream);

/* ------------------------------------------------------------------------
 * Small encoder ops (model/resnet50.py:134-140,176-204).
 * ---------------------------------------------------------------------- */
/* u8 NCHW RGB -> f32 NHWC4 BGR + Caffe mean (added, bug-for-bug), 4th channel 0. */
int crn_preprocess_image(const uint8_t* image, int32_t N, int32_t H, int32_t W, float* out,
                         void* stream);
/* ZeroPad2d(1) + MaxPool2d(3, stride 2) on NHWC (input is post-ReLU). idx: int8 argmax tap. */
int crn_maxpool_fwd(const float* x, int32_t N, int32_t H, int32_t W, int32_t C, float* y,
                    int8_t* idx, void* stream);
int crn_maxpool_bwd(const float* dy, const int8_t* idx, int32_t N, int32_t H, int32_t W,
                    int32_t C, float* dx, void* stream);
/* y[n][c] = mean over HW of x[n][hw][c];  bwd: dx[n][hw][c] (+)= dy[n][c]/HW */
int crn_spatial_mean_fwd(const float* x, int32_t N, int32_t HW, int32_t C, float* y, void* stream);
int crn_spatial_mean_bwd(const float* dy, int32_t N, int32_t HW, int32_t C, float* dx,
                         int32_t accumulate, void* stream);
/* out[c] = sum over rows of x[row][c] (double accumulators, out is float[C]). */
int crn_colsum(const float* x, int64_t rows, int32_t C, int32_t x_cs, int32_t x_co, float* out,
               double* scratch, void* stream);
/* planar variant: x is [N][C][S]; out[c] = sum_n sum_s */
int crn_colsum_planar(const float* x, int32_t N, int32_t C, int64_t S, float* out, double* scratch,
                      void* stream);
/* planar [N][C][S] -> channels-last rows [N*S][CP] (zero padded), for the logits gradient. */
int crn_planar_to_rows(const float* x, int32_t N, int32_t C, int64_t S, int32_t CP, float* out,
                       void* stream);

/* ------------------------------------------------------------------------
 * Ray-traced skip connection.  Replaces the index/gather part of
 * SampleGrid2d.forward, model/ray_traced_skip_connection.py:91-142, and the
 * torch.cat at model/reconstruction_decoder.py:117.
 *   map   : compressed 2-D map, channels-last [N, h, w, map_cs]
 *   m     : float[N*16] row-major layer matrix (v2s @ scale(res/g))
 *   offs  : float[N*3] voxel sample offsets
 *   out   : 3-D grid [N, g, g, g, out_cs]; writes channels [out_co, out_co+C)
 * Nearest-pixel with truncation toward zero, clamp into the 1-px border
 * (= outside_value 0), zero behind the camera: bit-exact data movement.
 * ---------------------------------------------------------------------- */
int crn_skip_sample_fwd(const float* map, int32_t N, int32_t h, int32_t w, int32_t C,
                        int32_t map_cs, const float* m, const float* offs, int32_t gD, int32_t gH,
                        int32_t gW, float* out, int32_t out_cs, int32_t out_co, void* stream);
/* dmap (zeroed by caller) += scatter of dout channels [out_co, out_co+C). */
int crn_skip_sample_bwd(const float* dout, int32_t out_cs, int32_t out_co, int32_t N, int32_t h,
                        int32_t w, int32_t C, int32_t map_cs, const float* m, const float* offs,
                        int32_t gD, int32_t gH, int32_t gW, float* dmap, void* stream);
/* Deterministic backward (no atomics): crn_skip_build_lists sorts the voxels of all N scenes by the pixel they sample
 * (stable radix sort of the keys n*h*w + pixel; voxels that sample nothing go last) into sorted_vox int32[N*g^3] and
 * writes starts int32[N*h*w + 1] (the voxels of pixel p are sorted_vox[starts[p] .. starts[p+1])).  The lists depend
 * only on (m, offs): build them once per step, off the critical path.  crn_skip_sample_bwd_sorted then WRITES
 * dmap[p] = sum over the list of p of dout channels [out_co, out_co+C), in list order: bit-reproducible.
 * The reference's backward is an atomic index_put (ray_traced_skip_connection.py:135-142 under autograd). */
int64_t crn_skip_lists_workspace_bytes(int32_t N, int32_t h, int32_t w, int32_t gD, int32_t gH, int32_t gW);
int crn_skip_build_lists(int32_t N, int32_t h, int32_t w, const float* m, const float* offs, int32_t gD, int32_t gH,
                         int32_t gW, void* workspace, int64_t workspace_bytes, int32_t* sorted_vox, int32_t* starts,
                         void* stream);
int crn_skip_sample_bwd_sorted(const float* dout, int32_t out_cs, int32_t out_co, int32_t N, int32_t h, int32_t w,
                               int32_t C, int32_t map_cs, const int32_t* sorted_vox, const int32_t* starts,
                               float* dmap, void* stream);
/* The int32 pixel index (iy*(w+2)+ix into the padded map, or -1 behind the
 * camera) for every voxel: test/debug hook for bit-exact index parity. */
int crn_skip_indices(int32_t N, int32_t h, int32_t w, const float* m, const float* offs,
                     int32_t gD, int32_t gH, int32_t gW, int32_t* idx, void* stream);

/* ------------------------------------------------------------------------
 * Losses on planar logits [N, C, S].  Replaces model/losses.py:64-114
 * (iou_fgbg) and :144-160 (xent_times_iou_agnostic), and its two factors on their own (:19-61 iou_agnostic,
 * :117-141 xent; modes 2 and 3 share the sums / backward passes of mode 1).  gt: int32 or int64 [N,S].
 * Pass 1 accumulates per-scene sums (double[N*4]: inter, union, xent, -);
 * the scalar combination is done by the host wrapper; pass 2 writes dlogits.
 * ---------------------------------------------------------------------- */
int crn_loss_sums(const float* logits, const void* gt, int32_t gt_is_i64, int32_t N, int32_t C,
                  int64_t S, int32_t mode /*0 iou_fgbg, 1 (1+iou_agnostic)(1+xent), 2 iou_agnostic, 3 xent*/, double* sums,
                  void* stream);
/* loss[0] and coef float[2N+1] = (dL/dI_n, dL/dU_n)..., dL/d(sum xent) from the per-scene sums. */
int crn_loss_finalize(const double* sums, int32_t N, int32_t C, int64_t S, int32_t mode, float* loss,
                      float* coef, void* stream);
/* dlogits = gscale * (coef-weighted derivative); gscale: device float (upstream grad) or NULL = 1. */
int crn_loss_bwd(const float* logits, const void* gt, int32_t gt_is_i64, int32_t N, int32_t C,
                 int64_t S, int32_t mode, const float* coef, const float* gscale, float* dlogits,
                 void* stream);
/* Eval: softmax over C (planar) and argmax -> confusion matrix counts
 * (evaluation_results.py:249-254, voxel_metrics.py:33-58).  cm: int64[C*C] zeroed by caller. */
int crn_softmax_planar(const float* logits, int32_t N, int32_t C, int64_t S, float* pmf,
                       void* stream);
int crn_argmax_confusion(const float* logits, const void* gt, int32_t gt_is_i64, int32_t N,
                         int32_t C, int64_t S, int64_t* cm, void* stream);
/* FG_BG evaluation form (evaluation_results.py:40-51): predicted and ground-truth labels (0/1) of scene n are
 * multiplied by scene_label[n] (device int32[N], values in [0, K/(C-1))) before counting into cm int64[K, K];
 * scene_label == NULL requires K == C and is crn_argmax_confusion. */
int crn_argmax_confusion_labeled(const float* logits, const void* gt, int32_t gt_is_i64, int32_t N, int32_t C,
                                 int64_t S, const int32_t* scene_label, int32_t K, int64_t* cm, void* stream);

/* Layout-aware forms: rows_cp == 0 -> planar logits [N, C, S] (the reference's NCDHW); rows_cp > 0 -> channels-last
 * rows [N*S, rows_cp] (rows_cp = C rounded up to a multiple of 4) as the tcgen05 epilogue of the last transposed
 * convolution writes them for C > 4: the training step then never materialises planar logits / logit gradients.
 * crn_loss_bwd_l writes dlogits in the layout d_rows_cp selects (pad channels written as 0). */
int crn_loss_sums_l(const float* logits, int32_t rows_cp, const void* gt, int32_t gt_is_i64, int32_t N, int32_t C,
                    int64_t S, int32_t mode, double* sums, void* stream);
int crn_loss_bwd_l(const float* logits, int32_t rows_cp, const void* gt, int32_t gt_is_i64, int32_t N, int32_t C,
                   int64_t S, int32_t mode, const float* coef, const float* gscale, float* dlogits,
                   int32_t d_rows_cp, void* stream);
int crn_softmax_l(const float* logits, int32_t rows_cp, int32_t N, int32_t C, int64_t S, float* pmf, void* stream);
int crn_argmax_confusion_l(const float* logits, int32_t rows_cp, const void* gt, int32_t gt_is_i64, int32_t N,
                           int32_t C, int64_t S, const int32_t* scene_label, int32_t K, int64_t* cm, void* stream);
/* channels-last rows [N*S][CP] -> planar [N][C][S] (the boundary layout of CoreNet.forward). */
int crn_rows_to_planar(const float* rows, int32_t N, int32_t C, int64_t S, int32_t CP, float* out, void* stream);

/* ------------------------------------------------------------------------
 * fill_inside_voxels.  Replaces cc/fill_voxels_gpu.cu:136-171 (kernels
 * :96-132) / cc/fill_voxels_cpu.cc:158-183.  grid: [N, D, H, W] of `elem_size`
 * byte elements of kind `dtype_kind` (0 = signed int, 1 = float, 2 = unsigned int);
 * any value > 0 is occupied.  Every voxel is overwritten with 0/1 in the
 * element type.  Empty voxels connected (6-neighbourhood) to the x=0, y=0 or
 * z=0 faces stay 0 (near-face rule, bug-for-bug); all others become 1.
 * workspace: crn_fill_workspace_bytes(N,D,H,W) bytes.
 * ---------------------------------------------------------------------- */
int64_t crn_fill_workspace_bytes(int32_t N, int32_t D, int32_t H, int32_t W);
int crn_fill_inside(const void* grid_in, void* grid_out, int32_t elem_size, int32_t dtype_kind,
                    int32_t N, int32_t D, int32_t H, int32_t W, void* workspace, void* stream);

/* ------------------------------------------------------------------------
 * voxelize_mesh.  Replaces geometry/voxelization.py:32-164 +
 * geometry/shaders/voxelize.geom:37-60 + voxelize.frag:29-58 (the EGL/GLSL
 * rasteriser).  triangles: float[T*9] view space; tri_mesh: int32[T] mesh id;
 * view2voxel: float[M*16] row-major.  grid: float[M, gD, gH, gW] ZEROED by the
 * caller, where (gD,gH,gW) = (D,H,W) or (2D+1,2H+1,2W+1) with sub_grid != 0.
 * ---------------------------------------------------------------------- */
int crn_voxelize_mesh(const float* triangles, const int32_t* tri_mesh, int32_t T,
                      const float* view2voxel, int32_t M, int32_t D, int32_t H, int32_t W,
                      int32_t image_resolution, int32_t depth_mult, int32_t sub_grid_side,
                      int32_t conservative, float* grid, void* stream);
/* grid[b] = max_m(label_m * occ_m) as int32 (data/batched_example.py:186-196).
 * mesh_scene: int32[M] scene index per mesh; labels: float[M]. out zeroed by caller. */
int crn_merge_mesh_grids(const float* mesh_grids, const int32_t* mesh_scene, const float* labels,
                         int32_t M, int64_t voxels, int32_t* out, void* stream);

/* Conv3d k=5 s=1 p=2 forward / dgrad on tcgen05 tensor cores (3xTF32, fp32 TMEM accumulators):
 * replaces the cuDNN call behind nn.Conv3d(k=5) at model/reconstruction_decoder.py:66,74,82,91.
 * Weights are pre-split (hi/lo) and packed per 8-channel pass by crn_tc5_pack (w is the PyTorch
 * parameter [Cout][Cin][5][5][5]; dgrad != 0 packs the flipped/transposed operator); out must hold
 * crn_tc5_packed_floats(K, N) floats.  kind 0: y = conv(x)+bias, kind 1: dx = conv^T(dy).
 * The grid must tile by 8 (x) x 16 (y) x 8 (z); N <= 64.  *status (device int) is set to 1 if an
 * internal barrier wait timed out (the result is then invalid). */
int64_t crn_tc5_packed_floats(int32_t K, int32_t N);
int crn_tc5_pack(const float* w, int32_t Cout, int32_t Cin, int32_t dgrad, float* out, void* stream);
int crn_conv5_tc(const crn_conv_desc* d, int32_t kind, const float* in, const float* wtc, const float* bias,
                 float* out, int32_t* status, void* stream);

/* Conv3d k=5 forward for Cout <= 16 with the 5 kz taps stacked into N (csrc/conv_tc5s.cu): one staged input plane
 * feeds the accumulators of the five output planes it contributes to in ONE tcgen05.mma (N = 5 x 32), which amortises
 * the shared-memory A fetch that bounds crn_conv5_tc at these channel counts.  Same contract as crn_conv5_tc kind 0;
 * the grid must tile by 8 (x) x 16 (y) x 4 (z).  Replaces the cuDNN call behind model/reconstruction_decoder.py:91. */
int64_t crn_tc5s_packed_floats(int32_t K);   /* K = reduction channels: Cin (forward) / Cout (dgrad) */
int crn_tc5s_pack(const float* w, int32_t Cout, int32_t Cin, float* out, void* stream);
int crn_conv5_tcs(const crn_conv_desc* d, const float* x, const float* wtc, const float* bias, float* y,
                  int32_t* status, void* stream);
/* General form: up to 32 output channels (17..32: the three hi/lo products as three stacked MMAs) and dgrad
 * (kind 1: dx = conv5^T(dy), weights packed with dgrad != 0 = flipped taps, transposed channels). */
int crn_tc5s_pack2(const float* w, int32_t Cout, int32_t Cin, int32_t dgrad, float* out, void* stream);
int crn_conv5_tcs2(const crn_conv_desc* d, int32_t kind, const float* in, const float* wtc, const float* bias,
                   float* out, int32_t* status, void* stream);

/* ConvTranspose3d k=7 s=2 p=3 output_padding=1 forward on the same tcgen05 kernel: seen from the input
 * grid all 8 output parity classes form one stride-1 4x4x4-tap convolution with 8*Cout columns, whose
 * epilogue scatters column (class, co) of input voxel i to output voxel 2i+class.  Replaces the cuDNN
 * call behind nn.ConvTranspose3d at model/reconstruction_decoder.py:69,77,85,95.  w is the PyTorch
 * parameter [Cin][Cout][7][7][7]; out of crn_tct_pack holds crn_tct_packed_floats(Cin, Cout) floats.
 * Input grid must tile by 8 (x) x 16 (y) x 8 (z); Cout <= 16; y may be channels-last or planar (desc). */
int64_t crn_tct_packed_floats(int32_t Cin, int32_t Cout, int32_t dgrad);
int crn_tct_pack(const float* w, int32_t Cin, int32_t Cout, int32_t dgrad, float* out, void* stream);
int crn_convt7_tc(const crn_conv_desc* d, const float* x, const float* wtc, const float* bias, float* y,
                  int32_t* status, void* stream);
/* ... and its dgrad, dx = convT^T(dy): the same kernel with K = 8*Cout class channels gathered from dy
 * (channels-last) and the taps mirrored (crn_tct_pack with dgrad != 0).  Cout % 4 == 0, Cin <= 64. */
int crn_convt7_tc_dgrad(const crn_conv_desc* d, const float* dy, const float* wtc, float* dx,
                        int32_t* status, void* stream);

/* Implicit-GEMM convolution forward / dgrad on tcgen05 for the wide layers (both channel counts >= 32): the
 * ResNet-50 encoder's Conv2d 1x1 / 3x3 (model/resnet50.py:61-70,94-108,122-131) and the coarse decoder Conv3d
 * layers (model/reconstruction_decoder.py:66,74).  3xTF32 with fp32 TMEM accumulators flushed every 72 MMAs.
 * Weights are pre-split / pre-packed per (output-channel tile, 16-channel K stage) by ONE crn_gemm_tc_pack launch
 * over a DEVICE item list: src is the PyTorch conv parameter [Cout][Cin][taps], dst holds
 * crn_gemm_tc_packed_floats(K, N, taps) floats with (K, N) = (Cin, Cout) for the forward operator and
 * (Cout, Cin) for dgrad != 0 (flipped taps, transposed channels).  offsets: DEVICE int64[n+1] prefix sums of
 * packed_floats/2 per item, total = offsets[n].
 * crn_conv_gemm_tc kind 0: y = conv(x) + bias (any stride); kind 1: dx = conv^T(dy), stride-1 convolutions only.
 * accumulate != 0 adds onto out (bias ignored).  *status is set to 1 if an internal barrier wait timed out. */
typedef struct {
  const float* src;
  float* dst;
  int32_t Cout, Cin, taps, dgrad;
} crn_gemm_tc_pack_item;
int64_t crn_gemm_tc_packed_floats(int32_t K, int32_t N, int32_t taps);
int crn_gemm_tc_pack(const crn_gemm_tc_pack_item* items, const int64_t* offsets, int32_t n, int64_t total,
                     void* stream);
int crn_conv_gemm_tc(const crn_conv_desc* d, int32_t kind, const float* in, const float* wtc, const float* bias,
                     float* out, int32_t accumulate, int32_t* status, void* stream);

/* Weight gradient of the wide layers on tcgen05 (3xTF32; both operands are MN-major tf32 UMMA operands in the
 * 128B-swizzle / 32B-base layout).  Same contract as crn_conv_wgrad: dWf[tap][ci][co] += sum_rows x * dy, dw zeroed
 * by the caller before the first call; plain and transposed convolutions, any stride.  Replaces the cuDNN
 * backward-filter calls behind model/resnet50.py:61-70,94-108,122-131 and model/reconstruction_decoder.py:66-77. */
int crn_conv_wgrad_tc(const crn_conv_desc* d, const float* x, const float* dy, float* dw_packed, int32_t* status,
                      void* stream);

/* Weight gradient of the narrow Conv3d k=5 s=1 p=2 layers (Cin <= 32 & Cout <= 16 at W = 64 / 32, or Cin <= 64 &
 * Cout <= 32 at W = 32) on tcgen05: filter taps are stacked into the M / N dimensions of the MMA through the
 * MN-major descriptor strides (csrc/conv_wgrad_line.cu).  Same contract as crn_conv_wgrad.  Replaces the cuDNN
 * backward-filter call behind model/reconstruction_decoder.py:82,91. */
int crn_conv_wgrad_line_supported(const crn_conv_desc* d);
int crn_conv_wgrad_line(const crn_conv_desc* d, const float* x, const float* dy, float* dw_packed, int32_t* status,
                        void* stream);

/* ... of the wide coarse Conv3d k=5 layers (Cin >= 64, 32 <= Cout <= 128 at W = 16 / 8: stage_4.c1, stage_3.c1): one
 * CTA per (kz, ky), the 5 kx taps are start-address shifts of the image row staged once. */
int crn_conv_wgrad_xline_supported(const crn_conv_desc* d);
int crn_conv_wgrad_xline(const crn_conv_desc* d, const float* x, const float* dy, float* dw_packed, int32_t* status,
                         void* stream);
/* ... and of ConvTranspose3d k=7 s=2 p=3 with Cin <= 32, Cout == 16 at coarse W = 32 / 16 (stage_5.t1,
 * model/reconstruction_decoder.py:85): the class-channel view of dy on the coarse grid makes it a stride-1 problem
 * whose N blocks are the fine lines themselves. */
int crn_convt7_wgrad_line_supported(const crn_conv_desc* d);
int crn_convt7_wgrad_line(const crn_conv_desc* d, const float* x, const float* dy, float* dw_packed, int32_t* status,
                          void* stream);

/* dgrad of ConvTranspose3d k=7 s=2 p=3 for Cin <= 32, Cout % 4 == 0 with the four jz taps STACKED INTO N (one staged
 * plane of the class-channel view of dy feeds the accumulators of four coarse output planes: N = 4 x 32 per
 * tcgen05.mma instead of 32; same idea as crn_conv5_tcs2).  Weights packed by crn_tcts_pack
 * (crn_tcts_packed_floats(Cout) floats).  Coarse grid must tile by 8 (x) x 16 (y) x 4 (z).
 * (model/reconstruction_decoder.py:85,93 backward.) */
/* ... and the FORWARD of that layer for Cout <= 2 (the FG_BG logits layer: 8 parity classes x 2 channels = 16
 * accumulator columns per coarse plane, four planes stacked per tcgen05.mma); y planar [N, Cout, 2D, 2H, 2W] or rows. */
int64_t crn_tctsf_packed_floats(int32_t Cin);
int crn_tctsf_pack(const float* w, int32_t Cin, int32_t Cout, float* out, void* stream);
int crn_convt7_tcs_fwd(const crn_conv_desc* d, const float* x, const float* wtc, const float* bias, float* y,
                       int32_t* status, void* stream);
int64_t crn_tcts_packed_floats(int32_t Cout);
int crn_tcts_pack(const float* w, int32_t Cin, int32_t Cout, float* out, void* stream);
int crn_convt7_tcs_dgrad(const crn_conv_desc* d, const float* dy, const float* wtc, float* dx, int32_t* status,
                         void* stream);

/* Fused Adam step over a flat list (state.py:65-66) — next-row (f2) op. */
int crn_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                  float beta2, float eps, int32_t step, float grad_scale, void* stream);

/* Batched fp64 -> fp32 copies: items / offsets are DEVICE arrays (offsets = int64[n+1] prefix sums of the item
 * lengths, total = offsets[n]); used to move all bias gradients (fp64 column sums) of a backward pass at once. */
typedef struct {
  const double* src;
  float* dst;
} crn_f64_copy_item;
int crn_gather_f64_to_f32(const crn_f64_copy_item* items, const int64_t* offsets, int32_t n, int64_t total,
                          void* stream);

/* Same with the step counter on the device (*step_dev is incremented by the call, then used for the bias
 * corrections): the form a captured CUDA graph can replay. */
int crn_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                      float beta2, float eps, int32_t* step_dev, float grad_scale, void* stream);
/* Guarded form: when *status (device word the tcgen05 kernels set on an mbarrier timeout; may be NULL) is non-zero the
 * step is skipped entirely (weights, moments and the counter stay as they were), so a failed step cannot reach the
 * parameters.  Works on a sub-range too: the counter is only bumped when bump_step != 0 (chunked updates call it once
 * per chunk with bump_step set on the first chunk only). */
int crn_adam_step_guarded(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                          float beta2, float eps, int32_t* step_dev, int32_t bump_step, float grad_scale,
                          const int32_t* status, void* stream);
/* Failure surfacing: if *status != 0 writes NaN to out[i * stride] for i < n (one element per scene of the logits is
 * enough to turn every loss into NaN).  No-op otherwise. */
int crn_status_poison(const int32_t* status, float* out, int64_t stride, int32_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CORENET_B200_H_ */
