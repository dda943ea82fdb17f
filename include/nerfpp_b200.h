/*
 * nerfpp_b200.h -- C ABI of the B200-native NeRF++ ray-marching hot path.
 *
 * Drop-in boundary for the one path BASELINE.json names: hierarchical
 * sampling -> positional-encoded 8x256 MLP -> transmittance composite -> rgb + depth-prior loss.
 * The reference (cwchenwang/outdoor-nerf-depth) has no FFI for this path -- it is plain PyTorch
 * modules -- so every entry point cites the reference Python symbol it replaces
 * (paths relative to nerf-methods/nerfplusplus/).  The reference-side binding is the ctypes stub in
 * outdoor-nerf-depth_b200/nerfpp_b200/_lib.py, see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes; no torch / C++ types cross this boundary.
 *   - every pointer is a DEVICE pointer to contiguous row-major fp32 unless stated otherwise.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     enqueued on it, nothing synchronises unless stated.
 *   - the library owns nothing past a call: outputs and workspaces are caller-allocated.
 *   - return value: 0 on success, >0 a cudaError_t, <0 an argument error; never throws.
 *     nerfpp_last_error() returns a static, thread-local text for the last non-zero return.
 *   - compiled for sm_100a only; there is no CPU path.
 */
#ifndef NERFPP_B200_H_
#define NERFPP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NERFPP_ABI_VERSION 1

/* Network shape: the only one the reference scripts use (configs/kitti.txt:36-40):
 * netdepth 8, netwidth 256, skips [4], max_freq_log2 10, max_freq_log2_viewdirs 4. */
#define NERFPP_WIDTH 256
#define NERFPP_DEPTH 8
#define NERFPP_NFREQ_POS 10
#define NERFPP_NFREQ_VIEW 4
#define NERFPP_NLAYERS 12 /* base 0..7, sigma, base_remap, rgb.0, rgb.2 */

/* Parameters of one MLPNet (nerf_network.py:70-118) as the torch state_dict holds them:
 * w[i] is nn.Linear.weight [out,in] row-major, b[i] is nn.Linear.bias [out].
 * order: base_layers.0..7.0, sigma_layers.0, base_remap_layers.0, rgb_layers.0, rgb_layers.2 */
typedef struct NerfppNetParams {
  const float* w[NERFPP_NLAYERS];
  const float* b[NERFPP_NLAYERS];
} NerfppNetParams;

/* Gradients wrt the same tensors (same shapes). Accumulated INTO (+=), caller zero-fills. */
typedef struct NerfppNetGrads {
  float* w[NERFPP_NLAYERS];
  float* b[NERFPP_NLAYERS];
} NerfppNetGrads;

/* Field evaluators. TC = tcgen05 tensor-core path (fp16 operands, fp32 accumulate in TMEM);
 * SIMT = plain fp32 FFMA path (bit-for-bit fp32 arithmetic, ~20x slower; the cross-check). */
enum {
  NERFPP_FIELD_TC = 0,        /* tcgen05 tensor cores, fp16 operands (one pass), fp32 accumulate: the fast path */
  NERFPP_FIELD_SIMT = 1,      /* plain fp32 on CUDA cores: cross-check */
  NERFPP_FIELD_TC_SPLIT = 2   /* tcgen05, every operand carried as a hi + lo fp16 pair (3 MMA passes, ~22-bit operands):
                                 the full-precision inference mode -- meets 1e-4 on trained weights at ~1/3 of the rate */
};

int nerfpp_abi_version(void);
const char* nerfpp_last_error(void);

/* ---- A1: intersect_sphere (ddp_train_nerf.py:51-66) -------------------------------------- */
/* out_far[n]; *out_unbounded (int32, device) is set non-zero when any ||p||^2 >= 1, the
 * condition on which the reference raises (ddp_train_nerf.py:62-63).  Caller zero-fills it. */
int nerfpp_intersect_sphere(const float* ray_o, const float* ray_d, int n_rays, float* out_far,
                            int32_t* out_unbounded, void* stream);

/* ---- A2+A3: coarse depths (ddp_train_nerf.py:441-449 train / 168-175 test) ---------------- */
/* fg: z[i] = near + i*((far-near)/(S-1)).  bg: z[i] = bg_base[i] (the reference's
 * torch.linspace(0,1,S), supplied by the caller so its rounding is the framework's).
 * t_rand_fg / t_rand_bg [n,S]: the torch.rand_like draw of perturb_samples
 * (ddp_train_nerf.py:69-78); NULL = no perturbation (test-time path). */
int nerfpp_coarse_depths(const float* near, const float* far, const float* bg_base, int n_rays,
                         int n_samples, const float* t_rand_fg, const float* t_rand_bg,
                         float* out_fg_z, float* out_bg_z, void* stream);

/* perturb_samples (ddp_train_nerf.py:69-78) on caller-supplied depths z [n,S]; out_z must not alias z. */
int nerfpp_perturb_samples(const float* z, const float* t_rand, int n_rays, int n_samples, float* out_z,
                           void* stream);

/* ---- A4: sample_pdf (ddp_train_nerf.py:81-130) ------------------------------------------- */
/* bins [n, M+1] (row stride bins_ld), weights [n, M] (row stride w_ld, so the trainer's
 * weights[..., 1:-1] slice needs no copy), u [n, Ns] with row stride u_ld (0 = one shared row:
 * the det=True linspace).  out_samples [n,Ns]; out_above (int32 [n,Ns]) and out_cdf
 * ([n,M+1]) optional (NULL to skip).  Index contract: above = #{j<M : u >= cdf[j]}. */
int nerfpp_sample_pdf(const float* bins, int bins_ld, const float* weights, int w_ld,
                      const float* u, int u_ld, int n_rays, int M, int n_new,
                      float* out_samples, int32_t* out_above, float* out_cdf, void* stream);
/* Same search + interpolation on a caller-supplied cdf [n, M+1]: given identical (cdf,u) the
 * indices and samples are bit-identical to the reference's. */
int nerfpp_sample_cdf(const float* bins, int bins_ld, const float* cdf, int cdf_ld,
                      const float* u, int u_ld, int n_rays, int M, int n_new,
                      float* out_samples, int32_t* out_above, void* stream);

/* ---- A4+A5 fused: one cascade refinement (ddp_train_nerf.py:452-457 / 460-465) ------------ */
/* mids of z_prev -> sample_pdf(weights_prev[1:-1]) -> sort(cat(z_prev, new)).
 * z_prev, w_prev [n,Sp]; u [n,Ns] (u_ld 0 = shared row); out_z [n, Sp+Ns] ascending. */
int nerfpp_resample_merge(const float* z_prev, const float* w_prev, const float* u, int u_ld,
                          int n_rays, int n_prev, int n_new, float* out_z, void* stream);
/* The foreground (:452-457) and background (:460-465) refinements of one level in ONE launch: 4096 rays fill less than
 * half of a B200's warp slots, the two nets' rays run side by side.  Same arithmetic as two nerfpp_resample_merge calls;
 * fg_u / bg_u share u_ld (0 = one shared row each). */
int nerfpp_resample_merge_pair(const float* fg_z_prev, const float* fg_w_prev, const float* fg_u, float* out_fg_z,
                               const float* bg_z_prev, const float* bg_w_prev, const float* bg_u, float* out_bg_z,
                               int u_ld, int n_rays, int n_prev, int n_new, void* stream);

/* ---- packed weights ---------------------------------------------------------------------- */
/* The field kernels read a repacked copy of one net's parameters (transposed / padded fp32
 * for SIMT, fp16 UMMA-swizzled K-major tiles + fp32 tails for TC). Size query then pack;
 * re-pack whenever the parameters change (the Python shim keys it on tensor._version). */
int64_t nerfpp_packed_bytes(int is_bg, int field_impl);
int nerfpp_pack_weights(const NerfppNetParams* params, int is_bg, int field_impl, void* out_packed,
                        void* stream);

/* ---- A6+A7+A8: embed + MLP per sample (nerf_network.py:42-60,120-142; ddp_model.py:16-45) -- */
/* Evaluates one net on every sample of every ray.
 *   fg (is_bg=0): pts = ray_o + z*ray_d, sample order as given.
 *   bg (is_bg=1): pts = depth2pts_outside(ray_o, ray_d, z) = (x',y',z',1/r); outputs are written
 *                 in the reference's FLIPPED order (ddp_model.py:116-117): out[., j] belongs to
 *                 z[., S-1-j]; out_depth_real [n,S] likewise (may be NULL for fg).
 * out_sigma [n,S] (after abs), out_rgb [n,S,3] (after sigmoid). */
int nerfpp_field_forward(const void* packed, int is_bg, int field_impl, const float* ray_o,
                         const float* ray_d, const float* z, int n_rays, int n_samples,
                         float* out_sigma, float* out_rgb, float* out_depth_real, void* stream);

/* Training-mode field evaluation (tensor-core path only): same outputs, and additionally saves what the backward
 * needs into `train_workspace` (nerfpp_field_train_workspace_bytes(n_rays, n_samples) bytes, ~5.1 KB per sample):
 * every layer's fp16 activations in the MMA operand layout, the encoded inputs, and sigma before the abs(). */
int64_t nerfpp_field_train_workspace_bytes(int n_rays, int n_samples);
int nerfpp_field_forward_train(const void* packed, int is_bg, const float* ray_o, const float* ray_d,
                               const float* z, int n_rays, int n_samples, float* out_sigma, float* out_rgb,
                               float* out_depth_real, void* train_workspace, void* stream);

/* depth2pts_outside (ddp_model.py:16-45) as a standalone op: ray_o, ray_d [n,3] and depth [n]
 * already expanded per sample -> out_pts [n,4] = (x',y',z',1/r), out_depth_real [n]. */
int nerfpp_depth2pts_outside(const float* ray_o, const float* ray_d, const float* depth, int64_t n,
                             float* out_pts, float* out_depth_real, void* stream);

/* ---- A9+A10: composite ("raw2outputs", ddp_model.py:95-134) ------------------------------- */
typedef struct NerfppRenderOut { /* the 10 keys of NerfNet.forward's OrderedDict */
  float* rgb;        /* [n,3] */
  float* fg_weights; /* [n,S_fg] */
  float* bg_weights; /* [n,S_bg] flipped order, as the reference returns it */
  float* fg_dists;   /* [n,S_fg] */
  float* fg_rgb;     /* [n,3] */
  float* fg_depth;   /* [n] */
  float* bg_rgb;     /* [n,3]  already scaled by bg_lambda */
  float* bg_depth;   /* [n]    already scaled by bg_lambda */
  float* bg_lambda;  /* [n] */
  float* depth;      /* [n] */
} NerfppRenderOut;

int nerfpp_composite(const float* ray_d, const float* fg_z_max, const float* fg_z,
                     const float* bg_z, const float* fg_sigma, const float* fg_rgb,
                     const float* bg_sigma, const float* bg_rgb, const float* bg_depth_real,
                     int n_rays, int s_fg, int s_bg, const NerfppRenderOut* out, void* stream);

/* Backward of nerfpp_composite (autograd of ddp_model.py:95-134).  `grads` holds the upstream gradients on the ten
 * outputs (const pointers in a NerfppRenderOut; NULL = zero; fg_dists carries none); `fwd` is the forward's output
 * struct (bg_lambda is read).  Writes gradients on the per-sample sigma (after abs) [n,S] and rgb (after sigmoid)
 * [n,S,3] of both nets, background in the forward's flipped order. */
int nerfpp_composite_backward(const float* ray_d, const float* fg_z_max, const float* fg_z, const float* bg_z,
                              const float* fg_sigma, const float* fg_rgb, const float* bg_sigma,
                              const float* bg_rgb, const float* bg_depth_real, int n_rays, int s_fg, int s_bg,
                              const NerfppRenderOut* fwd, const NerfppRenderOut* grads, float* d_fg_sigma,
                              float* d_fg_rgb, float* d_bg_sigma, float* d_bg_rgb, void* stream);

/* ---- A11: NerfNet.forward (ddp_model.py:74-147) in one call -------------------------------- */
/* workspace: nerfpp_forward_workspace_bytes(n, s_fg, s_bg) bytes of device memory; holds the
 * per-sample sigma/rgb/depth_real (kept for backward). */
int64_t nerfpp_forward_workspace_bytes(int n_rays, int s_fg, int s_bg);
int nerfpp_forward(const void* packed_fg, const void* packed_bg, int field_impl,
                   const float* ray_o, const float* ray_d, const float* fg_z_max,
                   const float* fg_z, const float* bg_z, int n_rays, int s_fg, int s_bg,
                   const NerfppRenderOut* out, void* workspace, void* stream);

/* ---- training: forward that keeps what the backward needs, and the backward (autograd of NerfNet.forward) ---- */
/* nerfpp_forward_train = nerfpp_forward + a training workspace (1024-byte aligned,
 * nerfpp_forward_train_workspace_bytes bytes, ~5.1 KB per sample) holding both nets' fp16 activations.
 * nerfpp_backward: upstream gradients on the ten outputs (`grads`, NULL members = zero) -> gradients of the 2 x 24
 * parameter tensors, ACCUMULATED into grads_fg / grads_bg (caller zero-fills).  Tensor-core path only: operands fp16
 * under a power-of-two loss scale, accumulation fp32.  `workspace` / `train_workspace` are the forward's. */
int64_t nerfpp_forward_train_workspace_bytes(int n_rays, int s_fg, int s_bg);
int nerfpp_forward_train(const void* packed_fg, const void* packed_bg, const float* ray_o, const float* ray_d,
                         const float* fg_z_max, const float* fg_z, const float* bg_z, int n_rays, int s_fg, int s_bg,
                         const NerfppRenderOut* out, void* workspace, void* train_workspace, void* stream);
int64_t nerfpp_backward_workspace_bytes(int n_rays, int s_fg, int s_bg);
int nerfpp_backward(const NerfppNetParams* params_fg, const NerfppNetParams* params_bg, const float* ray_d,
                    const float* fg_z_max, const float* fg_z, const float* bg_z, int n_rays, int s_fg, int s_bg,
                    const NerfppRenderOut* out, const NerfppRenderOut* grads, const void* workspace,
                    const void* train_workspace, const NerfppNetGrads* grads_fg, const NerfppNetGrads* grads_bg,
                    void* bwd_workspace, void* stream);

/* ---- A12-A14: losses (utils.py:12-16, depth_loss.py:4-44, ddp_train_nerf.py:481-493) ------- */
enum { NERFPP_DEPTH_NONE = 0, NERFPP_DEPTH_MSE = 1, NERFPP_DEPTH_L1 = 2, NERFPP_DEPTH_KL = 3 };
/* out_loss[4] (device) = { img2mse(rgb, rgb_gt), depth loss, rgb + lambda*depth, #valid rays }.
 * mse/l1: mean over rays with depth_sup > 0 (NaN when none, like the reference);
 * kl: depth_kl(fg_weights, depth_sup, fg_z, fg_dists, kl_sigma, fg_z_max): sum over rays with
 * 0 < depth_sup < fg_z_max, divided by S.  workspace: nerfpp_loss_workspace_bytes(). */
int64_t nerfpp_loss_workspace_bytes(void);
int nerfpp_loss(const float* rgb, const float* rgb_gt, const float* depth, const float* depth_sup,
                const float* fg_weights, const float* fg_z, const float* fg_dists,
                const float* fg_z_max, int n_rays, int s_fg, int depth_loss_type,
                float lambda_depth, float kl_sigma, float* out_loss, void* workspace,
                void* stream);

/* The depth term alone (the drop-in depth_loss.depth_mse/l1/kl entry points): same out_loss[4]
 * layout with out_loss[0] = 0; and its backward: out_grad = d(loss)/d(depth) [n] for mse/l1,
 * d(loss)/d(fg_weights) [n,S] for kl, times the upstream scalar *grad_out (device). fwd_out is
 * the forward's out_loss (supplies the valid-ray count). */
int nerfpp_depth_loss(const float* depth, const float* depth_sup, const float* fg_weights,
                      const float* fg_z, const float* fg_dists, const float* fg_z_max, int n_rays,
                      int s_fg, int depth_loss_type, float kl_sigma, float* out_loss, void* workspace,
                      void* stream);
int nerfpp_depth_loss_backward(const float* depth, const float* depth_sup, const float* fg_weights,
                               const float* fg_z, const float* fg_dists, const float* fg_z_max,
                               int n_rays, int s_fg, int depth_loss_type, float kl_sigma,
                               const float* fwd_out, const float* grad_out, float* out_grad,
                               void* stream);

/* ---- N1: ray generation + batch gather on the device -------------------------------- */
/* get_rays_single_image (nerf_sample_ray_split.py:10-34) for the selected pixels only, fused with the gathers of
 * RaySamplerSingleImage.random_sample / get_all (:131-221).  kinv_host = inverse of the 3x3 intrinsics (row-major,
 * HOST memory, 9 floats), c2w_host = 4x4 camera-to-world (row-major, HOST memory), cam_depth = inv(c2w)[2,3].
 * pixel_ids: int64 [n] device (row-major pixel index v*W+u; NULL = all n = H*W pixels in order).  img_*: the image's
 * device-resident arrays [H*W,3] / [H*W] (NULL = absent; min depth then defaults to 1e-4).  Outputs [n,..]; rgb,
 * depth, depth_sup, min_depth may be NULL. */
int nerfpp_gen_rays(const float* kinv_host, const float* c2w_host, float cam_depth, int W, const int64_t* pixel_ids,
                    int64_t n, const float* img_rgb, const float* img_depth_sup, const float* img_min_depth,
                    float* ray_o, float* ray_d, float* depth, float* rgb, float* depth_sup, float* min_depth,
                    void* stream);

/* ---- N2: pixel decode of the on-disk formats ------------------------------------------------ */
/* RaySamplerSingleImage.set_resolution_level (nerf_sample_ray_split.py:73-102) turns the PNGs that load_data_split
 * (data_loader_split.py:27-129) found into float arrays on the host: rgb/mask = u8/255, min depth = u8/255*max_depth+1e-4,
 * depth{,_<type>} = depth_scale * (u16/256).  Here the raw pixels are uploaded as they are (1-2 bytes each) and decoded on
 * the device: out[i] = ((src[i] / div) * mul) + add, each step rounded to fp32 in numpy's order (bit-exact).
 * src: device, src_bits = 8 (uint8) or 16 (uint16); n elements. */
int nerfpp_decode_pixels(const void* src, int src_bits, int64_t n, float div, float mul, float add, float* out,
                         void* stream);

/* ---- N3: image metrics of the test loop (ddp_train_nerf.py:556-600) -------------------- */
/* out_metrics[8] (device) = mse, psnr = -10 log10(mse + 1e-6), #valid depth pixels, rmse, rmse_log, abs_diff, abs_rel,
 * sq_rel; depth metrics over pixels with 1e-3 < gt/depth_scale < cap (cap = 80 m in the reference), both sides clipped to
 * [1e-3, cap].  rgb_gt / depth_gt may be NULL (their outputs are then 0 / NaN).  workspace:
 * nerfpp_loss_workspace_bytes() * 2 bytes. */
int nerfpp_image_metrics(const float* rgb, const float* rgb_gt, const float* depth, const float* depth_gt,
                         int64_t n_pixels, float depth_scale, float cap, float* out_metrics, void* workspace,
                         void* stream);

/* ---- A16: mipnerf360 twins (config 3; nerf-methods/mipnerf360/internal/) ------------------------ */
/* stepfun.sample_intervals (stepfun.py:214-263) with use_gpu_resampling=False: softmax(w_logits) ->
 * integrate_weights -> sorted_interp(u) -> interval fenceposts (midpoints, reflected + clamped ends).
 * t [n, M+1] sorted bin edges, w_logits [n, M], u [n, Ns] inverse-CDF ordinates with row stride u_ld
 * (0 = one shared row: the rng=None deterministic centres); out_t [n, Ns+1].  The reference draws the
 * jitter with jax.random; the caller supplies u.  M, Ns <= 256. */
int mip360_sample_intervals(const float* t, const float* w_logits, const float* u, int u_ld, int n_rays,
                            int n_bins, int n_samples, float domain_min, float domain_max, float* out_t,
                            void* stream);
/* render.compute_alpha_weights (render.py:130-151): density [n,S], tdist [n,S+1], dirs [n,3] ->
 * weights, alpha, trans [n,S] (alpha / trans may be NULL). */
int mip360_compute_alpha_weights(const float* density, const float* tdist, const float* dirs, int n_rays,
                                 int n_samples, int opaque_background, float* out_weights, float* out_alpha,
                                 float* out_trans, void* stream);
/* render.volumetric_rendering (render.py:154-216), compute_extras=True, extras=None.  rgbs [n,S,3],
 * weights [n,S], tdist [n,S+1], bg_rgbs [3] (bg_ld 0) or [n,3] (bg_ld 3), t_far [n].
 * out_rgb [n,3]; out_scalars [n,6] = acc, distance_mean, depth (the fork's addition, :199-201),
 * distance_percentile_5, distance_median, distance_percentile_95.  S <= 256. */
int mip360_volumetric_rendering(const float* rgbs, const float* weights, const float* tdist,
                                const float* bg_rgbs, int bg_ld, const float* t_far, int n_rays, int n_samples,
                                float* out_rgb, float* out_scalars, void* stream);
/* Depth-prior losses of the mipnerf360 trainer (train_utils.py:108-129): NERFPP_DEPTH_KL =
 * depth_loss.depth_loss(..., 'kl') (depth_loss.py:66-97 -> ds_nerf_depth_loss :5-26: 1e-7, /(2*sigma), mean over
 * rays x samples with invalid rays zeroed but counted); NERFPP_DEPTH_MSE / _L1 on predicted_depth =
 * rendering['distance_mean'] with the mask multiplied in and the mean over all rays.  out_loss[1] (device). */
int64_t mip360_depth_loss_workspace_bytes(int n_rays);
int mip360_depth_loss(const float* weights, const float* tdist, const float* termination_depth,
                      const float* predicted_depth, const float* dirs, int n_rays, int n_samples,
                      int depth_loss_type, float sigma, float* out_loss, void* workspace, void* stream);

/* ---- N4 (partial): the regularisers of the mipnerf360 trainer (train_utils.py:160-180) ---------- */
/* stepfun.lossfun_outer (stepfun.py:82-89; inner_outer :64-79, searchsorted :30-53): the proposal histogram (t_env
 * [n,P+1], w_env [n,P]) must be an upper envelope of the NeRF histogram (t [n,S+1], w [n,S]):
 * out_loss [n,S] = max(0, w - w_outer)^2 / (w + eps).  interlevel_loss is the mean of it, summed over proposal levels.
 * With grad_loss [n,S] (d/d out_loss) also out_grad_w_env [n,P] -- the only input interlevel_loss differentiates
 * (c and w pass through stop_gradient).  out_loss may be NULL in a backward-only call.  S, P <= 256. */
int mip360_lossfun_outer(const float* t, const float* w, const float* t_env, const float* w_env, int n_rays, int n_bins,
                         int n_env_bins, float eps, const float* grad_loss, float* out_loss, float* out_grad_w_env,
                         void* stream);
/* stepfun.lossfun_distortion (stepfun.py:266-276): out_loss [n] = sum_ij w_i w_j |u_i - u_j| + sum_i w_i^2 (t_i+1 - t_i)/3,
 * u = interval midpoints; distortion_loss is its mean.  With grad_loss [n]: out_grad_t [n,S+1], out_grad_w [n,S]. */
int mip360_lossfun_distortion(const float* t, const float* w, int n_rays, int n_bins, const float* grad_loss,
                              float* out_loss, float* out_grad_t, float* out_grad_w, void* stream);
/* stepfun.max_dilate (weights_mode 0) / max_dilate_weights (weights_mode 1) (stepfun.py:99-128): max-pool the step
 * function (t [n,M+1], w [n,M]) over +-dilation; models.py:150-169 applies it to the proposal histogram between levels.
 * out_t [n,3M+1] = clip(sort(cat(t, t[:-1]-d, t[1:]+d)), domain), out_w [n,3M]; renormalize (weights_mode only) divides
 * by max(eps, sum).  t must be sorted (non-decreasing), as the reference assumes.  M <= 85. */
int mip360_max_dilate(const float* t, const float* w, int n_rays, int n_bins, float dilation, float domain_min,
                      float domain_max, int weights_mode, int renormalize, float eps, float* out_t, float* out_w,
                      void* stream);

/* ---- config 3 (BASELINE.json configs[2]): the mipnerf360 field -- Model.__call__'s per-level body, models.py:200-231 ---- */
/* One level's  s_to_t -> render.cast_rays -> MLP  (models.py:204-231): coord.construct_ray_warps with
 * raydist_fn = reciprocal (coord.py:63-99), conical frusta as full-covariance Gaussians (render.py:21-85, ray_shape
 * 'cone', diag=False), then MLP.__call__ (models.py:398-611) as configs/360.gin sets it up: warp_fn = coord.contract
 * carried through coord.track_linearize (coord.py:22-60), lift_and_diagonalize on the 21-vector icosahedron basis
 * (geopoly.generate_basis('icosahedron', 2)), integrated_pos_enc degrees 0..11 (504 features, coord.py:107-126),
 * net_depth x net_width Dense + ReLU with the encoding concatenated back after layer 4, density = softplus(raw - 1);
 * and, when has_rgb (NerfMLP), bottleneck 256 | pos_enc(viewdirs, 0, 4) -> Dense 128 + ReLU -> Dense 3 -> sigmoid with
 * rgb_padding 0.001.  (net_depth, net_width) = (4, 256) PropMLP / (8, 1024) NerfMLP; disable_density_normals = True.
 *
 * Parameters arrive as flax holds them -- Dense_i kernels [in, out] row-major + biases, in construction order:
 * Dense_0..Dense_{depth-1} trunk, Dense_depth density head, then (has_rgb) bottleneck, view layer, rgb head -- and are
 * packed once into fp16 [out, in_padded] operand images (prec = 1: hi + lo halves, three MMA passes per layer). */
typedef struct Mip360MlpParams {
  const float* kernel[12];
  const float* bias[12];
} Mip360MlpParams;
int64_t mip360_mlp_packed_bytes(int net_depth, int net_width, int has_rgb, int prec);
int mip360_mlp_pack(const Mip360MlpParams* params, int net_depth, int net_width, int has_rgb, int prec, void* packed,
                    void* stream);
int64_t mip360_field_workspace_bytes(int64_t n_samples, int net_depth, int net_width, int has_rgb, int prec);
/* sdist [n, S+1] normalised fenceposts, near / far / radii [n], origins / directions / viewdirs [n,3] ->
 * out_tdist [n, S+1] (metric fenceposts), out_density [n, S], out_rgb [n, S, 3] (NULL unless has_rgb).  workspace:
 * mip360_field_workspace_bytes(n*S, ...) bytes, 256-byte aligned. */
int mip360_field_forward(const void* packed, int net_depth, int net_width, int has_rgb, int prec, const float* sdist,
                         const float* near, const float* far, const float* origins, const float* directions,
                         const float* viewdirs, const float* radii, int n_rays, int n_samples, float* out_tdist,
                         float* out_density, float* out_rgb, void* workspace, void* stream);
/* The first stage alone (tests): out_enc [n*S, 512] fp16 = the 504 integrated-positional-encoding features of every
 * sample (+ 8 zero columns), out_means [n*S,3] / out_covs [n*S,9] = the contracted Gaussians (NULL = skip). */
int mip360_cast_encode(const float* sdist, const float* near, const float* far, const float* origins,
                       const float* directions, const float* viewdirs, const float* radii, int n_rays, int n_samples,
                       float* out_tdist, void* out_enc, void* out_enc_lo, void* out_dir, void* out_dir_lo,
                       float* out_means, float* out_covs, void* stream);
/* One Dense layer on the tensor cores (tests, microbenchmarks): out[M,N] = act(a[M,K] w[N,K]^T + bias), fp16 operands
 * and output, fp32 accumulation; K % 8 == 0, N % 128 == 0. */
int mip360_dense_f16(const void* a, const void* w, const float* bias, void* out, int M, int N, int K, int relu,
                     void* stream);
/* The whole resampling step between two levels of Model.__call__ (models.py:171-200) in one launch: the logits above built
 * inline from the weights, then stepfun.sample_intervals.  t / weights are read with row strides t_ld / w_ld (pass
 * dilated_t + 1 and dilated_w + 1 with the dilated histogram's strides: its [1:-1] slices need no copy; n_bins = the sliced
 * bin count).  u: with jitter == NULL the ordinates [n, Ns] (u_ld = Ns) or one shared row (u_ld 0); with jitter [n] (the
 * single_jitter draw in [0,1)) u is the shared base row (stepfun.py:203-209) and the ordinates are base + jitter * max_jitter. */
int mip360_resample_level(const float* t, int t_ld, const float* weights, int w_ld, int n_rays, int n_bins, float anneal,
                          float resample_padding, const float* u, int u_ld, const float* jitter, float max_jitter,
                          int n_samples, float domain_min, float domain_max, float* out_t, void* stream);
/* Model.__call__'s logits for the next resampling (models.py:171-185): where(sdist[1:] > sdist[:-1],
 * anneal * log(weights + padding), -inf).  sdist [n, M+1], weights [n, M] -> out_logits [n, M]. */
int mip360_resample_logits(const float* sdist, const float* weights, int n_rays, int n_bins, float anneal,
                           float resample_padding, float* out_logits, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERFPP_B200_H_ */
