// rgb MSE and the three depth-prior losses as one deterministic
// two-stage reduction: per-block partial sums in a fixed order, then one block folds the
// partials in index order.  Sums are carried in fp64 (cheap: a few values per ray) so the
// result does not depend on the grid shape.
#include "common.cuh"

namespace npp {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_MAX_BLOCKS = 1024;
// partial layout: [block][4] doubles = { sum rgb sq err, sum depth term, valid count, unused }

__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = (lane < LOSS_THREADS / 32) ? sh[lane] : 0.0;
    t = warp_sum(t);
  }
  return t;   // valid in warp 0
}

__global__ void __launch_bounds__(LOSS_THREADS)
loss_partial_kernel(const float* __restrict__ rgb, const float* __restrict__ rgb_gt, const float* __restrict__ depth,
                    const float* __restrict__ depth_sup, const float* __restrict__ fg_w, const float* __restrict__ fg_z,
                    const float* __restrict__ fg_dists, const float* __restrict__ fg_far, int n, int S, int type,
                    float kl_sigma, double* __restrict__ partial) {
  __shared__ double sh[LOSS_THREADS / 32];
  double s_rgb = 0.0, s_dep = 0.0, s_cnt = 0.0;
  int tid = blockIdx.x * LOSS_THREADS + threadIdx.x, nthr = gridDim.x * LOSS_THREADS;
  // img2mse, utils.py:12-14: mean over n*3 of (x-y)*(x-y)
  if (rgb)
    for (int i = tid; i < n * 3; i += nthr) {
      float e = rgb[i] - rgb_gt[i];
      s_rgb += (double)(e * e);
    }
  if (type == NERFPP_DEPTH_MSE || type == NERFPP_DEPTH_L1) {
    // depth_loss.py:4-18: masked mean over rays with gt > 0
    for (int i = tid; i < n; i += nthr) {
      float gt = depth_sup[i];
      if (gt > 0.f) {
        float e = gt - depth[i];
        s_dep += (double)(type == NERFPP_DEPTH_MSE ? e * e : fabsf(e));
        s_cnt += 1.0;
      }
    }
  } else if (type == NERFPP_DEPTH_KL) {
    // depth_loss.py:39-44: -log(w+1e-5) * exp(-(z-t)^2/(2 sigma)) * dists, summed over valid rays
    float two_s = 2.f * kl_sigma;
    long long tot = (long long)n * S;
    for (long long i = tid; i < tot; i += nthr) {
      int r = (int)(i / S);
      float t = depth_sup[r];
      if (t > 0.f && t < fg_far[r]) {
        float dz = fg_z[i] - t;
        s_dep += (double)(-logf(fg_w[i] + 1e-5f) * expf(-(dz * dz) / two_s) * fg_dists[i]);
      }
    }
    for (int i = tid; i < n; i += nthr) {
      float t = depth_sup[i];
      if (t > 0.f && t < fg_far[i]) s_cnt += 1.0;
    }
  }
  double a = block_sum(s_rgb, sh), b = block_sum(s_dep, sh), c = block_sum(s_cnt, sh);
  if (threadIdx.x == 0) {
    partial[4 * blockIdx.x] = a; partial[4 * blockIdx.x + 1] = b; partial[4 * blockIdx.x + 2] = c;
  }
}

__global__ void loss_final_kernel(const double* __restrict__ partial, int nblocks, int n, int S, int type, float lambda_depth,
                                  float* __restrict__ out) {
  if (threadIdx.x != 0) return;
  double a = 0.0, b = 0.0, c = 0.0;
  for (int i = 0; i < nblocks; ++i) { a += partial[4 * i]; b += partial[4 * i + 1]; c += partial[4 * i + 2]; }
  float rgb_loss = (float)(a / (double)(3LL * n));
  float dl = 0.f;
  if (type == NERFPP_DEPTH_MSE || type == NERFPP_DEPTH_L1) dl = (float)(b / c);   // 0/0 = NaN like torch.mean of empty
  else if (type == NERFPP_DEPTH_KL) dl = (float)(b / (double)S);
  out[0] = rgb_loss;
  out[1] = dl;
  out[2] = (type == NERFPP_DEPTH_NONE) ? rgb_loss : rgb_loss + lambda_depth * dl;   // ddp_train_nerf.py:482,493
  out[3] = (float)c;
}

// d(depth loss)/d(depth_pred) for mse/l1, d/d(fg_weights) for kl; scaled by the upstream scalar grad.
__global__ void depth_loss_backward_kernel(const float* __restrict__ depth, const float* __restrict__ depth_sup,
                                           const float* __restrict__ fg_w, const float* __restrict__ fg_z,
                                           const float* __restrict__ fg_dists, const float* __restrict__ fg_far, int n, int S,
                                           int type, float kl_sigma, const float* __restrict__ fwd_out,
                                           const float* __restrict__ grad_out, float* __restrict__ grad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const float g = grad_out[0];
  if (type == NERFPP_DEPTH_KL) {
    if (i >= (long long)n * S) return;
    int r = (int)(i / S);
    float t = depth_sup[r], v = 0.f;
    if (t > 0.f && t < fg_far[r]) {
      float dz = fg_z[i] - t;
      v = -expf(-(dz * dz) / (2.f * kl_sigma)) * fg_dists[i] / (fg_w[i] + 1e-5f) / (float)S * g;
    }
    grad[i] = v;
  } else {
    if (i >= n) return;
    float gt = depth_sup[i], v = 0.f;
    if (gt > 0.f) {
      float e = depth[i] - gt, cnt = fwd_out[3];
      v = (type == NERFPP_DEPTH_MSE ? 2.f * e : (e > 0.f ? 1.f : e < 0.f ? -1.f : 0.f)) / cnt * g;
    }
    grad[i] = v;
  }
}

// ---- N3: image metrics of the test loop (ddp_train_nerf.py:556-600) ------------------------------------------------
// partial [block][8] doubles: sum sq rgb err, valid count, sum (gt-pred)^2, sum (log gt - log pred)^2, sum |gt-pred|,
// sum |gt-pred|/gt, sum (gt-pred)^2/gt, unused
__global__ void __launch_bounds__(LOSS_THREADS)
metrics_partial_kernel(const float* __restrict__ rgb, const float* __restrict__ rgb_gt, const float* __restrict__ depth,
                       const float* __restrict__ depth_gt, long long n, float depth_scale, float cap, double* __restrict__ partial) {
  __shared__ double sh[LOSS_THREADS / 32];
  double a[7] = {0, 0, 0, 0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (rgb_gt)
      for (int c = 0; c < 3; ++c) { const float d = rgb_gt[3 * i + c] - rgb[3 * i + c]; a[0] += (double)(d * d); }
    if (depth_gt) {
      const float gt = depth_gt[i] / depth_scale, pr0 = depth[i] / depth_scale;        // :567-569
      if (gt < cap && gt > 1e-3f) {                                                     // :575
        const float g = fminf(fmaxf(gt, 1e-3f), cap), p = fminf(fmaxf(pr0, 1e-3f), cap);
        const float d = g - p, dl = logf(g) - logf(p);
        a[1] += 1.0; a[2] += (double)(d * d); a[3] += (double)(dl * dl); a[4] += (double)fabsf(d);
        a[5] += (double)(fabsf(d) / g); a[6] += (double)(d * d / g);
      }
    }
  }
  for (int k = 0; k < 7; ++k) {
    const double t = block_sum(a[k], sh);
    if (threadIdx.x == 0) partial[8 * blockIdx.x + k] = t;
  }
}
// out[8] = mse, psnr (utils.py:31), #valid depth pixels, rmse, rmse_log, abs_diff, abs_rel, sq_rel
__global__ void metrics_final_kernel(const double* __restrict__ partial, int nblocks, long long n, float* __restrict__ out) {
  if (threadIdx.x != 0) return;
  double a[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < nblocks; ++i)
    for (int k = 0; k < 7; ++k) a[k] += partial[8 * i + k];
  const double mse = a[0] / (3.0 * (double)n), c = a[1];
  out[0] = (float)mse;
  out[1] = (float)(-10.0 * log(mse + 1e-6) / log(10.0));
  out[2] = (float)c;
  out[3] = (float)sqrt(a[2] / c);
  out[4] = (float)sqrt(a[3] / c);
  out[5] = (float)(a[4] / c);
  out[6] = (float)(a[5] / c);
  out[7] = (float)(a[6] / c);
}

}  // namespace npp

using namespace npp;

extern "C" int nerfpp_depth_loss(const float* depth, const float* depth_sup, const float* fg_weights, const float* fg_z,
                                 const float* fg_dists, const float* fg_z_max, int n_rays, int s_fg, int depth_loss_type,
                                 float kl_sigma, float* out_loss, void* workspace, void* stream) {
  NPP_CHECK_ARG(depth_loss_type >= NERFPP_DEPTH_MSE && depth_loss_type <= NERFPP_DEPTH_KL, "unknown depth loss type");
  return nerfpp_loss(nullptr, nullptr, depth, depth_sup, fg_weights, fg_z, fg_dists, fg_z_max, n_rays, s_fg, depth_loss_type,
                     0.f, kl_sigma, out_loss, workspace, stream);
}

extern "C" int nerfpp_depth_loss_backward(const float* depth, const float* depth_sup, const float* fg_weights, const float* fg_z,
                                          const float* fg_dists, const float* fg_z_max, int n_rays, int s_fg,
                                          int depth_loss_type, float kl_sigma, const float* fwd_out, const float* grad_out,
                                          float* out_grad, void* stream) {
  NPP_CHECK_ARG(n_rays >= 0 && depth_sup && fwd_out && grad_out && out_grad, "bad argument");
  NPP_CHECK_ARG(depth_loss_type >= NERFPP_DEPTH_MSE && depth_loss_type <= NERFPP_DEPTH_KL, "unknown depth loss type");
  if (depth_loss_type == NERFPP_DEPTH_KL) NPP_CHECK_ARG(fg_weights && fg_z && fg_dists && fg_z_max && s_fg >= 1, "kl needs weights/z/dists/far");
  else NPP_CHECK_ARG(depth != nullptr, "mse/l1 need depth");
  long long tot = depth_loss_type == NERFPP_DEPTH_KL ? (long long)n_rays * s_fg : n_rays;
  if (tot == 0) return 0;
  depth_loss_backward_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      depth, depth_sup, fg_weights, fg_z, fg_dists, fg_z_max, n_rays, s_fg, depth_loss_type, kl_sigma, fwd_out, grad_out, out_grad);
  NPP_CHECK_LAUNCH();
  return 0;
}

extern "C" int64_t nerfpp_loss_workspace_bytes(void) { return (int64_t)LOSS_MAX_BLOCKS * 4 * sizeof(double); }

extern "C" int nerfpp_loss(const float* rgb, const float* rgb_gt, const float* depth, const float* depth_sup,
                           const float* fg_weights, const float* fg_z, const float* fg_dists, const float* fg_z_max,
                           int n_rays, int s_fg, int depth_loss_type, float lambda_depth, float kl_sigma, float* out_loss,
                           void* workspace, void* stream) {
  NPP_CHECK_ARG(n_rays >= 0 && (rgb == nullptr) == (rgb_gt == nullptr) && out_loss && workspace, "bad argument");
  NPP_CHECK_ARG(depth_loss_type >= NERFPP_DEPTH_NONE && depth_loss_type <= NERFPP_DEPTH_KL, "unknown depth loss type");
  if (depth_loss_type == NERFPP_DEPTH_MSE || depth_loss_type == NERFPP_DEPTH_L1)
    NPP_CHECK_ARG(depth && depth_sup, "mse/l1 need depth and depth_sup");
  if (depth_loss_type == NERFPP_DEPTH_KL)
    NPP_CHECK_ARG(depth_sup && fg_weights && fg_z && fg_dists && fg_z_max && s_fg >= 1, "kl needs weights/z/dists/far");
  long long work = depth_loss_type == NERFPP_DEPTH_KL ? (long long)n_rays * s_fg : (long long)n_rays * 3;
  int blocks = (int)((work + LOSS_THREADS * 4 - 1) / (LOSS_THREADS * 4));
  blocks = blocks < 1 ? 1 : blocks > LOSS_MAX_BLOCKS ? LOSS_MAX_BLOCKS : blocks;
  loss_partial_kernel<<<blocks, LOSS_THREADS, 0, (cudaStream_t)stream>>>(rgb, rgb_gt, depth, depth_sup, fg_weights, fg_z,
                                                                        fg_dists, fg_z_max, n_rays, s_fg, depth_loss_type,
                                                                        kl_sigma, (double*)workspace);
  NPP_CHECK_LAUNCH();
  loss_final_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const double*)workspace, blocks, n_rays, s_fg, depth_loss_type,
                                                       lambda_depth, out_loss);
  NPP_CHECK_LAUNCH();
  return 0;
}

extern "C" int nerfpp_image_metrics(const float* rgb, const float* rgb_gt, const float* depth, const float* depth_gt, int64_t n_pixels,
                                    float depth_scale, float cap, float* out_metrics, void* workspace, void* stream) {
  NPP_CHECK_ARG(out_metrics && workspace && n_pixels >= 1, "bad argument");
  NPP_CHECK_ARG(!rgb_gt || rgb, "rgb_gt without rgb");
  NPP_CHECK_ARG(!depth_gt || depth, "depth_gt without depth");
  int blocks = (int)((n_pixels + LOSS_THREADS - 1) / LOSS_THREADS);
  if (blocks > LOSS_MAX_BLOCKS) blocks = LOSS_MAX_BLOCKS;
  metrics_partial_kernel<<<blocks, LOSS_THREADS, 0, (cudaStream_t)stream>>>(rgb, rgb_gt, depth, depth_gt, n_pixels, depth_scale, cap,
                                                                           (double*)workspace);
  NPP_CHECK_LAUNCH();
  metrics_final_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const double*)workspace, blocks, n_pixels, out_metrics);
  NPP_CHECK_LAUNCH();
  return 0;
}
