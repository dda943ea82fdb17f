// Transmittance composite of foreground and inverted-sphere background and their merge
// (the "raw2outputs" inlined in ddp_model.py:95-134).
// One warp per ray: samples are striped over lanes in chunks of 32, the exclusive
// product-scan of (1 - alpha + 1e-6) is a 5-step shuffle scan with a running carry between
// chunks, the weighted sums are shuffle reductions.  All loads/stores are lane-contiguous.
#include "common.cuh"

namespace npp {

constexpr int COMP_WARPS = 4;

struct CompAcc { float r, g, b, d; };

// Composites one side of one ray.  dist(i, z_i) supplies the interval length, val(i) the value
// whose expectation is the depth.  Returns the final transmittance (product over all samples).
template <class DistF, class ValF>
__device__ __forceinline__ float composite_side(const float* __restrict__ sigma, const float* __restrict__ rgb, int S, int lane,
                                                DistF dist, ValF val, float* __restrict__ out_w, float* __restrict__ out_dist,
                                                CompAcc& acc) {
  float carry = 1.f;
  acc = CompAcc{0.f, 0.f, 0.f, 0.f};
  for (int base = 0; base < S; base += 32) {
    int i = base + lane;
    bool ok = i < S;
    float dl = ok ? dist(i) : 0.f;
    float sg = ok ? sigma[i] : 0.f;
    float alpha = 1.f - expf(-sg * dl);              // ddp_model.py:99 / :121
    float x = ok ? (1.f - alpha + NPP_TINY) : 1.f;    // :100 / :124, the +1e-6 is inside the product
    float incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl *= t;
    }
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    float T = carry * excl;                            // exclusive transmittance (:102 / :125)
    float w = alpha * T;
    if (ok) {
      out_w[i] = w;
      if (out_dist) out_dist[i] = dl;
      acc.r += w * rgb[3 * i];
      acc.g += w * rgb[3 * i + 1];
      acc.b += w * rgb[3 * i + 2];
      acc.d += w * val(i);
    }
    carry *= __shfl_sync(0xffffffffu, incl, 31);
  }
  acc.r = warp_sum(acc.r); acc.g = warp_sum(acc.g); acc.b = warp_sum(acc.b); acc.d = warp_sum(acc.d);
  return carry;
}

__global__ void __launch_bounds__(COMP_WARPS * 32)
composite_kernel(const float* __restrict__ ray_d, const float* __restrict__ fg_z_max, const float* __restrict__ fg_z,
                 const float* __restrict__ bg_z, const float* __restrict__ fg_sigma, const float* __restrict__ fg_rgb,
                 const float* __restrict__ bg_sigma, const float* __restrict__ bg_rgb, const float* __restrict__ bg_depth_real,
                 int n, int Sf, int Sb, NerfppRenderOut out) {
  int lane = threadIdx.x & 31;
  int r = blockIdx.x * COMP_WARPS + (threadIdx.x >> 5);
  if (r >= n) return;
  float d0 = ray_d[3 * r], d1 = ray_d[3 * r + 1], d2 = ray_d[3 * r + 2];
  float dnorm = norm3(d0, d1, d2);
  float zmax = fg_z_max[r];
  const float* zf = fg_z + (size_t)r * Sf;
  const float* zb = bg_z + (size_t)r * Sb;

  CompAcc fa, ba;
  // foreground, ddp_model.py:95-105: dists scaled by ||ray_d||, last interval runs to fg_z_max
  float lam = composite_side(
      fg_sigma + (size_t)r * Sf, fg_rgb + (size_t)r * Sf * 3, Sf, lane,
      [&](int i) { return dnorm * ((i + 1 < Sf ? zf[i + 1] : zmax) - zf[i]); },
      [&](int i) { return zf[i]; }, out.fg_weights + (size_t)r * Sf, out.fg_dists + (size_t)r * Sf, fa);
  // background, ddp_model.py:116-128: flipped order j <-> bg_z[S-1-j], unscaled dists, last = 1e10
  const float* dr = bg_depth_real + (size_t)r * Sb;
  composite_side(
      bg_sigma + (size_t)r * Sb, bg_rgb + (size_t)r * Sb * 3, Sb, lane,
      [&](int j) { return (j + 1 < Sb) ? (zb[Sb - 1 - j] - zb[Sb - 2 - j]) : NPP_HUGE; },
      [&](int j) { return dr[j]; }, out.bg_weights + (size_t)r * Sb, (float*)nullptr, ba);
  if (lane == 0) {
    // merge, ddp_model.py:131-134
    float br = lam * ba.r, bg = lam * ba.g, bb = lam * ba.b, bd = lam * ba.d;
    out.fg_rgb[3 * r] = fa.r; out.fg_rgb[3 * r + 1] = fa.g; out.fg_rgb[3 * r + 2] = fa.b;
    out.bg_rgb[3 * r] = br; out.bg_rgb[3 * r + 1] = bg; out.bg_rgb[3 * r + 2] = bb;
    out.rgb[3 * r] = fa.r + br; out.rgb[3 * r + 1] = fa.g + bg; out.rgb[3 * r + 2] = fa.b + bb;
    out.fg_depth[r] = fa.d;
    out.bg_depth[r] = bd;
    out.bg_lambda[r] = lam;
    out.depth[r] = fa.d + bd;
  }
}

// ---- backward of the composite (autograd of ddp_model.py:95-134) --------------------------------
// One warp per ray.  With x_i = 1 - alpha_i + 1e-6, T_i = prod_{j<i} x_j, w_i = alpha_i T_i and an upstream
// gradient gw_i on every weight (its own + the rgb / depth sums it feeds):
//   dL/dalpha_i = gw_i T_i - (sum_{k>i} gw_k w_k + dlam * T_final) / x_i ,   dL/dsigma_i = dL/dalpha_i * delta_i * (1 - alpha_i)
// (dlam = gradient on the final transmittance: bg_lambda for the foreground, unused for the background).
struct SideGrad { float gc[3]; float gd; float dlam; };

template <class DistF, class ValF>
__device__ __forceinline__ void composite_side_backward(const float* __restrict__ sigma, const float* __restrict__ rgb, int S, int lane,
                                                        DistF dist, ValF val, const float* __restrict__ g_w, SideGrad g,
                                                        float* __restrict__ d_sigma, float* __restrict__ d_rgb,
                                                        float* sT, float* sA) {
  // forward recompute: exclusive transmittance and alpha of every sample into shared memory
  float carry = 1.f;
  for (int base = 0; base < S; base += 32) {
    int i = base + lane;
    bool ok = i < S;
    float dl = ok ? dist(i) : 0.f;
    float sg = ok ? sigma[i] : 0.f;
    float alpha = 1.f - expf(-sg * dl);
    float x = ok ? (1.f - alpha + NPP_TINY) : 1.f;
    float incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl *= t;
    }
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    if (ok) { sT[i] = carry * excl; sA[i] = alpha; }
    carry *= __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncwarp();
  // reverse pass: suffix sums of gw_k w_k
  float suffix = g.dlam * carry;                      // carry = T_final
  for (int base = ((S - 1) / 32) * 32; base >= 0; base -= 32) {
    int i = base + lane;
    bool ok = i < S;
    float T = ok ? sT[i] : 0.f, a = ok ? sA[i] : 0.f;
    float w = a * T;
    float gw = 0.f;
    if (ok) {
      gw = (g_w ? g_w[i] : 0.f) + g.gc[0] * rgb[3 * i] + g.gc[1] * rgb[3 * i + 1] + g.gc[2] * rgb[3 * i + 2] + g.gd * val(i);
      d_rgb[3 * i] = g.gc[0] * w; d_rgb[3 * i + 1] = g.gc[1] * w; d_rgb[3 * i + 2] = g.gc[2] * w;
    }
    float p = gw * w;
    float incl = p;                                   // inclusive suffix sum within the chunk (towards higher lanes)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float t = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < 32) incl += t;
    }
    float after = incl - p + suffix;                  // sum over k > i
    if (ok) {
      float x = 1.f - a + NPP_TINY;
      float dalpha = gw * T - after / x;
      d_sigma[i] = dalpha * dist(i) * (1.f - a);
    }
    suffix += __shfl_sync(0xffffffffu, incl, 0);
  }
}

__global__ void __launch_bounds__(COMP_WARPS * 32)
composite_backward_kernel(const float* __restrict__ ray_d, const float* __restrict__ fg_z_max, const float* __restrict__ fg_z,
                          const float* __restrict__ bg_z, const float* __restrict__ fg_sigma, const float* __restrict__ fg_rgb,
                          const float* __restrict__ bg_sigma, const float* __restrict__ bg_rgb, const float* __restrict__ bg_depth_real,
                          int n, int Sf, int Sb, NerfppRenderOut fwd, NerfppRenderOut g, float* __restrict__ d_fg_sigma,
                          float* __restrict__ d_fg_rgb, float* __restrict__ d_bg_sigma, float* __restrict__ d_bg_rgb, int smax) {
  extern __shared__ float s_buf[];
  int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  int r = blockIdx.x * COMP_WARPS + wl;
  if (r >= n) return;
  float* sT = s_buf + (size_t)wl * 2 * smax;
  float* sA = sT + smax;
  float d0 = ray_d[3 * r], d1 = ray_d[3 * r + 1], d2 = ray_d[3 * r + 2];
  float dnorm = norm3(d0, d1, d2);
  float zmax = fg_z_max[r];
  const float* zf = fg_z + (size_t)r * Sf;
  const float* zb = bg_z + (size_t)r * Sb;
  const float* dr = bg_depth_real + (size_t)r * Sb;
  auto G3 = [&](const float* p, int c) { return p ? p[3 * r + c] : 0.f; };
  auto G1 = [&](const float* p) { return p ? p[r] : 0.f; };
  float lam = fwd.bg_lambda[r];
  // bg_rgb / bg_depth as returned are already scaled by lambda (ddp_model.py:131-132)
  SideGrad gf, gb;
  float gs_c[3], gs_d = G1(g.depth) + G1(g.bg_depth);
  for (int c = 0; c < 3; ++c) { gf.gc[c] = G3(g.rgb, c) + G3(g.fg_rgb, c); gs_c[c] = G3(g.rgb, c) + G3(g.bg_rgb, c); gb.gc[c] = lam * gs_c[c]; }
  gf.gd = G1(g.depth) + G1(g.fg_depth);
  gb.gd = lam * gs_d;
  gb.dlam = 0.f;
  // d lambda = G_lambda + gs_c . bg_rgb_raw + gs_d * bg_depth_raw, with raw = scaled / lambda where lambda != 0;
  // recomputed from the samples below when lambda == 0 would need the raw sums, so accumulate them here
  float raw_c[3] = {0.f, 0.f, 0.f}, raw_d = 0.f;
  {
    float carry = 1.f;
    const float* sg = bg_sigma + (size_t)r * Sb;
    const float* cb = bg_rgb + (size_t)r * Sb * 3;
    for (int base = 0; base < Sb; base += 32) {
      int j = base + lane;
      bool ok = j < Sb;
      float dl = ok ? ((j + 1 < Sb) ? (zb[Sb - 1 - j] - zb[Sb - 2 - j]) : NPP_HUGE) : 0.f;
      float alpha = 1.f - expf(-(ok ? sg[j] : 0.f) * dl);
      float x = ok ? (1.f - alpha + NPP_TINY) : 1.f;
      float incl = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl *= t;
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.f;
      float w = alpha * carry * excl;
      if (ok) { raw_c[0] += w * cb[3 * j]; raw_c[1] += w * cb[3 * j + 1]; raw_c[2] += w * cb[3 * j + 2]; raw_d += w * dr[j]; }
      carry *= __shfl_sync(0xffffffffu, incl, 31);
    }
    raw_c[0] = warp_sum(raw_c[0]); raw_c[1] = warp_sum(raw_c[1]); raw_c[2] = warp_sum(raw_c[2]); raw_d = warp_sum(raw_d);
  }
  gf.dlam = G1(g.bg_lambda) + gs_c[0] * raw_c[0] + gs_c[1] * raw_c[1] + gs_c[2] * raw_c[2] + gs_d * raw_d;
  composite_side_backward(
      fg_sigma + (size_t)r * Sf, fg_rgb + (size_t)r * Sf * 3, Sf, lane,
      [&](int i) { return dnorm * ((i + 1 < Sf ? zf[i + 1] : zmax) - zf[i]); }, [&](int i) { return zf[i]; },
      g.fg_weights ? g.fg_weights + (size_t)r * Sf : nullptr, gf, d_fg_sigma + (size_t)r * Sf, d_fg_rgb + (size_t)r * Sf * 3, sT, sA);
  __syncwarp();
  composite_side_backward(
      bg_sigma + (size_t)r * Sb, bg_rgb + (size_t)r * Sb * 3, Sb, lane,
      [&](int j) { return (j + 1 < Sb) ? (zb[Sb - 1 - j] - zb[Sb - 2 - j]) : NPP_HUGE; }, [&](int j) { return dr[j]; },
      g.bg_weights ? g.bg_weights + (size_t)r * Sb : nullptr, gb, d_bg_sigma + (size_t)r * Sb, d_bg_rgb + (size_t)r * Sb * 3, sT, sA);
}

// depth2pts_outside (ddp_model.py:16-45) as a standalone op: one thread per (ray, depth) pair.
__global__ void depth2pts_kernel(const float* __restrict__ ray_o, const float* __restrict__ ray_d, const float* __restrict__ depth,
                                 long long n, float* __restrict__ pts, float* __restrict__ depth_real) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float o[3] = {ray_o[3 * i], ray_o[3 * i + 1], ray_o[3 * i + 2]};
  float d[3] = {ray_d[3 * i], ray_d[3 * i + 1], ray_d[3 * i + 2]};
  BgRay br = bg_ray_setup(o, d);
  float x[4];
  depth_real[i] = bg_point(br, depth[i], x);
  reinterpret_cast<float4*>(pts)[i] = make_float4(x[0], x[1], x[2], x[3]);
}

}  // namespace npp

using namespace npp;

extern "C" int nerfpp_depth2pts_outside(const float* ray_o, const float* ray_d, const float* depth, int64_t n, float* out_pts,
                                        float* out_depth_real, void* stream) {
  NPP_CHECK_ARG(n >= 0 && ray_o && ray_d && depth && out_pts && out_depth_real, "bad argument");
  if (n == 0) return 0;
  depth2pts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ray_o, ray_d, depth, n, out_pts, out_depth_real);
  NPP_CHECK_LAUNCH();
  return 0;
}

extern "C" int nerfpp_composite(const float* ray_d, const float* fg_z_max, const float* fg_z, const float* bg_z,
                                const float* fg_sigma, const float* fg_rgb, const float* bg_sigma, const float* bg_rgb,
                                const float* bg_depth_real, int n_rays, int s_fg, int s_bg, const NerfppRenderOut* out,
                                void* stream) {
  NPP_CHECK_ARG(n_rays >= 0 && s_fg >= 1 && s_bg >= 1 && out, "bad argument");
  NPP_CHECK_ARG(ray_d && fg_z_max && fg_z && bg_z && fg_sigma && fg_rgb && bg_sigma && bg_rgb && bg_depth_real, "null input");
  NPP_CHECK_ARG(out->rgb && out->fg_weights && out->bg_weights && out->fg_dists && out->fg_rgb && out->fg_depth &&
                out->bg_rgb && out->bg_depth && out->bg_lambda && out->depth, "null output");
  if (n_rays == 0) return 0;
  composite_kernel<<<(n_rays + COMP_WARPS - 1) / COMP_WARPS, COMP_WARPS * 32, 0, (cudaStream_t)stream>>>(
      ray_d, fg_z_max, fg_z, bg_z, fg_sigma, fg_rgb, bg_sigma, bg_rgb, bg_depth_real, n_rays, s_fg, s_bg, *out);
  NPP_CHECK_LAUNCH();
  return 0;
}

// Backward of nerfpp_composite: upstream gradients on the ten outputs (any pointer of `grads` may be NULL = zero;
// fg_dists carries no gradient) -> gradients on the per-sample sigma (after abs) and rgb (after sigmoid) of both nets,
// background in the same flipped order as the forward's arrays.  `fwd` needs bg_lambda only.
extern "C" int nerfpp_composite_backward(const float* ray_d, const float* fg_z_max, const float* fg_z, const float* bg_z,
                                         const float* fg_sigma, const float* fg_rgb, const float* bg_sigma, const float* bg_rgb,
                                         const float* bg_depth_real, int n_rays, int s_fg, int s_bg, const NerfppRenderOut* fwd,
                                         const NerfppRenderOut* grads, float* d_fg_sigma, float* d_fg_rgb, float* d_bg_sigma,
                                         float* d_bg_rgb, void* stream) {
  NPP_CHECK_ARG(n_rays >= 0 && s_fg >= 1 && s_bg >= 1 && fwd && grads, "bad argument");
  NPP_CHECK_ARG(ray_d && fg_z_max && fg_z && bg_z && fg_sigma && fg_rgb && bg_sigma && bg_rgb && bg_depth_real, "null input");
  NPP_CHECK_ARG(fwd->bg_lambda && d_fg_sigma && d_fg_rgb && d_bg_sigma && d_bg_rgb, "null output");
  if (n_rays == 0) return 0;
  const int smax = s_fg > s_bg ? s_fg : s_bg;
  const size_t smem = (size_t)COMP_WARPS * 2 * smax * sizeof(float);
  NPP_CHECK_ARG(smem <= 200 * 1024, "too many samples per ray");
  if (smem > 48 * 1024) cudaFuncSetAttribute(composite_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  composite_backward_kernel<<<(n_rays + COMP_WARPS - 1) / COMP_WARPS, COMP_WARPS * 32, smem, (cudaStream_t)stream>>>(
      ray_d, fg_z_max, fg_z, bg_z, fg_sigma, fg_rgb, bg_sigma, bg_rgb, bg_depth_real, n_rays, s_fg, s_bg, *fwd, *grads, d_fg_sigma,
      d_fg_rgb, d_bg_sigma, d_bg_rgb, smax);
  NPP_CHECK_LAUNCH();
  return 0;
}
