// nerfpp_backward: autograd of NerfNet.forward (ddp_model.py:74-147) w.r.t. the 48 parameter tensors, as CUDA kernels:
// composite backward (composite.cu) -> loss scale -> per net: dgrad chain (field_bwd_tc.cu) -> wgrad (wgrad_tc.cu).
#include <cstdio>
#include "bwd_common.cuh"

size_t npp_dgrad_packed_bytes();
int npp_pack_dgrad(const NerfppNetParams* p, bool bg, void* out, cudaStream_t st);
int npp_field_dgrad(const void* packed, const void* act, const void* mask, const float* rgb, const float* raw_sigma, const float* d_sigma,
                    const float* d_rgb, const float* scale, long long total, void* dz, float* d_raw_sigma, float* d_raw_rgb,
                    cudaStream_t st);
int npp_field_wgrad(bool bg, const void* act, const void* etiles, const void* dz, const float* d_raw_sigma, const float* d_raw_rgb,
                    const float* scale, long long total, const NerfppNetGrads* grads, void* ws, cudaStream_t st);
int npp_field_wgrad_heads(const void* act, const float* d_raw_sigma, const float* d_raw_rgb, const float* scale, long long total,
                          const NerfppNetGrads* grads, void* part_h, cudaStream_t st);
size_t npp_wgrad_ws_bytes(int dev);
size_t npp_tc_train_ws_bytes(long long n_samples);
// bwd_fused.cu: dgrad chain and weight-gradient GEMMs as one producer/consumer kernel (dZ never leaves L2)
size_t npp_bwd_fused_ws_bytes(int dev);
int npp_field_bwd_fused(bool bg, const void* packed, size_t packed_blob_bytes, const void* act, const void* etiles, const void* mask, const float* rgb,
                        const float* raw_sigma, const float* d_sigma, const float* d_rgb, const float* scale, long long total,
                        float* d_raw_sigma, float* d_raw_rgb, const NerfppNetGrads* grads, void* ws, cudaStream_t st);

int npp_field_dgrad_v2(const void* packed, size_t packed_blob_bytes, bool bg, const void* act, const void* mask, const float* rgb, const float* raw_sigma,
                       const float* d_sigma, const float* d_rgb, const float* scale, long long total, void* dz, float* d_raw_sigma,
                       float* d_raw_rgb, cudaStream_t st);

// How the field's backward runs (tests / diagnostics select through this hook; the product uses the default):
//   2 (default)  two kernels: the data-gradient chain with the storer-warp staging (bwd_fused.cu, producer role on every SM)
//                writes dZ to HBM, wgrad_tc_kernel reads it
//   1            the same with round 1's dgrad kernel (field_bwd_tc.cu)
//   0            ONE kernel: producers and weight-gradient consumers on disjoint SMs, dZ through an L2-resident slot ring.
//                Measured SLOWER than the two-kernel form on B200: a consumer SM can keep only ~192 KB
//                of operands in flight, which at L2/HBM latency feeds its tensor pipe at a third of the rate the MMAs need
static int g_bwd_mode = 2;
extern "C" void nerfpp_debug_set_bwd_mode(int mode) { g_bwd_mode = (mode >= 0 && mode <= 2) ? mode : 2; }

namespace npp {

// max |gradient entering the MLP| of one net, separately for its two entry points -- out[0]: the density head (d sigma),
// out[1]: the colour head (d rgb through the sigmoid) -- as bit patterns of non-negative floats (ordered like unsigned
// ints); out[2] != 0: a non-finite incoming gradient was seen.
__global__ void grad_absmax_kernel(const float* __restrict__ d_sigma, const float* __restrict__ d_rgb, const float* __restrict__ rgb,
                                   long long total, unsigned* __restrict__ out) {
  float ms = 0.f, mr = 0.f;
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float ds = d_sigma[i];
    bad |= !isfinite(ds);
    ms = fmaxf(ms, fabsf(ds));
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float cc = rgb[3 * i + c], dr = d_rgb[3 * i + c];
      bad |= !isfinite(dr);
      mr = fmaxf(mr, fabsf(dr * cc * (1.f - cc)));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { ms = fmaxf(ms, __shfl_xor_sync(0xffffffffu, ms, o)); mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, o)); }
  bad = __any_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
    if (ms > 0.f && isfinite(ms)) atomicMax(out, __float_as_uint(ms));
    if (mr > 0.f && isfinite(mr)) atomicMax(out + 1, __float_as_uint(mr));
    if (bad) atomicOr(out + 2, 1u);
  }
}
// Power-of-two loss scales (fp16 operands inside the backward kernels).  The gradient enters the MLP at two points whose
// magnitudes can differ by many orders -- a depth-prior loss acts on sigma only and can be 1e9 times the colour loss early
// in training -- and the layers above the point where the two paths join (rgb.2, rgb.0, base_remap) see the colour path
// alone.  So there are two scales per net:
//   scale[1] (colour path: d raw rgb, dG, d remap-out)  puts max |d raw rgb| at ~2^10,
//   scale[0] (everything below the join)                 puts max(|d raw sigma|, |d raw rgb|) at ~2^10 (<= scale[1]);
// the dgrad epilogue rescales by scale[0] / scale[1] where the density head joins the chain.  fp16 then keeps 6 binades of
// headroom above the largest entry (the epilogue saturates instead of overflowing) and ~34 below.  The exponent is
// clamped so that neither the scales nor their ratio leave fp32's range; a non-finite incoming gradient makes both
// scales NaN, so the parameter gradients come out NaN as they would from autograd (instead of silently finite).
__device__ __forceinline__ float pow2_scale(float m) {
  if (!(m > 0.f)) return 1.f;
  return exp2f(fminf(fmaxf(10.f - ceilf(log2f(m)), -60.f), 60.f));
}
__global__ void grad_scale_kernel(const unsigned* __restrict__ absmax, float* __restrict__ scale) {   // thread = net
  const unsigned* am = absmax + 4 * threadIdx.x;
  const float ms = __uint_as_float(am[0]), mr = __uint_as_float(am[1]);
  float sc = pow2_scale(fmaxf(ms, mr)), sr = pow2_scale(mr);
  if (am[2] != 0u) sc = sr = __int_as_float(0x7fc00000);
  scale[2 * threadIdx.x] = sc;
  scale[2 * threadIdx.x + 1] = sr;
}

}  // namespace npp

using namespace npp;

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
// d_raw_* / packed / dz exist once PER NET: the two nets' chains run concurrently on two streams
struct BwdWs { float *d_fg_sigma, *d_fg_rgb, *d_bg_sigma, *d_bg_rgb, *d_raw_sigma[2], *d_raw_rgb[2], *scale; unsigned* absmax; uint8_t *packed[2], *dz[2], *wg[2]; size_t bytes; };
static BwdWs carve_bwd(void* base, int n, int sf, int sb) {
  BwdWs w;
  char* p = (char*)base;
  size_t o = 0;
  const size_t tiles_net[2] = {((size_t)n * sf + tc::TILE - 1) / tc::TILE, ((size_t)n * sb + tc::TILE - 1) / tc::TILE};
  w.d_fg_sigma = (float*)(p + o); o += al256((size_t)n * sf * 4);
  w.d_fg_rgb = (float*)(p + o); o += al256((size_t)n * sf * 12);
  w.d_bg_sigma = (float*)(p + o); o += al256((size_t)n * sb * 4);
  w.d_bg_rgb = (float*)(p + o); o += al256((size_t)n * sb * 12);
  w.scale = (float*)(p + o); w.absmax = (unsigned*)(p + o + 32); o += 256;   // scale [2 nets][2], absmax [2 nets][4]
  int dev = 0;
  cudaGetDevice(&dev);
  for (int net = 0; net < 2; ++net) {
    w.d_raw_sigma[net] = (float*)(p + o); o += al256(tiles_net[net] * tc::TILE * 4);
    w.d_raw_rgb[net] = (float*)(p + o); o += al256(tiles_net[net] * tc::TILE * 12);
    w.packed[net] = (uint8_t*)(p + o); o += al256(npp_dgrad_packed_bytes());
    o = (o + 1023) & ~(size_t)1023;
    // fused: the producers' slot rings + counters; two-kernel form: every dZ of the net
    w.dz[net] = (uint8_t*)(p + o); o += g_bwd_mode == 0 ? npp_bwd_fused_ws_bytes(dev) : tc::act_bytes(tiles_net[net]);
    o = (o + 1023) & ~(size_t)1023;
    // per-CTA partials of the weight-gradient kernels (summed in a fixed order: the step is bit-reproducible)
    w.wg[net] = (uint8_t*)(p + o); o += npp_wgrad_ws_bytes(dev);
    o = (o + 1023) & ~(size_t)1023;
  }
  w.bytes = o;
  return w;
}

// layout of nerfpp_forward's per-sample workspace (capi.cu)
struct FwdWsView { const float *fg_sigma, *fg_rgb, *bg_sigma, *bg_rgb, *bg_dr; };
static FwdWsView view_fwd(const void* base, int n, int sf, int sb) {
  FwdWsView w;
  const char* p = (const char*)base;
  size_t o = 0;
  w.fg_sigma = (const float*)(p + o); o += al256((size_t)n * sf * 4);
  w.fg_rgb = (const float*)(p + o); o += al256((size_t)n * sf * 12);
  w.bg_sigma = (const float*)(p + o); o += al256((size_t)n * sb * 4);
  w.bg_rgb = (const float*)(p + o); o += al256((size_t)n * sb * 12);
  w.bg_dr = (const float*)(p + o);
  return w;
}

extern "C" int64_t nerfpp_backward_workspace_bytes(int n_rays, int s_fg, int s_bg) {
  if (n_rays < 0 || s_fg < 1 || s_bg < 1) return -1;
  return (int64_t)carve_bwd(nullptr, n_rays, s_fg, s_bg).bytes + 1024;
}

extern "C" int nerfpp_backward(const NerfppNetParams* params_fg, const NerfppNetParams* params_bg, const float* ray_d,
                               const float* fg_z_max, const float* fg_z, const float* bg_z, int n_rays, int s_fg, int s_bg,
                               const NerfppRenderOut* out, const NerfppRenderOut* grads, const void* workspace,
                               const void* train_workspace, const NerfppNetGrads* grads_fg, const NerfppNetGrads* grads_bg,
                               void* bwd_workspace, void* stream) {
  NPP_CHECK_ARG(params_fg && params_bg && ray_d && fg_z_max && fg_z && bg_z && out && grads && workspace && train_workspace &&
                grads_fg && grads_bg && bwd_workspace, "null argument");
  NPP_CHECK_ARG(n_rays >= 0 && s_fg >= 1 && s_bg >= 1, "bad shape");
  if (n_rays == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const FwdWsView f = view_fwd(workspace, n_rays, s_fg, s_bg);
  const BwdWs b = carve_bwd((void*)(((uintptr_t)bwd_workspace + 1023) & ~(uintptr_t)1023), n_rays, s_fg, s_bg);
  int rc = nerfpp_composite_backward(ray_d, fg_z_max, fg_z, bg_z, f.fg_sigma, f.fg_rgb, f.bg_sigma, f.bg_rgb, f.bg_dr, n_rays, s_fg,
                                     s_bg, out, grads, b.d_fg_sigma, b.d_fg_rgb, b.d_bg_sigma, b.d_bg_rgb, stream);
  if (rc) return rc;
  const long long tot_fg = (long long)n_rays * s_fg, tot_bg = (long long)n_rays * s_bg;
  cudaMemsetAsync(b.absmax, 0, 8 * sizeof(unsigned), st);
  grad_absmax_kernel<<<296, 256, 0, st>>>(b.d_fg_sigma, b.d_fg_rgb, f.fg_rgb, tot_fg, b.absmax);
  grad_absmax_kernel<<<296, 256, 0, st>>>(b.d_bg_sigma, b.d_bg_rgb, f.bg_rgb, tot_bg, b.absmax + 4);
  grad_scale_kernel<<<1, 2, 0, st>>>(b.absmax, b.scale);
  NPP_CHECK_LAUNCH();
  const size_t fg_train_bytes = (npp_tc_train_ws_bytes(tot_fg) + 1023) & ~(size_t)1023;
  // the foreground and the background chain are independent: the background's kernels go to the side stream
  cudaStream_t side = npp_fork(st);
  int rc_net[2] = {0, 0};
  for (int bg = 0; bg < 2; ++bg) {
    cudaStream_t s = bg ? side : st;
    const long long total = bg ? tot_bg : tot_fg;
    const size_t tiles = (size_t)((total + tc::TILE - 1) / tc::TILE);
    const uint8_t* tw = (const uint8_t*)train_workspace + (bg ? fg_train_bytes : 0);
    const uint8_t* act = tw;
    const uint8_t* etiles = tw + tc::train_ws_e_off(tiles);
    const float* raw_sigma = (const float*)(tw + tc::train_ws_sigma_off(tiles));
    const uint8_t* mask = tw + tc::train_ws_mask_off(tiles);
    const float* rgb = bg ? f.bg_rgb : f.fg_rgb;
    const float* d_sigma = bg ? b.d_bg_sigma : b.d_fg_sigma;
    const float* d_rgb = bg ? b.d_bg_rgb : b.d_fg_rgb;
    const NerfppNetGrads* grads_net = bg ? grads_bg : grads_fg;
    const float* scale = b.scale + 2 * bg;
    int& r = rc_net[bg];
    r = npp_pack_dgrad(bg ? params_bg : params_fg, bg != 0, b.packed[bg], s);
    if (r) continue;
    if (g_bwd_mode == 0) {
      r = npp_field_bwd_fused(bg != 0, b.packed[bg], (size_t)tcb::make_table().total, act, etiles, mask, rgb, raw_sigma, d_sigma, d_rgb, scale, total,
                              b.d_raw_sigma[bg], b.d_raw_rgb[bg], grads_net, b.dz[bg], s);
      if (!r) r = npp_field_wgrad_heads(act, b.d_raw_sigma[bg], b.d_raw_rgb[bg], scale, total, grads_net, b.wg[bg], s);
      continue;
    }
    if (g_bwd_mode == 2)
      r = npp_field_dgrad_v2(b.packed[bg], (size_t)tcb::make_table().total, bg != 0, act, mask, rgb, raw_sigma, d_sigma, d_rgb, scale, total, b.dz[bg],
                             b.d_raw_sigma[bg], b.d_raw_rgb[bg], s);
    else
      r = npp_field_dgrad(b.packed[bg], act, mask, rgb, raw_sigma, d_sigma, d_rgb, scale, total, b.dz[bg], b.d_raw_sigma[bg], b.d_raw_rgb[bg], s);
    if (!r) r = npp_field_wgrad(bg != 0, act, etiles, b.dz[bg], b.d_raw_sigma[bg], b.d_raw_rgb[bg], scale, total, grads_net, b.wg[bg], s);
  }
  npp_join(st, side);
  if (rc_net[0]) return rc_net[0];
  if (rc_net[1]) return rc_net[1];
  return 0;
}
