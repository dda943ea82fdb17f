// nerfpp_backward: autograd of NerfNet.forward (ddp_model.py:74-147) w.r.t. the 48 parameter tensors, as CUDA kernels:
// composite backward (composite.cu) -> loss scale -> per net: dgrad chain (field_bwd_tc.cu) -> wgrad (wgrad_tc.cu).
#include <cstdio>
#include "tc_common.cuh"

size_t npp_dgrad_packed_bytes();
int npp_pack_dgrad(const NerfppNetParams* p, bool bg, void* out, cudaStream_t st);
int npp_field_dgrad(const void* packed, const void* act, const void* mask, const float* rgb, const float* raw_sigma, const float* d_sigma,
                    const float* d_rgb, const float* scale, long long total, void* dz, float* d_raw_sigma, float* d_raw_rgb,
                    cudaStream_t st);
int npp_field_wgrad(bool bg, const void* act, const void* etiles, const void* dz, const float* d_raw_sigma, const float* d_raw_rgb,
                    const float* scale, long long total, const NerfppNetGrads* grads, cudaStream_t st);
size_t npp_tc_train_ws_bytes(long long n_samples);

namespace npp {

// max |gradient entering the MLP| over both nets (as the bit pattern of a non-negative float: ordered like unsigned ints)
__global__ void grad_absmax_kernel(const float* __restrict__ d_sigma, const float* __restrict__ d_rgb, const float* __restrict__ rgb,
                                   long long total, unsigned* __restrict__ out) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    m = fmaxf(m, fabsf(d_sigma[i]));
#pragma unroll
    for (int c = 0; c < 3; ++c) { const float cc = rgb[3 * i + c]; m = fmaxf(m, fabsf(d_rgb[3 * i + c] * cc * (1.f - cc))); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f && isfinite(m)) atomicMax(out, __float_as_uint(m));
}
// power-of-two loss scale that puts the largest incoming gradient at ~2^10: fp16 then keeps 6 binades of headroom above
// it (the dgrad epilogue saturates instead of overflowing) and 34 below -- the gradients shrink by ~2^10 on the way down
// the eight layers and must stay clear of fp16's subnormal range
__global__ void grad_scale_kernel(const unsigned* __restrict__ absmax, float* __restrict__ scale) {   // one scale per net
  const float m = __uint_as_float(absmax[threadIdx.x]);
  scale[threadIdx.x] = m > 0.f ? exp2f(10.f - ceilf(log2f(m))) : 1.f;
}

}  // namespace npp

using namespace npp;

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
struct BwdWs { float *d_fg_sigma, *d_fg_rgb, *d_bg_sigma, *d_bg_rgb, *d_raw_sigma, *d_raw_rgb, *scale; unsigned* absmax; uint8_t *packed, *dz; size_t bytes; };
static BwdWs carve_bwd(void* base, int n, int sf, int sb) {
  BwdWs w;
  char* p = (char*)base;
  size_t o = 0;
  const size_t smax = (size_t)n * (sf > sb ? sf : sb);
  const size_t tiles = (smax + tc::TILE - 1) / tc::TILE;
  w.d_fg_sigma = (float*)(p + o); o += al256((size_t)n * sf * 4);
  w.d_fg_rgb = (float*)(p + o); o += al256((size_t)n * sf * 12);
  w.d_bg_sigma = (float*)(p + o); o += al256((size_t)n * sb * 4);
  w.d_bg_rgb = (float*)(p + o); o += al256((size_t)n * sb * 12);
  w.d_raw_sigma = (float*)(p + o); o += al256(tiles * tc::TILE * 4);
  w.d_raw_rgb = (float*)(p + o); o += al256(tiles * tc::TILE * 12);
  w.scale = (float*)(p + o); w.absmax = (unsigned*)(p + o + 8); o += 256;   // [2] each: fg, bg
  w.packed = (uint8_t*)(p + o); o += al256(npp_dgrad_packed_bytes());
  o = (o + 1023) & ~(size_t)1023;
  w.dz = (uint8_t*)(p + o); o += tc::act_bytes(tiles);
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
  cudaMemsetAsync(b.absmax, 0, 2 * sizeof(unsigned), st);
  grad_absmax_kernel<<<296, 256, 0, st>>>(b.d_fg_sigma, b.d_fg_rgb, f.fg_rgb, tot_fg, b.absmax);
  grad_absmax_kernel<<<296, 256, 0, st>>>(b.d_bg_sigma, b.d_bg_rgb, f.bg_rgb, tot_bg, b.absmax + 1);
  grad_scale_kernel<<<1, 2, 0, st>>>(b.absmax, b.scale);
  NPP_CHECK_LAUNCH();
  const size_t fg_train_bytes = (npp_tc_train_ws_bytes(tot_fg) + 1023) & ~(size_t)1023;
  for (int bg = 0; bg < 2; ++bg) {
    const long long total = bg ? tot_bg : tot_fg;
    const size_t tiles = (size_t)((total + tc::TILE - 1) / tc::TILE);
    const uint8_t* tw = (const uint8_t*)train_workspace + (bg ? fg_train_bytes : 0);
    const uint8_t* act = tw;
    const uint8_t* etiles = tw + tc::train_ws_e_off(tiles);
    const float* raw_sigma = (const float*)(tw + tc::train_ws_sigma_off(tiles));
    rc = npp_pack_dgrad(bg ? params_bg : params_fg, bg != 0, b.packed, st);
    if (rc) return rc;
    rc = npp_field_dgrad(b.packed, act, tw + tc::train_ws_mask_off(tiles), bg ? f.bg_rgb : f.fg_rgb, raw_sigma, bg ? b.d_bg_sigma : b.d_fg_sigma,
                         bg ? b.d_bg_rgb : b.d_fg_rgb, b.scale + bg, total, b.dz, b.d_raw_sigma, b.d_raw_rgb, st);
    if (rc) return rc;
    rc = npp_field_wgrad(bg != 0, act, etiles, b.dz, b.d_raw_sigma, b.d_raw_rgb, b.scale + bg, total, bg ? grads_bg : grads_fg, st);
    if (rc) return rc;
  }
  return 0;
}
