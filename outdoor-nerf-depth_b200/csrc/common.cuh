// Shared device helpers and the packed-weight layouts of the NeRF++ hot path kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/nerfpp_b200.h"

#define NPP_TINY 1e-6f   // utils.py:8
#define NPP_HUGE 1e10f   // utils.py:7

namespace npp {

constexpr int W = NERFPP_WIDTH;       // 256
constexpr int NF_POS = NERFPP_NFREQ_POS;
constexpr int NF_VIEW = NERFPP_NFREQ_VIEW;
constexpr int VIEW_DIM = 3 * (1 + 2 * NF_VIEW);   // 27
constexpr int VIEW_PAD = 32;
constexpr int RGB_HID = W / 2;                    // 128

__host__ __device__ constexpr int pos_dim(bool bg) { return bg ? 4 : 3; }
__host__ __device__ constexpr int emb_dim(bool bg) { return pos_dim(bg) * (1 + 2 * NF_POS); }   // 63 / 84
// SIMT path pads the embedding to a multiple of 16 (one weight stage), TC path to 64 (one SW128 atom)
__host__ __device__ constexpr int emb_pad_simt(bool bg) { return bg ? 96 : 64; }
__host__ __device__ constexpr int emb_pad_tc(bool bg) { return bg ? 128 : 64; }

// layer ids in NerfppNetParams
enum { L_SIGMA = 8, L_REMAP = 9, L_RGB0 = 10, L_RGB2 = 11 };

// in-features of the nn.Linear of layer l as the state dict holds it (nerf_network.py:89-117)
__host__ __device__ constexpr int layer_in(int l, bool bg) {
  return l == 0 ? emb_dim(bg) : l == 5 ? emb_dim(bg) + W : l == L_RGB0 ? W + VIEW_DIM : l == L_RGB2 ? RGB_HID : W;
}
__host__ __device__ constexpr int layer_out(int l) {
  return l == L_SIGMA ? 1 : l == L_RGB0 ? RGB_HID : l == L_RGB2 ? 3 : W;
}

// ---- SIMT packed layout: fp32, transposed Wt[k][n], zero padded ------------------------------
struct SimtLayout {
  int w[NERFPP_NLAYERS];   // float offsets of Wt (sigma / rgb2 kept [out][in])
  int b[NERFPP_NLAYERS];
  int total;
};
__host__ __device__ constexpr int simt_rows(int l, bool bg) {   // padded K of layer l
  return l == 0 ? emb_pad_simt(bg) : l == 5 ? emb_pad_simt(bg) + W : l == L_RGB0 ? W + VIEW_PAD : l == L_RGB2 ? RGB_HID : W;
}
__host__ __device__ constexpr SimtLayout simt_layout(bool bg) {
  SimtLayout L{};
  int off = 0;
  for (int l = 0; l < NERFPP_NLAYERS; ++l) {
    L.w[l] = off;
    int n = (l == L_SIGMA) ? 1 : (l == L_RGB2) ? 3 : layer_out(l);
    off += simt_rows(l, bg) * n;
    off = (off + 3) & ~3;
  }
  for (int l = 0; l < NERFPP_NLAYERS; ++l) {
    L.b[l] = off;
    off += (layer_out(l) + 3) & ~3;
  }
  L.total = off;
  return L;
}

// ---- small device helpers -------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Per-ray constants of the inverted-sphere parametrisation (ddp_model.py:22-31): everything in
// depth2pts_outside that does not depend on the sample.
struct BgRay {
  float ps[3];      // p_sphere
  float ax[3];      // unit rotation axis
  float axps[3];    // cross(axis, p_sphere)
  float ax_dot_ps;  // sum(axis * p_sphere)
  float p_mid_norm, phi, d1, inv_len;
};
// torch.norm over 3 components as the CPU kernel rounds it: sqrt(fma(z,z,fma(y,y,x*x)))
__device__ __forceinline__ float norm3(float x, float y, float z) {
  return __fsqrt_rn(__fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x))));
}
__device__ __forceinline__ BgRay bg_ray_setup(const float o[3], const float d[3]) {
  BgRay r;
  float dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  float dO = d[0] * o[0] + d[1] * o[1] + d[2] * o[2];
  r.d1 = -dO / dd;
  float pm[3] = {o[0] + r.d1 * d[0], o[1] + r.d1 * d[1], o[2] + r.d1 * d[2]};
  r.p_mid_norm = norm3(pm[0], pm[1], pm[2]);
  r.inv_len = 1.f / norm3(d[0], d[1], d[2]);
  float d2 = sqrtf(1.f - r.p_mid_norm * r.p_mid_norm) * r.inv_len;
  float t = r.d1 + d2;
  r.ps[0] = o[0] + t * d[0]; r.ps[1] = o[1] + t * d[1]; r.ps[2] = o[2] + t * d[2];
  float ax[3] = {o[1] * r.ps[2] - o[2] * r.ps[1], o[2] * r.ps[0] - o[0] * r.ps[2], o[0] * r.ps[1] - o[1] * r.ps[0]};
  float an = norm3(ax[0], ax[1], ax[2]);
  r.ax[0] = ax[0] / an; r.ax[1] = ax[1] / an; r.ax[2] = ax[2] / an;
  r.axps[0] = r.ax[1] * r.ps[2] - r.ax[2] * r.ps[1];
  r.axps[1] = r.ax[2] * r.ps[0] - r.ax[0] * r.ps[2];
  r.axps[2] = r.ax[0] * r.ps[1] - r.ax[1] * r.ps[0];
  r.ax_dot_ps = r.ax[0] * r.ps[0] + r.ax[1] * r.ps[1] + r.ax[2] * r.ps[2];
  r.phi = asinf(r.p_mid_norm);
  return r;
}
// Per-sample part (ddp_model.py:32-44): pts = (x',y',z',1/r), returns depth_real.
__device__ __forceinline__ float bg_point(const BgRay& r, float depth, float pts[4]) {
  float theta = asinf(r.p_mid_norm * depth);
  float ang = r.phi - theta;
  float s, c;
  sincosf(ang, &s, &c);
  float q[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) q[i] = r.ps[i] * c + r.axps[i] * s + r.ax[i] * r.ax_dot_ps * (1.f - c);
  float qn = norm3(q[0], q[1], q[2]);
  pts[0] = q[0] / qn; pts[1] = q[1] / qn; pts[2] = q[2] / qn; pts[3] = depth;
  return 1.f / (depth + NPP_TINY) * cosf(theta) * r.inv_len + r.d1;
}

}  // namespace npp

// Fork / join of the caller's stream onto a library-owned side stream (one per device): the foreground and background
// nets of a NerfNet are independent until the composite, so their kernels are issued on two streams.  Each field kernel
// is a persistent one-CTA-per-SM grid, so they do not share SMs; what overlaps is the ragged last wave of one net with
// the first tiles of the other (and prologues with tails).  Works under stream capture: the side stream joins the
// capture through the fork event and must be joined back before the capture ends (npp_join does).
struct NppFork { cudaStream_t side; cudaEvent_t fork_ev, join_ev; };
NppFork* npp_fork_state();                       // capi.cu: lazily created per device; nullptr = overlap disabled
inline cudaStream_t npp_fork(cudaStream_t main) {
  NppFork* f = npp_fork_state();
  if (!f) return main;
  cudaEventRecord(f->fork_ev, main);
  cudaStreamWaitEvent(f->side, f->fork_ev, 0);
  return f->side;
}
inline void npp_join(cudaStream_t main, cudaStream_t side) {
  if (side == main) return;
  NppFork* f = npp_fork_state();
  cudaEventRecord(f->join_ev, side);
  cudaStreamWaitEvent(main, f->join_ev, 0);
}

// error plumbing shared by the .cu files
void npp_set_error(const char* fmt, ...);
#define NPP_CHECK_ARG(cond, msg) do { if (!(cond)) { npp_set_error("%s: %s", __func__, msg); return -1; } } while (0)
#define NPP_CHECK_LAUNCH() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
  npp_set_error("%s: %s", __func__, cudaGetErrorString(e_)); return (int)e_; } } while (0)
