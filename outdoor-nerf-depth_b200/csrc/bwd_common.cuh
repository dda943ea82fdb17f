// Tables and layouts shared by the backward kernels of the field MLP: the data-gradient chain (field_bwd_tc.cu), the
// weight-gradient GEMMs (wgrad_tc.cu) and the fused producer/consumer kernel that runs both at once (bwd_fused.cu).
#pragma once
#include "tc_common.cuh"

namespace npp {
namespace tcb {   // ---- data-gradient chain ---------------------------------------------------------------------------
using namespace npp::tc;

constexpr int NSTAGE = 4;
constexpr int STAGE_BYTES = 32768;          // one [256 N x 64 K] SW128 tile of W^T
constexpr int G_BYTES = 2 * CHUNK_BYTES;    // dG operand: 128 columns
constexpr int NUM_LAYERS = 9;               // t = 0: rgb.0 (remap part), 1: base_remap, 2..8: base 7..1

// forward layer whose weight the dgrad layer t multiplies with, and the forward activation its output is the gradient of
__host__ __device__ constexpr int weight_layer(int t) { return t == 0 ? L_RGB0 : t == 1 ? L_REMAP : 9 - t; }   // 2 -> base 7 ... 8 -> base 1
__host__ __device__ constexpr int target_act(int t) { return t == 0 ? 8 : 8 - t; }                             // ACT index: 8 = remap out, 7..0 = base outputs
// ... and back: the chain step whose output is the gradient of ACT index a (0..8)
__host__ __device__ constexpr int chain_step_of_act(int a) { return a == 8 ? 0 : 8 - a; }

struct Step { short t; short chunk; int blob_off; };
struct StepTable { Step s[40]; int n; int total; };
__host__ __device__ constexpr StepTable make_table() {
  StepTable tb{};
  int i = 0, off = 0;
  for (int t = 0; t < NUM_LAYERS; ++t)
    for (int c = 0; c < (t == 0 ? 2 : 4); ++c) { tb.s[i] = Step{(short)t, (short)c, off}; off += STAGE_BYTES; ++i; }
  tb.n = i; tb.total = off;
  return tb;
}

// fp32 tail of the packed buffer: rgb.2 weights [3][128], sigma-head weights [256]
constexpr int T_W2 = 0, T_WSIG = 3 * RGB_HID, T_TOTAL = T_WSIG + W;

}  // namespace tcb

namespace tcw {   // ---- weight-gradient GEMMs -------------------------------------------------------------------------
using namespace npp::tc;

// One GEMM: dW[:, col0 : col0+ncols] (+)= DZ[a_layer]^T * X
struct Job {
  int a_layer;        // DZ layer (rows of dW)
  int m_halves;       // 2 (256 outputs) or 1 (128)
  int x_is_e;         // X from the E tiles (1) or from ACT[x_layer] (0)
  int x_layer, x_chunk0, x_nchunks;
  int skip;           // leading columns of the first X chunk that map to no weight column (view dir starts at E col 96 = chunk 1 col 32)
  int ncols;          // weight columns written
  int w_index;        // NerfppNetGrads.w index
  int ld, col0;       // row stride (in-features of the layer) and first column in dW
  int bias;           // this job also sums DZ[a_layer] over the samples = the layer's bias gradient (one job per layer)
  int head;           // wgrad_tc_kernel's idle epilogue warps also take a small head's gradient from operands this job stages anyway:
                      // 1 = sigma head (weights = sum d_raw_sigma * h7: this job's X IS h7), 2 = rgb.2 (sum d_raw_rgb * rgb hidden: the
                      // 32 KB tile of ACT[9] rides along in the unused part of this job's X buffer)
};
constexpr int MAX_JOBS = 12;
struct JobTable { Job j[MAX_JOBS]; int n; };
__host__ __device__ constexpr JobTable make_jobs(bool bg) {
  JobTable t{};
  const int emb = emb_dim(bg), ech = bg ? 2 : 1;
  int i = 0;
  t.j[i++] = Job{0, 2, 1, 0, 0, ech, 0, emb, 0, emb, 0, 1};                               // base 0: embedding
  for (int l = 1; l < 8; ++l) {
    if (l == 5) t.j[i++] = Job{5, 2, 1, 0, 0, ech, 0, emb, 5, emb + W, 0, 0};             // base 5: [embedding | h4]
    t.j[i++] = Job{l, 2, 0, l - 1, 0, 4, 0, W, l, l == 5 ? emb + W : W, l == 5 ? emb : 0, 1};
  }
  t.j[i++] = Job{8, 2, 0, 7, 0, 4, 0, W, L_REMAP, W, 0, 1, 1};                            // base_remap (+ sigma head)
  t.j[i++] = Job{9, 1, 0, 8, 0, 4, 0, W, L_RGB0, W + VIEW_DIM, 0, 1};                     // rgb.0: remap part
  t.j[i++] = Job{9, 1, 1, 0, 1, 1, 32, VIEW_DIM, L_RGB0, W + VIEW_DIM, W, 0, 2};          // rgb.0: view-direction part (+ rgb.2 head)
  t.n = i;
  return t;
}

// MN-major SWIZZLE_128B operand: 64-feature blocks LBO = one chunk apart, 8-sample groups SBO = 1024 B apart
constexpr uint32_t MN_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t mn_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | ((uint32_t)(CHUNK_BYTES >> 4) << 16); }
__host__ __device__ constexpr uint32_t idesc_mn(int n) { return idesc_f16(n) | (1u << 15) | (1u << 16); }   // A and B MN-major

}  // namespace tcw
}  // namespace npp
