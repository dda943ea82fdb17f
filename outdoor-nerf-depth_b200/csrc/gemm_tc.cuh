// Dense layer  out = act(A * W^T + bias)  on the tcgen05 tensor cores for the 1024-wide network of config 3
// (mipnerf360's NerfMLP / PropMLP, models.py:342-466), where one sample tile's activations (128 x 1024 fp16 = 256 KB)
// fit neither TMEM nor shared memory: activations live in global memory (row-major fp16, L2/HBM) and every layer is one
// persistent GEMM launch.  Operands arrive through tensor-map TMA (cp.async.bulk.tensor, SWIZZLE_128B boxes = the UMMA
// K-major operand layout), accumulators are double-buffered in TMEM so that a tile's epilogue runs under the next tile's
// MMAs, the output goes back through shared memory and a tensor-map TMA store.
#pragma once
#include <cuda.h>
#include "tc_common.cuh"

namespace npp {
namespace gemm {

constexpr int BM = 128;            // rows (samples) per tile = MMA M
constexpr int BK = 64;             // K per stage = one 128-byte swizzle row of fp16
constexpr int MAX_SEG = 6;
constexpr int THREADS = 320;       // warp 0: TMA producer, warp 1: MMA issuer, warps 2..9: epilogue

// One K-segment of a layer: n_chunks x 64 columns of A source `a_map` starting at column a_col against columns w_col.. of
// weight image `w_map`.  A plain layer has one segment; the skip layer two ([h | encoding]); the split-precision mode runs
// every segment three times (A_hi W_hi + A_lo W_hi + A_hi W_lo).
struct Segment { int a_map, a_col, w_map, w_col, n_chunks; };

struct GemmArgs {
  CUtensorMap a[4];                // A sources: dims {K_cols, M}, box {64, 128}
  CUtensorMap w[2];                // weights [N, K_pad] fp16 (hi, lo): box {64, BN}
  CUtensorMap out[2];              // outputs [M, N] fp16 (hi, lo): box {64, 128}
  Segment seg[MAX_SEG];
  int n_seg;
  int M, N;
  int relu;
  const float* bias;               // [N] fp32, added in the epilogue
};

template <int BN, bool PREC> struct Cfg {
  static constexpr int STAGES = PREC ? 3 : 4;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OUT_BUFS = PREC ? 4 : 2;
  static constexpr int OUT_BYTES = BM * 64 * 2;
  static constexpr int OFF_OUT = STAGES * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_OUT + OUT_BUFS * OUT_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;     // + alignment slack
};

int launch_gemm(const GemmArgs& g, int bn, bool prec, cudaStream_t st);
// row-major fp16 matrix [outer, inner_valid] with a row pitch of `pitch_elems`; box = {64, box_outer}
int make_map(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems, uint32_t box_outer);

}  // namespace gemm
}  // namespace npp
