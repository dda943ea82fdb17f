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
  CUtensorMap out[2];              // outputs [M, N] fp16 (hi, lo): box {64, 32} (one epilogue warp's rows)
  Segment seg[MAX_SEG];
  int n_seg;
  int M, N;
  int relu;
  const float* bias;               // [N] fp32, added in the epilogue
  // optional fused head (the density head behind the last trunk layer, the rgb head behind the view layer): head_part[row][p][k],
  // p = 2 * (n0 / BN) + hh, = the dot product of the row's activated fp32 outputs in the warp's column chunks with head_w[k] --
  // summed over p in a fixed order by a small kernel.  no_store: the layer's own output is not needed (nothing is written).
  const float* head_w;             // [head_n][N] fp32 or NULL
  float* head_part;                // [M][2 * N / BN][head_n]
  int head_n;                      // 1 or 3
  int no_store;
  int fused3;                      // split precision, F3 kernel: `seg` lists the hi x hi pass only; A lo map = a_map + 1, W lo map = 1
};

// CTAS = 2: a CTA pair (cluster of two SMs of one TPC, tcgen05 cta_group::2) computes a 256 x BN tile -- each CTA stages its
// own 128 rows of A and HALF of the weight tile, the leader issues M = 256 MMAs that read both halves, each CTA's TMEM
// receives its 128 rows of the result.  Per SM that halves the weight stream (the L2 -> shared-memory traffic that bounds
// the one-CTA form at 96 B/clk) and leaves room for six pipeline stages.
// F3 (split precision, CTA pairs): a stage holds the hi AND lo tiles of both operands for one K-chunk and feeds three MMA
// groups (A_hi W_hi, A_lo W_hi, A_hi W_lo): four tiles staged per three groups instead of six -- per 128 MMA-cycles the SM then
// moves 107 B/clk through shared memory instead of the 128 B/clk it has.  The epilogue stages its hi and lo images one after
// the other through one 4 KB buffer per warp to make room for three 64 KB stages.
template <int BN, bool PREC, int CTAS, bool F3 = false> struct Cfg {
  static constexpr int B_ROWS = BN / CTAS;                    // weight rows each CTA loads per stage
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = (F3 ? 2 : 1) * (A_BYTES + B_BYTES);
  static constexpr int OUT_WARP_BYTES = ((PREC && !F3) ? 2 : 1) * 4096;   // per epilogue warp: [32 rows x 64 cols] fp16 (hi [, lo])
  static constexpr int OUT_BYTES = 8 * OUT_WARP_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int BUDGET = 227 * 1024 - 1024 - BAR_BYTES - OUT_BYTES;
  static constexpr int STAGES = BUDGET / STAGE_BYTES > 8 ? 8 : BUDGET / STAGE_BYTES;
  static constexpr int OFF_OUT = STAGES * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_OUT + OUT_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + BAR_BYTES + 1024;     // + alignment slack
};

constexpr int OUT_BOX_ROWS = 32;
void plan(int N, int K, int* bn, int* ctas);      // tile plan for an output width: column-block width, CTAs per tile (the W map's box has bn / ctas rows)
int launch_gemm(const GemmArgs& g, int bn, int ctas, bool prec, cudaStream_t st);
void set_pair_mode(int mode);      // debug / A-B: 0 = one-CTA tiles only, 1 = CTA pairs where the shape allows (default)
// row-major fp16 matrix [outer, inner_valid] with a row pitch of `pitch_elems`; box = {64, box_outer}
int make_map(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems, uint32_t box_outer);

}  // namespace gemm
}  // namespace npp
