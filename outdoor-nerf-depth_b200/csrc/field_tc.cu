// Field evaluation on the 5th-generation tensor cores.
//
// One persistent CTA per SM walks 128-sample tiles (samples of all rays packed along the MMA M
// axis).  Per tile the ten dense layers (base 0..7, base_remap, rgb.0) run as tcgen05.mma
// (M=128, N=256|128, K=16, fp16 operands, fp32 accumulators in TMEM); sigma (256->1) and rgb.2
// (128->3) are fp32 dot products on the fp32 accumulators.
// (N=128 column halves would let the epilogue of one half hide under the MMAs of the other; measured, the
// overlap is cancelled by the slower epilogue/issue under a busy pipe -- DESIGN.md 4.1 -- so layers stay N=256.)
//
// The bias is part of the contraction: the E operand carries two constant-one columns and every
// layer has a K=16 MMA against a [N x 16] tile holding the bias split into fp16 hi + lo, so the
// epilogue is TMEM load -> cvt.rn.relu.f16x2 -> TMEM store and nothing else; the bias and embedding
// MMAs of layer l+1 do not depend on layer l's epilogue and run under its latency.
//
//   warp 0-7   epilogue: per layer TMEM -> regs, ReLU + fp16 pack -> written back to TMEM in place as
//              the A operand of the next layer, 64 columns at a time; sigma head after layer 7
//   warp 8     one lane issues every tcgen05.mma (A from TMEM or the E tile, B from shared memory);
//              layer l+1's K-chunk j starts as soon as the epilogue of layer l has produced columns
//              [64j,64j+64) (two TMEM accumulators ping-pong between consecutive layers)
//   warp 9     one lane streams the pre-swizzled weight tiles (in MMA issue order) from L2 into a
//              4 x 40 KB shared-memory ring with cp.async.bulk (+cluster multicast) and mbarriers
//   warp 10-13 positional encoding of the NEXT tile into the double-buffered E operand (fp16, SW128);
//              colour head (ReLU, rgb.2, sigmoid) of the tile that has just finished, from TMEM
//
// Shared memory (bytes): E 2x32K | ring 4x40K | barriers | fp32 head weights.  TMEM: 2 x 256 columns.
// Precision: operands are rounded to fp16 (11-bit significand), products/sums are fp32; measured
// against the fp32 reference: rgb/depth within 3e-5 relative (tests/test_parity_gpu.py).
#include <cstdlib>
#include "tc_common.cuh"

namespace npp {
namespace tc {

constexpr int NSTAGE = 4;
constexpr int AUX_BYTES = 8192;             // head of a ring slot: the layer's bias tile (first stage of a layer only)
constexpr int STAGE_BYTES = AUX_BYTES + 32768;   // + one [256 N x 64 K] SW128 weight tile
constexpr int BIAS_ROW_BYTES = 32;          // bias tile: [N x 16 K] fp16, unswizzled 8x8 core matrices
constexpr int E_BYTES = 2 * CHUNK_BYTES;    // 128 columns: [0,emb) position, [96,123) view dir, 123/124 = 1; double-buffered
constexpr int OFF_E = 0, OFF_W = 2 * E_BYTES, OFF_BAR = OFF_W + NSTAGE * STAGE_BYTES, OFF_TAIL = OFF_BAR + 256;
constexpr int VIEW_COL = 96, ONE_COL = 123;  // ONE_COL, ONE_COL+1 hold 1.0 (bias hi / lo)
constexpr int NUM_EPI_WARPS = 8, MMA_WARP = 8, LOAD_WARP = 9, EMB_WARP0 = 10, NUM_EMB_WARPS = 4, THREADS = 448;
constexpr int NUM_MMA_LAYERS = 10;          // base 0..7, remap, rgb0

// barrier slots
enum { B_WFULL = 0, B_WEMPTY = NSTAGE, B_AREADY = 2 * NSTAGE, B_EFULL = B_AREADY + 4, B_EEMPTY = B_EFULL + 2,
       B_ACC = B_EEMPTY + 2, B_ACCH = B_ACC + 2, B_RGBREADY = B_ACCH + 2, B_RGBFREE = B_RGBREADY + 1, B_COUNT = B_RGBFREE + 1 };
static_assert(8 * B_COUNT + 8 <= 256, "barrier area");

// fp32 tail of the packed buffer (float offsets)
constexpr int T_WSIG = 0;                    // 256
constexpr int T_WRGB2 = 256;                 // 3 x 128
constexpr int T_BSIG = 256 + 384;
constexpr int T_BRGB2 = T_BSIG + 1;
constexpr int T_TOTAL = T_BRGB2 + 3;
constexpr int SMEM_BYTES = OFF_TAIL + T_TOTAL * 4;
// The training instantiation runs the weight ring with 3 slots and stages its activation stores (tc_common.cuh:
// stage_store_chunk, 32 KB) in the fourth slot.
constexpr int NSTAGE_TRAIN = 3;
constexpr int OFF_STG_TRAIN = OFF_W + NSTAGE_TRAIN * STAGE_BYTES;
static_assert(STG_BYTES <= STAGE_BYTES && OFF_STG_TRAIN % 1024 == 0, "staging fits the spare ring slot");   // the fp32 head weights live in shared memory for the whole kernel
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");

__host__ __device__ constexpr int param_layer(int m) { return m < 8 ? m : m == 8 ? L_REMAP : L_RGB0; }

// One ring stage = one [N x 64] weight tile of one layer (+ the layer's bias tile ahead of its first stage).
// The loader and the packer walk this table; the MMA issuer hard-codes the same order.
enum { SRC_E = 0, SRC_A = 1 };
struct Step {
  short layer;     // MMA layer 0..9
  short src;       // SRC_*
  short chunk;     // 64-column chunk within the E / A region
  short n;         // rows of the tile = outputs of the layer (256 or 128)
  short bias;      // 1: [n x 16] bias tile at the head of the slot (blob: bias tile then weight tile)
  short lo;        // split-precision mode: 1 = the tile holds the LOW halves  W - fp16(W)  of the weights
  int col0;        // first input feature (state-dict column) of the tile; -2 = view dir + bias columns
  int blob_off, blob_bytes;
};
struct StepTable { Step s[96]; int n; int total; };

// prec: the split-precision ("3-pass") variant -- every weight tile is followed by the tile of its low halves, and the
// kernel accumulates  A_hi W_hi + A_lo W_hi + A_hi W_lo  (operands carried as hi + lo fp16 pairs, ~22 bits)
__host__ __device__ constexpr StepTable make_table(bool bg, bool prec = false) {
  StepTable t{};
  int i = 0, off = 0;
  for (int m = 0; m < NUM_MMA_LAYERS; ++m) {
    const short n = (m == 9) ? 128 : 256;
    const int abase = (m == 5) ? emb_dim(bg) : 0;
    short bias = (m != 9);
    auto push = [&](short src, short chunk, int col0) {
      const int bytes = n * 128 + (bias ? n * BIAS_ROW_BYTES : 0);
      t.s[i] = Step{(short)m, src, chunk, n, bias, 0, col0, off, bytes};
      off += bytes; ++i; bias = 0;
      if (prec) {
        t.s[i] = Step{(short)m, src, chunk, n, 0, 1, col0, off, n * 128};
        off += n * 128; ++i;
      }
    };
    if (m == 0 || m == 5) for (int c = 0; c < (bg ? 2 : 1); ++c) push(SRC_E, (short)c, 64 * c);
    if (m == 9) push(SRC_E, 1, -2);   // view-direction columns [96,123) + the bias columns of E -> rgb.0 inputs 256..282 + bias
    if (m != 0) for (int c = 0; c < 4; ++c) push(SRC_A, (short)c, abase + 64 * c);
  }
  t.n = i;
  t.total = off;
  return t;
}

__constant__ StepTable c_tab[4] = {make_table(false), make_table(true), make_table(false, true), make_table(true, true)};   // [2 prec + bg]
static const StepTable h_tab[4] = {make_table(false), make_table(true), make_table(false, true), make_table(true, true)};

// sin/cos of x * 2^k for the fp16 operand: two-constant Cody-Waite reduction to [-pi, pi] (exact to
// ~3e-7 for |arg| < 2^10), then the SFU approximations (abs error ~5e-7, far below the fp16 rounding
// of the operand, 2.4e-4).
__device__ __forceinline__ void fast_sincos(float arg, float* sn, float* cs) {
  const float n = rintf(arg * 0.15915494309189535f);
  float r = fmaf(-n, 6.2831854820251465f, arg);
  r = fmaf(-n, -1.7484555e-7f, r);
  *sn = __sinf(r);
  *cs = __cosf(r);
}

// Encodes one row (sample) of the E operand: columns [col_base, +dim(1+2 nfreq)) = Embedder(x)
// (nerf_network.py:42-60); fp16, 128B-swizzled.
// LO_OFF != 0 (split-precision mode): the low half  v - fp16(v)  of every value goes to the same place LO_OFF bytes on,
// and sin / cos come from the accurate library routine (the SFU approximations are good to ~5e-7 absolute, which the
// fp16 operand hides but a 22-bit operand does not).
template <int LO_OFF>
__device__ __forceinline__ void put_split(uint8_t* region, uint32_t off, float v) {
  const __half h = __float2half_rn(v);
  *reinterpret_cast<__half*>(region + off) = h;
  if (LO_OFF) *reinterpret_cast<__half*>(region + off + LO_OFF) = __float2half_rn(v - __half2float(h));
}
template <int LO_OFF>
__device__ __forceinline__ void embed_vec(const float* x, int dim, int nfreq, uint8_t* region, int row, int col_base) {
  for (int c = 0; c < dim; ++c) put_split<LO_OFF>(region, sw128_off(row, col_base + c), x[c]);
#pragma unroll 1
  for (int k = 0; k < nfreq; ++k) {
    const float f = (float)(1 << k);
    for (int c = 0; c < dim; ++c) {
      float sn, cs;
      if (LO_OFF) sincosf(x[c] * f, &sn, &cs);
      else fast_sincos(x[c] * f, &sn, &cs);
      const int col = col_base + dim + 2 * k * dim + c;
      put_split<LO_OFF>(region, sw128_off(row, col), sn);
      put_split<LO_OFF>(region, sw128_off(row, col + dim), cs);
    }
  }
}

// Epilogue of one layer for one thread (= one accumulator row, TMEM lane).  The layer kinds are compile-time variants so
// that the per-chunk chain  tcgen05.ld -> cvt -> tcgen05.st -> fence -> mbarrier arrive  carries no layer tests: the
// tensor pipe idles for exactly this chain at every layer boundary (a generic loop with the layer tests inside ran
// ~3x longer per chunk than the bare sequence).
//   KIND 0  hidden layer: ReLU + fp16 pack, written back over the fp32 columns just read = next layer's A operand
//   KIND 1  base layer 7: the same, then the sigma head (nerf_network.py:133) on the fp32 values AFTER the arrive
//   KIND 2  base_remap: fp16 pack without ReLU (nerf_network.py:135)
//   NSPLIT  the layer's MMAs run as two N = 128 column halves (see the issuer): the caller has waited for columns [0,128)
//           only; columns [128,256) are complete when `full_bar` reaches parity `full_par`, awaited here between chunks 1 and 2
template <int KIND, bool SAVE, bool PREC, bool NSPLIT = false>
__device__ __forceinline__ void epilogue_layer(uint32_t acc_addr, uint32_t aready_bar, int lane, int row, int hh,
                                               const float* __restrict__ tail, uint8_t* act_chunk0, uint4* mask_dst, uint8_t* stg,
                                               uint32_t& stg_flip, float& sig_part, long long* probe_slot, int xflags = 0,
                                               uint32_t full_bar = 0, uint32_t full_par = 0) {
  uint32_t v[2][32];
  uint32_t mbits[4];
  tmem_ld32(acc_addr, v[0]);
#pragma unroll
  for (int j = 0; j < 4; ++j) {       // 64-column chunks of the layer output
    uint32_t (&cur)[32] = v[j & 1];
    tmem_ld_wait(cur);
    if (j + 1 < 4 && !(NSPLIT && j == 1)) tmem_ld32(acc_addr + 64u * (j + 1), v[(j + 1) & 1]);   // overlaps the work below
    uint32_t pk[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) pk[t] = pack_f16x2<KIND != 2>(cur[2 * t], cur[2 * t + 1]);
    tmem_st16(acc_addr + 64u * j, pk);
    if (PREC) {   // the low halves  v - fp16(v)  go to the 16 columns behind the high halves (free: 32 fp32 columns were read)
      uint32_t pl[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        float v0 = __uint_as_float(cur[2 * t]), v1 = __uint_as_float(cur[2 * t + 1]);
        if (KIND != 2) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&pk[t]));
        pl[t] = pack_f16x2<false>(__float_as_uint(v0 - hf.x), __float_as_uint(v1 - hf.y));
      }
      tmem_st16(acc_addr + 64u * j + 16u, pl);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(aready_bar + 8u * j);
    if (probe_slot) probe_slot[j] = clock64();
    if (SAVE) {   // training, behind the arrive: the fp16 activations leave through shared memory and one bulk copy per warp
                  // pair (tc_common.cuh), and their ReLU mask is kept as bits for the dgrad chain
      if (!(xflags & 1024)) stage_store_chunk(stg, stg_flip, act_chunk0 + (size_t)j * CHUNK_BYTES, row >> 5, lane, hh, pk, (xflags & 2048) != 0);
      if (KIND != 2) mbits[j] = relu_mask_bits(pk);
    }
    if (KIND == 1) {
      float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float4 w4 = reinterpret_cast<const float4*>(tail + T_WSIG + 64 * j + 32 * hh)[t];
        s[0] = fmaf(fmaxf(__uint_as_float(cur[4 * t]), 0.f), w4.x, s[0]);
        s[1] = fmaf(fmaxf(__uint_as_float(cur[4 * t + 1]), 0.f), w4.y, s[1]);
        s[2] = fmaf(fmaxf(__uint_as_float(cur[4 * t + 2]), 0.f), w4.z, s[2]);
        s[3] = fmaf(fmaxf(__uint_as_float(cur[4 * t + 3]), 0.f), w4.w, s[3]);
      }
      sig_part += (s[0] + s[1]) + (s[2] + s[3]);
    }
    if (NSPLIT && j == 1) {      // the second column half of the accumulator
      mbar_wait(full_bar, full_par);
      tc_fence_after();
      tmem_ld32(acc_addr + 128u, v[0]);
    }
  }
  if (SAVE && KIND != 2) *mask_dst = make_uint4(mbits[0], mbits[1], mbits[2], mbits[3]);
}

template <bool SAVE, bool PREC, bool NSPLIT = false>
__device__ __forceinline__ void epilogue_dispatch(int m, uint32_t acc_addr, uint32_t aready_bar, int lane, int row, int hh,
                                                  const float* __restrict__ tail, uint8_t* act_chunk0, uint4* mask_dst, uint8_t* stg,
                                                  uint32_t& stg_flip, float& sig_part, long long* probe_slot, int xflags = 0,
                                                  uint32_t full_bar = 0, uint32_t full_par = 0) {
  if (m < 7) epilogue_layer<0, SAVE, PREC, NSPLIT>(acc_addr, aready_bar, lane, row, hh, tail, act_chunk0, mask_dst, stg, stg_flip, sig_part, probe_slot, xflags, full_bar, full_par);
  else if (m == 7) epilogue_layer<1, SAVE, PREC, NSPLIT>(acc_addr, aready_bar, lane, row, hh, tail, act_chunk0, mask_dst, stg, stg_flip, sig_part, probe_slot, xflags, full_bar, full_par);
  else epilogue_layer<2, SAVE, PREC, NSPLIT>(acc_addr, aready_bar, lane, row, hh, tail, act_chunk0, mask_dst, stg, stg_flip, sig_part, probe_slot, xflags, full_bar, full_par);
}

// Colour head for one row: rgb.2 (nerf_network.py:114-117) as fp32 dot products over the 128 rgb.0 accumulators
// (ReLU applied here), then the sigmoid (nerf_network.py:139).  Run by the embedding warps, which are idle most of a
// tile and -- being four consecutive warps -- reach all four TMEM lane quadrants; the epilogue warps go straight on to
// the next tile.  `free_bar` is arrived on once the last accumulator column is in registers: layer 1 of the next tile
// (the next writer of this TMEM buffer) waits for it.
template <bool SAVE>
__device__ __forceinline__ void rgb_head(uint32_t acc_addr, uint32_t free_bar, int lane, int row, const float* __restrict__ tail,
                                         uint8_t* act_chunk0, float* __restrict__ out_rgb3) {
  float acc[3] = {tail[T_BRGB2], tail[T_BRGB2 + 1], tail[T_BRGB2 + 2]};
  uint32_t v[2][32];
  tmem_ld32(acc_addr, v[0]);
#pragma unroll
  for (int p = 0; p < 4; ++p) {       // 32-column pieces
    uint32_t (&cur)[32] = v[p & 1];
    tmem_ld_wait(cur);
    if (p + 1 < 4) {
      tmem_ld32(acc_addr + 32u * (p + 1), v[(p + 1) & 1]);
    } else {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(free_bar);
    }
    if (SAVE) {     // training keeps relu(rgb.0) like every other layer's activations
      uint32_t pk[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) pk[t] = pack_f16x2<true>(cur[2 * t], cur[2 * t + 1]);
      store_act_chunk(act_chunk0 + (size_t)(p >> 1) * CHUNK_BYTES, row, p & 1, pk);
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float a0 = fmaxf(__uint_as_float(cur[4 * t]), 0.f), a1 = fmaxf(__uint_as_float(cur[4 * t + 1]), 0.f);
      const float a2 = fmaxf(__uint_as_float(cur[4 * t + 2]), 0.f), a3 = fmaxf(__uint_as_float(cur[4 * t + 3]), 0.f);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float4 w4 = reinterpret_cast<const float4*>(tail + T_WRGB2 + c * RGB_HID + 32 * p)[t];
        acc[c] = fmaf(a0, w4.x, acc[c]);
        acc[c] = fmaf(a1, w4.y, acc[c]);
        acc[c] = fmaf(a2, w4.z, acc[c]);
        acc[c] = fmaf(a3, w4.w, acc[c]);
      }
    }
  }
  if (out_rgb3) {
#pragma unroll
    for (int c = 0; c < 3; ++c) out_rgb3[c] = 1.f / (1.f + expf(-acc[c]));
  }
}

// CLUSTER > 1: the CTAs of a cluster walk their tiles in lock step and share every weight tile: each
// CTA fetches 1/CLUSTER of it from L2 and multicasts that slice into all CLUSTER rings.
// TRAIN: the training-mode forward (saves activations / masks / E tiles for the backward) is a separate instantiation so
// that the inference kernel's register allocation never sees the save code.
// PREC: split-precision ("3-pass") inference variant, see make_table().  The E operand is then single-buffered: its two
// 32 KB buffers hold the high and the low halves of one tile's encoding.
// NSPLIT (fast inference kernel only): hidden layers issue their A-chunk MMAs as N = 128 column halves in the order
// [h0: K0 K1] [h1: K0 K1] [h0: K2 K3] [h1: K2 K3].  The epilogue converts columns [0,128) of a layer -- the next layer's
// K-chunks 0 and 1 -- while that layer's last eight MMAs (h1: K2 K3) still run, and columns [128,256) while the NEXT
// layer's first sixteen MMAs (both halves of K0, K1) run: the tensor pipe no longer idles for the conversion chain at
// every layer boundary.  Each weight stage serves both halves (rows 0-127, then 128-255) before it is released.
template <bool BG, int CLUSTER, bool TRAIN, bool PREC = false, bool NSPLIT = false>
__global__ void __launch_bounds__(THREADS, 1)
field_tc_kernel(const uint8_t* __restrict__ blobs, const float* __restrict__ tail, const float* __restrict__ ray_o,
                const float* __restrict__ ray_d, const float* __restrict__ z, int n, int S, float* __restrict__ out_sigma,
                float* __restrict__ out_rgb, float* __restrict__ out_depth_real, int num_tiles, TrainSave save, long long* __restrict__ dbg, int flags) {
  extern __shared__ __align__(1024) uint8_t smem[];
  static_assert(!PREC || (!TRAIN && CLUSTER == 1), "the split-precision variant is inference-only");
  static_assert(!NSPLIT || (!TRAIN && !PREC && CLUSTER == 1), "column-half issue order: fast inference kernel only");
  constexpr int D = BG ? 4 : 3;
  constexpr uint32_t NEB = PREC ? 1 : 2;       // E buffers in rotation
  const StepTable& tab = c_tab[(PREC ? 2 : 0) + (BG ? 1 : 0)];
  // read the thread coordinates once through volatile asm: otherwise ptxas re-reads %tid / %ctaid with S2R (tens of cycles
  // each) inside the epilogue's per-chunk loop, on the critical path of every mbarrier arrive
  uint32_t tid_pinned, cta_pinned;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_pinned));
  asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(cta_pinned));
  const int warp = (int)(tid_pinned >> 5), lane = (int)(tid_pinned & 31);
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar0 = s_base + OFF_BAR;
  auto bar = [&](int i) { return bar0 + 8u * i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * B_COUNT);
  if ((s_base & 1023u) != 0) __trap();
  const uint32_t cta_rank = CLUSTER > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t kMask = (uint16_t)((1u << CLUSTER) - 1u);
  // tiles: cluster c takes groups of CLUSTER consecutive tiles; every CTA of a cluster runs the same
  // number of (possibly empty) tiles so the shared weight ring stays in step
  const int n_groups = (num_tiles + CLUSTER - 1) / CLUSTER;
  const int group0 = (int)cta_pinned / CLUSTER, group_step = gridDim.x / CLUSTER;
  const bool probing = dbg != nullptr;                 // single-shot timestamps of one layer boundary (tile 5, layers 2-3)
  const bool timing = probing && !(flags & 128);      // + per-role cycle accumulators (perturbs the epilogue by ~2x)

  if (warp == MMA_WARP && lane == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar(B_WFULL + i), 1); mbar_init(bar(B_WEMPTY + i), CLUSTER); }
    for (int i = 0; i < 4; ++i) mbar_init(bar(B_AREADY + i), NUM_EPI_WARPS);
    for (int i = 0; i < 2; ++i) { mbar_init(bar(B_EFULL + i), NUM_EMB_WARPS); mbar_init(bar(B_EEMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(B_ACC + i), 1); mbar_init(bar(B_ACCH + i), 1); }
    mbar_init(bar(B_RGBREADY), 1);
    mbar_init(bar(B_RGBFREE), NUM_EMB_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == LOAD_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void*)tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp < NUM_EPI_WARPS) {   // zero both E buffers once (padding columns are never written again), then the two ones
    for (int i = (int)tid_pinned; i < 2 * E_BYTES / 16; i += NUM_EPI_WARPS * 32) reinterpret_cast<uint4*>(smem + OFF_E)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int eb = (int)tid_pinned >> 7, row = (int)tid_pinned & 127;
    if (!PREC || eb == 0) {        // (split precision: buffer 1 holds the low halves, whose constant columns stay 0)
      *reinterpret_cast<__half*>(smem + OFF_E + eb * E_BYTES + sw128_off(row, ONE_COL)) = __float2half_rn(1.f);
      *reinterpret_cast<__half*>(smem + OFF_E + eb * E_BYTES + sw128_off(row, ONE_COL + 1)) = __float2half_rn(1.f);
    }
    fence_proxy_async();
  }
  float* const tail_s = reinterpret_cast<float*>(smem + OFF_TAIL);
  for (int i = (int)tid_pinned; i < T_TOTAL; i += THREADS) tail_s[i] = tail[i];
  tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();   // peers' barriers are initialised before anything remote targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long total = (long long)n * S;

  if (warp == LOAD_WARP) {
    // ================= weight loader =================
    if (lane == 0) {
      uint32_t it = 0;
      for (int grp = group0; grp < n_groups; grp += group_step) {
        for (int i = 0; i < tab.n; ++i, ++it) {
          constexpr uint32_t NS = TRAIN ? NSTAGE_TRAIN : NSTAGE;
          const uint32_t st = it % NS, ph = (it / NS) & 1;
          const uint32_t bytes = (uint32_t)tab.s[i].blob_bytes;
          // a stage without bias tile lands straight on the weight area of the slot
          const uint32_t dst = s_base + OFF_W + st * STAGE_BYTES + (tab.s[i].bias ? (uint32_t)(AUX_BYTES - tab.s[i].n * BIAS_ROW_BYTES) : (uint32_t)AUX_BYTES);
          mbar_wait(bar(B_WEMPTY + st), ph ^ 1);          // every CTA of the cluster has consumed this stage
          if ((flags & 8) && it >= NS) { mbar_arrive(bar(B_WFULL + st)); continue; }   // timing experiment: no refill (wrong results)
          mbar_expect_tx(bar(B_WFULL + st), bytes);
          if (CLUSTER == 1) {
            bulk_g2s(dst, blobs + tab.s[i].blob_off, bytes, bar(B_WFULL + st));
          } else {
            const uint32_t part = bytes / CLUSTER, o = cta_rank * part;
            bulk_g2s_mcast(dst + o, blobs + tab.s[i].blob_off + o, part, bar(B_WFULL + st), kMask);
          }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ================= MMA issuer =================
    // The whole warp runs the control flow (waits, ring bookkeeping) converged; one elected lane issues.  Whatever
    // this warp executes between two MMAs is potential tensor-pipe idle time (a single warp cannot hide its own
    // instruction latencies), so the ring position is kept as running registers and descriptors as 32-bit words
    // that advance by constants.
    {
      // (flags & 64: timing experiment -- N=16 MMAs leave the tensor pipe idle, exposing the issue thread's own pace)
      const uint32_t ID256 = (flags & 64) ? idesc_f16(16) : idesc_f16(256), ID128 = (flags & 64) ? idesc_f16(16) : idesc_f16(128);
      constexpr int E_CHUNKS = BG ? 2 : 1;
      const uint32_t ring0 = s_base + OFF_W, wfull0 = bar(B_WFULL), wempty0 = bar(B_WEMPTY);
      uint32_t st = 0, ph = 0, slot = ring0, wfull = wfull0, wempty = wempty0;   // ring position
      uint32_t a_par = 0, tile_i = 0;
      const long long t0 = clock64();
      auto wait_stage = [&]() { mbar_wait(wfull, ph); tc_fence_after(); };
      auto release = [&](uint32_t wempty_bar) { if (CLUSTER == 1) tc_commit(wempty_bar); else tc_commit_mcast(wempty_bar, kMask); };
      auto advance = [&]() {
        ++st; slot += STAGE_BYTES; wfull += 8; wempty += 8;
        if (st == (TRAIN ? NSTAGE_TRAIN : NSTAGE)) { st = 0; ph ^= 1; slot = ring0; wfull = wfull0; wempty = wempty0; }
      };
      auto ts4 = [&](uint32_t d, uint32_t a0, uint32_t blo, uint32_t idesc) {   // one 64-column A chunk from TMEM
        mma_ts<1>(d, a0, blo, idesc);
        mma_ts<1>(d, a0 + 8u, blo + 2u, idesc);
        mma_ts<1>(d, a0 + 32u, blo + 4u, idesc);
        mma_ts<1>(d, a0 + 40u, blo + 6u, idesc);
      };
      for (int grp = group0; grp < n_groups; grp += group_step, ++tile_i) {
        const uint32_t eb = tile_i % NEB;
        const uint32_t e_addr = s_base + OFF_E + eb * E_BYTES;
        const uint32_t e_lo_addr = e_addr + E_BYTES;                           // PREC: the low halves of the encoding
        const uint32_t one_lo = sw128_lo(e_addr + CHUNK_BYTES + 3 * 32);   // E columns [112,128): the constant-one columns
        mbar_wait(bar(B_EFULL + eb), (tile_i / NEB) & 1);
        if (TRAIN && lane == 0) bulk_s2g(save.e + (size_t)(grp * CLUSTER + (int)cta_rank) * E_BYTES, e_addr, E_BYTES);   // training: keep the E operand
        // layer 9 of the previous tile reads its A operand from accumulator buffer 0, which layer 0 is about to
        // overwrite: wait until those MMAs have completed (that ACC barrier completes 5 times per tile; layer 9 is the 5th)
        if (tile_i > 0) mbar_wait(bar(B_ACC + 1), (tile_i - 1) & 1);
        tc_fence_after();
#pragma unroll 1
          for (int m = 0; m < NUM_MMA_LAYERS; ++m) {
            const uint32_t d_tmem = tmem_base + (uint32_t)(m & 1) * 256u;
            const uint32_t a_tmem = tmem_base + (uint32_t)((m - 1) & 1) * 256u;
            const uint32_t idesc = (m == 9) ? ID128 : ID256;
            if (m == 0 || m == 5) {      // embedding chunks (+ the bias tile with the first)
#pragma unroll
              for (int c = 0; c < E_CHUNKS; ++c) {
                wait_stage();
                if (elect_one()) {
                  if (c == 0) mma_ss<0>(d_tmem, one_lo, SW128_HI, bias_lo(slot, 256), NOSW_HI, idesc);
                  const uint32_t alo = sw128_lo(e_addr + c * CHUNK_BYTES), blo = sw128_lo(slot + AUX_BYTES);
#pragma unroll
                  for (int k = 0; k < (c == 1 ? 2 : 4); ++k) mma_ss<1>(d_tmem, alo + 2u * k, SW128_HI, blo + 2u * k, SW128_HI, idesc);
                  if (PREC) {            // + E_lo W_hi
                    const uint32_t allo = sw128_lo(e_lo_addr + c * CHUNK_BYTES);
#pragma unroll
                    for (int k = 0; k < (c == 1 ? 2 : 4); ++k) mma_ss<1>(d_tmem, allo + 2u * k, SW128_HI, blo + 2u * k, SW128_HI, idesc);
                  }
                  release(wempty);
                  if (!PREC && m == 0 && c == E_CHUNKS - 1) { if (NSPLIT) tc_commit(bar(B_ACCH)); tc_commit(bar(B_ACC)); }
                }
                __syncwarp();
                advance();
                if (PREC) {              // + E_hi W_lo (the next ring stage)
                  wait_stage();
                  if (elect_one()) {
                    const uint32_t alo = sw128_lo(e_addr + c * CHUNK_BYTES), blo = sw128_lo(slot + AUX_BYTES);
#pragma unroll
                    for (int k = 0; k < (c == 1 ? 2 : 4); ++k) mma_ss<1>(d_tmem, alo + 2u * k, SW128_HI, blo + 2u * k, SW128_HI, idesc);
                    release(wempty);
                    if (m == 0 && c == E_CHUNKS - 1) tc_commit(bar(B_ACC));
                  }
                  __syncwarp();
                  advance();
                }
              }
            }
            if (m == 9) {                // view-direction columns + bias columns of E against rgb.0's view/bias tile
              if (TRAIN && lane == 0) bulk_s2g_wait_read();
              __syncwarp();
              wait_stage();
              if (elect_one()) {
                const uint32_t alo = sw128_lo(e_addr + CHUNK_BYTES), blo = sw128_lo(slot + AUX_BYTES);
                mma_ss<0>(d_tmem, alo + 4u, SW128_HI, blo + 4u, SW128_HI, idesc);
                mma_ss<1>(d_tmem, alo + 6u, SW128_HI, blo + 6u, SW128_HI, idesc);
                if (PREC) {
                  const uint32_t allo = sw128_lo(e_lo_addr + CHUNK_BYTES);
                  mma_ss<1>(d_tmem, allo + 4u, SW128_HI, blo + 4u, SW128_HI, idesc);
                  mma_ss<1>(d_tmem, allo + 6u, SW128_HI, blo + 6u, SW128_HI, idesc);
                }
                release(wempty);
                if (!PREC) tc_commit(bar(B_EEMPTY + eb));      // last reader of this tile's E buffer (the bulk store below was waited for)
              }
              __syncwarp();
              advance();
              if (PREC) {
                wait_stage();
                if (elect_one()) {
                  const uint32_t alo = sw128_lo(e_addr + CHUNK_BYTES), blo = sw128_lo(slot + AUX_BYTES);
                  mma_ss<1>(d_tmem, alo + 4u, SW128_HI, blo + 4u, SW128_HI, idesc);
                  mma_ss<1>(d_tmem, alo + 6u, SW128_HI, blo + 6u, SW128_HI, idesc);
                  release(wempty);
                  tc_commit(bar(B_EEMPTY + eb));
                }
                __syncwarp();
                advance();
              }
            }
            if (NSPLIT && m != 0 && m != 9) {
              // ---- column-half issue order (see the kernel's header comment) ----
              wait_stage();                                   // chunk 0's stage; its head holds the layer's bias tile (not layer 5's)
              if (m == 1 && tile_i > 0) { mbar_wait(bar(B_RGBFREE), (tile_i - 1) & 1); tc_fence_after(); }   // previous tile's rgb.0 accumulators have been read
              if (m != 5) {
                // the bias MMA (N = 256, overwrites the whole accumulator) goes out right behind the previous layer's last MMA:
                // the tensor pipe executes in issue order, so it cannot overtake that layer's reads of this buffer
                if (elect_one()) mma_ss<0>(d_tmem, one_lo, SW128_HI, bias_lo(slot, 256), NOSW_HI, ID256);
                __syncwarp();
              }
              uint32_t hslot[2], hwempty[2];
#pragma unroll
              for (int half = 0; half < 2; ++half) {          // K-chunk pairs (0, 1) and (2, 3)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                  const int c = 2 * half + cc;
                  mbar_wait2(bar(B_AREADY + c), a_par, wfull, ph);
                  tc_fence_after();
                  if (elect_one()) ts4(d_tmem, a_tmem + 64u * c, sw128_lo(slot + AUX_BYTES), ID128);              // columns [0,128)
                  __syncwarp();
                  hslot[cc] = slot; hwempty[cc] = wempty;
                  advance();
                }
                if (elect_one()) {
                  if (half == 1) tc_commit(bar(B_ACCH + (m & 1)));                                                  // columns [0,128) complete
                  ts4(d_tmem + 128u, a_tmem + 64u * (2 * half), sw128_lo(hslot[0] + AUX_BYTES + 16384u), ID128);   // columns [128,256): weight rows 128..255
                  release(hwempty[0]);
                  ts4(d_tmem + 128u, a_tmem + 64u * (2 * half + 1), sw128_lo(hslot[1] + AUX_BYTES + 16384u), ID128);
                  release(hwempty[1]);
                  if (half == 1) tc_commit(bar(B_ACC + (m & 1)));
                }
                __syncwarp();
              }
              a_par ^= 1;
            } else if (m != 0) {
              const bool early_bias = (m != 1 && m != 5 && m != 9);
              if (early_bias) {
                // The bias MMA does not depend on the epilogue, so it runs under the epilogue's latency chain -- once the
                // previous layer has completed (it overwrites the accumulator buffer that layer's MMAs read their A operand
                // from).  Not for layer 1: its buffer still holds the previous tile's rgb.0 accumulators until the colour
                // head has read them (B_RGBFREE); there the bias goes with chunk 0.
                mbar_wait2(bar(B_ACC + ((m - 1) & 1)), (tile_i + (uint32_t)((m - 1) >> 1)) & 1u, wfull, ph);
                tc_fence_after();
                if (elect_one()) mma_ss<0>(d_tmem, one_lo, SW128_HI, bias_lo(slot, 256), NOSW_HI, idesc);
                __syncwarp();
              }
              if (m == 1 && tile_i > 0) mbar_wait(bar(B_RGBFREE), (tile_i - 1) & 1);   // previous tile's rgb.0 accumulators have been read
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                mbar_wait2(bar(B_AREADY + c), a_par, wfull, ph);
                tc_fence_after();
                if (probing && lane == 0 && tile_i == 5 && m == 3) dbg[14 * 148 + 16 * (int)cta_pinned + 4 + c] = clock64();
                if (elect_one()) {
                  if (c == 0 && m == 1) mma_ss<0>(d_tmem, one_lo, SW128_HI, bias_lo(slot, 256), NOSW_HI, idesc);
                  ts4(d_tmem, a_tmem + 64u * c, sw128_lo(slot + AUX_BYTES), idesc);
                  if (PREC) ts4(d_tmem, a_tmem + 64u * c + 16u, sw128_lo(slot + AUX_BYTES), idesc);     // + A_lo W_hi
                  release(wempty);
                  if (!PREC && c == 3) {
                    if (NSPLIT) tc_commit(bar(B_ACCH + (m & 1)));   // (only layer 9 comes here then: keeps the two barriers' phases in step)
                    tc_commit(bar(B_ACC + (m & 1)));
                    if (m == 9) tc_commit(bar(B_RGBREADY));   // the colour head (embedding warps) has its own barrier: one phase per tile
                  }
                }
                __syncwarp();
                advance();
                if (PREC) {              // + A_hi W_lo (the next ring stage)
                  wait_stage();
                  if (elect_one()) {
                    ts4(d_tmem, a_tmem + 64u * c, sw128_lo(slot + AUX_BYTES), idesc);
                    release(wempty);
                    if (c == 3) {
                      tc_commit(bar(B_ACC + (m & 1)));
                      if (m == 9) tc_commit(bar(B_RGBREADY));
                    }
                  }
                  __syncwarp();
                  advance();
                }
              }
              a_par ^= 1;
            }
            if (probing && lane == 0 && tile_i == 5 && (m == 2 || m == 3)) dbg[14 * 148 + 16 * (int)cta_pinned + (m == 2 ? 0 : 12)] = clock64();
            if (probing && lane == 0 && (tile_i == 5 || (tile_i == 6 && m < 2))) dbg[32 * 148 + 40 * (int)cta_pinned + m + (tile_i == 6 ? 30 : 0)] = clock64();   // whole-tile timeline
          }
      }
      if (probing && lane == 0) dbg[8 * (int)cta_pinned] = clock64() - t0;
    }
  } else if (warp >= EMB_WARP0) {
    // ================= embedding producers: E operand of the NEXT tile while the current one runs; =====
    // ================= then the colour head of the tile that has just finished                      =====
    const int row = (int)tid_pinned - EMB_WARP0 * 32;    // 0..127
    const int hq = warp & 3, hrow = hq * 32 + lane;      // colour head: the TMEM lane quadrant this warp can address
    const uint32_t rgb0_addr = tmem_base + ((uint32_t)(hq * 32) << 16) + 256u;   // rgb.0 accumulators: buffer 1, columns [0,128)
    uint32_t tile_i = 0;
    long long m_t0 = clock64(), m_wait = 0, mtt = 0;
    auto tile_of = [&](uint32_t ti) { return (group0 + (int)ti * group_step) * CLUSTER + (int)cta_rank; };
    auto colour_head = [&](int tile, uint32_t ti) {
      mbar_wait(bar(B_RGBREADY), ti & 1);
      tc_fence_after();
      const long long hg = (long long)tile * TILE + hrow;
      float* const dst = hg < total ? out_rgb + 3 * hg : nullptr;
      if (TRAIN && !(flags & 4096)) rgb_head<true>(rgb0_addr, bar(B_RGBFREE), lane, hrow, tail_s, save.act + act_chunk_off(9, (size_t)num_tiles, (size_t)tile, 0), dst);   // (4096: timing experiment, layer-9 activations not saved)
      else rgb_head<false>(rgb0_addr, bar(B_RGBFREE), lane, hrow, tail_s, nullptr, dst);
    };
    for (int grp = group0; grp < n_groups; grp += group_step, ++tile_i) {
      const uint32_t eb = tile_i % NEB;
      const int tile = grp * CLUSTER + (int)cta_rank;
      // tile_i - 2 finishes its layer 9 about when it releases the E buffer this iteration refills: the colour head is
      // the urgent one (layer 1 of tile_i - 1 waits for it), the embedding is not needed for another tile
      // (split precision: ONE E buffer, released by layer 9 of tile_i - 1, whose colour head therefore comes first)
      if (!PREC && tile_i >= 2) colour_head(tile_of(tile_i - 2), tile_i - 2);
      if (timing) mtt = clock64();
      mbar_wait(bar(B_EEMPTY + eb), ((tile_i / NEB) & 1) ^ 1);
      if (timing) m_wait += clock64() - mtt;
      uint8_t* sE = smem + OFF_E + eb * E_BYTES;
      long long g = (long long)tile * TILE + row;
      const bool valid = g < total;
      if (!valid) g = total - 1;
      const int r = (int)(g / S), j = (int)(g % S);
      float o[3] = {ray_o[3 * r], ray_o[3 * r + 1], ray_o[3 * r + 2]};
      float d[3] = {ray_d[3 * r], ray_d[3 * r + 1], ray_d[3 * r + 2]};
      float x[4];
      if (BG) {
        BgRay br = bg_ray_setup(o, d);
        float dr = bg_point(br, z[(size_t)r * S + (S - 1 - j)], x);   // flipped order, ddp_model.py:116-117
        if (valid) out_depth_real[g] = dr;
      } else {
        float zv = z[g];
        for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(o[c], __fmul_rn(zv, d[c]));   // ddp_model.py:91
      }
      float dn = norm3(d[0], d[1], d[2]);
      float vd[3] = {d[0] / dn, d[1] / dn, d[2] / dn};                              // ddp_model.py:82-83
      if (!(flags & 512)) {   // experiment: 512 = no embedding work (results wrong, timing only)
        embed_vec<PREC ? E_BYTES : 0>(x, D, NF_POS, sE, row, 0);
        embed_vec<PREC ? E_BYTES : 0>(vd, 3, NF_VIEW, sE, row, VIEW_COL);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_EFULL + eb));
      // split precision: the single E buffer is released early in the previous tile's layer 9 (its view-direction MMAs come
      // first), so the encoding above overlaps the rest of that layer; its colour head follows here
      if (PREC && tile_i >= 1) colour_head(tile_of(tile_i - 1), tile_i - 1);
    }
    for (uint32_t ti = tile_i >= NEB ? tile_i - NEB : 0; ti < tile_i; ++ti) colour_head(tile_of(ti), ti);
    if (timing && (int)tid_pinned == EMB_WARP0 * 32) { dbg[8 * (int)cta_pinned + 4] = clock64() - m_t0; dbg[8 * (int)cta_pinned + 7] = m_wait; }
  } else {
    // ================= epilogue warps =================
    // Thread = one accumulator row (TMEM lane); warps w and w+4 split each 64-column chunk.  The fp16
    // activations are written back IN PLACE over the first half of the fp32 columns just read
    // (chunk j, half hh: K values [32hh,32hh+32) -> TMEM columns [64j+32hh, 64j+32hh+16)), where the
    // next layer's MMAs read them as the A operand: no shared-memory round trip.
    const int q = warp & 3, hh = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t acc_par = 0, stg_flip = 0;
    long long e_wait = 0, e_t0 = clock64(), ett = 0;
    float pend = 0.f;            // raw sigma of the previous tile's row, written out under the next tile's layer-1 MMAs
    long long pend_g = -1;
    auto flush_pending = [&]() {
      if (pend_g >= 0) {
        out_sigma[pend_g] = fabsf(pend);
        if (TRAIN) save.raw_sigma[pend_g] = pend;
        pend_g = -1;
      }
    };
    for (int grp = group0; grp < n_groups; grp += group_step) {
      const int tile = grp * CLUSTER + (int)cta_rank;
      const long long g = (long long)tile * TILE + row;
      const bool valid = g < total;
      float sig_part = 0.f;
      long long* const probe_base = (probing && (int)tid_pinned == 0 && grp == group0 + 5 * group_step) ? dbg + 14 * 148 + 16 * (int)cta_pinned : nullptr;
      uint8_t* const act_tile = TRAIN ? save.act + act_chunk_off(0, (size_t)num_tiles, (size_t)tile, 0) : nullptr;
      const size_t act_layer_stride = act_layer_off(1, (size_t)num_tiles);
#pragma unroll 1
      for (int m = 0; m < NUM_MMA_LAYERS - 1; ++m) {       // rgb.0 (layer 9) is read by the colour head, not here
        const int ab = m & 1;
        if (timing) ett = clock64();
        const uint32_t acc_ph = (acc_par >> ab) & 1u;
        mbar_wait(bar((NSPLIT ? B_ACCH : B_ACC) + ab), acc_ph);      // NSPLIT: columns [0,128) first; the rest is awaited inside
        if (timing) e_wait += clock64() - ett;
        acc_par ^= 1u << ab;
        tc_fence_after();
        long long* const probe = (probe_base && m == 2) ? probe_base + 8 : nullptr;
        if (probe) probe_base[1] = clock64();
        if (probe_base) dbg[32 * 148 + 40 * (int)cta_pinned + 10 + m] = clock64();
        const uint32_t acc_addr = lane_addr + (uint32_t)(ab * 256 + 32 * hh);
        if (TRAIN) epilogue_dispatch<true, false>(m, acc_addr, bar(B_AREADY), lane, row, hh, tail_s, act_tile + m * act_layer_stride,
                                              reinterpret_cast<uint4*>(save.mask + mask_off(m & 7, (size_t)num_tiles, (size_t)tile, hh, row)),
                                              smem + OFF_STG_TRAIN, stg_flip, sig_part, probe, flags);
        else epilogue_dispatch<false, PREC, NSPLIT>(m, acc_addr, bar(B_AREADY), lane, row, hh, tail_s, nullptr, nullptr, nullptr, stg_flip, sig_part, probe,
                                                    0, bar(B_ACC + ab), acc_ph);
        if (probe_base) dbg[32 * 148 + 40 * (int)cta_pinned + 20 + m] = clock64();
        if (m == 0) flush_pending();
      }
      // layer 9's phase of accumulator barrier 1 is not waited for here; it has completed by the time this warp next
      // waits on that barrier (layer 1 of the next tile is issued behind the next layer 0, whose completion it awaits first)
      acc_par ^= 2u;
      // ---- combine the two column halves of sigma -------------------------------------------------------
      // Warps w and w+4 own the same TMEM lanes, so the hh = 1 partial sum crosses over through TMEM columns that are
      // free between base_remap and the next tile's layer 1 (buffer 1, columns 128-131).
      const uint32_t xaddr = lane_addr + 256u + 128u;
      if (hh == 1) {
        tmem_st4(xaddr, __float_as_uint(sig_part), 0u, 0u, 0u);
        tmem_st_wait();
      }
      tc_fence_before();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      tc_fence_after();
      if (hh == 0) {
        uint32_t p[4];
        tmem_ld4_sync(xaddr, p);
        pend = sig_part + __uint_as_float(p[0]) + tail_s[T_BSIG];
        pend_g = valid ? g : -1;
      }
    }
    flush_pending();
    if (TRAIN) stage_store_drain();
    if (timing && (int)tid_pinned == 0) {
      dbg[8 * (int)cta_pinned + 5] = clock64() - e_t0; dbg[8 * (int)cta_pinned + 6] = e_wait;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();   // no peer may still multicast into / arrive on this CTA
  if (warp == LOAD_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---- packer: state-dict tensors -> fp16 tiles in MMA issue order + fp32 tail ------------------------
__device__ __forceinline__ float low_half(float v) { return v - __half2float(__float2half_rn(v)); }

__global__ void pack_tc_kernel(NerfppNetParams p, bool bg, bool prec, uint8_t* __restrict__ out, int blob_total) {
  const StepTable& tab = c_tab[(prec ? 2 : 0) + (bg ? 1 : 0)];
  const int i = blockIdx.y;
  if (i < tab.n) {
    const Step s = tab.s[i];
    const int pl = param_layer(s.layer), nin = layer_in(pl, bg);
    const float* Wl = p.w[pl];
    const float* Bl = p.b[pl];
    const int bias_bytes = s.bias ? s.n * BIAS_ROW_BYTES : 0;
    if (s.bias) {
      __half* bt = reinterpret_cast<__half*>(out + s.blob_off);
      for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < s.n * 16; idx += gridDim.x * blockDim.x) {
        const int nn = idx >> 4, kk = idx & 15;
        const float b = Bl[nn];
        const __half hi = __float2half_rn(b);
        __half v = __float2half_rn(0.f);
        if (kk == ONE_COL - 112) v = hi;
        if (kk == ONE_COL + 1 - 112) v = __float2half_rn(b - __half2float(hi));
        bt[bias_tile_off(s.n, nn, kk) / 2] = v;
      }
    }
    __half* blob = reinterpret_cast<__half*>(out + s.blob_off + bias_bytes);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < s.n * 64; idx += gridDim.x * blockDim.x) {
      const int nn = idx >> 6, kk = idx & 63;
      float v = 0.f;
      if (s.src == SRC_E) {
        if (s.col0 == -2) {           // rgb.0: view part + bias columns
          const int c = 64 * s.chunk + kk;
          if (c >= VIEW_COL && c < VIEW_COL + VIEW_DIM) v = Wl[(size_t)nn * nin + W + (c - VIEW_COL)];
          else if ((c == ONE_COL || c == ONE_COL + 1) && !s.lo) {       // (a low-half tile carries no bias: its columns stay 0)
            const float b = Bl[nn];
            const float hi = __half2float(__float2half_rn(b));
            v = (c == ONE_COL) ? hi : b - hi;
          }
        } else {
          const int c = s.col0 + kk;
          if (c < emb_dim(bg)) v = Wl[(size_t)nn * nin + c];                                  // embedding part
        }
      } else {
        v = Wl[(size_t)nn * nin + s.col0 + kk];
      }
      if (s.lo && !(s.src == SRC_E && s.col0 == -2 && (64 * s.chunk + kk == ONE_COL || 64 * s.chunk + kk == ONE_COL + 1))) v = low_half(v);
      blob[((nn >> 3) * 1024 + (nn & 7) * 128 + (((kk >> 3) ^ (nn & 7)) << 4)) / 2 + (kk & 7)] = __float2half_rn(v);
    }
  } else {
    float* tail = reinterpret_cast<float*>(out + blob_total);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < T_TOTAL; idx += gridDim.x * blockDim.x) {
      float v = 0.f;
      if (idx < T_WRGB2) v = p.w[L_SIGMA][idx - T_WSIG];
      else if (idx < T_BSIG) v = p.w[L_RGB2][idx - T_WRGB2];
      else if (idx == T_BSIG) v = p.b[L_SIGMA][0];
      else v = p.b[L_RGB2][idx - T_BRGB2];
      tail[idx] = v;
    }
  }
}

}  // namespace tc
}  // namespace npp

using namespace npp;

static int g_cluster = -1;      // weight-sharing cluster size; NERFPP_TC_CLUSTER overrides (1 or 2)
static long long* g_dbg = nullptr;
static int g_flags = 0;          // experiment switches (diagnostics only)
static int g_nsplit = 0;         // column-half issue order of the fast inference kernel (A/B: nerfpp_debug_set_tc_nsplit)

static void tc_config() {
  static bool done = false;
  if (done) return;
  done = true;
  if (const char* f = getenv("NERFPP_TC_FLAGS")) g_flags = atoi(f);
  if (g_cluster < 0) {
    const char* c = getenv("NERFPP_TC_CLUSTER");
    g_cluster = c ? atoi(c) : 1;
  }
  if (g_cluster != 1 && g_cluster != 2) g_cluster = 1;
}

size_t npp_tc_packed_bytes(bool bg, bool prec) { tc_config(); return (size_t)tc::h_tab[(prec ? 2 : 0) + bg].total + tc::T_TOTAL * sizeof(float); }

int npp_pack_tc(const NerfppNetParams* p, bool bg, bool prec, void* out, cudaStream_t st) {
  tc_config();
  const tc::StepTable& t = tc::h_tab[(prec ? 2 : 0) + bg];
  tc::pack_tc_kernel<<<dim3(8, t.n + 1), 256, 0, st>>>(*p, bg, prec, (uint8_t*)out, t.total);
  NPP_CHECK_LAUNCH();
  return 0;
}

template <bool BG, int CLUSTER, bool TRAIN, bool PREC = false, bool NSPLIT = false>
static int launch_tc(int max_ctas, const uint8_t* blobs, const float* tail, const float* ray_o, const float* ray_d, const float* z,
                     int n, int S, float* out_sigma, float* out_rgb, float* out_dr, int num_tiles, tc::TrainSave save, cudaStream_t st) {
  auto kern = tc::field_tc_kernel<BG, CLUSTER, TRAIN, PREC, NSPLIT>;
  static bool configured_dev[64] = {false};       // the shared-memory opt-in is per device
  static int max_clusters = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  bool& configured = configured_dev[dev & 63];
  if (!configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
    cudaLaunchConfig_t q{};
    q.gridDim = dim3(max_ctas / CLUSTER * CLUSTER); q.blockDim = dim3(tc::THREADS); q.dynamicSmemBytes = tc::SMEM_BYTES;
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = CLUSTER; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    q.attrs = a; q.numAttrs = 1;
    if (CLUSTER == 1 || cudaOccupancyMaxActiveClusters(&max_clusters, kern, &q) != cudaSuccess || max_clusters <= 0) max_clusters = max_ctas / CLUSTER;
    cudaGetLastError();
    configured = true;
  }
  const int n_groups = (num_tiles + CLUSTER - 1) / CLUSTER;
  const int clusters = n_groups < max_clusters ? n_groups : max_clusters;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * CLUSTER); cfg.blockDim = dim3(tc::THREADS); cfg.dynamicSmemBytes = tc::SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, blobs, tail, ray_o, ray_d, z, n, S, out_sigma, out_rgb, out_dr, num_tiles, save, g_dbg, g_flags);
  if (e != cudaSuccess) { npp_set_error("field_tc launch (cluster %d): %s", CLUSTER, cudaGetErrorString(e)); return (int)e; }
  return 0;
}

// debug hooks (tests/diag only): device buffer receiving per-role cycle counters; cluster size override
extern "C" void nerfpp_debug_set_tc_timers(long long* dev_buf) { g_dbg = dev_buf; }
extern "C" void nerfpp_debug_set_tc_cluster(int c) { g_cluster = (c == 2) ? 2 : 1; }
extern "C" void nerfpp_debug_set_tc_flags(int f) { g_flags = f; }
extern "C" void nerfpp_debug_set_tc_nsplit(int on) { g_nsplit = on; }

size_t npp_tc_train_ws_bytes(long long n_samples) { return tc::train_ws_bytes((size_t)((n_samples + tc::TILE - 1) / tc::TILE)); }

int npp_field_tc(const void* packed, bool bg, bool prec, const float* ray_o, const float* ray_d, const float* z, int n, int S,
                 float* out_sigma, float* out_rgb, float* out_depth_real, void* train_ws, cudaStream_t st) {
  static int sms_dev[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& num_sms = sms_dev[dev & 63];
  if (num_sms == 0) cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  tc_config();
  const long long total = (long long)n * S;
  const int num_tiles = (int)((total + tc::TILE - 1) / tc::TILE);
  const uint8_t* blobs = (const uint8_t*)packed;
  const float* tail = (const float*)(blobs + tc::h_tab[(prec ? 2 : 0) + bg].total);
  tc::TrainSave save{nullptr, nullptr, nullptr, nullptr};
  if (train_ws) {
    uint8_t* w = (uint8_t*)train_ws;
    save.act = w; save.e = w + tc::train_ws_e_off((size_t)num_tiles); save.raw_sigma = (float*)(w + tc::train_ws_sigma_off((size_t)num_tiles));
    save.mask = w + tc::train_ws_mask_off((size_t)num_tiles);
  }
#define NPP_TC_LAUNCH(BG, C, T) launch_tc<BG, C, T>(num_sms, blobs, tail, ray_o, ray_d, z, n, S, out_sigma, out_rgb, out_depth_real, num_tiles, save, st)
  if (prec) {         // split precision: inference only
    if (train_ws) { npp_set_error("field_tc: the split-precision variant has no training mode"); return -1; }
    return bg ? launch_tc<true, 1, false, true>(num_sms, blobs, tail, ray_o, ray_d, z, n, S, out_sigma, out_rgb, out_depth_real, num_tiles, save, st)
              : launch_tc<false, 1, false, true>(num_sms, blobs, tail, ray_o, ray_d, z, n, S, out_sigma, out_rgb, out_depth_real, num_tiles, save, st);
  }
  if (train_ws) {     // (the cluster variant is an experiment: NERFPP_TC_CLUSTER=2)
    if (bg) return g_cluster == 1 ? NPP_TC_LAUNCH(true, 1, true) : NPP_TC_LAUNCH(true, 2, true);
    return g_cluster == 1 ? NPP_TC_LAUNCH(false, 1, true) : NPP_TC_LAUNCH(false, 2, true);
  }
  if (g_nsplit && g_cluster == 1)      // the fast inference kernel with the column-half issue order
    return bg ? launch_tc<true, 1, false, false, true>(num_sms, blobs, tail, ray_o, ray_d, z, n, S, out_sigma, out_rgb, out_depth_real, num_tiles, save, st)
              : launch_tc<false, 1, false, false, true>(num_sms, blobs, tail, ray_o, ray_d, z, n, S, out_sigma, out_rgb, out_depth_real, num_tiles, save, st);
  if (bg) return g_cluster == 1 ? NPP_TC_LAUNCH(true, 1, false) : NPP_TC_LAUNCH(true, 2, false);
  return g_cluster == 1 ? NPP_TC_LAUNCH(false, 1, false) : NPP_TC_LAUNCH(false, 2, false);
#undef NPP_TC_LAUNCH
}
