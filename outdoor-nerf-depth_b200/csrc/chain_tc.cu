// PropMLP (4 x 256, density only; configs/360.gin:10-14) as ONE persistent kernel: a 128-sample tile's activations never
// leave the SM.  Layer by layer through gemm_tc_kernel the 256-wide layers are bound by the activations' HBM round trip
// (128 FLOP per byte; measured 44 us per layer at 60 % of the DRAM peak, tensor pipe 43 % active); here a tile's encoding
// is the only thing read (1 KB per sample) and its density the only thing written.
//
//   * two tiles (slots X, Y) are in flight per SM: while the MMAs of one run, the epilogue of the other converts its
//     accumulator D[slot] (256 TMEM columns, fp32) into the next layer's A operand -- fp16, SWIZZLE_128B chunks in the
//     slot's 64 KB activation buffer in shared memory.  Issue order per pair of tiles: L0(X) L0(Y) L1(X) L1(Y) ... L3(Y).
//   * layer 0's A operand (the encoding, 8 chunks of [128 x 64]) streams by tensor-map TMA through the four chunk areas of
//     the slot's activation buffer (free until layer 0's epilogue writes it); every layer's weights stream through a
//     3-stage ring of [256 x 64] chunks, in MMA issue order.
//   * the density head (256 -> 1, softplus(raw - 1)) is evaluated by layer 3's epilogue on the fp32 accumulators.
#include "gemm_tc.cuh"

namespace npp {
namespace chain {
using namespace tc;

constexpr int WIDTH = 256, DEPTH = 4, K0_CHUNKS = 8, KL_CHUNKS = 4;
constexpr int THREADS = 320;                 // warp 0: TMA producer, warp 1: MMA issuer, warps 2..9: epilogue
constexpr int ACT_BYTES = 4 * CHUNK_BYTES;   // one slot: 4 chunks of [128 x 64] fp16
constexpr int W_STAGE = WIDTH * 64 * 2;      // 32 KB
constexpr int NW = 3;
constexpr int OFF_ACT = 0;
constexpr int OFF_W = 2 * ACT_BYTES;
constexpr int OFF_HEAD = OFF_W + NW * W_STAGE;            // 257 floats (+ 4 x 128 floats of pair exchange)
constexpr int HEAD_BYTES = 264 * 4 + 4 * 32 * 4;
constexpr int OFF_BAR = OFF_HEAD + HEAD_BYTES;
// barriers
constexpr int B_WFULL = 0, B_WEMPTY = NW, B_AFULL = 2 * NW, B_AEMPTY = 2 * NW + 8, B_DFULL = 2 * NW + 16, B_ACTRDY = 2 * NW + 18,
              B_COUNT = 2 * NW + 20;
constexpr int SMEM_BYTES = OFF_BAR + 8 * B_COUNT + 16 + 1024;

struct ChainArgs {
  CUtensorMap enc;            // [M, 512] fp16, box {64, 128}
  CUtensorMap w[DEPTH];       // [256, K_l] fp16, box {64, 256}
  const float* bias[DEPTH];
  const float* head;          // 256 weights + bias
  float* density;             // [M]
  int M;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void mma_rt(uint32_t d_tmem, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %5, 0;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n}" ::"r"(d_tmem), "r"(alo), "r"(blo), "r"(SW128_HI), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) prop_chain_kernel(const __grid_constant__ ChainArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t s_base = (raw + 1023u) & ~1023u;
  uint8_t* const smem = smem_raw + (s_base - raw);
  const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
  const uint32_t bar0 = s_base + OFF_BAR;
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * B_COUNT);
  float* const head_s = reinterpret_cast<float*>(smem + OFF_HEAD);
  float* const xchg = head_s + 264;                       // [4 quadrants][32 rows] partial head sums of the hh = 1 warps

  const int n_tiles = (g.M + TILE - 1) / TILE;
  // this CTA's tiles: blockIdx.x, + gridDim.x, ...; processed in pairs (slot 0, slot 1)
  const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int n_pairs = (my_tiles + 1) / 2;
  auto tile_of = [&](int pair, int slot) { const int k = 2 * pair + slot; return k < my_tiles ? (int)blockIdx.x + k * (int)gridDim.x : -1; };

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NW; ++i) { mbar_init(bar(B_WFULL + i), 1); mbar_init(bar(B_WEMPTY + i), 1); }
    for (int i = 0; i < 8; ++i) { mbar_init(bar(B_AFULL + i), 1); mbar_init(bar(B_AEMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(B_DFULL + i), 1); mbar_init(bar(B_ACTRDY + i), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void*)tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = (int)threadIdx.x; i < WIDTH + 1; i += THREADS) head_s[i] = g.head[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: encoding chunks (layer 0) and weight chunks, in MMA issue order =================
    if (lane == 0) {
      uint32_t wit = 0, ait[2] = {0, 0};
      for (int pair = 0; pair < n_pairs; ++pair)
        for (int l = 0; l < DEPTH; ++l)
          for (int slot = 0; slot < 2; ++slot) {
            const int tile = tile_of(pair, slot);
            if (tile < 0) continue;
            const int nch = l == 0 ? K0_CHUNKS : KL_CHUNKS;
            for (int c = 0; c < nch; ++c) {
              if (l == 0) {
                // area c & 3 of the slot's activation buffer: free once its last readers (the previous tile's layer 3, or this
                // tile's layer-0 chunk c - 4) have completed
                const uint32_t a = ait[slot]++, area = a & 3, ph = (a >> 2) & 1;
                const uint32_t ab = B_AFULL + slot * 4 + area, eb = B_AEMPTY + slot * 4 + area;
                mbar_wait(bar(eb), ph ^ 1);
                mbar_expect_tx(bar(ab), CHUNK_BYTES);
                tma_load_2d(s_base + OFF_ACT + slot * ACT_BYTES + area * CHUNK_BYTES, &g.enc, c * 64, tile * TILE, bar(ab));
              }
              const uint32_t st = wit % NW, ph = (wit / NW) & 1;
              ++wit;
              mbar_wait(bar(B_WEMPTY + st), ph ^ 1);
              mbar_expect_tx(bar(B_WFULL + st), W_STAGE);
              tma_load_2d(s_base + OFF_W + st * W_STAGE, &g.w[l], c * 64, 0, bar(B_WFULL + st));
            }
          }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = idesc_f16(WIDTH);
    uint32_t wit = 0, ait[2] = {0, 0}, rdy[2] = {0, 0};
    for (int pair = 0; pair < n_pairs; ++pair)
      for (int l = 0; l < DEPTH; ++l)
        for (int slot = 0; slot < 2; ++slot) {
          if (tile_of(pair, slot) < 0) continue;
          // D[slot] drained (and, l > 0, the slot's activation buffer written) by the epilogue of the previous layer / tile
          { const uint32_t r = rdy[slot]++; mbar_wait(bar(B_ACTRDY + slot), (r & 1) ^ 1); }
          tc_fence_after();
          const uint32_t d = tmem_base + slot * WIDTH;
          const int nch = l == 0 ? K0_CHUNKS : KL_CHUNKS;
          for (int c = 0; c < nch; ++c) {
            uint32_t area = c & 3;
            if (l == 0) {
              const uint32_t a = ait[slot]++;
              area = a & 3;
              mbar_wait(bar(B_AFULL + slot * 4 + area), (a >> 2) & 1);
            }
            const uint32_t st = wit % NW, ph = (wit / NW) & 1;
            ++wit;
            mbar_wait(bar(B_WFULL + st), ph);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t alo = sw128_lo(s_base + OFF_ACT + slot * ACT_BYTES + area * CHUNK_BYTES);
              const uint32_t blo = sw128_lo(s_base + OFF_W + st * W_STAGE);
              mma_rt(d, alo, blo, idesc, c > 0 ? 1u : 0u);
              mma_rt(d, alo + 2, blo + 2, idesc, 1u);
              mma_rt(d, alo + 4, blo + 4, idesc, 1u);
              mma_rt(d, alo + 6, blo + 6, idesc, 1u);
              tc_commit(bar(B_WEMPTY + st));
              // the area's next TMA load may land once these MMAs have read it: layer-0 chunks 0..3 (chunk c + 4 follows),
              // and layer 3 (the next tile's encoding follows)
              if ((l == 0 && c < 4) || l == DEPTH - 1) tc_commit(bar(B_AEMPTY + slot * 4 + area));
              if (c == nch - 1) tc_commit(bar(B_DFULL + slot));
            }
            __syncwarp();
          }
        }
  } else {
    // ================= epilogue =================
    // warp (q, hh): rows [32q, 32q+32) of the tile, 64-column chunks c = hh and hh + 2.  Layers 0..2: bias, ReLU, fp16 into
    // the slot's activation buffer (the next layer's A operand).  Layer 3: the density head on the fp32 values.
    const int e = warp - 2, q = warp & 3, hh = e >> 2;
    const int row = q * 32 + lane;
    uint32_t dph[2] = {0, 0};
    for (int pair = 0; pair < n_pairs; ++pair)
      for (int l = 0; l < DEPTH; ++l)
        for (int slot = 0; slot < 2; ++slot) {
          const int tile = tile_of(pair, slot);
          if (tile < 0) continue;
          { const uint32_t p = dph[slot]++; mbar_wait(bar(B_DFULL + slot), p & 1); }
          tc_fence_after();
          const float* bias = g.bias[l];
          uint8_t* const act = smem + OFF_ACT + slot * ACT_BYTES;
          float head_acc = 0.f;
#pragma unroll 1
          for (int c = hh; c < 4; c += 2) {
            uint32_t v0[32], v1[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + slot * WIDTH + c * 64;
            tmem_ld32(taddr, v0);
            tmem_ld32(taddr + 32, v1);
            const float4* bp = reinterpret_cast<const float4*>(bias + c * 64);
            tmem_ld_wait(v0);
            tmem_ld_wait(v1);
            if (l < DEPTH - 1) {
              uint32_t pk[32];
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                const uint32_t* v = half ? v1 : v0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 b4 = __ldg(bp + half * 8 + i);
                  pk[half * 16 + 2 * i] = pack_f16x2<true>(__float_as_uint(__uint_as_float(v[4 * i]) + b4.x), __float_as_uint(__uint_as_float(v[4 * i + 1]) + b4.y));
                  pk[half * 16 + 2 * i + 1] = pack_f16x2<true>(__float_as_uint(__uint_as_float(v[4 * i + 2]) + b4.z), __float_as_uint(__uint_as_float(v[4 * i + 3]) + b4.w));
                }
              }
              uint4* rowp = reinterpret_cast<uint4*>(act + c * CHUNK_BYTES + row * 128);
#pragma unroll
              for (int u = 0; u < 8; ++u) rowp[u ^ (row & 7)] = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
            } else {
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                const uint32_t* v = half ? v1 : v0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 b4 = __ldg(bp + half * 8 + i);
                  const float* hw = head_s + c * 64 + half * 32 + 4 * i;
                  head_acc += fmaxf(__uint_as_float(v[4 * i]) + b4.x, 0.f) * hw[0] + fmaxf(__uint_as_float(v[4 * i + 1]) + b4.y, 0.f) * hw[1]
                            + fmaxf(__uint_as_float(v[4 * i + 2]) + b4.z, 0.f) * hw[2] + fmaxf(__uint_as_float(v[4 * i + 3]) + b4.w, 0.f) * hw[3];
                }
              }
            }
          }
          if (l < DEPTH - 1) fence_proxy_async();          // generic-proxy writes -> visible to the MMAs' operand reads
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(B_ACTRDY + slot));
          if (l == DEPTH - 1) {
            // the two column halves of a row meet through shared memory (64-thread named barrier per lane quadrant)
            if (hh == 1) xchg[q * 32 + lane] = head_acc;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
            if (hh == 0) {
              const long long r = (long long)tile * TILE + row;
              if (r < g.M) {
                const float x = head_acc + xchg[q * 32 + lane] + head_s[WIDTH] - 1.f;      // density_bias = -1 (models.py:375)
                g.density[r] = fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));                    // softplus
              }
            }
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");                      // xchg free for the next tile
          }
        }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}


// ---- split-precision variant ---------------------------------------------------------------------------------------------
// Every operand as a hi + lo fp16 pair, three MMA groups per K-chunk (A_hi W_hi + A_lo W_hi + A_hi W_lo).  The activation
// buffer then takes 2 x 64 KB, so ONE tile is in flight; the layer boundary is bridged the way the NeRF++ field kernel does
// it instead: the epilogue hands the next layer's A operand over chunk by chunk (a barrier per 64-column chunk), so the next
// layer's MMAs start behind the first converted chunk, and the two accumulators alternate between consecutive layers.
constexpr int P_OFF_HI = 0, P_OFF_LO = ACT_BYTES;
constexpr int PB_WFULL = 0, PB_WEMPTY = NW, PB_AFULL = 2 * NW, PB_AEMPTY = 2 * NW + 4, PB_DFULL = 2 * NW + 8, PB_ACTRDY = 2 * NW + 9,
              PB_DRAINED = 2 * NW + 13, PB_COUNT = 2 * NW + 15;
static_assert(PB_COUNT <= B_COUNT, "the two variants share the barrier area");

struct ChainPrecArgs {
  CUtensorMap enc, enc_lo;
  CUtensorMap w[DEPTH], w_lo[DEPTH];
  const float* bias[DEPTH];
  const float* head;
  float* density;
  int M;
};

__global__ void __launch_bounds__(THREADS, 1) prop_chain_prec_kernel(const __grid_constant__ ChainPrecArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t s_base = (raw + 1023u) & ~1023u;
  uint8_t* const smem = smem_raw + (s_base - raw);
  const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
  const uint32_t bar0 = s_base + OFF_BAR;
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * B_COUNT);
  float* const head_s = reinterpret_cast<float*>(smem + OFF_HEAD);
  float* const xchg = head_s + 264;
  const int n_tiles = (g.M + TILE - 1) / TILE;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NW; ++i) { mbar_init(bar(PB_WFULL + i), 1); mbar_init(bar(PB_WEMPTY + i), 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(bar(PB_AFULL + i), 1); mbar_init(bar(PB_AEMPTY + i), 1); mbar_init(bar(PB_ACTRDY + i), 4); }
    mbar_init(bar(PB_DFULL), 1);
    for (int i = 0; i < 2; ++i) mbar_init(bar(PB_DRAINED + i), 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void*)tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = (int)threadIdx.x; i < WIDTH + 1; i += THREADS) head_s[i] = g.head[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: per K-chunk the encoding's hi + lo chunks (layer 0), then W_hi and W_lo =================
    if (lane == 0) {
      uint32_t wit = 0, ait = 0;
      for (int tile = (int)blockIdx.x; tile < n_tiles; tile += (int)gridDim.x)
        for (int l = 0; l < DEPTH; ++l) {
          const int nch = l == 0 ? K0_CHUNKS : KL_CHUNKS;
          for (int c = 0; c < nch; ++c) {
            if (l == 0) {
              const uint32_t a = ait++, area = a & 3, ph = (a >> 2) & 1;
              mbar_wait(bar(PB_AEMPTY + area), ph ^ 1);
              mbar_expect_tx(bar(PB_AFULL + area), 2 * CHUNK_BYTES);
              tma_load_2d(s_base + P_OFF_HI + area * CHUNK_BYTES, &g.enc, c * 64, tile * TILE, bar(PB_AFULL + area));
              tma_load_2d(s_base + P_OFF_LO + area * CHUNK_BYTES, &g.enc_lo, c * 64, tile * TILE, bar(PB_AFULL + area));
            }
            for (int half = 0; half < 2; ++half) {          // W_hi chunk, then W_lo chunk
              const uint32_t st = wit % NW, ph = (wit / NW) & 1;
              ++wit;
              mbar_wait(bar(PB_WEMPTY + st), ph ^ 1);
              mbar_expect_tx(bar(PB_WFULL + st), W_STAGE);
              tma_load_2d(s_base + OFF_W + st * W_STAGE, half ? &g.w_lo[l] : &g.w[l], c * 64, 0, bar(PB_WFULL + st));
            }
          }
        }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = idesc_f16(WIDTH);
    uint32_t wit = 0, ait = 0, use[2] = {0, 0}, rdy = 0;
    auto groups4 = [&](uint32_t d, uint32_t alo, uint32_t blo, uint32_t first_acc) {
      mma_rt(d, alo, blo, idesc, first_acc);
      mma_rt(d, alo + 2, blo + 2, idesc, 1u);
      mma_rt(d, alo + 4, blo + 4, idesc, 1u);
      mma_rt(d, alo + 6, blo + 6, idesc, 1u);
    };
    for (int tile = (int)blockIdx.x; tile < n_tiles; tile += (int)gridDim.x)
      for (int l = 0; l < DEPTH; ++l) {
        const uint32_t b = (uint32_t)l & 1u;
        // the epilogue of the layer that last used accumulator b has read all of it
        if (use[b] > 0) mbar_wait(bar(PB_DRAINED + b), (use[b] - 1) & 1);
        ++use[b];
        tc_fence_after();
        const uint32_t d = tmem_base + b * WIDTH;
        const int nch = l == 0 ? K0_CHUNKS : KL_CHUNKS;
        for (int c = 0; c < nch; ++c) {
          uint32_t area = (uint32_t)c & 3u;
          if (l == 0) {
            const uint32_t a = ait++;
            area = a & 3;
            mbar_wait(bar(PB_AFULL + area), (a >> 2) & 1);
          } else {
            // chunk c of this layer's A operand has been written by the previous layer's epilogue (layers 1..3: the
            // barrier completes three times per tile)
            mbar_wait(bar(PB_ACTRDY + c), (rdy + (uint32_t)(l - 1)) & 1);
          }
          const uint32_t st0 = wit % NW, ph0 = (wit / NW) & 1;
          ++wit;
          const uint32_t st1 = wit % NW, ph1 = (wit / NW) & 1;
          ++wit;
          mbar_wait(bar(PB_WFULL + st0), ph0);
          tc_fence_after();
          const uint32_t ahi = sw128_lo(s_base + P_OFF_HI + area * CHUNK_BYTES), alo_ = sw128_lo(s_base + P_OFF_LO + area * CHUNK_BYTES);
          if (elect_one()) {
            const uint32_t bhi = sw128_lo(s_base + OFF_W + st0 * W_STAGE);
            groups4(d, ahi, bhi, c > 0 ? 1u : 0u);       // A_hi W_hi
            groups4(d, alo_, bhi, 1u);                    // A_lo W_hi
            tc_commit(bar(PB_WEMPTY + st0));
          }
          __syncwarp();
          mbar_wait(bar(PB_WFULL + st1), ph1);
          tc_fence_after();
          if (elect_one()) {
            groups4(d, ahi, sw128_lo(s_base + OFF_W + st1 * W_STAGE), 1u);      // A_hi W_lo
            tc_commit(bar(PB_WEMPTY + st1));
            if ((l == 0 && c < 4) || l == DEPTH - 1) tc_commit(bar(PB_AEMPTY + area));
            if (c == nch - 1) tc_commit(bar(PB_DFULL));
          }
          __syncwarp();
        }
        if (l == DEPTH - 1) rdy += 3;        // three completions of every ACTRDY barrier per tile
      }
  } else {
    // ================= epilogue =================
    const int e = warp - 2, q = warp & 3, hh = e >> 2;
    const int row = q * 32 + lane;
    uint32_t dph = 0;
    for (int tile = (int)blockIdx.x; tile < n_tiles; tile += (int)gridDim.x)
      for (int l = 0; l < DEPTH; ++l) {
        mbar_wait(bar(PB_DFULL), dph & 1);
        ++dph;
        tc_fence_after();
        const float* bias = g.bias[l];
        const uint32_t b = (uint32_t)l & 1u;
        float head_acc = 0.f;
#pragma unroll 1
        for (int c = hh; c < 4; c += 2) {
          uint32_t v0[32], v1[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + b * WIDTH + c * 64;
          tmem_ld32(taddr, v0);
          tmem_ld32(taddr + 32, v1);
          const float4* bp = reinterpret_cast<const float4*>(bias + c * 64);
          tmem_ld_wait(v0);
          tmem_ld_wait(v1);
          if (c + 2 >= 4) {              // this warp's last read of the accumulator
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(PB_DRAINED + b));
          }
          if (l < DEPTH - 1) {
            uint32_t pk[32], pl[32];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const uint32_t* v = half ? v1 : v0;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 b4 = __ldg(bp + half * 8 + i);
                const float x0 = fmaxf(__uint_as_float(v[4 * i]) + b4.x, 0.f), x1 = fmaxf(__uint_as_float(v[4 * i + 1]) + b4.y, 0.f);
                const float x2 = fmaxf(__uint_as_float(v[4 * i + 2]) + b4.z, 0.f), x3 = fmaxf(__uint_as_float(v[4 * i + 3]) + b4.w, 0.f);
                const uint32_t p0 = pack_f16x2<false>(__float_as_uint(x0), __float_as_uint(x1));
                const uint32_t p1 = pack_f16x2<false>(__float_as_uint(x2), __float_as_uint(x3));
                const __half2 h0 = *reinterpret_cast<const __half2*>(&p0), h1 = *reinterpret_cast<const __half2*>(&p1);
                pk[half * 16 + 2 * i] = p0; pk[half * 16 + 2 * i + 1] = p1;
                pl[half * 16 + 2 * i] = pack_f16x2<false>(__float_as_uint(x0 - __low2float(h0)), __float_as_uint(x1 - __high2float(h0)));
                pl[half * 16 + 2 * i + 1] = pack_f16x2<false>(__float_as_uint(x2 - __low2float(h1)), __float_as_uint(x3 - __high2float(h1)));
              }
            }
            uint4* rowh = reinterpret_cast<uint4*>(smem + P_OFF_HI + c * CHUNK_BYTES + row * 128);
            uint4* rowl = reinterpret_cast<uint4*>(smem + P_OFF_LO + c * CHUNK_BYTES + row * 128);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              rowh[u ^ (row & 7)] = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
              rowl[u ^ (row & 7)] = make_uint4(pl[4 * u], pl[4 * u + 1], pl[4 * u + 2], pl[4 * u + 3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(PB_ACTRDY + c));          // the four quadrant warps of this column half: count 4
          } else {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const uint32_t* v = half ? v1 : v0;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 b4 = __ldg(bp + half * 8 + i);
                const float* hw = head_s + c * 64 + half * 32 + 4 * i;
                head_acc += fmaxf(__uint_as_float(v[4 * i]) + b4.x, 0.f) * hw[0] + fmaxf(__uint_as_float(v[4 * i + 1]) + b4.y, 0.f) * hw[1]
                          + fmaxf(__uint_as_float(v[4 * i + 2]) + b4.z, 0.f) * hw[2] + fmaxf(__uint_as_float(v[4 * i + 3]) + b4.w, 0.f) * hw[3];
              }
            }
          }
        }
        if (l == DEPTH - 1) {
          if (hh == 1) xchg[q * 32 + lane] = head_acc;
          asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
          if (hh == 0) {
            const long long r = (long long)tile * TILE + row;
            if (r < g.M) {
              const float x = head_acc + xchg[q * 32 + lane] + head_s[WIDTH] - 1.f;
              g.density[r] = fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
            }
          }
          asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        }
      }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

}  // namespace chain
}  // namespace npp

// enc [M, 512] fp16 (the kernel reads 504 + 8 zero columns), w[l] fp16 [256, k_pad[l]] (k_pad 512, 256, 256, 256), bias[l]
// fp32 [256], head = 256 weights + bias fp32, density [M].  enc_lo / w_lo non-NULL: the split-precision variant (low halves).
int npp_prop_chain(const void* enc, const void* enc_lo, const void* const* w, const void* const* w_lo, const int* k_pad,
                   const float* const* bias, const float* head, float* density, long long M, cudaStream_t st) {
  using namespace npp::chain;
  static bool configured_dev[64] = {false};
  static int sms_dev[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured_dev[dev & 63]) {
    cudaFuncSetAttribute(prop_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(prop_chain_prec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaDeviceGetAttribute(&sms_dev[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    configured_dev[dev & 63] = true;
  }
  const int n_tiles = (int)((M + npp::tc::TILE - 1) / npp::tc::TILE);
  if (n_tiles <= 0) return 0;
  if (enc_lo) {
    ChainPrecArgs g{};
    if (npp::gemm::make_map(&g.enc, enc, 512, (uint64_t)M, 512, npp::tc::TILE)) return -1;
    if (npp::gemm::make_map(&g.enc_lo, enc_lo, 512, (uint64_t)M, 512, npp::tc::TILE)) return -1;
    for (int l = 0; l < DEPTH; ++l) {
      if (npp::gemm::make_map(&g.w[l], w[l], (uint64_t)k_pad[l], WIDTH, (uint64_t)k_pad[l], WIDTH)) return -1;
      if (npp::gemm::make_map(&g.w_lo[l], w_lo[l], (uint64_t)k_pad[l], WIDTH, (uint64_t)k_pad[l], WIDTH)) return -1;
      g.bias[l] = bias[l];
    }
    g.head = head; g.density = density; g.M = (int)M;
    const int grid = n_tiles < sms_dev[dev & 63] ? n_tiles : sms_dev[dev & 63];
    prop_chain_prec_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(g);
  } else {
    ChainArgs g{};
    if (npp::gemm::make_map(&g.enc, enc, 512, (uint64_t)M, 512, npp::tc::TILE)) return -1;
    for (int l = 0; l < DEPTH; ++l) {
      if (npp::gemm::make_map(&g.w[l], w[l], (uint64_t)k_pad[l], WIDTH, (uint64_t)k_pad[l], WIDTH)) return -1;
      g.bias[l] = bias[l];
    }
    g.head = head; g.density = density; g.M = (int)M;
    // two tiles per CTA at a time: with fewer than 2 tiles per SM, use fewer CTAs so that every CTA has a pair to overlap
    const int grid = (n_tiles + 1) / 2 < sms_dev[dev & 63] ? (n_tiles + 1) / 2 : sms_dev[dev & 63];
    prop_chain_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(g);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { npp_set_error("prop_chain launch: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}
