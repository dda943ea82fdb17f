// Backward of the field MLP, data-gradient half, on the 5th-generation tensor cores.
//
// Same machine as field_tc.cu run in reverse: one persistent CTA per SM walks the 128-sample tiles; per tile nine
// dense layers  dX = dZ * W  (rgb.0 -> base_remap -> base 7 ... base 1, transposed weight tiles streamed through the
// same shared-memory ring) run as tcgen05.mma with fp32 accumulators in TMEM; the epilogue multiplies by the ReLU
// mask of the saved forward activation, rounds to fp16, writes the result back IN PLACE as the next layer's A operand
// and stores it to the DZ buffer (operand-chunk layout, tc_common.cuh) for the weight-gradient kernel.
//
//   warp 0-7   epilogue (TMEM -> regs -> mask -> fp16 -> TMEM in place + global DZ)
//   warp 8     MMA issuer (converged warp, elect.sync)
//   warp 9     weight-tile loader (cp.async.bulk ring)
//   warp 10-13 prologue of the NEXT tile: d(raw rgb) = d rgb * c (1 - c), d(raw sigma) = d sigma * sign, and the first
//              operand  dG = (d raw rgb . W_rgb2) * [rgb hidden > 0]  as an fp16 SW128 tile in shared memory
//
// All gradients inside this kernel are multiplied by a power-of-two loss scale (read from device memory) so that they
// survive the fp16 operand format; wgrad_tc.cu divides it out again.
#include <cstdlib>
#include "bwd_common.cuh"

namespace npp {
namespace tcb {

constexpr int OFF_G = 0, OFF_W = 2 * G_BYTES, OFF_BAR = OFF_W + NSTAGE * STAGE_BYTES, OFF_W2 = OFF_BAR + 512;   // dG operand double-buffered
constexpr int OFF_STG = OFF_W2 + 3 * RGB_HID * 4;   // staging of the DZ store (tc_common.cuh: stage_store_chunk), 1024-aligned
constexpr int SMEM_BYTES = OFF_STG + STG_BYTES;
static_assert(OFF_STG % 1024 == 0 && SMEM_BYTES <= 232448, "shared memory budget");
constexpr int NUM_EPI_WARPS = 8, MMA_WARP = 8, LOAD_WARP = 9, PRO_WARP0 = 10, NUM_PRO_WARPS = 4, THREADS = 448;

enum { B_WFULL = 0, B_WEMPTY = NSTAGE, B_AREADY = 2 * NSTAGE, B_GFULL = B_AREADY + 4, B_GEMPTY = B_GFULL + 2,
       B_ACC = B_GEMPTY + 2, B_COUNT = B_ACC + 2 };
static_assert(8 * B_COUNT + 8 <= 512, "barrier area");

__constant__ StepTable c_tab = make_table();
static const StepTable h_tab = make_table();

__global__ void __launch_bounds__(THREADS, 1)
field_dgrad_kernel(const uint8_t* __restrict__ blobs, const float* __restrict__ tail, const uint8_t* __restrict__ act,
                   const uint8_t* __restrict__ mask, const float* __restrict__ rgb, const float* __restrict__ raw_sigma, const float* __restrict__ d_sigma,
                   const float* __restrict__ d_rgb, const float* __restrict__ scale_ptr, long long total, int num_tiles,
                   uint8_t* __restrict__ dz, float* __restrict__ d_raw_sigma, float* __restrict__ d_raw_rgb, int flags) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar0 = s_base + OFF_BAR;
  auto bar = [&](int i) { return bar0 + 8u * i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * B_COUNT);
  if ((s_base & 1023u) != 0) __trap();
  const float scale = scale_ptr[0], scale_rgb = scale_ptr[1];   // below the join | colour path (backward.cu)
  const float join = scale / scale_rgb;
  const StepTable& tab = c_tab;

  if (warp == MMA_WARP && lane == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar(B_WFULL + i), 1); mbar_init(bar(B_WEMPTY + i), 1); }
    for (int i = 0; i < 4; ++i) mbar_init(bar(B_AREADY + i), NUM_EPI_WARPS);
    for (int i = 0; i < 2; ++i) { mbar_init(bar(B_GFULL + i), NUM_PRO_WARPS); mbar_init(bar(B_GEMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) mbar_init(bar(B_ACC + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == LOAD_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void*)tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 3 * RGB_HID; i += THREADS) reinterpret_cast<float*>(smem + OFF_W2)[i] = tail[T_W2 + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == LOAD_WARP) {
    // ================= weight loader =================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
        for (int i = 0; i < tab.n; ++i, ++it) {
          const uint32_t st = it % NSTAGE, ph = (it / NSTAGE) & 1;
          mbar_wait(bar(B_WEMPTY + st), ph ^ 1);
          if ((flags & 8) && it >= NSTAGE) { mbar_arrive(bar(B_WFULL + st)); continue; }   // timing experiment: no refill (wrong results)
          mbar_expect_tx(bar(B_WFULL + st), STAGE_BYTES);
          bulk_g2s(s_base + OFF_W + st * STAGE_BYTES, blobs + tab.s[i].blob_off, STAGE_BYTES, bar(B_WFULL + st));
        }
    }
  } else if (warp == MMA_WARP) {
    // ================= MMA issuer =================
    constexpr uint32_t ID256 = idesc_f16(256);
    const uint32_t ring0 = s_base + OFF_W, wfull0 = bar(B_WFULL), wempty0 = bar(B_WEMPTY);
    uint32_t st = 0, ph = 0, slot = ring0, wfull = wfull0, wempty = wempty0;
    uint32_t a_par = 0, tile_i = 0, acc_cnt[2] = {0, 0};
    auto advance = [&]() {
      ++st; slot += STAGE_BYTES; wfull += 8; wempty += 8;
      if (st == NSTAGE) { st = 0; ph ^= 1; slot = ring0; wfull = wfull0; wempty = wempty0; }
    };
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_i) {
      const uint32_t gb = tile_i & 1;
      const uint32_t g_addr = s_base + OFF_G + gb * G_BYTES;
      mbar_wait(bar(B_GFULL + gb), (tile_i >> 1) & 1);
      // the previous tile's last layer (buffer (8 + tile_i - 1) & 1) read its A operand from the buffer this tile's first
      // layer is about to overwrite: wait until those MMAs have completed
      if (tile_i > 0) { const uint32_t b8 = (8 + tile_i - 1) & 1; mbar_wait(bar(B_ACC + b8), (acc_cnt[b8] - 1) & 1); }
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < NUM_LAYERS; ++t) {
        const uint32_t buf = (t + tile_i) & 1;
        const uint32_t d_tmem = tmem_base + buf * 256u;
        const uint32_t a_tmem = tmem_base + (buf ^ 1u) * 256u;
        if (t == 0) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            mbar_wait(wfull, ph);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t alo = sw128_lo(g_addr + c * CHUNK_BYTES), blo = sw128_lo(slot);
              if (c == 0) mma_ss<0>(d_tmem, alo, SW128_HI, blo, SW128_HI, ID256); else mma_ss<1>(d_tmem, alo, SW128_HI, blo, SW128_HI, ID256);
              mma_ss<1>(d_tmem, alo + 2u, SW128_HI, blo + 2u, SW128_HI, ID256);
              mma_ss<1>(d_tmem, alo + 4u, SW128_HI, blo + 4u, SW128_HI, ID256);
              mma_ss<1>(d_tmem, alo + 6u, SW128_HI, blo + 6u, SW128_HI, ID256);
              tc_commit(wempty);
              if (c == 1) { tc_commit(bar(B_GEMPTY + gb)); tc_commit(bar(B_ACC + buf)); }
            }
            __syncwarp();
            advance();
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            mbar_wait2(bar(B_AREADY + c), a_par, wfull, ph);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t blo = sw128_lo(slot), a0 = a_tmem + 64u * c;
              if (c == 0) mma_ts<0>(d_tmem, a0, blo, ID256); else mma_ts<1>(d_tmem, a0, blo, ID256);
              mma_ts<1>(d_tmem, a0 + 8u, blo + 2u, ID256);
              mma_ts<1>(d_tmem, a0 + 32u, blo + 4u, ID256);
              mma_ts<1>(d_tmem, a0 + 40u, blo + 6u, ID256);
              tc_commit(wempty);
              if (c == 3) tc_commit(bar(B_ACC + buf));
            }
            __syncwarp();
            advance();
          }
          a_par ^= 1;
        }
        ++acc_cnt[buf];
      }
    }
  } else if (warp >= PRO_WARP0) {
    // ================= prologue producers: dG operand of the NEXT tile =================
    const int row = threadIdx.x - PRO_WARP0 * 32;    // 0..127
    const float* w2 = reinterpret_cast<const float*>(smem + OFF_W2);
    uint32_t tile_i = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_i) {
      const uint32_t gb = tile_i & 1;
      mbar_wait(bar(B_GEMPTY + gb), ((tile_i >> 1) & 1) ^ 1);
      uint8_t* sG = smem + OFF_G + gb * G_BYTES;
      const long long g = (long long)tile * TILE + row;
      float dr[3] = {0.f, 0.f, 0.f};
      if (g < total) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { const float cc = rgb[3 * g + c]; dr[c] = d_rgb[3 * g + c] * cc * (1.f - cc) * scale_rgb; }   // sigmoid'
        d_raw_rgb[3 * g] = dr[0]; d_raw_rgb[3 * g + 1] = dr[1]; d_raw_rgb[3 * g + 2] = dr[2];
        const float rs = raw_sigma[g];
        d_raw_sigma[g] = d_sigma[g] * (rs > 0.f ? 1.f : rs < 0.f ? -1.f : 0.f) * scale;                                          // abs'
      }
      // dG[k] = (dr . W_rgb2[:,k]) * [g_k > 0], g = saved rgb hidden (ACT layer 9, 2 chunks)
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {
        const uint8_t* ach = act + act_chunk_off(9, (size_t)num_tiles, (size_t)tile, j);
        uint8_t* dch = dz + act_chunk_off(9, (size_t)num_tiles, (size_t)tile, j);
        const uint32_t roff = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t uoff = roff + (uint32_t)((u ^ (row & 7)) << 4);
          const uint4 hv = *reinterpret_cast<const uint4*>(ach + uoff);
          const __half2* hp = reinterpret_cast<const __half2*>(&hv);
          uint4 ov;
          __half2* op = reinterpret_cast<__half2*>(&ov);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = 64 * j + 8 * u + 2 * e;
            const float v0 = dr[0] * w2[k] + dr[1] * w2[RGB_HID + k] + dr[2] * w2[2 * RGB_HID + k];
            const float v1 = dr[0] * w2[k + 1] + dr[1] * w2[RGB_HID + k + 1] + dr[2] * w2[2 * RGB_HID + k + 1];
            op[e] = __hmul2(__floats2half2_rn(v0, v1), __hgt2(hp[e], __float2half2_rn(0.f)));
          }
          *reinterpret_cast<uint4*>(sG + j * CHUNK_BYTES + uoff) = ov;
          *reinterpret_cast<uint4*>(dch + uoff) = ov;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_GFULL + gb));
    }
  } else {
    // ================= epilogue warps =================
    const int q = warp & 3, hh = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t roff = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
    uint32_t acc_par = 0, tile_i = 0, stg_flip = 0;
    uint8_t* const stg = smem + OFF_STG;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_i) {
      const long long g = (long long)tile * TILE + row;
      float dsr = 0.f;
      if (g < total) { const float rs = raw_sigma[g]; dsr = d_sigma[g] * (rs > 0.f ? 1.f : rs < 0.f ? -1.f : 0.f) * scale; }
#pragma unroll 1
      for (int t = 0; t < NUM_LAYERS; ++t) {
        const uint32_t buf = (t + tile_i) & 1;
        const int l = target_act(t);
        // ReLU mask of the forward activation this layer's output is the gradient of (base_remap, t == 0, had none):
        // the training forward left it as bits (MASK layout, tc_common.cuh), one 16-byte word group per thread and
        // layer.  It is fetched HERE, before the wait for the layer's accumulator, so the load flies while the tensor
        // pipe still runs this layer's MMAs.
        uint4 mb4 = make_uint4(0u, 0u, 0u, 0u);
        if (t != 0) mb4 = __ldg(reinterpret_cast<const uint4*>(mask + mask_off(l, (size_t)num_tiles, (size_t)tile, hh, row)));
        const uint32_t mb[4] = {mb4.x, mb4.y, mb4.z, mb4.w};
        mbar_wait(bar(B_ACC + buf), (acc_par >> buf) & 1u);
        acc_par ^= 1u << buf;
        tc_fence_after();
        const uint32_t acc_addr = lane_addr + buf * 256u + 32u * hh;
        uint32_t v[2][32];
        tmem_ld32(acc_addr, v[0]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t (&cur)[32] = v[j & 1];
          tmem_ld_wait(cur);
          if (j + 1 < 4) tmem_ld32(acc_addr + 64u * (j + 1), v[(j + 1) & 1]);
          if (t == 1) {       // the sigma head joins here: d h7 = (colour path, rescaled to the chain's loss scale) + d(raw sigma) * w_sigma   (nerf_network.py:133)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 w4 = __ldg(reinterpret_cast<const float4*>(tail + T_WSIG + 64 * j + 32 * hh) + e);
              cur[4 * e] = __float_as_uint(fmaf(dsr, w4.x, join * __uint_as_float(cur[4 * e])));
              cur[4 * e + 1] = __float_as_uint(fmaf(dsr, w4.y, join * __uint_as_float(cur[4 * e + 1])));
              cur[4 * e + 2] = __float_as_uint(fmaf(dsr, w4.z, join * __uint_as_float(cur[4 * e + 2])));
              cur[4 * e + 3] = __float_as_uint(fmaf(dsr, w4.w, join * __uint_as_float(cur[4 * e + 3])));
            }
          }
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) pk[e] = pack_f16x2_sat(cur[2 * e], cur[2 * e + 1]);
          if (t != 0) apply_relu_mask(pk, mb[j]);
          if (t != NUM_LAYERS - 1) {
            tmem_st16(acc_addr + 64u * j, pk);      // next layer's A operand, in place
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_AREADY + j));
          }
          // the copy for the weight-gradient kernel leaves behind the arrive, off the chain the tensor pipe waits for,
          // through shared memory and one bulk copy per warp pair
          if (!(flags & 1024)) stage_store_chunk(stg, stg_flip, dz + act_chunk_off(l, (size_t)num_tiles, (size_t)tile, j), q, lane, hh, pk, (flags & 2048) != 0);
        }
      }
    }
    stage_store_drain();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == LOAD_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// packed dgrad weights: for step (t, chunk c) the tile B[n = input feature i][k = output o in [64c, 64c+64)] = W[o][col0 + i]
__global__ void pack_dgrad_kernel(NerfppNetParams p, bool bg, uint8_t* __restrict__ out, int blob_total) {
  const StepTable& tab = c_tab;
  const int i = blockIdx.y;
  if (i < tab.n) {
    const Step s = tab.s[i];
    const int pl = weight_layer(s.t), nin = layer_in(pl, bg);
    const int col0 = (pl == 5) ? emb_dim(bg) : 0;         // base 5 takes [embedding, h4]: only the h4 part carries gradient
    const float* Wl = p.w[pl];
    __half* blob = reinterpret_cast<__half*>(out + s.blob_off);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 256 * 64; idx += gridDim.x * blockDim.x) {
      const int nn = idx >> 6, kk = idx & 63;
      const float v = Wl[(size_t)(64 * s.chunk + kk) * nin + col0 + nn];
      blob[((nn >> 3) * 1024 + (nn & 7) * 128 + (((kk >> 3) ^ (nn & 7)) << 4)) / 2 + (kk & 7)] = __float2half_rn(v);
    }
  } else {
    float* tail = reinterpret_cast<float*>(out + blob_total);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < T_TOTAL; idx += gridDim.x * blockDim.x)
      tail[idx] = idx < T_WSIG ? p.w[L_RGB2][idx] : p.w[L_SIGMA][idx - T_WSIG];
  }
}

}  // namespace tcb
}  // namespace npp

using namespace npp;

size_t npp_dgrad_packed_bytes() { return (size_t)tcb::h_tab.total + tcb::T_TOTAL * sizeof(float); }

int npp_pack_dgrad(const NerfppNetParams* p, bool bg, void* out, cudaStream_t st) {
  tcb::pack_dgrad_kernel<<<dim3(8, tcb::h_tab.n + 1), 256, 0, st>>>(*p, bg, (uint8_t*)out, tcb::h_tab.total);
  NPP_CHECK_LAUNCH();
  return 0;
}

// act / dz: ACT-layout buffers (tc_common.cuh) of the same tiling as the forward's training workspace; mask: its MASK region
int npp_field_dgrad(const void* packed, const void* act, const void* mask, const float* rgb, const float* raw_sigma, const float* d_sigma,
                    const float* d_rgb, const float* scale, long long total, void* dz, float* d_raw_sigma, float* d_raw_rgb,
                    cudaStream_t st) {
  static int sms_dev[64] = {0};               // both caches are per device (the shared-memory opt-in is a per-device attribute)
  static bool configured_dev[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  int& num_sms = sms_dev[dev & 63];
  bool& configured = configured_dev[dev & 63];
  if (num_sms == 0) cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (!configured) { cudaFuncSetAttribute(tcb::field_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcb::SMEM_BYTES); configured = true; }
  const int num_tiles = (int)((total + tc::TILE - 1) / tc::TILE);
  const uint8_t* blobs = (const uint8_t*)packed;
  const float* tail = (const float*)(blobs + tcb::h_tab.total);
  const int grid = num_tiles < num_sms ? num_tiles : num_sms;
  static int dflags = -1;          // experiment switches (diagnostics only), same meaning as field_tc.cu's
  if (dflags < 0) { const char* f = getenv("NERFPP_TC_FLAGS"); dflags = f ? atoi(f) : 0; }
  tcb::field_dgrad_kernel<<<grid, tcb::THREADS, tcb::SMEM_BYTES, st>>>(blobs, tail, (const uint8_t*)act, (const uint8_t*)mask, rgb, raw_sigma, d_sigma, d_rgb,
                                                                        scale, total, num_tiles, (uint8_t*)dz, d_raw_sigma, d_raw_rgb, dflags);
  NPP_CHECK_LAUNCH();
  return 0;
}
