// Persistent tcgen05 GEMM with fused bias / ReLU / fp16 (or hi + lo fp16) output -- see gemm_tc.cuh.
#include <stdio.h>
#include "gemm_tc.cuh"

namespace npp {
namespace gemm {
using namespace tc;

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
template <int CTAS>
__device__ __forceinline__ void mma_ss_rt(uint32_t d_tmem, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t acc) {
  if (CTAS == 1)
    asm volatile(
        "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %5, 0;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n}" ::"r"(d_tmem), "r"(alo), "r"(blo), "r"(SW128_HI), "r"(idesc), "r"(acc)
        : "memory");
  else
    asm volatile(
        "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %5, 0;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n}" ::"r"(d_tmem), "r"(alo), "r"(blo), "r"(SW128_HI), "r"(idesc), "r"(acc)
        : "memory");
}
// completion of this thread's MMAs -> one arrival on the barrier at the same offset in every CTA of the pair
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// pair form of the TMA load: the data lands in THIS CTA's shared memory, the bytes are counted on the leader's barrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(leader_bar) : "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <int M_TOTAL> __host__ __device__ constexpr uint32_t idesc_f16_m(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M_TOTAL >> 4) << 24);
}

// barrier slots
template <int STAGES> struct Bars {
  static constexpr int FULL = 0, EMPTY = STAGES, TFULL = 2 * STAGES, TEMPTY = 2 * STAGES + 2, COUNT = 2 * STAGES + 4;
};

template <int BN, bool PREC, int CTAS, bool F3 = false>
__global__ void __launch_bounds__(THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmArgs g) {
  static_assert(!F3 || (PREC && CTAS == 2), "the fused three-group stage is the split-precision CTA-pair form");
  using C = Cfg<BN, PREC, CTAS, F3>;
  using B = Bars<C::STAGES>;
  static_assert(8 * B::COUNT + 8 <= C::BAR_BYTES, "barrier area");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t s_base = (raw + 1023u) & ~1023u;
  uint8_t* const smem = smem_raw + (s_base - raw);
  const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
  const uint32_t bar0 = s_base + C::OFF_BAR;
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_BAR + 8 * B::COUNT);
  const uint32_t rank = CTAS > 1 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  // tiles: BM * CTAS rows x BN columns, walked by CTA pairs (or single CTAs) with the column block fastest
  const int n_blks = g.N / BN;
  const int m_blks = (g.M + BM * CTAS - 1) / (BM * CTAS);
  const int n_tiles = n_blks * m_blks;
  const int t0 = (int)blockIdx.x / CTAS, t_step = (int)gridDim.x / CTAS;
  int chunks_per_tile = 0;
  for (int s = 0; s < g.n_seg; ++s) chunks_per_tile += g.seg[s].n_chunks;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) { mbar_init(bar(B::FULL + i), 1); mbar_init(bar(B::EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(B::TFULL + i), 1); mbar_init(bar(B::TEMPTY + i), 8 * CTAS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CTAS == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void*)tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void*)tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS > 1) cluster_sync_all();      // the peer's barriers exist before anything remote targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (every CTA: its own A rows, its share of the weight rows) =================
    if (lane == 0) {
      for (int i = 0; i < 4; ++i) prefetch_map(&g.a[i]);
      prefetch_map(&g.w[0]);
      uint32_t it = 0;
      for (int t = t0; t < n_tiles; t += t_step) {
        const int m0 = (t / n_blks) * (BM * CTAS) + (int)rank * BM, n0 = (t % n_blks) * BN + (int)rank * C::B_ROWS;
        for (int s = 0; s < g.n_seg; ++s) {
          const Segment sg = g.seg[s];
          for (int c = 0; c < sg.n_chunks; ++c, ++it) {
            const uint32_t st = it % C::STAGES, ph = (it / C::STAGES) & 1;
            mbar_wait(bar(B::EMPTY + st), ph ^ 1);
            const uint32_t dst = s_base + st * C::STAGE_BYTES;
            if (CTAS == 1) {
              mbar_expect_tx(bar(B::FULL + st), C::STAGE_BYTES);
              tma_load_2d(dst, &g.a[sg.a_map], sg.a_col + c * BK, m0, bar(B::FULL + st));
              tma_load_2d(dst + C::A_BYTES, &g.w[sg.w_map], sg.w_col + c * BK, n0, bar(B::FULL + st));
            } else {
              if (leader) mbar_expect_tx(bar(B::FULL + st), C::STAGE_BYTES * CTAS);     // both CTAs' bytes land on the leader's barrier
              const uint32_t fb = map_to_cta(bar(B::FULL + st), 0);
              if (!F3) {
                tma_load_2d_pair(dst, &g.a[sg.a_map], sg.a_col + c * BK, m0, fb);
                tma_load_2d_pair(dst + C::A_BYTES, &g.w[sg.w_map], sg.w_col + c * BK, n0, fb);
              } else {            // A_hi | A_lo | W_hi | W_lo
                tma_load_2d_pair(dst, &g.a[sg.a_map], sg.a_col + c * BK, m0, fb);
                tma_load_2d_pair(dst + C::A_BYTES, &g.a[sg.a_map + 1], sg.a_col + c * BK, m0, fb);
                tma_load_2d_pair(dst + 2 * C::A_BYTES, &g.w[0], sg.w_col + c * BK, n0, fb);
                tma_load_2d_pair(dst + 2 * C::A_BYTES + C::B_BYTES, &g.w[1], sg.w_col + c * BK, n0, fb);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA; converged warp, one elected lane issues) =================
    if (leader) {
      constexpr uint32_t idesc = idesc_f16_m<BM * CTAS>(BN);
      uint32_t it = 0, tcount = 0;
      for (int t = t0; t < n_tiles; t += t_step, ++tcount) {
        const uint32_t buf = tcount & 1, tph = (tcount >> 1) & 1;
        mbar_wait(bar(B::TEMPTY + buf), tph ^ 1);          // the epilogues (of both CTAs) have drained this accumulator
        tc_fence_after();
        const uint32_t d = tmem_base + buf * BN;
        for (int c = 0; c < chunks_per_tile; ++c, ++it) {
          const uint32_t st = it % C::STAGES, ph = (it / C::STAGES) & 1;
          mbar_wait(bar(B::FULL + st), ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t alo = sw128_lo(s_base + st * C::STAGE_BYTES);
            const uint32_t blo = sw128_lo(s_base + st * C::STAGE_BYTES + (F3 ? 2 : 1) * C::A_BYTES);
            mma_ss_rt<CTAS>(d, alo, blo, idesc, c > 0 ? 1u : 0u);
            mma_ss_rt<CTAS>(d, alo + 2, blo + 2, idesc, 1u);
            mma_ss_rt<CTAS>(d, alo + 4, blo + 4, idesc, 1u);
            mma_ss_rt<CTAS>(d, alo + 6, blo + 6, idesc, 1u);
            if (F3) {
              const uint32_t al2 = sw128_lo(s_base + st * C::STAGE_BYTES + C::A_BYTES);                       // A_lo
              const uint32_t bl2 = sw128_lo(s_base + st * C::STAGE_BYTES + 2 * C::A_BYTES + C::B_BYTES);      // W_lo
#pragma unroll
              for (int k = 0; k < 4; ++k) mma_ss_rt<CTAS>(d, al2 + 2 * k, blo + 2 * k, idesc, 1u);            // A_lo W_hi
#pragma unroll
              for (int k = 0; k < 4; ++k) mma_ss_rt<CTAS>(d, alo + 2 * k, bl2 + 2 * k, idesc, 1u);            // A_hi W_lo
            }
            if (CTAS == 1) {
              tc_commit(bar(B::EMPTY + st));                  // stage free once these MMAs have read it
              if (c == chunks_per_tile - 1) tc_commit(bar(B::TFULL + buf));
            } else {
              tc_commit_pair(bar(B::EMPTY + st));
              if (c == chunks_per_tile - 1) tc_commit_pair(bar(B::TFULL + buf));
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================= epilogue: TMEM -> (+bias, ReLU) -> fp16 -> shared -> TMA store =================
    // warp (q, hh): TMEM lane quadrant q = rows [32q, 32q+32) of the CTA's 128, 64-column chunks c = hh, hh+2, ...: a
    // [32 x 64] sub-tile per chunk, staged in the warp's private 4 KB buffer (SWIZZLE_128B image) and stored by the
    // warp's own TMA store -- no block-wide barrier anywhere in the epilogue.
    const int e = warp - 2, q = warp & 3, hh = e >> 2;
    uint8_t* const stg = smem + C::OFF_OUT + e * C::OUT_WARP_BYTES;
    const uint32_t stg_s = s_base + C::OFF_OUT + e * C::OUT_WARP_BYTES;
    const uint32_t tempty_remote = CTAS > 1 ? map_to_cta(bar(B::TEMPTY), 0) : 0u;
    uint32_t tcount = 0;
    for (int t = t0; t < n_tiles; t += t_step, ++tcount) {
      const int m0 = (t / n_blks) * (BM * CTAS) + (int)rank * BM, n0 = (t % n_blks) * BN;
      const uint32_t buf = tcount & 1, tph = (tcount >> 1) & 1;
      mbar_wait(bar(B::TFULL + buf), tph);
      tc_fence_after();
      float hacc[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = hh; c < BN / 64; c += 2) {
        uint32_t v0[32], v1[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + c * 64;
        tmem_ld32(taddr, v0);
        tmem_ld32(taddr + 32, v1);
        const float4* bp = reinterpret_cast<const float4*>(g.bias + n0 + c * 64);
        tmem_ld_wait(v0);
        tmem_ld_wait(v1);
        if (c + 2 >= BN / 64) {                  // this warp's last read of the accumulator: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (CTAS == 1) mbar_arrive(bar(B::TEMPTY + buf)); else mbar_arrive_cluster(tempty_remote + 8u * buf); }
        }
        uint32_t pk[32], pl[PREC ? 32 : 1];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t* v = half ? v1 : v0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = __ldg(bp + half * 8 + i);
            float x0 = __uint_as_float(v[4 * i]) + b4.x, x1 = __uint_as_float(v[4 * i + 1]) + b4.y;
            float x2 = __uint_as_float(v[4 * i + 2]) + b4.z, x3 = __uint_as_float(v[4 * i + 3]) + b4.w;
            if (g.relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f); }
            if (g.head_w) {
              const float4* hw = reinterpret_cast<const float4*>(g.head_w + n0 + c * 64) + half * 8 + i;
              const float4 w4 = __ldg(hw);
              hacc[0] += x0 * w4.x + x1 * w4.y + x2 * w4.z + x3 * w4.w;
              if (g.head_n > 1) {
                const float4 w5 = __ldg(hw + g.N / 4), w6 = __ldg(hw + g.N / 2);
                hacc[1] += x0 * w5.x + x1 * w5.y + x2 * w5.z + x3 * w5.w;
                hacc[2] += x0 * w6.x + x1 * w6.y + x2 * w6.z + x3 * w6.w;
              }
            }
            const uint32_t p0 = pack_f16x2<false>(__float_as_uint(x0), __float_as_uint(x1));
            const uint32_t p1 = pack_f16x2<false>(__float_as_uint(x2), __float_as_uint(x3));
            pk[half * 16 + 2 * i] = p0; pk[half * 16 + 2 * i + 1] = p1;
            if (PREC) {
              const __half2 h0 = *reinterpret_cast<const __half2*>(&p0), h1 = *reinterpret_cast<const __half2*>(&p1);
              pl[half * 16 + 2 * i] = pack_f16x2<false>(__float_as_uint(x0 - __low2float(h0)), __float_as_uint(x1 - __high2float(h0)));
              pl[half * 16 + 2 * i + 1] = pack_f16x2<false>(__float_as_uint(x2 - __low2float(h1)), __float_as_uint(x3 - __high2float(h1)));
            }
          }
        }
        if (g.no_store) continue;                  // only the fused head's result is wanted
        if (lane == 0) bulk_s2g_wait_read();       // the warp's previous store has read the staging buffer
        __syncwarp();
        uint4* rowp = reinterpret_cast<uint4*>(stg + lane * 128);
#pragma unroll
        for (int u = 0; u < 8; ++u) rowp[u ^ (lane & 7)] = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        if (PREC && !F3) {
          uint4* rowl = reinterpret_cast<uint4*>(stg + 4096 + lane * 128);
#pragma unroll
          for (int u = 0; u < 8; ++u) rowl[u ^ (lane & 7)] = make_uint4(pl[4 * u], pl[4 * u + 1], pl[4 * u + 2], pl[4 * u + 3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&g.out[0], n0 + c * 64, m0 + q * 32, stg_s);
          if (PREC && !F3) tma_store_2d(&g.out[1], n0 + c * 64, m0 + q * 32, stg_s + 4096);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (F3) {                                  // the low halves through the same 4 KB buffer, once the hi store has read it
          if (lane == 0) bulk_s2g_wait_read();
          __syncwarp();
#pragma unroll
          for (int u = 0; u < 8; ++u) rowp[u ^ (lane & 7)] = make_uint4(pl[4 * u], pl[4 * u + 1], pl[4 * u + 2], pl[4 * u + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&g.out[1], n0 + c * 64, m0 + q * 32, stg_s);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      if (g.head_w) {
        const long long r = (long long)m0 + q * 32 + lane;
        if (r < g.M) {
          float* hp = g.head_part + (r * (2 * n_blks) + 2 * (n0 / BN) + hh) * g.head_n;
          hp[0] = hacc[0];
          if (g.head_n > 1) { hp[1] = hacc[1]; hp[2] = hacc[2]; }
        }
      }
    }
    if (lane == 0) stage_store_drain();
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS > 1) cluster_sync_all();      // the leader's MMAs read the peer's shared memory: nobody leaves early
  if (warp == 0) {
    tc_fence_after();
    if (CTAS == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_map(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems, uint32_t box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { npp_set_error("cuTensorMapEncodeTiled is not available from this driver"); return -1; }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {64, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { npp_set_error("cuTensorMapEncodeTiled failed (%d): inner %llu outer %llu pitch %llu", (int)r,
                                         (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_elems); return -1; }
  return 0;
}

static int g_pair_mode = 1;
void set_pair_mode(int mode) { g_pair_mode = mode; }
// tile plan for an output width N and a (padded) depth K: column-block width and CTAs per tile.  Pairs pay where the
// weight stream matters (measured on one box, 131072..262144 rows: N 1024, K 1024: 1244 against 1112 TFLOP/s; N 256,
// K 1024: 1018 / 946); the K = 256 layers of the PropMLP are bound by the activations' HBM traffic and lose with pairs
// (574 / 710).  mode 2 forces pairs (A/B runs).
void plan(int N, int K, int* bn, int* ctas) {
  *bn = (N % 256 == 0) ? 256 : 128;
  *ctas = (*bn == 256 && (g_pair_mode == 2 || (g_pair_mode == 1 && K >= 512))) ? 2 : 1;
}

template <int BN, bool PREC, int CTAS, bool F3 = false>
static int launch_t(const GemmArgs& g, cudaStream_t st) {
  using C = Cfg<BN, PREC, CTAS, F3>;
  auto kern = gemm_tc_kernel<BN, PREC, CTAS, F3>;
  static bool configured_dev[64] = {false};
  static int max_ctas_dev[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured_dev[dev & 63]) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int max_ctas = sms / CTAS * CTAS;
    if (CTAS > 1) {
      cudaLaunchConfig_t q{};
      q.gridDim = dim3(max_ctas); q.blockDim = dim3(THREADS); q.dynamicSmemBytes = C::SMEM_BYTES;
      cudaLaunchAttribute a[1];
      a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = CTAS; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
      q.attrs = a; q.numAttrs = 1;
      int clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&clusters, kern, &q) == cudaSuccess && clusters > 0 && clusters * CTAS < max_ctas) max_ctas = clusters * CTAS;
      cudaGetLastError();
    }
    max_ctas_dev[dev & 63] = max_ctas;
    configured_dev[dev & 63] = true;
  }
  const int n_tiles = (g.N / BN) * ((g.M + BM * CTAS - 1) / (BM * CTAS));
  int grid = n_tiles * CTAS < max_ctas_dev[dev & 63] ? n_tiles * CTAS : max_ctas_dev[dev & 63];
  if (grid <= 0) return 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = CTAS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, g);
  if (e != cudaSuccess) { npp_set_error("gemm_tc launch (BN %d, %d CTA/tile): %s", BN, CTAS, cudaGetErrorString(e)); return (int)e; }
  return 0;
}

int launch_gemm(const GemmArgs& g, int bn, int ctas, bool prec, cudaStream_t st) {
  if (g.N % bn != 0 || (bn != 256 && bn != 128) || (ctas == 2 && bn != 256)) { npp_set_error("gemm_tc: N %d / tile width %d / %d CTAs per tile", g.N, bn, ctas); return -1; }
  if (bn == 256 && ctas == 2) return prec ? (g.fused3 ? launch_t<256, true, 2, true>(g, st) : launch_t<256, true, 2>(g, st)) : launch_t<256, false, 2>(g, st);
  if (bn == 256) return prec ? launch_t<256, true, 1>(g, st) : launch_t<256, false, 1>(g, st);
  return prec ? launch_t<128, true, 1>(g, st) : launch_t<128, false, 1>(g, st);
}

}  // namespace gemm
}  // namespace npp

// Stand-alone entry point (tests, microbenchmarks): out[M,N] = act(A[M,K] W[N,K]^T + bias), fp16 operands and output,
// fp32 accumulation.  K must be a multiple of 8 (16-byte row pitch); columns beyond K are zero-filled by the TMA unit.
extern "C" int mip360_dense_f16(const void* a, const void* w, const float* bias, void* out, int M, int N, int K, int relu,
                                void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  using namespace npp::gemm;
  NPP_CHECK_ARG(a && w && bias && out, "null pointer");
  NPP_CHECK_ARG(M > 0 && N > 0 && K > 0 && K % 8 == 0 && N % 128 == 0, "need M, N, K > 0, K % 8 == 0, N % 128 == 0");
  int bn, ctas;
  plan(N, K, &bn, &ctas);
  GemmArgs g{};
  if (make_map(&g.a[0], a, (uint64_t)K, (uint64_t)M, (uint64_t)K, BM)) return -1;
  for (int i = 1; i < 4; ++i) g.a[i] = g.a[0];
  if (make_map(&g.w[0], w, (uint64_t)K, (uint64_t)N, (uint64_t)K, (uint32_t)(bn / ctas))) return -1;
  g.w[1] = g.w[0];
  if (make_map(&g.out[0], out, (uint64_t)N, (uint64_t)M, (uint64_t)N, OUT_BOX_ROWS)) return -1;
  g.out[1] = g.out[0];
  g.seg[0] = Segment{0, 0, 0, 0, (K + BK - 1) / BK};
  g.n_seg = 1; g.M = M; g.N = N; g.relu = relu; g.bias = bias;
  return launch_gemm(g, bn, ctas, false, st);
}

extern "C" void mip360_debug_set_pair_mode(int mode) { npp::gemm::set_pair_mode(mode); }
