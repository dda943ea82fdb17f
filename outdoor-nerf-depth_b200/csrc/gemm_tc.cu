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
__device__ __forceinline__ void mma_ss_rt(uint32_t d_tmem, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %5, 0;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n}" ::"r"(d_tmem), "r"(alo), "r"(blo), "r"(SW128_HI), "r"(idesc), "r"(acc)
      : "memory");
}

// barrier slots
template <int STAGES> struct Bars {
  static constexpr int FULL = 0, EMPTY = STAGES, TFULL = 2 * STAGES, TEMPTY = 2 * STAGES + 2, COUNT = 2 * STAGES + 4;
};

template <int BN, bool PREC>
__global__ void __launch_bounds__(THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmArgs g) {
  using C = Cfg<BN, PREC>;
  using B = Bars<C::STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t s_base = (raw + 1023u) & ~1023u;
  uint8_t* const smem = smem_raw + (s_base - raw);
  const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
  const uint32_t bar0 = s_base + C::OFF_BAR;
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_BAR + 8 * B::COUNT);

  const int n_blks = g.N / BN;
  const int m_blks = (g.M + BM - 1) / BM;
  const int n_tiles = n_blks * m_blks;
  int chunks_per_tile = 0;
  for (int s = 0; s < g.n_seg; ++s) chunks_per_tile += g.seg[s].n_chunks;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) { mbar_init(bar(B::FULL + i), 1); mbar_init(bar(B::EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(B::TFULL + i), 1); mbar_init(bar(B::TEMPTY + i), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void*)tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      for (int i = 0; i < 4; ++i) prefetch_map(&g.a[i]);
      prefetch_map(&g.w[0]);
      uint32_t it = 0;
      for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x) {
        const int m0 = (t / n_blks) * BM, n0 = (t % n_blks) * BN;
        for (int s = 0; s < g.n_seg; ++s) {
          const Segment sg = g.seg[s];
          for (int c = 0; c < sg.n_chunks; ++c, ++it) {
            const uint32_t st = it % C::STAGES, ph = (it / C::STAGES) & 1;
            mbar_wait(bar(B::EMPTY + st), ph ^ 1);
            mbar_expect_tx(bar(B::FULL + st), C::STAGE_BYTES);
            const uint32_t dst = s_base + st * C::STAGE_BYTES;
            tma_load_2d(dst, &g.a[sg.a_map], sg.a_col + c * BK, m0, bar(B::FULL + st));
            tma_load_2d(dst + C::A_BYTES, &g.w[sg.w_map], sg.w_col + c * BK, n0, bar(B::FULL + st));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (converged warp, one elected lane issues) =================
    constexpr uint32_t idesc = idesc_f16(BN);
    uint32_t it = 0, tcount = 0;
    for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x, ++tcount) {
      const uint32_t buf = tcount & 1, tph = (tcount >> 1) & 1;
      mbar_wait(bar(B::TEMPTY + buf), tph ^ 1);          // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d = tmem_base + buf * BN;
      for (int c = 0; c < chunks_per_tile; ++c, ++it) {
        const uint32_t st = it % C::STAGES, ph = (it / C::STAGES) & 1;
        mbar_wait(bar(B::FULL + st), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t alo = sw128_lo(s_base + st * C::STAGE_BYTES);
          const uint32_t blo = sw128_lo(s_base + st * C::STAGE_BYTES + C::A_BYTES);
          mma_ss_rt(d, alo, blo, idesc, c > 0 ? 1u : 0u);
          mma_ss_rt(d, alo + 2, blo + 2, idesc, 1u);
          mma_ss_rt(d, alo + 4, blo + 4, idesc, 1u);
          mma_ss_rt(d, alo + 6, blo + 6, idesc, 1u);
          tc_commit(bar(B::EMPTY + st));                  // stage free once these MMAs have read it
          if (c == chunks_per_tile - 1) tc_commit(bar(B::TFULL + buf));
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue: TMEM -> (+bias, ReLU) -> fp16 -> shared -> TMA store =================
    const int e = warp - 2, q = warp & 3, hh = e >> 2;
    const int row = q * 32 + lane;
    const bool issuer = (warp == 2) && (lane == 0);
    uint32_t tcount = 0, oc = 0;
    for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x, ++tcount) {
      const int m0 = (t / n_blks) * BM, n0 = (t % n_blks) * BN;
      const uint32_t buf = tcount & 1, tph = (tcount >> 1) & 1;
      mbar_wait(bar(B::TFULL + buf), tph);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c, ++oc) {
        const int col0 = c * 64 + hh * 32;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + col0, v);
        float bv[32];
        const float4* bp = reinterpret_cast<const float4*>(g.bias + n0 + col0);
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float4 b4 = __ldg(bp + i); bv[4 * i] = b4.x; bv[4 * i + 1] = b4.y; bv[4 * i + 2] = b4.z; bv[4 * i + 3] = b4.w; }
        tmem_ld_wait(v);
        uint32_t pk[16], pl[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x0 = __uint_as_float(v[2 * i]) + bv[2 * i], x1 = __uint_as_float(v[2 * i + 1]) + bv[2 * i + 1];
          if (g.relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
          pk[i] = pack_f16x2<false>(__float_as_uint(x0), __float_as_uint(x1));
          if (PREC) {
            const __half2 h = *reinterpret_cast<const __half2*>(&pk[i]);
            const float l0 = x0 - __low2float(h), l1 = x1 - __high2float(h);
            pl[i] = pack_f16x2<false>(__float_as_uint(l0), __float_as_uint(l1));
          }
        }
        const uint32_t ob = PREC ? (oc & 1) * 2 : (oc & 1);
        uint8_t* sb = smem + C::OFF_OUT + ob * C::OUT_BYTES;
        uint4* rowp = reinterpret_cast<uint4*>(sb + row * 128);
#pragma unroll
        for (int u = 0; u < 4; ++u) rowp[(4 * hh + u) ^ (row & 7)] = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        if (PREC) {
          uint4* rowl = reinterpret_cast<uint4*>(sb + C::OUT_BYTES + row * 128);
#pragma unroll
          for (int u = 0; u < 4; ++u) rowl[(4 * hh + u) ^ (row & 7)] = make_uint4(pl[4 * u], pl[4 * u + 1], pl[4 * u + 2], pl[4 * u + 3]);
        }
        fence_proxy_async();
        if (issuer) bulk_s2g_wait_read();      // the previous chunk's store has read its buffer: free for the next chunk
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (issuer) {
          tma_store_2d(&g.out[0], n0 + c * 64, m0, s_base + C::OFF_OUT + ob * C::OUT_BYTES);
          if (PREC) tma_store_2d(&g.out[1], n0 + c * 64, m0, s_base + C::OFF_OUT + (ob + 1) * C::OUT_BYTES);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B::TEMPTY + buf));
    }
    if (issuer) stage_store_drain();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
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

template <int BN, bool PREC>
static int launch_t(const GemmArgs& g, cudaStream_t st) {
  using C = Cfg<BN, PREC>;
  auto kern = gemm_tc_kernel<BN, PREC>;
  static bool configured_dev[64] = {false};
  static int sms_dev[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured_dev[dev & 63]) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    cudaDeviceGetAttribute(&sms_dev[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    configured_dev[dev & 63] = true;
  }
  const int n_tiles = (g.N / BN) * ((g.M + BM - 1) / BM);
  const int grid = n_tiles < sms_dev[dev & 63] ? n_tiles : sms_dev[dev & 63];
  if (grid <= 0) return 0;
  kern<<<grid, THREADS, C::SMEM_BYTES, st>>>(g);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { npp_set_error("gemm_tc launch: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

int launch_gemm(const GemmArgs& g, int bn, bool prec, cudaStream_t st) {
  if (g.N % bn != 0 || (bn != 256 && bn != 128)) { npp_set_error("gemm_tc: N %d is not a multiple of the tile width %d", g.N, bn); return -1; }
  if (bn == 256) return prec ? launch_t<256, true>(g, st) : launch_t<256, false>(g, st);
  return prec ? launch_t<128, true>(g, st) : launch_t<128, false>(g, st);
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
  const int bn = (N % 256 == 0) ? 256 : 128;
  GemmArgs g{};
  if (make_map(&g.a[0], a, (uint64_t)K, (uint64_t)M, (uint64_t)K, BM)) return -1;
  for (int i = 1; i < 4; ++i) g.a[i] = g.a[0];
  if (make_map(&g.w[0], w, (uint64_t)K, (uint64_t)N, (uint64_t)K, (uint32_t)bn)) return -1;
  g.w[1] = g.w[0];
  if (make_map(&g.out[0], out, (uint64_t)N, (uint64_t)M, (uint64_t)N, BM)) return -1;
  g.out[1] = g.out[0];
  g.seg[0] = Segment{0, 0, 0, 0, (K + BK - 1) / BK};
  g.n_seg = 1; g.M = M; g.N = N; g.relu = relu; g.bias = bias;
  return launch_gemm(g, bn, false, st);
}
