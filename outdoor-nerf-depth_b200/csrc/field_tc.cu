// Field evaluation on the 5th-generation tensor cores.
//
// One persistent CTA per SM walks 128-sample tiles (samples of all rays packed along the MMA M
// axis).  Per tile the ten dense layers (base 0..7, base_remap, rgb.0) run as tcgen05.mma
// (M=128, N=256|128, K=16, fp16 operands, fp32 accumulators in TMEM); sigma (256->1) and rgb.2
// (128->3) are fp32 dot products in the epilogue on the fp32 accumulators.
//
//   warp 0-7  epilogue: positional encoding -> E operand; per layer TMEM -> regs, +bias, ReLU,
//             fp16 pack -> A operand of the next layer (K-major, 128B swizzle), 64 columns at a time
//   warp 8    one lane issues every tcgen05.mma; layer l+1's K-chunk j starts as soon as the
//             epilogue of layer l has produced columns [64j,64j+64) (two TMEM accumulators
//             ping-pong between consecutive layers)
//   warp 9    one lane streams the pre-swizzled weight tiles (32 KB each, in MMA issue order)
//             from L2 into a 4-deep shared-memory ring with cp.async.bulk + mbarrier complete_tx
//
// Shared memory (bytes): E 32K | A 64K | ring 4x32K | barriers.  TMEM: 2 x 256 columns.
// Precision: operands are rounded to fp16 (11-bit significand), products/sums are fp32; measured
// against the fp32 reference: rgb/depth within 3e-5 relative (tests/test_parity_gpu.py).
#include <cuda_fp16.h>
#include <cstdlib>
#include "common.cuh"

namespace npp {
namespace tc {

constexpr int TILE = 128;
constexpr int NSTAGE = 4;
constexpr int STAGE_BYTES = 32768;
constexpr int CHUNK_BYTES = 16384;          // 128 rows x 64 fp16
constexpr int E_BYTES = 2 * CHUNK_BYTES;    // 128 columns: [0,emb) position, [96,123) view dir
constexpr int A_BYTES = 4 * CHUNK_BYTES;    // 256 columns
constexpr int OFF_E = 0, OFF_A = E_BYTES, OFF_W = E_BYTES + A_BYTES, OFF_BAR = OFF_W + NSTAGE * STAGE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256;
constexpr int VIEW_COL = 96;
constexpr int NUM_EPI_WARPS = 8, MMA_WARP = 8, LOAD_WARP = 9, THREADS = 320;
constexpr int NUM_MMA_LAYERS = 10;          // base 0..7, remap, rgb0

// barrier slots
enum { B_WFULL = 0, B_WEMPTY = NSTAGE, B_AREADY = 2 * NSTAGE, B_EREADY = 2 * NSTAGE + 4, B_ACC = 2 * NSTAGE + 5, B_COUNT = 2 * NSTAGE + 7 };

// fp32 tail of the packed buffer (float offsets)
constexpr int T_BIAS = 0;                    // 10 x 256
constexpr int T_WSIG = 10 * 256;             // 256
constexpr int T_WRGB2 = 11 * 256;            // 3 x 128
constexpr int T_BSIG = 11 * 256 + 384;
constexpr int T_BRGB2 = T_BSIG + 1;
constexpr int T_TOTAL = T_BRGB2 + 3;

__host__ __device__ constexpr int param_layer(int m) { return m < 8 ? m : m == 8 ? L_REMAP : L_RGB0; }

// One MMA step = one weight tile in the ring x one 64-column operand chunk.
struct Step {
  short layer;     // MMA layer 0..9
  short src;       // 0 = E region, 1 = A region
  short chunk;     // 64-column chunk within the region
  short k0, nk;    // K=16 sub-steps [k0, k0+nk) of the chunk
  short n;         // MMA N (256 or 128)
  short first, last;
  int col0;        // first input feature (state-dict column) this tile maps, -1 = zero tile part
  int blob_off, blob_bytes;
};
struct StepTable { Step s[48]; int n; int total; };

__host__ __device__ constexpr StepTable make_table(bool bg) {
  StepTable t{};
  int i = 0, off = 0;
  const int e_chunks = bg ? 2 : 1;
  for (int m = 0; m < NUM_MMA_LAYERS; ++m) {
    const int n = (m == 9) ? 128 : 256;
    const int bytes = n * 128;
    int first = 1;
    if (m == 0 || m == 5) {
      for (int c = 0; c < e_chunks; ++c) {
        t.s[i] = Step{(short)m, 0, (short)c, 0, (short)((bg && c == 1) ? 2 : 4), (short)n, (short)first, 0, c * 64, off, bytes};
        first = 0; off += bytes; ++i;
      }
    }
    if (m == 9) {   // view-direction columns [96,128) of E -> rgb.0 inputs 256..282
      t.s[i] = Step{(short)m, 0, 1, 2, 2, (short)n, (short)first, 0, -2, off, bytes};
      first = 0; off += bytes; ++i;
    }
    if (m != 0) {
      for (int c = 0; c < 4; ++c) {
        const int base = (m == 5) ? emb_dim(bg) : 0;
        t.s[i] = Step{(short)m, 1, (short)c, 0, 4, (short)n, (short)first, (short)(c == 3), base + c * 64, off, bytes};
        first = 0; off += bytes; ++i;
      }
    } else {
      t.s[i - 1].last = 1;
    }
  }
  t.n = i;
  t.total = off;
  return t;
}

__constant__ StepTable c_tab[2] = {make_table(false), make_table(true)};
static const StepTable h_tab[2] = {make_table(false), make_table(true)};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// one L2 read delivered to the same shared-memory offset (and mbarrier) of every CTA in cta_mask
__device__ __forceinline__ void bulk_g2s_mcast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t cta_mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mcast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 x fp16 -> fp32
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows of 64 fp16 (128 B), 8-row groups 1024 B apart.
// (cute::UMMA::SmemDescriptor: start>>4 | LBO=1<<16 | SBO=64<<32 | version=1<<46 | SWIZZLE_128B=2<<61)
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D=f32, A=B=f16, both K-major, M=128
__host__ __device__ constexpr uint32_t idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
// byte offset of element (row r, column c) of a [128 x 64k] region made of 64-column SW128 chunks
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)((c >> 6) * CHUNK_BYTES + (r >> 3) * 1024 + (r & 7) * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + (c & 7) * 2);
}

template <int D>
__device__ __forceinline__ void embed_rows(const float* x, int k_lo, int k_hi, bool raw, uint8_t* region, int row, int col_base) {
  if (raw)
    for (int c = 0; c < D; ++c) *reinterpret_cast<__half*>(region + sw128_off(row, col_base + c)) = __float2half_rn(x[c]);
  for (int k = k_lo; k < k_hi; ++k) {
    float f = (float)(1 << k);
    for (int c = 0; c < D; ++c) {
      float s, co;
      sincosf(x[c] * f, &s, &co);
      *reinterpret_cast<__half*>(region + sw128_off(row, col_base + D + 2 * k * D + c)) = __float2half_rn(s);
      *reinterpret_cast<__half*>(region + sw128_off(row, col_base + D + (2 * k + 1) * D + c)) = __float2half_rn(co);
    }
  }
}

// CLUSTER > 1: the CTAs of a cluster walk their tiles in lock step and share every weight tile: each
// CTA fetches 1/CLUSTER of it from L2 and multicasts that slice into all CLUSTER rings, so L2->SM
// traffic (the binding resource of the unshared version, profiles/r1_notes.md) drops CLUSTER-fold.
template <bool BG, int CLUSTER>
__global__ void __launch_bounds__(THREADS, 1)
field_tc_kernel(const uint8_t* __restrict__ blobs, const float* __restrict__ tail, const float* __restrict__ ray_o,
                const float* __restrict__ ray_d, const float* __restrict__ z, int n, int S, float* __restrict__ out_sigma,
                float* __restrict__ out_rgb, float* __restrict__ out_depth_real, int num_tiles, long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int D = BG ? 4 : 3;
  const StepTable& tab = c_tab[BG ? 1 : 0];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  const int group0 = blockIdx.x / CLUSTER, group_step = gridDim.x / CLUSTER;

  if (warp == MMA_WARP && lane == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar(B_WFULL + i), 1); mbar_init(bar(B_WEMPTY + i), CLUSTER); }
    for (int i = 0; i < 4; ++i) mbar_init(bar(B_AREADY + i), NUM_EPI_WARPS);
    mbar_init(bar(B_EREADY), NUM_EPI_WARPS);
    mbar_init(bar(B_ACC), 1); mbar_init(bar(B_ACC + 1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == LOAD_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void*)tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp < NUM_EPI_WARPS) {   // zero E once: padding columns are never written again
    for (int i = threadIdx.x; i < E_BYTES / 16; i += NUM_EPI_WARPS * 32) reinterpret_cast<uint4*>(smem + OFF_E)[i] = make_uint4(0, 0, 0, 0);
  }
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
          const uint32_t st = it % NSTAGE, ph = (it / NSTAGE) & 1;
          const uint32_t bytes = (uint32_t)tab.s[i].blob_bytes;
          mbar_wait(bar(B_WEMPTY + st), ph ^ 1);          // every CTA of the cluster has consumed this stage
          mbar_expect_tx(bar(B_WFULL + st), bytes);
          if (CLUSTER == 1) {
            bulk_g2s(s_base + OFF_W + st * STAGE_BYTES, blobs + tab.s[i].blob_off, bytes, bar(B_WFULL + st));
          } else {
            const uint32_t part = bytes / CLUSTER, o = cta_rank * part;
            bulk_g2s_mcast(s_base + OFF_W + st * STAGE_BYTES + o, blobs + tab.s[i].blob_off + o, part, bar(B_WFULL + st), kMask);
          }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ================= MMA issuer =================
    if (lane == 0) {
      uint32_t it = 0, a_cnt[4] = {0, 0, 0, 0}, e_cnt = 0;
      long long t_e = 0, t_a = 0, t_w = 0, t_i = 0, t0 = clock64(), tt;
      for (int grp = group0; grp < n_groups; grp += group_step) {
        tt = clock64();
        mbar_wait(bar(B_EREADY), e_cnt & 1);
        t_e += clock64() - tt;
        ++e_cnt;
        for (int i = 0; i < tab.n; ++i, ++it) {
          const Step s = tab.s[i];
          tt = clock64();
          if (s.src == 1) { mbar_wait(bar(B_AREADY + s.chunk), a_cnt[s.chunk] & 1); ++a_cnt[s.chunk]; }
          t_a += clock64() - tt;
          const uint32_t st = it % NSTAGE, ph = (it / NSTAGE) & 1;
          tt = clock64();
          mbar_wait(bar(B_WFULL + st), ph);
          t_w += clock64() - tt;
          tt = clock64();
          tc_fence_after();
          const uint32_t a_addr = s_base + (s.src ? OFF_A : OFF_E) + s.chunk * CHUNK_BYTES;
          const uint32_t b_addr = s_base + OFF_W + st * STAGE_BYTES;
          const uint32_t d_tmem = tmem_base + (uint32_t)(s.layer & 1) * 256u;
          const uint32_t idesc = idesc_f16(s.n);
          for (int k = s.k0; k < s.k0 + s.nk; ++k)
            umma_f16(d_tmem, sw128_desc(a_addr + k * 32), sw128_desc(b_addr + k * 32), idesc, (s.first && k == s.k0) ? 0u : 1u);
          if (CLUSTER == 1) tc_commit(bar(B_WEMPTY + st)); else tc_commit_mcast(bar(B_WEMPTY + st), kMask);
          if (s.last) tc_commit(bar(B_ACC + (s.layer & 1)));
          t_i += clock64() - tt;
        }
      }
      if (dbg) { dbg[8 * blockIdx.x] = clock64() - t0; dbg[8 * blockIdx.x + 1] = t_e; dbg[8 * blockIdx.x + 2] = t_a; dbg[8 * blockIdx.x + 3] = t_w; dbg[8 * blockIdx.x + 4] = t_i; }
    }
  } else {
    // ================= epilogue warps =================
    const int q = warp & 3, hh = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint8_t* sE = smem + OFF_E;
    uint8_t* sA = smem + OFF_A;
    uint32_t acc_cnt[2] = {0, 0};
    long long e_wait = 0, e_emb = 0, e_t0 = clock64(), ett;
    for (int grp = group0; grp < n_groups; grp += group_step) {
      ett = clock64();
      const int tile = grp * CLUSTER + (int)cta_rank;
      long long g = (long long)tile * TILE + row;
      const bool valid = g < total;
      if (!valid) g = total - 1;
      {  // ---- positions + encodings -> E (fp16). half 0: raw + freqs 0..4, half 1: freqs 5..9 + view dir
        const int r = (int)(g / S), j = (int)(g % S);
        float o[3] = {ray_o[3 * r], ray_o[3 * r + 1], ray_o[3 * r + 2]};
        float d[3] = {ray_d[3 * r], ray_d[3 * r + 1], ray_d[3 * r + 2]};
        float x[4];
        if (BG) {
          BgRay br = bg_ray_setup(o, d);
          float dr = bg_point(br, z[(size_t)r * S + (S - 1 - j)], x);
          if (hh == 0 && valid) out_depth_real[g] = dr;
        } else {
          float zv = z[g];
          for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(o[c], __fmul_rn(zv, d[c]));
        }
        if (hh == 0) {
          embed_rows<D>(x, 0, 5, true, sE, row, 0);
        } else {
          embed_rows<D>(x, 5, NF_POS, false, sE, row, 0);
          float dn = norm3(d[0], d[1], d[2]);
          float vd[3] = {d[0] / dn, d[1] / dn, d[2] / dn};
          embed_rows<3>(vd, 0, NF_VIEW, true, sE, row, VIEW_COL);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_EREADY));
        e_emb += clock64() - ett;
      }
      float sig_part = 0.f, rgb_part[3] = {0.f, 0.f, 0.f};
      for (int m = 0; m < NUM_MMA_LAYERS; ++m) {
        const int ab = m & 1;
        ett = clock64();
        mbar_wait(bar(B_ACC + ab), acc_cnt[ab] & 1);
        e_wait += clock64() - ett;
        ++acc_cnt[ab];
        tc_fence_after();
        const float* bias = tail + T_BIAS + m * 256;
        const int nchunk = (m == 9) ? 2 : 4;
        for (int jc = 0; jc < nchunk; ++jc) {
          const int col = 64 * jc + 32 * hh;
          uint32_t v[32];
          tmem_ld32(tmem_base + lane_addr + (uint32_t)(ab * 256 + col), v);
          float4 b4[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) b4[t] = __ldg(reinterpret_cast<const float4*>(bias + col) + t);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            f[4 * t + 0] = __uint_as_float(v[4 * t + 0]) + b4[t].x;
            f[4 * t + 1] = __uint_as_float(v[4 * t + 1]) + b4[t].y;
            f[4 * t + 2] = __uint_as_float(v[4 * t + 2]) + b4[t].z;
            f[4 * t + 3] = __uint_as_float(v[4 * t + 3]) + b4[t].w;
          }
          if (m != 8) {
#pragma unroll
            for (int t = 0; t < 32; ++t) f[t] = fmaxf(f[t], 0.f);
          }
          if (m == 7) {   // sigma head on the fp32 activations, nerf_network.py:133
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              float4 w4 = __ldg(reinterpret_cast<const float4*>(tail + T_WSIG + col) + t);
              sig_part = fmaf(f[4 * t], w4.x, sig_part); sig_part = fmaf(f[4 * t + 1], w4.y, sig_part);
              sig_part = fmaf(f[4 * t + 2], w4.z, sig_part); sig_part = fmaf(f[4 * t + 3], w4.w, sig_part);
            }
          }
          if (m == 9) {   // rgb.2 on the fp32 hidden colour features, nerf_network.py:114-117
#pragma unroll
            for (int c = 0; c < 3; ++c) {
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                float4 w4 = __ldg(reinterpret_cast<const float4*>(tail + T_WRGB2 + c * RGB_HID + col) + t);
                rgb_part[c] = fmaf(f[4 * t], w4.x, rgb_part[c]); rgb_part[c] = fmaf(f[4 * t + 1], w4.y, rgb_part[c]);
                rgb_part[c] = fmaf(f[4 * t + 2], w4.z, rgb_part[c]); rgb_part[c] = fmaf(f[4 * t + 3], w4.w, rgb_part[c]);
              }
            }
          } else {        // next layer's A operand, columns [col, col+32) of this row
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              __half2 h0 = __floats2half2_rn(f[8 * t + 0], f[8 * t + 1]), h1 = __floats2half2_rn(f[8 * t + 2], f[8 * t + 3]);
              __half2 h2 = __floats2half2_rn(f[8 * t + 4], f[8 * t + 5]), h3 = __floats2half2_rn(f[8 * t + 6], f[8 * t + 7]);
              uint4 pk = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                    *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
              *reinterpret_cast<uint4*>(sA + sw128_off(row, col + 8 * t)) = pk;
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_AREADY + jc));
          }
        }
      }
      // ---- combine the two column halves of each row, write sigma / rgb -------------------------------
      tc_fence_before();
      float4* scratch = reinterpret_cast<float4*>(sA);   // A is idle: every MMA of this tile has completed
      if (hh == 1) scratch[row] = make_float4(sig_part, rgb_part[0], rgb_part[1], rgb_part[2]);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (hh == 0 && valid) {
        float4 p = scratch[row];
        out_sigma[g] = fabsf(sig_part + p.x + tail[T_BSIG]);
        float c0 = rgb_part[0] + p.y + tail[T_BRGB2], c1 = rgb_part[1] + p.z + tail[T_BRGB2 + 1], c2 = rgb_part[2] + p.w + tail[T_BRGB2 + 2];
        out_rgb[3 * g] = 1.f / (1.f + expf(-c0));
        out_rgb[3 * g + 1] = 1.f / (1.f + expf(-c1));
        out_rgb[3 * g + 2] = 1.f / (1.f + expf(-c2));
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    if (dbg && threadIdx.x == 0) { dbg[8 * blockIdx.x + 5] = clock64() - e_t0; dbg[8 * blockIdx.x + 6] = e_wait; dbg[8 * blockIdx.x + 7] = e_emb; }
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();   // no peer may still multicast into / arrive on this CTA
  if (warp == LOAD_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---- packer: state-dict tensors -> swizzled fp16 tiles in MMA issue order + fp32 tail ---------------
__global__ void pack_tc_kernel(NerfppNetParams p, bool bg, uint8_t* __restrict__ out, int blob_total) {
  const StepTable& tab = c_tab[bg ? 1 : 0];
  const int i = blockIdx.y;
  if (i < tab.n) {
    const Step s = tab.s[i];
    const int pl = param_layer(s.layer), nin = layer_in(pl, bg);
    const float* Wl = p.w[pl];
    __half* blob = reinterpret_cast<__half*>(out + s.blob_off);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < s.n * 64; idx += gridDim.x * blockDim.x) {
      const int nn = idx >> 6, kk = idx & 63;
      int src = -1;
      if (s.src == 0) {
        if (s.col0 == -2) { int vc = kk - (VIEW_COL - 64); if (vc >= 0 && vc < VIEW_DIM) src = W + vc; }   // rgb.0 view part
        else { int c = s.col0 + kk; if (c < emb_dim(bg)) src = c; }                                       // embedding part
      } else {
        src = s.col0 + kk;
      }
      float v = src >= 0 ? Wl[(size_t)nn * nin + src] : 0.f;
      blob[((nn >> 3) * 1024 + (nn & 7) * 128 + (((kk >> 3) ^ (nn & 7)) << 4)) / 2 + (kk & 7)] = __float2half_rn(v);
    }
  } else {
    float* tail = reinterpret_cast<float*>(out + blob_total);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < T_TOTAL; idx += gridDim.x * blockDim.x) {
      float v = 0.f;
      if (idx < T_WSIG) { int m = idx / 256, c = idx % 256; int pl = param_layer(m); v = c < layer_out(pl) ? p.b[pl][c] : 0.f; }
      else if (idx < T_WRGB2) v = p.w[L_SIGMA][idx - T_WSIG];
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

size_t npp_tc_packed_bytes(bool bg) { return (size_t)tc::h_tab[bg].total + tc::T_TOTAL * sizeof(float); }

int npp_pack_tc(const NerfppNetParams* p, bool bg, void* out, cudaStream_t st) {
  const tc::StepTable& t = tc::h_tab[bg];
  tc::pack_tc_kernel<<<dim3(8, t.n + 1), 256, 0, st>>>(*p, bg, (uint8_t*)out, t.total);
  NPP_CHECK_LAUNCH();
  return 0;
}

static int g_cluster = -1;      // weight-sharing cluster size; NERFPP_TC_CLUSTER overrides (1, 2 or 4)
static long long* g_dbg = nullptr;

template <bool BG, int CLUSTER>
static int launch_tc(int max_ctas, const uint8_t* blobs, const float* tail, const float* ray_o, const float* ray_d, const float* z,
                     int n, int S, float* out_sigma, float* out_rgb, float* out_dr, int num_tiles, cudaStream_t st) {
  auto kern = tc::field_tc_kernel<BG, CLUSTER>;
  static bool configured = false;
  static int max_clusters = 0;
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, blobs, tail, ray_o, ray_d, z, n, S, out_sigma, out_rgb, out_dr, num_tiles, g_dbg);
  if (e != cudaSuccess) { npp_set_error("field_tc launch (cluster %d): %s", CLUSTER, cudaGetErrorString(e)); return (int)e; }
  return 0;
}

// debug hook (tests/diag only): device buffer of 4 x gridDim.x int64 receiving the MMA thread's cycle counters
extern "C" void nerfpp_debug_set_tc_timers(long long* dev_buf) { g_dbg = dev_buf; }
extern "C" void nerfpp_debug_set_tc_cluster(int c) { g_cluster = c; }

int npp_field_tc(const void* packed, bool bg, const float* ray_o, const float* ray_d, const float* z, int n, int S,
                 float* out_sigma, float* out_rgb, float* out_depth_real, cudaStream_t st) {
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (g_cluster < 0) {
    const char* e = getenv("NERFPP_TC_CLUSTER");
    g_cluster = e ? atoi(e) : 2;
    if (g_cluster != 1 && g_cluster != 2 && g_cluster != 4) g_cluster = 2;
  }
  const long long total = (long long)n * S;
  const int num_tiles = (int)((total + tc::TILE - 1) / tc::TILE);
  const uint8_t* blobs = (const uint8_t*)packed;
  const float* tail = (const float*)(blobs + tc::h_tab[bg].total);
#define NPP_TC_LAUNCH(BG, C) launch_tc<BG, C>(num_sms, blobs, tail, ray_o, ray_d, z, n, S, out_sigma, out_rgb, out_depth_real, num_tiles, st)
  if (bg) return g_cluster == 1 ? NPP_TC_LAUNCH(true, 1) : g_cluster == 2 ? NPP_TC_LAUNCH(true, 2) : NPP_TC_LAUNCH(true, 4);
  return g_cluster == 1 ? NPP_TC_LAUNCH(false, 1) : g_cluster == 2 ? NPP_TC_LAUNCH(false, 2) : NPP_TC_LAUNCH(false, 4);
#undef NPP_TC_LAUNCH
}
