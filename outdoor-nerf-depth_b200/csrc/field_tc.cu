// Field evaluation on the 5th-generation tensor cores.
//
// One persistent CTA per SM walks 128-sample tiles (samples of all rays packed along the MMA M
// axis).  Per tile the ten dense layers (base 0..7, base_remap, rgb.0) run as tcgen05.mma
// (M=128, N=256|128, K=16, fp16 operands, fp32 accumulators in TMEM); sigma (256->1) and rgb.2
// (128->3) are fp32 dot products in the epilogue on the fp32 accumulators.
//
//   warp 0-7   epilogue: per layer TMEM -> regs, +bias, ReLU, fp16 pack -> written back to TMEM in
//              place as the A operand of the next layer, 64 columns at a time
//   warp 8     one lane issues every tcgen05.mma (A from TMEM, B from shared memory); layer l+1's
//              K-chunk j starts as soon as the epilogue of layer l has produced columns [64j,64j+64)
//              (two TMEM accumulators ping-pong between consecutive layers)
//   warp 9     one lane streams the pre-swizzled weight tiles (32 KB each, in MMA issue order) from L2
//              into a 5-deep shared-memory ring with cp.async.bulk (+cluster multicast) and mbarriers
//   warp 10-11 positional encoding of the NEXT tile into the double-buffered E operand (fp16, SW128)
//
// Shared memory (bytes): E 2x32K | ring 5x32K | barriers.  TMEM: 2 x 256 columns.
// Precision: operands are rounded to fp16 (11-bit significand), products/sums are fp32; measured
// against the fp32 reference: rgb/depth within 3e-5 relative (tests/test_parity_gpu.py).
#include <cuda_fp16.h>
#include <cstdlib>
#include "common.cuh"

namespace npp {
namespace tc {

constexpr int TILE = 128;
constexpr int NSTAGE = 5;
constexpr int STAGE_BYTES = 32768;
constexpr int CHUNK_BYTES = 16384;          // 128 rows x 64 fp16
constexpr int E_BYTES = 2 * CHUNK_BYTES;    // 128 columns: [0,emb) position, [96,123) view dir; double-buffered
constexpr int OFF_E = 0, OFF_W = 2 * E_BYTES, OFF_BAR = OFF_W + NSTAGE * STAGE_BYTES, OFF_SCRATCH = OFF_BAR + 256;
constexpr int SMEM_BYTES = OFF_SCRATCH + 128 * 16;
constexpr int VIEW_COL = 96;
constexpr int NUM_EPI_WARPS = 8, MMA_WARP = 8, LOAD_WARP = 9, EMB_WARP0 = 10, NUM_EMB_WARPS = 2, THREADS = 384;
constexpr int NUM_MMA_LAYERS = 10;          // base 0..7, remap, rgb0

// barrier slots
enum { B_WFULL = 0, B_WEMPTY = NSTAGE, B_AREADY = 2 * NSTAGE, B_EFULL = 2 * NSTAGE + 4, B_EEMPTY = 2 * NSTAGE + 6,
       B_ACC = 2 * NSTAGE + 8, B_COUNT = 2 * NSTAGE + 10 };

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
// D[tmem] (+)= A[tmem] * B[smem]^T: A = fp16 pairs packed in 32-bit TMEM columns, row = lane
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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

// Encodes one row (sample) of the E operand: columns [0,emb) = Embedder(pts) (nerf_network.py:42-60),
// columns [96,123) = Embedder(viewdir); fp16, 128B-swizzled.  Deliberately a compact loop around ONE
// sincosf call site: this runs on the two producer warps, off the critical path, and the kernel's
// instruction footprint matters more (I-cache) than its speed.
__device__ __noinline__ void embed_vec(const float* x, int dim, int nfreq, uint8_t* region, int row, int col_base) {
  for (int c = 0; c < dim; ++c) *reinterpret_cast<__half*>(region + sw128_off(row, col_base + c)) = __float2half_rn(x[c]);
#pragma unroll 1
  for (int i = 0; i < nfreq * dim; ++i) {
    const int k = i / dim, c = i - k * dim;
    float sn, cs;
    sincosf(x[c] * (float)(1 << k), &sn, &cs);
    const int col = col_base + dim + 2 * k * dim + c;
    *reinterpret_cast<__half*>(region + sw128_off(row, col)) = __float2half_rn(sn);
    *reinterpret_cast<__half*>(region + sw128_off(row, col + dim)) = __float2half_rn(cs);
  }
}

// CLUSTER > 1: the CTAs of a cluster walk their tiles in lock step and share every weight tile: each
// CTA fetches 1/CLUSTER of it from L2 and multicasts that slice into all CLUSTER rings.
template <bool BG, int CLUSTER>
__global__ void __launch_bounds__(THREADS, 1)
field_tc_kernel(const uint8_t* __restrict__ blobs, const float* __restrict__ tail, const float* __restrict__ ray_o,
                const float* __restrict__ ray_d, const float* __restrict__ z, int n, int S, float* __restrict__ out_sigma,
                float* __restrict__ out_rgb, float* __restrict__ out_depth_real, int num_tiles, long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int D = BG ? 4 : 3;
  constexpr int E_CHUNKS = BG ? 2 : 1;
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
  const bool timing = dbg != nullptr;

  if (warp == MMA_WARP && lane == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar(B_WFULL + i), 1); mbar_init(bar(B_WEMPTY + i), CLUSTER); }
    for (int i = 0; i < 4; ++i) mbar_init(bar(B_AREADY + i), NUM_EPI_WARPS);
    for (int i = 0; i < 2; ++i) { mbar_init(bar(B_EFULL + i), NUM_EMB_WARPS); mbar_init(bar(B_EEMPTY + i), 1); }
    mbar_init(bar(B_ACC), 1); mbar_init(bar(B_ACC + 1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == LOAD_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void*)tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp < NUM_EPI_WARPS) {   // zero both E buffers once: padding columns are never written again
    for (int i = threadIdx.x; i < 2 * E_BYTES / 16; i += NUM_EPI_WARPS * 32) reinterpret_cast<uint4*>(smem + OFF_E)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
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
    // Issue is the critical resource: UTCHMMA issue blocks for about the execution time of the MMAs
    // ahead of it, so everything this thread does between two MMAs is tensor-pipe idle time.
    if (lane == 0) {
      uint32_t it = 0, a_par = 0, tile_i = 0;
      long long t_e = 0, t_a = 0, t_w = 0, t0 = clock64(), tt = 0;
      constexpr uint32_t ID256 = idesc_f16(256), ID128 = idesc_f16(128);
      // one ring stage of weights against an operand in shared memory (E region)
      auto step_ss = [&](uint32_t a_smem, int k0, int nk, uint32_t d_tmem, uint32_t idesc, bool first) {
        const uint32_t st = it % NSTAGE, ph = (it / NSTAGE) & 1;
        if (timing) tt = clock64();
        mbar_wait(bar(B_WFULL + st), ph);
        if (timing) t_w += clock64() - tt;
        tc_fence_after();
        const uint32_t b_addr = s_base + OFF_W + st * STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k >= k0 && k < k0 + nk)
            umma_f16(d_tmem, sw128_desc(a_smem + k * 32), sw128_desc(b_addr + k * 32), idesc, (first && k == k0) ? 0u : 1u);
        if (CLUSTER == 1) tc_commit(bar(B_WEMPTY + st)); else tc_commit_mcast(bar(B_WEMPTY + st), kMask);
        ++it;
      };
      // one ring stage against operand chunk c held in TMEM (the previous layer's accumulator, fp16 in place)
      auto step_ts = [&](uint32_t a_tmem, int c, uint32_t d_tmem, uint32_t idesc, bool first) {
        const uint32_t st = it % NSTAGE, ph = (it / NSTAGE) & 1;
        if (timing) tt = clock64();
        mbar_wait(bar(B_AREADY + c), a_par);
        if (timing) { long long t1 = clock64(); t_a += t1 - tt; tt = t1; }
        mbar_wait(bar(B_WFULL + st), ph);
        if (timing) t_w += clock64() - tt;
        tc_fence_after();
        const uint32_t b_addr = s_base + OFF_W + st * STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ts(d_tmem, a_tmem + 64u * c + 32u * (k >> 1) + 8u * (k & 1), sw128_desc(b_addr + k * 32), idesc, (first && k == 0) ? 0u : 1u);
        if (CLUSTER == 1) tc_commit(bar(B_WEMPTY + st)); else tc_commit_mcast(bar(B_WEMPTY + st), kMask);
        ++it;
      };
      for (int grp = group0; grp < n_groups; grp += group_step, ++tile_i) {
        const uint32_t eb = tile_i & 1;
        const uint32_t e_addr = s_base + OFF_E + eb * E_BYTES;
        if (timing) tt = clock64();
        mbar_wait(bar(B_EFULL + eb), (tile_i >> 1) & 1);
        if (timing) t_e += clock64() - tt;
#pragma unroll 1
        for (int m = 0; m < NUM_MMA_LAYERS; ++m) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(m & 1) * 256u;
          const uint32_t a_tmem = tmem_base + (uint32_t)((m - 1) & 1) * 256u;
          const uint32_t idesc = (m == 9) ? ID128 : ID256;
          bool first = true;
          if (m == 0 || m == 5) {
#pragma unroll 1
            for (int c = 0; c < E_CHUNKS; ++c) { step_ss(e_addr + c * CHUNK_BYTES, 0, (BG && c == 1) ? 2 : 4, d_tmem, idesc, first); first = false; }
          }
          if (m == 9) {   // view-direction columns; last reader of this tile's E buffer
            step_ss(e_addr + CHUNK_BYTES, 2, 2, d_tmem, idesc, first); first = false;
            tc_commit(bar(B_EEMPTY + eb));
          }
          if (m != 0) {
#pragma unroll 1
            for (int c = 0; c < 4; ++c) { step_ts(a_tmem, c, d_tmem, idesc, first); first = false; }
            a_par ^= 1;
          }
          tc_commit(bar(B_ACC + (m & 1)));
        }
      }
      if (timing) { dbg[8 * blockIdx.x] = clock64() - t0; dbg[8 * blockIdx.x + 1] = t_e; dbg[8 * blockIdx.x + 2] = t_a; dbg[8 * blockIdx.x + 3] = t_w; }
    }
  } else if (warp >= EMB_WARP0) {
    // ================= embedding producers: E operand of the NEXT tile while the current one runs ======
    const int et = threadIdx.x - EMB_WARP0 * 32;    // 0..63, two rows each
    uint32_t tile_i = 0;
    for (int grp = group0; grp < n_groups; grp += group_step, ++tile_i) {
      const uint32_t eb = tile_i & 1;
      const int tile = grp * CLUSTER + (int)cta_rank;
      mbar_wait(bar(B_EEMPTY + eb), ((tile_i >> 1) & 1) ^ 1);
      uint8_t* sE = smem + OFF_E + eb * E_BYTES;
#pragma unroll 1
      for (int rr = 0; rr < 2; ++rr) {
        const int row = et + 64 * rr;
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
        embed_vec(x, D, NF_POS, sE, row, 0);
        embed_vec(vd, 3, NF_VIEW, sE, row, VIEW_COL);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_EFULL + eb));
    }
  } else {
    // ================= epilogue warps =================
    // Thread = one accumulator row (TMEM lane); warps w and w+4 split each 64-column chunk.  The fp16
    // activations are written back IN PLACE over the first half of the fp32 columns just read
    // (chunk j, half hh: K values [32hh,32hh+32) -> TMEM columns [64j+32hh, 64j+32hh+16)), where the
    // next layer's MMAs read them as the A operand: no shared-memory round trip.
    const int q = warp & 3, hh = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t acc_par[2] = {0, 0};
    long long e_wait = 0, e_t0 = clock64(), ett = 0;
    for (int grp = group0; grp < n_groups; grp += group_step) {
      const int tile = grp * CLUSTER + (int)cta_rank;
      const long long g = (long long)tile * TILE + row;
      const bool valid = g < total;
      float sig_part = 0.f, rgb_part[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
      for (int m = 0; m < NUM_MMA_LAYERS; ++m) {
        const int ab = m & 1;
        if (timing) ett = clock64();
        mbar_wait(bar(B_ACC + ab), acc_par[ab]);
        if (timing) e_wait += clock64() - ett;
        acc_par[ab] ^= 1;
        tc_fence_after();
        const float* bias = tail + T_BIAS + m * 256 + 32 * hh;
        const uint32_t acc_addr = lane_addr + (uint32_t)(ab * 256 + 32 * hh);
        const int nchunk = (m == 9) ? 2 : 4;
        uint32_t v[2][32];
        tmem_ld32(acc_addr, v[0]);
#pragma unroll
        for (int jc = 0; jc < 4; ++jc) {
          if (jc < nchunk) {
            uint32_t (&cur)[32] = v[jc & 1];
            float4 b4[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) b4[t] = __ldg(reinterpret_cast<const float4*>(bias + 64 * jc) + t);
            tmem_ld_wait();
            if (jc + 1 < nchunk) tmem_ld32(acc_addr + 64u * (jc + 1), v[(jc + 1) & 1]);   // overlaps the maths below
            float f[32];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              f[4 * t + 0] = __uint_as_float(cur[4 * t + 0]) + b4[t].x;
              f[4 * t + 1] = __uint_as_float(cur[4 * t + 1]) + b4[t].y;
              f[4 * t + 2] = __uint_as_float(cur[4 * t + 2]) + b4[t].z;
              f[4 * t + 3] = __uint_as_float(cur[4 * t + 3]) + b4[t].w;
            }
            if (m != 8) {
#pragma unroll
              for (int t = 0; t < 32; ++t) f[t] = fmaxf(f[t], 0.f);
            }
            if (m == 7) {   // sigma head on the fp32 activations, nerf_network.py:133
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                float4 w4 = __ldg(reinterpret_cast<const float4*>(tail + T_WSIG + 64 * jc + 32 * hh) + t);
                sig_part = fmaf(f[4 * t], w4.x, sig_part); sig_part = fmaf(f[4 * t + 1], w4.y, sig_part);
                sig_part = fmaf(f[4 * t + 2], w4.z, sig_part); sig_part = fmaf(f[4 * t + 3], w4.w, sig_part);
              }
            }
            if (m == 9) {   // rgb.2 on the fp32 hidden colour features, nerf_network.py:114-117
#pragma unroll
              for (int c = 0; c < 3; ++c) {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                  float4 w4 = __ldg(reinterpret_cast<const float4*>(tail + T_WRGB2 + c * RGB_HID + 64 * jc + 32 * hh) + t);
                  rgb_part[c] = fmaf(f[4 * t], w4.x, rgb_part[c]); rgb_part[c] = fmaf(f[4 * t + 1], w4.y, rgb_part[c]);
                  rgb_part[c] = fmaf(f[4 * t + 2], w4.z, rgb_part[c]); rgb_part[c] = fmaf(f[4 * t + 3], w4.w, rgb_part[c]);
                }
              }
            } else {        // next layer's A operand: 32 fp16 = 16 packed columns, in place
              uint32_t pk[16];
#pragma unroll
              for (int t = 0; t < 16; ++t) {
                __half2 h = __floats2half2_rn(f[2 * t], f[2 * t + 1]);
                pk[t] = *reinterpret_cast<uint32_t*>(&h);
              }
              tmem_st16(acc_addr + 64u * jc, pk);
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar(B_AREADY + jc));
            }
          }
        }
      }
      // ---- combine the two column halves of each row, write sigma / rgb -------------------------------
      float4* scratch = reinterpret_cast<float4*>(smem + OFF_SCRATCH);
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
    if (timing && threadIdx.x == 0) { dbg[8 * blockIdx.x + 5] = clock64() - e_t0; dbg[8 * blockIdx.x + 6] = e_wait; }
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
