// Backward of the field MLP as ONE kernel: the data-gradient chain (producers) and the weight-gradient GEMMs (consumers)
// run at the same time on disjoint sets of SMs, and the gradients dZ that connect them never reach HBM.
//
// In the split form (field_bwd_tc.cu then wgrad_tc.cu) every dZ_l -- 0.6 MB per 128-sample tile, 10.5 GB per trainer
// step -- is written to HBM by one kernel and read back by the next, and the weight-gradient kernel is a memcpy with
// MMAs attached (HBM-bound, 21 GB of operands per step).  Here:
//
//   producer CTAs [0, P)      the dgrad machine of field_bwd_tc.cu, tile by tile.  Each dZ_l leaves the epilogue through
//                             a 4 x 16 KB shared-memory staging ring and a dedicated STORER warp, which bulk-copies it
//                             into this producer's slot ring in global memory (9 x 64 KB + 2 x 32 KB per producer,
//                             rewritten in place once per tile: it lives in L2) and publishes a counter.
//   consumer CTAs [P, P + C)  the split-K GEMMs of wgrad_tc.cu: one job (layer) and a subset of the producers per CTA.
//                             The loader polls the producers' counters, pulls dZ from the slot ring (L2) and the saved
//                             forward activation X from HBM, the MMAs accumulate dW in TMEM over ALL tiles of the subset,
//                             one flush with red.global.add at the end.  When a slot has been copied to shared memory
//                             the consumer bumps its own counter, which the producer's storer checks before reusing the slot.
//
// All CTAs must be resident at once (a producer waits for consumers and vice versa): the launch is cooperative, so an
// oversubscribed grid is refused by the driver instead of deadlocking, and every global spin loop carries a watchdog
// that traps.  HBM traffic per step: the saved activations, read once (11 GB) -- dZ (10.5 GB written + 10.5 GB read) is gone.
#include <cstdlib>
#include "bwd_common.cuh"

namespace npp {
namespace tcf {
using namespace npp::tc;
using tcb::NSTAGE;
using tcb::STAGE_BYTES;
using tcb::NUM_LAYERS;
using tcw::Job;
using tcw::JobTable;
using tcw::MAX_JOBS;
using tcw::MN_HI;
using tcw::mn_lo;
using tcw::idesc_mn;

constexpr int THREADS = 480;   // producer: warps 0-7 epilogue, 8 MMA, 9 weight loader, 10-13 prologue, 14 storer
                               // consumer: warps 0-3 epilogue / bias sums, 4 MMA, 5 loader (the rest idle)
// ---- slot ring and counters in global memory -------------------------------------------------------------------------
constexpr int DEPTH = NUM_LAYERS;                         // chain slots per producer: slot = chain step t, reused every tile
constexpr int UNIT_BYTES = 4 * CHUNK_BYTES;               // one dZ_l of one tile: [128 x 256] fp16
constexpr int DG_UNIT_BYTES = 2 * CHUNK_BYTES;            // dG (gradient of the rgb hidden layer): [128 x 128]
constexpr int DG_DEPTH = 2;
constexpr size_t RING_BYTES = (size_t)DEPTH * UNIT_BYTES + (size_t)DG_DEPTH * DG_UNIT_BYTES;   // per producer, 640 KB
constexpr int SYNC_WORDS = 32;                            // per producer: one 128-byte line
enum { S_UNITS = 0, S_DG = 1, S_DONE = 2 };               // chain units published | dG units published | S_DONE + job: units consumed
constexpr int MAX_CTAS = 160;
struct Roles {                                            // who does what (host-computed, launch parameter)
  int P;                                                  // producers
  short job[MAX_CTAS];                                    // consumer c = blockIdx.x - P: its job ...
  short split[MAX_CTAS];                                  // ... and which of the job's K[job] producer subsets (p % K == split)
  short K[MAX_JOBS];
};

// ---- producer shared memory ------------------------------------------------------------------------------------------
constexpr int G_BYTES = tcb::G_BYTES;                     // dG operand, SINGLE buffer: it is free again once layer t = 0 has run
constexpr int NSTG = 4;                                   // staging ring: whole 16 KB chunk images
constexpr int P_OFF_G = 0, P_OFF_W = G_BYTES, P_OFF_STG = P_OFF_W + NSTAGE * STAGE_BYTES, P_OFF_BAR = P_OFF_STG + NSTG * CHUNK_BYTES,
              P_OFF_W2 = P_OFF_BAR + 512, P_SMEM = P_OFF_W2 + 3 * RGB_HID * 4;
enum { B_WFULL = 0, B_WEMPTY = NSTAGE, B_AREADY = 2 * NSTAGE, B_GFULL = B_AREADY + 4, B_GEMPTY = B_GFULL + 1, B_ACC = B_GEMPTY + 1,
       B_SFULL = B_ACC + 2, B_SEMPTY = B_SFULL + NSTG, B_PCOUNT = B_SEMPTY + NSTG };
static_assert(8 * B_PCOUNT + 8 <= 512 && P_OFF_STG % 1024 == 0 && P_SMEM <= 232448, "producer shared memory");
constexpr int NUM_EPI_WARPS = 8, MMA_WARP = 8, LOAD_WARP = 9, PRO_WARP0 = 10, NUM_PRO_WARPS = 4, STORE_WARP = 14;
// ---- consumer shared memory ------------------------------------------------------------------------------------------
constexpr int X_BYTES = 4 * CHUNK_BYTES, AH_BYTES = 2 * CHUNK_BYTES;
constexpr int C_OFF_X = 0, C_OFF_A = 2 * X_BYTES, C_OFF_BAR = C_OFF_A + 2 * AH_BYTES, C_SMEM = C_OFF_BAR + 128;
enum { C_XFULL = 0, C_XEMPTY = 2, C_AFULL = 4, C_AEMPTY = 6, C_DONE = 8, C_COUNT = 9 };
constexpr int SMEM_BYTES = P_SMEM > C_SMEM ? P_SMEM : C_SMEM;

__constant__ tcb::StepTable c_tab = tcb::make_table();
__constant__ JobTable c_jobs[2] = {tcw::make_jobs(false), tcw::make_jobs(true)};
static const JobTable h_jobs[2] = {tcw::make_jobs(false), tcw::make_jobs(true)};

// ---- global-memory flags -----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// async-proxy accesses (bulk copies) on either side of a generic-proxy flag need a cross-proxy fence
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Spin until *flag >= need.  A producer and a consumer that wait for each other forever would hang the GPU: after ~2 s
// of polling the kernel traps (the launch fails with an error instead).
__device__ __forceinline__ void wait_flag(const uint32_t* flag, uint32_t need) {
  if (ld_acquire(flag) >= need) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (ld_acquire(flag) < need) {
    __nanosleep(100);
    if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void bulk_s2g_nocommit(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
// streamed-once operands (the saved activations): do not let them push the slot ring out of L2
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}

struct Args {
  // producer side (field_bwd_tc.cu's arguments)
  const uint8_t* blobs; const float* tail; const uint8_t* mask; const float* rgb; const float* raw_sigma; const float* d_sigma;
  const float* d_rgb; const float* scale_ptr; float* d_raw_sigma; float* d_raw_rgb;
  // consumer side (wgrad_tc.cu's)
  const uint8_t* act; const uint8_t* etiles; NerfppNetGrads grads;
  // shared
  uint8_t* ring; uint32_t* sync; long long total; int num_tiles; int bg;
  uint8_t* dz;        // non-null: producer-only launch -- every dZ goes to this ACT-layout buffer in HBM (tc_common.cuh), no
                      // slot ring, no counters; wgrad_tc_kernel reads it afterwards (the two-kernel form of the backward)
};

// ======================================================================================================================
// producer: field_dgrad_kernel's body with the dZ stores going to the slot ring through the storer warp
// ======================================================================================================================
__device__ __forceinline__ void producer(const Args& a, const Roles& roles, uint8_t* smem) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar0 = s_base + P_OFF_BAR;
  auto bar = [&](int i) { return bar0 + 8u * i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + P_OFF_BAR + 8 * B_PCOUNT);
  const float scale = a.scale_ptr[0], scale_rgb = a.scale_ptr[1];   // below the join | colour path (backward.cu)
  const float join = scale / scale_rgb;
  const tcb::StepTable& tab = c_tab;
  const int P = roles.P, p = blockIdx.x, num_tiles = a.num_tiles;
  uint8_t* const ring = a.ring + (size_t)p * RING_BYTES;
  uint32_t* const sync = a.sync + (size_t)p * SYNC_WORDS;
  const long long total = a.total;
  const float* __restrict__ tail = a.tail;

  if (warp == MMA_WARP && lane == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar(B_WFULL + i), 1); mbar_init(bar(B_WEMPTY + i), 1); }
    for (int i = 0; i < 4; ++i) mbar_init(bar(B_AREADY + i), NUM_EPI_WARPS);
    mbar_init(bar(B_GFULL), NUM_PRO_WARPS);
    mbar_init(bar(B_GEMPTY), 1);
    for (int i = 0; i < 2; ++i) mbar_init(bar(B_ACC + i), 1);
    for (int i = 0; i < NSTG; ++i) { mbar_init(bar(B_SFULL + i), NUM_EPI_WARPS); mbar_init(bar(B_SEMPTY + i), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == LOAD_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void*)tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 3 * RGB_HID; i += THREADS) reinterpret_cast<float*>(smem + P_OFF_W2)[i] = tail[tcb::T_W2 + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == LOAD_WARP) {
    // ================= weight loader =================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = p; tile < num_tiles; tile += P)
        for (int i = 0; i < tab.n; ++i, ++it) {
          const uint32_t st = it % NSTAGE, ph = (it / NSTAGE) & 1;
          mbar_wait(bar(B_WEMPTY + st), ph ^ 1);
          mbar_expect_tx(bar(B_WFULL + st), STAGE_BYTES);
          bulk_g2s(s_base + P_OFF_W + st * STAGE_BYTES, a.blobs + tab.s[i].blob_off, STAGE_BYTES, bar(B_WFULL + st));
        }
    }
  } else if (warp == MMA_WARP) {
    // ================= MMA issuer =================
    constexpr uint32_t ID256 = idesc_f16(256);
    const uint32_t ring0 = s_base + P_OFF_W, wfull0 = bar(B_WFULL), wempty0 = bar(B_WEMPTY);
    uint32_t st = 0, ph = 0, slot = ring0, wfull = wfull0, wempty = wempty0;
    uint32_t a_par = 0, tile_i = 0, acc_cnt[2] = {0, 0};
    auto advance = [&]() {
      ++st; slot += STAGE_BYTES; wfull += 8; wempty += 8;
      if (st == NSTAGE) { st = 0; ph ^= 1; slot = ring0; wfull = wfull0; wempty = wempty0; }
    };
    const uint32_t g_addr = s_base + P_OFF_G;
    for (int tile = p; tile < num_tiles; tile += P, ++tile_i) {
      mbar_wait(bar(B_GFULL), tile_i & 1);
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
              if (c == 1) { tc_commit(bar(B_GEMPTY)); tc_commit(bar(B_ACC + buf)); }
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
  } else if (warp >= PRO_WARP0 && warp < PRO_WARP0 + NUM_PRO_WARPS) {
    // ================= prologue producers: dG operand of the NEXT tile (shared memory) and its copy for the rgb.0 jobs =====
    const int row = threadIdx.x - PRO_WARP0 * 32;    // 0..127
    const float* w2 = reinterpret_cast<const float*>(smem + P_OFF_W2);
    const JobTable& jt = c_jobs[a.bg];
    uint32_t tile_i = 0;
    for (int tile = p; tile < num_tiles; tile += P, ++tile_i) {
      mbar_wait(bar(B_GEMPTY), (tile_i & 1) ^ 1);
      uint8_t* sG = smem + P_OFF_G;
      // the dG slot written two tiles ago must have been copied out by both rgb.0 jobs
      if (!a.dz) {
        if (tile_i >= DG_DEPTH && row == 0) {
          for (int j = 0; j < jt.n; ++j)
            if (jt.j[j].a_layer == 9) wait_flag(sync + S_DONE + j, tile_i - DG_DEPTH + 1);
        }
        asm volatile("bar.sync 6, 128;" ::: "memory");
      }
      uint8_t* const dslot = a.dz ? a.dz + act_chunk_off(9, (size_t)num_tiles, (size_t)tile, 0)
                                  : ring + (size_t)DEPTH * UNIT_BYTES + (size_t)(tile_i % DG_DEPTH) * DG_UNIT_BYTES;
      const long long g = (long long)tile * TILE + row;
      float dr[3] = {0.f, 0.f, 0.f};
      if (g < total) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { const float cc = a.rgb[3 * g + c]; dr[c] = a.d_rgb[3 * g + c] * cc * (1.f - cc) * scale_rgb; }   // sigmoid'
        a.d_raw_rgb[3 * g] = dr[0]; a.d_raw_rgb[3 * g + 1] = dr[1]; a.d_raw_rgb[3 * g + 2] = dr[2];
        const float rs = a.raw_sigma[g];
        a.d_raw_sigma[g] = a.d_sigma[g] * (rs > 0.f ? 1.f : rs < 0.f ? -1.f : 0.f) * scale;                                          // abs'
      }
      // dG[k] = (dr . W_rgb2[:,k]) * [g_k > 0], g = saved rgb hidden (ACT layer 9, 2 chunks)
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {
        const uint8_t* ach = a.act + act_chunk_off(9, (size_t)num_tiles, (size_t)tile, j);
        uint8_t* dch = dslot + (size_t)j * CHUNK_BYTES;
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
      if (lane == 0) mbar_arrive(bar(B_GFULL));
      // publish the global copy: every writer's stores are ordered before the barrier, the signalling thread's release
      // fence is cumulative over them
      if (!a.dz) {
        asm volatile("bar.sync 6, 128;" ::: "memory");
        if (row == 0) { __threadfence(); st_release(sync + S_DG, tile_i + 1); }
      }
    }
  } else if (warp == STORE_WARP) {
    // ================= storer: staging ring -> slot ring (global, L2-resident) + the published counter =================
    if (lane == 0) {
      const JobTable& jt = c_jobs[a.bg];
      const uint32_t n_tiles_mine = (uint32_t)((num_tiles - p + P - 1) / P);
      const uint32_t n_chunks = n_tiles_mine * NUM_LAYERS * 4;
      for (uint32_t c = 0; c < n_chunks; ++c) {
        const uint32_t k = c & 3u, seq = c >> 2, t = seq % NUM_LAYERS, tile_i = seq / NUM_LAYERS;
        if (k == 0 && tile_i > 0 && !a.dz) {
          // slot t still holds the previous tile's dZ until every job reading that layer has copied it to shared memory
          const int al = tcb::target_act((int)t);
          for (int j = 0; j < jt.n; ++j)
            if (jt.j[j].a_layer == al) wait_flag(sync + S_DONE + j, tile_i);
          fence_proxy_async_all();
        }
        const uint32_t sb = c % NSTG;
        mbar_wait(bar(B_SFULL + sb), (c / NSTG) & 1);
        uint8_t* const dst = a.dz ? a.dz + act_chunk_off(tcb::target_act((int)t), (size_t)num_tiles, (size_t)(p + (int)tile_i * P), (int)k)
                                  : ring + (size_t)t * UNIT_BYTES + (size_t)k * CHUNK_BYTES;
        bulk_s2g_nocommit(dst, s_base + P_OFF_STG + sb * CHUNK_BYTES, CHUNK_BYTES);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        // staging buffers are released two copies late (their shared-memory reads have finished by then without stalling us)
        asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
        if (c >= 2) mbar_arrive(bar(B_SEMPTY + ((c - 2) % NSTG)));
        if (k == 3 && !a.dz) {
          // ... and units are published one unit late: all copies but the last four have completed (are globally visible)
          asm volatile("cp.async.bulk.wait_group 4;" ::: "memory");
          if (seq > 0) { fence_proxy_async_all(); __threadfence(); st_release(sync + S_UNITS, seq); }
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      if (n_chunks > 0 && !a.dz) { fence_proxy_async_all(); __threadfence(); st_release(sync + S_UNITS, n_chunks >> 2); }
    }
  } else if (warp < NUM_EPI_WARPS) {
    // ================= epilogue warps =================
    const int q = warp & 3, hh = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t roff = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
    uint32_t acc_par = 0, tile_i = 0, cchunk = 0;
    for (int tile = p; tile < num_tiles; tile += P, ++tile_i) {
      const long long g = (long long)tile * TILE + row;
      float dsr = 0.f;
      if (g < total) { const float rs = a.raw_sigma[g]; dsr = a.d_sigma[g] * (rs > 0.f ? 1.f : rs < 0.f ? -1.f : 0.f) * scale; }
#pragma unroll 1
      for (int t = 0; t < NUM_LAYERS; ++t) {
        const uint32_t buf = (t + tile_i) & 1;
        const int l = tcb::target_act(t);
        uint4 mb4 = make_uint4(0u, 0u, 0u, 0u);
        if (t != 0) mb4 = __ldg(reinterpret_cast<const uint4*>(a.mask + mask_off(l, (size_t)num_tiles, (size_t)tile, hh, row)));
        const uint32_t mb[4] = {mb4.x, mb4.y, mb4.z, mb4.w};
        mbar_wait(bar(B_ACC + buf), (acc_par >> buf) & 1u);
        acc_par ^= 1u << buf;
        tc_fence_after();
        const uint32_t acc_addr = lane_addr + buf * 256u + 32u * hh;
        uint32_t v[2][32];
        tmem_ld32(acc_addr, v[0]);
#pragma unroll
        for (int j = 0; j < 4; ++j, ++cchunk) {
          uint32_t (&cur)[32] = v[j & 1];
          tmem_ld_wait(cur);
          if (j + 1 < 4) tmem_ld32(acc_addr + 64u * (j + 1), v[(j + 1) & 1]);
          if (t == 1) {       // the sigma head joins here: d h7 = (colour path, rescaled to the chain's loss scale) + d(raw sigma) * w_sigma   (nerf_network.py:133)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 w4 = __ldg(reinterpret_cast<const float4*>(tail + tcb::T_WSIG + 64 * j + 32 * hh) + e);
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
          // the copy for the weight-gradient consumers leaves behind the arrive, off the chain the tensor pipe waits for:
          // this thread's 64 bytes of the chunk image go into the staging ring, the storer warp does the rest
          const uint32_t sb = cchunk % NSTG;
          mbar_wait(bar(B_SEMPTY + sb), ((cchunk / NSTG) & 1u) ^ 1u);
          uint4* rowp = reinterpret_cast<uint4*>(smem + P_OFF_STG + sb * CHUNK_BYTES + roff);
#pragma unroll
          for (int u = 0; u < 4; ++u) rowp[(4 * hh + u) ^ (row & 7)] = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(B_SFULL + sb));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == LOAD_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ======================================================================================================================
// consumer: wgrad_tc_kernel's body over the units (producer p of its subset, that producer's r-th tile)
// ======================================================================================================================
__device__ __forceinline__ void consumer(const Args& a, const Roles& roles, uint8_t* smem) {
  __shared__ uint32_t tmem_slot_c;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t s_base = smem_u32(smem);
  auto bar = [&](int i) { return s_base + C_OFF_BAR + 8u * i; };
  const int P = roles.P, cidx = (int)blockIdx.x - P;
  const int job = roles.job[cidx], split = roles.split[cidx], K = roles.K[job];
  const Job jb = c_jobs[a.bg].j[job];
  const int num_tiles = a.num_tiles;
  const int N = 64 * jb.x_nchunks;
  const size_t nt = (size_t)num_tiles;
  const int R = (num_tiles + P - 1) / P;                    // tiles of the busiest producer
  // this CTA's units: for r = 0..R-1, for p = split, split + K, ... < P with p + r P < num_tiles
  int n_units = 0;
  for (int r = 0; r < R; ++r)
    for (int p = split; p < P && p + r * P < num_tiles; p += K) ++n_units;
  const bool chain = jb.a_layer != 9;
  const int t_chain = chain ? tcb::chain_step_of_act(jb.a_layer) : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(bar(C_XFULL + i), 1); mbar_init(bar(C_XEMPTY + i), 1); mbar_init(bar(C_AFULL + i), 1); mbar_init(bar(C_AEMPTY + i), jb.bias ? 5 : 1); }
    mbar_init(bar(C_DONE), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot_c)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot_c;

  if (warp == 5) {
    // ================= loader: wait for the producer, then X (HBM) and the dZ halves (slot ring, L2) =================
    if (lane == 0 && n_units > 0) {
      const uint64_t pol = policy_evict_first();
      uint32_t ix = 0, ia = 0;
      for (int r = 0; r < R; ++r)
        for (int p = split; p < P && p + r * P < num_tiles; p += K) {
          const int tile = p + r * P;
          const uint32_t xb = ix & 1;
          mbar_wait(bar(C_XEMPTY + xb), ((ix >> 1) & 1) ^ 1);
          const uint8_t* xsrc = jb.x_is_e ? a.etiles + ((size_t)tile * 2 + jb.x_chunk0) * CHUNK_BYTES
                                          : a.act + act_chunk_off(jb.x_layer, nt, (size_t)tile, jb.x_chunk0);
          mbar_expect_tx(bar(C_XFULL + xb), (uint32_t)(jb.x_nchunks * CHUNK_BYTES));
          bulk_g2s_hint(s_base + C_OFF_X + xb * X_BYTES, xsrc, (uint32_t)(jb.x_nchunks * CHUNK_BYTES), bar(C_XFULL + xb), pol);
          ++ix;
          // the producer has published this unit?
          const uint32_t* sync = a.sync + (size_t)p * SYNC_WORDS;
          const uint8_t* ring = a.ring + (size_t)p * RING_BYTES;
          const uint8_t* src;
          if (chain) { wait_flag(sync + S_UNITS, (uint32_t)(r * NUM_LAYERS + t_chain + 1)); src = ring + (size_t)t_chain * UNIT_BYTES; }
          else { wait_flag(sync + S_DG, (uint32_t)(r + 1)); src = ring + (size_t)DEPTH * UNIT_BYTES + (size_t)(r % DG_DEPTH) * DG_UNIT_BYTES; }
          fence_proxy_async_all();
          for (int h = 0; h < jb.m_halves; ++h, ++ia) {
            const uint32_t ab = ia & 1;
            mbar_wait(bar(C_AEMPTY + ab), ((ia >> 1) & 1) ^ 1);
            mbar_expect_tx(bar(C_AFULL + ab), AH_BYTES);
            bulk_g2s(s_base + C_OFF_A + ab * AH_BYTES, src + (size_t)h * AH_BYTES, AH_BYTES, bar(C_AFULL + ab));
          }
        }
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    const uint32_t idesc = idesc_mn(N);
    uint32_t ix = 0, ia = 0;
    int unit = 0;
    for (int r = 0; r < R; ++r)
      for (int p = split; p < P && p + r * P < num_tiles; p += K, ++ix, ++unit) {
        const uint32_t xb = ix & 1;
        mbar_wait(bar(C_XFULL + xb), (ix >> 1) & 1);
        for (int h = 0; h < jb.m_halves; ++h, ++ia) {
          const uint32_t ab = ia & 1;
          mbar_wait(bar(C_AFULL + ab), (ia >> 1) & 1);
          tc_fence_after();
          // the unit is in shared memory: its slot in the producer's ring may be overwritten
          if (h == jb.m_halves - 1 && lane == 0) st_release(a.sync + (size_t)p * SYNC_WORDS + S_DONE + job, (uint32_t)(r + 1));
          __syncwarp();
          if (elect_one()) {
            const uint32_t alo = mn_lo(s_base + C_OFF_A + ab * AH_BYTES), blo = mn_lo(s_base + C_OFF_X + xb * X_BYTES);
            const uint32_t d = tmem_base + 256u * (uint32_t)h;
#pragma unroll
            for (int k = 0; k < 8; ++k) {        // 16 samples per MMA = two 8-sample groups = 2048 B
              if (k == 0 && unit == 0) mma_ss<0>(d, alo, MN_HI, blo, MN_HI, idesc);
              else mma_ss<1>(d, alo + 128u * k, MN_HI, blo + 128u * k, MN_HI, idesc);
            }
            tc_commit(bar(C_AEMPTY + ab));
            if (h == jb.m_halves - 1) tc_commit(bar(C_XEMPTY + xb));
            if (unit == n_units - 1 && h == jb.m_halves - 1) tc_commit(bar(C_DONE));
          }
          __syncwarp();
        }
      }
  } else if (warp < 4) {
    // ================= epilogue: bias column sums while the MMAs run, then TMEM -> red.global.add into dW =================
    if (jb.bias && n_units > 0) {
      const int t = threadIdx.x, pp = t & 63, rh = t >> 6;
      const int col = 2 * (pp & 31);
      const uint32_t coff = (uint32_t)((pp >> 5) * CHUNK_BYTES + (col & 7) * 2);
      float bsum[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
      uint32_t ia = 0;
      for (int unit = 0; unit < n_units; ++unit) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h < jb.m_halves) {
            const uint32_t ab = ia & 1;
            mbar_wait(bar(C_AFULL + ab), (ia >> 1) & 1);
            const uint8_t* ah = smem + C_OFF_A + ab * AH_BYTES + coff;
#pragma unroll 8
            for (int k = 0; k < 64; ++k) {
              const int rr = rh * 64 + k;
              const __half2 v = *reinterpret_cast<const __half2*>(ah + (rr >> 3) * 1024 + (rr & 7) * 128 + (((col >> 3) ^ (rr & 7)) << 4));
              const float2 f = __half22float2(v);
              bsum[h][0] += f.x;
              bsum[h][1] += f.y;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(C_AEMPTY + ab));
            ++ia;
          }
        }
      }
      const float inv_scale_b = 1.f / a.scale_ptr[jb.a_layer >= 8 ? 1 : 0];
      float* db = a.grads.b[jb.a_layer < 8 ? jb.a_layer : jb.a_layer == 8 ? L_REMAP : L_RGB0];
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (h < jb.m_halves) {
          const int f0 = 128 * h + 64 * (pp >> 5) + col;
          atomicAdd(db + f0, bsum[h][0] * inv_scale_b);
          atomicAdd(db + f0 + 1, bsum[h][1] * inv_scale_b);
        }
    }
    if (n_units > 0) {
      mbar_wait(bar(C_DONE), 0);
      tc_fence_after();
      const float inv_scale = 1.f / a.scale_ptr[jb.a_layer >= 8 ? 1 : 0];     // colour-path layers carry scale[1] (backward.cu)
      float* dW = a.grads.w[jb.w_index];
      const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
      for (int h = 0; h < jb.m_halves; ++h) {
        const int orow = 128 * h + warp * 32 + lane;
        float* wrow = dW + (size_t)orow * jb.ld + jb.col0;
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(lane_addr + 256u * (uint32_t)h + (uint32_t)c0, v);
          tmem_ld_wait(v);
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int col = c0 + e - jb.skip;
            if (col >= 0 && col < jb.ncols) atomicAdd(wrow + col, __uint_as_float(v[e]) * inv_scale);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

__global__ void __launch_bounds__(THREADS, 1)
bwd_fused_kernel(const __grid_constant__ Args a, const __grid_constant__ Roles roles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if ((int)blockIdx.x < roles.P) producer(a, roles, smem);
  else consumer(a, roles, smem);
}

}  // namespace tcf
}  // namespace npp

using namespace npp;

// ---- host side ---------------------------------------------------------------------------------------------------------
static int g_cons_permille = 480;      // share of the grid given to the weight-gradient consumers (tests/diag may change it)
extern "C" void nerfpp_debug_set_bwd_consumers(int permille) { if (permille > 0 && permille < 1000) g_cons_permille = permille; }

static int fused_ctas(int dev) {
  static int sms[64] = {0};
  if (dev < 0 || dev >= 64) dev = 0;
  if (sms[dev] == 0) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
  return sms[dev] < tcf::MAX_CTAS ? sms[dev] : tcf::MAX_CTAS;
}

// Splits `ctas` CTAs into producers and consumers and the consumers among the jobs in proportion to what a job moves
// per unit (its operand bytes bound it, not its MMAs: 64-128 KB of shared-memory fill per 256-2048 tensor-pipe cycles).
static void plan_roles(bool bg, int ctas, int num_tiles, tcf::Roles* ro) {
  const tcw::JobTable& jt = tcf::h_jobs[bg];
  double w[tcw::MAX_JOBS], wsum = 0.0;
  for (int j = 0; j < jt.n; ++j) {
    const tcw::Job& jb = jt.j[j];
    const double bytes = (jb.m_halves * 2 + jb.x_nchunks) * (double)tc::CHUNK_BYTES;
    const double mma = jb.m_halves * 8 * 128.0 * (64.0 * jb.x_nchunks / 256.0);
    w[j] = bytes / 56.0 > mma ? bytes / 56.0 : mma;
    wsum += w[j];
  }
  int C = (int)(ctas * (g_cons_permille / 1000.0) + 0.5);
  if (C < jt.n) C = jt.n;
  int P = ctas - C;
  if (P > num_tiles) P = num_tiles;
  if (P < 1) { P = 1; C = ctas - 1; }
  // when there are few tiles, fewer consumers per job than producers is all that can be used
  int used = 0;
  for (int j = 0; j < jt.n; ++j) {
    int k = (int)(C * w[j] / wsum);
    if (k < 1) k = 1;
    if (k > P) k = P;
    ro->K[j] = (short)k;
    used += k;
  }
  for (int guard = 0; used < C && guard < 4 * C; ++guard) {       // hand out what the rounding left, largest load per CTA first
    int best = -1;
    double load = 0.0;
    for (int j = 0; j < jt.n; ++j)
      if (ro->K[j] < P && w[j] / ro->K[j] > load) { load = w[j] / ro->K[j]; best = j; }
    if (best < 0) break;
    ++ro->K[best];
    ++used;
  }
  while (used > C) {                                               // (only when C was raised to one CTA per job)
    int best = -1;
    double load = 1e300;
    for (int j = 0; j < jt.n; ++j)
      if (ro->K[j] > 1 && w[j] / ro->K[j] < load) { load = w[j] / ro->K[j]; best = j; }
    if (best < 0) break;
    --ro->K[best];
    --used;
  }
  ro->P = P;
  int c = 0;
  for (int j = 0; j < jt.n; ++j)
    for (int s = 0; s < ro->K[j]; ++s, ++c) { ro->job[c] = (short)j; ro->split[c] = (short)s; }
  for (int j = jt.n; j < tcw::MAX_JOBS; ++j) ro->K[j] = 0;
}

// The producer role alone on every SM, writing dZ to an ACT-layout buffer in HBM: the data-gradient kernel of the
// two-kernel backward (same arguments as field_bwd_tc.cu's npp_field_dgrad, which it supersedes: 64 KB of staging and a
// storer warp instead of 32 KB and a barrier per chunk in the epilogue warps).
int npp_field_dgrad_v2(const void* packed, size_t packed_blob_bytes, bool bg, const void* act, const void* mask, const float* rgb, const float* raw_sigma,
                       const float* d_sigma, const float* d_rgb, const float* scale, long long total, void* dz, float* d_raw_sigma,
                       float* d_raw_rgb, cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  static bool configured[64] = {false};
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(tcf::bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcf::SMEM_BYTES);
    if (e != cudaSuccess) { npp_set_error("dgrad: shared memory opt-in: %s", cudaGetErrorString(e)); return (int)e; }
    configured[dev & 63] = true;
  }
  const int num_tiles = (int)((total + tc::TILE - 1) / tc::TILE);
  const int ctas = fused_ctas(dev);
  tcf::Roles ro{};
  ro.P = num_tiles < ctas ? num_tiles : ctas;
  tcf::Args a{};
  const uint8_t* blobs = (const uint8_t*)packed;
  a.blobs = blobs; a.tail = (const float*)(blobs + packed_blob_bytes); a.mask = (const uint8_t*)mask; a.rgb = rgb; a.raw_sigma = raw_sigma;
  a.d_sigma = d_sigma; a.d_rgb = d_rgb; a.scale_ptr = scale; a.d_raw_sigma = d_raw_sigma; a.d_raw_rgb = d_raw_rgb;
  a.act = (const uint8_t*)act; a.total = total; a.num_tiles = num_tiles; a.bg = bg ? 1 : 0; a.dz = (uint8_t*)dz;
  tcf::bwd_fused_kernel<<<ro.P, tcf::THREADS, tcf::SMEM_BYTES, st>>>(a, ro);
  NPP_CHECK_LAUNCH();
  return 0;
}

size_t npp_bwd_fused_ws_bytes(int dev) {
  const int ctas = fused_ctas(dev);
  return (size_t)ctas * tcf::RING_BYTES + (size_t)ctas * tcf::SYNC_WORDS * sizeof(uint32_t) + 2048;
}

// One net's parameter gradients (+=) from the saved forward state and the gradients entering the MLP.  `ws`: workspace of
// npp_bwd_fused_ws_bytes() bytes, 1024-aligned.  `packed`: npp_pack_dgrad's buffer.
int npp_field_bwd_fused(bool bg, const void* packed, size_t packed_blob_bytes, const void* act, const void* etiles, const void* mask, const float* rgb,
                        const float* raw_sigma, const float* d_sigma, const float* d_rgb, const float* scale, long long total,
                        float* d_raw_sigma, float* d_raw_rgb, const NerfppNetGrads* grads, void* ws, cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  static bool configured[64] = {false};
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(tcf::bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcf::SMEM_BYTES);
    if (e != cudaSuccess) { npp_set_error("bwd_fused: shared memory opt-in: %s", cudaGetErrorString(e)); return (int)e; }
    configured[dev & 63] = true;
  }
  const int num_tiles = (int)((total + tc::TILE - 1) / tc::TILE);
  const int ctas = fused_ctas(dev);
  tcf::Roles ro{};
  plan_roles(bg, ctas, num_tiles, &ro);
  int C = 0;
  for (int j = 0; j < tcf::h_jobs[bg].n; ++j) C += ro.K[j];
  const int grid = ro.P + C;
  uint8_t* w = (uint8_t*)ws;
  tcf::Args a{};
  const uint8_t* blobs = (const uint8_t*)packed;
  a.blobs = blobs; a.tail = (const float*)(blobs + packed_blob_bytes); a.mask = (const uint8_t*)mask; a.rgb = rgb; a.raw_sigma = raw_sigma;
  a.d_sigma = d_sigma; a.d_rgb = d_rgb; a.scale_ptr = scale; a.d_raw_sigma = d_raw_sigma; a.d_raw_rgb = d_raw_rgb;
  a.act = (const uint8_t*)act; a.etiles = (const uint8_t*)etiles; a.grads = *grads;
  a.ring = w; a.sync = (uint32_t*)(w + (size_t)ctas * tcf::RING_BYTES); a.total = total; a.num_tiles = num_tiles; a.bg = bg ? 1 : 0;
  cudaMemsetAsync(a.sync, 0, (size_t)ctas * tcf::SYNC_WORDS * sizeof(uint32_t), st);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(tcf::THREADS); cfg.dynamicSmemBytes = tcf::SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;      // all CTAs resident, or the launch is refused
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tcf::bwd_fused_kernel, a, ro);
  if (e != cudaSuccess) { npp_set_error("bwd_fused launch (%d producers + %d consumers): %s", ro.P, C, cudaGetErrorString(e)); return (int)e; }
  return 0;
}
