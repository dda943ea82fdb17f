// Backward of the field MLP, weight-gradient half:  dW_l = dZ_l^T X_l  summed over all samples.
//
// wgrad_tc_kernel: tcgen05 split-K GEMMs.  M = output features of the layer (two halves of 128 = TMEM lanes),
// N = input features (<= 256 TMEM columns per half), K = samples.  Both operands are the buffers the forward (ACT, E)
// and the dgrad kernel (DZ) wrote in the operand-chunk layout: a [128 samples x 64 features] SWIZZLE_128B chunk read
// MN-major IS the [features x samples] operand, so tiles go HBM -> shared memory with plain bulk copies and no
// transposition anywhere.  Each CTA owns one (job, sample-range) pair, accumulates in fp32 in TMEM over its tiles and
// adds its partial into the fp32 gradient with red.global.add (the caller zero-fills or accumulates).
// This kernel is HBM-bound by construction (128 KB of operands per 2048 tensor-pipe cycles).
//
// The bias gradients (column sums of DZ) are taken by wgrad_tc_kernel's otherwise idle epilogue warps from the A tiles in
// shared memory.  wgrad_small_kernel: the two tiny heads (sigma 256->1, rgb.2 128->3) on CUDA cores.
#include "bwd_common.cuh"

namespace npp {
namespace tcw {
using namespace npp::tc;

constexpr int THREADS = 192;          // warps 0-3 epilogue, 4 MMA issuer, 5 loader
constexpr int X_BYTES = 4 * CHUNK_BYTES, AH_BYTES = 2 * CHUNK_BYTES;
constexpr int OFF_X = 0, OFF_A = 2 * X_BYTES, OFF_BAR = OFF_A + 2 * AH_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 128;
enum { B_XFULL = 0, B_XEMPTY = 2, B_AFULL = 4, B_AEMPTY = 6, B_DONE = 8, B_COUNT = 9 };

constexpr int HEAD_PART = 512;        // floats per CTA of a head job: <= 3 x 128 weight partials, bias partial(s) at 256 / 384..386
__constant__ JobTable c_jobs[2] = {make_jobs(false), make_jobs(true)};
static const JobTable h_jobs[2] = {make_jobs(false), make_jobs(true)};

__global__ void __launch_bounds__(THREADS, 1)
wgrad_tc_kernel(int bg, const uint8_t* __restrict__ act, const uint8_t* __restrict__ etiles, const uint8_t* __restrict__ dz,
                const float* __restrict__ scale_ptr, int num_tiles, int splits, NerfppNetGrads grads, float* __restrict__ part_w,
                float* __restrict__ part_b, float* __restrict__ part_s, const float* __restrict__ d_raw_sigma,
                const float* __restrict__ d_raw_rgb, long long total) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t s_base = smem_u32(smem);
  auto bar = [&](int i) { return s_base + OFF_BAR + 8u * i; };
  if ((s_base & 1023u) != 0) __trap();
  const Job jb = c_jobs[bg].j[blockIdx.x / splits];
  const int split = blockIdx.x % splits;
  const int t_begin = (int)((long long)num_tiles * split / splits), t_end = (int)((long long)num_tiles * (split + 1) / splits);
  const int N = 64 * jb.x_nchunks;
  const size_t nt = (size_t)num_tiles;
  const int head = part_s ? jb.head : 0;          // (the heads ride along only in the deterministic two-kernel backward)

  if (threadIdx.x == 0) {
    // an A half is released by the MMAs' commit and, in a bias job, also by each of the four column-sum warps; likewise
    // the X tile in a job that carries a head
    for (int i = 0; i < 2; ++i) { mbar_init(bar(B_XFULL + i), 1); mbar_init(bar(B_XEMPTY + i), head ? 5 : 1); mbar_init(bar(B_AFULL + i), 1); mbar_init(bar(B_AEMPTY + i), jb.bias ? 5 : 1); }
    mbar_init(bar(B_DONE), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 5) {
    // ================= loader: X tile, then the A halves, per sample tile =================
    if (lane == 0) {
      uint32_t ix = 0, ia = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        const uint32_t xb = ix & 1;
        mbar_wait(bar(B_XEMPTY + xb), ((ix >> 1) & 1) ^ 1);
        const uint8_t* xsrc = jb.x_is_e ? etiles + ((size_t)tile * 2 + jb.x_chunk0) * CHUNK_BYTES
                                        : act + act_chunk_off(jb.x_layer, nt, (size_t)tile, jb.x_chunk0);
        mbar_expect_tx(bar(B_XFULL + xb), (uint32_t)((jb.x_nchunks + (head == 2 ? 2 : 0)) * CHUNK_BYTES));
        bulk_g2s(s_base + OFF_X + xb * X_BYTES, xsrc, (uint32_t)(jb.x_nchunks * CHUNK_BYTES), bar(B_XFULL + xb));
        if (head == 2)      // the rgb hidden layer's tile, behind this job's single X chunk
          bulk_g2s(s_base + OFF_X + xb * X_BYTES + CHUNK_BYTES, act + act_chunk_off(9, nt, (size_t)tile, 0), 2 * CHUNK_BYTES, bar(B_XFULL + xb));
        ++ix;
        for (int h = 0; h < jb.m_halves; ++h, ++ia) {
          const uint32_t ab = ia & 1;
          mbar_wait(bar(B_AEMPTY + ab), ((ia >> 1) & 1) ^ 1);
          mbar_expect_tx(bar(B_AFULL + ab), AH_BYTES);
          bulk_g2s(s_base + OFF_A + ab * AH_BYTES, dz + act_chunk_off(jb.a_layer, nt, (size_t)tile, 2 * h), AH_BYTES, bar(B_AFULL + ab));
        }
      }
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    const uint32_t idesc = idesc_mn(N);
    uint32_t ix = 0, ia = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++ix) {
      const uint32_t xb = ix & 1;
      mbar_wait(bar(B_XFULL + xb), (ix >> 1) & 1);
      for (int h = 0; h < jb.m_halves; ++h, ++ia) {
        const uint32_t ab = ia & 1;
        mbar_wait(bar(B_AFULL + ab), (ia >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t alo = mn_lo(s_base + OFF_A + ab * AH_BYTES), blo = mn_lo(s_base + OFF_X + xb * X_BYTES);
          const uint32_t d = tmem_base + 256u * (uint32_t)h;
#pragma unroll
          for (int k = 0; k < 8; ++k) {        // 16 samples per MMA = two 8-sample groups = 2048 B
            if (k == 0 && tile == t_begin) mma_ss<0>(d, alo, MN_HI, blo, MN_HI, idesc);
            else mma_ss<1>(d, alo + 128u * k, MN_HI, blo + 128u * k, MN_HI, idesc);
          }
          tc_commit(bar(B_AEMPTY + ab));
          if (h == jb.m_halves - 1) tc_commit(bar(B_XEMPTY + xb));
          if (tile == t_end - 1 && h == jb.m_halves - 1) tc_commit(bar(B_DONE));
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue: TMEM -> red.global.add into dW =================
    // While the MMAs run these four warps are idle, and every DZ tile of the layer passes through shared memory as the
    // A operand: in a bias job they add it up over the samples (the bias gradient) instead of a second kernel reading
    // all of DZ from HBM again.  Thread = one pair of adjacent features x one half of the 128 sample rows; a warp reads
    // one 128-byte row of the swizzled image per step (conflict-free).
    if ((jb.bias || head) && t_end > t_begin) {
      const int t = threadIdx.x, p = t & 63, rh = t >> 6;
      const int col = 2 * (p & 31);
      const uint32_t coff = (uint32_t)((p >> 5) * CHUNK_BYTES + (col & 7) * 2);
      float bsum[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
      float hs[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}}, hb[3] = {0.f, 0.f, 0.f};     // head: weight-gradient partials, bias partial
      __shared__ float s_hw[2][3 * TILE];           // the tile's per-sample head gradients (double-buffered by tile parity)
      uint32_t ia = 0, ix = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        if (head) {
          // every thread fetches one sample's gradient(s); zero beyond the last sample
          const long long g = (long long)tile * TILE + t;
          float* hw = s_hw[ix & 1];
          if (head == 1) hw[t] = g < total ? d_raw_sigma[g] : 0.f;
          else {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) hw[3 * t + ch] = g < total ? d_raw_rgb[3 * g + ch] : 0.f;
          }
          const uint32_t xb = ix & 1;
          mbar_wait(bar(B_XFULL + xb), (ix >> 1) & 1);
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (head == 1) {
            // sigma head: columns (2 (t & 31), +1) of chunk t >> 5 of h7, all 128 rows; a warp reads one 128-byte row per step
            const uint8_t* xc = smem + OFF_X + xb * X_BYTES + (t >> 5) * CHUNK_BYTES + ((t & 31) & 3) * 4;
            const int unit = (t & 31) >> 2;
#pragma unroll 8
            for (int r = 0; r < TILE; ++r) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(xc + (r >> 3) * 1024 + (r & 7) * 128 + ((unit ^ (r & 7)) << 4)));
              const float w = hw[r];
              hs[0][0] = fmaf(w, f.x, hs[0][0]);
              hs[0][1] = fmaf(w, f.y, hs[0][1]);
              hb[0] += w;
            }
          } else {
            // rgb.2: columns (col, col + 1) of the rgb hidden tile (behind the E chunk), this thread's half of the rows
            const uint8_t* xc = smem + OFF_X + xb * X_BYTES + CHUNK_BYTES + coff;
#pragma unroll 4
            for (int k = 0; k < 64; ++k) {
              const int r = rh * 64 + k;
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(xc + (r >> 3) * 1024 + (r & 7) * 128 + (((col >> 3) ^ (r & 7)) << 4)));
#pragma unroll
              for (int ch = 0; ch < 3; ++ch) {
                const float w = hw[3 * r + ch];
                hs[ch][0] = fmaf(w, f.x, hs[ch][0]);
                hs[ch][1] = fmaf(w, f.y, hs[ch][1]);
                hb[ch] += w;
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(B_XEMPTY + xb));
          ++ix;
        }
        if (jb.bias) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h < jb.m_halves) {
              const uint32_t ab = ia & 1;
              mbar_wait(bar(B_AFULL + ab), (ia >> 1) & 1);
              const uint8_t* ah = smem + OFF_A + ab * AH_BYTES + coff;
#pragma unroll 8
              for (int k = 0; k < 64; ++k) {
                const int r = rh * 64 + k;
                const __half2 v = *reinterpret_cast<const __half2*>(ah + (r >> 3) * 1024 + (r & 7) * 128 + (((col >> 3) ^ (r & 7)) << 4));
                const float2 f = __half22float2(v);
                bsum[h][0] += f.x;
                bsum[h][1] += f.y;
              }
              __syncwarp();
              if (lane == 0) mbar_arrive(bar(B_AEMPTY + ab));
              ++ia;
            }
          }
        }
      }
      __shared__ float s_b[2][256];
      if (jb.bias) {
        const float inv_scale_b = 1.f / scale_ptr[jb.a_layer >= 8 ? 1 : 0];
        float* db = grads.b[jb.a_layer < 8 ? jb.a_layer : jb.a_layer == 8 ? L_REMAP : L_RGB0];
        // two threads (rh = 0 / 1: the two halves of the sample rows) hold partial sums of the same feature pair: they meet
        // in shared memory and are added in a fixed order
#pragma unroll
        for (int h = 0; h < 2; ++h)
          if (h < jb.m_halves) {
            const int f0 = 128 * h + 64 * (p >> 5) + col;
            s_b[rh][f0] = bsum[h][0];
            s_b[rh][f0 + 1] = bsum[h][1];
          }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int f = t; f < 128 * jb.m_halves; f += 128) {
          const float v = (s_b[0][f] + s_b[1][f]) * inv_scale_b;
          if (part_b) part_b[(size_t)blockIdx.x * 256 + f] = v;      // deterministic: summed over the splits by wgrad_reduce_kernel
          else atomicAdd(db + f, v);
        }
      }
      if (head == 1) {          // sigma head: 256 weight partials (one column pair per thread) + the bias partial
        const float inv = 1.f / scale_ptr[0];
        float* ps = part_s + (size_t)blockIdx.x * HEAD_PART;
        ps[2 * t] = hs[0][0] * inv;
        ps[2 * t + 1] = hs[0][1] * inv;
        if (t == 0) ps[256] = hb[0] * inv;
      } else if (head == 2) {   // rgb.2: the two row halves of a column pair meet in shared memory, fixed order
        const float inv = 1.f / scale_ptr[1];
        float* ps = part_s + (size_t)blockIdx.x * HEAD_PART;
        const int c0 = 64 * (p >> 5) + col;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          asm volatile("bar.sync 1, 128;" ::: "memory");         // the previous channel's sums have been read
          s_b[rh][c0] = hs[ch][0];
          s_b[rh][c0 + 1] = hs[ch][1];
          if (p == 0) s_b[rh][128] = hb[ch];                     // (every thread of a row half holds the same bias partial)
          asm volatile("bar.sync 1, 128;" ::: "memory");
          ps[ch * RGB_HID + t] = (s_b[0][t] + s_b[1][t]) * inv;
          if (t == 0) ps[384 + ch] = (s_b[0][128] + s_b[1][128]) * inv;
        }
      }
    }
    if (t_end > t_begin) {
      mbar_wait(bar(B_DONE), 0);
      tc_fence_after();
      const float inv_scale = 1.f / scale_ptr[jb.a_layer >= 8 ? 1 : 0];     // colour-path layers carry scale[1] (backward.cu)
      float* dW = grads.w[jb.w_index];
      const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
      for (int h = 0; h < jb.m_halves; ++h) {
        const int orow = 128 * h + warp * 32 + lane;
        float* wrow = dW + (size_t)orow * jb.ld + jb.col0;
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(lane_addr + 256u * (uint32_t)h + (uint32_t)c0, v);
          tmem_ld_wait(v);
          if (part_w) {     // deterministic: this CTA's partial [256 rows][256 columns], 128 contiguous bytes per thread
            float4* dst = reinterpret_cast<float4*>(part_w + ((size_t)blockIdx.x * 256 + orow) * 256 + c0);
#pragma unroll
            for (int e = 0; e < 8; ++e)
              dst[e] = make_float4(__uint_as_float(v[4 * e]) * inv_scale, __uint_as_float(v[4 * e + 1]) * inv_scale,
                                   __uint_as_float(v[4 * e + 2]) * inv_scale, __uint_as_float(v[4 * e + 3]) * inv_scale);
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int col = c0 + e - jb.skip;
              if (col >= 0 && col < jb.ncols) atomicAdd(wrow + col, __uint_as_float(v[e]) * inv_scale);
            }
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

// Sums the per-CTA partials of wgrad_tc_kernel over the sample splits IN ORDER and adds them to the gradients: every
// weight / bias gradient entry is produced by exactly one job, so the result does not depend on scheduling -- the
// training step is bit-reproducible (with red.global.add the order of the 24 partials per entry was not).
// ordered sum of `n` partials `stride` floats apart: the loads are issued eight at a time (independent), the additions stay in
// index order, so the result is the same whatever the scheduling
__device__ __forceinline__ float ordered_sum(const float* __restrict__ p, size_t stride, int n) {
  float acc = 0.f;
  int i = 0;
  for (; i + 8 <= n; i += 8) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(p + (size_t)(i + k) * stride);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += v[k];
  }
  for (; i < n; ++i) acc += __ldg(p + (size_t)i * stride);
  return acc;
}

__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(int bg, const float* __restrict__ part_w, const float* __restrict__ part_b, const float* __restrict__ part_s, int num_tiles,
                    int splits, NerfppNetGrads grads) {
  const JobTable& jt = c_jobs[bg];
  const int j = blockIdx.y;
  if (j >= jt.n) return;
  const Job jb = jt.j[j];
  const int rows = 128 * jb.m_halves;
  const int n_b = jb.bias ? rows : 0;
  const int n_h = !part_s ? 0 : jb.head == 1 ? W + 1 : jb.head == 2 ? 3 * RGB_HID + 3 : 0;      // head weights, then its bias(es)
  const int n_w = rows * jb.ncols, n_all = n_w + n_b + n_h;
  // splits with an empty sample range wrote nothing: they are the trailing ones only when splits > num_tiles, which the
  // launcher excludes (splits <= num_tiles), so every split has at least one tile
  float* dW = grads.w[jb.w_index];
  float* db = grads.b[jb.a_layer < 8 ? jb.a_layer : jb.a_layer == 8 ? L_REMAP : L_RGB0];
  const size_t cta0 = (size_t)j * splits;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_all; i += gridDim.x * blockDim.x) {
    if (i < n_w) {
      const int orow = i / jb.ncols, col = i - orow * jb.ncols;
      dW[(size_t)orow * jb.ld + jb.col0 + col] += ordered_sum(part_w + (cta0 * 256 + orow) * 256 + col + jb.skip, (size_t)256 * 256, splits);
    } else if (i < n_w + n_b) {
      const int f = i - n_w;
      db[f] += ordered_sum(part_b + cta0 * 256 + f, 256, splits);
    } else {
      const int f = i - n_w - n_b;
      const int nw = jb.head == 1 ? W : 3 * RGB_HID;
      const int src = f < nw ? f : (jb.head == 1 ? 256 : 384) + (f - nw);
      const float acc = ordered_sum(part_s + cta0 * HEAD_PART + src, HEAD_PART, splits);
      const int l = jb.head == 1 ? L_SIGMA : L_RGB2;
      if (f < nw) grads.w[l][f] += acc; else grads.b[l][f - nw] += acc;
    }
  }
}

// Weighted column sums over all samples of a chunked [tiles][nch][128 rows][64 cols] fp16 operand buffer:
//   out[ch][col] = inv_scale * sum_rows wgt[row][ch] * X[row][col]
// This is the sigma head (X = h7, wgt = d raw sigma) and rgb.2 (X = rgb hidden, wgt = d raw rgb, NCH = 3).  blockIdx.x = sample-range split; one thread owns one 16-byte unit (8 columns) of
// four rows of every chunk, so all loads are 16-byte and a warp reads 512 contiguous bytes; the 32 row groups' partials of
// a column meet in shared memory and are added in a FIXED order.  HBM-bound: every operand byte is read exactly once.
// The CTA's result goes to `part` (deterministic two-stage reduction: heads_reduce_kernel adds the CTAs in order).
template <int NCH>
__device__ __forceinline__ void weighted_colsum(const uint8_t* __restrict__ layer_base, int nch, const float* __restrict__ wgt,
                                                long long total, int t_begin, int t_end, float inv_scale, float* s_scr,
                                                float* __restrict__ part) {
  const int tid = threadIdx.x, u = tid & 7, r8 = tid >> 3;          // physical unit, row group: rows r8 + 32 k
  const int lu = u ^ (r8 & 7);                                       // logical unit (columns 8 lu .. 8 lu + 7) of all four rows
  for (int c = 0; c < nch; ++c) {
    float acc[NCH][8];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[ch][e] = 0.f;
    for (int tile = t_begin; tile < t_end; ++tile) {
      const uint8_t* chunk = layer_base + ((size_t)tile * nch + c) * (size_t)CHUNK_BYTES;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = r8 + 32 * k;
        const uint4 v = *reinterpret_cast<const uint4*>(chunk + (r >> 3) * 1024 + (r & 7) * 128 + u * 16);
        const __half2* hp = reinterpret_cast<const __half2*>(&v);
        float w[NCH];
        const long long g = (long long)tile * TILE + r;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) w[ch] = g < total ? wgt[(size_t)g * NCH + ch] : 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(hp[e]);
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) { acc[ch][2 * e] = fmaf(w[ch], f.x, acc[ch][2 * e]); acc[ch][2 * e + 1] = fmaf(w[ch], f.y, acc[ch][2 * e + 1]); }
        }
      }
    }
    // s_scr[r8][ch][64 columns of this chunk]; then column (ch, col) = sum over r8 = 0..31 in order
    __syncthreads();
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int e = 0; e < 8; ++e) s_scr[(r8 * NCH + ch) * 64 + lu * 8 + e] = acc[ch][e];
    __syncthreads();
    for (int i = tid; i < NCH * 64; i += 256) {
      const int ch = i >> 6, col = i & 63;
      float a = 0.f;
#pragma unroll 8
      for (int q = 0; q < 32; ++q) a += s_scr[(q * NCH + ch) * 64 + col];
      part[ch * 64 * nch + c * 64 + col] = a * inv_scale;
    }
  }
}

constexpr int HEADS_PART = 512;     // floats per (CTA, head): <= 3 x 128 weight partials, then the bias partial(s) at [384..387)

// blockIdx.y: 0 = sigma head weights (X = h7, weights d raw sigma), 1 = rgb.2 weights (X = rgb hidden, weights d raw rgb);
// the heads' bias gradients (plain sums of 1 / 3 numbers per sample) are done by the same CTAs.
__global__ void __launch_bounds__(256)
wgrad_small_kernel(const uint8_t* __restrict__ act, const float* __restrict__ d_raw_sigma, const float* __restrict__ d_raw_rgb,
                   const float* __restrict__ scale_ptr, long long total, int num_tiles, float* __restrict__ part_h) {
  __shared__ float s_scr[32 * 3 * 64];
  __shared__ float s_warp[8];
  const int what = 10 + blockIdx.y;     // (the layers' bias gradients are summed inside wgrad_tc_kernel)
  const int t_begin = (int)((long long)num_tiles * blockIdx.x / gridDim.x), t_end = (int)((long long)num_tiles * (blockIdx.x + 1) / gridDim.x);
  const float inv_scale = 1.f / scale_ptr[what == 10 ? 0 : 1];       // d raw sigma carries scale[0], d raw rgb scale[1]
  const size_t nt = (size_t)num_tiles;
  float* part = part_h + ((size_t)blockIdx.x * 2 + blockIdx.y) * HEADS_PART;
  if (what == 10) weighted_colsum<1>(act + act_layer_off(7, nt), 4, d_raw_sigma, total, t_begin, t_end, inv_scale, s_scr, part);
  else weighted_colsum<3>(act + act_layer_off(9, nt), 2, d_raw_rgb, total, t_begin, t_end, inv_scale, s_scr, part);
  // bias of the head: sum over the CTA's samples of d raw sigma / d raw rgb (warp sums, then the 8 warps in order)
  const int nch = what == 10 ? 1 : 3;
  const float* d = what == 10 ? d_raw_sigma : d_raw_rgb;
  long long lo = (long long)t_begin * TILE, hi = (long long)t_end * TILE;
  if (hi > total) hi = total;
  for (int ch = 0; ch < nch; ++ch) {
    float a = 0.f;
    for (long long g = lo + threadIdx.x; g < hi; g += 256) a += d[(size_t)g * nch + ch];
    a = warp_sum(a);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += s_warp[w];
      part[384 + ch] = t * inv_scale;
    }
  }
}

// adds the CTAs' partials of wgrad_small_kernel in order: block 0 = sigma head (256 weights + 1 bias), block 1 = rgb.2 (384 + 3)
__global__ void __launch_bounds__(512)
heads_reduce_kernel(const float* __restrict__ part_h, int n_ctas, NerfppNetGrads grads) {
  const int what = blockIdx.x, i = threadIdx.x;
  const int n_w = what == 0 ? W : 3 * RGB_HID, n_b = what == 0 ? 1 : 3;
  if (i >= n_w && !(i >= 384 && i < 384 + n_b)) return;
  const float a = ordered_sum(part_h + (size_t)what * HEADS_PART + i, (size_t)2 * HEADS_PART, n_ctas);
  if (i < n_w) grads.w[what == 0 ? L_SIGMA : L_RGB2][i] += a;
  else grads.b[what == 0 ? L_SIGMA : L_RGB2][i - 384] += a;
}

}  // namespace tcw
}  // namespace npp

using namespace npp;

// Workspace of the deterministic reductions: the per-CTA partials of wgrad_tc_kernel (weights, biases) and of the heads kernel.
static int wgrad_splits(int num_sms, int njobs, int num_tiles) {
  int splits = (2 * num_sms) / njobs;              // ~2 waves of CTAs; every job gets the same number of sample ranges
  if (splits > num_tiles) splits = num_tiles;
  return splits < 1 ? 1 : splits;
}
static int heads_ctas(int num_sms, int num_tiles) { return num_tiles < 4 * num_sms ? num_tiles : 4 * num_sms; }
size_t npp_wgrad_ws_bytes(int dev) {
  int num_sms = 148;
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t ctas = (size_t)2 * num_sms + tcw::MAX_JOBS;
  return ctas * 256 * 256 * 4 + ctas * 256 * 4 + ctas * tcw::HEAD_PART * 4 + (size_t)4 * num_sms * 2 * tcw::HEADS_PART * 4 + 1024;
}

int npp_field_wgrad_heads(const void* act, const float* d_raw_sigma, const float* d_raw_rgb, const float* scale, long long total,
                          const NerfppNetGrads* grads, void* ws, cudaStream_t st);

// The two small heads (sigma, rgb.2) can ride along in wgrad_tc_kernel's base_remap / view-direction jobs (Job::head: no
// second pass over ACT[7] / ACT[9], 1.7 GB per step) -- built, correct, and measured 0.2 ms per step SLOWER than their own
// HBM-bound kernel (10.41 vs 10.21 ms, same box, profiles/r2_experiments/exp18.log): the CUDA-core sums make those two jobs' CTAs the
// long pole of the launch.  Off by default; tests may switch it on.
static int g_heads_folded = 0;
extern "C" void nerfpp_debug_set_heads_folded(int on) { g_heads_folded = on ? 1 : 0; }

// Accumulates (+=) the gradients of one net's 24 parameter tensors.  `ws`: npp_wgrad_ws_bytes() bytes (256-byte aligned).
int npp_field_wgrad(bool bg, const void* act, const void* etiles, const void* dz, const float* d_raw_sigma, const float* d_raw_rgb,
                    const float* scale, long long total, const NerfppNetGrads* grads, void* ws, cudaStream_t st) {
  static int sms_dev[64] = {0};               // both caches are per device (the shared-memory opt-in is a per-device attribute)
  static bool configured_dev[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  int& num_sms = sms_dev[dev & 63];
  bool& configured = configured_dev[dev & 63];
  if (num_sms == 0) cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (!configured) { cudaFuncSetAttribute(tcw::wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcw::SMEM_BYTES); configured = true; }
  const int num_tiles = (int)((total + tc::TILE - 1) / tc::TILE);
  const int njobs = tcw::h_jobs[bg].n;
  const int splits = wgrad_splits(num_sms, njobs, num_tiles);
  const size_t ctas_max = (size_t)2 * num_sms + tcw::MAX_JOBS;
  float* part_w = (float*)ws;
  float* part_b = part_w + ctas_max * 256 * 256;
  float* part_s = part_b + ctas_max * 256;
  // the two small heads (sigma, rgb.2) ride along in the base_remap and view-direction jobs: no second pass over ACT[7] / ACT[9]
  tcw::wgrad_tc_kernel<<<njobs * splits, tcw::THREADS, tcw::SMEM_BYTES, st>>>(bg ? 1 : 0, (const uint8_t*)act, (const uint8_t*)etiles,
                                                                               (const uint8_t*)dz, scale, num_tiles, splits, *grads, part_w, part_b,
                                                                               g_heads_folded ? part_s : nullptr, d_raw_sigma, d_raw_rgb, total);
  NPP_CHECK_LAUNCH();
  tcw::wgrad_reduce_kernel<<<dim3(128, njobs), 256, 0, st>>>(bg ? 1 : 0, part_w, part_b, g_heads_folded ? part_s : nullptr, num_tiles, splits, *grads);
  NPP_CHECK_LAUNCH();
  if (!g_heads_folded) return npp_field_wgrad_heads(act, d_raw_sigma, d_raw_rgb, scale, total, grads, part_s + ctas_max * tcw::HEAD_PART, st);
  return 0;
}

// The two small heads (sigma: 256 -> 1, rgb.2: 128 -> 3): weighted column sums of h7 / the rgb hidden layer on CUDA cores.
// `part_h`: 4 * #SMs * 2 * 512 floats (inside npp_wgrad_ws_bytes' workspace).
int npp_field_wgrad_heads(const void* act, const float* d_raw_sigma, const float* d_raw_rgb, const float* scale, long long total,
                          const NerfppNetGrads* grads, void* part_h, cudaStream_t st) {
  int dev = 0, num_sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  const int num_tiles = (int)((total + tc::TILE - 1) / tc::TILE);
  // 96 KB of activations per tile, read once.  A thread's loop over its tiles is a chain of dependent round trips to HBM
  // (4 x 16 B in flight per thread), so the sample range is cut fine enough for ~8 CTAs per SM to cover the latency.
  const int sx = heads_ctas(num_sms, num_tiles);
  tcw::wgrad_small_kernel<<<dim3(sx, 2), 256, 0, st>>>((const uint8_t*)act, d_raw_sigma, d_raw_rgb, scale, total, num_tiles, (float*)part_h);
  NPP_CHECK_LAUNCH();
  tcw::heads_reduce_kernel<<<2, 512, 0, st>>>((const float*)part_h, sx, *grads);
  NPP_CHECK_LAUNCH();
  return 0;
}
