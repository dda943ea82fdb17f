// Microbenchmark (diagnostic, not a test): cycles per tcgen05.mma for the shapes field_tc.cu uses, one CTA
// per SM, one thread issuing a long run of MMAs back to back with a single commit at the end.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/bench_umma tests/bench_umma.cu && tests/bench_umma
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
  return pred != 0;
}
// issue pace of the thread itself (N=16 MMAs leave the pipe idle): VAR 0 = one thread of a diverged warp,
// VAR 1 = converged warp + elect.sync around each group of 4
template <int VAR>
__global__ void __launch_bounds__(128, 1) pace(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t a_s = smem_u32(smem), b_s = a_s + 16384;
  const uint64_t ad = sw128_desc(a_s), bd = sw128_desc(b_s);
  const uint32_t id = idesc_f16(16);
  if (VAR == 0 ? (threadIdx.x == 32) : (warp == 1)) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (VAR == 0 || elect_one()) {
        umma_ts(tmem, tmem + 256, bd, id, 1u);
        umma_ts(tmem, tmem + 264, bd + 2, id, 1u);
        umma_ts(tmem, tmem + 288, bd + 4, id, 1u);
        umma_ts(tmem, tmem + 296, bd + 6, id, 1u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      if (VAR == 1) __syncwarp();
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = t1 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// mode: 0 = SS N=256, 1 = TS N=256, 2 = SS N=128, 3 = TS N=128, 4 = SS N=256 with a commit + mbarrier wait every 4 MMAs
template <int mode>
__global__ void __launch_bounds__(128, 1) bench(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 32) {
    const uint32_t a_s = smem_u32(smem), b_s = a_s + 16384;
    const uint64_t ad = sw128_desc(a_s), bd = sw128_desc(b_s);
    const int n = (mode == 2 || mode == 3) ? 128 : 256;
    const uint32_t id = idesc_f16(n);
    uint32_t par = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == 1 || mode == 3) {
        umma_ts(tmem, tmem + 256, bd, id, 1u);
        umma_ts(tmem, tmem + 264, bd + 2, id, 1u);
        umma_ts(tmem, tmem + 288, bd + 4, id, 1u);
        umma_ts(tmem, tmem + 296, bd + 6, id, 1u);
      } else {
        umma_ss(tmem, ad, bd, id, 1u);
        umma_ss(tmem, ad + 2, bd + 2, id, 1u);
        umma_ss(tmem, ad + 4, bd + 4, id, 1u);
        umma_ss(tmem, ad + 6, bd + 6, id, 1u);
      }
      if (mode == 5 || (mode == 6 && (it & 3) == 3)) {   // commit only, never waited on
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      if (mode == 7 && it >= 2) {   // an (already complete) barrier wait per 4 MMAs, no commit
        asm volatile("{\n.reg .pred p;\nW7:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D7;\nbra W7;\nD7:\n}" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
      }
      if (mode == 4) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        if (it >= 2) {   // wait for the commit issued two iterations ago (keeps ~8 MMAs in flight)
          asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&bar)), "r"(par) : "memory");
          par ^= 1;
        }
      }
    }
    long long t1 = clock64();
    if (mode < 4 || mode == 7) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    if (mode < 4 || mode == 7) asm volatile("{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D2;\nbra W2;\nD2:\n}" ::"r"(smem_u32(&bar)) : "memory");
    long long t2 = clock64();
    out[2 * blockIdx.x] = t1 - t0;
    out[2 * blockIdx.x + 1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int M> void launch1(int grid, int iters, long long* out) {
  cudaFuncSetAttribute(bench<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768);
  bench<M><<<grid, 128, 16384 + 32768>>>(iters, out);
}
void launch(int mode, int grid, int iters, long long* out) {
  switch (mode) {
    case 0: launch1<0>(grid, iters, out); break; case 1: launch1<1>(grid, iters, out); break;
    case 2: launch1<2>(grid, iters, out); break; case 3: launch1<3>(grid, iters, out); break;
    case 4: launch1<4>(grid, iters, out); break; case 5: launch1<5>(grid, iters, out); break;
    case 6: launch1<6>(grid, iters, out); break; case 7: launch1<7>(grid, iters, out); break;
    case 8: cudaFuncSetAttribute(pace<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152); pace<0><<<grid, 128, 49152>>>(iters, out); break;
    case 9: cudaFuncSetAttribute(pace<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152); pace<1><<<grid, 128, 49152>>>(iters, out); break;
  }
}
int main() {
  long long* out;
  cudaMalloc(&out, 2 * 148 * sizeof(long long));
  const char* names[] = {"SS N=256", "TS N=256", "SS N=128", "TS N=128", "SS N=256 commit+wait/4", "SS N=256 commit/4", "SS N=256 commit/16", "SS N=256 wait/4", "pace: diverged thread, N=16 + commit/4", "pace: elect.sync, N=16 + commit/4"};
  for (int grid : {148}) {
    for (int mode = 0; mode < 10; ++mode) {
      const int iters = 2000;
      launch(mode, grid, iters, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[2 * 148];
      cudaMemcpy(h, out, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
      double issue = 0, total = 0;
      for (int i = 0; i < grid; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
      printf("grid %3d  %-24s issue %.1f cyc/MMA   complete %.1f cyc/MMA\n", grid, names[mode], issue / grid / (4.0 * iters), total / grid / (4.0 * iters));
    }
  }
  return 0;
}
