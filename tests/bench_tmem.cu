// Microbenchmark (diagnostic): tcgen05.ld throughput from TMEM, as the field kernel's epilogue uses it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/bench_tmem tests/bench_tmem.cu && tests/bench_tmem
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}

__device__ __forceinline__ void ld32x32b_x32(uint32_t a, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
                 "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(a));
}
__device__ __forceinline__ void ld32x32b_x16(uint32_t a, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]) : "r"(a));
}
__device__ __forceinline__ void ld16x256b_x8(uint32_t a, uint32_t (&v)[32]) {   // 16 lanes x 64 columns
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
                 "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(a));
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MODE 0: 32x32b.x32, wait after each   1: two x32 in flight   2: 32x32b.x16, wait after each   3: 16x256b.x8 (two per 32 lanes)
// 4: x32 + 16 cvt + tcgen05.st x16 + wait::st  (the epilogue's per-chunk sequence without barriers)
// MODE 5: MODE 4 + tcgen05.fence::before_thread_sync + __syncwarp + mbarrier.arrive (the complete per-chunk sequence)
// MODE 6: MODE 5 while another warp keeps the tensor pipe busy with N=256 MMAs into the other 256 TMEM columns
// MODE 7: like 6 but the MMAs read their A operand from TMEM columns [0,256), the buffer the readers work on (as in field_tc)
template <int MODE>
__global__ void __launch_bounds__(544, 1) bench(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ volatile int stop;
  extern __shared__ __align__(1024) uint8_t dsm[];
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int nwarps = (MODE >= 6) ? (blockDim.x >> 5) - 1 : (blockDim.x >> 5);
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(nwarps)); stop = 0; }
  __syncthreads();
  if (MODE >= 6 && warp == nwarps) {       // MMA feeder: D in columns [256,512), operands = whatever is in shared memory
    __syncthreads();                         // the readers' start line
    if ((threadIdx.x & 31) == 0) {
      const uint64_t ad = sw128_desc(smem_u32(dsm)), bd = sw128_desc(smem_u32(dsm) + 16384);
      const uint32_t id = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
      while (atomicAdd((int*)&stop, 0) < nwarps) { for (int k = 0; k < 8; ++k) { if (MODE == 7) umma_ts(tmem + 256, tmem + 8 * k + 64 * (k & 3), bd + 2 * (k & 3), id); else umma_ss(tmem + 256, ad + 2 * (k & 3), bd + 2 * (k & 3), id); } }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[3])) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    return;
  }
  const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t a = base + (uint32_t)((it & 3) * 64);
    if (MODE == 0) { uint32_t v[32]; ld32x32b_x32(a, v); ld_wait(); acc ^= v[0] ^ v[31]; }
    if (MODE == 1) { uint32_t v[32], w[32]; ld32x32b_x32(a, v); ld32x32b_x32(a ^ 64u, w); ld_wait(); acc ^= v[0] ^ w[31]; }
    if (MODE == 2) { uint32_t v[16]; ld32x32b_x16(a, v); ld_wait(); acc ^= v[0] ^ v[15]; }
    if (MODE == 3) { uint32_t v[32], w[32]; ld16x256b_x8(a, v); ld16x256b_x8(a + (16u << 16), w); ld_wait(); acc ^= v[0] ^ w[31]; }
    if (MODE >= 4) {
      uint32_t v[32]; ld32x32b_x32(a, v); ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(pk[t]) : "f"(__uint_as_float(v[2 * t + 1])), "f"(__uint_as_float(v[2 * t])));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                   ::"r"(a), "r"(pk[0]),"r"(pk[1]),"r"(pk[2]),"r"(pk[3]),"r"(pk[4]),"r"(pk[5]),"r"(pk[6]),"r"(pk[7]),"r"(pk[8]),"r"(pk[9]),"r"(pk[10]),"r"(pk[11]),"r"(pk[12]),"r"(pk[13]),"r"(pk[14]),"r"(pk[15]) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      if (MODE >= 5) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if ((threadIdx.x & 31) == 0) asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(&bars[it & 1])) : "memory");
      }
    }
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) { out[blockIdx.x * 16 + warp] = t1 - t0; if (MODE >= 6) atomicAdd((int*)&stop, 1); }
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();

  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int MODE> void run(const char* name, int warps, int bytes_per_iter_per_warp, long long* out, uint32_t* sink) {
  const int iters = 4000;
  cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  bench<MODE><<<148, (warps + (MODE >= 6 ? 1 : 0)) * 32, 65536>>>(iters, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(e)); return; }
  long long h[148 * 16];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = 0;
  for (int b = 0; b < 148; ++b) { long long m = 0; for (int w = 0; w < warps; ++w) m = h[b * 16 + w] > m ? h[b * 16 + w] : m; cyc += (double)m; }
  cyc /= 148;
  printf("%-46s warps %2d: %7.1f cycles/iter/warp   %7.1f B/cycle/SM\n", name, warps, cyc / iters, (double)warps * bytes_per_iter_per_warp * iters / cyc);
}

int main() {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 148 * 16 * sizeof(long long)); cudaMalloc(&sink, 4);
  for (int warps : {4, 8, 16}) {
    run<0>("32x32b.x32 + wait", warps, 4096, out, sink);
    run<1>("2 x 32x32b.x32 in flight + wait", warps, 8192, out, sink);
    run<2>("32x32b.x16 + wait", warps, 2048, out, sink);
    run<3>("2 x 16x256b.x8 (32 lanes x 64 cols) + wait", warps, 8192, out, sink);
    run<4>("x32 ld + 16 cvt + x16 st + wait::st", warps, 4096, out, sink);
    run<5>("  ... + fence + syncwarp + mbarrier.arrive", warps, 4096, out, sink);
    run<6>("  ... same, tensor pipe busy (N=256 MMAs)", warps, 4096, out, sink);
    run<7>("  ... same, MMAs read A from the same buffer", warps, 4096, out, sink);
  }
  return 0;
}
