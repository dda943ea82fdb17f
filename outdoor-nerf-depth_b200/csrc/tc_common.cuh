// PTX wrappers and operand-layout helpers shared by the tcgen05 kernels (field_tc.cu, field_bwd_tc.cu, wgrad_tc.cu).
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace npp {
namespace tc {

constexpr int TILE = 128;                   // samples per tile = MMA M = TMEM lanes
constexpr int CHUNK_BYTES = 16384;          // one [128 rows x 64 fp16] SWIZZLE_128B operand chunk

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
// waits on two barriers at once (their try_wait latencies overlap)
__device__ __forceinline__ void mbar_wait2(uint32_t bar_a, uint32_t par_a, uint32_t bar_b, uint32_t par_b) {
  asm volatile(
      "{\n.reg .pred p, q;\nWAIT2_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 q, [%2], %3;\n"
      "and.pred p, p, q;\n"
      "@p bra WAIT2_DONE;\nbra WAIT2_LOOP;\nWAIT2_DONE:\n}" ::"r"(bar_a), "r"(par_a), "r"(bar_b), "r"(par_b) : "memory");
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
// shared -> global bulk copy (async proxy), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_s2g_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
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
// true in exactly one lane of a converged warp.  tcgen05.mma / tcgen05.commit issued under it compile to bare
// UTCHMMA / UTCBAR; under a plain `lane == 0` branch every one of them is wrapped in a per-lane serialisation loop
// that nearly doubles the issuing thread's cost per MMA (tests/bench_umma.cu: 56 vs 33 cycles).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
  return pred != 0;
}
// Lean forms for the issue loop: descriptors as (lo, hi) words -- only lo changes between MMAs -- and a
// compile-time accumulate flag, so one MMA costs the issuing thread a couple of integer adds.
constexpr uint32_t SW128_HI = 64u | (1u << 14) | (2u << 29);       // SBO 1024 B, version 1, SWIZZLE_128B
constexpr uint32_t NOSW_HI = (128u >> 4) | (1u << 14);             // SBO 128 B, version 1, no swizzle
__device__ __forceinline__ uint32_t sw128_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint32_t bias_lo(uint32_t saddr, uint32_t n) { return ((saddr >> 4) & 0x3FFFu) | (n << 16); }
template <int ACC>
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}" ::"r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "n"(ACC)
      : "memory");
}
template <int ACC>
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t blo, uint32_t idesc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 db;\nsetp.ne.b32 p, %5, 0;\nmov.b64 db, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n}" ::"r"(d_tmem), "r"(a_tmem), "r"(blo), "r"(SW128_HI), "r"(idesc), "n"(ACC)
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
// tcgen05.ld is asynchronous: its destination registers are only valid after wait::ld.  Binding them to the
// wait as in/out operands keeps the compiler from reading or moving them across it.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                 "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                 "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :: "memory");
}
// four columns of the warp's 32 TMEM lanes (one value per lane and column)
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld4_sync(uint32_t taddr, uint32_t (&v)[4]) {   // load + wait::ld
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\ttcgen05.wait::ld.sync.aligned;"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
// two fp32 -> packed fp16x2 (lo in the low half), optionally through ReLU, one instruction
template <bool RELU>
__device__ __forceinline__ uint32_t pack_f16x2(uint32_t lo, uint32_t hi) {
  uint32_t d;
  // .satfinite: an activation beyond fp16's range (|h| > 65504, possible with trained weights) clamps to +-65504 instead
  // of turning into inf and poisoning every later layer with inf - inf = NaN; same single F2FP instruction
  if (RELU) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return d;
}

// same, saturating to +-65504 instead of overflowing to infinity (gradients under a loss scale)
__device__ __forceinline__ uint32_t pack_f16x2_sat(uint32_t lo, uint32_t hi) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return d;
}

// K-major, 128-byte-swizzled operand tile: rows of 64 fp16 (128 B), 8-row groups 1024 B apart.
// (cute::UMMA::SmemDescriptor: start>>4 | LBO=1<<16 | SBO=64<<32 | version=1<<46 | SWIZZLE_128B=2<<61)
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// K-major unswizzled [N x 16] tile: 8x8 core matrices of 128 B; N/8 along N (SBO = 128 B), 2 along K (LBO = 16 N B)
__device__ __forceinline__ uint64_t bias_desc(uint32_t saddr, int n) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)n << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr int bias_tile_off(int n_total, int n, int k) { return (k >> 3) * 16 * n_total + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2; }
// instruction descriptor: D=f32, A=B=f16, both K-major, M=128
__host__ __device__ constexpr uint32_t idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
// byte offset of element (row r, column c) of a [128 x 64k] region made of 64-column SW128 chunks
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)((c >> 6) * CHUNK_BYTES + (r >> 3) * 1024 + (r & 7) * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + (c & 7) * 2);
}


// ---- per-tile activation store (training): every layer's fp16 activations in the operand-chunk layout ------------
// ACT / DZ buffers: layer-major, [layer][tile][chunk][128 rows][64 cols] fp16, each chunk in the SWIZZLE_128B
// arrangement the MMAs consume (row r at (r>>3)*1024 + (r&7)*128, 16-byte units XOR-ed with r&7).  Read K-major it is
// the [samples x features] A operand of the dgrad MMAs, read MN-major it is the [features x samples] operand of the
// wgrad MMAs -- the same bytes.  Layers 0..8 have 4 chunks (256 features), layer 9 (rgb hidden / its gradient) 2.
__host__ __device__ inline size_t act_layer_off(int layer, size_t n_tiles) { return (size_t)layer * n_tiles * 4 * CHUNK_BYTES; }
__host__ __device__ inline size_t act_bytes(size_t n_tiles) { return (9 * 4 + 2) * n_tiles * (size_t)CHUNK_BYTES; }
__host__ __device__ inline size_t act_chunk_off(int layer, size_t n_tiles, size_t tile, int chunk) {
  return act_layer_off(layer, n_tiles) + (tile * (layer == 9 ? 2 : 4) + chunk) * (size_t)CHUNK_BYTES;
}

// training: one thread's 32 fp16 activations (row `row`, K columns [32 hh, 32 hh + 32) of a chunk) into the chunk's
// SWIZZLE_128B image in global memory: four 16-byte units, 64 contiguous bytes after the XOR
__device__ __forceinline__ void store_act_chunk(uint8_t* chunk, int row, int hh, const uint32_t (&pk)[16]) {
  uint4* rowp = reinterpret_cast<uint4*>(chunk + (row >> 3) * 1024 + (row & 7) * 128);
#pragma unroll
  for (int u = 0; u < 4; ++u) rowp[(4 * hh + u) ^ (row & 7)] = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
}

// The same store through shared memory and ONE bulk copy per warp pair.  store_act_chunk's st.global.v4 has every lane
// write 16 bytes of a different 128-byte line: 32 line transactions per instruction, 4096 per layer and tile, and the
// LSU retires about one per cycle -- measured as +4 000 cycles per layer in both kernels that save a tile this way.
// Here warps w and w+4 (the two column halves of rows [32q, 32q+32)) write their 4 KB slice of the chunk image into a
// shared-memory buffer (conflict-free: the image is 128-byte swizzled), meet on a 64-thread named barrier, and one
// thread issues cp.async.bulk shared -> global for the slice.  Two buffers per pair alternate; the issuer waits for its
// previous copy to have READ its buffer before the barrier, so after the barrier the other buffer is known to be free.
// Call sequence must be identical in both warps of a pair (it is: one call per chunk).  Named barriers 2..5.
constexpr int STG_SLICE_BYTES = 4096;                       // 32 rows x 128 B
constexpr int STG_BYTES = 4 * 2 * STG_SLICE_BYTES;          // 4 pairs x 2 buffers
__device__ __forceinline__ void stage_store_chunk(uint8_t* stg, uint32_t& flip, uint8_t* gchunk, int q, int lane, int hh,
                                                  const uint32_t (&pk)[16], bool no_copy = false) {
  uint8_t* sb = stg + (q * 2 + (int)flip) * STG_SLICE_BYTES;
  flip ^= 1u;
  uint4* rowp = reinterpret_cast<uint4*>(sb + (lane >> 3) * 1024 + (lane & 7) * 128);
#pragma unroll
  for (int u = 0; u < 4; ++u) rowp[(4 * hh + u) ^ (lane & 7)] = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
  fence_proxy_async();                                       // generic-proxy writes -> visible to the bulk copy engine
  const bool issuer = (hh == 0) && (lane == 0) && !no_copy;     // no_copy: timing experiment (nothing leaves the SM)
  if (issuer) bulk_s2g_wait_read();                          // my previous slice (the other buffer) has been read out
  asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
  if (issuer) bulk_s2g(gchunk + q * STG_SLICE_BYTES, smem_u32(sb), STG_SLICE_BYTES);
}
// before the kernel ends (shared memory is released with the CTA) every issuer drains its copies
__device__ __forceinline__ void stage_store_drain() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// What a training-mode forward saves for the backward (all NULL = inference): the fp16 activations of every layer
// (ACT layout above; layers 0..7 = base outputs after ReLU, 8 = base_remap output, 9 = rgb hidden after ReLU), the
// E operand tile of every sample tile ([tile][2 chunks]) and the sigma head's output before the abs().
// ... and the ReLU masks of base layers 0..7 as bits (MASK layout below): the data-gradient chain needs only the sign of
// those activations, 4 KB per layer and tile instead of the 64 KB of fp16 it would otherwise re-read.
struct TrainSave { uint8_t* act; uint8_t* e; float* raw_sigma; uint8_t* mask; };
__host__ __device__ inline size_t train_ws_e_off(size_t n_tiles) { return act_bytes(n_tiles); }
__host__ __device__ inline size_t train_ws_sigma_off(size_t n_tiles) { return act_bytes(n_tiles) + n_tiles * 2 * (size_t)CHUNK_BYTES; }
__host__ __device__ inline size_t train_ws_mask_off(size_t n_tiles) { return train_ws_sigma_off(n_tiles) + n_tiles * TILE * sizeof(float); }
// MASK layout: [layer 0..7][tile][hh 0..1][row 0..127] uint4 = the four chunk words of one epilogue thread (row, column
// half hh of every 64-column chunk), written and read as one coalesced 16-byte access per thread
constexpr int MASK_LAYERS = 8;
__host__ __device__ inline size_t mask_off(int layer, size_t n_tiles, size_t tile, int hh, int row) {
  return (((size_t)layer * n_tiles + tile) * 2 + (size_t)hh) * (TILE * 16) + (size_t)row * 16;
}
__host__ __device__ inline size_t train_ws_bytes(size_t n_tiles) { return train_ws_mask_off(n_tiles) + MASK_LAYERS * n_tiles * 2 * (size_t)(TILE * 16); }

// One chunk's ReLU mask as a word: the bit of packed pair e's low / high half sits where  (word << (e >> 1))  exposes it
// as the sign bit of byte 0 / 2 (e even) or byte 1 / 3 (e odd), so that ONE prmt with sign replication per pair turns it
// back into a 0xffff / 0x0000 mask per half (apply_relu_mask).
__device__ __forceinline__ uint32_t relu_mask_bits(const uint32_t (&pk)[16]) {
  uint32_t m = 0u;
  const __half2 zero2 = __float2half2_rn(0.f);
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const int sft = e >> 1;
    const uint32_t pos = (e & 1) ? ((1u << (15 - sft)) | (1u << (31 - sft))) : ((1u << (7 - sft)) | (1u << (23 - sft)));
    m |= __hgt2_mask(*reinterpret_cast<const __half2*>(&pk[e]), zero2) & pos;
  }
  return m;
}
__device__ __forceinline__ void apply_relu_mask(uint32_t (&pk)[16], uint32_t bits) {
#pragma unroll
  for (int e2 = 0; e2 < 8; ++e2) {
    const uint32_t sh = bits << e2;
    uint32_t m0, m1;
    asm("prmt.b32 %0, %1, %1, 0xAA88;" : "=r"(m0) : "r"(sh));
    asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(m1) : "r"(sh));
    pk[2 * e2] &= m0;
    pk[2 * e2 + 1] &= m1;
  }
}

}  // namespace tc
}  // namespace npp
