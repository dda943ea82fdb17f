// Ray sampling kernels: unit-sphere intersection, stratified coarse
// depths, inverse-CDF importance sampling and the sorted merge of old and new depths.
// One warp per ray; the search and the rank-merge run out of shared memory with warp shuffles
// for the reductions/scans.  All arithmetic that decides a sample position uses explicitly
// rounded fp32 ops (no FMA contraction) so positions are bit-identical to the reference's
// op-by-op torch arithmetic given the same cdf (R7/H2).
#include <cstdio>
#include "common.cuh"

namespace npp {

// ---- A1 ---------------------------------------------------------------------------------------
// ddp_train_nerf.py:51-66
__global__ void intersect_sphere_kernel(const float* __restrict__ ray_o, const float* __restrict__ ray_d, int n,
                                        float* __restrict__ out_far, int32_t* __restrict__ out_unbounded) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float o0 = ray_o[3 * i], o1 = ray_o[3 * i + 1], o2 = ray_o[3 * i + 2];
  float d0 = ray_d[3 * i], d1_ = ray_d[3 * i + 1], d2_ = ray_d[3 * i + 2];
  float dO = __fadd_rn(__fadd_rn(__fmul_rn(d0, o0), __fmul_rn(d1_, o1)), __fmul_rn(d2_, o2));
  float dd = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1_, d1_)), __fmul_rn(d2_, d2_));
  float d1 = __fdiv_rn(-dO, dd);
  float p0 = __fadd_rn(o0, __fmul_rn(d1, d0)), p1 = __fadd_rn(o1, __fmul_rn(d1, d1_)), p2 = __fadd_rn(o2, __fmul_rn(d1, d2_));
  // torch.norm (CPU) accumulates x*x with an FMA chain: sqrt(fma(z,z,fma(y,y,x*x)))
  float inv_len = __fdiv_rn(1.f, __fsqrt_rn(__fmaf_rn(d2_, d2_, __fmaf_rn(d1_, d1_, __fmul_rn(d0, d0)))));
  float psq = __fadd_rn(__fadd_rn(__fmul_rn(p0, p0), __fmul_rn(p1, p1)), __fmul_rn(p2, p2));
  if (psq >= 1.f) atomicOr(out_unbounded, 1);
  float d2 = __fmul_rn(__fsqrt_rn(__fsub_rn(1.f, psq)), inv_len);
  out_far[i] = __fadd_rn(d1, d2);
}

// ---- A2 + A3 ----------------------------------------------------------------------------------
// ddp_train_nerf.py:441-449 (train) / 168-175 (test), perturb_samples :69-78
__device__ __forceinline__ float perturb(float zl, float z, float zr, bool first, bool last, float t) {
  float lower = first ? z : __fmul_rn(0.5f, __fadd_rn(z, zl));
  float upper = last ? z : __fmul_rn(0.5f, __fadd_rn(zr, z));
  return __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t));
}
__global__ void coarse_depths_kernel(const float* __restrict__ near, const float* __restrict__ far,
                                     const float* __restrict__ bg_base, int n, int S,
                                     const float* __restrict__ t_fg, const float* __restrict__ t_bg,
                                     float* __restrict__ out_fg, float* __restrict__ out_bg) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * S) return;
  int r = (int)(idx / S), i = (int)(idx % S);
  float nr = near[r];
  float step = __fdiv_rn(__fsub_rn(far[r], nr), (float)(S - 1));
  auto zf = [&](int j) { return __fadd_rn(nr, __fmul_rn((float)j, step)); };
  float z = zf(i);
  if (t_fg) z = perturb(i > 0 ? zf(i - 1) : 0.f, z, i < S - 1 ? zf(i + 1) : 0.f, i == 0, i == S - 1, t_fg[idx]);
  out_fg[idx] = z;
  float b = bg_base[i];
  if (t_bg) b = perturb(i > 0 ? bg_base[i - 1] : 0.f, b, i < S - 1 ? bg_base[i + 1] : 0.f, i == 0, i == S - 1, t_bg[idx]);
  out_bg[idx] = b;
}

// generic perturb_samples (ddp_train_nerf.py:69-78) for caller-supplied depths
__global__ void perturb_kernel(const float* __restrict__ zin, const float* __restrict__ t, int n, int S, float* __restrict__ out) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * S) return;
  int i = (int)(idx % S);
  out[idx] = perturb(i > 0 ? zin[idx - 1] : 0.f, zin[idx], i < S - 1 ? zin[idx + 1] : 0.f, i == 0, i == S - 1, t[idx]);
}

// ---- A4 ---------------------------------------------------------------------------------------
// cdf of one ray into shared memory: w+1e-6, / sum, cumsum with a leading 0 (ddp_train_nerf.py:90-93).
// Sum and running sum are carried in fp64 and rounded to fp32 per element -- the accumulation the
// reference's CPU torch.cumsum performs.  Warp-cooperative, any M.
__device__ void warp_build_cdf(const float* __restrict__ w, int M, float* __restrict__ cdf_s, int lane) {
  double part = 0.0;
  for (int j = lane; j < M; j += 32) part += (double)__fadd_rn(w[j], NPP_TINY);
  float total = (float)warp_sum(part);
  double carry = 0.0;
  if (lane == 0) cdf_s[0] = 0.f;
  for (int base = 0; base < M; base += 32) {
    int j = base + lane;
    double v = (j < M) ? (double)__fdiv_rn(__fadd_rn(w[j], NPP_TINY), total) : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    v += carry;
    if (j < M) cdf_s[j + 1] = (float)v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
  __syncwarp();
}

// above = #{j < M : u >= cdf[j]} (ddp_train_nerf.py:111) == upper_bound over the sorted cdf[0..M)
__device__ __forceinline__ int cdf_above(const float* __restrict__ cdf_s, int M, float u) {
  int lo = 0, hi = M;   // first index with cdf[idx] > u
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (u >= cdf_s[mid]) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// ddp_train_nerf.py:114-128
__device__ __forceinline__ float cdf_interp(const float* __restrict__ cdf_s, const float* __restrict__ bins_s, int above, float u) {
  int below = max(above - 1, 0);
  float c0 = cdf_s[below], c1 = cdf_s[above];
  float b0 = bins_s[below], b1 = bins_s[above];
  float denom = __fsub_rn(c1, c0);
  if (denom < NPP_TINY) denom = 1.f;
  float t = __fdiv_rn(__fsub_rn(u, c0), denom);
  return __fadd_rn(b0, __fmul_rn(t, __fadd_rn(__fsub_rn(b1, b0), NPP_TINY)));
}

// warps per block; each warp owns smem_per_warp floats of dynamic shared memory
constexpr int SAMP_WARPS = 4;

// mode 0: cdf from weights; mode 1: cdf supplied
template <int MODE>
__global__ void __launch_bounds__(SAMP_WARPS * 32)
sample_pdf_kernel(const float* __restrict__ bins, int bins_ld, const float* __restrict__ wc, int wc_ld,
                  const float* __restrict__ u, int u_ld, int n, int M, int Ns, float* __restrict__ out,
                  int32_t* __restrict__ out_above, float* __restrict__ out_cdf) {
  extern __shared__ float smem[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int r = blockIdx.x * SAMP_WARPS + warp;
  if (r >= n) return;
  float* cdf_s = smem + warp * (2 * (M + 1));
  float* bins_s = cdf_s + (M + 1);
  for (int j = lane; j <= M; j += 32) bins_s[j] = bins[(size_t)r * bins_ld + j];
  if (MODE == 0) {
    warp_build_cdf(wc + (size_t)r * wc_ld, M, cdf_s, lane);
  } else {
    for (int j = lane; j <= M; j += 32) cdf_s[j] = wc[(size_t)r * wc_ld + j];
    __syncwarp();
  }
  if (out_cdf) for (int j = lane; j <= M; j += 32) out_cdf[(size_t)r * (M + 1) + j] = cdf_s[j];
  for (int i = lane; i < Ns; i += 32) {
    float ui = u[(size_t)r * u_ld + i];
    int above = cdf_above(cdf_s, M, ui);
    out[(size_t)r * Ns + i] = cdf_interp(cdf_s, bins_s, above, ui);
    if (out_above) out_above[(size_t)r * Ns + i] = above;
  }
}

// ---- A4 + A5 fused ----------------------------------------------------------------------------
// ddp_train_nerf.py:452-457: mids -> sample_pdf(w[1:-1]) -> sort(cat(z_prev, z_new)).
// Only the sorted VALUES are returned (:457 drops torch.sort's indices), so any exact sorting network gives the
// reference's result: a warp-wide bitonic sort in shared memory over the next power of two (padding = +inf),
// 36 compare-exchange stages of 4 pairs per lane at 192 values, instead of the O(S^2) rank count it replaces.
// One launch serves the foreground and the background net of a level (blockIdx.y): 4096 rays are 4096 warps, less
// than half of the 9472 warps a B200 holds, so the two halves run side by side.
struct ResampleJob { const float* z_prev; const float* w_prev; const float* u; float* out_z; };
__global__ void __launch_bounds__(SAMP_WARPS * 32)
resample_merge_kernel(ResampleJob job0, ResampleJob job1, int u_ld, int n, int Sp, int Ns, int P) {
  extern __shared__ float smem[];
  const ResampleJob job = blockIdx.y == 0 ? job0 : job1;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int r = blockIdx.x * SAMP_WARPS + warp;
  if (r >= n) return;
  const int M = Sp - 2, St = Sp + Ns;
  float* cdf_s = smem + warp * (2 * Sp + P);   // M+1 = Sp-1 used
  float* bins_s = cdf_s + Sp;
  float* all_s = bins_s + Sp;                  // [P]: Sp old, Ns new, padding
  const float* zr = job.z_prev + (size_t)r * Sp;
  for (int j = lane; j < Sp; j += 32) all_s[j] = zr[j];
  for (int j = St + lane; j < P; j += 32) all_s[j] = __int_as_float(0x7f800000);
  __syncwarp();
  for (int j = lane; j < Sp - 1; j += 32) bins_s[j] = __fmul_rn(0.5f, __fadd_rn(all_s[j + 1], all_s[j]));
  warp_build_cdf(job.w_prev + (size_t)r * Sp + 1, M, cdf_s, lane);
  for (int i = lane; i < Ns; i += 32) {
    float ui = job.u[(size_t)r * u_ld + i];
    all_s[Sp + i] = cdf_interp(cdf_s, bins_s, cdf_above(cdf_s, M, ui), ui);
  }
  __syncwarp();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (P >> 1); t += 32) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));      // the pair's lower index: bit j clear
        const int l = i | j;
        const float a = all_s[i], b = all_s[l];
        const bool up = (i & k) == 0;
        if ((a > b) == up && a != b) { all_s[i] = b; all_s[l] = a; }
      }
      __syncwarp();
    }
  }
  float* o = job.out_z + (size_t)r * St;
  for (int i = lane; i < St; i += 32) o[i] = all_s[i];
}

// ---- N1: ray generation + batch gather (nerf_sample_ray_split.py:10-34, 155-221) ------------------------------------
// ray_d = R_c2w (K^-1 [u+.5, v+.5, 1]) evaluated per selected pixel instead of for the whole image on the host; the
// per-image arrays the trainer gathers with numpy fancy indexing (rgb, depth prior, min depth) are gathered here too.
struct RayCam { float kinv[9]; float rot[9]; float org[3]; float depth; };
__global__ void gen_rays_kernel(RayCam cam, int W, const long long* __restrict__ ids, long long n, const float* __restrict__ img,
                                const float* __restrict__ img_depth_sup, const float* __restrict__ img_min_depth,
                                float* __restrict__ ray_o, float* __restrict__ ray_d, float* __restrict__ depth, float* __restrict__ rgb,
                                float* __restrict__ depth_sup, float* __restrict__ min_depth) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long id = ids ? ids[i] : i;
  const float u = (float)(id % W) + 0.5f, v = (float)(id / W) + 0.5f;     // :19-21, add half pixel
  float c[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) c[r] = __fmaf_rn(cam.kinv[3 * r + 2], 1.f, __fmaf_rn(cam.kinv[3 * r + 1], v, __fmul_rn(cam.kinv[3 * r], u)));
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    ray_d[3 * i + r] = __fmaf_rn(cam.rot[3 * r + 2], c[2], __fmaf_rn(cam.rot[3 * r + 1], c[1], __fmul_rn(cam.rot[3 * r], c[0])));
    ray_o[3 * i + r] = cam.org[r];
  }
  if (depth) depth[i] = cam.depth;
  if (rgb && img) { rgb[3 * i] = img[3 * id]; rgb[3 * i + 1] = img[3 * id + 1]; rgb[3 * i + 2] = img[3 * id + 2]; }
  if (depth_sup && img_depth_sup) depth_sup[i] = img_depth_sup[id];
  if (min_depth) min_depth[i] = img_min_depth ? img_min_depth[id] : 1e-4f;   // :194-197
}

}  // namespace npp

using namespace npp;

extern "C" int nerfpp_intersect_sphere(const float* ray_o, const float* ray_d, int n_rays, float* out_far,
                                       int32_t* out_unbounded, void* stream) {
  NPP_CHECK_ARG(n_rays >= 0 && ray_o && ray_d && out_far && out_unbounded, "bad argument");
  if (n_rays == 0) return 0;
  intersect_sphere_kernel<<<(n_rays + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ray_o, ray_d, n_rays, out_far, out_unbounded);
  NPP_CHECK_LAUNCH();
  return 0;
}

extern "C" int nerfpp_coarse_depths(const float* near, const float* far, const float* bg_base, int n_rays, int n_samples,
                                    const float* t_rand_fg, const float* t_rand_bg, float* out_fg_z, float* out_bg_z,
                                    void* stream) {
  NPP_CHECK_ARG(n_rays >= 0 && n_samples >= 2 && near && far && bg_base && out_fg_z && out_bg_z, "bad argument");
  if (n_rays == 0) return 0;
  long long tot = (long long)n_rays * n_samples;
  coarse_depths_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(near, far, bg_base, n_rays, n_samples,
                                                                                     t_rand_fg, t_rand_bg, out_fg_z, out_bg_z);
  NPP_CHECK_LAUNCH();
  return 0;
}

extern "C" int nerfpp_perturb_samples(const float* z, const float* t_rand, int n_rays, int n_samples, float* out_z, void* stream) {
  NPP_CHECK_ARG(n_rays >= 0 && n_samples >= 1 && z && t_rand && out_z && out_z != z, "bad argument");
  if (n_rays == 0) return 0;
  long long tot = (long long)n_rays * n_samples;
  perturb_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(z, t_rand, n_rays, n_samples, out_z);
  NPP_CHECK_LAUNCH();
  return 0;
}

static int sample_common(int mode, const float* bins, int bins_ld, const float* wc, int wc_ld, const float* u, int u_ld,
                         int n, int M, int Ns, float* out, int32_t* out_above, float* out_cdf, void* stream) {
  NPP_CHECK_ARG(n >= 0 && M >= 1 && Ns >= 1 && bins && wc && u && out, "bad argument");
  NPP_CHECK_ARG(M <= 4096, "M too large");
  if (n == 0) return 0;
  size_t smem = (size_t)SAMP_WARPS * 2 * (M + 1) * sizeof(float);
  dim3 grid((n + SAMP_WARPS - 1) / SAMP_WARPS), block(SAMP_WARPS * 32);
  if (mode == 0) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(sample_pdf_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sample_pdf_kernel<0><<<grid, block, smem, (cudaStream_t)stream>>>(bins, bins_ld, wc, wc_ld, u, u_ld, n, M, Ns, out, out_above, out_cdf);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(sample_pdf_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sample_pdf_kernel<1><<<grid, block, smem, (cudaStream_t)stream>>>(bins, bins_ld, wc, wc_ld, u, u_ld, n, M, Ns, out, out_above, out_cdf);
  }
  NPP_CHECK_LAUNCH();
  return 0;
}

extern "C" int nerfpp_sample_pdf(const float* bins, int bins_ld, const float* weights, int w_ld, const float* u, int u_ld,
                                 int n_rays, int M, int n_new, float* out_samples, int32_t* out_above, float* out_cdf,
                                 void* stream) {
  return sample_common(0, bins, bins_ld, weights, w_ld, u, u_ld, n_rays, M, n_new, out_samples, out_above, out_cdf, stream);
}

extern "C" int nerfpp_sample_cdf(const float* bins, int bins_ld, const float* cdf, int cdf_ld, const float* u, int u_ld,
                                 int n_rays, int M, int n_new, float* out_samples, int32_t* out_above, void* stream) {
  return sample_common(1, bins, bins_ld, cdf, cdf_ld, u, u_ld, n_rays, M, n_new, out_samples, out_above, nullptr, stream);
}

static int resample_launch(const ResampleJob& j0, const ResampleJob& j1, int jobs, int u_ld, int n_rays, int n_prev, int n_new,
                           void* stream) {
  NPP_CHECK_ARG(n_rays >= 0 && n_prev >= 3 && n_new >= 1, "bad argument");
  NPP_CHECK_ARG(n_prev + n_new <= 4096, "too many samples per ray");
  if (n_rays == 0) return 0;
  int P = 2;
  while (P < n_prev + n_new) P <<= 1;
  size_t smem = (size_t)SAMP_WARPS * (2 * n_prev + P) * sizeof(float);
  if (smem > 48 * 1024) cudaFuncSetAttribute(resample_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((n_rays + SAMP_WARPS - 1) / SAMP_WARPS, jobs);
  resample_merge_kernel<<<grid, SAMP_WARPS * 32, smem, (cudaStream_t)stream>>>(j0, j1, u_ld, n_rays, n_prev, n_new, P);
  NPP_CHECK_LAUNCH();
  return 0;
}

extern "C" int nerfpp_resample_merge(const float* z_prev, const float* w_prev, const float* u, int u_ld, int n_rays,
                                     int n_prev, int n_new, float* out_z, void* stream) {
  NPP_CHECK_ARG(z_prev && w_prev && u && out_z, "null argument");
  ResampleJob j{z_prev, w_prev, u, out_z};
  return resample_launch(j, j, 1, u_ld, n_rays, n_prev, n_new, stream);
}

extern "C" int nerfpp_resample_merge_pair(const float* fg_z_prev, const float* fg_w_prev, const float* fg_u, float* out_fg_z,
                                          const float* bg_z_prev, const float* bg_w_prev, const float* bg_u, float* out_bg_z,
                                          int u_ld, int n_rays, int n_prev, int n_new, void* stream) {
  NPP_CHECK_ARG(fg_z_prev && fg_w_prev && fg_u && out_fg_z && bg_z_prev && bg_w_prev && bg_u && out_bg_z, "null argument");
  ResampleJob j0{fg_z_prev, fg_w_prev, fg_u, out_fg_z}, j1{bg_z_prev, bg_w_prev, bg_u, out_bg_z};
  return resample_launch(j0, j1, 2, u_ld, n_rays, n_prev, n_new, stream);
}

extern "C" int nerfpp_gen_rays(const float* kinv_host, const float* c2w_host, float cam_depth, int W, const int64_t* pixel_ids,
                               int64_t n, const float* img_rgb, const float* img_depth_sup, const float* img_min_depth, float* ray_o,
                               float* ray_d, float* depth, float* rgb, float* depth_sup, float* min_depth, void* stream) {
  NPP_CHECK_ARG(kinv_host && c2w_host && ray_o && ray_d && W >= 1 && n >= 0, "bad argument");
  if (n == 0) return 0;
  RayCam cam;
  for (int i = 0; i < 9; ++i) cam.kinv[i] = kinv_host[i];
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) cam.rot[3 * r + c] = c2w_host[4 * r + c]; cam.org[r] = c2w_host[4 * r + 3]; }
  cam.depth = cam_depth;
  gen_rays_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(cam, W, (const long long*)pixel_ids, n, img_rgb, img_depth_sup,
                                                                              img_min_depth, ray_o, ray_d, depth, rgb, depth_sup, min_depth);
  NPP_CHECK_LAUNCH();
  return 0;
}

// ---- N2: pixel decode of the on-disk formats (nerf_sample_ray_split.py:73-102) -------------------------------------
// out = ((x / div) * mul) + add with every step rounded to fp32, the order numpy evaluates
//   rgb / mask  imread(...).astype(float32) / 255.                      (div 255, mul 1, add 0)      :73,80
//   min depth   imread(...).astype(float32) / 255. * max_depth + 1e-4   (div 255, mul max, add 1e-4) :87
//   depth       depth_scale * (imread(...).astype(float32) / 256.0)     (div 256, mul scale, add 0)  :94-102
// (x * 1 and x + 0 are exact, so one kernel serves all three.)
template <typename T>
__global__ void decode_pixels_kernel(const T* __restrict__ src, long long n, float div, float mul, float add, float* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __fadd_rn(__fmul_rn(__fdiv_rn((float)src[i], div), mul), add);
}

extern "C" int nerfpp_decode_pixels(const void* src, int src_bits, int64_t n, float div, float mul, float add, float* out, void* stream) {
  NPP_CHECK_ARG(n >= 0 && (src_bits == 8 || src_bits == 16) && div != 0.f, "bad argument");
  if (n == 0) return 0;                         // an empty image: nothing to launch (the pointers may be null)
  NPP_CHECK_ARG(src && out, "null pointer");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (src_bits == 8) decode_pixels_kernel<uint8_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)src, n, div, mul, add, out);
  else decode_pixels_kernel<uint16_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)src, n, div, mul, add, out);
  NPP_CHECK_LAUNCH();
  return 0;
}
