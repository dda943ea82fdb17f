// Config 3: the mipnerf360 field (models.py:200-231 per level; MLP.__call__ models.py:398-611) -- ray warp, conical
// frusta, contraction with linearised covariances, integrated positional encoding, then the Dense stack as tcgen05
// GEMM launches (gemm_tc.cu) and the two small heads on CUDA cores.
#include <cuda_fp16.h>
#include <float.h>
#include "gemm_tc.cuh"

namespace npp {
namespace m360 {

constexpr int NB = 21;                 // basis vectors (icosahedron, 2 subdivisions, symmetric copies removed)
constexpr int NDEG = 12;               // integrated_pos_enc degrees 0..11
constexpr int NPAIR = NB * NDEG;       // 252 (sin, cos) pairs -> 504 features
constexpr int ENC_LD = 512;            // row pitch of the encoding (504 + 8 zero columns)
constexpr int DIR_LD = 64;             // row pitch of the view-direction encoding (27 + zeros)
constexpr int DIR_DEG = 4;
constexpr int BOTTLENECK = 256, VIEW_W = 128;
constexpr int SPB = 128;               // samples per block of the encode kernel

// geopoly.generate_basis('icosahedron', 2) (geopoly.py:80-126), rows in the reference's order (golden table in
// tests/geopoly_test.py:79-99): a = 1/sqrt(1+phi^2), c = phi a, p = (phi-1)/2, q = phi/2
#define M360_A 0.5257311121191336
#define M360_C 0.85065080835204
#define M360_P 0.30901699437494745
#define M360_Q 0.8090169943749475
__constant__ float c_basis[NB][3] = {
    {(float)M360_C, 0.f, (float)M360_A},       {(float)M360_Q, 0.5f, (float)M360_P},    {(float)M360_A, (float)M360_C, 0.f},
    {1.f, 0.f, 0.f},                           {(float)M360_Q, 0.5f, (float)-M360_P},   {(float)M360_C, 0.f, (float)-M360_A},
    {(float)M360_P, (float)M360_Q, -0.5f},     {0.f, (float)M360_A, (float)-M360_C},    {0.5f, (float)M360_P, (float)-M360_Q},
    {0.f, 1.f, 0.f},                           {(float)-M360_A, (float)M360_C, 0.f},    {(float)-M360_P, (float)M360_Q, -0.5f},
    {0.f, (float)M360_A, (float)M360_C},       {(float)-M360_P, (float)M360_Q, 0.5f},   {(float)M360_P, (float)M360_Q, 0.5f},
    {0.5f, (float)M360_P, (float)M360_Q},      {0.5f, (float)-M360_P, (float)M360_Q},   {0.f, 0.f, 1.f},
    {-0.5f, (float)M360_P, (float)M360_Q},     {(float)-M360_Q, 0.5f, (float)M360_P},   {(float)-M360_Q, 0.5f, (float)-M360_P}};

// math.safe_sin (math.py:27-40): sin(where(|x| < 100 pi, x, x % (100 pi))), jnp's `%` being the floor modulus.
// The modulus is exact, as fmod is, but costs eight instructions instead of fmodf's loop: q = floor(x / T) from an IEEE
// division can only be one too LARGE (rounding to nearest never crosses an integer downwards), so r = x - q T lies in
// (-T, T); x and q T are multiples of ulp(T) = 2^-15 and |r| < 512, hence r is representable and the FMA returns it exactly,
// and so is the correction r + T.
// PRECISE (split-precision mode, whose features keep ~1e-7 through their low halves): the exact modulus and sin_cw (~1 ulp).
// Otherwise (features rounded to fp16, half an ulp = 1.2e-4): six instructions -- the modulus with a reciprocal multiply
// (a quotient off by one shifts the phase by T - 100 pi = 5.9e-6 only, T being 50 periods to that accuracy) and the SFU's
// sin on the reduced argument (|x| < 2 T: 1 / 2 pi scaling error <= 4e-5 rad, SFU 4e-7).
__device__ __forceinline__ float sin_sfu(float x) { return __sinf(x); }          // |x| < 8: view directions
// |x| < 2 * 100 pi: two-FMA Cody-Waite reduction by pi/2 (the first FMA is exact: both terms are multiples of 2^-23 and the
// difference is below 1) and the single-precision minimax polynomials for sin / cos on [-pi/4, pi/4] (~1 ulp): libm's sinf
// without its large-argument path, less than half the instructions
__device__ __forceinline__ float sin_cw(float x) {
  const float kf = rintf(x * 0.636619772367581f);
  float r = __fmaf_rn(-kf, 1.5707963705062866f, x);
  r = __fmaf_rn(-kf, -4.371138828673793e-8f, r);
  const int k = (int)kf;
  const float s = r * r;
  float v;
  if (k & 1) {
    v = fmaf(s, fmaf(s, fmaf(s, 2.443315711809948e-5f, -1.388731625493765e-3f), 4.166664568298827e-2f), -0.5f);
    v = fmaf(v, s, 1.f);
  } else {
    v = fmaf(s, fmaf(s, -1.9515295891e-4f, 8.3321608736e-3f), -1.6666654611e-1f);
    v = fmaf(v * s, r, r);
  }
  return (k & 2) ? -v : v;
}
template <bool PRECISE>
__device__ __forceinline__ float safe_sin(float x) {
  const float T = 314.15927f;      // float32(100 * pi)
  if (PRECISE) {
    if (!(fabsf(x) < T)) {
      const float q = floorf(__fdiv_rn(x, T));
      float r = __fmaf_rn(-q, T, x);
      if (r < 0.f) r = __fadd_rn(r, T);
      x = r;
    }
    return sin_cw(x);
  }
  if (fabsf(x) >= T) x = __fmaf_rn(-floorf(x * 0.0031830987f), T, x);
  return __sinf(x);
}

// coord.construct_ray_warps(reciprocal, near, far)[1] (coord.py:92-98): s_to_t(s) = 1 / (s / far + (1 - s) / near), every
// operation rounded separately as numpy / XLA do (no FMA contraction) so that the fenceposts are bit-exact
__device__ __forceinline__ float s_to_t(float s, float s_near, float s_far) {
  return __frcp_rn(__fadd_rn(__fmul_rn(s, s_far), __fmul_rn(__fsub_rn(1.f, s), s_near)));
}

struct Gauss { float mean[3]; float cov[3][3]; };

// Every operation below is rounded on its own (__f*_rn intrinsics are never contracted into FMAs) and sums run left to
// right: the evaluation order of the test suite's CPU restatement (lifted_gaussians_ordered), operation for operation.
// The path is badly conditioned -- feature sin(2^11 x) turns one ulp of a lifted mean into 5e-4, J cov J^T cancels ten
// digits for distant samples -- so "the same arithmetic" has to mean the same order, not just the same formula.
#define MUL(a, b) __fmul_rn((a), (b))
#define ADD(a, b) __fadd_rn((a), (b))
#define SUB(a, b) __fsub_rn((a), (b))
#define DIV(a, b) __fdiv_rn((a), (b))

// render.conical_frustum_to_gaussian (stable form, render.py:42-66) + lift_gaussian (:21-39, diag=False) + origin
__device__ __forceinline__ Gauss cast_cone(float t0, float t1, const float o[3], const float d[3], float radius) {
  const float eps = FLT_EPSILON;
  const float mu = DIV(ADD(t0, t1), 2.f), hw = DIV(SUB(t1, t0), 2.f);
  const float mu2 = MUL(mu, mu), hw2 = MUL(hw, hw), hw4 = MUL(hw2, hw2);
  const float denom = fmaxf(eps, ADD(MUL(3.f, mu2), hw2));
  const float t_mean = ADD(mu, DIV(MUL(MUL(2.f, mu), hw2), denom));
  const float t_var = SUB(DIV(hw2, 3.f), DIV(MUL(MUL((float)(4.0 / 15.0), hw4), SUB(MUL(12.f, mu2), hw2)), MUL(denom, denom)));
  float r_var = SUB(ADD(DIV(mu2, 4.f), MUL((float)(5.0 / 12.0), hw2)), DIV(MUL((float)(4.0 / 15.0), hw4), denom));
  r_var = MUL(r_var, MUL(radius, radius));
  const float dms = fmaxf(1e-10f, ADD(ADD(MUL(d[0], d[0]), MUL(d[1], d[1])), MUL(d[2], d[2])));
  Gauss g;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    g.mean[i] = ADD(MUL(d[i], t_mean), o[i]);
#pragma unroll
    for (int j = 0; j < 3; ++j)
      g.cov[i][j] = ADD(MUL(t_var, MUL(d[i], d[j])), MUL(r_var, SUB(i == j ? 1.f : 0.f, MUL(d[i], DIV(d[j], dms)))));
  }
  return g;
}

// coord.track_linearize(coord.contract, mean, cov) (coord.py:22-60): z = contract(x), cov' = J cov J^T with the Jacobian
// J = scale I + (k x) x^T outside the unit ball (s = |x|^2, scale = (2 sqrt(s) - 1) / s, k = 2 (1 - sqrt(s)) / s^2), I inside
__device__ __forceinline__ void contract_linearize(Gauss& g) {
  const float* x = g.mean;
  const float s = fmaxf(FLT_EPSILON, ADD(ADD(MUL(x[0], x[0]), MUL(x[1], x[1])), MUL(x[2], x[2])));
  if (s <= 1.f) return;
  const float rs = __fsqrt_rn(s);
  const float scale = DIV(SUB(MUL(2.f, rs), 1.f), s);
  const float k = DIV(MUL(2.f, SUB(1.f, rs)), MUL(s, s));
  float J[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) J[i][j] = ADD(i == j ? scale : 0.f, MUL(MUL(k, x[i]), x[j]));
  float t[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) t[i][j] = ADD(ADD(MUL(J[i][0], g.cov[0][j]), MUL(J[i][1], g.cov[1][j])), MUL(J[i][2], g.cov[2][j]));
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) g.cov[i][j] = ADD(ADD(MUL(t[i][0], J[j][0]), MUL(t[i][1], J[j][1])), MUL(t[i][2], J[j][2]));
  const float z0 = MUL(scale, x[0]), z1 = MUL(scale, x[1]), z2 = MUL(scale, x[2]);
  g.mean[0] = z0; g.mean[1] = z1; g.mean[2] = z2;
}

__device__ __forceinline__ uint32_t pack_hi(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_lo(float a, float b, uint32_t hi) {
  const __half2 h = *reinterpret_cast<const __half2*>(&hi);
  return pack_hi(a - __low2float(h), b - __high2float(h));
}

// One block = SPB consecutive samples, 256 threads.  Phase 1a: a thread per sample casts the frustum and contracts it
// (means and covariances -> shared memory); the other half of the block writes the samples' view-direction rows (one
// evaluation per sample of its ray's 24 sines -- a ray's samples share them, but the rows are what the view layer's TMA
// loads read).  Phase 1b: all threads project onto the basis (21 lifted means / variances per sample).  Phase 2: all
// threads sweep (sample, feature-pair) items so that a warp writes 128 contiguous bytes of a row.
template <bool PRECISE>
__global__ void __launch_bounds__(256) cast_encode_kernel(
    const float* __restrict__ sdist, const float* __restrict__ near, const float* __restrict__ far, const float* __restrict__ origins,
    const float* __restrict__ directions, const float* __restrict__ viewdirs, const float* __restrict__ radii, int n, int S,
    float* __restrict__ out_tdist, __half* __restrict__ enc, __half* __restrict__ enc_lo, __half* __restrict__ dir, __half* __restrict__ dir_lo,
    float* __restrict__ out_means, float* __restrict__ out_covs) {
  __shared__ float lm[SPB][NB + 1], lv[SPB][NB + 1], gs[SPB][13];       // gs: contracted mean (3) + covariance (9)
  const long long M = (long long)n * S;
  const long long g0 = (long long)blockIdx.x * SPB;
  const int tid = (int)threadIdx.x;
  if (tid < SPB) {
    const long long gi = g0 + tid;
    if (gi < M) {
      const int ray = (int)(gi / S), i = (int)(gi % S);
      const float s_near = __frcp_rn(near[ray]), s_far = __frcp_rn(far[ray]);
      const float t0 = s_to_t(sdist[(long long)ray * (S + 1) + i], s_near, s_far);
      const float t1 = s_to_t(sdist[(long long)ray * (S + 1) + i + 1], s_near, s_far);
      if (out_tdist) {
        out_tdist[(long long)ray * (S + 1) + i] = t0;
        if (i == S - 1) out_tdist[(long long)ray * (S + 1) + S] = t1;
      }
      const float o[3] = {origins[3 * ray], origins[3 * ray + 1], origins[3 * ray + 2]};
      const float d[3] = {directions[3 * ray], directions[3 * ray + 1], directions[3 * ray + 2]};
      Gauss g = cast_cone(t0, t1, o, d, radii[ray]);
      contract_linearize(g);
      if (out_means) { for (int k = 0; k < 3; ++k) out_means[gi * 3 + k] = g.mean[k]; }
      if (out_covs) { for (int k = 0; k < 9; ++k) out_covs[gi * 9 + k] = g.cov[k / 3][k % 3]; }
#pragma unroll
      for (int k = 0; k < 3; ++k) gs[tid][k] = g.mean[k];
#pragma unroll
      for (int k = 0; k < 9; ++k) gs[tid][3 + k] = g.cov[k / 3][k % 3];
    }
  } else if (dir != nullptr) {
    // coord.pos_enc(viewdirs, 0, 4, append_identity=True) (coord.py:136-148): [d, sin(d 2^j), sin(d 2^j + pi/2)], 27 wide
    const long long gi = g0 + (tid - SPB);
    if (gi < M) {
      const int ray = (int)(gi / S);
      const float d[3] = {viewdirs[3 * ray], viewdirs[3 * ray + 1], viewdirs[3 * ray + 2]};
      float f[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) f[k] = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) f[k] = d[k];
#pragma unroll
      for (int j = 0; j < DIR_DEG; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float x = d[k] * (float)(1 << j);
          f[3 + j * 3 + k] = PRECISE ? sin_cw(x) : sin_sfu(x);
          f[3 + 3 * DIR_DEG + j * 3 + k] = PRECISE ? sin_cw(x + 1.5707964f) : sin_sfu(x + 1.5707964f);
        }
      uint4* rowp = reinterpret_cast<uint4*>(dir + gi * DIR_LD);
      uint4* rowl = dir_lo ? reinterpret_cast<uint4*>(dir_lo + gi * DIR_LD) : nullptr;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { h[e] = pack_hi(f[8 * u + 2 * e], f[8 * u + 2 * e + 1]); l[e] = pack_lo(f[8 * u + 2 * e], f[8 * u + 2 * e + 1], h[e]); }
        rowp[u] = make_uint4(h[0], h[1], h[2], h[3]);
        if (rowl) rowl[u] = make_uint4(l[0], l[1], l[2], l[3]);
      }
#pragma unroll
      for (int u = 4; u < 8; ++u) { rowp[u] = make_uint4(0, 0, 0, 0); if (rowl) rowl[u] = make_uint4(0, 0, 0, 0); }
    }
  }
  __syncthreads();
  if (enc == nullptr) return;
  // coord.lift_and_diagonalize (coord.py:129-133): mean @ basis, diag(basis^T cov basis); item = (sample, basis vector)
  for (int w = tid; w < SPB * NB; w += 256) {
    const int smp = w / NB, b = w - smp * NB;
    if (g0 + smp >= M) break;
    const float* m = gs[smp];
    const float* c = gs[smp] + 3;
    const float b0 = c_basis[b][0], b1 = c_basis[b][1], b2 = c_basis[b][2];
    lm[smp][b] = ADD(ADD(MUL(m[0], b0), MUL(m[1], b1)), MUL(m[2], b2));
    const float c0 = ADD(ADD(MUL(c[0], b0), MUL(c[1], b1)), MUL(c[2], b2));
    const float c1 = ADD(ADD(MUL(c[3], b0), MUL(c[4], b1)), MUL(c[5], b2));
    const float c2 = ADD(ADD(MUL(c[6], b0), MUL(c[7], b1)), MUL(c[8], b2));
    lv[smp][b] = ADD(ADD(MUL(b0, c0), MUL(b1, c1)), MUL(b2, c2));
  }
  __syncthreads();
  // coord.integrated_pos_enc (coord.py:107-126): feature j*21+b = exp(-0.5 var 4^j) safe_sin(mean 2^j), the second half the
  // same with the argument shifted by float32(pi/2)
  // Thread t of a 128-thread half owns the two adjacent (degree, basis) pairs 2t and 2t+1 for every sample it visits
  // (126 of 128 lanes active): the degree -- hence the slow modulus path of safe_sin -- is uniform over most of a warp,
  // nothing is divided inside the loop, and a warp writes 128 contiguous bytes of a row per store.
  constexpr int ITEMS = NPAIR / 2;        // 126 items of two adjacent pairs per sample
  const int t = tid & 127;
  if (t >= ITEMS) return;
  int jb[2];
  float sc[2], sc2[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int p = 2 * t + e, j = p / NB;
    jb[e] = p - j * NB;
    sc[e] = (float)(1 << j);
    sc2[e] = sc[e] * sc[e];
  }
  for (int smp = tid >> 7; smp < SPB; smp += 2) {
    const long long gi = g0 + smp;
    if (gi >= M) break;
    float sv[2], cv[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float x = MUL(lm[smp][jb[e]], sc[e]);
      const float hv = MUL(-0.5f, MUL(lv[smp][jb[e]], sc2[e]));
      // exp(hv) < 2^-25 rounds to zero in fp16 (and its low half too): skip the two sines
      if (hv < -17.4f) { sv[e] = 0.f; cv[e] = 0.f; continue; }
      const float ex = PRECISE ? expf(hv) : __expf(hv);
      sv[e] = MUL(ex, safe_sin<PRECISE>(x));
      cv[e] = MUL(ex, safe_sin<PRECISE>(ADD(x, 1.5707964f)));
    }
    const uint32_t hs = pack_hi(sv[0], sv[1]), hc = pack_hi(cv[0], cv[1]);
    uint32_t* row = reinterpret_cast<uint32_t*>(enc + gi * ENC_LD);
    row[t] = hs;
    row[ITEMS + t] = hc;
    if (t == 0) *reinterpret_cast<uint4*>(enc + gi * ENC_LD + 2 * NPAIR) = make_uint4(0, 0, 0, 0);
    if (enc_lo) {
      uint32_t* rowl = reinterpret_cast<uint32_t*>(enc_lo + gi * ENC_LD);
      rowl[t] = pack_lo(sv[0], sv[1], hs);
      rowl[ITEMS + t] = pack_lo(cv[0], cv[1], hc);
      if (t == 0) *reinterpret_cast<uint4*>(enc_lo + gi * ENC_LD + 2 * NPAIR) = make_uint4(0, 0, 0, 0);
    }
  }
}

// density head: raw = h . w + b over K fp16 columns (hi [+ lo]); density = softplus(raw + density_bias), density_bias = -1
// (models.py:375,507).  One warp per row, 16-byte loads.
__global__ void __launch_bounds__(256) density_head_kernel(const __half* __restrict__ h, const __half* __restrict__ h_lo, long long M, int K,
                                                           const float* __restrict__ wb, float* __restrict__ out) {
  extern __shared__ float ws[];      // K weights + bias
  for (int i = (int)threadIdx.x; i <= K; i += (int)blockDim.x) ws[i] = wb[i];
  __syncthreads();
  const int lane = (int)(threadIdx.x & 31);
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < M; r += warps) {
    float acc = 0.f;
    for (int k = lane * 8; k < K; k += 256) {
      const uint4 v = *reinterpret_cast<const uint4*>(h + r * K + k);
      const __half2* hp = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(hp[e]); acc += f.x * ws[k + 2 * e] + f.y * ws[k + 2 * e + 1]; }
      if (h_lo) {
        const uint4 vl = *reinterpret_cast<const uint4*>(h_lo + r * K + k);
        const __half2* lp = reinterpret_cast<const __half2*>(&vl);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(lp[e]); acc += f.x * ws[k + 2 * e] + f.y * ws[k + 2 * e + 1]; }
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      const float x = acc + ws[K] - 1.f;
      out[r] = fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));      // jax.nn.softplus = logaddexp(x, 0)
    }
  }
}

// density from the partial head sums the last trunk layer's GEMM epilogue left (fixed summation order: reproducible)
__global__ void density_from_partials_kernel(const float* __restrict__ part, long long M, int P, const float* __restrict__ head_bias,
                                             float* __restrict__ out) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  float acc = 0.f;
  for (int i = 0; i < P; ++i) acc += part[r * P + i];
  const float x = acc + head_bias[0] - 1.f;
  out[r] = fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}

// colour from the partial head sums the view layer's GEMM epilogue left: sigmoid(sum + b) * (1 + 2 pad) - pad (models.py:589-609)
__global__ void rgb_from_partials_kernel(const float* __restrict__ part, long long M, int P, const float* __restrict__ head_bias, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * 3) return;
  const long long r = i / 3;
  const int k = (int)(i - r * 3);
  float acc = 0.f;
  for (int p = 0; p < P; ++p) acc += part[(r * P + p) * 3 + k];
  const float sg = 1.f / (1.f + expf(-(acc + head_bias[k])));
  out[i] = sg * (1.f + 2.f * 0.001f) - 0.001f;
}

// rgb head (models.py:589-609): sigmoid(h_view . W + b) * (1 + 2 pad) - pad, pad = 0.001.  Eight lanes per row.
__global__ void __launch_bounds__(256) rgb_head_kernel(const __half* __restrict__ h, const __half* __restrict__ h_lo, long long M,
                                                       const float* __restrict__ wb, float* __restrict__ out) {
  __shared__ float ws[3 * VIEW_W + 3];       // [3][128] + bias
  for (int i = (int)threadIdx.x; i < 3 * VIEW_W + 3; i += (int)blockDim.x) ws[i] = wb[i];
  __syncthreads();
  const int sub = (int)(threadIdx.x & 7);
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  float a[3] = {0.f, 0.f, 0.f};
  if (r < M) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int k = sub * 16 + u * 8;
      uint4 v = *reinterpret_cast<const uint4*>(h + r * VIEW_W + k);
      const __half2* hp = reinterpret_cast<const __half2*>(&v);
      float f[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 x = __half22float2(hp[e]); f[2 * e] = x.x; f[2 * e + 1] = x.y; }
      if (h_lo) {
        const uint4 vl = *reinterpret_cast<const uint4*>(h_lo + r * VIEW_W + k);
        const __half2* lp = reinterpret_cast<const __half2*>(&vl);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 x = __half22float2(lp[e]); f[2 * e] += x.x; f[2 * e + 1] += x.y; }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int e = 0; e < 8; ++e) a[c] += f[e] * ws[c * VIEW_W + k + e];
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    a[c] += __shfl_xor_sync(0xffffffffu, a[c], 1);
    a[c] += __shfl_xor_sync(0xffffffffu, a[c], 2);
    a[c] += __shfl_xor_sync(0xffffffffu, a[c], 4);
  }
  if (r < M && sub < 3) {
    const float x = a[sub] + ws[3 * VIEW_W + sub];
    const float sg = 1.f / (1.f + expf(-x));
    out[r * 3 + sub] = sg * (1.f + 2.f * 0.001f) - 0.001f;
  }
}

// flax Dense kernel [in, out] fp32 -> fp16 [out, k_pad] (hi, optionally lo), zero beyond `in`
__global__ void pack_dense_kernel(const float* __restrict__ src, int in, int out, int k_pad, __half* __restrict__ hi, __half* __restrict__ lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)out * k_pad) return;
  const int k = (int)(i % k_pad), o = (int)(i / k_pad);
  const float v = k < in ? src[(long long)k * out + o] : 0.f;
  const __half h = __float2half_rn(v);
  hi[i] = h;
  if (lo) lo[i] = __float2half_rn(v - __half2float(h));
}
// head weights: [in, out] -> fp32 [out][in] followed by the out biases
__global__ void pack_head_kernel(const float* __restrict__ kernel, const float* __restrict__ bias, int in, int out, float* __restrict__ dst) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i < in * out) { const int k = i % in, o = i / in; dst[i] = kernel[k * out + o]; }
  else if (i < in * out + out) dst[i] = bias[i - in * out];
}
__global__ void copy_f32_kernel(const float* __restrict__ src, int n, float* __restrict__ dst) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i < n) dst[i] = src[i];
}

__global__ void resample_logits_kernel(const float* __restrict__ sdist, const float* __restrict__ w, long long n, int M, float anneal, float pad,
                                       float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * M) return;
  const long long r = i / M;
  const int k = (int)(i - r * M);
  const float a = sdist[r * (M + 1) + k], b = sdist[r * (M + 1) + k + 1];
  out[i] = (b > a) ? __fmul_rn(anneal, logf(__fadd_rn(w[i], pad))) : -INFINITY;
}

// ---- packed-parameter and workspace layouts (host) ------------------------------------------------------------------
struct Layout {
  int depth, width, has_rgb, prec;
  int k_pad[12];            // padded in-features of the GEMM layers (trunk, bottleneck, view)
  size_t w_hi[12], w_lo[12], bias[12];   // byte offsets; heads: w_hi = fp32 [out][in] + bias
  size_t total;
};
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static bool make_layout(int depth, int width, int has_rgb, int prec, Layout& L) {
  if (!((depth == 4 && width == 256) || (depth == 8 && width == 1024) || (depth == 8 && width == 256) || (depth == 4 && width == 1024))) return false;
  L.depth = depth; L.width = width; L.has_rgb = has_rgb; L.prec = prec;
  size_t off = 0;
  auto gemm_layer = [&](int idx, int k_pad, int out) {
    L.k_pad[idx] = k_pad;
    L.w_hi[idx] = off; off = align256(off + (size_t)out * k_pad * 2);
    L.w_lo[idx] = off; if (prec) off = align256(off + (size_t)out * k_pad * 2);
    L.bias[idx] = off; off = align256(off + (size_t)out * 4);
  };
  for (int l = 0; l < depth; ++l) gemm_layer(l, l == 0 ? ENC_LD : (l == 5 ? width + ENC_LD : width), width);
  L.w_hi[depth] = off; off = align256(off + (size_t)(width + 1) * 4);           // density head
  if (has_rgb) {
    gemm_layer(depth + 1, width, BOTTLENECK);
    gemm_layer(depth + 2, BOTTLENECK + DIR_LD, VIEW_W);
    L.w_hi[depth + 3] = off; off = align256(off + (size_t)(3 * VIEW_W + 3) * 4);
  }
  L.total = off;
  return true;
}
static int dense_in(const Layout& L, int idx) {
  if (idx == 0) return 2 * NPAIR;
  if (idx < L.depth) return idx == 5 ? L.width + 2 * NPAIR : L.width;
  if (idx == L.depth) return L.width;
  if (idx == L.depth + 1) return L.width;
  if (idx == L.depth + 2) return BOTTLENECK + 3 + 6 * DIR_DEG;
  return VIEW_W;
}

struct Workspace { size_t enc, enc_lo, h[2], h_lo[2], bott, bott_lo, dir, dir_lo, hv, hv_lo, hpart, total; };
static Workspace make_ws(long long M, const Layout& L) {
  Workspace W{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align256(off + bytes); return o; };
  W.enc = take((size_t)M * ENC_LD * 2);
  W.enc_lo = L.prec ? take((size_t)M * ENC_LD * 2) : 0;
  for (int i = 0; i < 2; ++i) { W.h[i] = take((size_t)M * L.width * 2); W.h_lo[i] = L.prec ? take((size_t)M * L.width * 2) : 0; }
  if (L.has_rgb) {
    W.bott = take((size_t)M * BOTTLENECK * 2); W.bott_lo = L.prec ? take((size_t)M * BOTTLENECK * 2) : 0;
    W.dir = take((size_t)M * DIR_LD * 2); W.dir_lo = L.prec ? take((size_t)M * DIR_LD * 2) : 0;
    W.hv = take((size_t)M * VIEW_W * 2); W.hv_lo = L.prec ? take((size_t)M * VIEW_W * 2) : 0;
  }
  W.hpart = take((size_t)M * 2 * (L.width / 128 > 8 ? L.width / 128 : 8) * 4);
  W.total = off;
  return W;
}

static int g_fused3 = 1;      // A/B: mip360_debug_set_fused3
// one Dense layer: up to two A sources (a0 with k0 columns, a1 with k1 columns, lo images alongside in prec mode)
static int dense(const Layout& L, const uint8_t* packed, int idx, const uint8_t* a0, const uint8_t* a0_lo, int k0, int ld0,
                 const uint8_t* a1, const uint8_t* a1_lo, int k1, int ld1, uint8_t* out, uint8_t* out_lo, long long M, int N, int relu,
                 cudaStream_t st, const float* head_w = nullptr, float* head_part = nullptr, int head_n = 1, int no_store = 0) {
  using namespace gemm;
  GemmArgs g{};
  int bn, ctas;
  plan(N, (k0 + BK - 1) / BK * BK + (a1 ? (k1 + BK - 1) / BK * BK : 0), &bn, &ctas);
  const bool prec = L.prec != 0;
  if (make_map(&g.a[0], a0, (uint64_t)k0, (uint64_t)M, (uint64_t)ld0, BM)) return -1;
  g.a[1] = g.a[0]; g.a[2] = g.a[0]; g.a[3] = g.a[0];
  if (prec && make_map(&g.a[1], a0_lo, (uint64_t)k0, (uint64_t)M, (uint64_t)ld0, BM)) return -1;
  if (a1) {
    if (make_map(&g.a[2], a1, (uint64_t)k1, (uint64_t)M, (uint64_t)ld1, BM)) return -1;
    g.a[3] = g.a[2];
    if (prec && make_map(&g.a[3], a1_lo, (uint64_t)k1, (uint64_t)M, (uint64_t)ld1, BM)) return -1;
  }
  const int kp = L.k_pad[idx];
  if (make_map(&g.w[0], packed + L.w_hi[idx], (uint64_t)kp, (uint64_t)N, (uint64_t)kp, (uint32_t)(bn / ctas))) return -1;
  g.w[1] = g.w[0];
  if (prec && make_map(&g.w[1], packed + L.w_lo[idx], (uint64_t)kp, (uint64_t)N, (uint64_t)kp, (uint32_t)(bn / ctas))) return -1;
  if (make_map(&g.out[0], out, (uint64_t)N, (uint64_t)M, (uint64_t)N, OUT_BOX_ROWS)) return -1;
  g.out[1] = g.out[0];
  if (prec && make_map(&g.out[1], out_lo, (uint64_t)N, (uint64_t)M, (uint64_t)N, OUT_BOX_ROWS)) return -1;
  const int c0 = (k0 + BK - 1) / BK, c1 = a1 ? (k1 + BK - 1) / BK : 0;
  const int w1 = c0 * BK;     // weight column where the second source starts
  int ns = 0;
  // (A half, W half): hi x hi, then the two cross terms of the split-precision mode (lo x lo is below fp32 resolution)
  // CTA pairs in split precision: ONE pass whose stages carry hi and lo tiles of both operands (gemm_tc.cuh: F3)
  const bool fused3 = prec && ctas == 2 && g_fused3;
  g.fused3 = fused3;
  const int passes = (prec && !fused3) ? 3 : 1;
  for (int p = 0; p < passes; ++p) {
    const int ah = (p == 1) ? 1 : 0, wh = (p == 2) ? 1 : 0;
    g.seg[ns++] = Segment{0 + ah, 0, wh, 0, c0};
    if (a1) g.seg[ns++] = Segment{2 + ah, 0, wh, w1, c1};
  }
  g.n_seg = ns; g.M = (int)M; g.N = N; g.relu = relu;
  g.bias = reinterpret_cast<const float*>(packed + L.bias[idx]);
  g.head_w = head_w; g.head_part = head_part; g.head_n = head_n; g.no_store = no_store;
  return launch_gemm(g, bn, ctas, prec, st);
}

}  // namespace m360
}  // namespace npp

using namespace npp::m360;

// chain_tc.cu: the PropMLP's four layers + density head in one persistent kernel (activations stay on the SM)
int npp_prop_chain(const void* enc, const void* enc_lo, const void* const* w, const void* const* w_lo, const int* k_pad,
                   const float* const* bias, const float* head, float* density, long long M, cudaStream_t st);
static int g_chain = 1, g_fuse_head = 1;
extern "C" void mip360_debug_set_fused_head(int on) { g_fuse_head = on; }
extern "C" void mip360_debug_set_fused3(int on) { npp::m360::g_fused3 = on; }
extern "C" void mip360_debug_set_chain(int on) { g_chain = on; }      // A/B and tests: 0 = layer-by-layer GEMM launches

extern "C" int64_t mip360_mlp_packed_bytes(int net_depth, int net_width, int has_rgb, int prec) {
  Layout L;
  if (!make_layout(net_depth, net_width, has_rgb, prec, L)) { npp_set_error("mip360_mlp: unsupported network %d x %d", net_depth, net_width); return -1; }
  return (int64_t)L.total;
}

extern "C" int mip360_mlp_pack(const Mip360MlpParams* p, int net_depth, int net_width, int has_rgb, int prec, void* packed, void* stream) {
  NPP_CHECK_ARG(p && packed, "null pointer");
  Layout L;
  NPP_CHECK_ARG(make_layout(net_depth, net_width, has_rgb, prec, L), "unsupported network shape (4 or 8 layers of 256 or 1024)");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* out = (uint8_t*)packed;
  const int n_dense = net_depth + 1 + (has_rgb ? 3 : 0);
  for (int i = 0; i < n_dense; ++i) NPP_CHECK_ARG(p->kernel[i] && p->bias[i], "missing Dense parameters");
  auto gemm_layer = [&](int idx, int outf) {
    const int in = dense_in(L, idx), kp = L.k_pad[idx];
    const long long tot = (long long)outf * kp;
    pack_dense_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(p->kernel[idx], in, outf, kp, (__half*)(out + L.w_hi[idx]),
                                                                      prec ? (__half*)(out + L.w_lo[idx]) : nullptr);
    copy_f32_kernel<<<(outf + 255) / 256, 256, 0, st>>>(p->bias[idx], outf, (float*)(out + L.bias[idx]));
  };
  for (int l = 0; l < net_depth; ++l) gemm_layer(l, net_width);
  pack_head_kernel<<<(net_width + 1 + 255) / 256, 256, 0, st>>>(p->kernel[net_depth], p->bias[net_depth], net_width, 1, (float*)(out + L.w_hi[net_depth]));
  if (has_rgb) {
    gemm_layer(net_depth + 1, BOTTLENECK);
    gemm_layer(net_depth + 2, VIEW_W);
    pack_head_kernel<<<(3 * VIEW_W + 3 + 255) / 256, 256, 0, st>>>(p->kernel[net_depth + 3], p->bias[net_depth + 3], VIEW_W, 3,
                                                                   (float*)(out + L.w_hi[net_depth + 3]));
  }
  NPP_CHECK_LAUNCH();
  return 0;
}

extern "C" int64_t mip360_field_workspace_bytes(int64_t n_samples, int net_depth, int net_width, int has_rgb, int prec) {
  Layout L;
  if (!make_layout(net_depth, net_width, has_rgb, prec, L)) { npp_set_error("mip360_field: unsupported network %d x %d", net_depth, net_width); return -1; }
  return (int64_t)make_ws(n_samples, L).total;
}

extern "C" int mip360_cast_encode(const float* sdist, const float* near, const float* far, const float* origins, const float* directions,
                                  const float* viewdirs, const float* radii, int n_rays, int n_samples, float* out_tdist, void* out_enc,
                                  void* out_enc_lo, void* out_dir, void* out_dir_lo, float* out_means, float* out_covs, void* stream) {
  NPP_CHECK_ARG(sdist && near && far && origins && directions && radii, "null pointer");
  NPP_CHECK_ARG(n_rays > 0 && n_samples > 0, "empty batch");
  NPP_CHECK_ARG(!out_dir || viewdirs, "viewdirs needed for the direction encoding");
  const long long M = (long long)n_rays * n_samples;
  // the low halves are wanted: the split-precision mode, libm sin / exp; otherwise the SFU forms
  auto kern = out_enc_lo ? cast_encode_kernel<true> : cast_encode_kernel<false>;
  kern<<<(unsigned)((M + SPB - 1) / SPB), 256, 0, (cudaStream_t)stream>>>(
      sdist, near, far, origins, directions, viewdirs, radii, n_rays, n_samples, out_tdist, (__half*)out_enc, (__half*)out_enc_lo,
      (__half*)out_dir, (__half*)out_dir_lo, out_means, out_covs);
  NPP_CHECK_LAUNCH();
  return 0;
}

extern "C" int mip360_field_forward(const void* packed, int net_depth, int net_width, int has_rgb, int prec, const float* sdist,
                                    const float* near, const float* far, const float* origins, const float* directions,
                                    const float* viewdirs, const float* radii, int n_rays, int n_samples, float* out_tdist,
                                    float* out_density, float* out_rgb, void* workspace, void* stream) {
  NPP_CHECK_ARG(packed && workspace && out_density && out_tdist, "null pointer");
  NPP_CHECK_ARG(!has_rgb || (out_rgb && viewdirs), "the NerfMLP needs viewdirs and an rgb output");
  Layout L;
  NPP_CHECK_ARG(make_layout(net_depth, net_width, has_rgb, prec, L), "unsupported network shape (4 or 8 layers of 256 or 1024)");
  const long long M = (long long)n_rays * n_samples;
  NPP_CHECK_ARG(M > 0 && M < (1ll << 31), "n_rays * n_samples out of range");
  const Workspace W = make_ws(M, L);
  uint8_t* ws = (uint8_t*)workspace;
  const uint8_t* pk = (const uint8_t*)packed;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* enc = ws + W.enc;
  uint8_t* enc_lo = prec ? ws + W.enc_lo : nullptr;
  int rc = mip360_cast_encode(sdist, near, far, origins, directions, viewdirs, radii, n_rays, n_samples, out_tdist, enc, enc_lo,
                              has_rgb ? ws + W.dir : nullptr, (has_rgb && prec) ? ws + W.dir_lo : nullptr, nullptr, nullptr, stream);
  if (rc) return rc;
  if (g_chain && net_depth == 4 && net_width == 256 && !has_rgb) {
    const void* wp[4];
    const void* wl[4];
    const float* bp[4];
    int kp[4];
    for (int l = 0; l < 4; ++l) { wp[l] = pk + L.w_hi[l]; wl[l] = pk + L.w_lo[l]; bp[l] = (const float*)(pk + L.bias[l]); kp[l] = L.k_pad[l]; }
    return npp_prop_chain(enc, enc_lo, wp, prec ? wl : nullptr, kp, bp, (const float*)(pk + L.w_hi[net_depth]), out_density, M, st);
  }
  auto hbuf = [&](int i) { return ws + W.h[i]; };
  auto hlo = [&](int i) { return prec ? ws + W.h_lo[i] : (uint8_t*)nullptr; };
  const int Wd = net_width;
  const float* head_w = (const float*)(pk + L.w_hi[net_depth]);
  rc = dense(L, pk, 0, enc, enc_lo, 2 * NPAIR, ENC_LD, nullptr, nullptr, 0, 0, hbuf(0), hlo(0), M, Wd, 1, st);
  if (rc) return rc;
  int cur = 0;
  for (int l = 1; l < net_depth; ++l) {
    if (l == 5) rc = dense(L, pk, l, hbuf(cur), hlo(cur), Wd, Wd, enc, enc_lo, 2 * NPAIR, ENC_LD, hbuf(cur ^ 1), hlo(cur ^ 1), M, Wd, 1, st);
    else {
      // the last trunk layer's epilogue also evaluates the density head on its fp32 values (partials per column block)
      const bool last = g_fuse_head && l == net_depth - 1;
      rc = dense(L, pk, l, hbuf(cur), hlo(cur), Wd, Wd, nullptr, nullptr, 0, 0, hbuf(cur ^ 1), hlo(cur ^ 1), M, Wd, 1, st,
                 last ? head_w : nullptr, last ? (float*)(ws + W.hpart) : nullptr);
    }
    if (rc) return rc;
    cur ^= 1;
  }
  if (g_fuse_head && net_depth - 1 != 5) {
    int bn, ctas;
    npp::gemm::plan(Wd, Wd, &bn, &ctas);
    density_from_partials_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>((const float*)(ws + W.hpart), M, 2 * (Wd / bn), head_w + Wd, out_density);
  } else {
    const int blocks = (int)((M + 7) / 8 < 148 * 8 ? (M + 7) / 8 : 148 * 8);
    density_head_kernel<<<blocks, 256, (Wd + 1) * sizeof(float), st>>>((const __half*)hbuf(cur), (const __half*)hlo(cur), M, Wd, head_w, out_density);
  }
  if (has_rgb) {
    uint8_t* bott = ws + W.bott;
    uint8_t* bott_lo = prec ? ws + W.bott_lo : nullptr;
    rc = dense(L, pk, net_depth + 1, hbuf(cur), hlo(cur), Wd, Wd, nullptr, nullptr, 0, 0, bott, bott_lo, M, BOTTLENECK, 0, st);
    if (rc) return rc;
    // the view layer; its epilogue also evaluates the rgb head on the fp32 activated values (partials per column chunk), so the
    // 128-wide hidden layer is never written out
    const float* rgb_w = (const float*)(pk + L.w_hi[net_depth + 3]);
    float* rpart = (float*)(ws + W.hpart);
    rc = dense(L, pk, net_depth + 2, bott, bott_lo, BOTTLENECK, BOTTLENECK, ws + W.dir, prec ? ws + W.dir_lo : nullptr, DIR_LD, DIR_LD,
               ws + W.hv, prec ? ws + W.hv_lo : nullptr, M, VIEW_W, 1, st, g_fuse_head ? rgb_w : nullptr, g_fuse_head ? rpart : nullptr, 3, g_fuse_head);
    if (rc) return rc;
    if (g_fuse_head)
      rgb_from_partials_kernel<<<(unsigned)((M * 3 + 255) / 256), 256, 0, st>>>(rpart, M, 2, rgb_w + 3 * VIEW_W, out_rgb);
    else
      rgb_head_kernel<<<(unsigned)((M * 8 + 255) / 256), 256, 0, st>>>((const __half*)(ws + W.hv), prec ? (const __half*)(ws + W.hv_lo) : nullptr, M, rgb_w, out_rgb);
  }
  NPP_CHECK_LAUNCH();
  return 0;
}

extern "C" int mip360_resample_logits(const float* sdist, const float* weights, int n_rays, int n_bins, float anneal, float resample_padding,
                                      float* out_logits, void* stream) {
  NPP_CHECK_ARG(sdist && weights && out_logits, "null pointer");
  const long long tot = (long long)n_rays * n_bins;
  if (tot <= 0) return 0;
  resample_logits_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sdist, weights, n_rays, n_bins, anneal, resample_padding, out_logits);
  NPP_CHECK_LAUNCH();
  return 0;
}
