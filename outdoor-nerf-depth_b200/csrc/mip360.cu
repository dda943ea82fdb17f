// mipnerf360 twins of the hot path (config 3): interval resampling, alpha
// weights, volumetric rendering and the depth-prior losses of nerf-methods/mipnerf360/internal/.
// One warp per ray; the step function lives in shared memory, sums/scans are warp shuffles with a running carry,
// the inverse-CDF lookup is a binary search (equivalent to the reference's O(S*M) masked max/min, math.py:108-127).
#include <cfloat>
#include "common.cuh"

namespace npp {
namespace mip {

constexpr int WARPS = 4;
constexpr int MAX_BINS = 256;      // bins per ray the shared-memory staging is sized for

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// inclusive scan of one value per lane
__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
// jnp.nan_to_num(x, <positional copy>): NaN -> 0, +-inf -> +-FLT_MAX (math.py:125, render.py:196-200)
__device__ __forceinline__ float nan_to_num(float x) {
  if (isnan(x)) return 0.f;
  if (isinf(x)) return x > 0 ? FLT_MAX : -FLT_MAX;
  return x;
}
// sorted_interp (math.py:108-127) of one query into (xp, fp)[0..m): lo = last j with xp[j] <= x (0 if none),
// hi = first j with xp[j] > x (m-1 if none)
__device__ __forceinline__ float sorted_interp1(float x, const float* xp, const float* fp, int m) {
  int a = 0, b = m;                 // count of xp[j] <= x
  while (a < b) { int mid = (a + b) >> 1; if (xp[mid] <= x) a = mid + 1; else b = mid; }
  const int lo = a > 0 ? a - 1 : 0, hi = a < m ? a : m - 1;
  const float xp0 = xp[lo], xp1 = xp[hi], fp0 = fp[lo], fp1 = fp[hi];
  float off = nan_to_num(__fdiv_rn(__fsub_rn(x, xp0), __fsub_rn(xp1, xp0)));
  off = fminf(fmaxf(off, 0.f), 1.f);
  return __fadd_rn(fp0, __fmul_rn(off, __fsub_rn(fp1, fp0)));
}

// stepfun.sample_intervals (stepfun.py:214-263): t [B,M+1], w_logits [B,M], u [B,Ns] (u_ld 0 = one shared row)
// FUSED: the whole resampling step of Model.__call__ between two levels (models.py:171-200) in one launch -- the logits
// where(t[1:] > t[:-1], anneal * log(w + padding), -inf) are built here from the WEIGHTS (`logits` then points at them), rows
// of t / w are read with strides t_ld / w_ld (the dilated histogram's [1:-1] slices need no copy), and with `jitter` the
// ordinates are u[k] = base[k] + jitter[r] * max_jitter (stepfun.py:203-209, single_jitter) instead of a materialised [B,Ns].
template <bool FUSED>
__global__ void __launch_bounds__(WARPS * 32)
sample_intervals_kernel(const float* __restrict__ t, int t_ld, const float* __restrict__ logits, int w_ld, const float* __restrict__ u, int u_ld,
                        const float* __restrict__ jitter, float max_jitter, float anneal, float padding,
                        int B, int M, int Ns, float dmin, float dmax, float* __restrict__ out) {
  __shared__ float s_t[WARPS][MAX_BINS + 1], s_cw[WARPS][MAX_BINS + 1], s_c[WARPS][MAX_BINS + 1];
  __shared__ float s_lg[FUSED ? WARPS : 1][FUSED ? MAX_BINS : 1];
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int r = blockIdx.x * WARPS + wl;
  if (r >= B) return;
  const float* tr = t + (size_t)r * t_ld;
  const float* lr = logits + (size_t)r * w_ld;
  float* st = s_t[wl]; float* cw = s_cw[wl]; float* cen = s_c[wl];
  if (FUSED) {
    float* lg = s_lg[wl];
    for (int i = lane; i < M; i += 32)
      lg[i] = (tr[i + 1] > tr[i]) ? __fmul_rn(anneal, logf(__fadd_rn(lr[i], padding))) : -INFINITY;
    __syncwarp();
    lr = lg;
  }
  // softmax (jax.nn.softmax, stepfun.py:156)
  float mx = -INFINITY;
  for (int i = lane; i < M; i += 32) mx = fmaxf(mx, lr[i]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int i = lane; i < M; i += 32) sum += expf(lr[i] - mx);
  sum = warp_sum(sum);
  // integrate_weights (stepfun.py:131-150): cw0 = [0, min(1, cumsum(w[:-1])), 1]
  float carry = 0.f;
  for (int base = 0; base < M; base += 32) {
    const int i = base + lane;
    const float w = i < M ? expf(lr[i] - mx) / sum : 0.f;
    const float inc = warp_incl_scan(w, lane) + carry;
    if (i < M - 1) cw[i + 1] = fminf(1.f, inc);
    carry = __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) { cw[0] = 0.f; cw[M] = 1.f; }
  for (int i = lane; i <= M; i += 32) st[i] = tr[i];
  __syncwarp();
  // invert_cdf -> centres (stepfun.py:153-161)
  const float* ur = u + (size_t)r * u_ld;
  const float jit = (FUSED && jitter) ? __fmul_rn(jitter[r], max_jitter) : 0.f;
  for (int k = lane; k < Ns; k += 32) cen[k] = sorted_interp1((FUSED && jitter) ? __fadd_rn(ur[k], jit) : ur[k], cw, st, M + 1);
  __syncwarp();
  // fenceposts: midpoints, reflected and clamped ends (stepfun.py:250-262)
  float* o = out + (size_t)r * (Ns + 1);
  for (int k = lane; k <= Ns; k += 32) {
    float v;
    if (k == 0) v = fmaxf(dmin, __fsub_rn(__fmul_rn(2.f, cen[0]), __fdiv_rn(__fadd_rn(cen[1], cen[0]), 2.f)));
    else if (k == Ns) v = fminf(dmax, __fsub_rn(__fmul_rn(2.f, cen[Ns - 1]), __fdiv_rn(__fadd_rn(cen[Ns - 1], cen[Ns - 2]), 2.f)));
    else v = __fdiv_rn(__fadd_rn(cen[k], cen[k - 1]), 2.f);
    o[k] = v;
  }
}

// render.compute_alpha_weights (render.py:130-151)
__global__ void __launch_bounds__(WARPS * 32)
alpha_weights_kernel(const float* __restrict__ density, const float* __restrict__ tdist, const float* __restrict__ dirs, int B, int S,
                     int opaque, float* __restrict__ weights, float* __restrict__ alpha, float* __restrict__ trans) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (r >= B) return;
  const float dn = sqrtf(dirs[3 * r] * dirs[3 * r] + dirs[3 * r + 1] * dirs[3 * r + 1] + dirs[3 * r + 2] * dirs[3 * r + 2]);
  const float* td = tdist + (size_t)r * (S + 1);
  float carry = 0.f;
  for (int base = 0; base < S; base += 32) {
    const int i = base + lane;
    float dd = 0.f;
    if (i < S) {
      dd = density[(size_t)r * S + i] * ((td[i + 1] - td[i]) * dn);
      if (opaque && i == S - 1) dd = INFINITY;
    }
    const float incl = warp_incl_scan(i < S - 1 ? dd : 0.f, lane);      // cumsum(density_delta[:-1])
    const float excl = carry + incl - (i < S - 1 ? dd : 0.f);
    if (i < S) {
      const float a = 1.f - expf(-dd), tr = expf(-excl);
      const size_t o = (size_t)r * S + i;
      weights[o] = a * tr;
      if (alpha) alpha[o] = a;
      if (trans) trans[o] = tr;
    }
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
}

// render.volumetric_rendering (render.py:154-216), compute_extras=True, extras=None.
// out_scalars [B,6] = acc, distance_mean, depth, percentile_5, median, percentile_95
__global__ void __launch_bounds__(WARPS * 32)
volumetric_rendering_kernel(const float* __restrict__ rgbs, const float* __restrict__ weights, const float* __restrict__ tdist,
                            const float* __restrict__ bg_rgbs, int bg_ld, const float* __restrict__ t_far, int B, int S,
                            float* __restrict__ out_rgb, float* __restrict__ out_scalars) {
  __shared__ float s_cw[WARPS][MAX_BINS + 2], s_t[WARPS][MAX_BINS + 2];
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int r = blockIdx.x * WARPS + wl;
  if (r >= B) return;
  const float* w = weights + (size_t)r * S;
  const float* td = tdist + (size_t)r * (S + 1);
  const float* c = rgbs + (size_t)r * S * 3;
  float acc = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, elog = 0.f, emid = 0.f;
  for (int i = lane; i < S; i += 32) {
    const float wi = w[i], tm = 0.5f * (td[i] + td[i + 1]);
    acc += wi; cr += wi * c[3 * i]; cg += wi * c[3 * i + 1]; cb += wi * c[3 * i + 2];
    elog += wi * logf(tm); emid += wi * tm;
  }
  acc = warp_sum(acc); cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb); elog = warp_sum(elog); emid = warp_sum(emid);
  const float bgw = fmaxf(0.f, 1.f - acc);
  // percentiles: weighted_percentile(t_aug, w_aug, [5,50,95]) (stepfun.py:298-308) = jnp.interp into integrate_weights
  float* cw = s_cw[wl]; float* ta = s_t[wl];
  float carry = 0.f;
  for (int base = 0; base < S + 1; base += 32) {       // w_aug has S+1 entries; cumsum over the first S
    const int i = base + lane;
    const float wi = i < S ? w[i] : 0.f;
    const float inc = warp_incl_scan(wi, lane) + carry;
    if (i < S) cw[i + 1] = fminf(1.f, inc);
    carry = __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) { cw[0] = 0.f; cw[S + 1] = 1.f; ta[S + 1] = t_far[r]; }
  for (int i = lane; i <= S; i += 32) ta[i] = td[i];
  __syncwarp();
  if (lane < 3) {
    const float p = lane == 0 ? 0.05f : lane == 1 ? 0.5f : 0.95f;
    const int m = S + 2;
    int a = 0, b = m;                                    // np.interp: j = (#xp <= x) - 1, slope form
    while (a < b) { int mid = (a + b) >> 1; if (cw[mid] <= p) a = mid + 1; else b = mid; }
    float v;
    if (a == 0) v = ta[0];
    else if (a >= m) v = ta[m - 1];
    else { const int j = a - 1; v = ta[j] + (p - cw[j]) * ((ta[j + 1] - ta[j]) / (cw[j + 1] - cw[j])); }
    out_scalars[6 * r + 3 + lane] = v;
  }
  if (lane == 0) {
    const float* bg = bg_rgbs + (size_t)r * bg_ld;
    out_rgb[3 * r] = cr + bgw * bg[0]; out_rgb[3 * r + 1] = cg + bgw * bg[1]; out_rgb[3 * r + 2] = cb + bgw * bg[2];
    const float t0 = td[0], t1 = td[S];
    out_scalars[6 * r] = acc;
    out_scalars[6 * r + 1] = fminf(fmaxf(nan_to_num(expf(elog / fmaxf(FLT_EPSILON, acc))), t0), t1);
    out_scalars[6 * r + 2] = fminf(fmaxf(nan_to_num(emid), t0), t1);
  }
}

// depth losses: kl = ds_nerf_depth_loss (depth_loss.py:5-26 via :66-97) as the trainer reaches it (mean over
// rays x samples, invalid rays zeroed but counted); mse / l1 = train_utils.py:109-121 on distance_mean.
// partial [gridDim.x] fp64 sums, reduced by the host-side second launch.
__global__ void __launch_bounds__(WARPS * 32)
depth_loss_partial_kernel(const float* __restrict__ weights, const float* __restrict__ tdist, const float* __restrict__ prior,
                          const float* __restrict__ pred, const float* __restrict__ dirs, int B, int S, int type, float sigma,
                          double* __restrict__ partial) {
  __shared__ double s_part[WARPS];
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int r = blockIdx.x * WARPS + wl;
  double acc = 0.0;
  if (r < B) {
    const float td = prior[r];
    const float m = td > 0.f ? 1.f : 0.f;
    if (type == NERFPP_DEPTH_KL) {
      const float dn = sqrtf(dirs[3 * r] * dirs[3 * r] + dirs[3 * r + 1] * dirs[3 * r + 1] + dirs[3 * r + 2] * dirs[3 * r + 2]);
      const float* t = tdist + (size_t)r * (S + 1);
      float a = 0.f;
      for (int i = lane; i < S; i += 32) {
        const float step = 0.5f * (t[i] + t[i + 1]), len = (t[i + 1] - t[i]) * dn, d = step - td;
        a += -logf(weights[(size_t)r * S + i] + 1e-7f) * expf(-(d * d) / (2.f * sigma)) * len * m;
      }
      acc = (double)a;
    } else if (lane == 0) {
      const float d = m * pred[r] - m * td;
      acc = type == NERFPP_DEPTH_MSE ? (double)(d * d) : (double)fabsf(d);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) s_part[wl] = acc;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0; for (int i = 0; i < WARPS; ++i) s += s_part[i]; partial[blockIdx.x] = s; }
}
__global__ void depth_loss_final_kernel(const double* __restrict__ partial, int nblocks, double denom, float* __restrict__ out) {
  __shared__ double s[256];
  double a = 0;
  for (int i = threadIdx.x; i < nblocks; i += 256) a += partial[i];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) out[0] = (float)(s[0] / denom);
}


// ---- N4: the regularisers of the mipnerf360 trainer (train_utils.py:160-180) -----------------
// stepfun.lossfun_outer (stepfun.py:82-89) on top of inner_outer (:64-79) and searchsorted (:30-53), one warp per ray:
//   cy = [0, cumsum(w_env)];  for every fencepost v = t[i]:  c = #{j : t_env[j] <= v},  lo = max(c-1, 0),  hi = min(c, P)
//   w_outer[i] = cy[hi[i+1]] - cy[lo[i]] = sum of w_env[k] for k in [lo[i], hi[i+1]);   loss[i] = max(0, w[i]-w_outer[i])^2 / (w[i]+eps)
// Backward (g = d/d loss, may be NULL for the forward alone): the only differentiable input of interlevel_loss is w_env
// (c and w come through stop_gradient, the indices are piecewise constant):
//   d w_env[k] = sum_i [lo[i] <= k < hi[i+1]] * g[i] * (-2 max(0, w[i]-w_outer[i]) / (w[i]+eps))
__global__ void __launch_bounds__(WARPS * 32)
lossfun_outer_kernel(const float* __restrict__ t, const float* __restrict__ w, const float* __restrict__ t_env,
                     const float* __restrict__ w_env, int B, int S, int P, float eps, const float* __restrict__ g,
                     float* __restrict__ out_loss, float* __restrict__ out_dwenv) {
  __shared__ float s_te[WARPS][MAX_BINS + 1], s_cy[WARPS][MAX_BINS + 1], s_go[WARPS][MAX_BINS];
  __shared__ short s_lo[WARPS][MAX_BINS + 1], s_hi[WARPS][MAX_BINS + 1];
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int r = blockIdx.x * WARPS + wl;
  if (r >= B) return;
  float* te = s_te[wl]; float* cy = s_cy[wl]; float* go = s_go[wl];
  short* lo = s_lo[wl]; short* hi = s_hi[wl];
  for (int j = lane; j <= P; j += 32) te[j] = t_env[(size_t)r * (P + 1) + j];
  // cy[0] = 0, cy[k+1] = w_env[0] + ... + w_env[k]   (fp32 inclusive scan, 32 at a time with a running carry)
  float carry = 0.f;
  if (lane == 0) cy[0] = 0.f;
  for (int base = 0; base < P; base += 32) {
    const int k = base + lane;
    float v = k < P ? w_env[(size_t)r * P + k] : 0.f;
    v = warp_incl_scan(v, lane) + carry;
    if (k < P) cy[k + 1] = v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
  __syncwarp();
  for (int i = lane; i <= S; i += 32) {
    const float v = t[(size_t)r * (S + 1) + i];
    int a = 0, b = P + 1;                       // c = number of t_env[j] <= v
    while (a < b) { const int mid = (a + b) >> 1; if (te[mid] <= v) a = mid + 1; else b = mid; }
    lo[i] = (short)(a > 0 ? a - 1 : 0);
    hi[i] = (short)(a < P + 1 ? a : P);
  }
  __syncwarp();
  for (int i = lane; i < S; i += 32) {
    const float wi = w[(size_t)r * S + i];
    const float w_outer = cy[hi[i + 1]] - cy[lo[i]];
    const float ex = fmaxf(0.f, wi - w_outer);
    if (out_loss) out_loss[(size_t)r * S + i] = ex * ex / (wi + eps);
    go[i] = g ? g[(size_t)r * S + i] * (-2.f * ex / (wi + eps)) : 0.f;
  }
  __syncwarp();
  if (g && out_dwenv) {
    for (int k = lane; k < P; k += 32) {
      float acc = 0.f;
      for (int i = 0; i < S; ++i) acc += (lo[i] <= k && k < hi[i + 1]) ? go[i] : 0.f;
      out_dwenv[(size_t)r * P + k] = acc;
    }
  }
}

// stepfun.lossfun_distortion (stepfun.py:266-276), one warp per ray, O(S^2):
//   u_i = (t_i + t_{i+1}) / 2,  A_i = sum_j w_j |u_i - u_j|,  loss = sum_i w_i A_i + sum_i w_i^2 (t_{i+1} - t_i) / 3
// Backward (g = d/d loss [B], NULL for the forward alone), with B_i = sum_j w_j sign(u_i - u_j):
//   d w_i = g (2 A_i + (2/3) w_i (t_{i+1} - t_i)),   d u_i = 2 g w_i B_i,
//   d t_k = (d u_{k-1} + d u_k) / 2 + g (w_{k-1}^2 - w_k^2) / 3      (terms with an index outside [0, S) dropped)
__global__ void __launch_bounds__(WARPS * 32)
lossfun_distortion_kernel(const float* __restrict__ t, const float* __restrict__ w, int B, int S, const float* __restrict__ g,
                          float* __restrict__ out_loss, float* __restrict__ out_dt, float* __restrict__ out_dw) {
  __shared__ float s_t[WARPS][MAX_BINS + 1], s_u[WARPS][MAX_BINS], s_w[WARPS][MAX_BINS], s_du[WARPS][MAX_BINS];
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int r = blockIdx.x * WARPS + wl;
  if (r >= B) return;
  float* st = s_t[wl]; float* su = s_u[wl]; float* sw = s_w[wl]; float* du = s_du[wl];
  for (int i = lane; i <= S; i += 32) st[i] = t[(size_t)r * (S + 1) + i];
  for (int i = lane; i < S; i += 32) sw[i] = w[(size_t)r * S + i];
  __syncwarp();
  for (int i = lane; i < S; i += 32) su[i] = (st[i + 1] + st[i]) / 2.f;
  __syncwarp();
  const float gr = g ? g[r] : 0.f;
  float part = 0.f;
  for (int i = lane; i < S; i += 32) {
    const float ui = su[i], wi = sw[i];
    float A = 0.f, Bs = 0.f;
    for (int j = 0; j < S; ++j) {
      const float d = ui - su[j];
      A += sw[j] * fabsf(d);
      Bs += sw[j] * (d > 0.f ? 1.f : d < 0.f ? -1.f : 0.f);
    }
    const float delta = st[i + 1] - st[i];
    part += wi * A + wi * wi * delta / 3.f;
    if (g) {
      if (out_dw) out_dw[(size_t)r * S + i] = gr * (2.f * A + (2.f / 3.f) * wi * delta);
      du[i] = 2.f * gr * wi * Bs;
    }
  }
  part = warp_sum(part);
  if (lane == 0 && out_loss) out_loss[r] = part;
  __syncwarp();
  if (g && out_dt) {
    for (int k = lane; k <= S; k += 32) {
      float v = 0.f;
      if (k > 0) v += 0.5f * du[k - 1] + gr * sw[k - 1] * sw[k - 1] / 3.f;
      if (k < S) v += 0.5f * du[k] - gr * sw[k] * sw[k] / 3.f;
      out_dt[(size_t)r * (S + 1) + k] = v;
    }
  }
}


// stepfun.max_dilate / max_dilate_weights (stepfun.py:99-128; models.py:150-169 dilates the proposal histogram between
// levels), one warp per ray.  The 3M+1 fenceposts  sort(cat(t, t[:-1]-d, t[1:]+d))  come from a warp bitonic sort in
// shared memory (values only, padded with +inf), are clipped to the domain, and every dilated interval takes the max of
// the values whose widened interval [t_j - d, t_j+1 + d) contains its left fencepost.  weights_mode: the values are the
// pdf w / max(eps, dt) going in and are multiplied by the dilated interval widths coming out (weight_to_pdf /
// pdf_to_weight :89-96), optionally renormalised to sum 1.
constexpr int DIL_MAX_BINS = 85;              // 3 M + 1 <= 256
__global__ void __launch_bounds__(WARPS * 32)
max_dilate_kernel(const float* __restrict__ t, const float* __restrict__ w, int B, int M, float dilation, float dmin, float dmax,
                  int weights_mode, int renormalize, float eps, float* __restrict__ out_t, float* __restrict__ out_w) {
  __shared__ float s_all[WARPS][256], s_t0[WARPS][DIL_MAX_BINS], s_t1[WARPS][DIL_MAX_BINS], s_v[WARPS][DIL_MAX_BINS], s_wd[WARPS][256];
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int r = blockIdx.x * WARPS + wl;
  if (r >= B) return;
  float* all = s_all[wl]; float* t0 = s_t0[wl]; float* t1 = s_t1[wl]; float* v = s_v[wl]; float* wd = s_wd[wl];
  const int n_out = 3 * M + 1;
  int P = 2;
  while (P < n_out) P <<= 1;
  const float* tr = t + (size_t)r * (M + 1);
  for (int j = lane; j <= M; j += 32) all[j] = tr[j];
  for (int j = n_out + lane; j < P; j += 32) all[j] = INFINITY;
  __syncwarp();
  for (int j = lane; j < M; j += 32) {
    const float a = all[j], b = all[j + 1];
    const float lo = __fsub_rn(a, dilation), hi = __fadd_rn(b, dilation);
    t0[j] = lo; t1[j] = hi;
    all[M + 1 + j] = lo; all[2 * M + 1 + j] = hi;
    const float wj = w[(size_t)r * M + j];
    v[j] = weights_mode ? __fdiv_rn(wj, fmaxf(eps, __fsub_rn(b, a))) : wj;
  }
  __syncwarp();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int q = lane; q < (P >> 1); q += 32) {
        const int i = ((q & ~(j - 1)) << 1) | (q & (j - 1));
        const int l = i | j;
        const float a = all[i], b = all[l];
        const bool up = (i & k) == 0;
        if ((a > b) == up && a != b) { all[i] = b; all[l] = a; }
      }
      __syncwarp();
    }
  }
  for (int k = lane; k < n_out; k += 32) all[k] = fminf(fmaxf(all[k], dmin), dmax);      // jnp.clip(t_dilate, *domain)
  __syncwarp();
  float part = 0.f;
  for (int k = lane; k < n_out - 1; k += 32) {
    const float tk = all[k];
    float m = 0.f;
    // t is sorted, hence so are t0 = t[:-1] - d and t1 = t[1:] + d: the intervals with t0[j] <= tk < t1[j] are the contiguous
    // range [first j with t1[j] > tk, last j with t0[j] <= tk] -- two binary searches instead of a sweep over all M
    int lo = 0, hi = M;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (t1[mid] > tk) hi = mid; else lo = mid + 1; }
    const int jl = lo;
    lo = 0; hi = M;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (t0[mid] <= tk) lo = mid + 1; else hi = mid; }
    for (int j = jl; j < lo; ++j) m = fmaxf(m, v[j]);
    if (weights_mode) m = __fmul_rn(m, __fsub_rn(all[k + 1], tk));
    wd[k] = m;
    part += m;
  }
  part = warp_sum(part);
  const float inv = (weights_mode && renormalize) ? fmaxf(eps, part) : 1.f;
  __syncwarp();
  for (int k = lane; k < n_out; k += 32) out_t[(size_t)r * n_out + k] = all[k];
  for (int k = lane; k < n_out - 1; k += 32) out_w[(size_t)r * (n_out - 1) + k] = (weights_mode && renormalize) ? __fdiv_rn(wd[k], inv) : wd[k];
}

}  // namespace mip
}  // namespace npp

using namespace npp;

extern "C" int mip360_sample_intervals(const float* t, const float* w_logits, const float* u, int u_ld, int n_rays, int n_bins,
                                       int n_samples, float domain_min, float domain_max, float* out_t, void* stream) {
  NPP_CHECK_ARG(t && w_logits && u && out_t, "null argument");
  NPP_CHECK_ARG(n_rays >= 0 && n_bins >= 1 && n_bins <= mip::MAX_BINS && n_samples >= 2 && n_samples <= mip::MAX_BINS, "bad shape");
  if (n_rays == 0) return 0;
  mip::sample_intervals_kernel<false><<<(n_rays + mip::WARPS - 1) / mip::WARPS, mip::WARPS * 32, 0, (cudaStream_t)stream>>>(
      t, n_bins + 1, w_logits, n_bins, u, u_ld, nullptr, 0.f, 1.f, 0.f, n_rays, n_bins, n_samples, domain_min, domain_max, out_t);
  NPP_CHECK_LAUNCH();
  return 0;
}

extern "C" int mip360_resample_level(const float* t, int t_ld, const float* weights, int w_ld, int n_rays, int n_bins, float anneal,
                                     float resample_padding, const float* u, int u_ld, const float* jitter, float max_jitter, int n_samples,
                                     float domain_min, float domain_max, float* out_t, void* stream) {
  NPP_CHECK_ARG(t && weights && u && out_t, "null argument");
  NPP_CHECK_ARG(n_rays >= 0 && n_bins >= 1 && n_bins <= mip::MAX_BINS && n_samples >= 2 && n_samples <= mip::MAX_BINS, "bad shape");
  NPP_CHECK_ARG(t_ld >= n_bins + 1 && w_ld >= n_bins, "row strides shorter than the rows");
  if (n_rays == 0) return 0;
  mip::sample_intervals_kernel<true><<<(n_rays + mip::WARPS - 1) / mip::WARPS, mip::WARPS * 32, 0, (cudaStream_t)stream>>>(
      t, t_ld, weights, w_ld, u, u_ld, jitter, max_jitter, anneal, resample_padding, n_rays, n_bins, n_samples, domain_min, domain_max, out_t);
  NPP_CHECK_LAUNCH();
  return 0;
}

extern "C" int mip360_compute_alpha_weights(const float* density, const float* tdist, const float* dirs, int n_rays, int n_samples,
                                            int opaque_background, float* out_weights, float* out_alpha, float* out_trans, void* stream) {
  NPP_CHECK_ARG(density && tdist && dirs && out_weights, "null argument");
  NPP_CHECK_ARG(n_rays >= 0 && n_samples >= 1, "bad shape");
  if (n_rays == 0) return 0;
  mip::alpha_weights_kernel<<<(n_rays + mip::WARPS - 1) / mip::WARPS, mip::WARPS * 32, 0, (cudaStream_t)stream>>>(
      density, tdist, dirs, n_rays, n_samples, opaque_background, out_weights, out_alpha, out_trans);
  NPP_CHECK_LAUNCH();
  return 0;
}

extern "C" int mip360_volumetric_rendering(const float* rgbs, const float* weights, const float* tdist, const float* bg_rgbs, int bg_ld,
                                           const float* t_far, int n_rays, int n_samples, float* out_rgb, float* out_scalars, void* stream) {
  NPP_CHECK_ARG(rgbs && weights && tdist && bg_rgbs && t_far && out_rgb && out_scalars, "null argument");
  NPP_CHECK_ARG(n_rays >= 0 && n_samples >= 1 && n_samples <= mip::MAX_BINS, "bad shape");
  if (n_rays == 0) return 0;
  mip::volumetric_rendering_kernel<<<(n_rays + mip::WARPS - 1) / mip::WARPS, mip::WARPS * 32, 0, (cudaStream_t)stream>>>(
      rgbs, weights, tdist, bg_rgbs, bg_ld, t_far, n_rays, n_samples, out_rgb, out_scalars);
  NPP_CHECK_LAUNCH();
  return 0;
}

extern "C" int64_t mip360_depth_loss_workspace_bytes(int n_rays) {
  return n_rays < 0 ? -1 : (int64_t)((n_rays + mip::WARPS - 1) / mip::WARPS + 1) * (int64_t)sizeof(double);
}

extern "C" int mip360_depth_loss(const float* weights, const float* tdist, const float* termination_depth, const float* predicted_depth,
                                 const float* dirs, int n_rays, int n_samples, int depth_loss_type, float sigma, float* out_loss,
                                 void* workspace, void* stream) {
  NPP_CHECK_ARG(termination_depth && out_loss && workspace, "null argument");
  NPP_CHECK_ARG(n_rays >= 1 && n_samples >= 1, "bad shape");
  NPP_CHECK_ARG(depth_loss_type == NERFPP_DEPTH_KL || depth_loss_type == NERFPP_DEPTH_MSE || depth_loss_type == NERFPP_DEPTH_L1, "unknown depth_loss_type");
  if (depth_loss_type == NERFPP_DEPTH_KL) NPP_CHECK_ARG(weights && tdist && dirs, "kl needs weights, tdist, dirs");
  else NPP_CHECK_ARG(predicted_depth, "mse/l1 need predicted_depth");
  const int nb = (n_rays + mip::WARPS - 1) / mip::WARPS;
  mip::depth_loss_partial_kernel<<<nb, mip::WARPS * 32, 0, (cudaStream_t)stream>>>(weights, tdist, termination_depth, predicted_depth, dirs,
                                                                                  n_rays, n_samples, depth_loss_type, sigma, (double*)workspace);
  NPP_CHECK_LAUNCH();
  const double denom = depth_loss_type == NERFPP_DEPTH_KL ? (double)n_rays * n_samples : (double)n_rays;
  mip::depth_loss_final_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((const double*)workspace, nb, denom, out_loss);
  NPP_CHECK_LAUNCH();
  return 0;
}

extern "C" int mip360_lossfun_outer(const float* t, const float* w, const float* t_env, const float* w_env, int n_rays, int n_bins,
                                    int n_env_bins, float eps, const float* grad_loss, float* out_loss, float* out_grad_w_env,
                                    void* stream) {
  NPP_CHECK_ARG(t && w && t_env && w_env && (out_loss || (grad_loss && out_grad_w_env)), "null argument");
  NPP_CHECK_ARG(!out_grad_w_env || grad_loss, "out_grad_w_env needs grad_loss");
  NPP_CHECK_ARG(n_rays >= 0 && n_bins >= 1 && n_bins <= mip::MAX_BINS && n_env_bins >= 1 && n_env_bins <= mip::MAX_BINS, "bad shape");
  if (n_rays == 0) return 0;
  mip::lossfun_outer_kernel<<<(n_rays + mip::WARPS - 1) / mip::WARPS, mip::WARPS * 32, 0, (cudaStream_t)stream>>>(
      t, w, t_env, w_env, n_rays, n_bins, n_env_bins, eps, grad_loss, out_loss, out_grad_w_env);
  NPP_CHECK_LAUNCH();
  return 0;
}

extern "C" int mip360_lossfun_distortion(const float* t, const float* w, int n_rays, int n_bins, const float* grad_loss,
                                         float* out_loss, float* out_grad_t, float* out_grad_w, void* stream) {
  NPP_CHECK_ARG(t && w && (out_loss || grad_loss), "null argument");
  NPP_CHECK_ARG((!out_grad_t && !out_grad_w) || grad_loss, "gradients need grad_loss");
  NPP_CHECK_ARG(n_rays >= 0 && n_bins >= 1 && n_bins <= mip::MAX_BINS, "bad shape");
  if (n_rays == 0) return 0;
  mip::lossfun_distortion_kernel<<<(n_rays + mip::WARPS - 1) / mip::WARPS, mip::WARPS * 32, 0, (cudaStream_t)stream>>>(
      t, w, n_rays, n_bins, grad_loss, out_loss, out_grad_t, out_grad_w);
  NPP_CHECK_LAUNCH();
  return 0;
}

extern "C" int mip360_max_dilate(const float* t, const float* w, int n_rays, int n_bins, float dilation, float domain_min,
                                 float domain_max, int weights_mode, int renormalize, float eps, float* out_t, float* out_w,
                                 void* stream) {
  NPP_CHECK_ARG(t && w && out_t && out_w, "null argument");
  NPP_CHECK_ARG(n_rays >= 0 && n_bins >= 1 && n_bins <= mip::DIL_MAX_BINS, "bad shape (at most 85 bins: 3 M + 1 <= 256 fenceposts)");
  if (n_rays == 0) return 0;
  mip::max_dilate_kernel<<<(n_rays + mip::WARPS - 1) / mip::WARPS, mip::WARPS * 32, 0, (cudaStream_t)stream>>>(
      t, w, n_rays, n_bins, dilation, domain_min, domain_max, weights_mode, renormalize, eps, out_t, out_w);
  NPP_CHECK_LAUNCH();
  return 0;
}
