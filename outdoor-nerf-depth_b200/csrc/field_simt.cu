// Field evaluation, plain-fp32 variant: positional encoding, the 8x256
// MLP with its skip connection, sigma head and colour head, for a tile of 64 samples per CTA.
// Every product and sum is an fp32 FFMA: this is the arithmetic cross-check for the tcgen05
// variant (field_tc.cu) and the backing of NERFPP_FIELD_SIMT; it is not the fast path.
// Activations live in shared memory ([sample][feature]); weights stream through a
// double-buffered cp.async stage of 16 input rows x 256 outputs, already transposed by the packer.
#include "common.cuh"

namespace npp {

constexpr int TS = 64;        // samples per CTA
constexpr int KC = 16;        // weight rows per stage
constexpr int SIMT_THREADS = 256;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// out[s][n] = act(bias[n] + sum_k in[s][k] * Wt[k][n]) for the CTA's 64 samples; the input is the
// concatenation of two shared-memory segments (K0 then K1 features, each a multiple of KC).
template <int NOUT>
__device__ void dense(const float* __restrict__ Wt, const float* __restrict__ bias, const float* in0, int ld0, int K0,
                      const float* in1, int ld1, int K1, float* out, int ldo, bool relu, float* wst) {
  constexpr int NV = NOUT / 128;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  float acc[8][4 * NV];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4 * NV; ++j) acc[i][j] = 0.f;
  const int nchunks = (K0 + K1) / KC;
  auto issue = [&](int c) {
    const float* src = Wt + (size_t)c * KC * NOUT;
    float* dst = wst + (c & 1) * KC * NOUT;
    for (int i = tid; i < KC * NOUT / 4; i += SIMT_THREADS) cp_async16(dst + 4 * i, src + 4 * i);
    cp_async_commit();
  };
  issue(0);
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) { issue(c + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const int k0 = c * KC;
    const float* in = (k0 < K0) ? in0 + k0 : in1 + (k0 - K0);
    const int ld = (k0 < K0) ? ld0 : ld1;
    const float* wb = wst + (c & 1) * KC * NOUT;
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
      float4 a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(in + (ty * 8 + i) * ld + kk);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 w[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) w[v] = *reinterpret_cast<const float4*>(wb + (kk + j) * NOUT + v * 128 + tx * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float av = j == 0 ? a[i].x : j == 1 ? a[i].y : j == 2 ? a[i].z : a[i].w;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            acc[i][4 * v + 0] = fmaf(av, w[v].x, acc[i][4 * v + 0]);
            acc[i][4 * v + 1] = fmaf(av, w[v].y, acc[i][4 * v + 1]);
            acc[i][4 * v + 2] = fmaf(av, w[v].z, acc[i][4 * v + 2]);
            acc[i][4 * v + 3] = fmaf(av, w[v].w, acc[i][4 * v + 3]);
          }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    float4 b = *reinterpret_cast<const float4*>(bias + v * 128 + tx * 4);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 o = make_float4(acc[i][4 * v] + b.x, acc[i][4 * v + 1] + b.y, acc[i][4 * v + 2] + b.z, acc[i][4 * v + 3] + b.w);
      if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      *reinterpret_cast<float4*>(out + (ty * 8 + i) * ldo + v * 128 + tx * 4) = o;
    }
  }
  __syncthreads();
}

// Embedder.forward (nerf_network.py:42-60): [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...]; this thread
// fills the frequencies k = q, q+4, ... of one sample.
template <int D>
__device__ __forceinline__ void embed_part(const float* x, int nfreq, int q, float* row) {
  if (q == 0)
    for (int c = 0; c < D; ++c) row[c] = x[c];
  for (int k = q; k < nfreq; k += 4) {
    float f = (float)(1 << k);
    for (int c = 0; c < D; ++c) {
      float s, co;
      sincosf(x[c] * f, &s, &co);
      row[D + 2 * k * D + c] = s;
      row[D + (2 * k + 1) * D + c] = co;
    }
  }
}

template <bool BG>
__global__ void __launch_bounds__(SIMT_THREADS, 1)
field_simt_kernel(const float* __restrict__ packed, const float* __restrict__ ray_o, const float* __restrict__ ray_d,
                  const float* __restrict__ z, int n, int S, float* __restrict__ out_sigma, float* __restrict__ out_rgb,
                  float* __restrict__ out_depth_real) {
  constexpr int D = BG ? 4 : 3;
  constexpr int EP = emb_pad_simt(BG);
  constexpr SimtLayout L = simt_layout(BG);
  extern __shared__ __align__(16) float smem[];
  float* E = smem;                  // [TS][EP]
  float* V = E + TS * EP;           // [TS][VIEW_PAD]
  float* H0 = V + TS * VIEW_PAD;    // [TS][W]
  float* H1 = H0 + TS * W;          // [TS][W]
  float* wst = H1 + TS * W;         // [2][KC][W]
  const int tid = threadIdx.x;
  const long long total = (long long)n * S;
  const long long g0 = (long long)blockIdx.x * TS;

  {  // ---- positions and encodings -------------------------------------------------------------
    int s = tid >> 2, q = tid & 3;
    long long g = g0 + s;
    float* erow = E + s * EP;
    float* vrow = V + s * VIEW_PAD;
    if (q == 0) {
      for (int c = emb_dim(BG); c < EP; ++c) erow[c] = 0.f;
      for (int c = VIEW_DIM; c < VIEW_PAD; ++c) vrow[c] = 0.f;
    }
    if (g < total) {
      int r = (int)(g / S), j = (int)(g % S);
      float o[3] = {ray_o[3 * r], ray_o[3 * r + 1], ray_o[3 * r + 2]};
      float d[3] = {ray_d[3 * r], ray_d[3 * r + 1], ray_d[3 * r + 2]};
      float x[4];
      if (BG) {
        float zv = z[(size_t)r * S + (S - 1 - j)];   // flipped order, ddp_model.py:116-117
        BgRay br = bg_ray_setup(o, d);
        float dr = bg_point(br, zv, x);
        if (q == 0) out_depth_real[g] = dr;
      } else {
        float zv = z[g];
        for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(o[c], __fmul_rn(zv, d[c]));   // ddp_model.py:91
      }
      embed_part<D>(x, NF_POS, q, erow);
      float dn = norm3(d[0], d[1], d[2]);
      float vd[3] = {d[0] / dn, d[1] / dn, d[2] / dn};                              // ddp_model.py:82-83
      embed_part<3>(vd, NF_VIEW, q, vrow);
    } else {
      for (int c = q; c < emb_dim(BG); c += 4) erow[c] = 0.f;
      for (int c = q; c < VIEW_DIM; c += 4) vrow[c] = 0.f;
    }
  }
  __syncthreads();

  // ---- base layers, nerf_network.py:126-131 ---------------------------------------------------
  dense<W>(packed + L.w[0], packed + L.b[0], E, EP, EP, nullptr, 0, 0, H0, W, true, wst);
  float* cur = H0;
  float* nxt = H1;
  for (int l = 1; l < 8; ++l) {
    if (l == 5) dense<W>(packed + L.w[l], packed + L.b[l], E, EP, EP, cur, W, W, nxt, W, true, wst);   // cat(input_pts, base)
    else dense<W>(packed + L.w[l], packed + L.b[l], cur, W, W, nullptr, 0, 0, nxt, W, true, wst);
    float* t = cur; cur = nxt; nxt = t;
  }
  // ---- sigma head, nerf_network.py:133-134 ----------------------------------------------------
  {
    int s = tid >> 2, q = tid & 3;
    const float* h = cur + s * W + q * 64;
    const float* ws = packed + L.w[L_SIGMA] + q * 64;
    float a = 0.f;
#pragma unroll 8
    for (int k = 0; k < 64; ++k) a = fmaf(h[k], ws[k], a);
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    long long g = g0 + s;
    if (q == 0 && g < total) out_sigma[g] = fabsf(a + packed[L.b[L_SIGMA]]);
  }
  // ---- colour head, nerf_network.py:136-138 ---------------------------------------------------
  dense<W>(packed + L.w[L_REMAP], packed + L.b[L_REMAP], cur, W, W, nullptr, 0, 0, nxt, W, false, wst);
  dense<RGB_HID>(packed + L.w[L_RGB0], packed + L.b[L_RGB0], nxt, W, W, V, VIEW_PAD, VIEW_PAD, cur, RGB_HID, true, wst);
  {
    int s = tid >> 2, q = tid & 3;
    const float* h = cur + s * RGB_HID + q * 32;
    float a[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < 3; ++c) {
      const float* w2 = packed + L.w[L_RGB2] + c * RGB_HID + q * 32;
#pragma unroll 8
      for (int k = 0; k < 32; ++k) a[c] = fmaf(h[k], w2[k], a[c]);
      a[c] += __shfl_xor_sync(0xffffffffu, a[c], 1);
      a[c] += __shfl_xor_sync(0xffffffffu, a[c], 2);
    }
    long long g = g0 + s;
    if (q < 3 && g < total) {
      float v = (q == 0 ? a[0] : q == 1 ? a[1] : a[2]) + packed[L.b[L_RGB2] + q];
      out_rgb[3 * g + q] = 1.f / (1.f + expf(-v));
    }
  }
}

// ---- packer --------------------------------------------------------------------------------------
__device__ __forceinline__ int simt_src_col(int l, int k, bool bg) {
  const int E = emb_pad_simt(bg), e = emb_dim(bg);
  if (l == 0) return k < e ? k : -1;
  if (l == 5) return k < E ? (k < e ? k : -1) : e + (k - E);
  if (l == L_RGB0) return k < W ? k : (k - W < VIEW_DIM ? k : -1);
  return k;
}
__global__ void pack_simt_kernel(NerfppNetParams p, bool bg, float* __restrict__ out) {
  const int l = blockIdx.y;
  const SimtLayout L = simt_layout(bg);
  const int rows = simt_rows(l, bg), nout = layer_out(l), nin = layer_in(l, bg);
  const float* Wl = p.w[l];
  const int tot = rows * nout;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < tot; idx += gridDim.x * blockDim.x) {
    if (l == L_SIGMA || l == L_RGB2) {
      out[L.w[l] + idx] = Wl[idx];   // kept [out][in]
    } else {
      int k = idx / nout, nn = idx % nout;
      int sk = simt_src_col(l, k, bg);
      out[L.w[l] + idx] = sk >= 0 ? Wl[(size_t)nn * nin + sk] : 0.f;
    }
  }
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < ((nout + 3) & ~3); i += blockDim.x) out[L.b[l] + i] = i < nout ? p.b[l][i] : 0.f;
}

}  // namespace npp

using namespace npp;

size_t npp_simt_packed_bytes(bool bg) { return (size_t)simt_layout(bg).total * sizeof(float); }

int npp_pack_simt(const NerfppNetParams* p, bool bg, void* out, cudaStream_t st) {
  pack_simt_kernel<<<dim3(64, NERFPP_NLAYERS), 256, 0, st>>>(*p, bg, (float*)out);
  NPP_CHECK_LAUNCH();
  return 0;
}

int npp_field_simt(const void* packed, bool bg, const float* ray_o, const float* ray_d, const float* z, int n, int S,
                   float* out_sigma, float* out_rgb, float* out_depth_real, cudaStream_t st) {
  long long total = (long long)n * S;
  unsigned grid = (unsigned)((total + TS - 1) / TS);
  size_t smem = (size_t)(TS * emb_pad_simt(bg) + TS * VIEW_PAD + 2 * TS * W + 2 * KC * W) * sizeof(float);
  if (bg) {
    cudaFuncSetAttribute(field_simt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    field_simt_kernel<true><<<grid, SIMT_THREADS, smem, st>>>((const float*)packed, ray_o, ray_d, z, n, S, out_sigma, out_rgb, out_depth_real);
  } else {
    cudaFuncSetAttribute(field_simt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    field_simt_kernel<false><<<grid, SIMT_THREADS, smem, st>>>((const float*)packed, ray_o, ray_d, z, n, S, out_sigma, out_rgb, out_depth_real);
  }
  NPP_CHECK_LAUNCH();
  return 0;
}
