// extern "C" entry points that tie the kernels together (include/nerfpp_b200.h).
#include <cstdarg>
#include <cstdio>
#include "common.cuh"

// defined in field_simt.cu / field_tc.cu
size_t npp_simt_packed_bytes(bool bg);
int npp_pack_simt(const NerfppNetParams* p, bool bg, void* out, cudaStream_t st);
int npp_field_simt(const void* packed, bool bg, const float* ray_o, const float* ray_d, const float* z, int n, int S,
                   float* out_sigma, float* out_rgb, float* out_depth_real, cudaStream_t st);
size_t npp_tc_packed_bytes(bool bg, bool prec);
int npp_pack_tc(const NerfppNetParams* p, bool bg, bool prec, void* out, cudaStream_t st);
int npp_field_tc(const void* packed, bool bg, bool prec, const float* ray_o, const float* ray_d, const float* z, int n, int S,
                 float* out_sigma, float* out_rgb, float* out_depth_real, void* train_ws, cudaStream_t st);
size_t npp_tc_train_ws_bytes(long long n_samples);

static thread_local char g_err[512] = "";

void npp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// fg / bg nets on two streams: 0 never, 1 (default) in the training entry points and for SMALL inference batches,
// 2 always (tests and diagnostics select 0 / 2)
static int g_overlap = 1;
extern "C" void nerfpp_debug_set_overlap(int mode) { g_overlap = (mode >= 0 && mode <= 2) ? mode : 1; }
static int g_overlap_tiles = 12 * 148;       // inference: fork when a net's launch has fewer 128-sample tiles than this
extern "C" void nerfpp_debug_set_overlap_tiles(int tiles) { g_overlap_tiles = tiles; }
NppFork* npp_fork_state() {
  static NppFork st[64];
  static bool made[64] = {false};
  if (!g_overlap) return nullptr;
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!made[dev]) {
    if (cudaStreamCreateWithFlags(&st[dev].side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    cudaEventCreateWithFlags(&st[dev].fork_ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&st[dev].join_ev, cudaEventDisableTiming);
    made[dev] = true;
  }
  return &st[dev];
}

extern "C" int nerfpp_abi_version(void) { return NERFPP_ABI_VERSION; }
extern "C" const char* nerfpp_last_error(void) { return g_err; }

extern "C" int64_t nerfpp_packed_bytes(int is_bg, int field_impl) {
  if (field_impl == NERFPP_FIELD_SIMT) return (int64_t)npp_simt_packed_bytes(is_bg != 0);
  if (field_impl == NERFPP_FIELD_TC || field_impl == NERFPP_FIELD_TC_SPLIT) return (int64_t)npp_tc_packed_bytes(is_bg != 0, field_impl == NERFPP_FIELD_TC_SPLIT);
  return -1;
}

extern "C" int nerfpp_pack_weights(const NerfppNetParams* params, int is_bg, int field_impl, void* out_packed, void* stream) {
  NPP_CHECK_ARG(params && out_packed, "null argument");
  for (int l = 0; l < NERFPP_NLAYERS; ++l) NPP_CHECK_ARG(params->w[l] && params->b[l], "null parameter tensor");
  if (field_impl == NERFPP_FIELD_SIMT) return npp_pack_simt(params, is_bg != 0, out_packed, (cudaStream_t)stream);
  if (field_impl == NERFPP_FIELD_TC || field_impl == NERFPP_FIELD_TC_SPLIT)
    return npp_pack_tc(params, is_bg != 0, field_impl == NERFPP_FIELD_TC_SPLIT, out_packed, (cudaStream_t)stream);
  NPP_CHECK_ARG(false, "unknown field_impl");
}

extern "C" int nerfpp_field_forward(const void* packed, int is_bg, int field_impl, const float* ray_o, const float* ray_d,
                                    const float* z, int n_rays, int n_samples, float* out_sigma, float* out_rgb,
                                    float* out_depth_real, void* stream) {
  NPP_CHECK_ARG(packed && ray_o && ray_d && z && out_sigma && out_rgb, "null argument");
  NPP_CHECK_ARG(n_rays >= 0 && n_samples >= 1, "bad shape");
  NPP_CHECK_ARG(!is_bg || out_depth_real, "background needs out_depth_real");
  if (n_rays == 0) return 0;
  if (field_impl == NERFPP_FIELD_SIMT)
    return npp_field_simt(packed, is_bg != 0, ray_o, ray_d, z, n_rays, n_samples, out_sigma, out_rgb, out_depth_real, (cudaStream_t)stream);
  if (field_impl == NERFPP_FIELD_TC || field_impl == NERFPP_FIELD_TC_SPLIT)
    return npp_field_tc(packed, is_bg != 0, field_impl == NERFPP_FIELD_TC_SPLIT, ray_o, ray_d, z, n_rays, n_samples, out_sigma, out_rgb, out_depth_real,
                        nullptr, (cudaStream_t)stream);
  NPP_CHECK_ARG(false, "unknown field_impl");
}

extern "C" int64_t nerfpp_field_train_workspace_bytes(int n_rays, int n_samples) {
  if (n_rays < 0 || n_samples < 1) return -1;
  return (int64_t)npp_tc_train_ws_bytes((long long)n_rays * n_samples);
}

extern "C" int nerfpp_field_forward_train(const void* packed, int is_bg, const float* ray_o, const float* ray_d, const float* z,
                                          int n_rays, int n_samples, float* out_sigma, float* out_rgb, float* out_depth_real,
                                          void* train_workspace, void* stream) {
  NPP_CHECK_ARG(packed && ray_o && ray_d && z && out_sigma && out_rgb && train_workspace, "null argument");
  NPP_CHECK_ARG(n_rays >= 0 && n_samples >= 1, "bad shape");
  NPP_CHECK_ARG(!is_bg || out_depth_real, "background needs out_depth_real");
  if (n_rays == 0) return 0;
  return npp_field_tc(packed, is_bg != 0, false, ray_o, ray_d, z, n_rays, n_samples, out_sigma, out_rgb, out_depth_real, train_workspace,
                      (cudaStream_t)stream);
}

// workspace of nerfpp_forward: fg sigma [n,Sf], fg rgb [n,Sf,3], bg sigma [n,Sb], bg rgb [n,Sb,3], bg depth_real [n,Sb]
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
struct FwdWs { float *fg_sigma, *fg_rgb, *bg_sigma, *bg_rgb, *bg_dr; size_t bytes; };
static FwdWs carve(void* base, int n, int sf, int sb) {
  FwdWs w;
  char* p = (char*)base;
  size_t o = 0;
  w.fg_sigma = (float*)(p + o); o += align256((size_t)n * sf * 4);
  w.fg_rgb = (float*)(p + o); o += align256((size_t)n * sf * 12);
  w.bg_sigma = (float*)(p + o); o += align256((size_t)n * sb * 4);
  w.bg_rgb = (float*)(p + o); o += align256((size_t)n * sb * 12);
  w.bg_dr = (float*)(p + o); o += align256((size_t)n * sb * 4);
  w.bytes = o;
  return w;
}

extern "C" int64_t nerfpp_forward_workspace_bytes(int n_rays, int s_fg, int s_bg) {
  if (n_rays < 0 || s_fg < 1 || s_bg < 1) return -1;
  return (int64_t)carve(nullptr, n_rays, s_fg, s_bg).bytes;
}

extern "C" int nerfpp_forward(const void* packed_fg, const void* packed_bg, int field_impl, const float* ray_o,
                              const float* ray_d, const float* fg_z_max, const float* fg_z, const float* bg_z, int n_rays,
                              int s_fg, int s_bg, const NerfppRenderOut* out, void* workspace, void* stream) {
  NPP_CHECK_ARG(packed_fg && packed_bg && workspace && out, "null argument");
  NPP_CHECK_ARG(n_rays >= 0 && s_fg >= 1 && s_bg >= 1, "bad shape");
  if (n_rays == 0) return 0;
  FwdWs w = carve(workspace, n_rays, s_fg, s_bg);
  // Inference: at 4096 rays per launch (14-42 waves of tiles) issuing the two nets on two streams measured neutral
  // (1.93 vs 1.93 ms per step) -- the tails are short; with few waves per launch (strong scaling: 512 rays per GPU =
  // 1.7 waves at the coarse level) the ragged last wave and the prologue are a large share, and the other net fills them.
  const long long tiles_fg = ((long long)n_rays * s_fg + 127) / 128;
  const bool fork = g_overlap == 2 || (g_overlap == 1 && tiles_fg < g_overlap_tiles);
  cudaStream_t side = fork ? npp_fork((cudaStream_t)stream) : (cudaStream_t)stream;
  int rc = nerfpp_field_forward(packed_fg, 0, field_impl, ray_o, ray_d, fg_z, n_rays, s_fg, w.fg_sigma, w.fg_rgb, nullptr, stream);
  int rc2 = nerfpp_field_forward(packed_bg, 1, field_impl, ray_o, ray_d, bg_z, n_rays, s_bg, w.bg_sigma, w.bg_rgb, w.bg_dr, (void*)side);
  if (fork) npp_join((cudaStream_t)stream, side);
  if (rc) return rc;
  if (rc2) return rc2;
  return nerfpp_composite(ray_d, fg_z_max, fg_z, bg_z, w.fg_sigma, w.fg_rgb, w.bg_sigma, w.bg_rgb, w.bg_dr, n_rays, s_fg,
                          s_bg, out, stream);
}

// ---- training-mode NerfNet.forward: also fills the training workspace [fg net | bg net] the backward reads ----
extern "C" int64_t nerfpp_forward_train_workspace_bytes(int n_rays, int s_fg, int s_bg) {
  if (n_rays < 0 || s_fg < 1 || s_bg < 1) return -1;
  const size_t fg = (npp_tc_train_ws_bytes((long long)n_rays * s_fg) + 1023) & ~(size_t)1023;
  return (int64_t)(fg + npp_tc_train_ws_bytes((long long)n_rays * s_bg) + 1024);
}

extern "C" int nerfpp_forward_train(const void* packed_fg, const void* packed_bg, const float* ray_o, const float* ray_d,
                                    const float* fg_z_max, const float* fg_z, const float* bg_z, int n_rays, int s_fg, int s_bg,
                                    const NerfppRenderOut* out, void* workspace, void* train_workspace, void* stream) {
  NPP_CHECK_ARG(packed_fg && packed_bg && workspace && train_workspace && out, "null argument");
  NPP_CHECK_ARG(n_rays >= 0 && s_fg >= 1 && s_bg >= 1, "bad shape");
  NPP_CHECK_ARG(((uintptr_t)train_workspace & 1023) == 0, "train_workspace must be 1024-byte aligned");
  if (n_rays == 0) return 0;
  FwdWs w = carve(workspace, n_rays, s_fg, s_bg);
  const size_t fg_bytes = (npp_tc_train_ws_bytes((long long)n_rays * s_fg) + 1023) & ~(size_t)1023;
  cudaStream_t side = npp_fork((cudaStream_t)stream);         // the two nets meet again in the composite (measured: -2 % per training step)
  int rc = nerfpp_field_forward_train(packed_fg, 0, ray_o, ray_d, fg_z, n_rays, s_fg, w.fg_sigma, w.fg_rgb, nullptr, train_workspace, stream);
  int rc2 = nerfpp_field_forward_train(packed_bg, 1, ray_o, ray_d, bg_z, n_rays, s_bg, w.bg_sigma, w.bg_rgb, w.bg_dr,
                                       (char*)train_workspace + fg_bytes, (void*)side);
  npp_join((cudaStream_t)stream, side);
  if (rc) return rc;
  if (rc2) return rc2;
  return nerfpp_composite(ray_d, fg_z_max, fg_z, bg_z, w.fg_sigma, w.fg_rgb, w.bg_sigma, w.bg_rgb, w.bg_dr, n_rays, s_fg,
                          s_bg, out, stream);
}
