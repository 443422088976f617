// C-ABI glue: error text, MLP dispatch, the render_rays chunk pipeline (DS_NeRF/run_nerf.py:593-737)
// and the host-buffer end-to-end entry point.
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace spn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error in %s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return SPN_E_CUDA;
}

int sm_count() {
  // per device: a process may drive several GPUs (one cached value would size every grid after the first device it saw)
  static std::atomic<int> cached[64];
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  const int c = cached[dev].load(std::memory_order_relaxed);
  if (c > 0) return c;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached[dev].store(n, std::memory_order_relaxed);
  return n;
}

std::atomic<long long> g_launches{0};

// ---- in-library kernel timing (bench.py's roofline leg): cudaEvent pairs on the launching stream ----
namespace {
constexpr int kProfMax = 4096;
struct ProfState {
  bool on = false;
  int n[PROF_KINDS] = {0, 0, 0};
  cudaEvent_t ev[PROF_KINDS][kProfMax][2];
  int created[PROF_KINDS] = {0, 0, 0};
} g_prof;
}  // namespace

void prof_begin(int kind, cudaStream_t st) {
  if (!g_prof.on || g_prof.n[kind] >= kProfMax) return;
  int i = g_prof.n[kind];
  if (i >= g_prof.created[kind]) {
    cudaEventCreate(&g_prof.ev[kind][i][0]);
    cudaEventCreate(&g_prof.ev[kind][i][1]);
    g_prof.created[kind] = i + 1;
  }
  cudaEventRecord(g_prof.ev[kind][i][0], st);
}
void prof_end(int kind, cudaStream_t st) {
  if (!g_prof.on || g_prof.n[kind] >= kProfMax) return;
  cudaEventRecord(g_prof.ev[kind][g_prof.n[kind]][1], st);
  ++g_prof.n[kind];
}

}  // namespace spn

using namespace spn;

extern "C" int spn_profile_enable(int on) {
  g_prof.on = on != 0;
  for (int k = 0; k < PROF_KINDS; ++k) g_prof.n[k] = 0;
  return SPN_OK;
}

extern "C" int spn_profile_read(int kind, int* launches, float* total_ms) {
  SPN_CHECK_ARG(kind >= 0 && kind < PROF_KINDS && launches && total_ms, "spn_profile_read: bad arguments");
  float tot = 0.f;
  for (int i = 0; i < g_prof.n[kind]; ++i) {
    float ms = 0.f;
    SPN_CUDA(cudaEventSynchronize(g_prof.ev[kind][i][1]));
    SPN_CUDA(cudaEventElapsedTime(&ms, g_prof.ev[kind][i][0], g_prof.ev[kind][i][1]));
    tot += ms;
  }
  *launches = g_prof.n[kind];
  *total_ms = tot;
  return SPN_OK;
}

extern "C" long long spn_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

extern "C" int spn_version(void) { return SPN_VERSION; }
extern "C" const char* spn_last_error(void) { return g_err; }

extern "C" int spn_device_info(int* sms, int* cc_major, int* cc_minor) {
  int dev = 0;
  SPN_CUDA(cudaGetDevice(&dev));
  int a = 0, b = 0, c = 0;
  SPN_CUDA(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
  SPN_CUDA(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
  SPN_CUDA(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
  if (sms) *sms = a;
  if (cc_major) *cc_major = b;
  if (cc_minor) *cc_minor = c;
  return SPN_OK;
}

extern "C" int spn_mlp_param_offsets(int64_t* offsets25_host) {
  SPN_CHECK_ARG(offsets25_host, "spn_mlp_param_offsets: null pointer");
  ParamOffsets po = param_offsets();
  for (int i = 0; i <= kNTensors; ++i) offsets25_host[i] = po.off[i];
  return SPN_OK;
}

extern "C" size_t spn_mlp_packed_bytes(void) { return mlp_tc_packed_bytes(); }

extern "C" int spn_mlp_pack_weights(const float* params_flat, void* packed, void* stream) {
  SPN_CHECK_ARG(params_flat && packed, "spn_mlp_pack_weights: null pointer");
  return mlp_tc_pack(params_flat, packed, as_stream(stream));
}

extern "C" int spn_tc_selftest_gemm(const float* A, const float* B, float* D, int N, int K, void* stream) {
  return tc_selftest_gemm(A, B, D, N, K, as_stream(stream));
}

extern "C" int spn_tc_set_trace(long long* stamps_dev) {
  tc_set_trace(stamps_dev);
  return SPN_OK;
}

extern "C" int spn_tc_tmem_ld_rate(int nwarps, int reps, int with_mma, long long* out_dev, void* stream) {
  return tc_tmem_ld_rate(nwarps, reps, with_mma, out_dev, as_stream(stream));
}

extern "C" int spn_tc_e4m3_decode(const uint8_t* codes_dev, uint16_t* bf16_out_dev, int n, void* stream) {
  return tc_e4m3_decode(codes_dev, bf16_out_dev, n, as_stream(stream));
}
extern "C" int spn_tc_mma_rate_pair(int ts, int n, int reps, int nacc, int ld_warps, long long* out_dev, void* stream) {
  return tc_mma_rate_pair(ts, n, reps, nacc, ld_warps, out_dev, as_stream(stream));
}
extern "C" int spn_tc_mma_rate(int a_mn_major, int b_mn_major, int n, int reps, long long* cycles_dev, void* stream) {
  SPN_CHECK_ARG(cycles_dev && reps > 0 && (n == 64 || n == 128 || n == 256), "spn_tc_mma_rate: bad arguments");
  return tc_mma_rate(a_mn_major, b_mn_major, n, reps, cycles_dev, as_stream(stream));
}

extern "C" int spn_tc_bulk_rate(const void* src, size_t src_bytes, int copy_bytes, int depth, int iters, int grid,
                                int lanes, long long* cycles_dev, void* stream) {
  return tc_bulk_rate(src, src_bytes, copy_bytes, depth, iters, grid, lanes, cycles_dev, as_stream(stream));
}

extern "C" size_t spn_mlp_stash_bytes(int64_t m, int precision) {
  return precision == SPN_PREC_FP32 ? mlp_fp32_stash_bytes(m) : mlp_tc_stash_bytes(m);
}
extern "C" size_t spn_mlp_bwd_workspace_bytes(int64_t m, int precision) {
  return precision == SPN_PREC_FP32 ? mlp_fp32_bwd_ws_bytes(m) : mlp_tc_bwd_ws_bytes(m);
}

static int mlp_fwd(const float* params, const void* packed, const SampleSource& src, int64_t m, float* raw,
                   void* stash, int precision, cudaStream_t st) {
  SPN_CHECK_ARG(raw && m >= 0, "spn_mlp_fwd: bad arguments");
  SPN_CHECK_ARG(((uintptr_t)raw & 15) == 0, "spn_mlp_fwd: raw must be 16-byte aligned");
  if (m == 0) return SPN_OK;
  if (precision == SPN_PREC_FP32) return mlp_fp32_fwd(params, src, m, raw, stash, st);
  SPN_CHECK_ARG(precision == SPN_PREC_BF16, "spn_mlp_fwd: unknown precision %d", precision);
  SPN_CHECK_ARG(packed, "spn_mlp_fwd: BF16 mode needs the packed weight image (spn_mlp_pack_weights)");
  return mlp_tc_fwd(packed, src, m, raw, stash, st);
}

extern "C" int spn_mlp_fwd_points(const float* params_flat, const void* packed, const float* x6, int64_t m,
                                  float* raw, void* stash, int precision, void* stream) {
  if (m == 0) return SPN_OK;
  SPN_CHECK_ARG(x6, "spn_mlp_fwd_points: null x6");
  SampleSource src{x6, nullptr, nullptr, 0, 1};
  return mlp_fwd(params_flat, packed, src, m, raw, stash, precision, as_stream(stream));
}

extern "C" int spn_mlp_fwd_rays(const float* params_flat, const void* packed, const float* rays, int ncols,
                                const float* z, int n, int S, float* raw, void* stash, int precision,
                                void* stream) {
  SPN_CHECK_ARG(rays && z && ncols >= 11 && n >= 0 && S >= 1, "spn_mlp_fwd_rays: bad arguments (ncols=%d)", ncols);
  SampleSource src{nullptr, rays, z, ncols, S};
  return mlp_fwd(params_flat, packed, src, (int64_t)n * S, raw, stash, precision, as_stream(stream));
}

extern "C" int spn_mlp_bwd(const float* params_flat, const void* packed, const void* stash, const float* d_raw,
                           int64_t m, float* grads_flat, void* workspace, int precision, void* stream) {
  SPN_CHECK_ARG(stash && d_raw && grads_flat && m >= 0, "spn_mlp_bwd: null pointer");
  if (m == 0) return SPN_OK;
  if (precision == SPN_PREC_FP32)
    return mlp_fp32_bwd(params_flat, stash, d_raw, m, grads_flat, workspace, as_stream(stream));
  SPN_CHECK_ARG(precision == SPN_PREC_BF16 && packed, "spn_mlp_bwd: BF16 mode needs the packed weight image");
  return mlp_tc_bwd(packed, stash, d_raw, m, grads_flat, workspace, as_stream(stream));
}

// ---------------------------------------------------------------------------------------------
// render_rays, one chunk
// ---------------------------------------------------------------------------------------------
#define RET_IF(x) do { int rc__ = (x); if (rc__ != SPN_OK) return rc__; } while (0)

static int check_cfg(const spn_render_cfg* c, const spn_render_io* io) {
  SPN_CHECK_ARG(c && io, "spn_render_rays: null cfg/io");
  SPN_CHECK_ARG(c->n_rays >= 0 && (c->ncols == 11 || c->ncols == 12), "spn_render_rays: ncols=%d (need 11|12: viewdirs are required)", c->ncols);
  SPN_CHECK_ARG(c->n_samples >= 2 && c->n_importance >= 0, "spn_render_rays: bad sample counts");
  SPN_CHECK_ARG(io->rays && io->rgb_map && io->disp_map && io->acc_map && io->depth_map && io->weights &&
                io->z_vals && io->raw, "spn_render_rays: null required buffer");
  if (c->n_importance > 0)
    SPN_CHECK_ARG(io->z_coarse && io->raw_coarse && io->rgb0 && io->disp0 && io->acc0 && io->z_std,
                  "spn_render_rays: the fine pass needs z_coarse/raw_coarse/rgb0/disp0/acc0/z_std buffers");
  if (c->flags & SPN_F_PERTURB)
    SPN_CHECK_ARG(io->t_rand && (c->n_importance == 0 || io->u), "spn_render_rays: SPN_F_PERTURB needs t_rand and u");
  if (c->raw_noise_std > 0.f)
    SPN_CHECK_ARG(io->noise0 && (c->n_importance == 0 || io->noise1), "spn_render_rays: raw_noise_std>0 needs noise draws");
  if (c->flags & SPN_F_NEED_ALPHA)   // the reference raises NameError here (run_nerf.py:719-721)
    SPN_CHECK_ARG(c->n_importance > 0 && io->alpha && io->alpha0, "spn_render_rays: need_alpha requires N_importance>0");
  return SPN_OK;
}

extern "C" int spn_render_rays_fwd(const spn_render_cfg* c, const spn_render_io* io, void* stream) {
  RET_IF(check_cfg(c, io));
  cudaStream_t st = as_stream(stream);
  const int n = c->n_rays, S = c->n_samples, NI = c->n_importance, S2 = S + NI;
  if (n == 0) return SPN_OK;
  const int white = (c->flags & SPN_F_WHITE_BKGD) != 0, lindisp = (c->flags & SPN_F_LINDISP) != 0;
  const bool perturb = (c->flags & SPN_F_PERTURB) != 0;
  const bool noisy = c->raw_noise_std > 0.f;
  const float* rays_d = io->rays + 3;
  float* zc = NI > 0 ? io->z_coarse : io->z_vals;
  float* rawc = NI > 0 ? io->raw_coarse : io->raw;
  RET_IF(spn_sample_z(io->rays, n, c->ncols, S, lindisp, perturb ? io->t_rand : nullptr, zc, stream));
  RET_IF(spn_mlp_fwd_rays(io->params_coarse, io->packed_coarse, io->rays, c->ncols, zc, n, S, rawc, io->stash_coarse,
                          c->precision, stream));
  if (NI == 0) {
    return composite_fwd(rawc, zc, rays_d, c->ncols, noisy ? io->noise0 : nullptr, c->raw_noise_std, n, S, white,
                         io->rgb_map, io->disp_map, io->acc_map, io->weights, io->depth_map, nullptr, st);
  }
  // coarse maps; its weights [n,S] live temporarily in io->weights, depth in io->depth_map
  RET_IF(composite_fwd(rawc, zc, rays_d, c->ncols, noisy ? io->noise0 : nullptr, c->raw_noise_std, n, S, white,
                       io->rgb0, io->disp0, io->acc0, io->weights, io->depth_map,
                       (c->flags & SPN_F_NEED_ALPHA) ? io->alpha0 : nullptr, st));
  RET_IF(spn_resample(zc, io->weights, perturb ? io->u : nullptr, n, S, NI, io->z_vals, nullptr, io->z_std, nullptr,
                      stream));
  const float* pf = io->params_fine ? io->params_fine : io->params_coarse;
  const void* kf = io->params_fine ? io->packed_fine : io->packed_coarse;
  RET_IF(spn_mlp_fwd_rays(pf, kf, io->rays, c->ncols, io->z_vals, n, S2, io->raw, io->stash_fine, c->precision, stream));
  return composite_fwd(io->raw, io->z_vals, rays_d, c->ncols, noisy ? io->noise1 : nullptr, c->raw_noise_std, n, S2,
                       white, io->rgb_map, io->disp_map, io->acc_map, io->weights, io->depth_map,
                       (c->flags & SPN_F_NEED_ALPHA) ? io->alpha : nullptr, st);
}

extern "C" int spn_render_rays_bwd(const spn_render_cfg* c, const spn_render_io* io, const spn_render_grads* g,
                                   void* stream) {
  RET_IF(check_cfg(c, io));
  SPN_CHECK_ARG(g && g->grads_coarse && g->d_raw_scratch, "spn_render_rays_bwd: null gradient buffers");
  cudaStream_t st = as_stream(stream);
  const int n = c->n_rays, S = c->n_samples, NI = c->n_importance, S2 = S + NI;
  if (n == 0) return SPN_OK;
  const int white = (c->flags & SPN_F_WHITE_BKGD) != 0, detach = (c->flags & SPN_F_DETACH_WEIGHTS) != 0;
  const bool noisy = c->raw_noise_std > 0.f;
  const float* rays_d = io->rays + 3;
  if (NI == 0) {
    SPN_CHECK_ARG(io->stash_coarse, "spn_render_rays_bwd: forward ran without a stash");
    RET_IF(composite_bwd(io->raw, io->z_vals, rays_d, c->ncols, noisy ? io->noise0 : nullptr, c->raw_noise_std, n, S,
                         white, detach, g->detach_begin, g->detach_end, g->g_rgb, g->g_disp, g->g_acc, g->g_weights, g->g_depth, g->d_raw_scratch, st));
    return spn_mlp_bwd(io->params_coarse, io->packed_coarse, io->stash_coarse, g->d_raw_scratch, (int64_t)n * S,
                       g->grads_coarse, g->workspace, c->precision, stream);
  }
  SPN_CHECK_ARG(io->stash_coarse && io->stash_fine, "spn_render_rays_bwd: forward ran without a stash");
  const float* pf = io->params_fine ? io->params_fine : io->params_coarse;
  const void* kf = io->params_fine ? io->packed_fine : io->packed_coarse;
  float* gf = io->params_fine ? g->grads_fine : g->grads_coarse;
  SPN_CHECK_ARG(gf, "spn_render_rays_bwd: null grads_fine");
  if (g->g_rgb || g->g_disp || g->g_acc || g->g_weights || g->g_depth) {
    RET_IF(composite_bwd(io->raw, io->z_vals, rays_d, c->ncols, noisy ? io->noise1 : nullptr, c->raw_noise_std, n, S2,
                         white, detach, g->detach_begin, g->detach_end, g->g_rgb, g->g_disp, g->g_acc, g->g_weights, g->g_depth, g->d_raw_scratch, st));
    RET_IF(spn_mlp_bwd(pf, kf, io->stash_fine, g->d_raw_scratch, (int64_t)n * S2, gf, g->workspace, c->precision, stream));
  }
  // coarse pass: z_samples are detached (run_nerf.py:700) so only rgb0/disp0/acc0 carry gradient
  if (g->g_rgb0 || g->g_disp0 || g->g_acc0) {
    RET_IF(composite_bwd(io->raw_coarse, io->z_coarse, rays_d, c->ncols, noisy ? io->noise0 : nullptr,
                         c->raw_noise_std, n, S, white, detach, g->detach_begin, g->detach_end, g->g_rgb0, g->g_disp0, g->g_acc0, nullptr, nullptr,
                         g->d_raw_scratch, st));
    RET_IF(spn_mlp_bwd(io->params_coarse, io->packed_coarse, io->stash_coarse, g->d_raw_scratch, (int64_t)n * S,
                       g->grads_coarse, g->workspace, c->precision, stream));
  }
  return SPN_OK;
}

// ---------------------------------------------------------------------------------------------
// host-buffer entry point
// ---------------------------------------------------------------------------------------------
namespace {
struct HostArena {
  char* base = nullptr;
  size_t cap = 0, used = 0;
  std::mutex mu;
  int reserve(size_t bytes) {
    if (bytes <= cap) return SPN_OK;
    if (base) cudaFree(base);
    base = nullptr; cap = 0;
    SPN_CUDA(cudaMalloc(&base, bytes));
    cap = bytes;
    return SPN_OK;
  }
  void* take(size_t bytes) {
    size_t a = (used + 255) & ~(size_t)255;
    used = a + bytes;
    return base + a;
  }
};
HostArena g_arena;
}  // namespace

extern "C" int spn_render_host(const spn_render_cfg* c, const float* rays_host, const float* params_coarse_host,
                               const float* params_fine_host, float* rgb_host, float* disp_host, float* acc_host,
                               float* depth_host) {
  SPN_CHECK_ARG(c && rays_host && params_coarse_host && rgb_host && disp_host && acc_host && depth_host,
                "spn_render_host: null pointer");
  SPN_CHECK_ARG(!(c->flags & SPN_F_PERTURB) && c->raw_noise_std == 0.f && !(c->flags & SPN_F_NEED_ALPHA),
                "spn_render_host: deterministic rendering only (render_kwargs_test, run_nerf.py:485-488)");
  std::lock_guard<std::mutex> lock(g_arena.mu);
  const int64_t n = c->n_rays, S = c->n_samples, NI = c->n_importance, S2 = S + NI;
  if (n == 0) return SPN_OK;
  const size_t f = sizeof(float);
  size_t stash = c->precision == SPN_PREC_FP32 ? mlp_fp32_stash_bytes(n * S2) : 0;
  size_t need = 64 * 256 + n * c->ncols * f + 2 * SPN_MLP_NPARAMS * f + 2 * mlp_tc_packed_bytes() +
                n * (3 + 3 + 3 + 3 + 1) * f * 2 + n * S2 * 6 * f + n * S * 5 * f + stash;
  RET_IF(g_arena.reserve(need));
  g_arena.used = 0;
  cudaStream_t st = 0;
  spn_render_io io;
  memset(&io, 0, sizeof(io));
  float* rays = (float*)g_arena.take(n * c->ncols * f);
  float* pc = (float*)g_arena.take(SPN_MLP_NPARAMS * f);
  float* pf = params_fine_host ? (float*)g_arena.take(SPN_MLP_NPARAMS * f) : nullptr;
  SPN_CUDA(cudaMemcpyAsync(rays, rays_host, n * c->ncols * f, cudaMemcpyHostToDevice, st));
  SPN_CUDA(cudaMemcpyAsync(pc, params_coarse_host, SPN_MLP_NPARAMS * f, cudaMemcpyHostToDevice, st));
  if (pf) SPN_CUDA(cudaMemcpyAsync(pf, params_fine_host, SPN_MLP_NPARAMS * f, cudaMemcpyHostToDevice, st));
  io.rays = rays; io.params_coarse = pc; io.params_fine = pf;
  if (c->precision == SPN_PREC_BF16) {
    void* kc = g_arena.take(mlp_tc_packed_bytes());
    RET_IF(mlp_tc_pack(pc, kc, st));
    io.packed_coarse = kc;
    if (pf) {
      void* kf = g_arena.take(mlp_tc_packed_bytes());
      RET_IF(mlp_tc_pack(pf, kf, st));
      io.packed_fine = kf;
    }
  } else {
    io.stash_coarse = g_arena.take(stash);   // fp32 mode uses the stash as its activation workspace
    io.stash_fine = io.stash_coarse;
  }
  io.rgb_map = (float*)g_arena.take(n * 3 * f); io.disp_map = (float*)g_arena.take(n * f);
  io.acc_map = (float*)g_arena.take(n * f); io.depth_map = (float*)g_arena.take(n * f);
  io.weights = (float*)g_arena.take(n * S2 * f); io.z_vals = (float*)g_arena.take(n * S2 * f);
  io.raw = (float*)g_arena.take(n * S2 * 4 * f);
  io.rgb0 = (float*)g_arena.take(n * 3 * f); io.disp0 = (float*)g_arena.take(n * f);
  io.acc0 = (float*)g_arena.take(n * f); io.z_std = (float*)g_arena.take(n * f);
  io.z_coarse = (float*)g_arena.take(n * S * f); io.raw_coarse = (float*)g_arena.take(n * S * 4 * f);
  RET_IF(spn_render_rays_fwd(c, &io, st));
  SPN_CUDA(cudaMemcpyAsync(rgb_host, io.rgb_map, n * 3 * f, cudaMemcpyDeviceToHost, st));
  SPN_CUDA(cudaMemcpyAsync(disp_host, io.disp_map, n * f, cudaMemcpyDeviceToHost, st));
  SPN_CUDA(cudaMemcpyAsync(acc_host, io.acc_map, n * f, cudaMemcpyDeviceToHost, st));
  SPN_CUDA(cudaMemcpyAsync(depth_host, io.depth_map, n * f, cudaMemcpyDeviceToHost, st));
  SPN_CUDA(cudaStreamSynchronize(st));
  return SPN_OK;
}
