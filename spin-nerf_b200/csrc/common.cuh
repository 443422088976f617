// Shared declarations for the spinnerf_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/spinnerf_b200.h"

namespace spn {

// ---- error plumbing (thread-local text behind spn_last_error) ---------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define SPN_CHECK_ARG(cond, ...)                      \
  do {                                                \
    if (!(cond)) {                                    \
      spn::set_error(__VA_ARGS__);                    \
      return SPN_E_ARG;                               \
    }                                                 \
  } while (0)

#define SPN_CUDA(call)                                               \
  do {                                                               \
    cudaError_t e__ = (call);                                        \
    if (e__ != cudaSuccess) return spn::cuda_fail(e__, #call);       \
  } while (0)

#define SPN_LAUNCH_CHECK(name)                                       \
  do {                                                               \
    ++spn::g_launches;                                               \
    cudaError_t e__ = cudaGetLastError();                            \
    if (e__ != cudaSuccess) return spn::cuda_fail(e__, name);        \
  } while (0)

extern std::atomic<long long> g_launches;   // kernels launched by this library (spn_launch_count); host threads may launch concurrently

// optional in-library CUDA-event timing of the dominant kernels, on the launching stream
enum ProfKind : int { PROF_MLP_FWD = 0, PROF_MLP_DGRAD = 1, PROF_MLP_WGRAD = 2, PROF_KINDS = 3 };
void prof_begin(int kind, cudaStream_t st);
void prof_end(int kind, cudaStream_t st);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
int sm_count();

// ---- MLP geometry (DS_NeRF/run_nerf_helpers.py:86-102, use_viewdirs, D=8, W=256) ---------
constexpr int kW = 256;        // hidden width
constexpr int kEncP = 63;      // gamma(pts), L=10
constexpr int kEncD = 27;      // gamma(viewdir), L=4
constexpr int kWV = 128;       // views_linears.0 width
constexpr int kNTensors = SPN_MLP_NTENSORS;

struct ParamOffsets {
  int64_t off[kNTensors + 1];
};
// tensor ids in flat order
enum : int {
  T_W0 = 0, T_B0 = 1,            // pts_linears.i -> 2i, 2i+1
  T_WV = 16, T_BV = 17, T_WF = 18, T_BF = 19, T_WA = 20, T_BA = 21, T_WR = 22, T_BR = 23
};
__host__ __device__ inline ParamOffsets param_offsets() {
  ParamOffsets p;
  int64_t o = 0;
  for (int i = 0; i < 8; ++i) {
    int in = (i == 0) ? kEncP : (i == 5 ? kW + kEncP : kW);
    p.off[2 * i] = o; o += (int64_t)kW * in;
    p.off[2 * i + 1] = o; o += kW;
  }
  p.off[T_WV] = o; o += (int64_t)kWV * (kW + kEncD);
  p.off[T_BV] = o; o += kWV;
  p.off[T_WF] = o; o += (int64_t)kW * kW;
  p.off[T_BF] = o; o += kW;
  p.off[T_WA] = o; o += kW;
  p.off[T_BA] = o; o += 1;
  p.off[T_WR] = o; o += 3 * kWV;
  p.off[T_BR] = o; o += 3;
  p.off[kNTensors] = o;
  return p;
}

// ---- how the MLP kernels fetch a sample's 3-D point and view direction ---------------------
// mode POINTS: x6[m,6] rows = [pt, dir];  mode RAYS: rays[n,ncols] + z[n,S], m = n*S + s.
struct SampleSource {
  const float* x6;     // POINTS
  const float* rays;   // RAYS
  const float* z;
  int ncols;
  int S;
};
__device__ __forceinline__ void fetch_sample(const SampleSource& src, int64_t row, float pt[3],
                                             float dir[3]) {
  if (src.x6) {
    const float* p = src.x6 + row * 6;
    pt[0] = p[0]; pt[1] = p[1]; pt[2] = p[2];
    dir[0] = p[3]; dir[1] = p[4]; dir[2] = p[5];
  } else {
    int64_t r = row / src.S;
    const float* ray = src.rays + r * src.ncols;
    float zz = src.z[row];
    // pts = rays_o + rays_d * z   (run_nerf.py:670) — mul and add rounded separately like torch
    pt[0] = __fadd_rn(ray[0], __fmul_rn(ray[3], zz));
    pt[1] = __fadd_rn(ray[1], __fmul_rn(ray[4], zz));
    pt[2] = __fadd_rn(ray[2], __fmul_rn(ray[5], zz));
    dir[0] = ray[src.ncols - 3]; dir[1] = ray[src.ncols - 2]; dir[2] = ray[src.ncols - 1];
  }
}

// compositing launchers with the raw_noise_std scale applied in-kernel (ops_render.cu)
int composite_fwd(const float* raw, const float* z, const float* rays_d, int ld_d, const float* noise,
                  float noise_scale, int n, int S, int white_bkgd, float* rgb_map, float* disp_map, float* acc_map,
                  float* weights, float* depth_map, float* alpha, cudaStream_t stream);
int composite_bwd(const float* raw, const float* z, const float* rays_d, int ld_d, const float* noise,
                  float noise_scale, int n, int S, int white_bkgd, int detach_weights, int detach_begin, int detach_end,
                  const float* g_rgb, const float* g_disp, const float* g_acc, const float* g_weights,
                  const float* g_depth, float* d_raw, cudaStream_t stream);

// ---- kernels implemented in the other translation units -------------------------------------
// fp32 CUDA-core MLP (mlp_fp32.cu)
size_t mlp_fp32_stash_bytes(int64_t m);
size_t mlp_fp32_bwd_ws_bytes(int64_t m);
int mlp_fp32_fwd(const float* params, const SampleSource& src, int64_t m, float* raw, void* stash,
                 cudaStream_t st);
int mlp_fp32_bwd(const float* params, const void* stash, const float* d_raw, int64_t m,
                 float* grads, void* ws, cudaStream_t st);
// tcgen05 MLP (mlp_tc.cu)
size_t mlp_tc_packed_bytes();
void tc_set_trace(long long* dev);
long long* tc_get_trace();
size_t mlp_tc_stash_bytes(int64_t m);
size_t mlp_tc_bwd_ws_bytes(int64_t m);
int mlp_tc_pack(const float* params, void* packed, cudaStream_t st);
int mlp_tc_fwd(const void* packed, const SampleSource& src, int64_t m, float* raw, void* stash,
               cudaStream_t st);
int tc_bulk_rate(const void* src, size_t src_bytes, int copy_bytes, int depth, int iters, int grid, int lanes,
                 long long* out, cudaStream_t st);
int tc_tmem_ld_rate(int nwarps, int reps, int with_mma, long long* out, cudaStream_t st);
int tc_mma_rate(int a_mn, int b_mn, int n, int reps, long long* cycles_dev, cudaStream_t st);
int tc_e4m3_decode(const uint8_t* codes, uint16_t* out, int n, cudaStream_t st);
int tc_mma_rate_pair(int ts, int n, int reps, int nacc, int ld_warps, long long* out, cudaStream_t st);
int tc_selftest_gemm(const float* A, const float* B, float* D, int N, int K, cudaStream_t st);
int mlp_tc_bwd(const void* packed, const void* stash, const float* d_raw, int64_t m, float* grads,
               void* ws, cudaStream_t st);

}  // namespace spn
