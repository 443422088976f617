// Ray generation, stratified depths, alpha compositing (fwd/bwd), inverse-CDF resampling,
// positional encoding and flat Adam for the SPIn-NeRF render hot path.  sm_100a.
//
// These are the HBM/latency-bound pieces around the MLP: one warp per ray, samples strided
// over lanes (coalesced 128 B / 512 B rows), warp-shuffle scans for the transmittance product
// and its reverse-mode suffix sums.  Parity-critical arithmetic uses __fmul_rn/__fadd_rn so
// nvcc does not contract it into FMAs the reference's separate torch ops do not have.
#include <math.h>

#include "common.cuh"

namespace spn {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// torch.linspace(0,1,n)[i] as torch's CUDA kernel evaluates it (symmetric halves, one fma)
__device__ __forceinline__ float linspace01(int i, int n) {
  float step = __fdiv_rn(1.0f, (float)(n - 1));
  return (i < n / 2) ? __fmul_rn(step, (float)i) : fmaf(-step, (float)(n - 1 - i), 1.0f);
}

// ------------------------------------------------------------------------------------------
// a9 get_rays / ndc_rays / ray-matrix assembly
// ------------------------------------------------------------------------------------------
__global__ void get_rays_kernel(const float* __restrict__ c2w, int H, int W, float focal, int i0,
                                int j0, int h, int w, float* __restrict__ ro,
                                float* __restrict__ rd) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= h * w) return;
  int row = i0 + idx / w, col = j0 + idx % w;   // patch slices rows first (run_nerf.py:121-123)
  float dx = __fdiv_rn(__fsub_rn((float)col, __fmul_rn((float)W, 0.5f)), focal);
  float dy = -__fdiv_rn(__fsub_rn((float)row, __fmul_rn((float)H, 0.5f)), focal);
  float dz = -1.0f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float* r = c2w + 4 * k;
    float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, r[0]), __fmul_rn(dy, r[1])), __fmul_rn(dz, r[2]));
    rd[(int64_t)idx * 3 + k] = s;
    ro[(int64_t)idx * 3 + k] = r[3];
  }
}

__device__ __forceinline__ void ndc_one(int H, int W, float focal, float near_, const float o[3],
                                        const float d[3], float oo[3], float dd[3]) {
  // helpers:285-298
  float t = -(near_ + o[2]) / d[2];
  float ox = o[0] + t * d[0], oy = o[1] + t * d[1], oz = o[2] + t * d[2];
  float sx = -1.0f / (W / (2.0f * focal)), sy = -1.0f / (H / (2.0f * focal));
  oo[0] = sx * ox / oz;
  oo[1] = sy * oy / oz;
  oo[2] = 1.0f + 2.0f * near_ / oz;
  dd[0] = sx * (d[0] / d[2] - ox / oz);
  dd[1] = sy * (d[1] / d[2] - oy / oz);
  dd[2] = -2.0f * near_ / oz;
}

__global__ void ndc_rays_kernel(int n, int H, int W, float focal, float near_,
                                const float* __restrict__ ro, const float* __restrict__ rd,
                                float* __restrict__ oo, float* __restrict__ od) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float o[3] = {ro[3 * i], ro[3 * i + 1], ro[3 * i + 2]};
  float d[3] = {rd[3 * i], rd[3 * i + 1], rd[3 * i + 2]};
  float a[3], b[3];
  ndc_one(H, W, focal, near_, o, d, a, b);
  for (int k = 0; k < 3; ++k) { oo[3 * i + k] = a[k]; od[3 * i + k] = b[k]; }
}

__global__ void build_ray_batch_kernel(int n, const float* __restrict__ ro,
                                       const float* __restrict__ rd, float near_, float far_,
                                       int ndc, int H, int W, float focal,
                                       float* __restrict__ rays) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float o[3] = {ro[3 * i], ro[3 * i + 1], ro[3 * i + 2]};
  float d[3] = {rd[3 * i], rd[3 * i + 1], rd[3 * i + 2]};
  // viewdirs = rays_d / ||rays_d||  BEFORE the NDC map (run_nerf.py:128-140)
  float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  float* r = rays + (int64_t)i * 11;
  r[8] = d[0] / nrm; r[9] = d[1] / nrm; r[10] = d[2] / nrm;
  if (ndc) {
    float a[3], b[3];
    ndc_one(H, W, focal, 1.0f, o, d, a, b);
    for (int k = 0; k < 3; ++k) { o[k] = a[k]; d[k] = b[k]; }
  }
  r[0] = o[0]; r[1] = o[1]; r[2] = o[2];
  r[3] = d[0]; r[4] = d[1]; r[5] = d[2];
  r[6] = near_; r[7] = far_;
}

// batch assembly from a device-resident ray pool: what RayDataset.__getitem__ + DataLoader collation + render()'s ray
// matrix (data.py:4-15, run_nerf.py:126-153, 1367-1413) produce for N_rand sampled rays, in one launch
__global__ void gather_ray_batch_kernel(int n, const float* __restrict__ pool_o, const float* __restrict__ pool_d,
                                        const int64_t* __restrict__ idx, float near_, float far_, int ndc, int H, int W,
                                        float focal, float* __restrict__ rays, const float* __restrict__ rgb_pool,
                                        float* __restrict__ rgb_out, int n_rgb, const float* __restrict__ disp_pool,
                                        float* __restrict__ disp_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t j = idx[i];
  float o[3] = {pool_o[3 * j], pool_o[3 * j + 1], pool_o[3 * j + 2]};
  float d[3] = {pool_d[3 * j], pool_d[3 * j + 1], pool_d[3 * j + 2]};
  float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  float* r = rays + (int64_t)i * 11;
  r[8] = d[0] / nrm; r[9] = d[1] / nrm; r[10] = d[2] / nrm;
  if (ndc) {
    float a[3], b[3];
    ndc_one(H, W, focal, 1.0f, o, d, a, b);
    for (int k = 0; k < 3; ++k) { o[k] = a[k]; d[k] = b[k]; }
  }
  r[0] = o[0]; r[1] = o[1]; r[2] = o[2];
  r[3] = d[0]; r[4] = d[1]; r[5] = d[2];
  r[6] = near_; r[7] = far_;
  if (i < n_rgb) {
    if (rgb_pool) { rgb_out[3 * i] = rgb_pool[3 * j]; rgb_out[3 * i + 1] = rgb_pool[3 * j + 1]; rgb_out[3 * i + 2] = rgb_pool[3 * j + 2]; }
  } else if (disp_pool) {
    disp_out[i - n_rgb] = disp_pool[j];
  }
}

// ------------------------------------------------------------------------------------------
// a5 positional encoding (standalone op; the MLP kernels encode in-kernel)
// ------------------------------------------------------------------------------------------
__global__ void embed_kernel(const float* __restrict__ x, int64_t m, int L, float* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int width = 3 + 6 * L;
  if (idx >= m * width) return;
  int64_t row = idx / width;
  int c = (int)(idx % width);
  float v;
  if (c < 3) {
    v = x[row * 3 + c];
  } else {
    int k = (c - 3) / 6, r = (c - 3) % 6;
    float a = __fmul_rn(x[row * 3 + (r % 3)], (float)(1 << k));   // exact power-of-two scale
    v = (r < 3) ? sinf(a) : cosf(a);
  }
  out[idx] = v;
}

// ------------------------------------------------------------------------------------------
// a3 stratified depths
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float z_at(float near_, float far_, int i, int S, int lindisp) {
  float t = linspace01(i, S);
  float omt = __fsub_rn(1.0f, t);
  if (!lindisp) return __fadd_rn(__fmul_rn(near_, omt), __fmul_rn(far_, t));
  float a = __fmul_rn(__fdiv_rn(1.0f, near_), omt);
  float b = __fmul_rn(__fdiv_rn(1.0f, far_), t);
  return __fdiv_rn(1.0f, __fadd_rn(a, b));
}

__global__ void sample_z_kernel(const float* __restrict__ rays, int n, int ncols, int S,
                                int lindisp, const float* __restrict__ t_rand,
                                float* __restrict__ z) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * S) return;
  int r = (int)(idx / S), i = (int)(idx % S);
  float near_ = rays[(int64_t)r * ncols + 6], far_ = rays[(int64_t)r * ncols + 7];
  float zi = z_at(near_, far_, i, S, lindisp);
  if (t_rand) {   // run_nerf.py:654-668
    float lo = zi, hi = zi;
    if (i > 0) lo = __fmul_rn(0.5f, __fadd_rn(zi, z_at(near_, far_, i - 1, S, lindisp)));
    if (i < S - 1) hi = __fmul_rn(0.5f, __fadd_rn(z_at(near_, far_, i + 1, S, lindisp), zi));
    zi = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), t_rand[idx]));
  }
  z[idx] = zi;
}

// ------------------------------------------------------------------------------------------
// a7 raw2outputs: one warp per ray, front-to-back warp-scan of the transmittance product
// ------------------------------------------------------------------------------------------
constexpr int kMaxChunks = 8;   // S <= 256

struct RaySample {
  float alpha, trans, w, e, dist, t_in;
};

// forward sweep shared by fwd and bwd.  carry = prod of (1-alpha+1e-10) over earlier chunks.
__device__ __forceinline__ RaySample composite_step(float sigma, float dist, float& carry, int lane) {
  RaySample r;
  r.dist = dist;
  r.e = expf(-fmaxf(sigma, 0.0f) * dist);
  r.alpha = 1.0f - r.e;
  r.t_in = 1.0f - r.alpha + 1e-10f;
  float p = r.t_in;   // inclusive product scan over the 32 lanes
#pragma unroll
  for (int o = 1; o < kWarp; o <<= 1) {
    float q = __shfl_up_sync(kFull, p, o);
    if (lane >= o) p *= q;
  }
  float excl = __shfl_up_sync(kFull, p, 1);
  if (lane == 0) excl = 1.0f;
  r.trans = carry * excl;
  r.w = r.alpha * r.trans;
  carry *= __shfl_sync(kFull, p, kWarp - 1);
  return r;
}

__global__ void __launch_bounds__(256)
raw2outputs_fwd_kernel(const float4* __restrict__ raw, const float* __restrict__ z,
                       const float* __restrict__ rays_d, int ld_d, const float* __restrict__ noise,
                       float noise_scale, int n, int S, int white_bkgd, float* __restrict__ rgb_map,
                       float* __restrict__ disp_map, float* __restrict__ acc_map,
                       float* __restrict__ weights, float* __restrict__ depth_map,
                       float* __restrict__ alpha_out) {
  int ray = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  int lane = threadIdx.x % kWarp;
  if (ray >= n) return;
  const float* d = rays_d + (int64_t)ray * ld_d;
  float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);   // torch.norm (helpers:369)
  float carry = 1.0f, sr = 0, sg = 0, sb = 0, sdepth = 0, sacc = 0;
  for (int base = 0; base < S; base += kWarp) {
    int s = base + lane;
    bool live = s < S;
    int64_t idx = (int64_t)ray * S + (live ? s : S - 1);
    float4 rw = raw[idx];
    float zi = z[idx];
    float znext = (live && s + 1 < S) ? z[idx + 1] : 0.0f;
    float dist = (s + 1 < S) ? __fsub_rn(znext, zi) : 1e10f;   // helpers:366-367
    dist = __fmul_rn(dist, nrm);
    float sigma = rw.w + (noise ? noise[idx] * noise_scale : 0.0f);
    if (!live) { sigma = 0.0f; dist = 0.0f; }                   // alpha = 0, t = 1 (+1e-10)
    RaySample r = composite_step(sigma, dist, carry, lane);
    if (live) {
      weights[idx] = r.w;
      if (alpha_out) alpha_out[idx] = r.alpha;
      sr += r.w * (1.0f / (1.0f + expf(-rw.x)));
      sg += r.w * (1.0f / (1.0f + expf(-rw.y)));
      sb += r.w * (1.0f / (1.0f + expf(-rw.z)));
      sdepth += r.w * zi;
      sacc += r.w;
    }
  }
  sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb);
  sdepth = warp_sum(sdepth); sacc = warp_sum(sacc);
  if (lane == 0) {
    float ratio = sdepth / sacc;                 // NaN when acc == 0, like the reference (:391)
    float m = (ratio != ratio) ? ratio : fmaxf(1e-10f, ratio);
    disp_map[ray] = 1.0f / m;
    acc_map[ray] = sacc;
    depth_map[ray] = sdepth;
    float bg = white_bkgd ? (1.0f - sacc) : 0.0f;   // helpers:394-395
    rgb_map[3 * (int64_t)ray + 0] = sr + bg;
    rgb_map[3 * (int64_t)ray + 1] = sg + bg;
    rgb_map[3 * (int64_t)ray + 2] = sb + bg;
  }
}

__global__ void __launch_bounds__(256)
raw2outputs_bwd_kernel(const float4* __restrict__ raw, const float* __restrict__ z,
                       const float* __restrict__ rays_d, int ld_d, const float* __restrict__ noise,
                       float noise_scale, int n, int S, int white_bkgd, int detach_all, int det0, int det1,
                       const float* __restrict__ g_rgb, const float* __restrict__ g_disp,
                       const float* __restrict__ g_acc, const float* __restrict__ g_w,
                       const float* __restrict__ g_depth, float4* __restrict__ d_raw) {
  int ray = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  int lane = threadIdx.x % kWarp;
  if (ray >= n) return;
  const bool detach_weights = detach_all || (ray >= det0 && ray < det1);   // helpers:385-388, per ray
  const float* d = rays_d + (int64_t)ray * ld_d;
  float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  int nch = (S + kWarp - 1) / kWarp;
  RaySample smp[kMaxChunks];
  float rgbv[kMaxChunks][3], zv[kMaxChunks], sig[kMaxChunks];
  float carry = 1.0f, sdepth = 0, sacc = 0;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    if (c < nch) {
      int s = c * kWarp + lane;
      bool live = s < S;
      int64_t idx = (int64_t)ray * S + (live ? s : S - 1);
      float4 rw = raw[idx];
      float zi = z[idx];
      float znext = (live && s + 1 < S) ? z[idx + 1] : 0.0f;
      float dist = (s + 1 < S) ? __fsub_rn(znext, zi) : 1e10f;
      dist = __fmul_rn(dist, nrm);
      float sigma = rw.w + (noise ? noise[idx] * noise_scale : 0.0f);
      if (!live) { sigma = 0.0f; dist = 0.0f; }
      smp[c] = composite_step(sigma, dist, carry, lane);
      if (!live) smp[c].w = 0.0f;
      sig[c] = sigma;
      zv[c] = zi;
      rgbv[c][0] = 1.0f / (1.0f + expf(-rw.x));
      rgbv[c][1] = 1.0f / (1.0f + expf(-rw.y));
      rgbv[c][2] = 1.0f / (1.0f + expf(-rw.z));
      sdepth += smp[c].w * zi;
      sacc += smp[c].w;
    }
  }
  sdepth = warp_sum(sdepth); sacc = warp_sum(sacc);
  float gr[3] = {0, 0, 0};
  if (g_rgb) { gr[0] = g_rgb[3 * (int64_t)ray]; gr[1] = g_rgb[3 * (int64_t)ray + 1]; gr[2] = g_rgb[3 * (int64_t)ray + 2]; }
  float gacc = g_acc ? g_acc[ray] : 0.0f;
  if (white_bkgd) gacc -= gr[0] + gr[1] + gr[2];
  float gdepth = g_depth ? g_depth[ray] : 0.0f;
  float gdisp = g_disp ? g_disp[ray] : 0.0f;
  float ratio = sdepth / sacc;
  // disp = 1/max(1e-10, ratio): gradient flows through ratio only where ratio wins the max
  // (skipped entirely when disp carries no gradient: autograd never visits that branch, so an empty ray
  //  (acc == 0, ratio = NaN) must not poison the other outputs' gradients)
  if (gdisp != 0.0f) {
    float gratio = (ratio > 1e-10f) ? -gdisp / (ratio * ratio) : 0.0f;
    if (ratio != ratio) gratio = ratio;   // NaN propagates like autograd
    gdepth += gratio / sacc;
    gacc -= gratio * sdepth / (sacc * sacc);
  }
  // reverse sweep: suffix_i = sum_{k>i} gw_k * w_k, small tail terms accumulated first
  float tail = 0.0f;
#pragma unroll
  for (int c = kMaxChunks - 1; c >= 0; --c) {
    if (c < nch) {
      int s = c * kWarp + lane;
      bool live = s < S;
      int64_t idx = (int64_t)ray * S + (live ? s : S - 1);
      const RaySample& r = smp[c];
      float gw = (g_w ? g_w[idx] : 0.0f) + gdepth * zv[c] + gacc;
      if (!detach_weights) gw += gr[0] * rgbv[c][0] + gr[1] * rgbv[c][1] + gr[2] * rgbv[c][2];
      float gwk = live ? gw * r.w : 0.0f;
      float p = gwk;   // inclusive suffix scan over lanes (towards higher lanes)
#pragma unroll
      for (int o = 1; o < kWarp; o <<= 1) {
        float q = __shfl_down_sync(kFull, p, o);
        if (lane + o < kWarp) p += q;
      }
      float suffix = (p - gwk) + tail;
      tail += __shfl_sync(kFull, p, 0);
      float g_alpha = gw * r.trans - suffix / r.t_in;
      float d_sigma = (sig[c] > 0.0f) ? g_alpha * r.e * r.dist : 0.0f;
      if (live) {
        float4 o4;
        o4.x = gr[0] * r.w * rgbv[c][0] * (1.0f - rgbv[c][0]);
        o4.y = gr[1] * r.w * rgbv[c][1] * (1.0f - rgbv[c][1]);
        o4.z = gr[2] * r.w * rgbv[c][2] * (1.0f - rgbv[c][2]);
        o4.w = d_sigma;
        d_raw[idx] = o4;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// a8 sample_pdf + merge: one warp per ray, cdf in shared memory
// ------------------------------------------------------------------------------------------
constexpr int kMaxBins = 256;
constexpr int kMaxMerge = 512;
constexpr int kPdfWarps = 4;

// cdf[0..nb-1] from weights[0..nb-2] exactly as helpers:306-309 on the CPU oracle:
// fp32 (w+1e-5), fp32 total, fp32 division, prefix sum accumulated in fp64 and rounded per prefix.
__device__ __forceinline__ void build_cdf(const float* __restrict__ w, int nb, float* cdf, int lane) {
  float part = 0.0f;
  for (int i = lane; i < nb - 1; i += kWarp) part += __fadd_rn(w[i], 1e-5f);
  float tot = warp_sum(part);
  for (int i = lane; i < nb - 1; i += kWarp) cdf[i + 1] = __fdiv_rn(__fadd_rn(w[i], 1e-5f), tot);
  __syncwarp();
  if (lane == 0) {
    double acc = 0.0;
    cdf[0] = 0.0f;
    for (int i = 1; i < nb; ++i) { acc += (double)cdf[i]; cdf[i] = (float)acc; }
  }
  __syncwarp();
}

// searchsorted(cdf, u, right=True) (helpers:331) + the lerp of helpers:332-345
__device__ __forceinline__ float invert_cdf(const float* cdf, const float* bins, int nb, float u,
                                            int* ind_out) {
  int lo = 0, hi = nb;   // first index with cdf[idx] > u
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  *ind_out = lo;
  int below = max(0, lo - 1), above = min(nb - 1, lo);
  float c0 = cdf[below], c1 = cdf[above];
  float denom = __fsub_rn(c1, c0);
  if (denom < 1e-5f) denom = 1.0f;
  float t = __fdiv_rn(__fsub_rn(u, c0), denom);
  float b0 = bins[below], b1 = bins[above];
  return __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
}

__global__ void __launch_bounds__(kPdfWarps * kWarp)
sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights,
                  const float* __restrict__ u, int n, int nb, int ns, float* __restrict__ samples,
                  int64_t* __restrict__ inds, float* __restrict__ cdf_out) {
  __shared__ float s_cdf[kPdfWarps][kMaxBins];
  __shared__ float s_bins[kPdfWarps][kMaxBins];
  int wid = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  int ray = blockIdx.x * kPdfWarps + wid;
  if (ray >= n) return;
  float* cdf = s_cdf[wid];
  float* bn = s_bins[wid];
  for (int i = lane; i < nb; i += kWarp) bn[i] = bins[(int64_t)ray * nb + i];
  build_cdf(weights + (int64_t)ray * (nb - 1), nb, cdf, lane);
  if (cdf_out) for (int i = lane; i < nb; i += kWarp) cdf_out[(int64_t)ray * nb + i] = cdf[i];
  for (int j = lane; j < ns; j += kWarp) {
    float uj = u ? u[(int64_t)ray * ns + j] : linspace01(j, ns);
    int ind;
    float v = invert_cdf(cdf, bn, nb, uj, &ind);
    samples[(int64_t)ray * ns + j] = v;
    if (inds) inds[(int64_t)ray * ns + j] = ind;
  }
}

// bitonic sort of `cnt` (power of two, padded with +inf) floats in shared memory by one warp
__device__ __forceinline__ void warp_bitonic(float* v, int cnt, int lane) {
  for (int k = 2; k <= cnt; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < cnt; i += kWarp) {
        int p = i ^ j;
        if (p > i) {
          bool up = (i & k) == 0;
          float a = v[i], b = v[p];
          if ((a > b) == up) { v[i] = b; v[p] = a; }
        }
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(kPdfWarps * kWarp)
merge_sorted_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, int sa, int sb,
                    float* __restrict__ out) {
  __shared__ float s_v[kPdfWarps][kMaxMerge];
  int wid = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  int ray = blockIdx.x * kPdfWarps + wid;
  if (ray >= n) return;
  int tot = sa + sb, cnt = 1;
  while (cnt < tot) cnt <<= 1;
  float* v = s_v[wid];
  for (int i = lane; i < cnt; i += kWarp)
    v[i] = i < sa ? a[(int64_t)ray * sa + i] : (i < tot ? b[(int64_t)ray * sb + (i - sa)] : INFINITY);
  __syncwarp();
  warp_bitonic(v, cnt, lane);
  for (int i = lane; i < tot; i += kWarp) out[(int64_t)ray * tot + i] = v[i];
}

// run_nerf.py:696-702,726 in one launch
__global__ void __launch_bounds__(kPdfWarps * kWarp)
resample_kernel(const float* __restrict__ z, const float* __restrict__ weights,
                const float* __restrict__ u, int n, int S, int n_imp, float* __restrict__ z_out,
                float* __restrict__ z_samples, float* __restrict__ z_std,
                int64_t* __restrict__ inds) {
  __shared__ float s_cdf[kPdfWarps][kMaxBins];
  __shared__ float s_bins[kPdfWarps][kMaxBins];
  __shared__ float s_v[kPdfWarps][kMaxMerge];
  int wid = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  int ray = blockIdx.x * kPdfWarps + wid;
  if (ray >= n) return;
  float* cdf = s_cdf[wid];
  float* bn = s_bins[wid];
  float* v = s_v[wid];
  const float* zr = z + (int64_t)ray * S;
  int nb = S - 1;
  for (int i = lane; i < nb; i += kWarp) bn[i] = __fmul_rn(0.5f, __fadd_rn(zr[i + 1], zr[i]));   // z_vals_mid
  build_cdf(weights + (int64_t)ray * S + 1, nb, cdf, lane);                                      // weights[...,1:-1]
  int tot = S + n_imp, cnt = 1;
  while (cnt < tot) cnt <<= 1;
  float sum = 0.0f;
  for (int j = lane; j < n_imp; j += kWarp) {
    float uj = u ? u[(int64_t)ray * n_imp + j] : linspace01(j, n_imp);
    int ind;
    float smp = invert_cdf(cdf, bn, nb, uj, &ind);
    v[S + j] = smp;
    sum += smp;
    if (z_samples) z_samples[(int64_t)ray * n_imp + j] = smp;
    if (inds) inds[(int64_t)ray * n_imp + j] = ind;
  }
  for (int i = lane; i < S; i += kWarp) v[i] = zr[i];
  for (int i = tot + lane; i < cnt; i += kWarp) v[i] = INFINITY;
  __syncwarp();
  if (z_std) {   // torch.std(unbiased=False) (run_nerf.py:726)
    float mean = warp_sum(sum) / (float)n_imp, var = 0.0f;
    for (int j = lane; j < n_imp; j += kWarp) { float dlt = v[S + j] - mean; var += dlt * dlt; }
    var = warp_sum(var) / (float)n_imp;
    if (lane == 0) z_std[ray] = sqrtf(var);
  }
  warp_bitonic(v, cnt, lane);
  for (int i = lane; i < tot; i += kWarp) z_out[(int64_t)ray * tot + i] = v[i];
}

// ------------------------------------------------------------------------------------------
// a12 flat Adam (torch.optim.Adam single-tensor semantics)
// ------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                            float bc1, float bc2_sqrt, float gscale) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i] * gscale;
  float mi = m[i] + (gi - m[i]) * (1.0f - b1);          // exp_avg.lerp_(grad, 1-beta1)
  float vi = v[i] * b2 + (1.0f - b2) * gi * gi;          // mul_(beta2).addcmul_(g, g, 1-beta2)
  m[i] = mi; v[i] = vi;
  float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = p[i] - (lr / bc1) * (mi / denom);
}

// Batched row-wise searchsorted, the function of the reference's one native component (DS_NeRF/torchsearchsorted/src/cuda/
// searchsorted_cuda_kernel.cu:41-107, dead code on its hot path: run_nerf_helpers.py:10 uses torch.searchsorted).  One thread
// per query; numpy semantics like its unit test demands (test/test_searchsorted.py): left -> first i with a[i] >= v,
// right -> first i with a[i] > v.  Either operand may have a single row that is shared by all rows of the other.
__global__ void searchsorted_kernel(const float* __restrict__ a, const float* __restrict__ v, int64_t* __restrict__ out,
                                    int nrow_a, int nrow_v, int ncol_a, int ncol_v, int side_left) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nrow = nrow_a > nrow_v ? nrow_a : nrow_v;
  if (q >= (int64_t)nrow * ncol_v) return;
  const int row = (int)(q / ncol_v), col = (int)(q % ncol_v);
  const float* ar = a + (int64_t)(nrow_a == 1 ? 0 : row) * ncol_a;
  const float x = v[(int64_t)(nrow_v == 1 ? 0 : row) * ncol_v + col];
  int lo = 0, hi = ncol_a;                       // invariant: answer in [lo, hi]
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const float m = ar[mid];
    if (side_left ? (m < x) : (m <= x)) lo = mid + 1; else hi = mid;
  }
  out[q] = lo;
}

// Adam with the step counter and schedule resident on the device, so a whole train step can be replayed as a CUDA graph:
// state = {step, lr / bias_correction1, sqrt(bias_correction2), lr}; adam_tick advances it once per step.
__global__ void adam_tick_kernel(float* __restrict__ state, float lr0, float decay_base, float decay_steps, float b1,
                                 float b2) {
  const float k = state[0] + 1.0f;                                  // 1-based step about to be applied
  // run_nerf.py:1611-1622: the rate is updated AFTER optimizer.step() from the 0-based global_step, so step k runs at
  // lr0 * base^((k-2)/decay_steps) (steps 1 and 2 both at lr0)
  const float lr = lr0 * powf(decay_base, fmaxf(k - 2.0f, 0.0f) / decay_steps);
  state[0] = k;
  state[1] = lr / (1.0f - powf(b1, k));
  state[2] = sqrtf(1.0f - powf(b2, k));
  state[3] = lr;
}

__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, int64_t n, const float* __restrict__ state, float b1, float b2,
                                float eps, float gscale) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lr_bc1 = state[1], bc2_sqrt = state[2];
  float gi = g[i] * gscale;
  float mi = m[i] + (gi - m[i]) * (1.0f - b1);
  float vi = v[i] * b2 + (1.0f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = p[i] - lr_bc1 * (mi / denom);
}

}  // namespace spn

// ============================================================================================
// C ABI
// ============================================================================================
using namespace spn;

static inline int blocks_for(int64_t work, int threads) { return (int)((work + threads - 1) / threads); }

extern "C" int spn_get_rays(const float* c2w, int H, int W, float focal, int i0, int j0, int h, int w,
                            float* rays_o, float* rays_d, void* stream) {
  SPN_CHECK_ARG(c2w && rays_o && rays_d, "spn_get_rays: null pointer");
  SPN_CHECK_ARG(H > 0 && W > 0 && h > 0 && w > 0 && i0 >= 0 && j0 >= 0 && i0 + h <= H && j0 + w <= W,
                "spn_get_rays: window [%d+%d, %d+%d] outside %dx%d", i0, h, j0, w, H, W);
  get_rays_kernel<<<blocks_for((int64_t)h * w, 256), 256, 0, as_stream(stream)>>>(c2w, H, W, focal, i0, j0, h, w,
                                                                                 rays_o, rays_d);
  SPN_LAUNCH_CHECK("get_rays_kernel");
  return SPN_OK;
}

extern "C" int spn_ndc_rays(int n, int H, int W, float focal, float near_plane, const float* rays_o,
                            const float* rays_d, float* out_o, float* out_d, void* stream) {
  SPN_CHECK_ARG(n >= 0 && rays_o && rays_d && out_o && out_d, "spn_ndc_rays: bad arguments");
  if (n == 0) return SPN_OK;
  ndc_rays_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(n, H, W, focal, near_plane, rays_o, rays_d,
                                                                     out_o, out_d);
  SPN_LAUNCH_CHECK("ndc_rays_kernel");
  return SPN_OK;
}

extern "C" int spn_build_ray_batch(int n, const float* rays_o, const float* rays_d, float near_, float far_,
                                   int ndc, int H, int W, float focal, float* rays, void* stream) {
  SPN_CHECK_ARG(n >= 0 && rays_o && rays_d && rays, "spn_build_ray_batch: bad arguments");
  if (n == 0) return SPN_OK;
  build_ray_batch_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(n, rays_o, rays_d, near_, far_, ndc, H,
                                                                            W, focal, rays);
  SPN_LAUNCH_CHECK("build_ray_batch_kernel");
  return SPN_OK;
}

extern "C" int spn_gather_ray_batch(int n, const float* pool_o, const float* pool_d, const int64_t* idx, float near_,
                                    float far_, int ndc, int H, int W, float focal, float* rays, const float* rgb_pool,
                                    float* rgb_out, int n_rgb, const float* disp_pool, float* disp_out, void* stream) {
  SPN_CHECK_ARG(n >= 0 && pool_o && pool_d && idx && rays && n_rgb >= 0 && n_rgb <= n && (!rgb_pool || rgb_out || n_rgb == 0) &&
                (!disp_pool || disp_out || n_rgb == n), "spn_gather_ray_batch: bad arguments");
  if (n == 0) return SPN_OK;
  gather_ray_batch_kernel<<<blocks_for(n, 128), 128, 0, as_stream(stream)>>>(n, pool_o, pool_d, idx, near_, far_, ndc, H, W,
                                                                             focal, rays, rgb_pool, rgb_out, n_rgb,
                                                                             disp_pool, disp_out);
  SPN_LAUNCH_CHECK("gather_ray_batch_kernel");
  return SPN_OK;
}

extern "C" int spn_embed(const float* x, int64_t m, int n_freqs, float* out, void* stream) {
  SPN_CHECK_ARG(x && out && m >= 0 && n_freqs >= 0 && n_freqs <= 16, "spn_embed: bad arguments");
  if (m == 0) return SPN_OK;
  embed_kernel<<<blocks_for(m * (3 + 6 * n_freqs), 256), 256, 0, as_stream(stream)>>>(x, m, n_freqs, out);
  SPN_LAUNCH_CHECK("embed_kernel");
  return SPN_OK;
}

extern "C" int spn_sample_z(const float* rays, int n, int ncols, int S, int lindisp, const float* t_rand,
                            float* z, void* stream) {
  SPN_CHECK_ARG(rays && z && n >= 0 && ncols >= 8 && S >= 2, "spn_sample_z: bad arguments (n=%d ncols=%d S=%d)", n,
                ncols, S);
  if (n == 0) return SPN_OK;
  sample_z_kernel<<<blocks_for((int64_t)n * S, 256), 256, 0, as_stream(stream)>>>(rays, n, ncols, S, lindisp, t_rand, z);
  SPN_LAUNCH_CHECK("sample_z_kernel");
  return SPN_OK;
}

extern "C" int spn_raw2outputs_fwd(const float* raw, const float* z, const float* rays_d, int ld_d,
                                   const float* noise, int n, int S, int white_bkgd, float* rgb_map,
                                   float* disp_map, float* acc_map, float* weights, float* depth_map,
                                   float* alpha, void* stream) {
  return spn::composite_fwd(raw, z, rays_d, ld_d, noise, 1.0f, n, S, white_bkgd, rgb_map, disp_map, acc_map, weights,
                            depth_map, alpha, as_stream(stream));
}

int spn::composite_fwd(const float* raw, const float* z, const float* rays_d, int ld_d, const float* noise,
                       float noise_scale, int n, int S, int white_bkgd, float* rgb_map, float* disp_map,
                       float* acc_map, float* weights, float* depth_map, float* alpha, cudaStream_t stream) {
  if (n == 0) return SPN_OK;   // empty batches carry null data pointers
  SPN_CHECK_ARG(raw && z && rays_d && rgb_map && disp_map && acc_map && weights && depth_map,
                "spn_raw2outputs_fwd: null pointer");
  SPN_CHECK_ARG(n >= 0 && S >= 1 && ld_d >= 3, "spn_raw2outputs_fwd: bad shape n=%d S=%d", n, S);
  SPN_CHECK_ARG(((uintptr_t)raw & 15) == 0, "spn_raw2outputs_fwd: raw must be 16-byte aligned");
  if (n == 0) return SPN_OK;
  raw2outputs_fwd_kernel<<<blocks_for((int64_t)n * 32, 256), 256, 0, stream>>>(
      (const float4*)raw, z, rays_d, ld_d, noise, noise_scale, n, S, white_bkgd, rgb_map, disp_map, acc_map, weights,
      depth_map, alpha);
  SPN_LAUNCH_CHECK("raw2outputs_fwd_kernel");
  return SPN_OK;
}

extern "C" int spn_raw2outputs_bwd(const float* raw, const float* z, const float* rays_d, int ld_d,
                                   const float* noise, int n, int S, int white_bkgd, int detach_weights,
                                   const float* g_rgb, const float* g_disp, const float* g_acc,
                                   const float* g_weights, const float* g_depth, float* d_raw, void* stream) {
  return spn::composite_bwd(raw, z, rays_d, ld_d, noise, 1.0f, n, S, white_bkgd, detach_weights, 0, 0, g_rgb, g_disp, g_acc,
                            g_weights, g_depth, d_raw, as_stream(stream));
}

int spn::composite_bwd(const float* raw, const float* z, const float* rays_d, int ld_d, const float* noise,
                       float noise_scale, int n, int S, int white_bkgd, int detach_weights, int detach_begin,
                       int detach_end, const float* g_rgb, const float* g_disp, const float* g_acc,
                       const float* g_weights, const float* g_depth, float* d_raw, cudaStream_t stream) {
  if (n == 0) return SPN_OK;
  SPN_CHECK_ARG(raw && z && rays_d && d_raw, "spn_raw2outputs_bwd: null pointer");
  SPN_CHECK_ARG(n >= 0 && S >= 1 && S <= kMaxChunks * 32 && ld_d >= 3,
                "spn_raw2outputs_bwd: S=%d outside [1,%d]", S, kMaxChunks * 32);
  SPN_CHECK_ARG((((uintptr_t)raw | (uintptr_t)d_raw) & 15) == 0, "spn_raw2outputs_bwd: raw/d_raw must be 16-byte aligned");
  if (n == 0) return SPN_OK;
  raw2outputs_bwd_kernel<<<blocks_for((int64_t)n * 32, 256), 256, 0, stream>>>(
      (const float4*)raw, z, rays_d, ld_d, noise, noise_scale, n, S, white_bkgd, detach_weights, detach_begin, detach_end,
      g_rgb, g_disp, g_acc, g_weights, g_depth, (float4*)d_raw);
  SPN_LAUNCH_CHECK("raw2outputs_bwd_kernel");
  return SPN_OK;
}

extern "C" int spn_sample_pdf_cdf(const float* bins, const float* weights, const float* u, int n, int nb, int ns,
                                  float* samples, int64_t* inds, float* cdf_out, void* stream) {
  SPN_CHECK_ARG(bins && weights && samples, "spn_sample_pdf: null pointer");
  SPN_CHECK_ARG(n >= 0 && nb >= 2 && nb <= kMaxBins && ns >= 1, "spn_sample_pdf: nb=%d outside [2,%d]", nb, kMaxBins);
  if (n == 0) return SPN_OK;
  sample_pdf_kernel<<<blocks_for(n, kPdfWarps), kPdfWarps * 32, 0, as_stream(stream)>>>(bins, weights, u, n, nb, ns,
                                                                                      samples, inds, cdf_out);
  SPN_LAUNCH_CHECK("sample_pdf_kernel");
  return SPN_OK;
}

extern "C" int spn_sample_pdf(const float* bins, const float* weights, const float* u, int n, int nb, int ns,
                              float* samples, int64_t* inds, void* stream) {
  return spn_sample_pdf_cdf(bins, weights, u, n, nb, ns, samples, inds, nullptr, stream);
}

extern "C" int spn_searchsorted(const float* a, const float* v, int64_t* out, int nrow_a, int nrow_v, int ncol_a, int ncol_v,
                                int side_left, void* stream) {
  SPN_CHECK_ARG(a && v && out && nrow_a >= 1 && nrow_v >= 1 && ncol_a >= 0 && ncol_v >= 0 &&
                (nrow_a == nrow_v || nrow_a == 1 || nrow_v == 1), "spn_searchsorted: bad shapes (%d x %d, %d x %d)", nrow_a,
                ncol_a, nrow_v, ncol_v);
  const int64_t work = (int64_t)(nrow_a > nrow_v ? nrow_a : nrow_v) * ncol_v;
  if (work == 0) return SPN_OK;
  searchsorted_kernel<<<blocks_for(work, 256), 256, 0, as_stream(stream)>>>(a, v, out, nrow_a, nrow_v, ncol_a, ncol_v, side_left);
  SPN_LAUNCH_CHECK("searchsorted_kernel");
  return SPN_OK;
}

extern "C" int spn_merge_sorted(const float* a, const float* b, int n, int sa, int sb, float* out, void* stream) {
  SPN_CHECK_ARG(a && b && out && n >= 0 && sa >= 0 && sb >= 0 && sa + sb >= 1 && sa + sb <= kMaxMerge,
                "spn_merge_sorted: sa+sb=%d outside [1,%d]", sa + sb, kMaxMerge);
  if (n == 0) return SPN_OK;
  merge_sorted_kernel<<<blocks_for(n, kPdfWarps), kPdfWarps * 32, 0, as_stream(stream)>>>(a, b, n, sa, sb, out);
  SPN_LAUNCH_CHECK("merge_sorted_kernel");
  return SPN_OK;
}

extern "C" int spn_resample(const float* z, const float* weights, const float* u, int n, int S, int n_imp,
                            float* z_out, float* z_samples, float* z_std, int64_t* inds, void* stream) {
  SPN_CHECK_ARG(z && weights && z_out, "spn_resample: null pointer");
  SPN_CHECK_ARG(n >= 0 && S >= 3 && S - 1 <= kMaxBins && n_imp >= 1 && S + n_imp <= kMaxMerge,
                "spn_resample: S=%d n_imp=%d unsupported", S, n_imp);
  if (n == 0) return SPN_OK;
  resample_kernel<<<blocks_for(n, kPdfWarps), kPdfWarps * 32, 0, as_stream(stream)>>>(z, weights, u, n, S, n_imp,
                                                                                    z_out, z_samples, z_std, inds);
  SPN_LAUNCH_CHECK("resample_kernel");
  return SPN_OK;
}


// ---------------------------------------------------------------------------------------------
// train-step losses (run_nerf.py:1481-1521, default flags) for a batch that concatenates the step's three ray
// groups: [0,n1) unmasked rays and [n1,n1+n2) masked rays of the kept view -> img2mse(rgb, target) + img2mse(rgb0,
// target) per group; [n1+n2, n) inpainted-disparity rays -> img2mse(disp, target) + img2mse(disp0, target), dropped
// (loss and gradient) when NaN (`if not inp_loss.isnan()`, :1520).  Two launches: sums, then gradients + scalars.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
train_loss_sums_kernel(const float* __restrict__ rgb, const float* __restrict__ rgb0, const float* __restrict__ disp,
                       const float* __restrict__ disp0, const float* __restrict__ tgt_rgb,
                       const float* __restrict__ tgt_disp, int n1, int n2, int n3, float* __restrict__ sums) {
  // element space: 3*(n1+n2) colour entries, then n3 disparity entries
  const int ncol = 3 * (n1 + n2), total = ncol + n3;
  float s[6] = {0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < ncol) {
      const float t = tgt_rgb[i], a = rgb[i] - t, b = rgb0[i] - t;
      const int k = i < 3 * n1 ? 0 : 2;
      s[k] += a * a; s[k + 1] += b * b;
    } else {
      const int r = i - ncol;
      const float t = tgt_disp[r], a = disp[n1 + n2 + r] - t, b = disp0[n1 + n2 + r] - t;
      s[4] += a * a; s[5] += b * b;
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float v = warp_sum(s[k]);
    if ((threadIdx.x & 31) == 0 && v != 0.0f) atomicAdd(sums + k, v);
    if ((threadIdx.x & 31) == 0 && v != v) atomicAdd(sums + k, v);   // NaN must reach the sum
  }
}

__global__ void __launch_bounds__(256)
train_loss_grads_kernel(const float* __restrict__ rgb, const float* __restrict__ rgb0, const float* __restrict__ disp,
                        const float* __restrict__ disp0, const float* __restrict__ tgt_rgb,
                        const float* __restrict__ tgt_disp, int n1, int n2, int n3, const float* __restrict__ sums,
                        float* __restrict__ g_rgb, float* __restrict__ g_rgb0, float* __restrict__ g_disp,
                        float* __restrict__ g_disp0, float* __restrict__ out) {
  const int n = n1 + n2 + n3, ncol = 3 * n;
  const float ld = n3 > 0 ? sums[4] / n3 : 0.0f, ld0 = n3 > 0 ? sums[5] / n3 : 0.0f;
  const bool bad = (ld + ld0) != (ld + ld0);
  const float c1 = n1 > 0 ? 2.0f / (3.0f * n1) : 0.0f, c2 = n2 > 0 ? 2.0f / (3.0f * n2) : 0.0f;
  const float c3 = (n3 > 0 && !bad) ? 2.0f / n3 : 0.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ncol + n; i += gridDim.x * blockDim.x) {
    if (i < ncol) {
      float a = 0.0f, b = 0.0f;
      if (i < 3 * (n1 + n2)) {
        const float t = tgt_rgb[i], c = i < 3 * n1 ? c1 : c2;
        a = c * (rgb[i] - t); b = c * (rgb0[i] - t);
      }
      g_rgb[i] = a; g_rgb0[i] = b;
    } else {
      const int r = i - ncol;
      float a = 0.0f, b = 0.0f;
      if (r >= n1 + n2 && c3 != 0.0f) {
        const float t = tgt_disp[r - n1 - n2];
        a = c3 * (disp[r] - t); b = c3 * (disp0[r] - t);
      }
      g_disp[r] = a; g_disp0[r] = b;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const float l0 = n1 > 0 ? sums[0] / (3.0f * n1) : 0.0f, l1 = n1 > 0 ? sums[1] / (3.0f * n1) : 0.0f;
    const float l2 = n2 > 0 ? sums[2] / (3.0f * n2) : 0.0f, l3 = n2 > 0 ? sums[3] / (3.0f * n2) : 0.0f;
    out[0] = l0 + l1 + l2 + l3 + (bad ? 0.0f : ld + ld0);   // the step's loss
    out[1] = -10.0f * log10f(l0);                           // psnr of the unmasked fine rgb (mse2psnr)
    out[2] = l0; out[3] = l1; out[4] = l2; out[5] = l3; out[6] = bad ? 0.0f : ld; out[7] = bad ? 0.0f : ld0;
  }
}

extern "C" int spn_train_losses(const float* rgb_map, const float* rgb0, const float* disp_map, const float* disp0,
                                const float* target_rgb, const float* target_disp, int n1, int n2, int n3,
                                float* sums6_zeroed, float* g_rgb, float* g_rgb0, float* g_disp, float* g_disp0,
                                float* out8, void* stream) {
  SPN_CHECK_ARG(rgb_map && rgb0 && disp_map && disp0 && sums6_zeroed && g_rgb && g_rgb0 && g_disp && g_disp0 && out8 &&
                n1 >= 0 && n2 >= 0 && n3 >= 0 && (n1 + n2 == 0 || target_rgb) && (n3 == 0 || target_disp),
                "spn_train_losses: bad arguments");
  const int n = n1 + n2 + n3;
  if (n == 0) return SPN_OK;
  cudaStream_t st = as_stream(stream);
  const int blocks = blocks_for((int64_t)4 * n, 256) < 296 ? blocks_for((int64_t)4 * n, 256) : 296;
  train_loss_sums_kernel<<<blocks, 256, 0, st>>>(rgb_map, rgb0, disp_map, disp0, target_rgb, target_disp, n1, n2, n3, sums6_zeroed);
  SPN_LAUNCH_CHECK("train_loss_sums_kernel");
  train_loss_grads_kernel<<<blocks, 256, 0, st>>>(rgb_map, rgb0, disp_map, disp0, target_rgb, target_disp, n1, n2, n3,
                                                 sums6_zeroed, g_rgb, g_rgb0, g_disp, g_disp0, out8);
  SPN_LAUNCH_CHECK("train_loss_grads_kernel");
  return SPN_OK;
}

extern "C" int spn_adam_tick(float* state4, float lr0, float decay_base, float decay_steps, float beta1, float beta2,
                             void* stream) {
  SPN_CHECK_ARG(state4 && decay_steps > 0.f, "spn_adam_tick: bad arguments");
  adam_tick_kernel<<<1, 1, 0, as_stream(stream)>>>(state4, lr0, decay_base, decay_steps, beta1, beta2);
  SPN_LAUNCH_CHECK("adam_tick_kernel");
  return SPN_OK;
}

extern "C" int spn_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                 const float* state4, float beta1, float beta2, float eps, float grad_scale,
                                 void* stream) {
  SPN_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && state4 && n >= 0, "spn_adam_step_dev: bad arguments");
  if (n == 0) return SPN_OK;
  adam_dev_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, state4, beta1,
                                                                     beta2, eps, grad_scale);
  SPN_LAUNCH_CHECK("adam_dev_kernel");
  return SPN_OK;
}

extern "C" int spn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                             float lr, float beta1, float beta2, float eps, int step, float grad_scale,
                             void* stream) {
  SPN_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "spn_adam_step: bad arguments");
  if (n == 0) return SPN_OK;
  float bc1 = 1.0f - powf(beta1, (float)step);
  float bc2s = sqrtf(1.0f - powf(beta2, (float)step));
  adam_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1,
                                                                 beta2, eps, bc1, bc2s, grad_scale);
  SPN_LAUNCH_CHECK("adam_kernel");
  return SPN_OK;
}
