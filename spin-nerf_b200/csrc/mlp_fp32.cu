// NeRF MLP (DS_NeRF/run_nerf_helpers.py:74-127) in fp32 on CUDA cores — the tight-parity mode
// (SPN_PREC_FP32).  Activations are materialised in HBM; every linear layer is one launch of a
// strided SIMT GEMM.  This is the reference-faithful arithmetic (true fp32 products and sums,
// like the reference's TF32-off cuBLAS SGEMMs), NOT the fast path: the fast path is mlp_tc.cu.
#include "common.cuh"

namespace spn {

// C[m,n] (op)= sum_k A(m,k) * B(k,n)   with arbitrary element strides.
//   epilogue: (+ C if accumulate) (+ bias[n]) then relu, or multiply by (mask[m,n] > 0).
//   gridDim.z > 1 splits K and atomically adds raw partial sums (used for weight gradients).
struct GemmArgs {
  const float* A; int64_t sAm, sAk;
  const float* B; int64_t sBk, sBn;
  float* C; int64_t sCm, sCn;
  const float* bias;
  const float* mask; int64_t sMm;   // mask row stride (column stride 1)
  int64_t M; int N; int64_t K;
  int accumulate, relu;
};

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmArgs a) {
  __shared__ float As[TK][TM + 1];
  __shared__ float Bs[TK][TN + 1];
  int tid = threadIdx.x;
  int tx = tid % 16, ty = tid / 16;
  int64_t m0 = (int64_t)blockIdx.x * TM;
  int n0 = blockIdx.y * TN;
  int64_t kchunk = (a.K + gridDim.z - 1) / gridDim.z;
  kchunk = (kchunk + TK - 1) / TK * TK;
  int64_t kbeg = (int64_t)blockIdx.z * kchunk;
  int64_t kend = kbeg + kchunk < a.K ? kbeg + kchunk : a.K;
  float acc[4][4] = {};
  for (int64_t k0 = kbeg; k0 < kend; k0 += TK) {
    // pick the load order that walks the unit-stride dimension with consecutive threads
    for (int i = tid; i < TM * TK; i += 256) {
      int mm, kk;
      if (a.sAk == 1) { kk = i % TK; mm = i / TK; } else { mm = i % TM; kk = i / TM; }
      int64_t gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < a.M && gk < kend) ? a.A[gm * a.sAm + gk * a.sAk] : 0.0f;
    }
    for (int i = tid; i < TN * TK; i += 256) {
      int nn, kk;
      if (a.sBk == 1) { kk = i % TK; nn = i / TK; } else { nn = i % TN; kk = i / TN; }
      int gn = n0 + nn; int64_t gk = k0 + kk;
      Bs[kk][nn] = (gn < a.N && gk < kend) ? a.B[gk * a.sBk + (int64_t)gn * a.sBn] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t gm = m0 + ty * 4 + i;
    if (gm >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= a.N) continue;
      float* c = a.C + gm * a.sCm + (int64_t)gn * a.sCn;
      float v = acc[i][j];
      if (gridDim.z > 1) { atomicAdd(c, v); continue; }
      if (a.accumulate) v += *c;
      if (a.bias) v += a.bias[gn];
      if (a.relu) v = fmaxf(v, 0.0f);
      if (a.mask) v = a.mask[gm * a.sMm + gn] > 0.0f ? v : 0.0f;
      *c = v;
    }
  }
}

static int gemm(cudaStream_t st, const float* A, int64_t sAm, int64_t sAk, const float* B, int64_t sBk, int64_t sBn,
                float* C, int64_t sCm, int64_t sCn, int64_t M, int N, int64_t K, const float* bias, int relu,
                int accumulate, const float* mask, int64_t sMm, int splitk) {
  GemmArgs a{A, sAm, sAk, B, sBk, sBn, C, sCm, sCn, bias, mask, sMm, M, N, K, accumulate, relu};
  dim3 grid((unsigned)((M + TM - 1) / TM), (unsigned)((N + TN - 1) / TN), (unsigned)splitk);
  gemm_f32_kernel<<<grid, 256, 0, st>>>(a);
  SPN_LAUNCH_CHECK("gemm_f32_kernel");
  return SPN_OK;
}

// Y[m, n] = act(X[m,k] W[n,k]^T + b)    (nn.Linear)
static int linear(cudaStream_t st, const float* X, int64_t ldx, const float* W, int64_t ldw, const float* b, float* Y,
                  int64_t ldy, int64_t M, int N, int K, int relu, int accumulate) {
  return gemm(st, X, ldx, 1, W, 1, ldw, Y, ldy, 1, M, N, K, b, relu, accumulate, nullptr, 0, 1);
}
// dX[m,k] = (dY[m,n] W[n,k]) (* mask)
static int dgrad(cudaStream_t st, const float* dY, int64_t ldy, const float* W, int64_t ldw, float* dX, int64_t ldx,
                 int64_t M, int N, int K, int accumulate, const float* mask, int64_t ldm) {
  return gemm(st, dY, ldy, 1, W, ldw, 1, dX, ldx, 1, M, K, N, nullptr, 0, accumulate, mask, ldm, 1);
}
// dW[n,k] += sum_m dY[m,n] X[m,k]   (split over samples, atomic accumulate)
static int wgrad(cudaStream_t st, const float* dY, int64_t ldy, const float* X, int64_t ldx, float* dW, int64_t ldw,
                 int64_t M, int N, int K) {
  int tiles = ((N + TM - 1) / TM) * ((K + TN - 1) / TN);
  int64_t want = (2 * (int64_t)sm_count() + tiles - 1) / tiles;
  int64_t maxsplit = (M + 4 * TK - 1) / (4 * TK);
  int splitk = (int)(want < maxsplit ? want : maxsplit);
  if (splitk < 2) splitk = 2;   // always take the atomic path: dW accumulates into existing grads
  return gemm(st, dY, 1, ldy, X, ldx, 1, dW, ldw, 1, N, K, M, nullptr, 0, 0, nullptr, 0, splitk);
}

__global__ void colsum_kernel(const float* __restrict__ dY, int64_t ldy, int64_t M, int N, float* __restrict__ db) {
  // grid: (ceil(N/32), chunks of rows); block 32x8
  int n = blockIdx.x * 32 + threadIdx.x;
  int64_t rows_per = (M + gridDim.y - 1) / gridDim.y;
  int64_t r0 = (int64_t)blockIdx.y * rows_per, r1 = r0 + rows_per < M ? r0 + rows_per : M;
  float s = 0.0f;
  if (n < N)
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) s += dY[r * ldy + n];
  __shared__ float red[8][33];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    for (int i = 1; i < 8; ++i) s += red[i][threadIdx.x];
    atomicAdd(db + n, s);
  }
}
static int colsum(cudaStream_t st, const float* dY, int64_t ldy, int64_t M, int N, float* db) {
  int chunks = (int)((M + 2047) / 2048);
  if (chunks < 1) chunks = 1;
  if (chunks > 512) chunks = 512;
  colsum_kernel<<<dim3((N + 31) / 32, chunks), dim3(32, 8), 0, st>>>(dY, ldy, M, N, db);
  SPN_LAUNCH_CHECK("colsum_kernel");
  return SPN_OK;
}

// gamma(pt) [m,63] and gamma(dir) [m,27] (helpers:22-70), straight into the stash
__global__ void encode_kernel(SampleSource src, int64_t m, float* __restrict__ xp, float* __restrict__ xd) {
  int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= m) return;
  float pt[3], dir[3];
  fetch_sample(src, row, pt, dir);
  float* p = xp + row * kEncP;
  float* d = xd + row * kEncD;
  for (int c = 0; c < 3; ++c) { p[c] = pt[c]; d[c] = dir[c]; }
  for (int k = 0; k < 10; ++k)
    for (int c = 0; c < 3; ++c) {
      float a = __fmul_rn(pt[c], (float)(1 << k));
      p[3 + 6 * k + c] = sinf(a);
      p[3 + 6 * k + 3 + c] = cosf(a);
    }
  for (int k = 0; k < 4; ++k)
    for (int c = 0; c < 3; ++c) {
      float a = __fmul_rn(dir[c], (float)(1 << k));
      d[3 + 6 * k + c] = sinf(a);
      d[3 + 6 * k + 3 + c] = cosf(a);
    }
}

// stash layout (floats): XP[m,63] XD[m,27] H0..H7[m,256] FEAT[m,256] HV[m,128]
struct Fp32Stash {
  float *xp, *xd, *h[8], *feat, *hv;
};
static Fp32Stash carve(void* base, int64_t m) {
  Fp32Stash s;
  float* p = (float*)base;
  s.xp = p; p += m * kEncP;
  s.xd = p; p += m * kEncD;
  for (int i = 0; i < 8; ++i) { s.h[i] = p; p += m * kW; }
  s.feat = p; p += m * kW;
  s.hv = p; p += m * kWV;
  return s;
}
size_t mlp_fp32_stash_bytes(int64_t m) { return (size_t)m * (kEncP + kEncD + 9 * kW + kWV) * sizeof(float); }
size_t mlp_fp32_bwd_ws_bytes(int64_t m) { return (size_t)m * (2 * kW + kWV) * sizeof(float); }

#define RET_IF(x) do { int rc__ = (x); if (rc__ != SPN_OK) return rc__; } while (0)

int mlp_fp32_fwd(const float* P, const SampleSource& src, int64_t m, float* raw, void* stash, cudaStream_t st) {
  SPN_CHECK_ARG(P && raw && stash, "mlp_fp32_fwd: fp32 mode needs params and a stash/workspace buffer");
  ParamOffsets po = param_offsets();
  Fp32Stash s = carve(stash, m);
  encode_kernel<<<(unsigned)((m + 127) / 128), 128, 0, st>>>(src, m, s.xp, s.xd);
  SPN_LAUNCH_CHECK("encode_kernel");
  const float* h = s.xp;
  int k = kEncP;
  for (int i = 0; i < 8; ++i) {
    const float* W = P + po.off[2 * i];
    const float* b = P + po.off[2 * i + 1];
    if (i == 5) {   // [input_pts, h4] @ W5^T  (skip concat, helpers:110-111)
      int ldw = kEncP + kW;
      RET_IF(linear(st, s.xp, kEncP, W, ldw, nullptr, s.h[i], kW, m, kW, kEncP, 0, 0));
      RET_IF(linear(st, h, kW, W + kEncP, ldw, b, s.h[i], kW, m, kW, kW, 1, 1));
    } else {
      RET_IF(linear(st, h, k, W, k, b, s.h[i], kW, m, kW, k, 1, 0));
    }
    h = s.h[i];
    k = kW;
  }
  // alpha / feature / views / rgb heads (helpers:113-123); raw = [rgb(3), alpha]
  RET_IF(linear(st, h, kW, P + po.off[T_WA], kW, P + po.off[T_BA], raw + 3, 4, m, 1, kW, 0, 0));
  RET_IF(linear(st, h, kW, P + po.off[T_WF], kW, P + po.off[T_BF], s.feat, kW, m, kW, kW, 0, 0));
  int ldv = kW + kEncD;
  RET_IF(linear(st, s.feat, kW, P + po.off[T_WV], ldv, nullptr, s.hv, kWV, m, kWV, kW, 0, 0));
  RET_IF(linear(st, s.xd, kEncD, P + po.off[T_WV] + kW, ldv, P + po.off[T_BV], s.hv, kWV, m, kWV, kEncD, 1, 1));
  RET_IF(linear(st, s.hv, kWV, P + po.off[T_WR], kWV, P + po.off[T_BR], raw, 4, m, 3, kWV, 0, 0));
  return SPN_OK;
}

int mlp_fp32_bwd(const float* P, const void* stash, const float* d_raw, int64_t m, float* G, void* ws,
                 cudaStream_t st) {
  SPN_CHECK_ARG(P && stash && d_raw && G && ws, "mlp_fp32_bwd: null pointer");
  ParamOffsets po = param_offsets();
  Fp32Stash s = carve(const_cast<void*>(stash), m);
  float* dA = (float*)ws;
  float* dB = dA + m * kW;
  float* dHV = dB + m * kW;
  const float* d_rgb = d_raw;       // ld 4
  const float* d_alpha = d_raw + 3; // ld 4
  int ldv = kW + kEncD;
  // rgb head
  RET_IF(wgrad(st, d_rgb, 4, s.hv, kWV, G + po.off[T_WR], kWV, m, 3, kWV));
  RET_IF(colsum(st, d_rgb, 4, m, 3, G + po.off[T_BR]));
  RET_IF(dgrad(st, d_rgb, 4, P + po.off[T_WR], kWV, dHV, kWV, m, 3, kWV, 0, s.hv, kWV));
  // views layer
  RET_IF(wgrad(st, dHV, kWV, s.feat, kW, G + po.off[T_WV], ldv, m, kWV, kW));
  RET_IF(wgrad(st, dHV, kWV, s.xd, kEncD, G + po.off[T_WV] + kW, ldv, m, kWV, kEncD));
  RET_IF(colsum(st, dHV, kWV, m, kWV, G + po.off[T_BV]));
  RET_IF(dgrad(st, dHV, kWV, P + po.off[T_WV], ldv, dA, kW, m, kWV, kW, 0, nullptr, 0));   // d_feat
  // feature + alpha heads
  RET_IF(wgrad(st, dA, kW, s.h[7], kW, G + po.off[T_WF], kW, m, kW, kW));
  RET_IF(colsum(st, dA, kW, m, kW, G + po.off[T_BF]));
  RET_IF(wgrad(st, d_alpha, 4, s.h[7], kW, G + po.off[T_WA], kW, m, 1, kW));
  RET_IF(colsum(st, d_alpha, 4, m, 1, G + po.off[T_BA]));
  RET_IF(dgrad(st, dA, kW, P + po.off[T_WF], kW, dB, kW, m, kW, kW, 0, nullptr, 0));
  RET_IF(dgrad(st, d_alpha, 4, P + po.off[T_WA], kW, dB, kW, m, 1, kW, 1, s.h[7], kW));   // += then mask(h7>0)
  float* dcur = dB;
  float* dnext = dA;
  for (int i = 7; i >= 0; --i) {
    const float* W = P + po.off[2 * i];
    float* gW = G + po.off[2 * i];
    RET_IF(colsum(st, dcur, kW, m, kW, G + po.off[2 * i + 1]));
    if (i == 0) {
      RET_IF(wgrad(st, dcur, kW, s.xp, kEncP, gW, kEncP, m, kW, kEncP));
    } else if (i == 5) {
      int ldw = kEncP + kW;
      RET_IF(wgrad(st, dcur, kW, s.xp, kEncP, gW, ldw, m, kW, kEncP));
      RET_IF(wgrad(st, dcur, kW, s.h[4], kW, gW + kEncP, ldw, m, kW, kW));
      RET_IF(dgrad(st, dcur, kW, W + kEncP, ldw, dnext, kW, m, kW, kW, 0, s.h[4], kW));
    } else {
      RET_IF(wgrad(st, dcur, kW, s.h[i - 1], kW, gW, kW, m, kW, kW));
      RET_IF(dgrad(st, dcur, kW, W, kW, dnext, kW, m, kW, kW, 0, s.h[i - 1], kW));
    }
    float* t = dcur; dcur = dnext; dnext = t;
  }
  return SPN_OK;
}

}  // namespace spn
