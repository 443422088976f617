// NeRF MLP (DS_NeRF/run_nerf_helpers.py:74-127) fused with sampling-point generation and positional
// encoding (run_nerf.py:56-71, 670; helpers:22-70) on the 5th-generation tensor cores (tcgen05).
//
// One persistent CTA per SM, 320 threads:
//   warp 0      weight producer: streams pre-swizzled bf16 weight chunks (32 KB = [256 out x 64 in]) from
//               L2/HBM into a 3-stage shared-memory ring with cp.async.bulk (TMA engine) + mbarrier tx counts
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma (M=128, N=256|128, K=16), accumulators in TMEM
//   warps 2-5   epilogue / prologue of tile A (rows 0..127 -> TMEM lanes 0..127)
//   warps 6-9   epilogue / prologue of tile B
// Two 128-sample tiles are in flight per CTA and ping-pong on the tensor pipe: while tile A's epilogue turns
// its fp32 accumulator (TMEM, 256 columns) into the next layer's bf16 A-operand (bias, ReLU, cast, 128B-swizzled
// store to shared memory, in place), tile B's layer runs on the tensor cores, and vice versa.  Activations
// never leave the SM; HBM sees 24 B in + 16 B out per sample (plus the bf16 stash in training).
//
// The 63-wide skip input of layer 5 and the 27-wide view encoding of the views layer are applied as a
// second accumulating pass (K=64 / K=32) after the 256-wide pass, so the 128x256 activation tile can be
// updated in place; the encodings stay packed in registers in between.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "mlp_tc.cuh"

namespace spn {
using namespace tc;

size_t mlp_tc_packed_bytes() { return kPackedBytes; }

// bytes per bulk copy of the weight producers (SPN_W_PIECE = 4096 | 8192 | 16384 for timing experiments)
// diagnostic timeline buffer (device pointer, kTraceSlots int64): see tools/trace_fwd.py
static long long* g_trace = nullptr;
void tc_set_trace(long long* dev) { g_trace = dev; }
constexpr int kTraceRounds = 3, kTraceEvents = 16;
constexpr int kTraceSlots = kTraceRounds * 12 * 2 * kTraceEvents;
__device__ __forceinline__ void trace_stamp(long long* tr, int it, int s, int t, int e) {
  if (tr && blockIdx.x == 0 && it < kTraceRounds) tr[((it * 12 + s) * 2 + t) * kTraceEvents + e] = clock64();
}

int weight_piece_bytes() {
  static int v = 0;
  if (!v) {
    const char* e = getenv("SPN_W_PIECE");
    v = e ? atoi(e) : 16384;
    if (v != 4096 && v != 8192 && v != 16384) v = 16384;
  }
  return v;
}

struct ChunkDesc {
  int src_off;   // float offset of the tensor inside the flat parameter vector (+ n0 for transposed chunks)
  int ld;        // row stride of the source matrix
  int trans;     // 0: val(n,k) = W[n*ld + k0+k]   1: val(n,k) = W[(k0+k)*ld + n]
  int nrows;     // N rows of the chunk (256 | 128)
  int k0;
  int kvalid;    // columns >= kvalid are zero padding
  int dst_off;   // byte offset inside the packed image
};
struct PackTable {
  ChunkDesc c[kFwdChunks + kBwdChunks];
};

static PackTable build_pack_table() {
  PackTable t;
  ParamOffsets po = param_offsets();
  int n = 0, dst = 0;
  auto add = [&](int src, int ld, int trans, int nrows, int k0, int kvalid) {
    t.c[n++] = ChunkDesc{src, ld, trans, nrows, k0, kvalid, dst};
    dst += nrows * 128;
  };
  auto W = [&](int i) { return (int)po.off[2 * i]; };
  // ---- forward, in consumption order
  add(W(0), kEncP, 0, 256, 0, kEncP);
  for (int i = 1; i <= 4; ++i) for (int k = 0; k < 4; ++k) add(W(i), kW, 0, 256, 64 * k, 64);
  for (int k = 0; k < 4; ++k) add(W(5), kW + kEncP, 0, 256, kEncP + 64 * k, 64);
  add(W(5), kW + kEncP, 0, 256, 0, kEncP);
  for (int i = 6; i <= 7; ++i) for (int k = 0; k < 4; ++k) add(W(i), kW, 0, 256, 64 * k, 64);
  for (int k = 0; k < 4; ++k) add((int)po.off[T_WF], kW, 0, 256, 64 * k, 64);
  for (int k = 0; k < 4; ++k) add((int)po.off[T_WV], kW + kEncD, 0, 128, 64 * k, 64);
  add((int)po.off[T_WV], kW + kEncD, 0, 128, kW, kEncD);
  // ---- backward (dgrad): B[n = input feature][k = output feature] = W[k][n]
  for (int k = 0; k < 2; ++k) add((int)po.off[T_WV], kW + kEncD, 1, 256, 64 * k, 64);
  for (int k = 0; k < 4; ++k) add((int)po.off[T_WF], kW, 1, 256, 64 * k, 64);
  for (int i = 7; i >= 1; --i)
    for (int k = 0; k < 4; ++k) add(W(i) + (i == 5 ? kEncP : 0), i == 5 ? kW + kEncP : kW, 1, 256, 64 * k, 64);
  return t;
}

__global__ void pack_kernel(const float* __restrict__ P, uint8_t* __restrict__ out, PackTable tab, int total16) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;   // one 16-byte (8 x bf16) piece per thread
  if (idx >= total16) return;
  // locate the chunk: pieces per chunk = nrows*8
  int c = 0, base = 0;
  while (true) {
    int pieces = tab.c[c].nrows * 8;
    if (idx < base + pieces) break;
    base += pieces; ++c;
  }
  const ChunkDesc d = tab.c[c];
  int local = idx - base;
  int n = local >> 3, j = local & 7;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    int k = j * 8 + e;
    float x = 0.0f;
    if (k < d.kvalid) x = d.trans ? P[d.src_off + (int64_t)(d.k0 + k) * d.ld + n] : P[d.src_off + (int64_t)n * d.ld + d.k0 + k];
    v[e] = x;
  }
  uint4 q;
  q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]); q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(out + d.dst_off + sw128_off(n, j)) = q;
}

__global__ void pack_consts_kernel(const float* __restrict__ P, float* __restrict__ cst, ParamOffsets po) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kConstFloats) return;
  float v = 0.0f;
  if (i < C_BF) v = P[po.off[2 * (i / 256) + 1] + (i % 256)];
  else if (i < C_BV) v = P[po.off[T_BF] + (i - C_BF)];
  else if (i < C_WA) v = P[po.off[T_BV] + (i - C_BV)];
  else if (i < C_BA) v = P[po.off[T_WA] + (i - C_WA)];
  else if (i == C_BA) v = P[po.off[T_BA]];
  else if (i >= C_WR && i < C_BR) v = P[po.off[T_WR] + (i - C_WR)];
  else if (i >= C_BR && i < C_BR + 3) v = P[po.off[T_BR] + (i - C_BR)];
  cst[i] = v;
}

int mlp_tc_pack(const float* params, void* packed, cudaStream_t st) {
  static const PackTable tab = build_pack_table();
  SPN_CHECK_ARG(((uintptr_t)packed & 15) == 0, "spn_mlp_pack_weights: packed image must be 16-byte aligned");
  int total16 = (int)((kFwdBytes + kBwdBytes) / 16);
  pack_kernel<<<(total16 + 255) / 256, 256, 0, st>>>(params, (uint8_t*)packed, tab, total16);
  SPN_LAUNCH_CHECK("pack_kernel");
  pack_consts_kernel<<<(kConstFloats + 255) / 256, 256, 0, st>>>(
      params, (float*)((uint8_t*)packed + kFwdBytes + kBwdBytes), param_offsets());
  SPN_LAUNCH_CHECK("pack_consts_kernel");
  return SPN_OK;
}

// ---- kernel geometry: see mlp_tc.cuh ---------------------------------------------------------------
constexpr int kNumSteps = 12;

// step tables (forward).  A step = one accumulation pass on the tensor cores followed by an epilogue action.
enum EpiAction : int { EPI_RELU = 0, EPI_WRITE_ENC = 1, EPI_LINEAR = 2, EPI_WRITE_DENC = 3, EPI_FINAL = 4, EPI_RELU_ALPHA = 5 };
__constant__ int c_step_chunks[kNumSteps] = {1, 4, 4, 4, 4, 4, 1, 4, 4, 4, 4, 1};
__constant__ int c_step_n[kNumSteps] = {256, 256, 256, 256, 256, 256, 256, 256, 256, 256, 128, 128};
__constant__ int c_step_acc[kNumSteps] = {0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
__constant__ int c_step_ksteps[kNumSteps] = {4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 2};
__constant__ int c_step_epi[kNumSteps] = {EPI_RELU, EPI_RELU, EPI_RELU, EPI_RELU, EPI_RELU, EPI_WRITE_ENC, EPI_RELU,
                                          EPI_RELU, EPI_RELU_ALPHA, EPI_LINEAR, EPI_WRITE_DENC, EPI_FINAL};
__constant__ int c_step_bias[kNumSteps] = {C_B + 0, C_B + 256, C_B + 512, C_B + 768, C_B + 1024, 0, C_B + 1280,
                                           C_B + 1536, C_B + 1792, C_BF, 0, C_BV};
// stash slot written after the step's epilogue (-1: none).  Slots are 16 KB atoms inside the per-tile stash.
__constant__ int c_step_stash_atom[kNumSteps] = {1, 5, 9, 13, 17, -1, 21, 25, 29, 33, -1, 37};
__constant__ int c_step_mask_slot[kNumSteps] = {0, 1, 2, 3, 4, -1, 5, 6, 7, -1, -1, 8};

size_t mlp_tc_stash_bytes(int64_t m) {
  int64_t tiles = (m + 2 * kTileM - 1) / (2 * kTileM) * 2;   // tiles are processed in pairs
  return (size_t)tiles * kStashTileBytes + 256;
}

struct FwdParams {
  const uint8_t* packed;
  SampleSource src;
  int64_t m;
  float* raw;
  uint8_t* stash;   // nullable
  int num_pairs;
  int piece;        // bytes per cp.async.bulk of the weight producer (a chunk is split into pieces)
  long long* trace; // diagnostic (spn_tc_set_trace): clock64 stamps of CTA 0's pipeline events, NULL = off
};

// sin/cos for the bf16 encodings: two-constant Cody-Waite reduction to [-pi, pi] (exact product via fma) followed by
// the SFU approximations.  |abs error| < 1e-6 for |x| < 1e4, three orders below the bf16 rounding (4e-3) that follows;
// no local-memory slow path like sincosf's Payne-Hanek branch.  (The fp32 mode uses the accurate sinf/cosf.)
__device__ __forceinline__ void fast_sincos(float x, float& s, float& c) {
  const float k = rintf(x * 0.15915494309189535f);
  float r = fmaf(-k, 6.2831854820251465f, x);
  r = fmaf(k, 1.7484556e-7f, r);
  s = __sinf(r);
  c = __cosf(r);
}

// packs [v, sin(2^k v), cos(2^k v)]_k (helpers:28-52 order) as bf16 pairs; unused tail = 0
template <int NFREQ, int NWORDS>
__device__ __forceinline__ void encode_point(const float v[3], uint32_t (&out)[NWORDS]) {
  float vals[2 * NWORDS];
#pragma unroll
  for (int i = 0; i < 2 * NWORDS; ++i) vals[i] = 0.0f;
  vals[0] = v[0]; vals[1] = v[1]; vals[2] = v[2];
#pragma unroll
  for (int k = 0; k < NFREQ; ++k) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float s, c;
      fast_sincos(__fmul_rn(v[a], (float)(1 << k)), s, c);
      vals[3 + 6 * k + a] = s;
      vals[3 + 6 * k + 3 + a] = c;
    }
  }
#pragma unroll
  for (int i = 0; i < NWORDS; ++i) out[i] = pack_bf16(vals[2 * i], vals[2 * i + 1]);
}

// One 32-element half (HALF = 0: elements 0..31, 1: elements 32..63) of gamma(pts) (L = 10, 63 values + zero pad) as
// 16 packed bf16 pairs.  The two column halves of a tile's epilogue group each build one half: 15 / 16 sincos per thread.
template <int HALF>
__device__ __forceinline__ void encode_pts_half(const float v[3], uint32_t (&w)[16]) {
  float vals[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) vals[i] = 0.0f;
  if (HALF == 0) { vals[0] = v[0]; vals[1] = v[1]; vals[2] = v[2]; }
#pragma unroll
  for (int k = 0; k < 10; ++k) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int is = 3 + 6 * k + a, ic = 6 + 6 * k + a;
      const bool ns = (is >> 5) == HALF, nc = (ic >> 5) == HALF;
      if (ns || nc) {
        float sn, cs;
        fast_sincos(__fmul_rn(v[a], (float)(1 << k)), sn, cs);
        if (ns) vals[is & 31] = sn;
        if (nc) vals[ic & 31] = cs;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = pack_bf16(vals[2 * i], vals[2 * i + 1]);
}

// 16-byte chunks [j0, j0+NJ) of row r of a [128 x 64] swizzled atom <- packed words (zeros for dead rows / beyond the words)
template <int NW, int NJ, bool kStash>
__device__ __forceinline__ void store_enc_chunks(const uint32_t (&w)[NW], int j0, bool live, uint32_t atom_a, uint8_t* stash_atom,
                                                 int r) {
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) {
    uint4 q4 = make_uint4(0, 0, 0, 0);
    if (4 * jj < NW && live) q4 = make_uint4(w[(4 * jj) % NW], w[(4 * jj + 1) % NW], w[(4 * jj + 2) % NW], w[(4 * jj + 3) % NW]);
    const uint32_t off = sw128_off((uint32_t)r, (uint32_t)(j0 + jj));
    sts128(atom_a + off, q4.x, q4.y, q4.z, q4.w);
    if (kStash) *reinterpret_cast<uint4*>(stash_atom + off) = q4;
  }
}

// 32 accumulator columns of a hidden layer: h = acc + bias (ReLU), bf16, swizzled store into the A tile; returns the 32
// ReLU mask bits (bit = column).  MODE 0: ReLU, 1: ReLU + sigma-head partial from the fp32 h, 2: linear (feature layer).
//   bias_a: shared address of this thread's first bias entry (column cl = 0)    wa_a: bf16 sigma weights, same origin
//   row_a:  shared address of (this tile, this column half, row r, chunk 0)     rx: (r & 7) << 4
template <bool kTrain, int MODE>
__device__ __forceinline__ uint32_t epi_cols32(const uint32_t (&v)[32], const int cl, const uint32_t bias_a, const uint32_t wa_a,
                                               float& alpha, const uint32_t row_a, const uint32_t rx) {
  uint32_t mb = 0;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int c = cl + g * 8;                       // column inside this thread's 128 (compile-time after unrolling)
    const float4 b0 = lds128f(bias_a + c * 4), b1 = lds128f(bias_a + c * 4 + 16);
    const float2 h01 = __fadd2_rn(make_float2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1])), make_float2(b0.x, b0.y));
    const float2 h23 = __fadd2_rn(make_float2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3])), make_float2(b0.z, b0.w));
    const float2 h45 = __fadd2_rn(make_float2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5])), make_float2(b1.x, b1.y));
    const float2 h67 = __fadd2_rn(make_float2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7])), make_float2(b1.z, b1.w));
    float h[8] = {h01.x, h01.y, h23.x, h23.y, h45.x, h45.y, h67.x, h67.y};
    if (kTrain && MODE != 2) {
#pragma unroll
      for (int e = 0; e < 8; ++e) mb |= (h[e] > 0.f ? 1u : 0u) << (g * 8 + e);
    }
    if (MODE == 1) {
      const uint4 wq = lds128u(wa_a + c * 2);       // 8 bf16 sigma weights
      const uint32_t ww[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        h[e] = fmaxf(h[e], 0.f); h[e + 1] = fmaxf(h[e + 1], 0.f);
        alpha = fmaf(h[e], __uint_as_float(ww[e / 2] << 16), alpha);
        alpha = fmaf(h[e + 1], __uint_as_float(ww[e / 2] & 0xffff0000u), alpha);
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      o[e] = (MODE == 0) ? pack_relu_bf16(h[2 * e], h[2 * e + 1]) : pack_bf16(h[2 * e], h[2 * e + 1]);
    sts128(row_a + (uint32_t)(c / 64) * kAtomBytes + ((uint32_t)(((c % 64) / 8) << 4) ^ rx), o[0], o[1], o[2], o[3]);
  }
  return mb;
}

// one 256-wide layer epilogue for this thread: 128 columns as four 32-column TMEM loads, two in flight
template <bool kTrain, int MODE>
__device__ __forceinline__ void epi_layer(const uint32_t tmem_a, const uint32_t bias_a, const uint32_t wa_a, float& alpha,
                                          const uint32_t row_a, const uint32_t rx, uint32_t (&mk)[4]) {
  uint32_t va[32], vb[32];
  tmem_ld32(tmem_a, va);
  tmem_ld32(tmem_a + 32, vb);
  tmem_ld_wait_dep(va);
  tmem_ld_wait_dep(vb);
  mk[0] = epi_cols32<kTrain, MODE>(va, 0, bias_a, wa_a, alpha, row_a, rx);
  tmem_ld32(tmem_a + 64, va);
  mk[1] = epi_cols32<kTrain, MODE>(vb, 32, bias_a, wa_a, alpha, row_a, rx);
  tmem_ld32(tmem_a + 96, vb);
  tmem_ld_wait_dep(va);
  tmem_ld_wait_dep(vb);
  mk[2] = epi_cols32<kTrain, MODE>(va, 64, bias_a, wa_a, alpha, row_a, rx);
  mk[3] = epi_cols32<kTrain, MODE>(vb, 96, bias_a, wa_a, alpha, row_a, rx);
}

// 32 columns of the views layer: hv = relu(acc + bv), rgb += Wr[:, col] hv (fp32 partial), hv stashed in training.
//   col0: first column (0..127)   bias_a / wr_a: shared addresses of bias[0] / Wr[0][0] ([3][128] fp32)
template <bool kTrain>
__device__ __forceinline__ uint32_t epi_final32(const uint32_t (&v)[32], const int col0, const uint32_t bias_a, const uint32_t wr_a,
                                                float (&rgb)[3], uint8_t* stash_tile, const int r) {
  uint32_t mb = 0;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int col = col0 + g * 8;
    float h[8];
#pragma unroll
    for (int e = 0; e < 8; e += 4) {
      const float4 b4 = lds128f(bias_a + (col + e) * 4);
      const float4 r0 = lds128f(wr_a + (col + e) * 4);
      const float4 r1 = lds128f(wr_a + (128 + col + e) * 4);
      const float4 r2 = lds128f(wr_a + (256 + col + e) * 4);
      h[e + 0] = fmaxf(__uint_as_float(v[g * 8 + e + 0]) + b4.x, 0.f);
      h[e + 1] = fmaxf(__uint_as_float(v[g * 8 + e + 1]) + b4.y, 0.f);
      h[e + 2] = fmaxf(__uint_as_float(v[g * 8 + e + 2]) + b4.z, 0.f);
      h[e + 3] = fmaxf(__uint_as_float(v[g * 8 + e + 3]) + b4.w, 0.f);
      rgb[0] = fmaf(h[e], r0.x, fmaf(h[e + 1], r0.y, fmaf(h[e + 2], r0.z, fmaf(h[e + 3], r0.w, rgb[0]))));
      rgb[1] = fmaf(h[e], r1.x, fmaf(h[e + 1], r1.y, fmaf(h[e + 2], r1.z, fmaf(h[e + 3], r1.w, rgb[1]))));
      rgb[2] = fmaf(h[e], r2.x, fmaf(h[e + 1], r2.y, fmaf(h[e + 2], r2.z, fmaf(h[e + 3], r2.w, rgb[2]))));
    }
    if (kTrain) {
#pragma unroll
      for (int e = 0; e < 8; ++e) mb |= (h[e] > 0.f ? 1u : 0u) << (g * 8 + e);
      const uint4 v4 = make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]), pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7]));
      *reinterpret_cast<uint4*>(stash_tile + (size_t)(SA_HV + col / 64) * kAtomBytes + sw128_off(r, (col % 64) / 8)) = v4;
    }
  }
  return mb;
}

// Weight-ring schedule shared by the producer and the MMA issuer.  Both tiles of a CTA run the same layer back to back,
// so a chunk is loaded ONCE per layer where the 3-slot ring allows it: a 4-chunk layer is loaded as c0 c1 c2 c3 c0 —
// tile 0 consumes c0..c3 (c0's slot is recycled for c3), tile 1 consumes c1 c2 c3 from the slots tile 0 left behind and
// then the reloaded c0 — 5 loads instead of 8; layers of <= 3 chunks are loaded once for both tiles.  Slots are still
// released in load order, so the ring stays a FIFO with one full/empty mbarrier pair per slot.
__device__ __forceinline__ int ring_loads(int nch) { return nch == 4 ? 5 : nch; }

template <bool kTrain>
__global__ void __launch_bounds__(kThreads, 1) mlp_fwd_kernel(const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (smem != smem_raw) __trap();   // kSmemBytes has no alignment slack: the dynamic window must start 1024-aligned
  // barriers: full[3] empty[3] acc_full[2] act_ready[2]
  const uint32_t bar_full = sbase + SM_BAR, bar_empty = bar_full + 8 * kStages;
  const uint32_t bar_acc = bar_empty + 8 * kStages, bar_act = bar_acc + 16;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SM_TMEMPTR);
  const float* cst = reinterpret_cast<const float*>(p.packed + kFwdBytes + kBwdBytes);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int t = 0; t < 2; ++t) { mbar_init(bar_acc + 8 * t, 1); mbar_init(bar_act + 8 * t, kEpiThreads); }
    fence_mbar_init();
  }
  if (warp == 1) {   // TMEM: 512 columns = two 128x256 fp32 accumulators
    tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 256) {   // sigma-head weights as bf16, read by the layer-7 epilogues
    const int i = threadIdx.x - 64;
    const __nv_bfloat16 w = __float2bfloat16_rn(__ldg(cst + C_WA + i));
    sts16(sbase + SM_WA + 2 * i, *reinterpret_cast<const uint16_t*>(&w));
  }
  tcgen05_fence_before_sync();
  __syncthreads();
  tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int my_pairs = (p.num_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    // ================= weight producer =================
    // Bulk copies have a fixed per-copy cost and ONE thread retires at most one per ~700 cycles (tools/bulk_rate.py):
    // 4 KB pieces top out at 32 B/cycle/SM, 16 KB pieces reach ~70.  A chunk is therefore moved as 16 KB pieces and
    // every (ring stage, piece) has its own issuing lane, so no lane issues more often than once per ring revolution.
    const uint32_t piece = (uint32_t)p.piece;
    uint32_t stage = 0, phase = 0;
    for (int it = 0; it < my_pairs; ++it) {
      const uint8_t* src = p.packed;
      for (int s = 0; s < kNumSteps; ++s) {
        const uint32_t bytes = (uint32_t)c_step_n[s] * 128u;
        const int npieces = (int)(bytes / piece);
        const int nch = c_step_chunks[s], nl = ring_loads(nch);
        for (int j = 0; j < nl; ++j) {
          const uint8_t* sp = src + (size_t)(j & 3) * bytes;     // load j carries chunk j mod 4
          if (lane == 0) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            mbar_arrive_expect_tx(bar_full + 8 * stage, bytes);
            trace_stamp(p.trace, it, s, j >> 2, 12 + (j & 3));
          }
          __syncwarp();
          const int pi = lane - (int)stage * (int)(kChunkBig / piece);
          if (pi >= 0 && pi < npieces)
            bulk_g2s(sbase + SM_RING + stage * kChunkBig + pi * piece, sp + pi * piece, piece, bar_full + 8 * stage);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        src += (size_t)nch * bytes;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      uint32_t ld_idx = 0;                                  // loads consumed so far: load i sits in slot i % 3, phase (i / 3) & 1
      uint32_t act_phase[2] = {0, 0};
      for (int it = 0; it < my_pairs; ++it) {
        for (int s = 0; s < kNumSteps; ++s) {
          const int nch = c_step_chunks[s], n = c_step_n[s], ksteps = c_step_ksteps[s];
          const uint32_t idesc = make_idesc(kTileM, n, 0, 0);
          for (int t = 0; t < 2; ++t) {
            mbar_wait(bar_act + 8 * t, act_phase[t]);   // A operand written, accumulator drained
            act_phase[t] ^= 1;
            tcgen05_fence_after_sync();
            trace_stamp(p.trace, it, s, t, 0);
            const uint32_t d_tmem = tmem_base + (uint32_t)t * 256u;
            uint32_t accumulate = (uint32_t)c_step_acc[s];
            for (int ci = 0; ci < nch; ++ci) {
              // which load / chunk this MMA group uses (see ring_loads)
              const int j = (nch == 4 && t == 1) ? ci + 1 : ci;
              const int c = (nch == 4) ? (j & 3) : ci;
              const bool first_use = (t == 0) || (nch == 4 && ci == 3);
              const bool last_use = (t == 1) || (nch == 4 && ci == 0);
              const uint32_t li = ld_idx + (uint32_t)j;
              const uint32_t stage = li % kStages, phase = (li / kStages) & 1u;
              if (first_use) {
                mbar_wait(bar_full + 8 * stage, phase);
                tcgen05_fence_after_sync();
              }
              if (ci == 0) trace_stamp(p.trace, it, s, t, 1);
              if (ci == nch - 1) trace_stamp(p.trace, it, s, t, 2);
              trace_stamp(p.trace, it, s, t, 8 + ci);
              const uint32_t a_addr = sbase + SM_ACT + t * kActBytes + c * kAtomBytes;
              const uint32_t b_addr = sbase + SM_RING + stage * kChunkBig;
              const uint64_t a_desc = make_smem_desc(a_addr, 16, 1024);
              const uint64_t b_desc = make_smem_desc(b_addr, 16, 1024);
              for (int k = 0; k < ksteps; ++k) {
                // +32 bytes (16 bf16) along K inside the 128-byte swizzle atom: start-address field += 2
                umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, accumulate);
                accumulate = 1;
              }
              if (last_use) umma_commit(bar_empty + 8 * stage);   // ring slot free once these MMAs have read it
            }
            umma_commit(bar_acc + 8 * t);              // accumulator complete -> epilogue of tile t
            trace_stamp(p.trace, it, s, t, 3);
          }
          ld_idx += (uint32_t)ring_loads(nch);
        }
      }
    }
  } else {
    // ================= prologue + epilogue warps: 8 per tile =================
    // warp (q, cg): TMEM lane quarter q = warp % 4 (rows 32q..32q+31), column half cg (columns 128cg..128cg+127)
    const int ew = warp - 2;
    const int t = ew >> 3;                               // tile slot 0/1
    const int cg = (ew >> 2) & 1;
    const int q = warp & 3;
    const int r = q * 32 + lane;                         // row inside the tile
    const int tix = cg * 128 + r;                        // 0..255 inside the tile's epilogue group
    uint8_t* act = smem + SM_ACT + t * kActBytes;
    const uint32_t act_a = sbase + SM_ACT + t * kActBytes;
    const uint32_t rx = (uint32_t)(r & 7) << 4;
    const uint32_t row_a = act_a + (uint32_t)cg * 2u * kAtomBytes + (uint32_t)r * 128u;   // this thread's row, its column half
    const uint32_t bias_row = sbase + SM_BIAS + (uint32_t)t * 1024u;
    const uint32_t wa_a = sbase + SM_WA + (uint32_t)cg * 256u;
    const uint32_t wr_a = act_a + 2 * kAtomBytes;        // FINAL step: Wr [3][128] fp32 staged in the (dead) tile
    float4* xchg = reinterpret_cast<float4*>(act + 3 * kAtomBytes);   // FINAL-step scratch (tile is dead by then)
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)t * 256u;
    const bool store_lane = kTrain && lane == 0 && cg == 0;          // issues this warp's 16 KB stash stores
    uint32_t acc_phase = 0;
    for (int it = 0; it < my_pairs; ++it) {
      const int64_t tile = 2 * ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) + t;
      const int64_t row = tile * kTileM + r;
      const bool live = row < p.m;
      uint8_t* stash_tile = kTrain ? p.stash + (size_t)tile * kStashTileBytes : nullptr;
      // ---- prologue: sample point -> gamma(pts) -> A atom 0, one 32-element half per column half.
      //      Only the 3-D point / direction stay in registers; encodings are re-derived when a pass needs them.
      float pt[3] = {0, 0, 0}, dir[3] = {0, 0, 0};
      if (live) fetch_sample(p.src, row, pt, dir);
      {
        uint32_t w[16];
        if (cg == 0) encode_pts_half<0>(pt, w); else encode_pts_half<1>(pt, w);
        store_enc_chunks<16, 4, kTrain>(w, 4 * cg, live, act_a, stash_tile + (size_t)SA_ENC * kAtomBytes, r);
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_act + 8 * t);
      float alpha = 0.0f;                                // this half's partial of the sigma head
      int pending_atom = -1;                             // stash atom of the layer output still to be streamed out
      for (int s = 0; s < kNumSteps; ++s) {
        const int epi = c_step_epi[s];
        // (1) the whole group has finished the previous epilogue: its bias row is free, its output tile complete
        named_bar_sync(1 + t, kEpiThreads);
        if (kTrain && pending_atom >= 0 && store_lane) {      // one 16 KB atom per warp: bulk-copy issue is serialised per thread
          bulk_s2g(stash_tile + (size_t)(pending_atom + q) * kAtomBytes, act_a + q * kAtomBytes, kAtomBytes);
          bulk_commit();
        }
        // (2) everything that can be done while this step's MMAs are still running: bias row -> shared memory (there
        //     is no L1 left with 227 KB carved out, so a __ldg costs an L2 round trip), encodings, head weights
        uint32_t encw[16];
        float wr0 = 0.f, wr1 = 0.f;
        float4 brba = make_float4(0.f, 0.f, 0.f, 0.f);
        if (epi == EPI_WRITE_ENC) {
          if (cg == 0) encode_pts_half<0>(pt, encw); else encode_pts_half<1>(pt, encw);
        } else if (epi == EPI_WRITE_DENC) {
          if (cg == 1) encode_point<4, 16>(dir, encw);
        } else {
          sts32f(bias_row + 4 * tix, __ldg(cst + c_step_bias[s] + (epi == EPI_FINAL ? (tix & 127) : tix)));
          if (epi == EPI_FINAL) {
            wr0 = __ldg(cst + C_WR + tix);
            if (tix < 128) wr1 = __ldg(cst + C_WR + 256 + tix);
            if (cg == 0) brba = make_float4(__ldg(cst + C_BR), __ldg(cst + C_BR + 1), __ldg(cst + C_BR + 2), __ldg(cst + C_BA));
          }
        }
        if (store_lane && pending_atom >= 0) bulk_wait_read0();   // the stash store has finished READING the tile
        pending_atom = -1;
        named_bar_sync(1 + t, kEpiThreads);                   // bias row visible, tile free to overwrite after the MMAs
        mbar_wait(bar_acc + 8 * t, acc_phase);
        acc_phase ^= 1;
        tcgen05_fence_after_sync();
        if (tix == 0) trace_stamp(p.trace, it, s, t, 4);
        if (epi == EPI_WRITE_ENC || epi == EPI_WRITE_DENC) {
          // pass 1 has finished reading the tile: overwrite atom 0 with the second-pass operand
          if (epi == EPI_WRITE_ENC) store_enc_chunks<16, 4, false>(encw, 4 * cg, live, act_a, nullptr, r);
          else if (cg == 1) store_enc_chunks<16, 8, kTrain>(encw, 0, live, act_a, stash_tile + (size_t)SA_DENC * kAtomBytes, r);
          fence_proxy_async_smem();
          mbar_arrive(bar_act + 8 * t);
          if (tix == 0) trace_stamp(p.trace, it, s, t, 6);
          continue;
        }
        if (epi == EPI_FINAL) {
          // hv = relu(acc + bv) [128];  rgb = Wr hv + br;  raw = [rgb, alpha]   (helpers:117-123); 64 columns per half
          sts32f(wr_a + 4 * tix, wr0);
          if (tix < 128) sts32f(wr_a + 4 * (256 + tix), wr1);
          named_bar_sync(1 + t, kEpiThreads);
          float rgb[3] = {0.f, 0.f, 0.f};
          const int c0 = cg * 64;
          uint32_t va[32], vb[32];
          tmem_ld32(tmem_lane + c0, va);
          tmem_ld32(tmem_lane + c0 + 32, vb);
          tmem_ld_wait_dep(va);
          tmem_ld_wait_dep(vb);
          const uint32_t m0 = epi_final32<kTrain>(va, c0, bias_row, wr_a, rgb, stash_tile, r);
          const uint32_t m1 = epi_final32<kTrain>(vb, c0 + 32, bias_row, wr_a, rgb, stash_tile, r);
          if (kTrain) {
            uint32_t* mrow = reinterpret_cast<uint32_t*>(stash_tile + kStashMaskOff) + (8 * 128 + r) * 8 + cg * 2;
            *reinterpret_cast<uint2*>(mrow) = make_uint2(m0, m1);
          }
          if (cg == 1) xchg[r] = make_float4(rgb[0], rgb[1], rgb[2], alpha);
          tcgen05_fence_before_sync();
          named_bar_sync(1 + t, kEpiThreads);
          if (cg == 0 && live) {
            const float4 o = xchg[r];
            *reinterpret_cast<float4*>(p.raw + row * 4) =
                make_float4(rgb[0] + o.x + brba.x, rgb[1] + o.y + brba.y, rgb[2] + o.z + brba.z, alpha + o.w + brba.w);
          }
          if (tix == 0) trace_stamp(p.trace, it, s, t, 6);
          continue;   // next arrival on act_ready comes from the next tile's prologue
        }
        // ---- bias (+ReLU) -> bf16 -> swizzled in-place store; layer 7 also accumulates sigma from fp32 h7
        uint32_t mk[4];
        const uint32_t my_bias = bias_row + (uint32_t)cg * 512u;
        const uint32_t my_tmem = tmem_lane + (uint32_t)cg * 128u;
        if (epi == EPI_RELU) epi_layer<kTrain, 0>(my_tmem, my_bias, wa_a, alpha, row_a, rx, mk);
        else if (epi == EPI_RELU_ALPHA) epi_layer<kTrain, 1>(my_tmem, my_bias, wa_a, alpha, row_a, rx, mk);
        else epi_layer<kTrain, 2>(my_tmem, my_bias, wa_a, alpha, row_a, rx, mk);
        if (kTrain && c_step_mask_slot[s] >= 0) {
          uint32_t* mrow = reinterpret_cast<uint32_t*>(stash_tile + kStashMaskOff) + (c_step_mask_slot[s] * 128 + r) * 8 + cg * 4;
          *reinterpret_cast<uint4*>(mrow) = make_uint4(mk[0], mk[1], mk[2], mk[3]);
        }
        tcgen05_fence_before_sync();
        fence_proxy_async_smem();
        mbar_arrive(bar_act + 8 * t);
        if (tix == 0) trace_stamp(p.trace, it, s, t, 6);
        if (kTrain) pending_atom = c_step_stash_atom[s];
      }
    }
    if (store_lane) bulk_wait0();   // every stash store has landed before the kernel exits
  }

  tcgen05_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

static int check_arch() {
  static int ok = -1;
  if (ok < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
      major = 0;
    ok = (major == 10) ? 1 : 0;
  }
  if (!ok) {
    set_error("the tcgen05 MLP kernels need an sm_100 device (B200)");
    return SPN_E_ARCH;
  }
  return SPN_OK;
}

int mlp_tc_fwd(const void* packed, const SampleSource& src, int64_t m, float* raw, void* stash, cudaStream_t st) {
  int rc = check_arch();
  if (rc != SPN_OK) return rc;
  SPN_CHECK_ARG(((uintptr_t)packed & 15) == 0 && (!stash || ((uintptr_t)stash & 15) == 0), "mlp_tc_fwd: unaligned buffer");
  FwdParams p;
  p.packed = (const uint8_t*)packed; p.src = src; p.m = m; p.raw = raw; p.stash = (uint8_t*)stash;
  int64_t tiles = (m + kTileM - 1) / kTileM;
  p.num_pairs = (int)((tiles + 1) / 2);
  p.piece = weight_piece_bytes();
  p.trace = g_trace;
  int grid = p.num_pairs < sm_count() ? p.num_pairs : sm_count();
  auto kern = stash ? mlp_fwd_kernel<true> : mlp_fwd_kernel<false>;
  static bool attr_set[2] = {false, false};
  if (!attr_set[stash ? 1 : 0]) {
    SPN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set[stash ? 1 : 0] = true;
  }
  prof_begin(PROF_MLP_FWD, st);
  kern<<<grid, kThreads, kSmemBytes, st>>>(p);
  prof_end(PROF_MLP_FWD, st);
  SPN_LAUNCH_CHECK("mlp_fwd_kernel");
  return SPN_OK;
}

// ---- diagnostic: one UMMA GEMM  D[128,N] = A[128,K] * B[N,K]^T  (bf16 operands, fp32 accumulate) -------------
// Exercises exactly the descriptor / swizzle / TMEM conventions the MLP kernels rely on, in isolation.
__global__ void __launch_bounds__(128, 1) selftest_gemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                                float* __restrict__ D, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  uint8_t* sA = smem;                         // K/64 atoms of [128 x 64]
  uint8_t* sB = smem + 4 * kAtomBytes;        // K/64 chunks of [N x 64]
  const uint32_t bar = sbase + 4 * kAtomBytes + 4 * kChunkBig;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 4 * kAtomBytes + 4 * kChunkBig + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int katoms = K / 64;
  for (int i = threadIdx.x; i < 128 * katoms * 8; i += 128) {      // A pieces
    int j = i & 7, rr = (i >> 3) & 127, at = i >> 10;
    const float* s = A + (size_t)rr * K + at * 64 + j * 8;
    *reinterpret_cast<uint4*>(sA + at * kAtomBytes + sw128_off(rr, j)) =
        make_uint4(pack_bf16(s[0], s[1]), pack_bf16(s[2], s[3]), pack_bf16(s[4], s[5]), pack_bf16(s[6], s[7]));
  }
  for (int i = threadIdx.x; i < N * katoms * 8; i += 128) {        // B pieces
    int j = i & 7, rr = (i >> 3) % N, at = (i >> 3) / N;
    const float* s = B + (size_t)rr * K + at * 64 + j * 8;
    *reinterpret_cast<uint4*>(sB + at * (N * 128) + sw128_off(rr, j)) =
        make_uint4(pack_bf16(s[0], s[1]), pack_bf16(s[2], s[3]), pack_bf16(s[4], s[5]), pack_bf16(s[6], s[7]));
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(tptr), 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tcgen05_fence_before_sync();
  __syncthreads();
  tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(128, N, 0, 0);
    uint32_t acc = 0;
    for (int at = 0; at < katoms; ++at) {
      const uint64_t a_desc = make_smem_desc(sbase + at * kAtomBytes, 16, 1024);
      const uint64_t b_desc = make_smem_desc(sbase + 4 * kAtomBytes + at * (N * 128), 16, 1024);
      for (int k = 0; k < 4; ++k) { umma_bf16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, acc); acc = 1; }
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tcgen05_fence_after_sync();
  for (int cb = 0; cb < N / 32; ++cb) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + cb * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(size_t)(warp * 32 + lane) * N + cb * 32 + j] = __uint_as_float(v[j]);
  }
  tcgen05_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after_sync(); tmem_dealloc(tmem_base, 256); }
}

// ---- diagnostic: issue rate of back-to-back tcgen05.mma for K-major / MN-major operand combinations ---------------
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int a_mn, int b_mn, int n, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  for (int i = threadIdx.x; i < (64 * 1024) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  const uint32_t bar = sbase + 64 * 1024;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 64 * 1024 + 64);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(tptr), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tcgen05_fence_before_sync();
  __syncthreads();
  tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(128, n, a_mn, b_mn);
    // K-major: [rows x 64] atoms, +32 B per k-step; MN-major: 64-wide blocks 4 KB apart, 2 KB per k-step
    const uint64_t a0 = a_mn ? make_smem_desc(sbase, 4096, 1024) : make_smem_desc(sbase, 16, 1024);
    const uint64_t b0 = b_mn ? make_smem_desc(sbase + 16384, 4096, 1024) : make_smem_desc(sbase + 16384, 16, 1024);
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      const int k = i & 1;
      umma_bf16(tmem_base, a0 + (uint64_t)(a_mn ? 128 * k : 2 * k), b0 + (uint64_t)(b_mn ? 128 * k : 2 * k), idesc, 1);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  tcgen05_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after_sync(); tmem_dealloc(tmem_base, 512); }
}

int tc_mma_rate(int a_mn, int b_mn, int n, int reps, long long* cycles_dev, cudaStream_t st) {
  int rc = check_arch();
  if (rc != SPN_OK) return rc;
  const int smem_bytes = 64 * 1024 + 256 + 1024;
  SPN_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  mma_rate_kernel<<<1, 128, smem_bytes, st>>>(a_mn, b_mn, n, reps, cycles_dev);
  SPN_LAUNCH_CHECK("mma_rate_kernel");
  return SPN_OK;
}

// ---- diagnostic: cp.async.bulk global->shared throughput per SM vs copy size and copies in flight ------------------
__global__ void __launch_bounds__(32, 1) bulk_rate_kernel(const uint8_t* __restrict__ src, size_t src_bytes, int copy_bytes,
                                                          int depth, int iters, int lanes_arg, long long* out) {
  const int shared_bar = lanes_arg >= 200;     // lanes + 200: all lanes' copies of a slot land on one mbarrier
  const int poll = !shared_bar && lanes_arg >= 100;   // lanes + 100 selects the polling (test_wait) variant
  const int lanes = lanes_arg % 100;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + 200 * 1024;
  if (threadIdx.x == 0) {
    for (int d = 0; d < depth * lanes; ++d) mbar_init(bar0 + 8 * d, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if ((int)threadIdx.x < lanes) {
    // every issuing lane of every CTA streams its own region, so nothing is shared in L2; each lane owns `depth` slots
    const size_t region = src_bytes / ((size_t)gridDim.x * lanes) / 1024 * 1024;
    const uint8_t* base = src + ((size_t)blockIdx.x * lanes + threadIdx.x) * region;
    const uint32_t bar0 = sbase + 200 * 1024 + 8 * depth * threadIdx.x;
    const uint32_t sbase = smem_u32(smem) + (uint32_t)threadIdx.x * depth * copy_bytes;
    const size_t per = (size_t)copy_bytes;
    const size_t wrap = region / per;
    long long t0 = clock64();
    if (shared_bar) {
      // wgrad's pattern: all lanes' copies of a slot complete on ONE mbarrier (lane 0 posts the byte count)
      const uint32_t barS = smem_u32(smem) + 200 * 1024;
      for (int i = 0; i < iters + depth; ++i) {
        const int slot = i % depth;
        if (i >= depth) mbar_wait(barS + 8 * slot, (uint32_t)(((i - depth) / depth) & 1));
        if (i < iters) {
          if (threadIdx.x == 0) mbar_arrive_expect_tx(barS + 8 * slot, copy_bytes * lanes);
          __syncwarp((1u << lanes) - 1);
          bulk_g2s(sbase + slot * copy_bytes, base + ((size_t)i % wrap) * per, copy_bytes, barS + 8 * slot);
        }
      }
    } else
    for (int i = 0; i < iters + depth; ++i) {
      const int slot = i % depth;
      if (i >= depth) {                                    // retire the slot's previous copy
        if (poll) mbar_wait_poll(bar0 + 8 * slot, (uint32_t)(((i - depth) / depth) & 1));
        else mbar_wait(bar0 + 8 * slot, (uint32_t)(((i - depth) / depth) & 1));
      }
      if (i < iters) {
        mbar_arrive_expect_tx(bar0 + 8 * slot, copy_bytes);
        bulk_g2s(sbase + slot * copy_bytes, base + ((size_t)i % wrap) * per, copy_bytes, bar0 + 8 * slot);
      }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
}

int tc_bulk_rate(const void* src, size_t src_bytes, int copy_bytes, int depth, int iters, int grid, int lanes,
                 long long* out, cudaStream_t st) {
  SPN_CHECK_ARG(src && out && copy_bytes >= 1024 && copy_bytes % 1024 == 0 && depth >= 1 && depth <= 32 &&
                lanes >= 1 && (lanes % 100) * depth <= 32 && (size_t)copy_bytes * depth * (lanes % 100) <= 200 * 1024 && grid >= 1,
                "spn_tc_bulk_rate: bad arguments");
  const int smem_bytes = 200 * 1024 + 512 + 1024;
  SPN_CUDA(cudaFuncSetAttribute(bulk_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  bulk_rate_kernel<<<grid, 32, smem_bytes, st>>>((const uint8_t*)src, src_bytes, copy_bytes, depth, iters, lanes, out);
  SPN_LAUNCH_CHECK("bulk_rate_kernel");
  return SPN_OK;
}

int tc_selftest_gemm(const float* A, const float* B, float* D, int N, int K, cudaStream_t st) {
  int rc = check_arch();
  if (rc != SPN_OK) return rc;
  SPN_CHECK_ARG(A && B && D && (N == 128 || N == 256) && K >= 64 && K <= 256 && K % 64 == 0, "spn_tc_selftest_gemm: N in {128,256}, K in {64..256}");
  const int smem_bytes = 4 * kAtomBytes + 4 * kChunkBig + 128 + 1024;
  SPN_CUDA(cudaFuncSetAttribute(selftest_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  selftest_gemm_kernel<<<1, 128, smem_bytes, st>>>(A, B, D, N, K);
  SPN_LAUNCH_CHECK("selftest_gemm_kernel");
  return SPN_OK;
}

}  // namespace spn
