// NeRF MLP (DS_NeRF/run_nerf_helpers.py:74-127) fused with sampling-point generation and positional
// encoding (run_nerf.py:56-71, 670; helpers:22-70) on the 5th-generation tensor cores (tcgen05).
//
// Persistent CTA PAIRS (thread-block clusters of 2 on neighbouring SMs, cta_group::2 MMAs), 576 threads per CTA:
//   warp 0      weight producer: this CTA's HALF (its 128 of 256 output rows) of every pre-swizzled bf16 weight chunk,
//               cp.async.bulk (TMA engine) into a 3-slot ring of two-chunk groups, mbarrier tx counts
//   warp 1      leader CTA: one lane issues the pair's tcgen05.mma (M = 256 = one tile slot of both CTAs, N = 256|128,
//               K = 16) and the multicast commits; peer CTA: relays "my weight halves have landed" to the leader
//   warps 2-9   epilogue / prologue of tile slot 0 (rows 0..127 -> TMEM lanes 0..127, 2 column halves x 4 lane quarters)
//   warps 10-17 epilogue / prologue of tile slot 1
// Two 128-sample tiles per CTA ping-pong on the tensor pipe: while slot 0's epilogue turns its fp32 accumulator (TMEM,
// 256 columns) into the next layer's bf16 A operand (bias, ReLU, cast, 128B-swizzled store to shared memory, in place),
// slot 1's layer runs on the tensor cores, and vice versa.  Activations never leave the SM; HBM sees 24 B in + 16 B out
// per sample (plus the bf16 stash in training).  See the comment above mlp_fwd_kernel for the barrier protocol.
//
// The 63-wide skip input of layer 5 and the 27-wide view encoding of the views layer are applied as a
// second accumulating pass (K=64 / K=32) after the 256-wide pass, so the 128x256 activation tile can be
// updated in place; the encodings are re-derived from the 3-D point while the previous pass runs.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "mlp_tc.cuh"

namespace spn {
using namespace tc;

size_t mlp_tc_packed_bytes() { return kPackedBytes; }

// diagnostic timeline buffer (device pointer, kTraceSlots int64): see tools/trace_fwd.py
static long long* g_trace = nullptr;
void tc_set_trace(long long* dev) { g_trace = dev; }
long long* tc_get_trace() { return g_trace; }
constexpr int kTraceRounds = 3, kTraceEvents = 24;
constexpr int kTraceSlots = kTraceRounds * 12 * 2 * kTraceEvents;
__device__ __forceinline__ void trace_stamp(long long* tr, int it, int s, int t, int e) {
  if (it < kTraceRounds) tr[((it * 12 + s) * 2 + t) * kTraceEvents + e] = clock64();
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
  int64_t tiles = (m + 4 * kTileM - 1) / (4 * kTileM) * 4;   // a CTA pair processes 4 tiles per round
  return (size_t)tiles * kStashTileBytes + 256;
}

struct FwdParams {
  const uint8_t* packed;
  SampleSource src;
  int64_t m;
  float* raw;
  uint8_t* stash;   // nullable
  int num_quads;    // groups of 4 tiles: one round of a CTA pair (2 tile slots per CTA)
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

// ReLU mask of 32 columns, built from the 16 packed post-ReLU bf16 pairs in column order: a non-negative bf16 is non-zero
// iff adding 0x7FFF carries into bit 15, so (w + 0x7FFF7FFF) & 0x80008000 flags both halves at once and shifting the
// accumulator right by one per pair leaves pair i's flags at bits i (column 2i) and 16+i (column 2i+1): 3 instructions
// per pair instead of a compare + select + or per element.  relu_mask_bit(c) is where column c of the block ends up.
__device__ __forceinline__ uint32_t relu_mask_push(uint32_t acc, uint32_t packed_pair) {
  return (acc >> 1) | ((packed_pair + 0x7FFF7FFFu) & 0x80008000u);
}

// 32 accumulator columns of a hidden layer: h = acc + bias (ReLU), bf16, swizzled store into the A tile; returns the 32
// ReLU mask bits (layout: relu_mask_push).  MODE 0: ReLU, 1: ReLU + sigma-head partial from the fp32 h, 2: linear (feature layer).
//   bias_a: shared address of this thread's first bias entry (column cl = 0)    wa_a: bf16 sigma weights, same origin
//   row_a:  shared address of (this tile, this column half, row r, chunk 0)     rx: (r & 7) << 4
template <bool kTrain, int MODE>
__device__ __forceinline__ uint32_t epi_cols32(const uint32_t (&v)[32], const int cl, const uint32_t bias_a, const uint32_t wa_a,
                                               float& alpha, const uint32_t row_a, const uint32_t rx) {
  uint32_t mb = 0;
  // the bias row is read one 8-column group ahead: the shared-memory latency hides behind the previous group's math
  float4 nb0 = lds128f(bias_a + cl * 4), nb1 = lds128f(bias_a + cl * 4 + 16);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int c = cl + g * 8;                       // column inside this thread's 128 (compile-time after unrolling)
    const float4 b0 = nb0, b1 = nb1;
    if (g < 3) { nb0 = lds128f(bias_a + (c + 8) * 4); nb1 = lds128f(bias_a + (c + 8) * 4 + 16); }
    const float2 h01 = __fadd2_rn(make_float2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1])), make_float2(b0.x, b0.y));
    const float2 h23 = __fadd2_rn(make_float2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3])), make_float2(b0.z, b0.w));
    const float2 h45 = __fadd2_rn(make_float2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5])), make_float2(b1.x, b1.y));
    const float2 h67 = __fadd2_rn(make_float2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7])), make_float2(b1.z, b1.w));
    float h[8] = {h01.x, h01.y, h23.x, h23.y, h45.x, h45.y, h67.x, h67.y};
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
    if (kTrain && MODE != 2) {
#pragma unroll
      for (int e = 0; e < 4; ++e) mb = relu_mask_push(mb, o[e]);
    }
    sts128(row_a + (uint32_t)(c / 64) * kAtomBytes + ((uint32_t)(((c % 64) / 8) << 4) ^ rx), o[0], o[1], o[2], o[3]);
  }
  return mb;
}

// one 256-wide layer epilogue for this thread: 128 columns as four 32-column TMEM loads, two in flight
template <bool kTrain, int MODE>
__device__ __forceinline__ void epi_layer(const uint32_t tmem_a, const uint32_t bias_a, const uint32_t wa_a, float& alpha,
                                          const uint32_t row_a, const uint32_t rx, uint32_t (&mk)[4], long long* tr) {
  uint32_t va[32], vb[32];
  tmem_ld32(tmem_a, va);
  tmem_ld32(tmem_a + 32, vb);
  tmem_ld_wait_dep(va);
  tmem_ld_wait_dep(vb);
  if (tr) tr[16] = clock64();
  mk[0] = epi_cols32<kTrain, MODE>(va, 0, bias_a, wa_a, alpha, row_a, rx);
  if (tr) tr[17] = clock64();
  tmem_ld32(tmem_a + 64, va);
  mk[1] = epi_cols32<kTrain, MODE>(vb, 32, bias_a, wa_a, alpha, row_a, rx);
  if (tr) tr[18] = clock64();
  tmem_ld32(tmem_a + 96, vb);
  tmem_ld_wait_dep(va);
  tmem_ld_wait_dep(vb);
  if (tr) tr[19] = clock64();
  mk[2] = epi_cols32<kTrain, MODE>(va, 64, bias_a, wa_a, alpha, row_a, rx);
  if (tr) tr[20] = clock64();
  mk[3] = epi_cols32<kTrain, MODE>(vb, 96, bias_a, wa_a, alpha, row_a, rx);
  if (tr) tr[21] = clock64();
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
      const uint4 v4 = make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]), pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7]));
      mb = relu_mask_push(relu_mask_push(relu_mask_push(relu_mask_push(mb, v4.x), v4.y), v4.z), v4.w);
      *reinterpret_cast<uint4*>(stash_tile + (size_t)(SA_HV + col / 64) * kAtomBytes + sw128_off(r, (col % 64) / 8)) = v4;
    }
  }
  return mb;
}

// CTA pairs (cta_group::2).  Two CTAs on neighbouring SMs form a cluster; each owns two 128-sample tile slots that
// ping-pong as before, but every tcgen05.mma spans the pair (M = 256: slot t of both CTAs), so each SM reads only ITS
// half of the weight chunk from shared memory (128 of the 256 output rows) and loads only that half from L2: per-SM
// shared-memory traffic per layer drops from 336 KB to 224 KB (operand reads 192 -> 128, weight writes 80 -> 32), which
// is what bounded the single-CTA version (tools/trace_fwd.py: MMAs ran at ~190 instead of 128 cycles with the
// 128 B/cycle/SM shared-memory pipe saturated).  With 16 KB half-chunks the 96 KB ring holds 6 of them: a layer's four
// chunks are loaded ONCE per round and stay resident for both tile slots, the rest prefetches the next layer (see the
// ring / barrier comment inside the kernel; tests/test_ring_protocol_model.py is an executable model of the protocol).
//   leader CTA (cluster rank 0): warp 1 lane 0 issues the MMAs and the multicast commits (ring slot free / accumulator
//     ready arrive on the same barrier offsets in both CTAs)
//   peer CTA: warp 1 lane 0 relays "my halves of this weight group have landed" to the leader's group barrier
//   both: warp 0 loads the CTA's halves; the 16 epilogue warps signal "A tile written" to the leader's act barrier
//     (one elected lane per warp; remote arrive from the peer)
template <bool kTrain>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1) mlp_fwd_kernel(const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0: leader (issues the pair's MMAs)

  if (smem != smem_raw) __trap();   // kSmemBytes has no alignment slack: the dynamic window must start 1024-aligned
  // Weight ring: 3 slots, each one GROUP = two 16 KB half-chunks (chunks 2g, 2g+1 of a layer).  Every chunk is loaded
  // once per round and used by BOTH tile slots; a layer occupies two slots, the third prefetches group 0 of the next
  // layer, whose group 1 follows as soon as tile slot 1 is through group 0 of the current one.  FIFO.
  // barriers: group_full[2][4]  group g of step s uses barrier [g][s % 4] — the ring never holds groups of two steps that
  //             are 4 apart.  Armed by the local producer and, on the leader, by the peer's relay as well (count 2).
  //           empty[3] (one multicast commit per group, after tile slot 1 has used it)   acc_full[2], act_ready[2]
  // The MMA-issuing thread is the scarce resource (tools/trace_fwd.py): tcgen05.mma issue blocks once a few MMAs are
  // queued, a tcgen05.commit costs it ~130 cycles and even a satisfied mbarrier wait 200-400, during which the tensor
  // pipe drains.  Hence few, merged barriers: per layer it waits on group 0 (before the A tile, off the critical path),
  // the two A tiles and group 1 (mid-layer, tile slot 0 only), and commits 4 times.
  static_assert(kNumSteps % 4 == 0, "group barrier phase bookkeeping assumes a multiple of 4 steps per round");
  const uint32_t bar_full = sbase + SM_BAR, bar_empty = bar_full + 64;
  const uint32_t bar_acc = bar_empty + 8 * kSlots, bar_act = bar_acc + 16;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SM_TMEMPTR);
  const float* cst = reinterpret_cast<const float*>(p.packed + kFwdBytes + kBwdBytes);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) mbar_init(bar_empty + 8 * s, 1);
    for (int b = 0; b < 8; ++b) mbar_init(bar_full + 8 * b, rank == 0 ? 2 : 1);
    for (int t = 0; t < 2; ++t) { mbar_init(bar_acc + 8 * t, 1); mbar_init(bar_act + 8 * t, 2 * kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 1) {   // TMEM: 512 columns = two 128x256 fp32 accumulators, in both CTAs of the pair
    tmem_alloc2(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish2();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 256) {   // sigma-head weights as bf16, read by the layer-7 epilogues
    const int i = threadIdx.x - 64;
    const __nv_bfloat16 w = __float2bfloat16_rn(__ldg(cst + C_WA + i));
    sts16(sbase + SM_WA + 2 * i, *reinterpret_cast<const uint16_t*>(&w));
  }
  tcgen05_fence_before_sync();
  cluster_sync_all();               // barriers of both CTAs initialised before anyone signals across the pair
  tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int cid = (int)blockIdx.x >> 1, ncl = (int)gridDim.x >> 1;
  const int my_rounds = (p.num_quads - cid + ncl - 1) / ncl;
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;

  if (warp == 0) {
    // ================= weight producer (this CTA's half of every chunk) =================
    // One 16 KB (8 KB for the 128-wide views layer) bulk copy per chunk; the two chunks of a group are issued by two
    // different lanes (slot, slot + 3): a thread retires at most one cp.async.bulk per ~700 cycles (tools/bulk_rate.py).
    uint32_t stage = 0, phase = 0;
    for (int it = 0; it < my_rounds; ++it) {
      const uint8_t* src = p.packed;
      for (int s = 0; s < kNumSteps; ++s) {
        const uint32_t half = (uint32_t)c_step_n[s] * 64u;       // bytes of this CTA's half chunk
        const int nch = c_step_chunks[s];
        for (int g = 0; 2 * g < nch; ++g) {
          const int in_group = nch - 2 * g < 2 ? nch - 2 * g : 2;
          const int sub = lane >= kSlots ? 1 : 0;                // which chunk of the group this lane copies
          if (lane == (int)stage || lane == (int)stage + kSlots) {
            const uint32_t gbar = bar_full + 8 * (s & 3) + 32 * g;
            // BOTH lanes of the slot wait for every one of its releases, also when the group has a single chunk: a lane
            // that skipped a revolution would find its next parity wait satisfied by the release before last
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            if (sub == 0) {
              mbar_arrive_expect_tx(gbar, (uint32_t)in_group * half);
              if (g == 0 && nch <= 2) mbar_arrive(gbar + 32);   // no second group: its barrier still advances one phase per step
              if (tracing) trace_stamp(p.trace, it, s, 0, 12 + g);
            }
            const int c = 2 * g + sub;
            if (sub < in_group)
              bulk_g2s(sbase + SM_RING + stage * kSlotBytes + sub * (kSlotBytes / 2), src + (size_t)c * 2 * half + (size_t)rank * half,
                       half, gbar);
          }
          if (++stage == kSlots) { stage = 0; phase ^= 1; }
        }
        src += (size_t)nch * 2 * half;
      }
    }
  } else if (warp == 1 && rank != 0) {
    // ================= peer CTA: relay "my half landed" to the leader =================
    if (lane == 0) {
      const uint32_t full_leader = mapa_cluster(bar_full, 0);
      for (int it = 0; it < my_rounds; ++it)
        for (int s = 0; s < kNumSteps; ++s) {
          const uint32_t lph = (uint32_t)(it * (kNumSteps / 4) + (s >> 2)) & 1u;
          for (int g = 0; g < 2; ++g) {
            mbar_wait(bar_full + 32 * g + 8 * (s & 3), lph);
            mbar_arrive_cluster(full_leader + 32 * g + 8 * (s & 3));
          }
        }
    }
  } else if (warp == 1) {
    // ================= leader CTA: MMA issuer for the pair =================
    if (lane == 0) {
      uint32_t grp = 0;                                     // groups consumed so far: group i sits in ring slot i % 3
      uint32_t act_phase[2] = {0, 0};
      for (int it = 0; it < my_rounds; ++it) {
        for (int s = 0; s < kNumSteps; ++s) {
          const int nch = c_step_chunks[s], n = c_step_n[s], ksteps = c_step_ksteps[s];
          const uint32_t idesc = make_idesc(2 * kTileM, n, 0, 0);
          const uint32_t lph = (uint32_t)(it * (kNumSteps / 4) + (s >> 2)) & 1u;
          const uint32_t gbar = bar_full + 8 * (s & 3);
          mbar_wait_cluster(gbar, lph);                       // group 0 of both CTAs (prefetched a layer ahead)
          if (tracing) trace_stamp(p.trace, it, s, 0, 1);
          for (int t = 0; t < 2; ++t) {
            mbar_wait_cluster(bar_act + 8 * t, act_phase[t]);   // A operands of both CTAs written, accumulators drained
            act_phase[t] ^= 1;
            tcgen05_fence_after_sync();
            if (tracing) trace_stamp(p.trace, it, s, t, 0);
            const uint32_t d_tmem = tmem_base + (uint32_t)t * 256u;
            uint32_t accumulate = (uint32_t)c_step_acc[s];
            for (int c = 0; c < nch; ++c) {
              const uint32_t stage = (grp + (uint32_t)(c >> 1)) % kSlots;
              if (t == 0 && c == 2) {                           // group 1: landed while group 0 ran (tile slot 1 follows slot 0)
                mbar_wait_cluster(gbar + 32, lph);
                tcgen05_fence_after_sync();
                if (tracing) trace_stamp(p.trace, it, s, 0, 2);
              }
              if (tracing) trace_stamp(p.trace, it, s, t, 8 + c);
              const uint32_t a_addr = sbase + SM_ACT + t * kActBytes + (nch == 1 ? 0 : c) * kAtomBytes;
              const uint32_t b_addr = sbase + SM_RING + stage * kSlotBytes + (c & 1) * (kSlotBytes / 2);
              const uint64_t a_desc = make_smem_desc(a_addr, 16, 1024);
              const uint64_t b_desc = make_smem_desc(b_addr, 16, 1024);
              for (int k = 0; k < ksteps; ++k) {
                // +32 bytes (16 bf16) along K inside the 128-byte swizzle atom: start-address field += 2
                umma_bf16_2cta(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, accumulate);
                accumulate = 1;
              }
              // group free in both CTAs once tile slot 1's MMAs have read it
              if (t == 1 && ((c & 1) || c == nch - 1)) umma_commit_2cta(bar_empty + 8 * stage, 3);
            }
            umma_commit_2cta(bar_acc + 8 * t, 3);            // accumulators complete -> epilogues of slot t in both CTAs
            if (tracing) trace_stamp(p.trace, it, s, t, 3);
          }
          grp += (uint32_t)((nch + 1) >> 1);
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
    const uint32_t act_leader = mapa_cluster(bar_act + 8 * t, 0);    // the leader's "A tile written" barrier for slot t
    uint32_t acc_phase = 0;
    for (int it = 0; it < my_rounds; ++it) {
      const int64_t tile = 4 * ((int64_t)cid + (int64_t)it * ncl) + 2 * (int64_t)rank + t;
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
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(act_leader);
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
        if (tix == 0 && tracing) trace_stamp(p.trace, it, s, t, 4);
        if (epi == EPI_WRITE_ENC || epi == EPI_WRITE_DENC) {
          // pass 1 has finished reading the tile: overwrite atom 0 with the second-pass operand
          if (epi == EPI_WRITE_ENC) store_enc_chunks<16, 4, false>(encw, 4 * cg, live, act_a, nullptr, r);
          else if (cg == 1) store_enc_chunks<16, 8, kTrain>(encw, 0, live, act_a, stash_tile + (size_t)SA_DENC * kAtomBytes, r);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(act_leader);
          if (tix == 0 && tracing) trace_stamp(p.trace, it, s, t, 6);
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
          if (tix == 0 && tracing) trace_stamp(p.trace, it, s, t, 6);
          continue;   // next arrival on act_ready comes from the next tile's prologue
        }
        // ---- bias (+ReLU) -> bf16 -> swizzled in-place store; layer 7 also accumulates sigma from fp32 h7
        uint32_t mk[4];
        const uint32_t my_bias = bias_row + (uint32_t)cg * 512u;
        const uint32_t my_tmem = tmem_lane + (uint32_t)cg * 128u;
        long long* etr = (tix == 0 && tracing && it < kTraceRounds) ? p.trace + ((it * 12 + s) * 2 + t) * kTraceEvents : nullptr;
        if (epi == EPI_RELU) epi_layer<kTrain, 0>(my_tmem, my_bias, wa_a, alpha, row_a, rx, mk, etr);
        else if (epi == EPI_RELU_ALPHA) epi_layer<kTrain, 1>(my_tmem, my_bias, wa_a, alpha, row_a, rx, mk, etr);
        else epi_layer<kTrain, 2>(my_tmem, my_bias, wa_a, alpha, row_a, rx, mk, etr);
        tcgen05_fence_before_sync();
        if (etr) etr[22] = clock64();
        fence_proxy_async_smem();
        if (etr) etr[23] = clock64();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(act_leader);
        if (tix == 0 && tracing) trace_stamp(p.trace, it, s, t, 6);
        // the mask words go out AFTER the hand-over: a global store in front of it sat ~900 cycles on the critical path
        if (kTrain && c_step_mask_slot[s] >= 0) {
          uint32_t* mrow = reinterpret_cast<uint32_t*>(stash_tile + kStashMaskOff) + (c_step_mask_slot[s] * 128 + r) * 8 + cg * 4;
          *reinterpret_cast<uint4*>(mrow) = make_uint4(mk[0], mk[1], mk[2], mk[3]);
        }
        if (kTrain) pending_atom = c_step_stash_atom[s];
      }
    }
    if (store_lane) bulk_wait0();   // every stash store has landed before the kernel exits
  }

  __syncwarp();                     // single-lane roles rejoin their warp before the aligned cluster barrier
  tcgen05_fence_before_sync();
  cluster_sync_all();               // nobody signals into a CTA that has exited; all MMAs of the pair have completed
  if (warp == 1) {
    tcgen05_fence_after_sync();
    tmem_dealloc2(tmem_base, 512);
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
  p.num_quads = (int)((tiles + 3) / 4);
  p.trace = g_trace;
  const int pairs = sm_count() / 2;
  int grid = 2 * (p.num_quads < pairs ? p.num_quads : pairs);
  auto kern = stash ? mlp_fwd_kernel<true> : mlp_fwd_kernel<false>;
  static bool attr_set[2] = {false, false};
  if (!attr_set[stash ? 1 : 0]) {
    SPN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set[stash ? 1 : 0] = true;
  }
  prof_begin(PROF_MLP_FWD, st);
  kern<<<grid, kPairThreads, kSmemBytes, st>>>(p);
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


// ---- diagnostic: TMEM -> register drain rate (tcgen05.ld 32x32b.x32), optionally while MMAs run on the same SM ---------
// warps 0..nwarps-1 (warp w reads lane quarter w % 4) each drain 128 columns `reps` times; with_mma != 0: warp 16 lane 0
// keeps issuing M=128 N=256 K=16 MMAs into columns 256..511 for the whole time.  out[0] = cycles of the slowest drain
// warp, out[1] = MMAs retired meanwhile.
__global__ void __launch_bounds__(17 * 32, 1) tmem_ld_rate_kernel(int nwarps, int reps, int with_mma, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  for (int i = threadIdx.x; i < (64 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  const uint32_t bar = sbase + 64 * 1024;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 64 * 1024 + 64);
  volatile int* stop = reinterpret_cast<volatile int*>(smem + 64 * 1024 + 128);
  __shared__ long long s_cyc[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); *stop = 0; }
  if (warp == 0) { tmem_alloc(smem_u32(tptr), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tcgen05_fence_before_sync();
  __syncthreads();
  tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tptr;
  if (warp < nwarps) {
    const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 1) * 128u;
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      uint32_t va[32], vb[32];
      tmem_ld32(ta, va); tmem_ld32(ta + 32, vb);
      tmem_ld_wait_dep(va); tmem_ld_wait_dep(vb);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= va[j] + vb[j];
      tmem_ld32(ta + 64, va); tmem_ld32(ta + 96, vb);
      tmem_ld_wait_dep(va); tmem_ld_wait_dep(vb);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= va[j] + vb[j];
    }
    const long long t1 = clock64();
    if (lane == 0) s_cyc[warp] = t1 - t0 + (acc == 0x12345u ? 1 : 0);
    if (warp == 0 && lane == 0) *stop = 1;
  } else if (warp == 16 && lane == 0 && with_mma) {
    const uint32_t idesc = make_idesc(128, 256, 0, 0);
    const uint64_t a0 = make_smem_desc(sbase, 16, 1024), b0 = make_smem_desc(sbase + 16384, 16, 1024);
    long long n = 0;
    while (!*stop) {
      for (int i = 0; i < 16; ++i) umma_bf16(tmem_base + 256u, a0 + (uint64_t)(2 * (i & 3)), b0 + (uint64_t)(2 * (i & 3)), idesc, 1);
      n += 16;
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    out[1] = n;
  }
  tcgen05_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) {
    long long m = 0;
    for (int w = 0; w < nwarps; ++w) m = s_cyc[w] > m ? s_cyc[w] : m;
    out[0] = m;
    if (!with_mma) out[1] = 0;
  }
  if (warp == 0) { tcgen05_fence_after_sync(); tmem_dealloc(tmem_base, 512); }
}

int tc_tmem_ld_rate(int nwarps, int reps, int with_mma, long long* out, cudaStream_t st) {
  int rc = check_arch();
  if (rc != SPN_OK) return rc;
  SPN_CHECK_ARG(out && nwarps >= 1 && nwarps <= 16 && reps > 0, "spn_tc_tmem_ld_rate: bad arguments");
  const int smem_bytes = 64 * 1024 + 256 + 1024;
  SPN_CUDA(cudaFuncSetAttribute(tmem_ld_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  tmem_ld_rate_kernel<<<1, 17 * 32, smem_bytes, st>>>(nwarps, reps, with_mma, out);
  SPN_LAUNCH_CHECK("tmem_ld_rate_kernel");
  return SPN_OK;
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
