// NeRF MLP (DS_NeRF/run_nerf_helpers.py:74-127) fused with sampling-point generation and positional
// encoding (run_nerf.py:56-71, 670; helpers:22-70) on the 5th-generation tensor cores (tcgen05).
//
// Persistent CTA PAIRS (thread-block clusters of 2 on neighbouring SMs, cta_group::2 MMAs), 576 threads per CTA:
//   warp 0      weight producer: this CTA's half (its 128 of 256 output rows) of every pre-swizzled bf16 weight chunk,
//               cp.async.bulk into a 4-slot ring of two-chunk groups; one more lane streams the stash atoms out
//   warp 1      leader CTA: issues the pair's tcgen05.mma (M = 256, N = 128, K = 16; A from TMEM, B from the ring) in uniform
//               control flow; peer CTA: relays "my weight halves have landed" to the leader
//   warps 2-17  prologue (sample point -> gamma(pts)) and epilogues: accumulator half -> bias, ReLU, packed bf16 back into
//               TENSOR MEMORY as the next layer's A operand (tcgen05.st); in training also E4M3 -> staging -> stash, ReLU masks
// Activations never leave the SM and never touch shared memory; HBM sees 24 B in + 16 B out per sample (plus the stash in
// training).  See the comment block above mlp_fwd_ts_kernel for the TMEM budget, the batch order and the barrier protocol.
//
// The 63-wide skip input of layer 5 and the 27-wide view encoding of the views layer are extra K = 64 / K = 32 MMAs of the
// same batch whose A operand is the slot's encoding atom in shared memory (written once per round; gamma(dir) replaces
// gamma(pts) after the skip layer).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "mlp_tc.cuh"

namespace spn {
using namespace tc;

size_t mlp_tc_packed_bytes() { return kPackedBytes; }

// diagnostic timeline buffer (device pointer, kTraceSlots int64): see tools/trace_fwd_ts.py
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
  int trace_block;  // SPN_TRACE_BLOCK: whose epilogue stamps are recorded (0 = leader of pair 0, 1 = its peer); MMA stamps always come from block 0
  int debug;        // SPN_FWD_DEBUG (timing experiments, results are wrong): 1 = no wait for the stash copies, 2 = no mask stores, 4 = no stash copies
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

// values [I0, I0 + CNT) of [v, sin(2^k v), cos(2^k v)]_k (helpers:28-52 order; entries past 3 + 6 NFREQ are zero padding) as
// CNT / 2 packed bf16 pairs: one quarter of gamma(pts) (L = 10, 16 values) or of gamma(dir) (L = 4, 8 values) per thread
template <int NFREQ, int I0, int CNT>
__device__ __forceinline__ void encode_range(const float v[3], uint32_t (&w)[CNT / 2]) {
  float vals[CNT];
#pragma unroll
  for (int i = 0; i < CNT; ++i) vals[i] = 0.0f;
#pragma unroll
  for (int a = 0; a < 3; ++a)
    if (a >= I0 && a < I0 + CNT) vals[(a - I0) % CNT] = v[a];
#pragma unroll
  for (int k = 0; k < NFREQ; ++k) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int is = 3 + 6 * k + a, ic = is + 3;
      const bool ns = is >= I0 && is < I0 + CNT, nc = ic >= I0 && ic < I0 + CNT;
      if (ns || nc) {
        float sn, cs;
        fast_sincos(__fmul_rn(v[a], (float)(1 << k)), sn, cs);
        if (ns) vals[(is - I0) % CNT] = sn;
        if (nc) vals[(ic - I0) % CNT] = cs;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < CNT / 2; ++i) w[i] = pack_bf16(vals[2 * i], vals[2 * i + 1]);
}
__device__ __forceinline__ void encode_pts_quarter(const int cq, const float v[3], uint32_t (&w)[8]) {
  if (cq == 0) encode_range<10, 0, 16>(v, w);
  else if (cq == 1) encode_range<10, 16, 16>(v, w);
  else if (cq == 2) encode_range<10, 32, 16>(v, w);
  else encode_range<10, 48, 16>(v, w);
}
__device__ __forceinline__ void encode_dir_quarter(const int cq, const float v[3], uint32_t (&w)[4]) {
  if (cq == 0) encode_range<4, 0, 8>(v, w);
  else if (cq == 1) encode_range<4, 8, 8>(v, w);
  else if (cq == 2) encode_range<4, 16, 8>(v, w);
  else encode_range<4, 24, 8>(v, w);
}

// =====================================================================================================================
// "TS" forward kernel: activations never touch shared memory.
//
// What bounded the kernel above (profiles/r02b_trace_*.txt): with both operands in shared memory every MMA reads 8 KB per SM
// (A 4 KB + its half of B 4 KB = 64 B/cycle), the epilogue writes the next A tile back (64 KB per tile and layer) and reads the
// bias row, the ring is refilled and — in training — the TMA store reads the tile once more: ~2050 (inference) / ~2560 (training)
// cycles of the 128 B/cycle shared-memory pipe per tile and layer against 2048 cycles of MMA, so the MMAs ran at ~178 instead of
// 128 cycles however the epilogue was organised.  Here the A operand lives in TENSOR MEMORY (tcgen05.mma with A in TMEM): the
// epilogue writes relu(acc + b) as packed bf16 straight back to TMEM with tcgen05.st and the next layer's MMAs read it from
// there.  Shared memory carries only the weights (B: 32 B/cycle), the 64-wide encodings and, in training, the staging of the
// stash copies.  TMEM budget (512 columns): two 128x128 fp32 accumulators X, Y (N is split in halves so that an accumulator
// half drains while the other half's MMAs run) + the two tile slots' activations A0, A1 (256 bf16 per row = 128 columns each).
//
// Per 256-wide layer the tensor pipe runs four batches back to back: (slot 0, half a) -> X, (0, b) -> Y, (1, a) -> X, (1, b) -> Y,
// each 16 MMAs of M = 256 (pair) x N = 128 x K = 16.  All 16 epilogue warps follow the same order:
//   acc_full[X]: tcgen05.ld 32 columns -> arrive acc_free[X] at once (the accumulator is in registers) -> bias/ReLU/pack -> keep
//   acc_full[Y] (all MMAs that read the slot's old activations are done): tcgen05.st both halves into A_t -> arrive a_ready[t]
// so every hand-over has a full batch (1024 cycles) of slack.  With the round-1 weight packing (each CTA holds output rows
// 128c..128c+127 of a chunk) half h takes rows 64h..64h+63 of both CTAs: accumulator column j of half h is output feature
// 64h + (j & 63) + 128 (j >> 6) — a permutation the epilogue undoes when it picks bias, TMEM columns and stash atoms.
// =====================================================================================================================
constexpr int kTsSlots = 4;                                  // ring: 4 groups of two 16 KB half-chunks = 2 layers resident
constexpr int TS_RING = 0;
constexpr int TS_GAMMA = kTsSlots * kSlotBytes;              // 2 x 16 KB: gamma(pts) / gamma(dir) atom of each tile slot (SS operand)
constexpr int TS_STAGE = TS_GAMMA + 2 * kAtomBytes;          // training: 4 x 16 KB staging of a half layer's E4M3 stash atom (buffer 2 t + h)
constexpr int TS_BAR = TS_STAGE + 4 * kAtomBytes;            // = 229376
constexpr int TS_TMEMPTR = TS_BAR + 240;
constexpr int TS_BIAS = TS_BAR + 256;                        // 2 rows of 256 floats (layer parity)
constexpr int TS_WA = TS_BIAS + 2048;                        // sigma-head weights, 256 x bf16
constexpr int kTsSmemBytes = TS_WA + 512;                    // 232192
static_assert(kTsSmemBytes <= 232448, "shared memory budget");
constexpr int kTsLayers = 10;
__constant__ int c_l_step0[kTsLayers] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 10};     // first step (mlp step table) of each layer
__constant__ int c_l_parts[kTsLayers] = {1, 1, 1, 1, 1, 2, 1, 1, 1, 2};      // layer 5 = h part + gamma(pts) part, views = feature part + gamma(dir) part
// weight groups of a round in ring order (two 64-wide K chunks each); three empty groups pad the sequence to 6 revolutions of
// the 4-slot ring so that every layer starts in slot 0 or 2:  L0 | - | L1 L1 | L2 L2 | L3 L3 | L4 L4 | L5 L5 | L5g | - | L6 L6 |
// L7 L7 | feat feat | views views | views-dir | -
constexpr int kTsGroups = 24;
__constant__ int c_g_nch[kTsGroups] = {1, 0, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 0, 2, 2, 2, 2, 2, 2, 2, 2, 1, 0};
__constant__ int c_g_half[kTsGroups] = {16384, 0, 16384, 16384, 16384, 16384, 16384, 16384, 16384, 16384, 16384, 16384, 16384, 0,
                                        16384, 16384, 16384, 16384, 16384, 16384, 8192, 8192, 8192, 0};
enum TsLayerKind : int { LK_L0 = 0, LK_WIDE = 1, LK_L5 = 2, LK_VIEWS = 3 };
__constant__ int c_l_kind[kTsLayers] = {LK_L0, LK_WIDE, LK_WIDE, LK_WIDE, LK_WIDE, LK_L5, LK_WIDE, LK_WIDE, LK_WIDE, LK_VIEWS};
__constant__ int c_l_group0[kTsLayers] = {0, 2, 4, 6, 8, 10, 14, 16, 18, 20};   // first group of the layer in the round's sequence

// 32 accumulator columns = 32 consecutive output features: h = acc + bias (ReLU), packed bf16 -> out[16] (the TMEM image of the
// next A operand) and, in training, E4M3 -> the staging tile of the stash copy (32 bytes: chunks chunk0, chunk0 + 1 of the row);
// returns the 32 ReLU mask bits (relu_mask_push layout).  MODE 0: ReLU, 1: ReLU + sigma-head partial from the fp32 h, 2: linear.
template <bool kTrain, int MODE>
__device__ __forceinline__ uint32_t epi32_ts(const uint32_t (&v)[32], const uint32_t bias_a, const uint32_t wa_a, float& alpha,
                                             uint32_t (&out)[16], const uint32_t stage_row, const uint32_t chunk0, const uint32_t rx) {
  uint32_t mb = 0;
  uint32_t q8[4];
  float4 nb0 = lds128f(bias_a), nb1 = lds128f(bias_a + 16);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int c = g * 8;
    const float4 b0 = nb0, b1 = nb1;
    if (g < 3) { nb0 = lds128f(bias_a + (c + 8) * 4); nb1 = lds128f(bias_a + (c + 8) * 4 + 16); }
    const float2 h01 = __fadd2_rn(make_float2(__uint_as_float(v[c + 0]), __uint_as_float(v[c + 1])), make_float2(b0.x, b0.y));
    const float2 h23 = __fadd2_rn(make_float2(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])), make_float2(b0.z, b0.w));
    const float2 h45 = __fadd2_rn(make_float2(__uint_as_float(v[c + 4]), __uint_as_float(v[c + 5])), make_float2(b1.x, b1.y));
    const float2 h67 = __fadd2_rn(make_float2(__uint_as_float(v[c + 6]), __uint_as_float(v[c + 7])), make_float2(b1.z, b1.w));
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
#pragma unroll
    for (int e = 0; e < 4; ++e)
      out[4 * g + e] = (MODE == 0) ? pack_relu_bf16(h[2 * e], h[2 * e + 1]) : pack_bf16(h[2 * e], h[2 * e + 1]);
    if (kTrain) {
      if (MODE != 2) {
#pragma unroll
        for (int e = 0; e < 4; ++e) mb = relu_mask_push(mb, out[4 * g + e]);
      }
      // the stash copy is E4M3 of the fp32 value (MODE 0: the ReLU rides on the conversion)
      const uint32_t w0 = (MODE == 0) ? pack_relu_e4m3x4(h[0], h[1], h[2], h[3]) : pack_e4m3x4(h[0], h[1], h[2], h[3]);
      const uint32_t w1 = (MODE == 0) ? pack_relu_e4m3x4(h[4], h[5], h[6], h[7]) : pack_e4m3x4(h[4], h[5], h[6], h[7]);
      q8[(2 * g) & 3] = w0; q8[(2 * g + 1) & 3] = w1;
      if (g & 1) sts128(stage_row + (((chunk0 + (uint32_t)(g >> 1)) << 4) ^ rx), q8[0], q8[1], q8[2], q8[3]);
    }
  }
  return mb;
}

template <bool kTrain>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1) mlp_fwd_ts_kernel(const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;   // warp index the compiler can prove uniform
  const uint32_t rank = uniform_u32(cluster_ctarank());   // 0: leader (issues the pair's MMAs)
  if (smem != smem_raw) __trap();                   // no alignment slack in kTsSmemBytes
  // Warp 0: weight producer / stash store lane, warp 1: MMA issuer, warps 2-17: epilogue.  The issuer must NOT be the highest
  // warp id of its scheduler: the arbiter prefers the highest eligible warp id (B300_MICROARCH.md), and an issuer that polls
  // its barriers from warp 17 can starve the very epilogue warps it is waiting for — with that mapping 8-GPU render / LPIPS
  // runs died with launch failures on single ranks (profiles/r02g_n8_*), while it bought nothing on the timeline.
  constexpr int kProdWarp = 0, kMmaWarp = 1;
  // barriers: full[4] / empty[4] per ring slot (slot s is used once per revolution: parity = revolution & 1),
  //   acc_full[2] (commit after each batch), acc_free[2] (32 warps: accumulator loaded into registers, leader only),
  //   a_ready[2] (32 warps: slot t's activations / gamma atom written, leader only),
  //   written[4] / free[4] (training: staging buffer 2 t + h filled by the 16 warps / copied out by its store lane)
  const uint32_t bar_full = sbase + TS_BAR, bar_empty = bar_full + 8 * kTsSlots;
  const uint32_t bar_accfull = bar_empty + 8 * kTsSlots, bar_accfree = bar_accfull + 16, bar_aready = bar_accfree + 16;
  const uint32_t bar_written = bar_aready + 16, bar_free = bar_written + 32;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + TS_TMEMPTR);
  const float* cst = reinterpret_cast<const float*>(p.packed + kFwdBytes + kBwdBytes);

  if (threadIdx.x == 0) {
    // full: both producer lanes of the slot (+ the peer's relay on the leader)
    for (int s = 0; s < kTsSlots; ++s) { mbar_init(bar_empty + 8 * s, 1); mbar_init(bar_full + 8 * s, rank == 0 ? 3 : 2); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_accfull + 8 * t, 1);
      mbar_init(bar_accfree + 8 * t, 2 * kFwdEpiWarps);
      mbar_init(bar_aready + 8 * t, 2 * kFwdEpiWarps);
    }
    for (int b = 0; b < 4; ++b) { mbar_init(bar_written + 8 * b, kFwdEpiWarps); mbar_init(bar_free + 8 * b, 1); }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) { tmem_alloc2(smem_u32(tmem_ptr_smem), 512); tmem_relinquish2(); }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 256) {   // sigma-head weights as bf16, read by the layer-7 epilogues
    const int i = threadIdx.x - 64;
    const __nv_bfloat16 w = __float2bfloat16_rn(__ldg(cst + C_WA + i));
    sts16(sbase + TS_WA + 2 * i, *reinterpret_cast<const uint16_t*>(&w));
  }
  tcgen05_fence_before_sync();
  cluster_sync_all();
  tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  constexpr uint32_t kAccCols = 128, kA0 = 256;     // X at +0, Y at +128, A_t at +256 + 128 t

  const int cid = (int)blockIdx.x >> 1, ncl = (int)gridDim.x >> 1;
  const int my_rounds = (p.num_quads - cid + ncl - 1) / ncl;
  const bool tracing = p.trace != nullptr && (int)blockIdx.x == (warp >= 2 ? p.trace_block : 0);

  if (warp == kProdWarp) {
    if (lane < 2 * kTsSlots) {
      // ================= weight producer: this CTA's half of every chunk, groups of two chunks per ring slot =================
      // The group sequence of a round is padded to 24 (c_g_nch: three empty groups), so every group sits in a ring slot that is
      // known at compile time (the MMA issuer addresses the ring with immediates) and slot s is used once per revolution.
      // Two lanes per slot (a thread retires one cp.async.bulk per ~700 cycles): lane (slot, 0) arms the group's byte count and
      // copies chunk 0, lane (slot, 1) copies chunk 1.  BOTH take part in EVERY phase of the slot's two barriers — both wait
      // for the release, both arrive on the full barrier, also for one-chunk and empty groups — and the issuer waits for every
      // group's full barrier, also an empty group's, before it releases the slot: the two barriers then alternate strictly.
      // Round 2's first version let lane (slot, 1) sit out one-chunk / empty groups and released empty groups unseen; a lane
      // that was not scheduled for the ~2 us between two consecutive releases of its slot then waited for a parity that had
      // already come round again: one deadlock per ~1000 train steps (profiles/r02_summary.md).
      const int slot = lane & 3, sub = lane >> 2;
      for (int it = 0; it < my_rounds; ++it) {
        const uint8_t* src = p.packed;
        for (int g = 0; g < kTsGroups; ++g) {
          const int nch = c_g_nch[g];
          const uint32_t half = (uint32_t)c_g_half[g];
          if ((g & 3) == slot) {
            const uint32_t rev = (uint32_t)(it * (kTsGroups / 4) + (g >> 2));
            mbar_wait(bar_empty + 8 * slot, (rev & 1u) ^ 1u);
            if (sub == 0 && nch) mbar_arrive_expect_tx(bar_full + 8 * slot, (uint32_t)nch * half);
            else mbar_arrive(bar_full + 8 * slot);
            if (sub < nch)
              bulk_g2s(sbase + TS_RING + slot * kSlotBytes + sub * (kSlotBytes / 2), src + (size_t)sub * 2 * half + (size_t)rank * half, half,
                       bar_full + 8 * slot);
          }
          src += (size_t)nch * 2 * half;
        }
      }
    } else if (kTrain && lane == 8) {
      // ================= stash store lane: copies the E4M3 atom of every half epilogue (staging buffer 2 t + h) to the stash =================
      // One lane, buffers in the order the epilogue fills them (0, 1, 2, 3, 0, ...); bulk groups complete their reads in order,
      // so after committing copy k "at most one group pending" means copy k - 1 has finished reading its buffer.
      uint32_t k = 0;
      for (int it = 0; it < my_rounds; ++it) {
        const int64_t tile0 = 4 * ((int64_t)cid + (int64_t)it * ncl) + 2 * (int64_t)rank;
        for (int L = 0; L < kTsLayers - 1; ++L) {
          for (int b = 0; b < 4; ++b, ++k) {
            mbar_wait(bar_written + 8 * b, (k >> 2) & 1u);
            if (!(p.debug & 4)) {
              bulk_s2g(p.stash + (size_t)(tile0 + (b >> 1)) * kStashTileBytes + (size_t)stash_x_atom(L, b & 1) * kAtomBytes,
                       sbase + TS_STAGE + (uint32_t)b * kAtomBytes, kAtomBytes);
              bulk_commit();
              asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            }
            if (k > 0) mbar_arrive(bar_free + 8 * ((b + 3) & 3));      // buffer of copy k - 1
          }
        }
      }
      bulk_wait_read0();
      if (k > 0) mbar_arrive(bar_free + 8 * 3);
      bulk_wait0();
    }
  } else if (warp == kMmaWarp && rank != 0) {
    if (lane == 0) {   // ================= peer CTA: relay "my halves landed" to the leader =================
      const uint32_t full_leader = mapa_cluster(bar_full, 0);
      for (int it = 0; it < my_rounds; ++it)
        for (int g = 0; g < kTsGroups; ++g) {
          mbar_wait(bar_full + 8 * (g & 3), (uint32_t)(it * (kTsGroups / 4) + (g >> 2)) & 1u);
          mbar_arrive_cluster(full_leader + 8 * (g & 3));
        }
    }
  } else if (warp == kMmaWarp) {
    // ================= leader CTA: MMA issuer for the pair — the whole warp, uniform control flow (tc_common.cuh: elect_one) =================
    // Ring slots, chunk offsets and TMEM columns are immediates relative to three uniform bases (ring descriptor, gamma-atom
    // descriptor, TMEM base): an MMA costs two uniform adds.
    uint32_t batch = 0, ar_ph = 0, af_n0 = 0, af_n1 = 0;
    const uint32_t idesc = make_idesc(2 * kTileM, 128, 0, 0);
    const uint32_t tmem_u = uniform_u32(tmem_base);
    const uint64_t ring_desc = make_smem_desc(sbase + TS_RING, 16, 1024);     // slot s: + 2048 s, chunk c of a group: + 1024 c, half h: + 512 h
    const uint64_t gam_desc = make_smem_desc(sbase + TS_GAMMA, 16, 1024);     // tile slot t: + 1024 t
    auto full_wait = [&](const int slot, const uint32_t rev) {
      mbar_wait_cluster(bar_full + 8 * slot, rev & 1u);
      tcgen05_fence_after_sync();
    };
    auto empty_commit = [&](const int slot) {
      if (elect_one()) umma_commit_2cta(bar_empty + 8 * slot, 3);
    };
    for (int it = 0; it < my_rounds; ++it) {
      const uint32_t rev0 = (uint32_t)(it * (kTsGroups / 4));
#pragma unroll 1
      for (int L = 0; L < kTsLayers; ++L) {
        const int kind = c_l_kind[L];
        const int g0 = c_l_group0[L];
        const int s0 = g0 & 3;                               // first ring slot of the layer (0 or 2)
        const uint32_t rev = rev0 + (uint32_t)(g0 >> 2);
        const int nb = kind == LK_VIEWS ? 2 : 4;
        full_wait(s0, rev);                                  // the layer's first group, prefetched a layer ahead
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
          const int t = kind == LK_VIEWS ? b : (b >> 1), h = kind == LK_VIEWS ? 0 : (b & 1);
          const bool first = b == 0, last = b == nb - 1;
          const uint32_t acc = batch & 1u;
          ++batch;
          if (h == 0) {                                      // slot t's activations (and gamma atom) of both CTAs are written
            mbar_wait_cluster(bar_aready + 8 * t, (ar_ph >> t) & 1u);
            ar_ph ^= 1u << t;
          }
          {                                                  // accumulator `acc` is in the epilogue's registers
            const uint32_t n = acc ? af_n1 : af_n0;
            mbar_wait_cluster(bar_accfree + 8 * acc, (n & 1u) ^ 1u);
            if (acc) ++af_n1; else ++af_n0;
          }
          tcgen05_fence_after_sync();
          if (tracing && lane == 0) trace_stamp(p.trace, it, c_l_step0[L] + c_l_parts[L] - 1, t, h);
          const uint32_t d_tmem = tmem_u + acc * kAccCols;
          const uint32_t a_tmem = tmem_u + kA0 + (uint32_t)t * 128u;
          const uint64_t a_gam = gam_desc + (uint64_t)(1024 * t);
          const uint64_t b_desc = ring_desc + (uint64_t)(2048 * s0 + 512 * h);
          if (kind == LK_L0) {
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16_2cta(d_tmem, a_gam + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
            }
            if (last) { full_wait(1, rev); empty_commit(0); empty_commit(1); }       // slot 1: the empty group behind layer 0
          } else {
            // four 64-wide K chunks from the slot's activations in TMEM: two ring slots of two chunks
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                umma_bf16_ts_2cta(d_tmem, a_tmem + (uint32_t)(8 * j), b_desc + (uint64_t)(1024 * (j >> 2) + 2 * (j & 3)), idesc, j ? 1u : 0u);
            }
            if (last) empty_commit(s0);
            if (first) full_wait(s0 + 1, rev);
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                umma_bf16_ts_2cta(d_tmem, a_tmem + (uint32_t)(64 + 8 * j), b_desc + (uint64_t)(2048 + 1024 * (j >> 2) + 2 * (j & 3)), idesc, 1u);
            }
            if (last) empty_commit(s0 + 1);
            if (kind == LK_L5) {                             // + gamma(pts) part of the skip layer: ring slot 0 of the next revolution
              if (first) full_wait(0, rev + 1);
              const uint64_t bg = ring_desc + (uint64_t)(512 * h);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_2cta(d_tmem, a_gam + (uint64_t)(2 * k), bg + (uint64_t)(2 * k), idesc, 1u);
              }
              if (last) { full_wait(1, rev + 1); empty_commit(0); empty_commit(1); }   // slot 1: the empty group behind the skip layer
            } else if (kind == LK_VIEWS) {                   // + gamma(dir) part (K = 32): ring slot 2
              if (first) full_wait(2, rev);
              const uint64_t bg = ring_desc + (uint64_t)(2048 * 2);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 2; ++k) umma_bf16_2cta(d_tmem, a_gam + (uint64_t)(2 * k), bg + (uint64_t)(2 * k), idesc, 1u);
              }
              if (last) { full_wait(3, rev); empty_commit(2); empty_commit(3); }       // slot 3: the empty group that ends the round
            }
          }
          if (elect_one()) umma_commit_2cta(bar_accfull + 8 * acc, 3);
          if (tracing && lane == 0) trace_stamp(p.trace, it, c_l_step0[L] + c_l_parts[L] - 1, t, 2 + h);
        }
      }
    }
  } else {
    // ================= prologue + epilogue warps (all 16 follow the batch order of the tensor pipe) =================
    // warp (q, cq): TMEM lane quarter q = warp % 4 (rows 32q..32q+31), accumulator columns 32cq..32cq+31 of every half
    const int ew = warp - 2;
    const int cq = ew >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int tix = cq * 128 + r;
    const uint32_t rx = (uint32_t)(r & 7) << 4;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t bias_base = sbase + TS_BIAS;
    const uint32_t accfree_leader = mapa_cluster(bar_accfree, 0), aready_leader = mapa_cluster(bar_aready, 0);
    const int fsub = 32 * (cq & 1) + 128 * (cq >> 1);    // output feature of this thread's first column: 64 h + fsub
    uint32_t full_ph = 0, free_ph = 0, pend = 0, batch = 0;
    for (int it = 0; it < my_rounds; ++it) {
      const int64_t tile0 = 4 * ((int64_t)cid + (int64_t)it * ncl) + 2 * (int64_t)rank;
      // ---- prologue: gamma(pts) -> the slot's gamma atom (SS operand of layers 0 and 5), one quarter per thread; layer 0's bias row.
      //      The previous round's FINAL used the gamma atoms as scratch: everybody is through with it first.
      if (it > 0) named_bar_sync(1, kFwdEpiThreads);
      if (tix < 256) sts32f(bias_base + 4 * tix, __ldg(cst + c_step_bias[0] + tix));
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const int64_t row = (tile0 + t) * kTileM + r;
        const bool live = row < p.m;
        float pt[3] = {0, 0, 0}, dir[3] = {0, 0, 0};
        if (live) fetch_sample(p.src, row, pt, dir);
        uint32_t w[8];
        encode_pts_quarter(cq, pt, w);
        store_enc_chunks<8, 2, kTrain>(w, 2 * cq, live, sbase + TS_GAMMA + t * kAtomBytes,
                                       kTrain ? p.stash + (size_t)(tile0 + t) * kStashTileBytes + (size_t)SA_ENC * kAtomBytes : nullptr, r);
        tcgen05_fence_before_sync();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(aready_leader + 8 * t);
      }
      float alpha0 = 0.0f, alpha1 = 0.0f;
#pragma unroll 1
      for (int L = 0; L < kTsLayers; ++L) {
        const int sl = c_l_step0[L] + c_l_parts[L] - 1;   // the layer's last step carries its epilogue / bias / stash entries
        const int epi = c_step_epi[sl];
        // once per layer: all warps are through the previous layer's epilogues -> the next layer's bias row (parity buffer) may be staged
        named_bar_sync(1, kFwdEpiThreads);
        if (L + 1 < kTsLayers) {
          const int sn = c_l_step0[L + 1] + c_l_parts[L + 1] - 1;
          if (tix < (L + 1 == kTsLayers - 1 ? 128 : 256))
            sts32f(bias_base + (uint32_t)((L + 1) & 1) * 1024u + 4 * tix, __ldg(cst + c_step_bias[sn] + tix));
        }
        const uint32_t bias_row = bias_base + (uint32_t)(L & 1) * 1024u;
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          uint8_t* const stash_tile = kTrain ? p.stash + (size_t)(tile0 + t) * kStashTileBytes : nullptr;
          const uint32_t a_tmem = tmem_base + lane_off + kA0 + (uint32_t)t * 128u;
          if (epi == EPI_FINAL) {
            // ---- views layer (N = 128, one batch per slot): hv = relu(acc + bv); rgb = Wr hv + br; raw = [rgb, alpha] (helpers:117-123)
            const uint32_t acc = batch & 1u;
            ++batch;
            mbar_wait(bar_accfull + 8 * acc, (full_ph >> acc) & 1u);
            full_ph ^= 1u << acc;
            tcgen05_fence_after_sync();
            // the slot's gamma atom is dead now: Wr [3][128] fp32 is staged in its first 1.5 KB, the partial exchange behind it
            const uint32_t wr_a = sbase + TS_GAMMA + t * kAtomBytes;
            if (tix < 384) sts32f(wr_a + 4 * tix, __ldg(cst + C_WR + tix));
            uint32_t va[32];
            tmem_ld32(tmem_base + lane_off + acc * kAccCols + cq * 32, va);
            tmem_ld_wait_dep(va);
            tcgen05_fence_before_sync();
            named_bar_sync(1, kFwdEpiThreads);
            if (lane == 0) mbar_arrive_cluster(accfree_leader + 8 * acc);
            float rgb[3] = {0.f, 0.f, 0.f};
            const uint32_t m0 = epi_final32<kTrain>(va, cq * 32, bias_row, wr_a, rgb, stash_tile, r);
            if (kTrain) reinterpret_cast<uint32_t*>(stash_tile + kStashMaskOff)[(8 * 128 + r) * 8 + cq] = m0;
            float4* xchg = reinterpret_cast<float4*>(smem + TS_GAMMA + t * kAtomBytes + 2048);
            const float al = t ? alpha1 : alpha0;
            if (cq != 0) xchg[(cq - 1) * 128 + r] = make_float4(rgb[0], rgb[1], rgb[2], al);
            named_bar_sync(1, kFwdEpiThreads);
            const int64_t row = (tile0 + t) * kTileM + r;
            if (cq == 0 && row < p.m) {
              const float4 o1 = xchg[r], o2 = xchg[128 + r], o3 = xchg[256 + r];
              *reinterpret_cast<float4*>(p.raw + row * 4) =
                  make_float4(rgb[0] + o1.x + o2.x + o3.x + __ldg(cst + C_BR), rgb[1] + o1.y + o2.y + o3.y + __ldg(cst + C_BR + 1),
                              rgb[2] + o1.z + o2.z + o3.z + __ldg(cst + C_BR + 2), al + o1.w + o2.w + o3.w + __ldg(cst + C_BA));
            }
            continue;   // the slot's next a_ready arrival comes from the next round's prologue
          }
          // ---- 256-wide layer: half a (accumulator X), then half b (accumulator Y)
          uint32_t ra[16], rb[16];
          uint32_t mka, mkb;
          float al = t ? alpha1 : alpha0;
          {
            const uint32_t acc = batch & 1u;
            ++batch;
            mbar_wait(bar_accfull + 8 * acc, (full_ph >> acc) & 1u);
            full_ph ^= 1u << acc;
            tcgen05_fence_after_sync();
            if (tix == 0 && tracing) trace_stamp(p.trace, it, sl, t, 4);
            uint32_t v[32];
            tmem_ld32(tmem_base + lane_off + acc * kAccCols + cq * 32, v);
            const uint32_t sb = 2u * (uint32_t)t;                 // staging buffer of (slot t, half a)
            // the staging buffer's previous stash copy has been read (four half epilogues ago): checked under the TMEM load's latency
            if (kTrain && ((pend >> sb) & 1u)) { mbar_wait(bar_free + 8 * sb, (free_ph >> sb) & 1u); free_ph ^= 1u << sb; pend &= ~(1u << sb); }
            tmem_ld_wait_dep(v);
            tcgen05_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(accfree_leader + 8 * acc);      // the accumulator may be overwritten
            if (tix == 0 && tracing) trace_stamp(p.trace, it, sl, t, 9);
            const uint32_t stage_row = sbase + TS_STAGE + sb * kAtomBytes + (uint32_t)r * 128u;
            const uint32_t ba = bias_row + 4u * (uint32_t)fsub, wa = sbase + TS_WA + 2u * (uint32_t)fsub;
            if (epi == EPI_RELU) mka = epi32_ts<kTrain, 0>(v, ba, wa, al, ra, stage_row, 2u * cq, rx);
            else if (epi == EPI_RELU_ALPHA) mka = epi32_ts<kTrain, 1>(v, ba, wa, al, ra, stage_row, 2u * cq, rx);
            else mka = epi32_ts<kTrain, 2>(v, ba, wa, al, ra, stage_row, 2u * cq, rx);
            if (tix == 0 && tracing) trace_stamp(p.trace, it, sl, t, 6);
            if (kTrain) pend |= 1u << sb;                         // handed to the store lane together with half b's buffer (one proxy fence)
          }
          {
            const uint32_t acc = batch & 1u;
            ++batch;
            mbar_wait(bar_accfull + 8 * acc, (full_ph >> acc) & 1u);   // every MMA that read the slot's old activations has completed
            full_ph ^= 1u << acc;
            tcgen05_fence_after_sync();
            if (tix == 0 && tracing) trace_stamp(p.trace, it, sl, t, 5);
            tmem_st16(a_tmem + (uint32_t)(fsub >> 1), ra);
            uint32_t v[32];
            tmem_ld32(tmem_base + lane_off + acc * kAccCols + cq * 32, v);
            const uint32_t sb = 2u * (uint32_t)t + 1u;            // staging buffer of (slot t, half b)
            if (kTrain && ((pend >> sb) & 1u)) { mbar_wait(bar_free + 8 * sb, (free_ph >> sb) & 1u); free_ph ^= 1u << sb; pend &= ~(1u << sb); }
            tmem_ld_wait_dep(v);
            tcgen05_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(accfree_leader + 8 * acc);
            if (tix == 0 && tracing) trace_stamp(p.trace, it, sl, t, 12);
            const uint32_t stage_row = sbase + TS_STAGE + sb * kAtomBytes + (uint32_t)r * 128u;
            const uint32_t ba = bias_row + 4u * (uint32_t)(64 + fsub), wa = sbase + TS_WA + 2u * (uint32_t)(64 + fsub);
            if (epi == EPI_RELU) mkb = epi32_ts<kTrain, 0>(v, ba, wa, al, rb, stage_row, 2u * cq, rx);
            else if (epi == EPI_RELU_ALPHA) mkb = epi32_ts<kTrain, 1>(v, ba, wa, al, rb, stage_row, 2u * cq, rx);
            else mkb = epi32_ts<kTrain, 2>(v, ba, wa, al, rb, stage_row, 2u * cq, rx);
            tmem_st16(a_tmem + (uint32_t)((64 + fsub) >> 1), rb);
            if (tix == 0 && tracing) trace_stamp(p.trace, it, sl, t, 13);
            if (epi == EPI_RELU_ALPHA) { if (t) alpha1 = al; else alpha0 = al; }
            if (sl == 6) {
              // the skip layer's MMAs of this slot are done: gamma(pts) -> gamma(dir) in the slot's gamma atom (views layer, K = 32)
              const int64_t row = (tile0 + t) * kTileM + r;
              const bool live = row < p.m;
              float pt[3] = {0, 0, 0}, dir[3] = {0, 0, 0};
              if (live) fetch_sample(p.src, row, pt, dir);
              uint32_t dw[4];
              encode_dir_quarter(cq, dir, dw);
              store_enc_chunks<4, 1, kTrain>(dw, cq, live, sbase + TS_GAMMA + t * kAtomBytes, stash_tile + (size_t)SA_DENC * kAtomBytes, r);
              if (kTrain)   // wgrad reads all 64 columns of the stashed atom: columns 32..63 are zero padding
                *reinterpret_cast<uint4*>(stash_tile + (size_t)SA_DENC * kAtomBytes + sw128_off((uint32_t)r, (uint32_t)(4 + cq))) = make_uint4(0, 0, 0, 0);
            }
            tmem_st_wait();
            tcgen05_fence_before_sync();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive_cluster(aready_leader + 8 * t);
              if (kTrain) { mbar_arrive(bar_written + 8 * (sb - 1u)); mbar_arrive(bar_written + 8 * sb); }
            }
            if (kTrain) pend |= 1u << sb;
            if (tix == 0 && tracing) trace_stamp(p.trace, it, sl, t, 7);
          }
          // mask words after the hand-over: word w covers features 32w..32w+31
          if (kTrain && c_step_mask_slot[sl] >= 0 && !(p.debug & 2)) {
            uint32_t* mrow = reinterpret_cast<uint32_t*>(stash_tile + kStashMaskOff) + (c_step_mask_slot[sl] * 128 + r) * 8;
            mrow[fsub >> 5] = mka;
            mrow[(64 + fsub) >> 5] = mkb;
          }
        }
      }
    }
  }

  __syncwarp();
  tcgen05_fence_before_sync();
  cluster_sync_all();
  if (warp == kMmaWarp) {
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
  static const int fwd_debug = getenv("SPN_FWD_DEBUG") ? atoi(getenv("SPN_FWD_DEBUG")) : 0;
  p.debug = fwd_debug;
  static const int trace_block = getenv("SPN_TRACE_BLOCK") ? atoi(getenv("SPN_TRACE_BLOCK")) : 0;
  p.trace_block = trace_block;
  const int pairs = sm_count() / 2;
  int grid = 2 * (p.num_quads < pairs ? p.num_quads : pairs);
  auto kern = stash ? mlp_fwd_ts_kernel<true> : mlp_fwd_ts_kernel<false>;
  const int smem_bytes = kTsSmemBytes;
  static bool attr_set[2] = {false, false};
  if (!attr_set[stash ? 1 : 0]) {
    SPN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set[stash ? 1 : 0] = true;
  }
  prof_begin(PROF_MLP_FWD, st);
  kern<<<grid, kPairThreads, smem_bytes, st>>>(p);
  prof_end(PROF_MLP_FWD, st);
  SPN_LAUNCH_CHECK("mlp_fwd_ts_kernel");
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

// ---- diagnostic: rate of CTA-pair MMAs (cta_group::2, M = 256) with A from shared memory (ts = 0) or from TMEM (ts = 1), -------
// round-robin over `nacc` accumulators; `ld_warps` warps of both CTAs meanwhile drain / refill OTHER TMEM columns the way
// the TS epilogue does (tcgen05.ld 32 columns, tcgen05.st 16 columns).  out[0] = cycles for `reps` MMAs (leader thread).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(17 * 32, 1)
mma_rate_pair_kernel(int ts, int n, int reps, int nacc, int ld_warps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  for (int i = threadIdx.x; i < (64 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  const uint32_t bar = sbase + 64 * 1024;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 64 * 1024 + 64);
  volatile int* stop = reinterpret_cast<volatile int*>(smem + 64 * 1024 + 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); *stop = 0; }
  if (warp == 0) { tmem_alloc2(smem_u32(tptr), 512); tmem_relinquish2(); }
  fence_proxy_async_smem();
  tcgen05_fence_before_sync();
  cluster_sync_all();
  tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tptr;
  if (warp == 16) {
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = make_idesc(256, n, 0, 0);
      const uint64_t a0 = make_smem_desc(sbase, 16, 1024), b0 = make_smem_desc(sbase + 32768, 16, 1024);
      const long long t0 = clock64();
      for (int i = 0; i < reps; ++i) {
        const uint32_t d = tmem_base + (uint32_t)(i % nacc) * (uint32_t)n;
        if (ts) umma_bf16_ts_2cta(d, tmem_base + 384u + (uint32_t)(8 * (i & 15)), b0 + (uint64_t)(2 * (i & 3)), idesc, 1);
        else umma_bf16_2cta(d, a0 + (uint64_t)(2 * (i & 3)), b0 + (uint64_t)(2 * (i & 3)), idesc, 1);
      }
      umma_commit_2cta(bar, 3);
      mbar_wait(bar, 0);
      out[0] = clock64() - t0;
      *stop = 1;
      *reinterpret_cast<volatile int*>(smem + 64 * 1024 + 132) = 1;
    } else if (lane == 0) {
      mbar_wait(bar, 0);
      *stop = 1;
    }
  } else if (warp < ld_warps) {
    // epilogue-like TMEM traffic on columns 256..383 (never touched by the MMAs above when n * nacc <= 256)
    const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256u + (uint32_t)((warp >> 2) & 3) * 32u;
    uint32_t acc = 0;
    while (!*stop) {
      uint32_t v[32], w[16];
      tmem_ld32(ta, v);
      tmem_ld_wait_dep(v);
#pragma unroll
      for (int j = 0; j < 16; ++j) { w[j] = v[2 * j] + v[2 * j + 1]; acc ^= w[j]; }
      tmem_st16(ta, w);
      tmem_st_wait();
    }
    if (acc == 0x1234567u) out[1] = 1;
  }
  __syncwarp();
  tcgen05_fence_before_sync();
  cluster_sync_all();
  if (warp == 0) { tcgen05_fence_after_sync(); tmem_dealloc2(tmem_base, 512); }
}

int tc_mma_rate_pair(int ts, int n, int reps, int nacc, int ld_warps, long long* out, cudaStream_t st) {
  int rc = check_arch();
  if (rc != SPN_OK) return rc;
  SPN_CHECK_ARG(out && reps > 0 && (n == 64 || n == 128 || n == 256) && nacc >= 1 && n * nacc <= 256 && ld_warps >= 0 && ld_warps <= 16,
                "spn_tc_mma_rate_pair: bad arguments");
  const int smem_bytes = 64 * 1024 + 256 + 1024;
  SPN_CUDA(cudaFuncSetAttribute(mma_rate_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  mma_rate_pair_kernel<<<2, 17 * 32, smem_bytes, st>>>(ts, n, reps, nacc, ld_warps, out);
  SPN_LAUNCH_CHECK("mma_rate_pair_kernel");
  return SPN_OK;
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
