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
#include "common.cuh"
#include "tc_common.cuh"
#include "mlp_tc.cuh"

namespace spn {
using namespace tc;

size_t mlp_tc_packed_bytes() { return kPackedBytes; }

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

// gamma(v) (3 + 6*NFREQ values, zero padded to 64) as bf16 -> row r of a [128 x 64] swizzled atom in shared memory
// (and, in training, the same image in the stash).  Dead rows (beyond m) get zeros.
template <int NFREQ, bool kStash>
__device__ __forceinline__ void write_encoding(const float v[3], bool live, uint8_t* act_atom, uint8_t* stash_atom, int r) {
  constexpr int NW = NFREQ == 10 ? 32 : 16;
  uint32_t w[NW];
  encode_point<NFREQ, NW>(v, w);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint4 q4 = make_uint4(0, 0, 0, 0);
    if (4 * j < NW && live) q4 = make_uint4(w[(4 * j) % NW], w[(4 * j + 1) % NW], w[(4 * j + 2) % NW], w[(4 * j + 3) % NW]);
    *reinterpret_cast<uint4*>(act_atom + sw128_off(r, j)) = q4;
    if (kStash) *reinterpret_cast<uint4*>(stash_atom + sw128_off(r, j)) = q4;
  }
}

// 16 accumulator columns of a hidden layer: h = acc + bias (ReLU), bf16, swizzled store; returns 16 ReLU mask bits
template <bool kTrain, bool kRelu, bool kAlpha>
__device__ __forceinline__ uint32_t epi_cols16(const uint32_t (&v)[16], int col0, const float* bias,
                                               const float* __restrict__ cst, float& alpha, uint8_t* act, int r) {
  uint32_t mb = 0;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int col = col0 + g * 8;
    const float4 b0 = *reinterpret_cast<const float4*>(bias + col);        // staged in shared memory
    const float4 b1 = *reinterpret_cast<const float4*>(bias + col + 4);
    float h[8] = {__uint_as_float(v[g * 8 + 0]) + b0.x, __uint_as_float(v[g * 8 + 1]) + b0.y,
                  __uint_as_float(v[g * 8 + 2]) + b0.z, __uint_as_float(v[g * 8 + 3]) + b0.w,
                  __uint_as_float(v[g * 8 + 4]) + b1.x, __uint_as_float(v[g * 8 + 5]) + b1.y,
                  __uint_as_float(v[g * 8 + 6]) + b1.z, __uint_as_float(v[g * 8 + 7]) + b1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (kRelu) h[e] = fmaxf(h[e], 0.f);
      if (kTrain) mb |= (h[e] > 0.f ? 1u : 0u) << (g * 8 + e);
    }
    if (kAlpha) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(cst + C_WA + col));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(cst + C_WA + col + 4));
      alpha = fmaf(h[0], w0.x, fmaf(h[1], w0.y, fmaf(h[2], w0.z, fmaf(h[3], w0.w, alpha))));
      alpha = fmaf(h[4], w1.x, fmaf(h[5], w1.y, fmaf(h[6], w1.z, fmaf(h[7], w1.w, alpha))));
    }
    const uint4 v4 = make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]), pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7]));
    *reinterpret_cast<uint4*>(act + (uint32_t)(col / 64) * kAtomBytes + sw128_off(r, (col % 64) / 8)) = v4;
  }
  return mb;
}

// 16 columns of the views layer: hv = relu(acc + bv), rgb += Wr[:, col] hv (fp32 partial), hv stashed in training
template <bool kTrain>
__device__ __forceinline__ uint32_t epi_final16(const uint32_t (&v)[16], int col0, const float* bias,
                                                const float* __restrict__ cst, float (&rgb)[3], uint8_t* stash_tile, int r) {
  uint32_t mb = 0;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int col = col0 + g * 8;
    float h[8];
#pragma unroll
    for (int e = 0; e < 8; e += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias + col + e);   // staged in shared memory
      const float4 r0 = __ldg(reinterpret_cast<const float4*>(cst + C_WR + col + e));
      const float4 r1 = __ldg(reinterpret_cast<const float4*>(cst + C_WR + 128 + col + e));
      const float4 r2 = __ldg(reinterpret_cast<const float4*>(cst + C_WR + 256 + col + e));
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

template <bool kTrain>
__global__ void __launch_bounds__(kThreads, 1) mlp_fwd_kernel(const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (smem != smem_raw) __trap();   // kSmemBytes has no alignment slack: the dynamic window must start 1024-aligned
  // barriers: full[3] empty[3] acc_full[2] act_ready[2]; tmem pointer after them
  const uint32_t bar_full = sbase + SM_BAR, bar_empty = bar_full + 8 * kStages;
  const uint32_t bar_acc = bar_empty + 8 * kStages, bar_act = bar_acc + 16;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SM_BAR + 128);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int t = 0; t < 2; ++t) { mbar_init(bar_acc + 8 * t, 1); mbar_init(bar_act + 8 * t, kEpiThreads); }
    fence_mbar_init();
  }
  if (warp == 1) {   // TMEM: 512 columns = two 128x256 fp32 accumulators
    tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish();
  }
  tcgen05_fence_before_sync();
  __syncthreads();
  tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const float* cst = reinterpret_cast<const float*>(p.packed + kFwdBytes + kBwdBytes);
  const int my_pairs = (p.num_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    // ================= weight producer =================
    // One thread retires at most one cp.async.bulk per ~700 cycles (tools/bulk_rate.py), so a chunk is split over
    // kCopyLanes lanes that issue their pieces concurrently on the same mbarrier.
    constexpr int kCopyLanes = 8;
    uint32_t stage = 0, phase = 0;
    for (int it = 0; it < my_pairs; ++it) {
      const uint8_t* src = p.packed;
      for (int s = 0; s < kNumSteps; ++s) {
        const uint32_t bytes = (uint32_t)c_step_n[s] * 128u;
        const uint32_t piece = bytes / kCopyLanes;
        for (int t = 0; t < 2; ++t) {
          const uint8_t* sp = src;
          for (int c = 0; c < c_step_chunks[s]; ++c) {
            if (lane == 0) {
              mbar_wait(bar_empty + 8 * stage, phase ^ 1);
              mbar_arrive_expect_tx(bar_full + 8 * stage, bytes);
            }
            __syncwarp();
            if (lane < kCopyLanes)
              bulk_g2s(sbase + SM_RING + stage * kChunkBig + lane * piece, sp + lane * piece, piece, bar_full + 8 * stage);
            sp += bytes;
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
        src += (size_t)c_step_chunks[s] * bytes;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      uint32_t act_phase[2] = {0, 0};
      for (int it = 0; it < my_pairs; ++it) {
        for (int s = 0; s < kNumSteps; ++s) {
          const int nch = c_step_chunks[s], n = c_step_n[s], ksteps = c_step_ksteps[s];
          const uint32_t idesc = make_idesc(kTileM, n, 0, 0);
          for (int t = 0; t < 2; ++t) {
            mbar_wait(bar_act + 8 * t, act_phase[t]);   // A operand written, accumulator drained
            act_phase[t] ^= 1;
            tcgen05_fence_after_sync();
            const uint32_t d_tmem = tmem_base + (uint32_t)t * 256u;
            uint32_t accumulate = (uint32_t)c_step_acc[s];
            for (int c = 0; c < nch; ++c) {
              mbar_wait(bar_full + 8 * stage, phase);
              tcgen05_fence_after_sync();
              const uint32_t a_addr = sbase + SM_ACT + t * kActBytes + (nch == 1 ? 0 : c) * kAtomBytes;
              const uint32_t b_addr = sbase + SM_RING + stage * kChunkBig;
              const uint64_t a_desc = make_smem_desc(a_addr, 16, 1024);
              const uint64_t b_desc = make_smem_desc(b_addr, 16, 1024);
              for (int k = 0; k < ksteps; ++k) {
                // +32 bytes (16 bf16) along K inside the 128-byte swizzle atom: start-address field += 2
                umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, accumulate);
                accumulate = 1;
              }
              umma_commit(bar_empty + 8 * stage);      // ring slot free once these MMAs have read it
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            umma_commit(bar_acc + 8 * t);              // accumulator complete -> epilogue of tile t
          }
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
    float4* xchg = reinterpret_cast<float4*>(act + 3 * kAtomBytes);   // FINAL-step scratch (tile is dead by then)
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)t * 256u;
    uint32_t acc_phase = 0;
    for (int it = 0; it < my_pairs; ++it) {
      const int64_t tile = 2 * ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) + t;
      const int64_t row = tile * kTileM + r;
      const bool live = row < p.m;
      uint8_t* stash_tile = kTrain ? p.stash + (size_t)tile * kStashTileBytes : nullptr;
      // ---- prologue: sample point -> gamma(pts) -> A atom 0 (column half 0; half 1 later writes gamma(dir)).
      //      Only the 3-D point / direction stay in registers; encodings are re-derived when a pass needs them.
      float pt[3] = {0, 0, 0}, dir[3] = {0, 0, 0};
      if (live) fetch_sample(p.src, row, pt, dir);
      if (cg == 0) write_encoding<10, kTrain>(pt, live, act, stash_tile + (size_t)SA_ENC * kAtomBytes, r);
      fence_proxy_async_smem();
      mbar_arrive(bar_act + 8 * t);
      float alpha = 0.0f;                                // this half's partial of the sigma head
      float* bias_s = reinterpret_cast<float*>(smem + SM_BIAS) + t * 256;
      for (int s = 0; s < kNumSteps; ++s) {
        const int epi = c_step_epi[s];
        // Stage this step's bias row in shared memory while the MMAs of the step are still running: with 227 KB of
        // the SM carved out as shared memory there is no L1 left, so a __ldg in the epilogue costs an L2 round trip.
        named_bar_sync(1 + t, kEpiThreads);                   // everyone is done reading the previous bias row
        if (epi != EPI_WRITE_ENC && epi != EPI_WRITE_DENC)
          bias_s[tix] = __ldg(cst + c_step_bias[s] + (epi == EPI_FINAL ? (tix & 127) : tix));
        mbar_wait(bar_acc + 8 * t, acc_phase);
        acc_phase ^= 1;
        tcgen05_fence_after_sync();
        if (kTrain && lane == 0 && cg == 0) bulk_wait_read0();   // previous stash stores have finished READING the tile
        named_bar_sync(1 + t, kEpiThreads);
        if (epi == EPI_WRITE_ENC || epi == EPI_WRITE_DENC) {
          // pass 1 has finished reading the tile: overwrite atom 0 with the second-pass operand
          if (epi == EPI_WRITE_ENC) { if (cg == 0) write_encoding<10, false>(pt, live, act, nullptr, r); }
          else if (cg == 1) write_encoding<4, kTrain>(dir, live, act, stash_tile + (size_t)SA_DENC * kAtomBytes, r);
          fence_proxy_async_smem();
          mbar_arrive(bar_act + 8 * t);
          continue;
        }
        const float* bias = bias_s;
        uint32_t va[16], vb[16];
        if (epi == EPI_FINAL) {
          // hv = relu(acc + bv) [128];  rgb = Wr hv + br;  raw = [rgb, alpha]   (helpers:117-123); 64 columns per half
          float rgb[3] = {0.f, 0.f, 0.f};
          uint32_t mask[2];
          const int c0 = cg * 64;
          tmem_ld16(tmem_lane + c0, va);
#pragma unroll
          for (int sb = 0; sb < 4; sb += 2) {
            tmem_ld_wait_dep16(va);
            tmem_ld16(tmem_lane + c0 + (sb + 1) * 16, vb);
            const uint32_t m0 = epi_final16<kTrain>(va, c0 + sb * 16, bias, cst, rgb, stash_tile, r);
            tmem_ld_wait_dep16(vb);
            if (sb + 2 < 4) tmem_ld16(tmem_lane + c0 + (sb + 2) * 16, va);
            const uint32_t m1 = epi_final16<kTrain>(vb, c0 + (sb + 1) * 16, bias, cst, rgb, stash_tile, r);
            mask[sb / 2] = m0 | (m1 << 16);
          }
          if (kTrain) {
            uint32_t* mrow = reinterpret_cast<uint32_t*>(stash_tile + kStashMaskOff) + (8 * 128 + r) * 8 + cg * 2;
            *reinterpret_cast<uint2*>(mrow) = make_uint2(mask[0], mask[1]);
          }
          if (cg == 1) xchg[r] = make_float4(rgb[0], rgb[1], rgb[2], alpha);
          tcgen05_fence_before_sync();
          named_bar_sync(1 + t, kEpiThreads);
          if (cg == 0 && live) {
            const float4 o = xchg[r];
            *reinterpret_cast<float4*>(p.raw + row * 4) =
                make_float4(rgb[0] + o.x + cst[C_BR], rgb[1] + o.y + cst[C_BR + 1], rgb[2] + o.z + cst[C_BR + 2],
                            alpha + o.w + cst[C_BA]);
          }
          continue;   // next arrival on act_ready comes from the next tile's prologue
        }
        // ---- bias (+ReLU) -> bf16 -> swizzled in-place store; layer 7 also accumulates sigma from fp32 h7.
        //      128 columns per warp in 16-column TMEM loads, double-buffered.
        uint8_t* stash_layer = kTrain ? stash_tile + (size_t)c_step_stash_atom[s] * kAtomBytes : nullptr;
        uint32_t maskw[4];
        const int c0 = cg * 128;
        tmem_ld16(tmem_lane + c0, va);
#pragma unroll
        for (int sb = 0; sb < 8; sb += 2) {
          uint32_t m0, m1;
          tmem_ld_wait_dep16(va);
          tmem_ld16(tmem_lane + c0 + (sb + 1) * 16, vb);
          if (epi == EPI_RELU) m0 = epi_cols16<kTrain, true, false>(va, c0 + sb * 16, bias, cst, alpha, act, r);
          else if (epi == EPI_RELU_ALPHA) m0 = epi_cols16<kTrain, true, true>(va, c0 + sb * 16, bias, cst, alpha, act, r);
          else m0 = epi_cols16<kTrain, false, false>(va, c0 + sb * 16, bias, cst, alpha, act, r);
          tmem_ld_wait_dep16(vb);
          if (sb + 2 < 8) tmem_ld16(tmem_lane + c0 + (sb + 2) * 16, va);
          if (epi == EPI_RELU) m1 = epi_cols16<kTrain, true, false>(vb, c0 + (sb + 1) * 16, bias, cst, alpha, act, r);
          else if (epi == EPI_RELU_ALPHA) m1 = epi_cols16<kTrain, true, true>(vb, c0 + (sb + 1) * 16, bias, cst, alpha, act, r);
          else m1 = epi_cols16<kTrain, false, false>(vb, c0 + (sb + 1) * 16, bias, cst, alpha, act, r);
          maskw[sb / 2] = m0 | (m1 << 16);
        }
        if (kTrain && c_step_mask_slot[s] >= 0) {
          uint32_t* mrow = reinterpret_cast<uint32_t*>(stash_tile + kStashMaskOff) + (c_step_mask_slot[s] * 128 + r) * 8 + cg * 4;
          *reinterpret_cast<uint4*>(mrow) = make_uint4(maskw[0], maskw[1], maskw[2], maskw[3]);
        }
        tcgen05_fence_before_sync();
        fence_proxy_async_smem();
        mbar_arrive(bar_act + 8 * t);
        if (kTrain) {
          named_bar_sync(1 + t, kEpiThreads);    // every row of the tile is written and fenced
          if (lane == 0 && cg == 0) {            // one 16 KB atom per warp: bulk-copy issue is serialised per thread
            bulk_s2g(stash_layer + q * kAtomBytes, smem_u32(act) + q * kAtomBytes, kAtomBytes);
            bulk_commit();
          }
        }
      }
      if (kTrain && lane == 0 && cg == 0) bulk_wait0();   // stash complete before the tile slot is reused / kernel exit
    }
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
