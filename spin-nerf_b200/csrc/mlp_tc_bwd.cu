// Backward of the NeRF MLP (what autograd does for DS_NeRF/run_nerf_helpers.py:104-127) on tcgen05.
//
// Three kernels, all reading the bf16 activation stash the training forward wrote (mlp_tc.cu):
//   mlp_dgrad_kernel   per 128-sample tile, the chain  d_hv -> d_feat -> d_h7 -> ... -> d_h0  as nine fused GEMMs
//                      against the transposed weight images (same ping-pong / TMEM / bulk-copy-ring skeleton as
//                      the forward); ReLU masks come from the 1-bit-per-activation mask stash; every
//                      d(pre-activation) tile is written once as a bf16 SWIZZLE_128B image (the "dstash").
//   mlp_wgrad_kernel   dW_l = dpre_l^T . in_l  as a split-K GEMM over ALL samples: operands are the stash /
//                      dstash tiles used directly as MN-major UMMA operands (K = sample index), fp32 accumulators
//                      for a full 256x256 weight gradient live in the 512 TMEM columns, one red.global flush per
//                      (layer, sample-range) segment.  Bias gradients are column sums taken from the same
//                      shared-memory slabs by otherwise idle warps.
//   mlp_heads_wgrad_kernel   the 3x128 rgb / 1x256 sigma heads on CUDA cores.
// No gradient flows to the sampled points (z_samples are detached, run_nerf.py:700).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "mlp_tc.cuh"

namespace spn {
using namespace tc;

static int check_arch_bwd() {
  static int ok = -1;
  if (ok < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
      major = 0;
    ok = (major == 10) ? 1 : 0;
  }
  if (!ok) {
    set_error("the tcgen05 MLP kernels need an sm_100 device (B200)");
    return SPN_E_ARCH;
  }
  return SPN_OK;
}

static inline int64_t quad_tiles(int64_t m) { return (m + 4 * kTileM - 1) / (4 * kTileM) * 4; }   // a CTA pair works on 4 tiles
constexpr int kWgMaxCtas = 192;   // >= SM count of any sm_100 part
size_t mlp_tc_bwd_ws_bytes(int64_t m) {
  return (size_t)quad_tiles(m) * kDstashTileBytes + (size_t)kWgMaxCtas * 256 * 256 * sizeof(float) + 256;
}

// =====================================================================================================
// dgrad: CTA pairs (cta_group::2), same skeleton as the forward kernel (mlp_tc.cu: ring, barriers, roles)
// =====================================================================================================
constexpr int kDgSteps = 9;
__constant__ int c_dg_chunks[kDgSteps] = {2, 4, 4, 4, 4, 4, 4, 4, 4};
// ReLU-mask slot applied by the step's epilogue (-1: none) and destination atom inside the dstash tile
__constant__ int c_dg_mask[kDgSteps] = {-1, 7, 6, 5, 4, 3, 2, 1, 0};
__constant__ int c_dg_dst[kDgSteps] = {DA_FEAT, DA_H7, DA_H7 + 4, DA_H7 + 8, DA_H7 + 12, DA_H7 + 16, DA_H7 + 20,
                                       DA_H7 + 24, DA_H7 + 28};

struct DgradParams {
  const uint8_t* packed;
  const uint8_t* stash;
  const float* d_raw;
  uint8_t* dstash;
  int64_t m;
  int num_quads;
};

// 32 accumulator columns of a dgrad epilogue: (+ d_sigma * Wa) -> ReLU mask (bit = column) -> bf16 -> A operand of the
// next GEMM.   wa_a: shared address of this thread's first bf16 sigma weight    row_a / rx: as in the forward epilogue
template <bool kAlpha>
__device__ __forceinline__ void dg_cols32(const uint32_t (&v)[32], const int cl, const uint32_t mb, const float dalpha,
                                          const uint32_t wa_a, const uint32_t row_a, const uint32_t rx) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int c = cl + g * 8;
    float h[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) h[e] = __uint_as_float(v[g * 8 + e]);
    if (kAlpha) {
      const uint4 wq = lds128u(wa_a + c * 2);
      const uint32_t ww[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        h[e] = fmaf(dalpha, __uint_as_float(ww[e / 2] << 16), h[e]);
        h[e + 1] = fmaf(dalpha, __uint_as_float(ww[e / 2] & 0xffff0000u), h[e + 1]);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) h[e] = ((mb >> relu_mask_bit(g * 8 + e)) & 1u) ? h[e] : 0.f;
    sts128(row_a + (uint32_t)(c / 64) * kAtomBytes + ((uint32_t)(((c % 64) / 8) << 4) ^ rx), pack_bf16(h[0], h[1]),
           pack_bf16(h[2], h[3]), pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7]));
  }
}

template <bool kAlpha>
__device__ __forceinline__ void dg_layer(const uint32_t tmem_a, const uint32_t (&mw)[4], const float dalpha, const uint32_t wa_a,
                                         const uint32_t row_a, const uint32_t rx) {
  uint32_t va[32], vb[32];
  tmem_ld32(tmem_a, va);
  tmem_ld32(tmem_a + 32, vb);
  tmem_ld_wait_dep(va);
  tmem_ld_wait_dep(vb);
  dg_cols32<kAlpha>(va, 0, mw[0], dalpha, wa_a, row_a, rx);
  tmem_ld32(tmem_a + 64, va);
  dg_cols32<kAlpha>(vb, 32, mw[1], dalpha, wa_a, row_a, rx);
  tmem_ld32(tmem_a + 96, vb);
  tmem_ld_wait_dep(va);
  tmem_ld_wait_dep(vb);
  dg_cols32<kAlpha>(va, 64, mw[2], dalpha, wa_a, row_a, rx);
  dg_cols32<kAlpha>(vb, 96, mw[3], dalpha, wa_a, row_a, rx);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1) mlp_dgrad_kernel(const DgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (smem != smem_raw) __trap();   // kSmemBytes has no alignment slack
  // barriers as in mlp_fwd_kernel: group_full[2][4] (step counter % 4), empty[3], acc_full[2], act_ready[2]
  const uint32_t bar_full = sbase + SM_BAR, bar_empty = bar_full + 64;
  const uint32_t bar_acc = bar_empty + 8 * kSlots, bar_act = bar_acc + 16;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SM_TMEMPTR);
  const float* cst = reinterpret_cast<const float*>(p.packed + kFwdBytes + kBwdBytes);
  const uint32_t wr_s = sbase + SM_BIAS;        // rgb-head weights [3][128] fp32 (the bias rows are unused here)

  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) mbar_init(bar_empty + 8 * s, 1);
    for (int b = 0; b < 8; ++b) mbar_init(bar_full + 8 * b, rank == 0 ? 2 : 1);
    for (int t = 0; t < 2; ++t) { mbar_init(bar_acc + 8 * t, 1); mbar_init(bar_act + 8 * t, 2 * kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc2(smem_u32(tmem_ptr_smem), 512); tmem_relinquish2(); }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 256) {
    const int i = threadIdx.x - 64;
    const __nv_bfloat16 w = __float2bfloat16_rn(__ldg(cst + C_WA + i));          // the forward's bf16 sigma weights
    sts16(sbase + SM_WA + 2 * i, *reinterpret_cast<const uint16_t*>(&w));
    sts32f(wr_s + 4 * i, __ldg(cst + C_WR + i));
    if (i < 128) sts32f(wr_s + 4 * (256 + i), __ldg(cst + C_WR + 256 + i));
  }
  tcgen05_fence_before_sync();
  cluster_sync_all();
  tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int cid = (int)blockIdx.x >> 1, ncl = (int)gridDim.x >> 1;
  const int my_rounds = (p.num_quads - cid + ncl - 1) / ncl;

  if (warp == 0) {
    // ---- weight producer: this CTA's half (128 of 256 rows) of every transposed chunk; groups of two chunks per ring slot
    uint32_t stage = 0, phase = 0, gs = 0;
    for (int it = 0; it < my_rounds; ++it) {
      const uint8_t* src = p.packed + kFwdBytes;
      for (int s = 0; s < kDgSteps; ++s, ++gs) {
        const int nch = c_dg_chunks[s];
        for (int g = 0; 2 * g < nch; ++g) {
          const int sub = lane >= kSlots ? 1 : 0;
          if (lane == (int)stage || lane == (int)stage + kSlots) {
            const uint32_t gbar = bar_full + 8 * (gs & 3) + 32 * g;
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);       // both lanes, every release (see mlp_fwd_kernel)
            if (sub == 0) {
              mbar_arrive_expect_tx(gbar, 2u * (kChunkBig / 2));
              if (g == 0 && nch <= 2) mbar_arrive(gbar + 32);
            }
            bulk_g2s(sbase + SM_RING + stage * kSlotBytes + sub * (kSlotBytes / 2),
                     src + (size_t)(2 * g + sub) * kChunkBig + (size_t)rank * (kChunkBig / 2), kChunkBig / 2, gbar);
          }
          if (++stage == kSlots) { stage = 0; phase ^= 1; }
        }
        src += (size_t)nch * kChunkBig;
      }
    }
  } else if (warp == 1 && rank != 0) {
    if (lane == 0) {   // ---- peer: relay "my halves of this group have landed" to the leader's group barrier
      const uint32_t full_leader = mapa_cluster(bar_full, 0);
      uint32_t gs = 0;
      for (int it = 0; it < my_rounds; ++it)
        for (int s = 0; s < kDgSteps; ++s, ++gs)
          for (int g = 0; g < 2; ++g) {
            mbar_wait(bar_full + 32 * g + 8 * (gs & 3), (gs >> 2) & 1u);
            mbar_arrive_cluster(full_leader + 32 * g + 8 * (gs & 3));
          }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ---- leader: MMA issuer for the pair
      uint32_t grp = 0, gs = 0;
      uint32_t act_phase[2] = {0, 0};
      const uint32_t idesc = make_idesc(2 * kTileM, 256, 0, 0);
      for (int it = 0; it < my_rounds; ++it) {
        for (int s = 0; s < kDgSteps; ++s, ++gs) {
          const int nch = c_dg_chunks[s];
          const uint32_t lph = (gs >> 2) & 1u, gbar = bar_full + 8 * (gs & 3);
          mbar_wait_cluster(gbar, lph);
          for (int t = 0; t < 2; ++t) {
            mbar_wait_cluster(bar_act + 8 * t, act_phase[t]);
            act_phase[t] ^= 1;
            tcgen05_fence_after_sync();
            const uint32_t d_tmem = tmem_base + (uint32_t)t * 256u;
            uint32_t accumulate = 0;
            for (int c = 0; c < nch; ++c) {
              const uint32_t stage = (grp + (uint32_t)(c >> 1)) % kSlots;
              if (t == 0 && c == 2) { mbar_wait_cluster(gbar + 32, lph); tcgen05_fence_after_sync(); }
              const uint64_t a_desc = make_smem_desc(sbase + SM_ACT + t * kActBytes + c * kAtomBytes, 16, 1024);
              const uint64_t b_desc = make_smem_desc(sbase + SM_RING + stage * kSlotBytes + (c & 1) * (kSlotBytes / 2), 16, 1024);
              for (int k = 0; k < 4; ++k) {
                umma_bf16_2cta(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, accumulate);
                accumulate = 1;
              }
              if (t == 1 && ((c & 1) || c == nch - 1)) umma_commit_2cta(bar_empty + 8 * stage, 3);
            }
            umma_commit_2cta(bar_acc + 8 * t, 3);
          }
          grp += (uint32_t)((nch + 1) >> 1);
        }
      }
    }
  } else {
    // ---- prologue + epilogue warps: 8 per tile; warp (q, cg) owns rows 32q..32q+31, columns 128cg..128cg+127
    const int ew = warp - 2;
    const int t = ew >> 3;
    const int cg = (ew >> 2) & 1;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t act_a = sbase + SM_ACT + t * kActBytes;
    const uint32_t rx = (uint32_t)(r & 7) << 4;
    const uint32_t row_a = act_a + (uint32_t)cg * 2u * kAtomBytes + (uint32_t)r * 128u;
    const uint32_t wa_a = sbase + SM_WA + (uint32_t)cg * 256u;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)t * 256u + (uint32_t)cg * 128u;
    const uint32_t act_leader = mapa_cluster(bar_act + 8 * t, 0);
    const bool store_lane = lane == 0 && cg == 0;             // issues this warp's 16 KB dstash stores
    uint32_t acc_phase = 0;
    for (int it = 0; it < my_rounds; ++it) {
      const int64_t tile = 4 * ((int64_t)cid + (int64_t)it * ncl) + 2 * (int64_t)rank + t;
      const int64_t row = tile * kTileM + r;
      const bool live = row < p.m;
      const uint8_t* stash_tile = p.stash + (size_t)tile * kStashTileBytes;
      uint8_t* dst_tile = p.dstash + (size_t)tile * kDstashTileBytes;
      const uint32_t* masks = reinterpret_cast<const uint32_t*>(stash_tile + kStashMaskOff);
      // prologue: d_hv = (d_rgb . Wr) * (hv > 0)   [128 wide, 64 columns per half]  -> A atoms 0-1 and dstash atoms 0-1
      const float4 dr = live ? *reinterpret_cast<const float4*>(p.d_raw + row * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      const uint2 mk = *reinterpret_cast<const uint2*>(masks + (8 * 128 + r) * 8 + cg * 2);
      named_bar_sync(1 + t, kEpiThreads);      // the previous round's last dstash store was issued ...
      if (store_lane) bulk_wait_read0();       // ... and has finished reading the tile (all four issuing lanes wait)
      named_bar_sync(1 + t, kEpiThreads);
      {
        const uint32_t mw2[2] = {mk.x, mk.y};
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const int col = cg * 64 + cb * 32 + g8 * 8;       // column of d_hv (0..127)
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; e += 4) {
              const float4 w0 = lds128f(wr_s + 4 * (col + e)), w1 = lds128f(wr_s + 4 * (128 + col + e)),
                           w2 = lds128f(wr_s + 4 * (256 + col + e));
              v[e + 0] = dr.x * w0.x + dr.y * w1.x + dr.z * w2.x;
              v[e + 1] = dr.x * w0.y + dr.y * w1.y + dr.z * w2.y;
              v[e + 2] = dr.x * w0.z + dr.y * w1.z + dr.z * w2.z;
              v[e + 3] = dr.x * w0.w + dr.y * w1.w + dr.z * w2.w;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = ((mw2[cb] >> relu_mask_bit(g8 * 8 + e)) & 1u) ? v[e] : 0.f;
            sts128(act_a + (uint32_t)(col / 64) * kAtomBytes + sw128_off((uint32_t)r, (uint32_t)((col % 64) / 8)),
                   pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(act_leader);
      int pending_atom = DA_HV, pending_n = 2;                  // d_hv: atoms 0-1 of the tile
      for (int s = 0; s < kDgSteps; ++s) {
        named_bar_sync(1 + t, kEpiThreads);                     // the whole group has written the previous output
        if (store_lane && q < pending_n) {                      // one 16 KB atom per warp: bulk-copy issue is serialised per thread
          bulk_s2g(dst_tile + (size_t)(pending_atom + q) * kAtomBytes, act_a + q * kAtomBytes, kAtomBytes);
          bulk_commit();
        }
        // the step's ReLU mask words are fetched while its MMAs are still running (no L1 left: every load is an L2 trip)
        const int mslot = c_dg_mask[s];
        uint32_t mw[4] = {~0u, ~0u, ~0u, ~0u};
        if (mslot >= 0) {
          const uint4 m0 = *reinterpret_cast<const uint4*>(masks + (mslot * 128 + r) * 8 + cg * 4);
          mw[0] = m0.x; mw[1] = m0.y; mw[2] = m0.z; mw[3] = m0.w;
        }
        if (store_lane && q < pending_n) bulk_wait_read0();     // the store has finished READING the tile we overwrite
        named_bar_sync(1 + t, kEpiThreads);
        mbar_wait(bar_acc + 8 * t, acc_phase);
        acc_phase ^= 1;
        tcgen05_fence_after_sync();
        if (s == 1) dg_layer<true>(tmem_lane, mw, dr.w, wa_a, row_a, rx);   // d_h7 += d_sigma * Wa (alpha_linear, helpers:113)
        else dg_layer<false>(tmem_lane, mw, 0.f, wa_a, row_a, rx);
        tcgen05_fence_before_sync();
        fence_proxy_async_smem();
        __syncwarp();
        if (s != kDgSteps - 1 && lane == 0) mbar_arrive_cluster(act_leader);
        pending_atom = c_dg_dst[s]; pending_n = 4;
      }
      // the last layer's output (d_h0) is streamed out at the top of the next round / below
      named_bar_sync(1 + t, kEpiThreads);
      if (store_lane) {
        bulk_s2g(dst_tile + (size_t)(pending_atom + q) * kAtomBytes, act_a + q * kAtomBytes, kAtomBytes);
        bulk_commit();
      }
    }
    if (store_lane) bulk_wait0();    // dstash complete before the kernel can exit
  }
  __syncwarp();
  tcgen05_fence_before_sync();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after_sync();
    tmem_dealloc2(tmem_base, 512);
  }
}

// =====================================================================================================
// wgrad
// =====================================================================================================
// dW[out, in] = sum over samples of dpre[sample, out] * x[sample, in]: a split-K GEMM with K = sample.  A = dpre slabs from the
// dstash (bf16 SWIZZLE_128B images, used MN-major), B = the layer input.  The nine 256-wide inputs (h0..h7, feature) are
// stashed as E4M3 (mlp_tc.cuh): four otherwise idle warps widen each 64-sample slab to bf16 (exact) in the MN-major SWIZZLE_128B
// image the tensor core reads, while the previous slab's MMAs run (a bf16 x f16 MMA would save half the conversion work, but
// kind::f16 with different A and B formats raised "illegal instruction" on the B200).  The 64-wide encodings stay bf16 and are
// used as they land.
constexpr int kWgUnits = 12;
constexpr int kWgSlabRows = 64;                 // samples per pipeline stage
constexpr int kWgStages = 3;
constexpr int kWgSlabBytes = kWgSlabRows * 128; // one atom's slab: 8 KB
constexpr int kWgStageBytes = 6 * kWgSlabBytes; // A: 4 slabs (32 KB) at +0;  B as it lands: 2 E4M3 slabs or 1 bf16 slab at +32 KB
constexpr int WG_B16 = kWgStages * kWgStageBytes;          // 2 x 32 KB: the widened (bf16) B operand, double-buffered
constexpr int WG_BAR = WG_B16 + 2 * 4 * kWgSlabBytes;
constexpr int kWgConvWarps = 8;                   // warps 2..9 widen the E4M3 slabs; warps 2..5 also take the bias column sums and the flush
constexpr int kWgThreads = 32 * (2 + kWgConvWarps);
constexpr int kWgSmemBytes = WG_BAR + 256 + 1024;
static_assert(kWgSmemBytes <= 232448, "wgrad shared memory budget");

struct WgUnit {
  int d_atom;      // first atom of dpre inside the dstash tile
  int m_out;       // 256 | 128 output features
  int in_atom;     // first atom of the layer input inside the forward stash tile
  int n_in;        // 256 | 64 (padded) input features
  int n_valid;     // real input features (<= n_in)
  int w_off;       // float offset of dW[0][0] in the flat gradient
  int ld;          // row stride of dW
  int b_off;       // float offset of the bias gradient, -1 if another unit owns it
  int cost;        // relative CTA time per tile
  int b_fp8;       // 1: the input is the two-atom E4M3 image of a 256-wide layer output, 0: a 64-wide bf16 encoding atom
};
struct WgTable { WgUnit u[kWgUnits]; int total_cost; };

static WgTable build_wg_table() {
  WgTable t;
  ParamOffsets po = param_offsets();
  auto W = [&](int i) { return (int)po.off[2 * i]; };
  auto B = [&](int i) { return (int)po.off[2 * i + 1]; };
  auto DH = [&](int i) { return DA_H7 + 4 * (7 - i); };
  auto X = [&](int layer) { return stash_x_atom(layer, 0); };
  int n = 0;
  // cost = measured CTA-cycles per tile (tools/trace_wgrad.py, profiles/r02d_wgrad_cta_balance.txt), wide unit = 127
  t.u[n++] = WgUnit{DH(0), 256, SA_ENC, 64, kEncP, W(0), kEncP, B(0), 86, 0};
  for (int i = 1; i <= 4; ++i) t.u[n++] = WgUnit{DH(i), 256, X(i - 1), 256, 256, W(i), kW, B(i), 127, 1};
  t.u[n++] = WgUnit{DH(5), 256, SA_ENC, 64, kEncP, W(5), kW + kEncP, -1, 75, 0};
  t.u[n++] = WgUnit{DH(5), 256, X(4), 256, 256, W(5) + kEncP, kW + kEncP, B(5), 127, 1};
  t.u[n++] = WgUnit{DH(6), 256, X(5), 256, 256, W(6), kW, B(6), 127, 1};
  t.u[n++] = WgUnit{DH(7), 256, X(6), 256, 256, W(7), kW, B(7), 127, 1};
  t.u[n++] = WgUnit{DA_FEAT, 256, X(7), 256, 256, (int)po.off[T_WF], kW, (int)po.off[T_BF], 127, 1};
  t.u[n++] = WgUnit{DA_HV, 128, X(8), 256, 256, (int)po.off[T_WV], kW + kEncD, (int)po.off[T_BV], 113, 1};
  t.u[n++] = WgUnit{DA_HV, 128, SA_DENC, 64, kEncD, (int)po.off[T_WV] + kW, kW + kEncD, -1, 68, 0};
  t.total_cost = 0;
  for (int i = 0; i < kWgUnits; ++i) t.total_cost += t.u[i].cost;
  return t;
}

struct WgradParams {
  const uint8_t* stash;
  const uint8_t* dstash;
  float* grads;
  int64_t tiles;      // number of 128-sample tiles that hold real samples
  WgTable tab;
  int cta_begin[kWgUnits + 1];   // CTAs [cta_begin[u], cta_begin[u+1]) split unit u's tiles evenly
  float* partial;                // [gridDim.x][256][256] fp32 per-CTA weight-gradient partials
  long long* cta_cycles;         // diagnostic (spn_tc_set_trace): [2*cta] = unit, [2*cta+1] = cycles of the CTA's segment
  int debug;                     // SPN_WG_DEBUG bit0: skip MMAs, bit1: skip column sums, bit2: skip the copies (timing experiments)
};

__device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__global__ void __launch_bounds__(kWgThreads, 1) mlp_wgrad_kernel(const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const uint32_t bar_full = sbase + WG_BAR, bar_empty = bar_full + 8 * kWgStages;
  const uint32_t bar_acc_full = bar_empty + 8 * kWgStages, bar_acc_empty = bar_acc_full + 8;
  const uint32_t bar_conv = bar_acc_empty + 8, bar_b16free = bar_conv + 16;     // [2] each: widened B ready / consumed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + WG_BAR + 192);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1 + kWgConvWarps); }
    mbar_init(bar_acc_full, 1);
    mbar_init(bar_acc_empty, 128);
    for (int b = 0; b < 2; ++b) { mbar_init(bar_conv + 8 * b, kWgConvWarps); mbar_init(bar_b16free + 8 * b, 1); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_ptr_smem), 512); tmem_relinquish(); }
  tcgen05_fence_before_sync();
  __syncthreads();
  tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // one (unit, tile range) segment per CTA: the host gave every unit a CTA count proportional to its cost
  const int64_t T = p.tiles;
  uint32_t stage = 0, phase = 0;          // ring position: every role walks the same sequence
  uint32_t seg_phase = 0;
  uint32_t nconv = 0;                     // widened slabs so far (buffer = nconv & 1, use = nconv >> 1)
  for (int ui = 0; ui < kWgUnits; ++ui) {
    if ((int)blockIdx.x < p.cta_begin[ui] || (int)blockIdx.x >= p.cta_begin[ui + 1]) continue;
    const WgUnit u = p.tab.u[ui];
    const int64_t gu = p.cta_begin[ui + 1] - p.cta_begin[ui], ju = (int)blockIdx.x - p.cta_begin[ui];
    const int64_t t0 = T * ju / gu, t1 = T * (ju + 1) / gu;
    const int a_atoms = u.m_out / 64, b_pieces = u.b_fp8 ? 2 : 1;
    const uint32_t stage_bytes = (uint32_t)(a_atoms + b_pieces) * kWgSlabBytes;
    const int64_t nslabs = (t1 - t0) * (kTileM / kWgSlabRows);
    const long long seg_t0 = clock64();

    if (warp == 0) {
      // ---- producer: 64-sample slabs of dpre (A) and of the layer input (B).  Each of the <= 6 slab copies of a stage is
      //      issued by its own lane (one thread retires at most one cp.async.bulk per ~700 cycles, tools/bulk_rate.py)
      for (int64_t sl = 0; sl < nslabs; ++sl) {
        const int64_t tile = t0 + sl / (kTileM / kWgSlabRows);
        const int j = (int)(sl % (kTileM / kWgSlabRows));
        if (lane == 0) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          if (p.debug & 4) mbar_arrive(bar_full + 8 * stage);
          else mbar_arrive_expect_tx(bar_full + 8 * stage, stage_bytes);
        }
        __syncwarp();
        const uint32_t dstA = sbase + stage * kWgStageBytes, dstB = dstA + 4 * kWgSlabBytes;
        if (p.debug & 4) {
        } else if (lane < a_atoms) {
          const uint8_t* dsrc = p.dstash + (size_t)tile * kDstashTileBytes + (size_t)(u.d_atom + lane) * kAtomBytes + j * kWgSlabBytes;
          bulk_g2s(dstA + lane * kWgSlabBytes, dsrc, kWgSlabBytes, bar_full + 8 * stage);
        } else if (lane < a_atoms + b_pieces) {
          const int at = lane - a_atoms;
          const uint8_t* isrc = p.stash + (size_t)tile * kStashTileBytes + (size_t)(u.in_atom + at) * kAtomBytes + j * kWgSlabBytes;
          bulk_g2s(dstB + at * kWgSlabBytes, isrc, kWgSlabBytes, bar_full + 8 * stage);
        }
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
    } else if (warp == 1) {
      // ---- MMA issuer (whole warp, uniform control flow: tc_common.cuh elect_one): D[out, in] += dpre^T[out, k] . in[k, in],  k = sample
      const uint32_t idesc = make_idesc(128, u.n_in, 1, 1);
      const uint32_t tmem_u = uniform_u32(tmem_base);
      const int halves = u.m_out / 128;
      mbar_wait(bar_acc_empty, seg_phase ^ 1);     // previous segment's accumulators flushed
      tcgen05_fence_after_sync();
      uint32_t accumulate = 0;
      for (int64_t sl = 0; sl < nslabs; ++sl) {
        mbar_wait(bar_full + 8 * stage, phase);
        uint32_t aB = sbase + stage * kWgStageBytes + 4 * kWgSlabBytes;
        if (u.b_fp8) {
          const uint32_t cb = nconv & 1u;
          mbar_wait(bar_conv + 8 * cb, (nconv >> 1) & 1u);      // the converter warps have widened this slab
          aB = sbase + WG_B16 + cb * (4 * kWgSlabBytes);
        }
        tcgen05_fence_after_sync();
        const uint64_t a_desc = make_smem_desc(sbase + stage * kWgStageBytes, kWgSlabBytes, 1024);
        const uint64_t b_desc = make_smem_desc(aB, kWgSlabBytes, 1024);
        if (!(p.debug & 1) && elect_one()) {
#pragma unroll
          for (int k = 0; k < kWgSlabRows / 16; ++k) {
            // 16 samples = 2 KB inside every 64-feature block; the second 128 output rows are two blocks (16 KB) further
            umma_bf16(tmem_u, a_desc + (uint64_t)(128 * k), b_desc + (uint64_t)(128 * k), idesc, k ? 1u : accumulate);
            if (halves == 2) umma_bf16(tmem_u + 256u, a_desc + (uint64_t)(1024 + 128 * k), b_desc + (uint64_t)(128 * k), idesc, k ? 1u : accumulate);
          }
        }
        accumulate = 1;
        if (elect_one()) {
          umma_commit(bar_empty + 8 * stage);
          if (u.b_fp8) umma_commit(bar_b16free + 8 * (nconv & 1u));
        }
        if (u.b_fp8) ++nconv;
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
      if (nslabs > 0 && elect_one()) umma_commit(bar_acc_full);
    } else {
      // ---- converter warps (2..9) + column sums (bias gradient, warps 2..5) from the dpre slabs, then the accumulator flush (warps 2..5)
      const int tid = threadIdx.x - 64;           // 0..255
      const int q = warp & 3;
      const int cw = warp - 2;                    // 0..7
      const int j = tid & 31, rg = (tid >> 5) & 3;
      float bs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const bool do_bias = cw < 4 && u.b_off >= 0 && 8 * j < u.m_out;
      // widening: lane = (row rr of an 8-row block, 16-byte chunk jl of a 64-byte run); warp cw, iteration i -> combination cw * 4 + i
      // of (atom hh, upper/lower 64 bytes jh, row block rb).  Stores cover all 8 chunk positions of two rows per quarter warp.
      const int rr = lane >> 2, jl = lane & 3;
      for (int64_t sl = 0; sl < nslabs; ++sl) {
        mbar_wait(bar_full + 8 * stage, phase);
        if (do_bias && !(p.debug & 2)) {
          // thread = (row group rg of 8 samples) x (16-byte chunk j = 8 output features): 8 x LDS.128 per 32 rows
          const uint32_t abase = sbase + stage * kWgStageBytes + (j >> 3) * kWgSlabBytes;
#pragma unroll
          for (int sub = 0; sub < kWgSlabRows / 32; ++sub) {
            uint4 w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t addr = abase + sw128_off((uint32_t)(sub * 32 + rg * 8 + i), (uint32_t)(j & 7));
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[i].x), "=r"(w[i].y), "=r"(w[i].z), "=r"(w[i].w) : "r"(addr));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              bs[0] += __uint_as_float(w[i].x << 16); bs[1] += __uint_as_float(w[i].x & 0xffff0000u);
              bs[2] += __uint_as_float(w[i].y << 16); bs[3] += __uint_as_float(w[i].y & 0xffff0000u);
              bs[4] += __uint_as_float(w[i].z << 16); bs[5] += __uint_as_float(w[i].z & 0xffff0000u);
              bs[6] += __uint_as_float(w[i].w << 16); bs[7] += __uint_as_float(w[i].w & 0xffff0000u);
            }
          }
        }
        if (u.b_fp8) {
          const uint32_t cb = nconv & 1u;
          mbar_wait(bar_b16free + 8 * cb, ((nconv >> 1) & 1u) ^ 1u);   // the MMAs of the slab before last have read this buffer
          const uint32_t raw = sbase + stage * kWgStageBytes + 4 * kWgSlabBytes;
          const uint32_t dst = sbase + WG_B16 + cb * (4 * kWgSlabBytes);
#pragma unroll
          for (int i = 0; i < 32 / kWgConvWarps; ++i) {
            const int combo = cw * (32 / kWgConvWarps) + i;
            const int hh = combo & 1, jh = (combo >> 1) & 1, rb = combo >> 2;
            const uint32_t r = (uint32_t)(rb * 8 + rr);
            const uint4 v = lds128u(raw + (uint32_t)hh * kWgSlabBytes + r * 128u + ((uint32_t)((jh * 4 + jl) ^ rr) << 4));
            uint32_t o[8];
            e4m3x4_to_bf16x4(v.x, o[0], o[1]); e4m3x4_to_bf16x4(v.y, o[2], o[3]);
            e4m3x4_to_bf16x4(v.z, o[4], o[5]); e4m3x4_to_bf16x4(v.w, o[6], o[7]);
            // features 64 hh + 16 jl + 128 jh ... + 15: 64-feature block hh + 2 jh, chunks 2 jl and 2 jl + 1 of row r
            const uint32_t row_a = dst + (uint32_t)(hh + 2 * jh) * kWgSlabBytes + r * 128u;
            sts128(row_a + ((uint32_t)((2 * jl) ^ rr) << 4), o[0], o[1], o[2], o[3]);
            sts128(row_a + ((uint32_t)((2 * jl + 1) ^ rr) << 4), o[4], o[5], o[6], o[7]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_conv + 8 * cb);
          ++nconv;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * stage);
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
      if (cw >= 4) { seg_phase ^= 1; continue; }      // the flush below is the first four warps' (one per TMEM lane quarter)
      if (do_bias) {
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(p.grads + u.b_off + 8 * j + e, bs[e]);
      }
      if (t0 < t1) mbar_wait(bar_acc_full, seg_phase);
      tcgen05_fence_after_sync();
      float* part = p.partial + (size_t)blockIdx.x * 256 * 256;
      for (int h = 0; h < u.m_out / 128; ++h) {
        const int out = h * 128 + q * 32 + lane;
        float4* prow = reinterpret_cast<float4*>(part + (size_t)out * 256);
        for (int cb = 0; cb < u.n_in / 32; ++cb) {
          uint32_t v[32];
          if (t0 < t1) {
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)h * 256u + cb * 32, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) v[jj] = 0u;
          }
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)   // one full 128-byte line per thread and batch, no atomics
            prow[cb * 8 + jj] = make_float4(__uint_as_float(v[4 * jj]), __uint_as_float(v[4 * jj + 1]),
                                            __uint_as_float(v[4 * jj + 2]), __uint_as_float(v[4 * jj + 3]));
        }
      }
      tcgen05_fence_before_sync();
      mbar_arrive(bar_acc_empty);
      if (p.cta_cycles && threadIdx.x == 64) { p.cta_cycles[2 * blockIdx.x] = ui; p.cta_cycles[2 * blockIdx.x + 1] = clock64() - seg_t0; }
    }
    seg_phase ^= 1;
  }
  tcgen05_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// dW[unit] += sum over the unit's CTAs of their partials (deterministic, coalesced; replaces ~10^7 atomics/launch)
__global__ void __launch_bounds__(256) mlp_wgrad_reduce_kernel(const WgradParams p) {
  const WgUnit u = p.tab.u[blockIdx.y];
  const int c0 = p.cta_begin[blockIdx.y], c1 = p.cta_begin[blockIdx.y + 1];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // over m_out * n_in
  if (idx >= u.m_out * u.n_in) return;
  const int o = idx / u.n_in, i = idx % u.n_in;
  if (i >= u.n_valid) return;
  float acc = 0.f;
  for (int c = c0; c < c1; ++c) acc += p.partial[(size_t)c * 256 * 256 + (size_t)o * 256 + i];
  p.grads[u.w_off + (int64_t)o * u.ld + i] += acc;
}

// =====================================================================================================
// rgb / sigma heads: dWr[3,128] += d_rgb^T hv,  dbr,  dWa[256] += d_sigma h7,  dba
// =====================================================================================================
// 8 warps per block, warp w owns rows 16w..16w+15 of a tile.  Lanes 0..15 read h7 — the E4M3 stash image, two 128-byte rows per
// sample, one 16-byte chunk (16 features) per lane — and accumulate d alpha_linear.weight; lanes 16..31 read hv (bf16, one
// 16-byte chunk = 8 features per lane) and accumulate d rgb_linear.weight: every load instruction of a warp covers whole
// 128-byte lines.  Partials live in registers across the block's tiles and are combined through shared memory once, then 643
// atomics per block.
__global__ void __launch_bounds__(256, 2) mlp_heads_wgrad_kernel(const uint8_t* __restrict__ stash,
                                                              const float* __restrict__ d_raw, int64_t m,
                                                              int64_t tiles, float* __restrict__ g_wr,
                                                              float* __restrict__ g_br, float* __restrict__ g_wa,
                                                              float* __restrict__ g_ba) {
  __shared__ float4 s_d[kTileM];
  __shared__ float s_red[8][32][8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // acc: lanes 0..15 -> d Wa for 16 features; lanes 16..31 -> d Wr rows 0 and 1 for 8 features (row 2 in acc2)
  float acc[16], acc2[8];
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) acc2[e] = 0.f;
  float sb[4] = {0, 0, 0, 0};
  const int l16 = lane & 15;
  const bool is_h7 = lane < 16;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    __syncthreads();
    if (tid < kTileM) {
      const int64_t row = tile * kTileM + tid;
      s_d[tid] = row < m ? *reinterpret_cast<const float4*>(d_raw + row * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const uint8_t* st = stash + (size_t)tile * kStashTileBytes;
    const uint8_t* src = st + (size_t)(is_h7 ? stash_x_atom(7, l16 >> 3) : SA_HV + (l16 >> 3)) * kAtomBytes;   // lane -> (atom, chunk l16 & 7)
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {       // 8 rows in flight per lane
      uint4 qq[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        qq[i] = __ldg(reinterpret_cast<const uint4*>(src + sw128_off((uint32_t)(warp * 16 + half * 8 + i), (uint32_t)(lane & 7))));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 d = s_d[warp * 16 + half * 8 + i];
        const uint32_t w4[4] = {qq[i].x, qq[i].y, qq[i].z, qq[i].w};
        if (is_h7) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            uint32_t lo, hi;
            e4m3x4_to_bf16x4(w4[e], lo, hi);
            acc[4 * e + 0] = fmaf(d.w, __uint_as_float(lo << 16), acc[4 * e + 0]);
            acc[4 * e + 1] = fmaf(d.w, __uint_as_float(lo & 0xffff0000u), acc[4 * e + 1]);
            acc[4 * e + 2] = fmaf(d.w, __uint_as_float(hi << 16), acc[4 * e + 2]);
            acc[4 * e + 3] = fmaf(d.w, __uint_as_float(hi & 0xffff0000u), acc[4 * e + 3]);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float h0 = __uint_as_float(w4[e] << 16), h1 = __uint_as_float(w4[e] & 0xffff0000u);
            acc[2 * e] = fmaf(d.x, h0, acc[2 * e]); acc[2 * e + 1] = fmaf(d.x, h1, acc[2 * e + 1]);
            acc[8 + 2 * e] = fmaf(d.y, h0, acc[8 + 2 * e]); acc[8 + 2 * e + 1] = fmaf(d.y, h1, acc[8 + 2 * e + 1]);
            acc2[2 * e] = fmaf(d.z, h0, acc2[2 * e]); acc2[2 * e + 1] = fmaf(d.z, h1, acc2[2 * e + 1]);
          }
        }
        if (lane == 15) { sb[0] += d.x; sb[1] += d.y; sb[2] += d.z; sb[3] += d.w; }
      }
    }
  }
  // block reduction over the 8 warps, one quantity at a time through the same shared buffer; lanes [l0, l0 + 16) own dst[off(l) + e]
  auto reduce8 = [&](const float* v, float* dst, int l0, auto off) {
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e) s_red[warp][lane][e] = v[e];
    __syncthreads();
    if (warp == 0 && lane >= l0 && lane < l0 + 16) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) a += s_red[w][lane][e];
        atomicAdd(dst + off(lane - l0) + e, a);
      }
    }
  };
  // d alpha_linear.weight [256]: lane (half, chunk) holds features stash_x_feature(half, chunk) .. + 15
  reduce8(acc, g_wa, 0, [](int l) { return stash_x_feature(l >> 3, l & 7); });
  reduce8(acc + 8, g_wa, 0, [](int l) { return stash_x_feature(l >> 3, l & 7) + 8; });
  reduce8(acc, g_wr, 16, [](int l) { return 8 * l; });                        // d rgb_linear.weight [3][128]
  reduce8(acc + 8, g_wr + kWV, 16, [](int l) { return 8 * l; });
  reduce8(acc2, g_wr + 2 * kWV, 16, [](int l) { return 8 * l; });
  {
    const float v[8] = {sb[0], sb[1], sb[2], sb[3], 0.f, 0.f, 0.f, 0.f};
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e) s_red[warp][lane][e] = v[e];
    __syncthreads();
    if (tid < 4) {
      float a = 0.f;
      for (int w = 0; w < 8; ++w) a += s_red[w][15][tid];
      atomicAdd(tid < 3 ? g_br + tid : g_ba, a);
    }
  }
}

// ---- diagnostic: the weight-gradient kernel's E4M3 -> bf16 widening, four codes per thread (tests: all 256 codes) ----
__global__ void e4m3_decode_kernel(const uint8_t* __restrict__ codes, uint16_t* __restrict__ out, int n) {
  const int i = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n) return;
  uint32_t v = 0;
  for (int e = 0; e < 4 && i + e < n; ++e) v |= (uint32_t)codes[i + e] << (8 * e);
  uint32_t lo, hi;
  e4m3x4_to_bf16x4(v, lo, hi);
  const uint16_t r[4] = {(uint16_t)(lo & 0xffff), (uint16_t)(lo >> 16), (uint16_t)(hi & 0xffff), (uint16_t)(hi >> 16)};
  for (int e = 0; e < 4 && i + e < n; ++e) out[i + e] = r[e];
}
int tc_e4m3_decode(const uint8_t* codes, uint16_t* out, int n, cudaStream_t st) {
  SPN_CHECK_ARG(codes && out && n > 0, "spn_tc_e4m3_decode: bad arguments");
  e4m3_decode_kernel<<<(n / 4 + 128) / 128, 128, 0, st>>>(codes, out, n);
  SPN_LAUNCH_CHECK("e4m3_decode_kernel");
  return SPN_OK;
}

// =====================================================================================================
int mlp_tc_bwd(const void* packed, const void* stash, const float* d_raw, int64_t m, float* grads, void* ws,
               cudaStream_t st) {
  int rc = check_arch_bwd();
  if (rc != SPN_OK) return rc;
  SPN_CHECK_ARG(ws, "spn_mlp_bwd: BF16 mode needs a workspace of spn_mlp_bwd_workspace_bytes(m)");
  SPN_CHECK_ARG((((uintptr_t)packed | (uintptr_t)stash | (uintptr_t)ws | (uintptr_t)d_raw) & 15) == 0,
                "spn_mlp_bwd: buffers must be 16-byte aligned");
  static const WgTable wg_tab = build_wg_table();
  static bool attr_set = false;
  if (!attr_set) {
    SPN_CUDA(cudaFuncSetAttribute(mlp_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    SPN_CUDA(cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes));
    attr_set = true;
  }
  const int64_t tiles = (m + kTileM - 1) / kTileM;
  ParamOffsets po = param_offsets();
  // 1. dgrad chain
  DgradParams dp;
  dp.packed = (const uint8_t*)packed; dp.stash = (const uint8_t*)stash; dp.d_raw = d_raw;
  dp.dstash = (uint8_t*)ws; dp.m = m; dp.num_quads = (int)((tiles + 3) / 4);
  const int pairs = sm_count() / 2;
  int grid = 2 * (dp.num_quads < pairs ? dp.num_quads : pairs);
  prof_begin(PROF_MLP_DGRAD, st);
  mlp_dgrad_kernel<<<grid, kPairThreads, kSmemBytes, st>>>(dp);
  prof_end(PROF_MLP_DGRAD, st);
  SPN_LAUNCH_CHECK("mlp_dgrad_kernel");
  // 2. weight / bias gradients of the ten wide layers
  WgradParams wp;
  wp.stash = (const uint8_t*)stash; wp.dstash = (const uint8_t*)ws; wp.grads = grads; wp.tiles = tiles; wp.tab = wg_tab;
  // Every CTA owns exactly one (unit, tile range).  CTAs are handed out greedily to the unit whose CTAs would otherwise
  // run longest (cost / CTAs), which minimises the slowest CTA; a unit never gets more CTAs than tiles.
  int target = sm_count();
  {
    const int64_t enough = (int64_t)wg_tab.total_cost * tiles / (2 * 127);   // ~2 tiles of a big unit per CTA at least
    if (enough < target) target = (int)enough;
    if (target < kWgUnits) target = kWgUnits;
  }
  int per_unit[kWgUnits];
  for (int ui = 0; ui < kWgUnits; ++ui) per_unit[ui] = 1;
  for (int assigned_n = kWgUnits; assigned_n < target; ++assigned_n) {
    int best = -1;
    double worst = -1.0;
    for (int ui = 0; ui < kWgUnits; ++ui) {
      if (per_unit[ui] >= tiles) continue;
      const double tpc = (double)wg_tab.u[ui].cost / per_unit[ui];
      if (tpc > worst) { worst = tpc; best = ui; }
    }
    if (best < 0) break;
    ++per_unit[best];
  }
  int assigned = 0;
  wp.cta_begin[0] = 0;
  for (int ui = 0; ui < kWgUnits; ++ui) {
    assigned += per_unit[ui];
    wp.cta_begin[ui + 1] = assigned;
  }
  const int wgrid = assigned;
  wp.partial = reinterpret_cast<float*>((uint8_t*)ws + (size_t)quad_tiles(m) * kDstashTileBytes);
  static const int wg_debug = getenv("SPN_WG_DEBUG") ? atoi(getenv("SPN_WG_DEBUG")) : 0;
  wp.debug = wg_debug;
  wp.cta_cycles = tc_get_trace();
  prof_begin(PROF_MLP_WGRAD, st);
  mlp_wgrad_kernel<<<wgrid, kWgThreads, kWgSmemBytes, st>>>(wp);
  SPN_LAUNCH_CHECK("mlp_wgrad_kernel");
  mlp_wgrad_reduce_kernel<<<dim3(256, kWgUnits), 256, 0, st>>>(wp);
  prof_end(PROF_MLP_WGRAD, st);
  SPN_LAUNCH_CHECK("mlp_wgrad_reduce_kernel");
  // 3. heads
  const int hmax = 4 * sm_count();                    // 4 resident blocks per SM; every block ends with 643 atomics
  int hgrid = (int)(tiles < hmax ? tiles : hmax);
  mlp_heads_wgrad_kernel<<<hgrid, 256, 0, st>>>((const uint8_t*)stash, d_raw, m, tiles, grads + po.off[T_WR],
                                                grads + po.off[T_BR], grads + po.off[T_WA], grads + po.off[T_BA]);
  SPN_LAUNCH_CHECK("mlp_heads_wgrad_kernel");
  return SPN_OK;
}

}  // namespace spn
