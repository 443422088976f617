// Layout constants shared by the tcgen05 MLP kernels (forward: mlp_tc.cu, backward: mlp_tc_bwd.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace spn {

// ---- packed weight image (spn_mlp_pack_weights) ------------------------------------------------------
constexpr int kChunkBig = 256 * 128;   // [256 rows x 64 bf16] SWIZZLE_128B = 32 KB
constexpr int kChunkV = 128 * 128;     // [128 rows x 64 bf16] = 16 KB
constexpr int kFwdChunks = 39;
constexpr int kBwdChunks = 34;
constexpr size_t kFwdBytes = 34 * (size_t)kChunkBig + 5 * (size_t)kChunkV;
constexpr size_t kBwdBytes = (size_t)kBwdChunks * kChunkBig;
// fp32 constants the epilogues read (float offsets inside the constant block that follows the chunks)
constexpr int C_B = 0;          // b0..b7 [8][256]
constexpr int C_BF = 2048;      // feature bias [256]
constexpr int C_BV = 2304;      // views bias [128]
constexpr int C_WA = 2432;      // alpha weight [256]
constexpr int C_BA = 2688;      // alpha bias (padded to 4)
constexpr int C_WR = 2692;      // rgb weight [3][128]
constexpr int C_BR = 3076;      // rgb bias (padded to 4)
constexpr int kConstFloats = 3080;
constexpr size_t kPackedBytes = kFwdBytes + kBwdBytes + kConstFloats * sizeof(float);

// ---- tile geometry ------------------------------------------------------------------------------------
constexpr int kTileM = 128;
constexpr int kAtomBytes = kTileM * 128;        // [128 rows x 64 bf16] swizzle atom = 16 KB
constexpr int kActBytes = 4 * kAtomBytes;       // 256-wide activation tile = 64 KB
constexpr int kStages = 3;
constexpr int kEpiWarps = 8;                               // epilogue warps per tile (2 column halves x 4 lane quarters)
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 32 * (2 + 2 * kEpiWarps);         // producer + MMA issuer + 16 epilogue warps = 576
constexpr int kFwdEpiWarps = 2 * kEpiWarps;                // forward: all 16 epilogue warps serve tile slot 0, then tile slot 1
constexpr int kFwdEpiThreads = kFwdEpiWarps * 32;
constexpr int SM_ACT = 0;                                  // 2 tiles x 64 KB
constexpr int SM_RING = 2 * kActBytes;                     // 3 x 32 KB
constexpr int SM_BAR = SM_RING + kStages * kChunkBig;      // mbarriers (<= 24 x 8 B)
constexpr int SM_TMEMPTR = SM_BAR + 240;                   // TMEM base address written by tcgen05.alloc
constexpr int SM_BIAS = SM_BAR + 256;                      // 2 x 256 floats: the current layer's bias row per tile slot
constexpr int SM_WA = SM_BIAS + 2048;                      // forward only: sigma-head weights, 256 x bf16
constexpr int kSmemBytes = SM_WA + 512;                    // = 232192 <= 232448; starts 1024-byte aligned (checked)
// CTA-pair kernels (cta_group::2): every CTA holds HALF of each weight chunk (its 128 of the 256 output rows), so the
// same 96 KB ring holds 6 half-chunks: a whole layer (4) stays resident for both tile slots + 2 of the next layer
constexpr int kPairThreads = kThreads;
constexpr int kSlots = 3;            // ring slots: one GROUP (two half-chunks) each
constexpr int kSlotBytes = 32768;

// ---- per-tile forward stash (training), 16 KB atoms -------------------------------------------------------------
//   atom 0        gamma(pts) (63 + pad), bf16 SWIZZLE_128B image [128 rows x 64]
//   atoms 1..18   the nine 256-wide layer outputs h0..h7, feature as E4M3 (fp8) — two atoms per layer, atom h = the 128 output
//                 features the forward epilogue handles in accumulator half h: row r = 128 bytes, 16-byte chunk j' stored at
//                 position j' ^ (r & 7), chunk j' = features 64 h + 16 (j' & 3) + 128 (j' >> 2) ... + 15, one byte each.
//                 The weight-gradient GEMM is the only reader: it widens them to bf16 in shared memory (exact) for its MMAs.  Halves the bytes of the training step's largest HBM stream (round 1 stashed bf16).
//   atoms 19-20   hv (128 wide), bf16 image        atom 21   gamma(viewdir) (27 + pad), bf16 image
// followed by ReLU masks: 9 slots (h0..h7, hv) x 128 rows x 8 words; word w covers columns 32w..32w+31 with column
// 32w + c at bit relu_mask_bit(c) (pairs are pushed as packed bf16x2 words, see relu_mask_push in mlp_tc.cu)
__host__ __device__ constexpr int relu_mask_bit(int c) { return ((c & 1) << 4) | (c >> 1); }
constexpr int kStashAtoms = 22;
constexpr int SA_ENC = 0, SA_X0 = 1, SA_HV = 19, SA_DENC = 21;
__host__ __device__ constexpr int stash_x_atom(int layer, int half) { return SA_X0 + 2 * layer + half; }   // layer 0..7 = h0..h7, 8 = feature
__host__ __device__ constexpr int stash_x_feature(int half, int chunk) { return 64 * half + 16 * (chunk & 3) + 128 * (chunk >> 2); }
constexpr size_t kStashMaskOff = (size_t)kStashAtoms * kAtomBytes;
constexpr size_t kStashTileBytes = kStashMaskOff + 9 * 128 * 32;   // 397312 B per 128 samples (round 1: 692224)

// ---- per-tile backward stash (dgrad -> wgrad): d(pre-activation) as bf16 images -----------------------------
//   atoms 0-1  d_hv (128 wide)    atoms 2-5  d_feat     atoms 6+4*(7-i) .. : d_h{i} for i = 7..0
constexpr int kDstashAtoms = 38;
constexpr int DA_HV = 0, DA_FEAT = 2, DA_H7 = 6;
constexpr size_t kDstashTileBytes = (size_t)kDstashAtoms * kAtomBytes;   // 622592 B per 128 samples

}  // namespace spn
