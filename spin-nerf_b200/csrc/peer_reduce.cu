// Gradient exchange of the ray-sharded train step over NVLink peer memory, fused with the optimiser
// (reference: none — DS_NeRF is single-GPU, run_nerf.py:39; the update itself is run_nerf.py:433-434,
// 1611-1622 as in adam_kernel).
//
// STATUS: opt-in (SPN_P2P_ALLREDUCE=1, spin-nerf_b200/peer.py).  Written after the round's GPU budget was spent: it compiles for
// sm_100a and its host side is exercised without a GPU, but it has NOT run on hardware yet; the default multi-GPU path is the
// NCCL all-reduce + flat Adam of Trainer.apply_gradients.
//
// Every rank owns one cudaMalloc'ed, IPC-exported region  [ flags | G: its flat gradients (both networks) | R: reduced slice ]
// mapped into every other rank of the box.  One optimisation step is two kernels on each rank:
//
//   peer_reduce_slice_kernel   signal "my G of step k is complete" into every peer's flag row 0, wait for all peers' signals,
//                              then sum slice r = [r n/W, (r+1) n/W) of all W gradient vectors (W-1 of them through NVLink
//                              loads) into this rank's R                                            — the reduce-scatter
//   peer_gather_adam_kernel    signal "my R of step k is complete" into every peer's flag row 1, wait likewise, then read each
//                              element's reduced gradient from its owner's R and apply Adam to the local replica of both
//                              networks                                                             — all-gather + optimiser
//
// Per rank and step (W-1)/W * 4.8 MB cross NVLink twice instead of NCCL's ring / tree schedule plus two Adam launches, and the
// optimiser reads the reduced gradients straight from peer memory.  No buffer is double-buffered: a rank overwrites G (next
// step's backward) only after its gather kernel has passed barrier 1 of step k, which every peer signals after its reduce
// kernel — the only reader of foreign G — has finished; it overwrites R (next step's reduce kernel) only after barrier 0 of
// step k+1, which every peer signals after its gather kernel of step k — the only reader of foreign R — has finished.
// Flags are monotonically increasing step numbers, written with st.release.sys after a system-scope fence and polled with
// ld.acquire.sys; peer data is read with ld.global.cv (never from a stale L1 line).  A poll that lasts longer than ~4 s traps
// (a dead peer becomes a launch error instead of a hang).
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace spn {
namespace {

constexpr int kMaxPeers = 8;
constexpr size_t kFlagBytes = 4096;            // 2 rows x kMaxPeers flags, one 128-byte line each
constexpr int kThreads = 256;

struct PeerBases {
  uint8_t* base[kMaxPeers];                    // region of every rank as mapped in THIS process (own region at [rank])
};

__host__ __device__ inline size_t region_g_off() { return kFlagBytes; }
__host__ __device__ inline size_t region_r_off(int64_t n) { return kFlagBytes + (((size_t)n * 4 + 255) & ~(size_t)255); }

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// A peer that never arrives must not hang the GPU for good (that would cost the box), but ranks DO arrive late by seconds when
// rank 0 alone renders a test set or writes a checkpoint between steps: tools/run_nerf_fused.py therefore ends every rank-0-only
// section with a process-group barrier when this path is on, and the device-side limit is two minutes, not the 4 s of round 1.
constexpr unsigned long long kPeerTimeoutNs = 120ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ uint32_t* flag_of(uint8_t* base, int row, int writer) {
  return reinterpret_cast<uint32_t*>(base + ((size_t)row * kMaxPeers + writer) * 128);
}

// Cross-GPU barrier `row` of step `epoch`, executed by every block (each block needs the guarantee; only block 0 signals).
__device__ void peer_barrier(const PeerBases& pb, int world, int rank, int row, uint32_t epoch) {
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      __threadfence_system();                  // this GPU's earlier kernels' writes (G or R) visible system-wide
      for (int p = 0; p < world; ++p)
        if (p != rank) st_release_sys(flag_of(pb.base[p], row, rank), epoch);
    }
    const unsigned long long t0 = globaltimer_ns();
    for (int p = 0; p < world; ++p) {
      if (p == rank) continue;
      const uint32_t* f = flag_of(pb.base[rank], row, p);
      // epochs are compared as a wrapping distance so that a 32-bit step counter may overflow
      while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
        if (globaltimer_ns() - t0 > kPeerTimeoutNs) __trap();   // 120 s: a rank may legitimately be minutes late only if the caller forgot its barrier
        __nanosleep(200);
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads)
peer_reduce_slice_kernel(PeerBases pb, int world, int rank, uint32_t epoch, int64_t n) {
  peer_barrier(pb, world, rank, 0, epoch);
  const int64_t n4 = n >> 2;                                       // n is a multiple of 4 (checked on the host)
  const int64_t lo = n4 * rank / world, hi = n4 * (rank + 1) / world;
  float4* out = reinterpret_cast<float4*>(pb.base[rank] + region_r_off(n));
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < world; ++p) {                              // fixed summation order: identical result on every rank
      const float4 g = __ldcv(reinterpret_cast<const float4*>(pb.base[p] + region_g_off()) + i);
      acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
    }
    out[i] = acc;
  }
}

struct AdamNet {
  float* p; float* m; float* v;
};

__global__ void __launch_bounds__(kThreads)
peer_gather_adam_kernel(PeerBases pb, int world, int rank, uint32_t epoch, int64_t n, int64_t n_params, int64_t stride,
                        AdamNet net0, AdamNet net1, float lr_bc1, float bc2_sqrt, float b1, float b2, float eps,
                        float gscale) {
  peer_barrier(pb, world, rank, 1, epoch);
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    // owner of float4 i: the rank whose slice [n4 r / W, n4 (r+1) / W) contains it
    int owner = (int)(((i + 1) * world - 1) / n4);
    while (owner > 0 && n4 * owner / world > i) --owner;
    while (owner < world - 1 && n4 * (owner + 1) / world <= i) ++owner;
    const float4 g4 = __ldcv(reinterpret_cast<const float4*>(pb.base[owner] + region_r_off(n)) + i);
    const float gs[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t e = 4 * i + k;                                 // element of the [coarse | pad | fine | pad] gradient layout
      const int which = e >= stride ? 1 : 0;
      const int64_t j = e - which * stride;
      if (j >= n_params) continue;                                 // padding
      const AdamNet& a = which ? net1 : net0;
      const float gi = gs[k] * gscale;
      const float mi = a.m[j] + (gi - a.m[j]) * (1.0f - b1);       // same arithmetic as adam_kernel (ops_render.cu)
      const float vi = a.v[j] * b2 + (1.0f - b2) * gi * gi;
      a.m[j] = mi; a.v[j] = vi;
      a.p[j] = a.p[j] - lr_bc1 * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    }
  }
}

}  // namespace
}  // namespace spn

using namespace spn;

extern "C" size_t spn_peer_region_bytes(int64_t n_floats) {
  return region_r_off(n_floats) + (((size_t)n_floats * 4 + 255) & ~(size_t)255);
}

extern "C" int spn_peer_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64) {
  SPN_CHECK_ARG(dev_ptr && handle64 && bytes >= kFlagBytes, "spn_peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  SPN_CUDA(cudaMalloc(&p, bytes));
  SPN_CUDA(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  SPN_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(handle64, &h, 64);
  SPN_CUDA(cudaDeviceSynchronize());
  *dev_ptr = p;
  return SPN_OK;
}

extern "C" int spn_peer_open(const unsigned char* handle64, void** dev_ptr) {
  SPN_CHECK_ARG(handle64 && dev_ptr, "spn_peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  SPN_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return SPN_OK;
}

extern "C" int spn_peer_close(void* dev_ptr) {
  if (dev_ptr) SPN_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return SPN_OK;
}

extern "C" int spn_peer_free(void* dev_ptr) {
  if (dev_ptr) SPN_CUDA(cudaFree(dev_ptr));
  return SPN_OK;
}

extern "C" void* spn_peer_grad_ptr(void* region) { return region ? (uint8_t*)region + region_g_off() : nullptr; }

extern "C" int spn_peer_allreduce_adam(void* const* regions, int world, int rank, unsigned int epoch, int64_t n_floats,
                                       int64_t n_params, int64_t stride, float* param_c, float* m_c, float* v_c,
                                       float* param_f, float* m_f, float* v_f, float lr, float beta1, float beta2,
                                       float eps, int step, float grad_scale, void* stream) {
  SPN_CHECK_ARG(regions && world >= 2 && world <= kMaxPeers && rank >= 0 && rank < world && step >= 1 && epoch >= 1,
                "spn_peer_allreduce_adam: bad arguments (world=%d rank=%d)", world, rank);
  SPN_CHECK_ARG(n_floats > 0 && n_floats % 4 == 0 && stride % 4 == 0 && n_params <= stride && 2 * stride <= n_floats,
                "spn_peer_allreduce_adam: gradient layout (n=%lld stride=%lld params=%lld)", (long long)n_floats,
                (long long)stride, (long long)n_params);
  SPN_CHECK_ARG(param_c && m_c && v_c && param_f && m_f && v_f, "spn_peer_allreduce_adam: null optimiser buffers");
  PeerBases pb;
  for (int p = 0; p < kMaxPeers; ++p) pb.base[p] = p < world ? (uint8_t*)regions[p] : nullptr;
  for (int p = 0; p < world; ++p) SPN_CHECK_ARG(pb.base[p], "spn_peer_allreduce_adam: region %d not mapped", p);
  cudaStream_t st = as_stream(stream);
  const int grid = sm_count() < 64 ? sm_count() : 64;            // a few blocks saturate NVLink; all are co-resident
  peer_reduce_slice_kernel<<<grid, kThreads, 0, st>>>(pb, world, rank, epoch, n_floats);
  SPN_LAUNCH_CHECK("peer_reduce_slice_kernel");
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2s = sqrtf(1.0f - powf(beta2, (float)step));
  AdamNet a0{param_c, m_c, v_c}, a1{param_f, m_f, v_f};
  peer_gather_adam_kernel<<<grid, kThreads, 0, st>>>(pb, world, rank, epoch, n_floats, n_params, stride, a0, a1, lr / bc1,
                                                     bc2s, beta1, beta2, eps, grad_scale);
  SPN_LAUNCH_CHECK("peer_gather_adam_kernel");
  return SPN_OK;
}
