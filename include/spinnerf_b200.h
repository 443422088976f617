/* spinnerf_b200 — C ABI of the B200-native volumetric-rendering hot path of SPIn-NeRF.
 *
 * Every entry point replaces a PyTorch-op sequence of the reference (file:line cited per
 * function, paths relative to the reference repo).  The reference has no FFI of its own
 * for this path (its only plugin seam is `from run_nerf_helpers import *`,
 * DS_NeRF/run_nerf.py:23); the calling convention below follows the one native precedent
 * in the tree, torchsearchsorted (DS_NeRF/torchsearchsorted/src/cuda/searchsorted_cuda_wrapper.cpp:5-17):
 * caller-allocated outputs, contiguous device buffers, launch on the caller's stream, no sync.
 *
 * Conventions
 *   - all data pointers are DEVICE pointers unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - return value: 0 = OK, negative = error (SPN_E_*); spn_last_error() gives the text
 *     (thread-local).  Nothing throws, nothing allocates device memory, nothing syncs —
 *     except the *_host entry points, which own their H2D/D2H copies and synchronise.
 *   - fp32 row-major everywhere; int64 for sample_pdf indices (torch.searchsorted dtype).
 *   - re-entrant per stream for the compute entry points.  The DIAGNOSTIC switches (spn_profile_enable / _read,
 *     spn_launch_count, spn_tc_set_trace) are process-wide and unsynchronised: one measuring thread at a time.
 */
#ifndef SPINNERF_B200_H
#define SPINNERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPN_VERSION 100

enum {
  SPN_OK = 0,
  SPN_E_ARG = -1,      /* bad shape / null pointer / unsupported option            */
  SPN_E_CUDA = -2,     /* a CUDA runtime call or launch failed                      */
  SPN_E_ARCH = -3,     /* device is not sm_100 (tcgen05 kernels need a B200)        */
};

/* MLP arithmetic.  BF16 = tcgen05.mma kind::f16, bf16 operands, fp32 accumulate in TMEM.
 * FP32 = CUDA-core fp32 GEMMs that materialise activations (bit-faithful-to-tolerance mode
 * used for the tight parity gate; not the fast path). */
enum { SPN_PREC_BF16 = 0, SPN_PREC_FP32 = 1 };

/* flags for the render entry points */
enum {
  SPN_F_LINDISP = 1,        /* sample linearly in disparity       run_nerf.py:649-652 */
  SPN_F_WHITE_BKGD = 2,     /* rgb += 1 - acc                     helpers:394-395      */
  SPN_F_DETACH_WEIGHTS = 4, /* rgb_map uses weights.detach()      helpers:385-388      */
  SPN_F_PERTURB = 8,        /* stratified jitter from t_rand/u    run_nerf.py:654-668  */
  SPN_F_NEED_ALPHA = 16,
};

int spn_version(void);
const char* spn_last_error(void);
/* SM count / arch of the current device (0 on failure). */
int spn_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Measurement hooks (bench.py).  spn_profile_enable(1) brackets every MLP kernel launch with a cudaEvent pair
 * on the launching stream; spn_profile_read(kind, ...) synchronises them and returns the launch count and the
 * summed device time in ms (kind 0 = fused MLP forward, 1 = MLP dgrad, 2 = MLP wgrad).
 * spn_launch_count returns the number of kernels this library has launched (reset != 0 zeroes it). */
int spn_profile_enable(int on);
int spn_profile_read(int kind, int* launches, float* total_ms);
long long spn_launch_count(int reset);

/* ---- a9  rays ------------------------------------------------------------------------- */
/* get_rays (helpers:249-260) for the pixel window [i0,i0+h) x [j0,j0+w) of an H x W image
 * (render()'s `patch`, run_nerf.py:120-123).  c2w: [3,4] row-major. rays_o/rays_d: [h,w,3]. */
int spn_get_rays(const float* c2w, int H, int W, float focal, int i0, int j0, int h, int w,
                 float* rays_o, float* rays_d, void* stream);
/* ndc_rays (helpers:283-300), n rays, in-place allowed. */
int spn_ndc_rays(int n, int H, int W, float focal, float near_plane, const float* rays_o,
                 const float* rays_d, float* out_o, float* out_d, void* stream);
/* render()'s ray-matrix assembly (run_nerf.py:126-153): rays[n,11] =
 * [o(3) d(3) near far viewdir(3)], viewdir = d/|d| taken BEFORE the optional NDC map. */
int spn_build_ray_batch(int n, const float* rays_o, const float* rays_d, float near_, float far_,
                        int ndc, int H, int W, float focal, float* rays, void* stream);
/* Batch assembly from a device-resident ray pool (what RayDataset.__getitem__ + DataLoader
 * collation, DS_NeRF/data.py:4-15 / run_nerf.py:1367-1413, and render()'s ray matrix produce for the sampled rays):
 * row i of rays [n,11] is pool ray idx[i] (int64) laid out like spn_build_ray_batch; the first n_rgb rays also gather
 * their colour target rgb_pool[idx[i]] -> rgb_out [n_rgb,3], the remaining ones their inpainted-disparity target
 * disp_pool[idx[i]] -> disp_out [n - n_rgb] (either pool may be NULL). */
int spn_gather_ray_batch(int n, const float* pool_o, const float* pool_d, const int64_t* idx, float near_, float far_,
                         int ndc, int H, int W, float focal, float* rays, const float* rgb_pool, float* rgb_out,
                         int n_rgb, const float* disp_pool, float* disp_out, void* stream);

/* ---- a5  positional encoding (helpers:22-70) -------------------------------------------- */
/* x [m,3] -> out [m, 3+6*n_freqs] = [x, sin(2^k x), cos(2^k x)...]. */
int spn_embed(const float* x, int64_t m, int n_freqs, float* out, void* stream);

/* ---- a3  stratified depths (run_nerf.py:646-668) ---------------------------------------- */
/* rays [n, ncols] (near = col 6, far = col 7).  t_rand [n,S] or NULL (no jitter). z [n,S]. */
int spn_sample_z(const float* rays, int n, int ncols, int S, int lindisp, const float* t_rand,
                 float* z, void* stream);

/* ---- a7  raw2outputs (helpers:350-401) --------------------------------------------------- */
/* raw [n,S,4], z [n,S], rays_d: pointer to the first direction, row stride `ld_d` floats.
 * noise [n,S] (already scaled by raw_noise_std) or NULL.  alpha may be NULL. */
int spn_raw2outputs_fwd(const float* raw, const float* z, const float* rays_d, int ld_d,
                        const float* noise, int n, int S, int white_bkgd, float* rgb_map,
                        float* disp_map, float* acc_map, float* weights, float* depth_map,
                        float* alpha, void* stream);
/* upstream grads may each be NULL (= zero).  d_raw [n,S,4] is overwritten. */
int spn_raw2outputs_bwd(const float* raw, const float* z, const float* rays_d, int ld_d,
                        const float* noise, int n, int S, int white_bkgd, int detach_weights,
                        const float* g_rgb, const float* g_disp, const float* g_acc,
                        const float* g_weights, const float* g_depth, float* d_raw, void* stream);

/* ---- a8  sample_pdf (helpers:304-347) + sort(cat) (run_nerf.py:702) ----------------------- */
/* bins [n,nb], weights [n,nb-1], u [n,ns] or NULL (det: u = linspace(0,1,ns)).
 * samples [n,ns]; inds [n,ns] int64 or NULL (= searchsorted(cdf,u,right=True)). */
int spn_sample_pdf(const float* bins, const float* weights, const float* u, int n, int nb, int ns,
                   float* samples, int64_t* inds, void* stream);
/* same, additionally writing the cdf [n,nb] it built (helpers:306-309) — lets a test assert
 * "indices bit-exact GIVEN identical cdf and u". cdf_out may be NULL. */
int spn_sample_pdf_cdf(const float* bins, const float* weights, const float* u, int n, int nb, int ns,
                       float* samples, int64_t* inds, float* cdf_out, void* stream);
/* out[n,sa+sb] = sort(cat(a[n,sa], b[n,sb])) (b is unsorted when u is random: warp bitonic sort). */
int spn_merge_sorted(const float* a, const float* b, int n, int sa, int sb, float* out, void* stream);
/* Batched row-wise searchsorted with numpy semantics — the function of the reference's only native component
 * (DS_NeRF/torchsearchsorted/src/cuda/searchsorted_cuda_kernel.cu:83-142; wrapper src/torchsearchsorted/searchsorted.py:20-53),
 * which its hot path does not call (run_nerf_helpers.py:10 imports torch.searchsorted).  a [nrow_a, ncol_a] sorted rows,
 * v [nrow_v, ncol_v] queries, out [max(nrow_a, nrow_v), ncol_v] int64; nrow_a == nrow_v or one of them 1. */
int spn_searchsorted(const float* a, const float* v, int64_t* out, int nrow_a, int nrow_v, int ncol_a, int ncol_v,
                     int side_left, void* stream);
/* render_rays' whole resampling block (run_nerf.py:696-702,726) in one launch:
 * mids -> sample_pdf(mids, weights[:,1:-1]) -> merge with z -> population std of the samples. */
int spn_resample(const float* z, const float* weights, const float* u, int n, int S, int n_imp,
                 float* z_out, float* z_samples, float* z_std, int64_t* inds, void* stream);

/* ---- a6  NeRF MLP (helpers:74-127) fused with a4/a5 (run_nerf.py:56-71) -------------------- */
#define SPN_MLP_NPARAMS 595844   /* floats per network, torch .parameters() order */
#define SPN_MLP_NTENSORS 24
/* Flat parameter layout = concatenation, in nn.Module registration order (helpers:86-102):
 * pts_linears.{0..7}.{weight,bias}, views_linears.0.{weight,bias}, feature_linear.*,
 * alpha_linear.*, rgb_linear.*   (weights row-major [out,in]).  offsets[i] for tensor i. */
int spn_mlp_param_offsets(int64_t* offsets25_host);
/* bytes of the packed (bf16, UMMA-swizzled, fwd + transposed-for-dgrad) weight image */
size_t spn_mlp_packed_bytes(void);
int spn_mlp_pack_weights(const float* params_flat, void* packed, void* stream);
/* activation stash written by a training forward and consumed by spn_mlp_bwd */
size_t spn_mlp_stash_bytes(int64_t m, int precision);
size_t spn_mlp_bwd_workspace_bytes(int64_t m, int precision);
/* Input modes: (a) points: x6 [m,6] = [pt(3), viewdir(3)] per sample (what run_network feeds
 * NeRF.forward through the lazy embedder); (b) rays: rays [n,ncols] + z [n,S], m = n*S,
 * pts = o + d*z computed in-kernel (run_nerf.py:670) and viewdir = last 3 columns.
 * raw [m,4] = [rgb(3), sigma].  stash may be NULL (inference). `packed` is only read for
 * SPN_PREC_BF16, params_flat only for SPN_PREC_FP32 (either may be NULL otherwise). */
int spn_mlp_fwd_points(const float* params_flat, const void* packed, const float* x6, int64_t m,
                       float* raw, void* stash, int precision, void* stream);
int spn_mlp_fwd_rays(const float* params_flat, const void* packed, const float* rays, int ncols,
                     const float* z, int n, int S, float* raw, void* stash, int precision,
                     void* stream);
/* grads_flat [SPN_MLP_NPARAMS] += dL/dparams (accumulates: zero it first if needed). */
int spn_mlp_bwd(const float* params_flat, const void* packed, const void* stash, const float* d_raw,
                int64_t m, float* grads_flat, void* workspace, int precision, void* stream);

/* Diagnostic: a single tcgen05 GEMM D[128,N] = A[128,K] B[N,K]^T (fp32 in/out, bf16 operands on the tensor
 * cores) with exactly the shared-memory/TMEM conventions of the MLP kernels. N in {128,256}, K in {64,128,192,256}. */
int spn_tc_selftest_gemm(const float* A, const float* B, float* D, int N, int K, void* stream);

/* Diagnostic: cycles (clock64, written to cycles_dev[0]) for `reps` back-to-back tcgen05.mma M=128 x N x K=16 with
 * K-major (0) or MN-major (1) shared-memory operands — the measurement behind DESIGN.md's wgrad layout choice. */
/* Diagnostic: the weight-gradient kernel's E4M3 -> bf16 widening of the activation stash applied to n codes (device pointers). */
int spn_tc_e4m3_decode(const uint8_t* codes_dev, uint16_t* bf16_out_dev, int n, void* stream);
/* CTA-pair (cta_group::2, M = 256) MMA rate: ts = 1 takes the A operand from tensor memory; nacc accumulators in turn;
 * ld_warps warps per CTA generate epilogue-like tcgen05.ld / tcgen05.st traffic meanwhile.  out_dev[0] = cycles for reps MMAs. */
int spn_tc_mma_rate_pair(int ts, int n, int reps, int nacc, int ld_warps, long long* out_dev, void* stream);
int spn_tc_mma_rate(int a_mn_major, int b_mn_major, int n, int reps, long long* cycles_dev, void* stream);
/* Diagnostic: TMEM -> register drain rate.  `nwarps` (1..16) warps each read 128 accumulator columns of their lane quarter
 * `reps` times with tcgen05.ld 32x32b.x32; with_mma != 0 keeps M=128 N=256 MMAs running on the same SM meanwhile.
 * out_dev[0] = cycles of the slowest warp, out_dev[1] = MMAs issued meanwhile (tools/tmem_rate.py). */
int spn_tc_tmem_ld_rate(int nwarps, int reps, int with_mma, long long* out_dev, void* stream);
/* Diagnostic: per-CTA cycles (cycles_dev[grid]) to stream `iters` cp.async.bulk copies of `copy_bytes` each from `src`
 * into shared memory with `depth` copies in flight, one thread per CTA. */
int spn_tc_bulk_rate(const void* src, size_t src_bytes, int copy_bytes, int depth, int iters, int grid, int lanes,
                     long long* cycles_dev, void* stream);

/* Diagnostic: when stamps_dev != NULL the next spn_mlp_fwd_* launches (BF16 mode) record clock64 stamps of CTA 0's
 * pipeline events into stamps_dev[3 rounds][12 steps][2 tiles][24 events] (1728 int64; see tools/trace_fwd.py). NULL = off. */
int spn_tc_set_trace(long long* stamps_dev);

/* ---- a11  train-step losses (run_nerf.py:15, 1481-1521, default flags) for a chunk that concatenates the step's three
 * ray groups: [0,n1) unmasked rays, [n1,n1+n2) masked rays of the kept view (both: img2mse of rgb_map and rgb0 against
 * target_rgb [n1+n2,3]), [n1+n2, n1+n2+n3) inpainted-disparity rays (img2mse of disp_map and disp0 against target_disp
 * [n3]; dropped, loss and gradient, when NaN — run_nerf.py:1520).  Writes the upstream gradients g_* ([n,3] / [n]) for
 * spn_render_rays_bwd and out8 = {loss, psnr, the six loss terms}.  sums6_zeroed: 6 floats, zero on entry. */
int spn_train_losses(const float* rgb_map, const float* rgb0, const float* disp_map, const float* disp0,
                     const float* target_rgb, const float* target_disp, int n1, int n2, int n3,
                     float* sums6_zeroed, float* g_rgb, float* g_rgb0, float* g_disp, float* g_disp0,
                     float* out8, void* stream);

/* ---- a12  Adam (run_nerf.py:433-434, 1611-1622), one flat launch --------------------------- */
int spn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                  float lr, float beta1, float beta2, float eps, int step, float grad_scale,
                  void* stream);
/* The same update with the step counter and learning-rate schedule on the device, so that a whole train step can be
 * captured once and replayed as a CUDA graph.  state4 = {step, lr/bias_correction1, sqrt(bias_correction2), lr} (floats,
 * zero-initialised).  spn_adam_tick advances it once per optimisation step: step += 1, lr = lr0 * decay_base^(max(step-2, 0) /
 * decay_steps) — the reference updates the rate AFTER optimizer.step() from its 0-based global_step (run_nerf.py:1611-1622,
 * 1703), so steps 1 and 2 both run at lr0; decay_base 0.1, decay_steps lrate_decay*1000.  spn_adam_step_dev applies it. */
int spn_adam_tick(float* state4, float lr0, float decay_base, float decay_steps, float beta1, float beta2, void* stream);
int spn_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                      const float* state4, float beta1, float beta2, float eps, float grad_scale, void* stream);

/* ---- (e)  multi-GPU gradient exchange over NVLink peer memory, fused with the optimiser ------------------------------
 * No counterpart in the reference (single GPU, run_nerf.py:39); replaces "NCCL all-reduce of the flat gradients, then
 * spn_adam_step per network" of the ray-sharded step.  OPT-IN and not yet run on hardware
 * (csrc/peer_reduce.cu header); the default path is NCCL.
 * Each rank allocates one region of spn_peer_region_bytes(n_floats) with spn_peer_alloc (cudaMalloc + IPC handle, zeroed),
 * exchanges the 64-byte handles out of band, maps the others with spn_peer_open and, after a host-side barrier, calls
 * spn_peer_allreduce_adam once per optimisation step with the same monotonically increasing `epoch` (>= 1) on every rank.
 * The rank's flat gradients live INSIDE its region (spn_peer_grad_ptr): [coarse n_params | pad | fine n_params | pad] with
 * the fine vector at float offset `stride`, n_floats >= 2*stride, all multiples of 4.  `regions` is a HOST array of `world`
 * device pointers (own region at [rank]).  lr / betas / eps / step / grad_scale as in spn_adam_step (both networks). */
size_t spn_peer_region_bytes(int64_t n_floats);
int spn_peer_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64);
int spn_peer_open(const unsigned char* handle64, void** dev_ptr);
int spn_peer_close(void* dev_ptr);
int spn_peer_free(void* dev_ptr);
void* spn_peer_grad_ptr(void* region);
int spn_peer_allreduce_adam(void* const* regions, int world, int rank, unsigned int epoch, int64_t n_floats,
                            int64_t n_params, int64_t stride, float* param_c, float* m_c, float* v_c,
                            float* param_f, float* m_f, float* v_f, float lr, float beta1, float beta2,
                            float eps, int step, float grad_scale, void* stream);

/* ---- a1-a3  render_rays, whole chunk (run_nerf.py:593-737) --------------------------------- */
typedef struct {
  int n_rays;          /* rays in this chunk                                         */
  int ncols;           /* 11 or 12 (run_nerf.py:149-153)                             */
  int n_samples;       /* S  (coarse)                                                */
  int n_importance;    /* 0 or n_imp (fine pass on S+n_imp merged samples)           */
  int flags;           /* SPN_F_*                                                    */
  int precision;       /* SPN_PREC_*                                                 */
  float raw_noise_std; /* >0: noise0/noise1 must be given (unscaled N(0,1) draws)    */
} spn_render_cfg;

typedef struct {       /* device pointers; optional ones may be NULL                 */
  const float* rays;            /* [n, ncols]                                         */
  const float* params_coarse;   /* flat fp32 [SPN_MLP_NPARAMS]                        */
  const float* params_fine;     /* NULL -> fine pass reuses coarse (run_nerf.py:706)  */
  const void* packed_coarse;    /* spn_mlp_pack_weights images (BF16 mode)            */
  const void* packed_fine;
  const float* t_rand;          /* [n,S]    U[0,1) for SPN_F_PERTURB                  */
  const float* u;               /* [n,n_imp] U[0,1) for SPN_F_PERTURB                 */
  const float* noise0;          /* [n,S]    N(0,1)                                    */
  const float* noise1;          /* [n,S+n_imp]                                        */
  /* outputs */
  float* rgb_map; float* disp_map; float* acc_map; float* depth_map;   /* [n,3] [n] [n] [n] */
  float* weights; float* z_vals;                                       /* [n,S'] */
  float* raw;                                                          /* [n,S',4] (needed for bwd) */
  float* alpha; float* alpha0;                                         /* optional */
  float* rgb0; float* disp0; float* acc0; float* z_std;                /* coarse-pass maps */
  /* saved for backward (training): caller-allocated, may be NULL for inference */
  float* z_coarse; float* raw_coarse;                                   /* [n,S] [n,S,4] */
  void* stash_coarse; void* stash_fine;
} spn_render_io;

int spn_render_rays_fwd(const spn_render_cfg* cfg, const spn_render_io* io, void* stream);

typedef struct {       /* upstream gradients (NULL = zero) and gradient outputs       */
  const float* g_rgb; const float* g_disp; const float* g_acc; const float* g_depth;
  const float* g_weights;
  const float* g_rgb0; const float* g_disp0; const float* g_acc0;
  float* grads_coarse;  /* [SPN_MLP_NPARAMS], accumulated                              */
  float* grads_fine;    /* may alias grads_coarse when there is no fine net            */
  float* d_raw_scratch; /* [n, S', 4]                                                  */
  void* workspace;      /* spn_mlp_bwd_workspace_bytes(n*S')                           */
  int detach_begin;     /* rays [detach_begin, detach_end) of the chunk are rendered with detach_weights=True */
  int detach_end;       /* (helpers:385-388) in addition to SPN_F_DETACH_WEIGHTS (= all rays); 0,0 = none: lets the  */
                        /* step's three render calls (run_nerf.py:1455-1470) share one chunk                        */
} spn_render_grads;

int spn_render_rays_bwd(const spn_render_cfg* cfg, const spn_render_io* io,
                        const spn_render_grads* g, void* stream);

/* ---- end-to-end, HOST buffers (pinned or pageable): H2D, render, D2H inside ---------------- */
/* Renders n rays given as a host ray matrix [n,ncols]; writes host rgb[n,3], disp[n], acc[n],
 * depth[n].  params_*_host are flat fp32 parameter vectors.  Synchronises.  (bench.py's `e2e` leg times the train
 * step — host batches -> Trainer.step_graphed -> loss read back — not this render-only call; the render workload's
 * e2e goes through render_path_sharded.) */
int spn_render_host(const spn_render_cfg* cfg, const float* rays_host,
                    const float* params_coarse_host, const float* params_fine_host,
                    float* rgb_host, float* disp_host, float* acc_host, float* depth_host);

#ifdef __cplusplus
}
#endif
#endif /* SPINNERF_B200_H */
