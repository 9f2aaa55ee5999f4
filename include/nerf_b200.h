/*
 * nerf_b200.h -- C ABI of libnerf_b200.so: the B200 (sm_100a) implementation of torch-NeRF's per-ray
 * rendering hot path.  Plain pointers and sizes only; no torch types cross this boundary.
 *
 * The reference (DveloperY0115/torch-NeRF) is pure Python and has no FFI of its own: its "plugin API"
 * is the Python class protocol selected by config strings in torch_nerf/runners/runner_utils.py:526-660.
 * Each entry point below states the reference interface it stands behind (paths relative to the
 * reference root).  The Python mirror of those classes lives in torch-nerf_b200/ and binds these
 * symbols with ctypes (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - every pointer named *_dev / documented "device" is a CUDA device pointer owned by the caller
 *     (the Python side allocates with torch); the library never frees or keeps caller buffers.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no hidden syncs.
 *   - return value: 0 = ok, NERF_ERR_ARG = bad argument (the Python mirror raises ValueError with the
 *     reference's semantics before calling), NERF_ERR_CUDA = CUDA failure (RuntimeError).
 *     nerf_last_error() returns a thread-local message for the last non-zero return.
 *   - float = IEEE binary32.  Row-major, contiguous unless a leading dimension is given.
 */
#ifndef NERF_B200_H
#define NERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NERF_OK 0
#define NERF_ERR_ARG 1
#define NERF_ERR_CUDA 2

typedef void* nerf_stream_t; /* cudaStream_t */

/* ------------------------------------------------------------------------------------------------
 * library / device
 * ---------------------------------------------------------------------------------------------- */
int nerf_version(void);
const char* nerf_last_error(void);
/* fills SM count and compute capability of the current device */
int nerf_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------
 * K1 ray generation
 *   replaces RaySamplerBase.generate_rays  (src/renderer/ray_samplers/sampler_base.py:134-197,
 *   _get_ray_directions :70-113, map_rays_to_ndc :199-257) fed from PerspectiveCamera
 *   (src/renderer/cameras.py:84-153).  Directions are not normalised; NDC keeps the reference's
 *   formula (no origin shift to the near plane).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  float fx, fy, cx, cy;  /* intrinsic[0][0], [1][1], [0][2], [1][2] */
  float rot[9];          /* c2w[:3,:3], row-major */
  float trans[3];        /* c2w[:3,3] */
  float ndc_sx;          /* (float)(-(2*focal/img_w))   (python-double arithmetic, then float) */
  float ndc_sy;          /* (float)(-(2*focal/img_h)) */
  float ndc_two_near;    /* (float)(2*t_near) */
  int32_t img_w, img_h;
  int32_t project_to_ndc;
} nerf_camera_t;

/* coords_dev: (N,2) int64 screen coordinates (u=col, v=H-1-row), as VolumeRenderer.screen_coords holds them */
int nerf_generate_rays(const int64_t* coords_dev, int64_t n, const nerf_camera_t* cam, float* ray_o_dev,
                       float* ray_d_dev, nerf_stream_t stream);
/* pixel_idx_dev: (N,) int64 flat pixel ids p = row*W + col (volume_renderer.py:171-190 folded in);
 * NULL means p = first_pixel + i (whole-frame render, volume_renderer.py:129-133). */
int nerf_generate_rays_from_pixels(const int64_t* pixel_idx_dev, int64_t first_pixel, int64_t n,
                                   const nerf_camera_t* cam, float* ray_o_dev, float* ray_d_dev,
                                   nerf_stream_t stream);

/* The same with the camera in DEVICE memory, for callers that capture the training iteration (runners/train.py:130-218)
 * into a CUDA graph: the reference builds a new PerspectiveCamera per iteration (train.py:136-145), so the camera must
 * not be baked into the captured launch parameters.  nerf_upload_camera writes *cam to cam_dev on the stream (the
 * struct travels in the launch parameters; cam may be reused by the host as soon as the call returns). */
int nerf_upload_camera(nerf_camera_t* cam_dev, const nerf_camera_t* cam, nerf_stream_t stream);
int nerf_generate_rays_from_pixels_devcam(const int64_t* pixel_idx_dev, int64_t first_pixel, int64_t n,
                                          const nerf_camera_t* cam_dev, float* ray_o_dev, float* ray_d_dev,
                                          nerf_stream_t stream);

/* RaySamplerBase.map_rays_to_ndc (sampler_base.py:199-257) on its own: projects world-frame rays (N,3) to NDC;
 * out_* may alias the inputs.  (nerf_generate_rays* apply the same projection when camera.project_to_ndc is set.) */
int nerf_map_rays_to_ndc(const float* ray_o_dev, const float* ray_d_dev, int64_t n, double focal_length, double z_near,
                         int img_height, int img_width, float* out_o_dev, float* out_d_dev, nerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K2 / K3 sampling along rays
 *   replaces StratifiedSampler.sample_along_rays (src/renderer/ray_samplers/stratified_sampler.py:17-128)
 *   and sample_pdf (src/renderer/ray_samplers/utils.py:8-58).
 *   The uniforms are INPUTS (drawn by the caller in the reference's order: coarse pass rand_like(N,Sc);
 *   fine pass rand_like(N,Sc), rand(N,Sf), rand_like(N,Sf)) so results are replayable bit for bit.
 * ---------------------------------------------------------------------------------------------- */
/* host helper: bins[i] of torch.linspace(t_near, t_far, S+1)[:-1] (stratified_sampler.py:156-161) and
 * step = (float)((t_far - t_near) / S) (:162).  bins_host has room for num_partitions floats. */
int nerf_make_bins(double t_near, double t_far, int num_partitions, float* bins_host, float* step_out);

/* The sampling kernels evaluate the bins in-kernel with the same float32 formula as nerf_make_bins, so they
 * take (t_near, t_far) instead of a table.
 *
 * coarse branch (:91-128).  Any of t/pts/dirs/delta outputs may be NULL (not materialised). */
int nerf_sample_coarse(const float* ray_o_dev, const float* ray_d_dev, int64_t n, int num_samples, double t_near,
                       double t_far, const float* u_dev, float* t_dev, float* pts_dev, float* dirs_dev,
                       float* delta_dev, nerf_stream_t stream);

/* sample_pdf (utils.py:8-58).  weights_dev (N,Sc) is modified IN PLACE (+= 1e-5), like the reference.
 * idx_dev (N,Sf) int64 = searchsorted(cdf, u1, right=True) - 1, may be NULL. */
int nerf_sample_pdf(double t_near, double t_far, float* weights_dev, const float* u1_dev, const float* u2_dev,
                    int64_t n, int num_coarse, int num_fine, float* t_fine_dev, int64_t* idx_dev,
                    nerf_stream_t stream);

/* hierarchical branch (:57-90 + :112-128): fresh coarse draw (u0), importance samples (u1,u2), sort,
 * deltas, points.  Outputs have S = num_coarse + num_fine samples per ray; any may be NULL. */
int nerf_sample_fine(const float* ray_o_dev, const float* ray_d_dev, int64_t n, int num_coarse, int num_fine,
                     double t_near, double t_far, float* weights_dev, const float* u0_dev, const float* u1_dev,
                     const float* u2_dev, int64_t* idx_dev, float* t_dev, float* pts_dev, float* dirs_dev,
                     float* delta_dev, nerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K4 positional encoding
 *   replaces PositionalEncoder.encode (src/signal_encoder/positional_encoder.py:49-104):
 *   [x | sin(2^0 x) | cos(2^0 x) | ...], no pi.  out row stride ld_out >= out_dim floats.
 * ---------------------------------------------------------------------------------------------- */
int nerf_posenc(const float* x_dev, int64_t m, int in_dim, int embed_level, int include_input, float* out_dev,
                int64_t ld_out, nerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K7 / K8 alpha compositing
 *   replaces QuadratureIntegrator.integrate_along_rays (src/renderer/integrators/quadrature_integrator.py:14-67)
 *   and its autograd backward.  depth = sum w_i t_i and opacity = sum w_i are additions with no reference
 *   counterpart (t_dev/depth_dev/opacity_dev may be NULL).
 * ---------------------------------------------------------------------------------------------- */
int nerf_composite_fwd(const float* sigma_dev, const float* radiance_dev, const float* delta_dev,
                       const float* t_dev, int64_t n, int s, float* rgb_dev, float* w_dev, float* depth_dev,
                       float* opacity_dev, nerf_stream_t stream);
/* g_w_dev (N,S) may be NULL (zero upstream gradient on the weights, the training case) */
int nerf_composite_bwd(const float* sigma_dev, const float* radiance_dev, const float* delta_dev,
                       const float* g_rgb_dev, const float* g_w_dev, int64_t n, int s, float* g_sigma_dev,
                       float* g_radiance_dev, nerf_stream_t stream);

/* Training-step form of the above (train.py:180/202 + runner_utils.py:731): rgb_dev (N,3) is the rendered colour and
 * target_dev (N,3) the ground truth; the MSE head is folded in -- g_rgb = 2/(3N)(rgb - target) never goes to memory and
 * *loss_accum_dev += mean((rgb - target)^2).  s <= 256. */
int nerf_composite_bwd_mse(const float* sigma_dev, const float* radiance_dev, const float* delta_dev, const float* rgb_dev,
                           const float* target_dev, int64_t n, int s, float* g_sigma_dev, float* g_radiance_dev,
                           float* loss_accum_dev, nerf_stream_t stream);

/* Adam update of a flat fp32 parameter buffer (torch.optim.Adam as set up at runners/runner_utils.py:691-695: lr, eps
 * given; betas, weight_decay = 0, amsgrad = False defaults), update number `step` (1-based):
 *   g = grad * grad_scale;  m += (1 - beta1)(g - m);  v = beta2 v + (1 - beta2) g^2;
 *   p -= lr / (1 - beta1^step) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps)
 * grad_scale folds the 1/world of the data-parallel gradient average into the same pass.  Buffers 16-byte aligned. */
int nerf_adam_step(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t n, double lr,
                   double beta1, double beta2, double eps, int64_t step, double grad_scale, nerf_stream_t stream);

/* MSE loss head of one render pass (runners/runner_utils.py:731 nn.MSELoss as used at runners/train.py:180,202):
 * *loss_accum_dev += mean((rgb - target)^2) over the N*3 elements; g_rgb = 2/(3N) (rgb - target). */
int nerf_mse_loss(const float* rgb_dev, const float* target_dev, int64_t n, float* g_rgb_dev,
                  float* loss_accum_dev, nerf_stream_t stream);

/* Data-parallel step over NVLink peer memory: gradient exchange fused with the optimizer (csrc/dp_exchange.cu; SURVEY 8e --
 * the reference is single-device, runner_utils.py:431-453).  ONE launch per rank and step does what "all-reduce the flat
 * gradient buffer, then nerf_adam_step" does: entry barrier over peer-mapped flag words, rank k sums slice k of every rank's
 * gradient buffer with P2P loads (fixed rank order: bit-identical sums on all replicas) and stores it into slice k of every
 * buffer with P2P stores, barrier, then Adam over ALL parameters from the reduced local buffer (math of nerf_adam_step).
 *   grad_ptrs[r] : rank r's flat gradient buffer as mapped into this process (symmetric / peer memory), n floats padded to
 *                  a multiple of 4, 16-byte aligned;  flag_ptrs[r] : rank r's pad of 2*world 32-bit flag words, zeroed once
 *   seq          : 1, 2, 3, ... identical on all ranks, incremented every call;  counter_dev : one zeroed local word
 * param_dev == NULL: exchange only (the buffers end up holding the sum, no update) -- used by the N-rank parity check.
 * Every rank of the group must make the call; a peer that never arrives traps the kernel after ~70 s instead of hanging. */
int nerf_dp_exchange_adam(float* const* grad_ptrs, uint32_t* const* flag_ptrs, int rank, int world, float* param_dev,
                          float* exp_avg_dev, float* exp_avg_sq_dev, int64_t n, double lr, double beta1, double beta2,
                          double eps, int64_t step, double grad_scale, uint32_t seq, uint32_t* counter_dev,
                          nerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K5 / K6 NeRF MLP
 *   replaces NeRF.forward (src/network/nerf.py:65-121) and its autograd backward.
 *   params: host array of 22 device pointers in state_dict order
 *     fc_in.weight, fc_in.bias, fc_1.weight, fc_1.bias, ... fc_9.weight, fc_9.bias, fc_out.weight, fc_out.bias
 *   weights are (out,in) row-major fp32, exactly nn.Linear's layout (nerf.py:49-59).
 * ---------------------------------------------------------------------------------------------- */
#define NERF_NUM_PARAM_TENSORS 22

typedef struct {
  int32_t pos_dim;  /* 63 */
  int32_t view_dim; /* 27 */
  int32_t feat_dim; /* 256 */
} nerf_mlp_dims_t;

/* -- fp32 validation mode (CUDA-core SGEMM chain; the <=1e-3 parity gate runs on this) -- */
/* floats of activation cache per call of m rows (forward writes it, backward reads it) */
size_t nerf_mlp_f32_cache_floats(const nerf_mlp_dims_t* dims, int64_t m);
/* floats of scratch needed by backward */
size_t nerf_mlp_f32_bwd_scratch_floats(const nerf_mlp_dims_t* dims, int64_t m);
/* pos_dev (M,pos_dim), view_dev (M,view_dim): ENCODED inputs, as NeRF.forward takes them */
int nerf_mlp_f32_forward(const nerf_mlp_dims_t* dims, const float* const* params, const float* pos_dev,
                         const float* view_dev, int64_t m, float* sigma_dev, float* rgb_dev, float* cache_dev,
                         nerf_stream_t stream);
/* grads: host array of 22 device pointers (same order/shapes as params); OVERWRITTEN with dL/dparam */
int nerf_mlp_f32_backward(const nerf_mlp_dims_t* dims, const float* const* params, const float* cache_dev,
                          const float* rgb_dev, int64_t m, const float* g_sigma_dev, const float* g_rgb_dev,
                          float* const* grads, float* scratch_dev, nerf_stream_t stream);

/* -- bf16 tensor-core mode (tcgen05 / TMEM, fp32 accumulation; pos 63 / view 27 / feat 256 only) -- */
/* bytes of the packed weight image (bf16, K-major 128B-swizzled tiles, forward + transposed copies) */
size_t nerf_mlp_bf16_packed_bytes(void);
/* re-pack after every optimizer step */
int nerf_mlp_bf16_pack(const float* const* params, void* packed_dev, nerf_stream_t stream);
/* Prologue of a training iteration in ONE launch: nerf_mlp_bf16_pack of TWO networks (the coarse and the fine one,
 * runner_utils.py:569-660) plus the zero fill of two float buffers (the flat gradient buffer the weight-gradient kernels
 * accumulate into, the loss accumulators); either zero region may be empty. */
int nerf_train_prologue(const float* const* params_a, void* packed_a_dev, const float* const* params_b, void* packed_b_dev,
                        float* zero0_dev, int64_t zero0_count, float* zero1_dev, int64_t zero1_count, nerf_stream_t stream);
/* bytes of per-call training cache for m rows (saved activations in tile-image form + relu masks) */
size_t nerf_mlp_bf16_cache_bytes(int64_t m);
/* Fused query: positional encoding of pts/dirs (K4) is computed inside the kernel as the first layer's
 * operand (cube.py:62-74 + nerf.py:65-121).  pts_dev/dirs_dev (M,3) fp32 RAW (un-encoded).
 * Alternatively pts_dev == NULL and rays are given: row r = ray (r / s), sample (r % s), point =
 * o + t*d, view dir = d  (then t_dev (N,S), ray_o_dev/ray_d_dev (N,3), s = samples per ray).
 * cache_dev == NULL -> inference (nothing saved). */
int nerf_mlp_bf16_forward(const void* packed_dev, const float* pts_dev, const float* dirs_dev,
                          const float* ray_o_dev, const float* ray_d_dev, const float* t_dev, int s, int64_t m,
                          float* sigma_dev, float* rgb_dev, void* cache_dev, nerf_stream_t stream);
size_t nerf_mlp_bf16_bwd_scratch_bytes(int64_t m);
int nerf_mlp_bf16_backward(const void* packed_dev, const void* cache_dev, const float* rgb_dev, int64_t m,
                           const float* g_sigma_dev, const float* g_rgb_dev, float* const* grads,
                           void* scratch_dev, nerf_stream_t stream);

/* The same backward in pieces, for callers that pipeline it: `phases` is a mask (bit 0 zero the gradient tensors, bit 1
 * activation-gradient chain, bit 2 weight gradients -- accumulated onto `grads` with atomics, so several ranges add
 * up), restricted to the 128-row tiles [tile0, tile1) of the m rows (tile0 even) and to `ctas` thread blocks (0 = one
 * per SM).  Lets the engine run the weight gradients of one tile range on some SMs while the chain of the next range
 * runs on the others (engine.py), and lets bench.py time one kernel alone.  nerf_mlp_bf16_backward(...) is
 * phases = 7 over all tiles on all SMs. */
int nerf_mlp_bf16_backward_part(const void* packed_dev, const void* cache_dev, const float* rgb_dev, int64_t m,
                                const float* g_sigma_dev, const float* g_rgb_dev, float* const* grads,
                                void* scratch_dev, int phases, int64_t tile0, int64_t tile1, int ctas,
                                nerf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NERF_B200_H */
