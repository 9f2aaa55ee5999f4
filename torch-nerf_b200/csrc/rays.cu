// K1 ray generation, K2 stratified coarse sampling, K3 hierarchical (inverse-CDF) sampling.
//
// Reference behaviour restated (paths relative to the reference root, torch_nerf/src/...):
//   renderer/ray_samplers/sampler_base.py:70-113,134-197,199-257   rays (+NDC)
//   renderer/ray_samplers/stratified_sampler.py:57-128,130-164     coarse / hierarchical sampling, deltas, points
//   renderer/ray_samplers/utils.py:8-58                             sample_pdf
//
// All three are HBM-bound streaming kernels: one warp owns one ray, lanes own samples, every global access
// is a 128-byte coalesced row segment; (N,S,3) outputs are staged through shared memory so stores stay
// coalesced.  Arithmetic that decides the fine-sample bin index (the bit-exact gate) uses explicit
// round-to-nearest intrinsics in the reference's CPU order: no FMA contraction, IEEE division,
// torch.sum's 4x8-lane accumulation order, torch.cumsum's float64 running sum.
#include <math.h>

#include "common.cuh"

namespace nerf {

// ------------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) raygen_kernel(const int64_t* __restrict__ coords,
                                                      const int64_t* __restrict__ pix, int64_t first_pixel,
                                                      int64_t n, nerf_camera_t cam, float* __restrict__ ray_o,
                                                      float* __restrict__ ray_d) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float u, v;
  if (coords != nullptr) {
    u = (float)coords[2 * i];
    v = (float)coords[2 * i + 1];
  } else {
    int64_t p = pix != nullptr ? pix[i] : first_pixel + i;
    int64_t row = p / cam.img_w;
    int64_t col = p - row * cam.img_w;
    u = (float)col;                        // volume_renderer.py:179-188: (u = col, v = H-1-row)
    v = (float)((int64_t)cam.img_h - 1 - row);
  }
  // sampler_base.py:92-94
  float x = __fdiv_rn(__fsub_rn(u, cam.cx), cam.fx);
  float y = __fdiv_rn(__fsub_rn(v, cam.cy), cam.fy);
  // sampler_base.py:164  d = [x, y, -1] @ R^T
  float d[3], o[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float acc = __fmul_rn(x, cam.rot[3 * j + 0]);
    acc = fmaf(y, cam.rot[3 * j + 1], acc);
    acc = fmaf(-1.0f, cam.rot[3 * j + 2], acc);
    d[j] = acc;
    o[j] = cam.trans[j];  // sampler_base.py:165
  }
  if (cam.project_to_ndc) {
    // sampler_base.py:236-255
    float oxz = __fdiv_rn(o[0], o[2]);
    float oyz = __fdiv_rn(o[1], o[2]);
    float nz = __fdiv_rn(cam.ndc_two_near, o[2]);
    float no0 = __fmul_rn(cam.ndc_sx, oxz);
    float no1 = __fmul_rn(cam.ndc_sy, oyz);
    float no2 = __fadd_rn(1.0f, nz);
    float nd0 = __fmul_rn(cam.ndc_sx, __fsub_rn(__fdiv_rn(d[0], d[2]), oxz));
    float nd1 = __fmul_rn(cam.ndc_sy, __fsub_rn(__fdiv_rn(d[1], d[2]), oyz));
    float nd2 = -nz;
    o[0] = no0, o[1] = no1, o[2] = no2;
    d[0] = nd0, d[1] = nd1, d[2] = nd2;
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    ray_o[3 * i + j] = o[j];
    ray_d[3 * i + j] = d[j];
  }
}

// ------------------------------------------------------------------------------------------------
// shared tail of K2/K3: given sorted t[0..S) of one ray in shared memory, emit t, delta, pts, dirs
// ------------------------------------------------------------------------------------------------
constexpr int kWarpsPerBlock = 4;

// stratified_sampler.py:130-164.  bins = torch.linspace(near, far, P+1)[:-1] evaluated in-kernel with the
// float32 CPU formula (first half start + lin*i, second half end - lin*(steps-1-i)); `step` is the python
// float (far-near)/P applied as a float32 scalar.
struct BinSpec {
  float start, end, lin, step;
  int steps, half;
};

__device__ __forceinline__ float bin_at(const BinSpec& b, int i) {
  return i < b.half ? __fadd_rn(b.start, __fmul_rn(b.lin, (float)i))
                    : __fsub_rn(b.end, __fmul_rn(b.lin, (float)(b.steps - 1 - i)));
}

static BinSpec make_bin_spec(double t_near, double t_far, int num_partitions) {
  BinSpec b;
  b.steps = num_partitions + 1;
  b.half = b.steps / 2;
  b.start = (float)t_near;
  b.end = (float)t_far;
  volatile float lin = (b.end - b.start) / (float)(b.steps - 1);
  b.lin = lin;
  b.step = (float)((t_far - t_near) / (double)num_partitions);
  return b;
}

__device__ __forceinline__ void emit_samples(const float* ts, int s, int64_t ray, const float* __restrict__ ray_o,
                                             const float* __restrict__ ray_d, float* __restrict__ t_out,
                                             float* __restrict__ pts, float* __restrict__ dirs,
                                             float* __restrict__ delta, float* stage) {
  const int lane = lane_id();
  float o[3], d[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    o[j] = __ldg(ray_o + 3 * ray + j);
    d[j] = __ldg(ray_d + 3 * ray + j);
  }
  for (int base = 0; base < s; base += 32) {
    int i = base + lane;
    bool ok = i < s;
    float t = ok ? ts[i] : 0.f;
    if (ok && t_out) t_out[ray * s + i] = t;
    if (ok && delta) {
      // stratified_sampler.py:112-119: diff([t, 1e8])
      float nxt = (i + 1 < s) ? ts[i + 1] : 1e8f;
      delta[ray * s + i] = __fsub_rn(nxt, t);
    }
    int cnt = min(32, s - base);
    if (pts) {
      // stratified_sampler.py:126: o + t*d, product and sum rounded separately
      __syncwarp();
      if (ok) {
#pragma unroll
        for (int j = 0; j < 3; ++j) stage[3 * lane + j] = __fadd_rn(o[j], __fmul_rn(t, d[j]));
      }
      __syncwarp();
      float* dst = pts + ((int64_t)ray * s + base) * 3;
      for (int e = lane; e < 3 * cnt; e += 32) dst[e] = stage[e];
    }
    if (dirs) {
      float* dst = dirs + ((int64_t)ray * s + base) * 3;
      for (int e = lane; e < 3 * cnt; e += 32) dst[e] = d[e % 3];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K2: t = bins + step*u  (stratified_sampler.py:99-109)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    sample_coarse_kernel(const float* __restrict__ ray_o, const float* __restrict__ ray_d, int64_t n, int s,
                         BinSpec bins, const float* __restrict__ u,
                         float* __restrict__ t_out, float* __restrict__ pts, float* __restrict__ dirs,
                         float* __restrict__ delta) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  float* ts = smem + warp * (s + 96);
  float* stage = ts + s;
  int64_t ray = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (ray >= n) return;
  for (int i = lane; i < s; i += 32) ts[i] = __fadd_rn(bin_at(bins, i), __fmul_rn(bins.step, __ldg(u + ray * s + i)));
  __syncwarp();
  emit_samples(ts, s, ray, ray_o, ray_d, t_out, pts, dirs, delta, stage);
}

// ------------------------------------------------------------------------------------------------
// K3
// ------------------------------------------------------------------------------------------------
// torch.sum(dim=-1) on CPU float32 (utils.py:32), restated: lane i accumulates w[i], w[i+32], ... in order;
// lanes l, l+8, l+16, l+24 are combined as ((a0+a1)+a2)+a3; the 8 results are added left to right.
__device__ __forceinline__ float torch_cpu_row_sum(const float* w, int s) {
  const int lane = lane_id();
  float acc = 0.f;
  const int nfull = s / 32;
  for (int c = 0; c < nfull; ++c) acc = __fadd_rn(acc, w[32 * c + lane]);
  float a1 = __shfl_sync(0xffffffffu, acc, (lane + 8) & 31);
  float a2 = __shfl_sync(0xffffffffu, acc, (lane + 16) & 31);
  float a3 = __shfl_sync(0xffffffffu, acc, (lane + 24) & 31);
  float tot = __fadd_rn(__fadd_rn(__fadd_rn(acc, a1), a2), a3);  // valid on lanes 0..7
  // vector-of-8 tail, then the lane reduction, then a scalar tail (not pinned for s % 32 != 0)
  int off = nfull * 32;
  while (s - off >= 8) {
    float v = w[off + (lane & 7)];
    tot = __fadd_rn(tot, v);
    off += 8;
  }
  float r = __shfl_sync(0xffffffffu, tot, 0);
#pragma unroll
  for (int l = 1; l < 8; ++l) r = __fadd_rn(r, __shfl_sync(0xffffffffu, tot, l));
  for (int i = off; i < s; ++i) r = __fadd_rn(r, w[i]);
  return r;
}

// Builds the exclusive CDF of one ray in shared memory; weights updated in place (+= 1e-5).
__device__ __forceinline__ void build_cdf(float* __restrict__ weights_row, int sc, float* ws, float* cdf) {
  const int lane = lane_id();
  for (int i = lane; i < sc; i += 32) {
    float w = __fadd_rn(weights_row[i], 1e-5f);  // utils.py:31 (in place)
    weights_row[i] = w;
    ws[i] = w;
  }
  __syncwarp();
  float z = torch_cpu_row_sum(ws, sc);  // utils.py:32
  __syncwarp();
  for (int i = lane; i < sc; i += 32) ws[i] = __fdiv_rn(ws[i], z);  // utils.py:33
  __syncwarp();
  if (lane == 0) {
    // utils.py:36-40: cumsum with a float64 running sum rounded per element, shifted to exclusive
    double run = 0.0;
    cdf[0] = 0.f;
    for (int j = 0; j + 1 < sc; ++j) {
      run += (double)ws[j];
      cdf[j + 1] = (float)run;
    }
  }
  __syncwarp();
}

// idx = searchsorted(cdf, u, right=True) - 1 = #(cdf_j <= u) - 1   (utils.py:47-54)
__device__ __forceinline__ int upper_bound_minus1(const float* cdf, int sc, float u) {
  int lo = 0, hi = sc;  // first j with cdf[j] > u
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cdf[mid] <= u) lo = mid + 1;
    else hi = mid;
  }
  return lo - 1;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    sample_pdf_kernel(BinSpec bins, float* __restrict__ weights,
                      const float* __restrict__ u1, const float* __restrict__ u2, int64_t n, int sc, int sf,
                      float* __restrict__ t_fine, int64_t* __restrict__ idx_out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  float* ws = smem + warp * 2 * sc;
  float* cdf = ws + sc;
  int64_t ray = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (ray >= n) return;
  build_cdf(weights + ray * sc, sc, ws, cdf);
  for (int f = lane; f < sf; f += 32) {
    int idx = upper_bound_minus1(cdf, sc, __ldg(u1 + ray * sf + f));
    if (idx_out) idx_out[ray * sf + f] = idx;
    t_fine[ray * sf + f] = __fadd_rn(bin_at(bins, idx), __fmul_rn(bins.step, __ldg(u2 + ray * sf + f)));
  }
}

// warp-cooperative bitonic sort of p (power of two) floats in shared memory, ascending
__device__ __forceinline__ void warp_bitonic_sort(float* a, int p) {
  const int lane = lane_id();
  for (int k = 2; k <= p; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int q = lane; q < (p >> 1); q += 32) {
        int i = ((q & ~(j - 1)) << 1) | (q & (j - 1));
        int l = i | j;
        float x = a[i], y = a[l];
        bool up = (i & k) == 0;
        if ((x > y) == up) {
          a[i] = y;
          a[l] = x;
        }
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    sample_fine_kernel(const float* __restrict__ ray_o, const float* __restrict__ ray_d, int64_t n, int sc, int sf,
                       int p2, BinSpec bins, float* __restrict__ weights,
                       const float* __restrict__ u0, const float* __restrict__ u1, const float* __restrict__ u2,
                       int64_t* __restrict__ idx_out, float* __restrict__ t_out, float* __restrict__ pts,
                       float* __restrict__ dirs, float* __restrict__ delta) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  float* ts = smem + warp * (p2 + 2 * sc + 96);
  float* ws = ts + p2;
  float* cdf = ws + sc;
  float* stage = cdf + sc;
  int64_t ray = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (ray >= n) return;
  const int s = sc + sf;
  build_cdf(weights + ray * sc, sc, ws, cdf);
  // stratified_sampler.py:77: a fresh stratified draw, not the coarse pass's samples
  for (int i = lane; i < sc; i += 32) ts[i] = __fadd_rn(bin_at(bins, i), __fmul_rn(bins.step, __ldg(u0 + ray * sc + i)));
  for (int f = lane; f < sf; f += 32) {
    int idx = upper_bound_minus1(cdf, sc, __ldg(u1 + ray * sf + f));
    if (idx_out) idx_out[ray * sf + f] = idx;
    // utils.py:55-56: uniform inside the chosen bin
    ts[sc + f] = __fadd_rn(bin_at(bins, idx), __fmul_rn(bins.step, __ldg(u2 + ray * sf + f)));
  }
  for (int i = s + lane; i < p2; i += 32) ts[i] = INFINITY;
  __syncwarp();
  warp_bitonic_sort(ts, p2);  // stratified_sampler.py:87-90
  emit_samples(ts, s, ray, ray_o, ray_d, t_out, pts, dirs, delta, stage);
}

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace nerf

using namespace nerf;

extern "C" {

int nerf_generate_rays(const int64_t* coords_dev, int64_t n, const nerf_camera_t* cam, float* ray_o_dev,
                       float* ray_d_dev, nerf_stream_t stream) {
  NERF_CHECK_ARG(n >= 0, "nerf_generate_rays: negative ray count");
  if (n == 0) return NERF_OK;
  NERF_CHECK_ARG(cam && coords_dev && ray_o_dev && ray_d_dev, "nerf_generate_rays: null pointer");
  raygen_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(coords_dev, nullptr, 0, n, *cam,
                                                                             ray_o_dev, ray_d_dev);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_generate_rays_from_pixels(const int64_t* pixel_idx_dev, int64_t first_pixel, int64_t n,
                                   const nerf_camera_t* cam, float* ray_o_dev, float* ray_d_dev,
                                   nerf_stream_t stream) {
  NERF_CHECK_ARG(n >= 0, "nerf_generate_rays_from_pixels: negative ray count");
  if (n == 0) return NERF_OK;
  NERF_CHECK_ARG(cam && ray_o_dev && ray_d_dev, "nerf_generate_rays_from_pixels: null pointer");
  NERF_CHECK_ARG(cam->img_w > 0 && cam->img_h > 0, "nerf_generate_rays_from_pixels: bad image size");
  raygen_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(nullptr, pixel_idx_dev, first_pixel, n,
                                                                             *cam, ray_o_dev, ray_d_dev);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_make_bins(double t_near, double t_far, int num_partitions, float* bins_host, float* step_out) {
  NERF_CHECK_ARG(num_partitions > 0 && bins_host && step_out, "nerf_make_bins: bad arguments");
  // torch.linspace(start, end, P+1) float32 on CPU: step = (end-start)/(steps-1) in float32; first half
  // start + step*i, second half end - step*(steps-1-i)  (stratified_sampler.py:156-161)
  const int steps = num_partitions + 1;
  const float start = (float)t_near, end = (float)t_far;
  const float step = (end - start) / (float)(steps - 1);
  const int half = steps / 2;
  for (int i = 0; i < num_partitions; ++i) {
    volatile float prod = (i < half) ? step * (float)i : step * (float)(steps - 1 - i);
    bins_host[i] = (i < half) ? start + prod : end - prod;
  }
  *step_out = (float)((t_far - t_near) / (double)num_partitions);  // stratified_sampler.py:162
  return NERF_OK;
}

int nerf_sample_coarse(const float* ray_o_dev, const float* ray_d_dev, int64_t n, int num_samples, double t_near,
                       double t_far, const float* u_dev, float* t_dev, float* pts_dev, float* dirs_dev,
                       float* delta_dev, nerf_stream_t stream) {
  NERF_CHECK_ARG(num_samples > 0 && num_samples <= 1024, "nerf_sample_coarse: num_samples must be in [1,1024]");
  NERF_CHECK_ARG(n >= 0, "nerf_sample_coarse: negative ray count");
  if (n == 0) return NERF_OK;  // empty ray set: nothing to do (zero-size tensors have null data pointers)
  NERF_CHECK_ARG(ray_o_dev && ray_d_dev && u_dev, "nerf_sample_coarse: null pointer");
  size_t smem = sizeof(float) * kWarpsPerBlock * (num_samples + 96);
  sample_coarse_kernel<<<(unsigned)ceil_div64(n, kWarpsPerBlock), kWarpsPerBlock * 32, smem, as_stream(stream)>>>(
      ray_o_dev, ray_d_dev, n, num_samples, make_bin_spec(t_near, t_far, num_samples), u_dev, t_dev, pts_dev, dirs_dev,
      delta_dev);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_sample_pdf(double t_near, double t_far, float* weights_dev, const float* u1_dev, const float* u2_dev,
                    int64_t n, int num_coarse, int num_fine, float* t_fine_dev, int64_t* idx_dev,
                    nerf_stream_t stream) {
  NERF_CHECK_ARG(num_coarse > 0 && num_coarse <= 1024 && num_fine > 0, "nerf_sample_pdf: bad sample counts");
  NERF_CHECK_ARG(n >= 0, "nerf_sample_pdf: negative ray count");
  if (n == 0) return NERF_OK;
  NERF_CHECK_ARG(weights_dev && u1_dev && u2_dev && t_fine_dev, "nerf_sample_pdf: null pointer");
  size_t smem = sizeof(float) * kWarpsPerBlock * 2 * num_coarse;
  sample_pdf_kernel<<<(unsigned)ceil_div64(n, kWarpsPerBlock), kWarpsPerBlock * 32, smem, as_stream(stream)>>>(
      make_bin_spec(t_near, t_far, num_coarse), weights_dev, u1_dev, u2_dev, n, num_coarse, num_fine, t_fine_dev,
      idx_dev);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_sample_fine(const float* ray_o_dev, const float* ray_d_dev, int64_t n, int num_coarse, int num_fine,
                     double t_near, double t_far, float* weights_dev, const float* u0_dev, const float* u1_dev,
                     const float* u2_dev, int64_t* idx_dev, float* t_dev, float* pts_dev, float* dirs_dev,
                     float* delta_dev, nerf_stream_t stream) {
  NERF_CHECK_ARG(num_coarse > 0 && num_fine > 0 && num_coarse + num_fine <= 2048,
                 "nerf_sample_fine: num_coarse + num_fine must be <= 2048");
  NERF_CHECK_ARG(n >= 0, "nerf_sample_fine: negative ray count");
  if (n == 0) return NERF_OK;
  NERF_CHECK_ARG(ray_o_dev && ray_d_dev && weights_dev && u0_dev && u1_dev && u2_dev,
                 "nerf_sample_fine: null pointer");
  const int p2 = next_pow2(num_coarse + num_fine);
  size_t smem = sizeof(float) * kWarpsPerBlock * (p2 + 2 * num_coarse + 96);
  if (smem > 48 * 1024) {
    NERF_CUDA(cudaFuncSetAttribute(sample_fine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  sample_fine_kernel<<<(unsigned)ceil_div64(n, kWarpsPerBlock), kWarpsPerBlock * 32, smem, as_stream(stream)>>>(
      ray_o_dev, ray_d_dev, n, num_coarse, num_fine, p2, make_bin_spec(t_near, t_far, num_coarse), weights_dev, u0_dev,
      u1_dev, u2_dev, idx_dev, t_dev, pts_dev, dirs_dev, delta_dev);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

}  // extern "C"
