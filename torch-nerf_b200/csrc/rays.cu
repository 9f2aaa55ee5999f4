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
// kRPT rays per thread: 4 rays = 12 floats = three 16-byte stores per output (frame-sized launches); small training
// batches use one ray per thread so that a 4096-ray launch still fills more than a handful of warps

// sampler_base.py:236-255 (map_rays_to_ndc): sx = -(2 f / W), sy = -(2 f / H), two_near = 2 z_near; in place
__device__ __forceinline__ void ndc_project(float sx, float sy, float two_near, float* o, float* d) {
  float oxz = __fdiv_rn(o[0], o[2]);
  float oyz = __fdiv_rn(o[1], o[2]);
  float nz = __fdiv_rn(two_near, o[2]);
  float no0 = __fmul_rn(sx, oxz);
  float no1 = __fmul_rn(sy, oyz);
  float no2 = __fadd_rn(1.0f, nz);
  float nd0 = __fmul_rn(sx, __fsub_rn(__fdiv_rn(d[0], d[2]), oxz));
  float nd1 = __fmul_rn(sy, __fsub_rn(__fdiv_rn(d[1], d[2]), oyz));
  float nd2 = -nz;
  o[0] = no0, o[1] = no1, o[2] = no2;
  d[0] = nd0, d[1] = nd1, d[2] = nd2;
}

__global__ void __launch_bounds__(256) ndc_kernel(const float* __restrict__ ray_o, const float* __restrict__ ray_d, int64_t n,
                                                   float sx, float sy, float two_near, float* __restrict__ out_o,
                                                   float* __restrict__ out_d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float o[3] = {ray_o[3 * i], ray_o[3 * i + 1], ray_o[3 * i + 2]};
  float d[3] = {ray_d[3 * i], ray_d[3 * i + 1], ray_d[3 * i + 2]};
  ndc_project(sx, sy, two_near, o, d);
#pragma unroll
  for (int j = 0; j < 3; ++j) out_o[3 * i + j] = o[j], out_d[3 * i + j] = d[j];
}

__device__ __forceinline__ void one_ray(const int64_t* __restrict__ coords, const int64_t* __restrict__ pix,
                                        int64_t first_pixel, int64_t i, const nerf_camera_t& cam, float* o, float* d) {
  float u, v;
  if (coords != nullptr) {
    u = (float)coords[2 * i];
    v = (float)coords[2 * i + 1];
  } else {
    const int64_t p = pix != nullptr ? pix[i] : first_pixel + i;
    int64_t row, col;
    if ((uint64_t)p <= 0xffffffffu) {  // 32-bit division: a fraction of the instructions of the 64-bit one
      row = (uint32_t)p / (uint32_t)cam.img_w;
      col = (uint32_t)p - (uint32_t)row * (uint32_t)cam.img_w;
    } else {
      row = p / cam.img_w;
      col = p - row * cam.img_w;
    }
    u = (float)col;                        // volume_renderer.py:179-188: (u = col, v = H-1-row)
    v = (float)((int64_t)cam.img_h - 1 - row);
  }
  // sampler_base.py:92-94
  float x = __fdiv_rn(__fsub_rn(u, cam.cx), cam.fx);
  float y = __fdiv_rn(__fsub_rn(v, cam.cy), cam.fy);
  // sampler_base.py:164  d = [x, y, -1] @ R^T
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float acc = __fmul_rn(x, cam.rot[3 * j + 0]);
    acc = fmaf(y, cam.rot[3 * j + 1], acc);
    acc = fmaf(-1.0f, cam.rot[3 * j + 2], acc);
    d[j] = acc;
    o[j] = cam.trans[j];  // sampler_base.py:165
  }
  if (cam.project_to_ndc) ndc_project(cam.ndc_sx, cam.ndc_sy, cam.ndc_two_near, o, d);
}

// `cam_dev` != null: the camera is read from device memory (a captured CUDA graph replays with whatever camera
// nerf_upload_camera put there); otherwise it travels by value in the launch parameters.
template <int kRPT>
__global__ void __launch_bounds__(256) raygen_kernel(const int64_t* __restrict__ coords,
                                                      const int64_t* __restrict__ pix, int64_t first_pixel,
                                                      int64_t n, nerf_camera_t cam_arg,
                                                      const nerf_camera_t* __restrict__ cam_dev,
                                                      float* __restrict__ ray_o, float* __restrict__ ray_d, int vec_ok) {
  nerf_camera_t cam = cam_arg;
  if (cam_dev != nullptr) cam = *cam_dev;
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kRPT;
  if (i0 >= n) return;
  if (kRPT == 4 && vec_ok && i0 + kRPT <= n) {
    float o[3 * kRPT], d[3 * kRPT];
#pragma unroll
    for (int r = 0; r < kRPT; ++r) one_ray(coords, pix, first_pixel, i0 + r, cam, o + 3 * r, d + 3 * r);
    float4* dst_o = reinterpret_cast<float4*>(ray_o + 3 * i0);  // 48-byte slabs of a 16-byte aligned array
    float4* dst_d = reinterpret_cast<float4*>(ray_d + 3 * i0);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      dst_o[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
      dst_d[k] = make_float4(d[4 * k], d[4 * k + 1], d[4 * k + 2], d[4 * k + 3]);
    }
  } else {
    for (int64_t i = i0; i < min(i0 + kRPT, n); ++i) {
      float o[3], d[3];
      one_ray(coords, pix, first_pixel, i, cam, o, d);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        ray_o[3 * i + j] = o[j];
        ray_d[3 * i + j] = d[j];
      }
    }
  }
}

__global__ void store_camera_kernel(nerf_camera_t* dst, nerf_camera_t cam) {
  if (threadIdx.x == 0) *dst = cam;
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

// K2 without the (N,S,3) outputs (the fused engine's form): nothing couples the samples of a ray except the forward
// difference, so this is an elementwise stream -- one thread per four consecutive samples, 16-byte loads and stores;
// the first sample of the next quad is recomputed from one extra scalar load (same sector, no extra HBM traffic).
__global__ void __launch_bounds__(256)
    sample_coarse_flat_kernel(int64_t quads, int s, BinSpec bins, const float* __restrict__ u,
                              float* __restrict__ t_out, float* __restrict__ delta) {
  const int64_t q4 = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (q4 >= quads) return;
  const int s4 = s >> 2;
  const int i = 4 * (quads <= 0x7fffffff ? (int)((unsigned)q4 % (unsigned)s4) : (int)(q4 % s4));
  const int64_t q = 4 * q4;
  const float4 uv = __ldg(reinterpret_cast<const float4*>(u + q));
  float t[5];
  t[0] = __fadd_rn(bin_at(bins, i), __fmul_rn(bins.step, uv.x));
  t[1] = __fadd_rn(bin_at(bins, i + 1), __fmul_rn(bins.step, uv.y));
  t[2] = __fadd_rn(bin_at(bins, i + 2), __fmul_rn(bins.step, uv.z));
  t[3] = __fadd_rn(bin_at(bins, i + 3), __fmul_rn(bins.step, uv.w));
  t[4] = (i + 4 < s) ? __fadd_rn(bin_at(bins, i + 4), __fmul_rn(bins.step, __ldg(u + q + 4))) : 1e8f;  // :112-119
  if (t_out) *reinterpret_cast<float4*>(t_out + q) = make_float4(t[0], t[1], t[2], t[3]);
  if (delta)
    *reinterpret_cast<float4*>(delta + q) =
        make_float4(__fsub_rn(t[1], t[0]), __fsub_rn(t[2], t[1]), __fsub_rn(t[3], t[2]), __fsub_rn(t[4], t[3]));
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
  bool exact = true;
  for (int i = lane; i < sc; i += 32) {
    const float pdf = __fdiv_rn(ws[i], z);  // utils.py:33
    ws[i] = pdf;
    exact = exact && (pdf >= 0x1p-28f) && (pdf <= 1.0f);
  }
  __syncwarp();
  // utils.py:36-40: cumsum with a float64 running sum rounded per element, shifted to exclusive.
  // When every pdf value is in [2^-28, 1] (always the case for compositing weights: w + 1e-5 over a sum <= ~1) each is
  // a multiple of 2^-51; all pdf > 0 means all weights share the sign of their float32 sum z, whose relative error
  // is at most ~sc * 2^-24, so every partial sum stays below 2 and fits 52 significant bits: ALL float64 partial sums
  // are exact whatever the order, and a warp scan returns exactly the sequential running sum.  Anything else (huge
  // dynamic range, mixed signs, NaN) takes the sequential loop.
  if (__all_sync(0xffffffffu, exact) && sc <= 1024) {
    double carry = 0.0;
    if (lane == 0) cdf[0] = 0.f;
    for (int base = 0; base < sc; base += 32) {
      const int j = base + lane;
      double inc = j < sc ? (double)ws[j] : 0.0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double up = __hiloint2double(__shfl_up_sync(0xffffffffu, __double2hiint(inc), d),
                                           __shfl_up_sync(0xffffffffu, __double2loint(inc), d));
        if (lane >= d) inc += up;
      }
      inc += carry;
      if (j + 1 < sc) cdf[j + 1] = (float)inc;
      carry = __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(inc), 31),
                               __shfl_sync(0xffffffffu, __double2loint(inc), 31));
    }
  } else if (lane == 0) {
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

// ------------------------------------------------------------------------------------------------
// K3 fast path for the reference's default counts (64 coarse + 128 fine): the 256-wide bitonic network in shared
// memory is what the general kernel spends its time on (36 stages, ~1700 warp instructions per ray).  Here the two sets
// are sorted separately IN REGISTERS (the stratified draws only if they are out of order, the importance draws four per
// lane in blocked order) and then merged by ONE register bitonic merge of [fine ascending | +inf | coarse descending],
// eight values per lane; in the fused form t and delta are stored straight from those registers.  Same multiset, so
// the output is bit-identical to the full sort.
// ------------------------------------------------------------------------------------------------
template <int NSLOT>
__device__ __forceinline__ void warp_bitonic_sort_regs(float (&v)[NSLOT]) {
  const int lane = lane_id();
#pragma unroll
  for (int k = 2; k <= 32 * NSLOT; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int jr = j >> 5;
#pragma unroll
        for (int r = 0; r < NSLOT; ++r) {
          if ((r & jr) == 0) {
            const bool up = (((r << 5) & k) == 0);  // k >= 64 here: the direction depends on the slot only
            const float lo = fminf(v[r], v[r | jr]), hi = fmaxf(v[r], v[r | jr]);
            v[r] = up ? lo : hi;
            v[r | jr] = up ? hi : lo;
          }
        }
      } else {
        const bool lower = (lane & j) == 0;
#pragma unroll
        for (int r = 0; r < NSLOT; ++r) {
          const float other = __shfl_xor_sync(0xffffffffu, v[r], j);
          const bool up = ((((r << 5) | lane) & k) == 0);
          // keep the minimum iff up == lower; one compare (xor-ed with the direction) + one select per exchange
          const bool take = (v[r] > other) != (up != lower);
          v[r] = take ? other : v[r];
        }
      }
    }
  }
}

// Ascending sort of 128 floats held four per lane in BLOCKED order (element e = 4 * lane + slot).  It is the bitonic
// network in its all-ascending form: the first step of every merge pairs e with its mirror image e ^ (k - 1) inside the
// k-block, the remaining steps pair e with e ^ j, and the smaller value always goes to the smaller index.  Steps with
// j < 4 stay inside the lane (13 of the 28 stages: two min/max pairs each, no shuffle, no select); the other 15 cost one
// shuffle + compare + select per element.
__device__ __forceinline__ void warp_sort128_blocked(float (&v)[4]) {
  const int lane = lane_id();
  auto cx = [&](int a, int b) {  // in-lane compare-exchange, a < b
    const float lo = fminf(v[a], v[b]), hi = fmaxf(v[a], v[b]);
    v[a] = lo, v[b] = hi;
  };
#pragma unroll
  for (int k = 2; k <= 128; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const bool mirror = (j == (k >> 1));
      if (j >= 4) {
        const int lane_mask = mirror ? ((k - 1) >> 2) : (j >> 2);
        const bool keep_min = (lane & (j >> 2)) == 0;  // this lane holds the smaller index of each pair
        float other[4];
#pragma unroll
        for (int r = 0; r < 4; ++r)  // a mirrored partner sits in the opposite slot: offer slot r ^ 3, receive the partner's
          other[r] = __shfl_xor_sync(0xffffffffu, mirror ? v[r ^ 3] : v[r], lane_mask);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const bool take = (v[r] > other[r]) == keep_min;  // smaller index keeps the minimum (ties: same value either way)
          v[r] = take ? other[r] : v[r];
        }
      } else if (j == 2) {
        if (mirror) {  // k == 4: pairs (0,3) (1,2)
          cx(0, 3);
          cx(1, 2);
        } else {
          cx(0, 2);
          cx(1, 3);
        }
      } else {  // j == 1 (k == 2: the mirror of e is e ^ 1 as well)
        cx(0, 1);
        cx(2, 3);
      }
    }
  }
}

constexpr int kFastSc = 64, kFastSf = 128;

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    sample_fine_64_128_kernel(const float* __restrict__ ray_o, const float* __restrict__ ray_d, int64_t n, BinSpec bins,
                              float* __restrict__ weights, const float* __restrict__ u0, const float* __restrict__ u1,
                              const float* __restrict__ u2, int64_t* __restrict__ idx_out, float* __restrict__ t_out,
                              float* __restrict__ pts, float* __restrict__ dirs, float* __restrict__ delta, int vec_ok) {
  constexpr int sc = kFastSc, sf = kFastSf, s = sc + sf;
  // per warp: merged t (192) | sorted coarse (64) | sorted fine (128) | normalised weights (64) | cdf (64) | staging (96)
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  float* ts = smem + warp * (s + sc + sf + 2 * sc + 96);
  float* sa = ts + s;
  float* sb = sa + sc;
  float* ws = sb + sf;
  float* cdf = ws + sc;
  float* stage = cdf + sc;
  const int64_t ray = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (ray >= n) return;
  // every global load of the ray is issued before the CDF is built
  float uc[2], ua[4], ub[4];
#pragma unroll
  for (int r = 0; r < 2; ++r) uc[r] = __ldg(u0 + ray * sc + 32 * r + lane);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    ua[r] = __ldg(u1 + ray * sf + 32 * r + lane);
    ub[r] = __ldg(u2 + ray * sf + 32 * r + lane);
  }
  build_cdf(weights + ray * sc, sc, ws, cdf);
  float tc[2], tf[4];
#pragma unroll
  for (int r = 0; r < 2; ++r)  // stratified_sampler.py:77: a fresh stratified draw, not the coarse pass's samples
    tc[r] = __fadd_rn(bin_at(bins, 32 * r + lane), __fmul_rn(bins.step, uc[r]));
  {
    // idx = #(cdf_j <= u) - 1 (utils.py:47-54); cdf[0] = 0 <= u for every u in [0,1), so count over cdf[1..63] instead
    int idx[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) idx[r] = 0;
#pragma unroll
    for (int step = sc / 2; step > 0; step >>= 1) {
#pragma unroll
      for (int r = 0; r < 4; ++r) idx[r] += (cdf[idx[r] + step] <= ua[r]) ? step : 0;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (!(0.f <= ua[r])) idx[r] = -1;  // out-of-contract draws (negative, NaN) behave like the general kernel
      if (idx_out) idx_out[ray * sf + 32 * r + lane] = idx[r];
      tf[r] = __fadd_rn(bin_at(bins, idx[r]), __fmul_rn(bins.step, ub[r]));  // utils.py:55-56
    }
  }
  // the stratified draw is ascending by construction whenever the bins are exact in float32 (the shipped scene bounds);
  // sort it only if a neighbour pair is out of order
  {
    const float nxt0 = __shfl_down_sync(0xffffffffu, tc[0], 1), first1 = __shfl_sync(0xffffffffu, tc[1], 0);
    const float nxt1 = __shfl_down_sync(0xffffffffu, tc[1], 1);
    const bool bad = (tc[0] > (lane == 31 ? first1 : nxt0)) || (lane < 31 && tc[1] > nxt1);
    if (__any_sync(0xffffffffu, bad)) warp_bitonic_sort_regs<2>(tc);
  }
  // which draw sits in which (lane, slot) is irrelevant before the sort: read the striped registers as blocked
  warp_sort128_blocked(tf);
#pragma unroll
  for (int r = 0; r < 2; ++r) sa[32 * r + lane] = tc[r];
  *reinterpret_cast<float4*>(sb + 4 * lane) = make_float4(tf[0], tf[1], tf[2], tf[3]);
  __syncwarp();
  // merge (stratified_sampler.py:87-90 sorts the concatenation) as ONE bitonic merge of 256 values held eight per
  // lane in blocked order: [fine ascending (128) | +inf (64) | coarse descending (64)] is a bitonic sequence, so
  // log2(256) = 8 compare-exchange stages sort it: five across lanes (shuffle), three inside the lane.
  float v[8];
  if (lane < 16) {
    const float4 a = *reinterpret_cast<const float4*>(sb + 8 * lane), b = *reinterpret_cast<const float4*>(sb + 8 * lane + 4);
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
  } else if (lane < 24) {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = INFINITY;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = sa[255 - 8 * lane - k];
  }
#pragma unroll
  for (int j = 128; j >= 8; j >>= 1) {
    const bool keep_min = (lane & (j >> 3)) == 0;
    float other[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) other[k] = __shfl_xor_sync(0xffffffffu, v[k], j >> 3);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const bool take = (v[k] > other[k]) == keep_min;
      v[k] = take ? other[k] : v[k];
    }
  }
#pragma unroll
  for (int j = 4; j >= 1; j >>= 1) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if ((k & j) == 0) {
        const float lo = fminf(v[k], v[k | j]), hi = fmaxf(v[k], v[k | j]);
        v[k] = lo, v[k | j] = hi;
      }
    }
  }
  // positions 8 lane + k < 192 (lanes 0..23) hold the sorted samples: the +inf padding ends up behind them
  if (!pts && !dirs && vec_ok) {
    // fused form: t and the forward differences (stratified_sampler.py:112-119) go out straight from the registers
    const float nxt = __shfl_down_sync(0xffffffffu, v[0], 1);
    if (lane < 24) {
      float d[8];
#pragma unroll
      for (int k = 0; k < 7; ++k) d[k] = __fsub_rn(v[k + 1], v[k]);
      d[7] = __fsub_rn(lane == 23 ? 1e8f : nxt, v[7]);
      if (t_out) {
        float4* dst = reinterpret_cast<float4*>(t_out + ray * s + 8 * lane);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
      if (delta) {
        float4* dst = reinterpret_cast<float4*>(delta + ray * s + 8 * lane);
        dst[0] = make_float4(d[0], d[1], d[2], d[3]);
        dst[1] = make_float4(d[4], d[5], d[6], d[7]);
      }
    }
    return;
  }
  if (lane < 24) {
    *reinterpret_cast<float4*>(ts + 8 * lane) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(ts + 8 * lane + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  __syncwarp();
  emit_samples(ts, s, ray, ray_o, ray_d, t_out, pts, dirs, delta, stage);
}

static int aligned16(const void* a, const void* b) {
  return (((uintptr_t)a | (uintptr_t)b) & 15) == 0;
}

static void launch_raygen(const int64_t* coords, const int64_t* pix, int64_t first_pixel, int64_t n, const nerf_camera_t& cam,
                          const nerf_camera_t* cam_dev, float* ray_o, float* ray_d, cudaStream_t stream) {
  if (n >= (int64_t)1 << 18)
    raygen_kernel<4><<<(unsigned)ceil_div64(n, 256 * 4), 256, 0, stream>>>(coords, pix, first_pixel, n, cam, cam_dev, ray_o, ray_d,
                                                                         aligned16(ray_o, ray_d));
  else
    raygen_kernel<1><<<(unsigned)ceil_div64(n, 256), 256, 0, stream>>>(coords, pix, first_pixel, n, cam, cam_dev, ray_o, ray_d, 0);
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
  launch_raygen(coords_dev, nullptr, 0, n, *cam, nullptr, ray_o_dev, ray_d_dev, as_stream(stream));
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
  launch_raygen(nullptr, pixel_idx_dev, first_pixel, n, *cam, nullptr, ray_o_dev, ray_d_dev, as_stream(stream));
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_map_rays_to_ndc(const float* ray_o_dev, const float* ray_d_dev, int64_t n, double focal_length, double z_near,
                         int img_height, int img_width, float* out_o_dev, float* out_d_dev, nerf_stream_t stream) {
  NERF_CHECK_ARG(n >= 0, "nerf_map_rays_to_ndc: negative ray count");
  NERF_CHECK_ARG(z_near >= 0, "nerf_map_rays_to_ndc: z_near must be >= 0");
  NERF_CHECK_ARG(img_height > 0 && img_width > 0, "nerf_map_rays_to_ndc: bad image size");
  if (n == 0) return NERF_OK;
  NERF_CHECK_ARG(ray_o_dev && ray_d_dev && out_o_dev && out_d_dev, "nerf_map_rays_to_ndc: null pointer");
  // the scale factors are Python floats in the reference (evaluated in double, applied as float32 scalars)
  const float sx = (float)(-(2 * focal_length / img_width)), sy = (float)(-(2 * focal_length / img_height));
  ndc_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(ray_o_dev, ray_d_dev, n, sx, sy, (float)(2 * z_near),
                                                                        out_o_dev, out_d_dev);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_upload_camera(nerf_camera_t* cam_dev, const nerf_camera_t* cam, nerf_stream_t stream) {
  NERF_CHECK_ARG(cam_dev && cam, "nerf_upload_camera: null pointer");
  // the struct rides in the launch parameters (copied at launch time): no pinned staging buffer whose lifetime the
  // caller would have to manage, and the write is ordered on the stream like any kernel
  store_camera_kernel<<<1, 32, 0, as_stream(stream)>>>(cam_dev, *cam);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_generate_rays_from_pixels_devcam(const int64_t* pixel_idx_dev, int64_t first_pixel, int64_t n,
                                          const nerf_camera_t* cam_dev, float* ray_o_dev, float* ray_d_dev,
                                          nerf_stream_t stream) {
  NERF_CHECK_ARG(n >= 0, "nerf_generate_rays_from_pixels_devcam: negative ray count");
  if (n == 0) return NERF_OK;
  NERF_CHECK_ARG(cam_dev && ray_o_dev && ray_d_dev, "nerf_generate_rays_from_pixels_devcam: null pointer");
  nerf_camera_t unused = {};
  launch_raygen(nullptr, pixel_idx_dev, first_pixel, n, unused, cam_dev, ray_o_dev, ray_d_dev, as_stream(stream));
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
  if (!pts_dev && !dirs_dev && num_samples % 4 == 0 && ((uintptr_t)u_dev & 15) == 0 &&
      aligned16(t_dev, delta_dev)) {
    const int64_t quads = n * (num_samples / 4);
    sample_coarse_flat_kernel<<<(unsigned)ceil_div64(quads, 256), 256, 0, as_stream(stream)>>>(
        quads, num_samples, make_bin_spec(t_near, t_far, num_samples), u_dev, t_dev, delta_dev);
    NERF_LAUNCH_CHECK();
    return NERF_OK;
  }
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
  if (num_coarse == kFastSc && num_fine == kFastSf) {
    const size_t smem_fast = sizeof(float) * kWarpsPerBlock * ((kFastSc + kFastSf) * 2 + 2 * kFastSc + 96);
    sample_fine_64_128_kernel<<<(unsigned)ceil_div64(n, kWarpsPerBlock), kWarpsPerBlock * 32, smem_fast, as_stream(stream)>>>(
        ray_o_dev, ray_d_dev, n, make_bin_spec(t_near, t_far, num_coarse), weights_dev, u0_dev, u1_dev, u2_dev, idx_dev, t_dev,
        pts_dev, dirs_dev, delta_dev, aligned16(t_dev, delta_dev));
    NERF_LAUNCH_CHECK();
    return NERF_OK;
  }
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
