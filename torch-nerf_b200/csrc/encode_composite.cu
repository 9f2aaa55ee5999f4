// K4 positional encoding (standalone, for the drop-in encoder and parity checks) and K7/K8 alpha compositing.
//
// Reference behaviour restated (paths relative to the reference root, torch_nerf/src/...):
//   signal_encoder/positional_encoder.py:49-104          [x | sin(2^l x) | cos(2^l x)]_l, no pi
//   renderer/integrators/quadrature_integrator.py:14-67  x = sigma*delta, T = exp(-exclusive cumsum), w = T(1-e^-x)
//
// Both are HBM-bound: 384 B/sample (encoding), 24 B/sample forward and 40 B/sample backward (compositing).
// Compositing is one warp per ray; the transmittance prefix is a float64 shuffle scan whose per-element
// partial sums are rounded to float32 exactly like torch's CPU cumsum, and it is a TRUE exclusive scan
// (shifted), never inclusive-minus-own: delta_last = 1e8 would cancel catastrophically.
#include <math.h>

#include "common.cuh"

namespace nerf {

// ------------------------------------------------------------------------------------------------
// K4
// ------------------------------------------------------------------------------------------------
constexpr int kEncRows = 64;  // rows per CTA

__global__ void __launch_bounds__(256)
    posenc_kernel(const float* __restrict__ x, int64_t m, int c, int levels, int include_input,
                  float* __restrict__ out, int64_t ld_out) {
  extern __shared__ float tile[];  // [kEncRows][out_dim]
  const int out_dim = c * (2 * levels + (include_input ? 1 : 0));
  const int64_t row0 = (int64_t)blockIdx.x * kEncRows;
  const int rows = (int)min((int64_t)kEncRows, m - row0);
  const int shift = include_input ? c : 0;
  // one thread per (row, channel): coalesced read of the (rows, c) slab
  for (int e = threadIdx.x; e < rows * c; e += blockDim.x) {
    int r = e / c, ch = e - r * c;
    float v = __ldg(x + (row0 * c) + e);
    float* dst = tile + r * out_dim;
    if (include_input) dst[ch] = v;
    float f = 1.0f;
    for (int l = 0; l < levels; ++l) {
      float s, co;
      sincosf(__fmul_rn(f, v), &s, &co);  // freq * x is exact (power of two); full-range reduction
      dst[shift + (2 * l) * c + ch] = s;
      dst[shift + (2 * l + 1) * c + ch] = co;
      f *= 2.0f;
    }
  }
  __syncthreads();
  if (ld_out == out_dim) {
    float* dst = out + row0 * ld_out;
    for (int e = threadIdx.x; e < rows * out_dim; e += blockDim.x) dst[e] = tile[e];
  } else {
    for (int e = threadIdx.x; e < rows * out_dim; e += blockDim.x) {
      int r = e / out_dim, k = e - r * out_dim;
      out[(row0 + r) * ld_out + k] = tile[e];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K7 / K8
// ------------------------------------------------------------------------------------------------
constexpr int kCompWarps = 4;

__device__ __forceinline__ double shfl_up_f64(double v, int d) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_up_sync(0xffffffffu, lo, d);
  hi = __shfl_up_sync(0xffffffffu, hi, d);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_down_f64(double v, int d) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_down_sync(0xffffffffu, lo, d);
  hi = __shfl_down_sync(0xffffffffu, hi, d);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_f64(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(0xffffffffu, lo, src);
  hi = __shfl_sync(0xffffffffu, hi, src);
  return __hiloint2double(hi, lo);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// loads the 32x3 radiance block of one chunk through shared memory (coalesced), returns this lane's rgb
__device__ __forceinline__ void load_rgb_chunk(const float* __restrict__ src, int cnt, float* stage, float& r,
                                               float& g, float& b) {
  const int lane = lane_id();
  __syncwarp();
  for (int e = lane; e < 3 * cnt; e += 32) stage[e] = __ldg(src + e);
  __syncwarp();
  if (lane < cnt) {
    r = stage[3 * lane], g = stage[3 * lane + 1], b = stage[3 * lane + 2];
  } else {
    r = g = b = 0.f;
  }
}

__global__ void __launch_bounds__(kCompWarps * 32)
    composite_fwd_kernel(const float* __restrict__ sigma, const float* __restrict__ radiance,
                         const float* __restrict__ delta, const float* __restrict__ tvals, int64_t n, int s,
                         float* __restrict__ rgb_out, float* __restrict__ w_out, float* __restrict__ depth_out,
                         float* __restrict__ opacity_out) {
  __shared__ float stage_all[kCompWarps][96];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  float* stage = stage_all[warp];
  int64_t ray = (int64_t)blockIdx.x * kCompWarps + warp;
  if (ray >= n) return;
  double carry = 0.0;   // float64 running sum of x over previous chunks
  float prev_csum = 0.f;  // float32-rounded inclusive sum up to the previous chunk's last element
  float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_d = 0.f, acc_o = 0.f;
  for (int base = 0; base < s; base += 32) {
    const int i = base + lane;
    const bool ok = i < s;
    const int cnt = min(32, s - base);
    float x = 0.f;
    if (ok) x = __fmul_rn(__ldg(sigma + ray * s + i), __ldg(delta + ray * s + i));  // :41
    // inclusive float64 scan of x within the chunk
    double inc = (double)x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      double up = shfl_up_f64(inc, d);
      if (lane >= d) inc += up;
    }
    inc += carry;
    float csum = (float)inc;  // torch cumsum output element (float32)
    // exclusive = previous element's (rounded) inclusive sum  (:44-52)
    float excl = __shfl_up_sync(0xffffffffu, csum, 1);
    if (lane == 0) excl = prev_csum;
    float trans = expf(-excl);
    float alpha = __fsub_rn(1.0f, expf(-x));  // :55
    float w = ok ? __fmul_rn(trans, alpha) : 0.f;  // :58
    if (ok && w_out) w_out[ray * s + i] = w;
    float r, g, b;
    load_rgb_chunk(radiance + (ray * s + base) * 3, cnt, stage, r, g, b);
    acc_r = fmaf(w, r, acc_r);
    acc_g = fmaf(w, g, acc_g);
    acc_b = fmaf(w, b, acc_b);
    if (depth_out && ok) acc_d = fmaf(w, __ldg(tvals + ray * s + i), acc_d);
    acc_o += w;
    carry = shfl_f64(inc, 31);
    prev_csum = __shfl_sync(0xffffffffu, csum, 31);
  }
  acc_r = warp_sum(acc_r), acc_g = warp_sum(acc_g), acc_b = warp_sum(acc_b);
  if (depth_out) acc_d = warp_sum(acc_d);
  if (opacity_out) acc_o = warp_sum(acc_o);
  if (lane == 0) {
    rgb_out[3 * ray + 0] = acc_r;
    rgb_out[3 * ray + 1] = acc_g;
    rgb_out[3 * ray + 2] = acc_b;
    if (depth_out) depth_out[ray] = acc_d;
    if (opacity_out) opacity_out[ray] = acc_o;
  }
}

// Backward of w_i = T_i (1 - e^{-x_i}), rgb = sum w_i c_i, x = sigma*delta:
//   g_c_i = w_i g_rgb;  g_w_i = g_rgb . c_i (+ external);  g_x_i = g_w_i T_{i+1} - sum_{k>i} g_w_k w_k;  g_sigma_i = delta_i g_x_i
// with T_{i+1} = T_i e^{-x_i}.  Forward quantities are recomputed (24 B/sample read instead of 28).
// Chunks are walked from the far end so the suffix sum is a running carry (true exclusive suffix scan);
// the prefix sums of x needed for T are taken from a first pass that stores the chunk totals.
__global__ void __launch_bounds__(kCompWarps * 32)
    composite_bwd_kernel(const float* __restrict__ sigma, const float* __restrict__ radiance,
                         const float* __restrict__ delta, const float* __restrict__ g_rgb,
                         const float* __restrict__ g_w_ext, int64_t n, int s, float* __restrict__ g_sigma,
                         float* __restrict__ g_radiance) {
  __shared__ float stage_all[kCompWarps][96];
  __shared__ double chunk_tot[kCompWarps][64];  // up to 2048 samples per ray
  const int warp = threadIdx.x >> 5, lane = lane_id();
  float* stage = stage_all[warp];
  int64_t ray = (int64_t)blockIdx.x * kCompWarps + warp;
  if (ray >= n) return;
  const float gr = __ldg(g_rgb + 3 * ray), gg = __ldg(g_rgb + 3 * ray + 1), gb = __ldg(g_rgb + 3 * ray + 2);
  const int nchunk = (s + 31) / 32;
  // pass 1: float64 totals of x per chunk -> exclusive chunk offsets
  {
    double run = 0.0;
    for (int c = 0; c < nchunk; ++c) {
      int i = c * 32 + lane;
      float x = (i < s) ? __fmul_rn(__ldg(sigma + ray * s + i), __ldg(delta + ray * s + i)) : 0.f;
      double v = (double)x;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += shfl_down_f64(v, d);
      v = shfl_f64(v, 0);
      if (lane == 0) chunk_tot[warp][c] = run;  // sum of all x before this chunk
      run += v;
    }
  }
  __syncwarp();
  double suffix = 0.0;  // sum_{k > last element of this chunk} g_w_k w_k
  for (int c = nchunk - 1; c >= 0; --c) {
    const int base = c * 32;
    const int i = base + lane;
    const bool ok = i < s;
    const int cnt = min(32, s - base);
    float sg = 0.f, dl = 0.f;
    if (ok) {
      sg = __ldg(sigma + ray * s + i);
      dl = __ldg(delta + ray * s + i);
    }
    float x = __fmul_rn(sg, dl);
    double inc = (double)x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      double up = shfl_up_f64(inc, d);
      if (lane >= d) inc += up;
    }
    inc += chunk_tot[warp][c];
    float csum = (float)inc;
    float excl = __shfl_up_sync(0xffffffffu, csum, 1);
    if (lane == 0) excl = (c == 0) ? 0.f : (float)chunk_tot[warp][c];
    float trans = expf(-excl);
    float ex = expf(-x);
    float w = ok ? __fmul_rn(trans, __fsub_rn(1.0f, ex)) : 0.f;
    float r, g, b;
    load_rgb_chunk(radiance + (ray * s + base) * 3, cnt, stage, r, g, b);
    float gw = gr * r + gg * g + gb * b;
    if (g_w_ext && ok) gw += __ldg(g_w_ext + ray * s + i);
    // exclusive suffix scan of gw*w inside the chunk (float64), plus the carry from later chunks
    double p = ok ? (double)gw * (double)w : 0.0;
    double sfx = p;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      double dn = shfl_down_f64(sfx, d);
      if (lane + d < 32) sfx += dn;
    }
    double sfx_excl = shfl_down_f64(sfx, 1);
    if (lane == 31) sfx_excl = 0.0;
    sfx_excl += suffix;
    if (ok) {
      float gx = (float)((double)gw * (double)(trans * ex) - sfx_excl);
      g_sigma[ray * s + i] = dl * gx;
    }
    // g_c = w * g_rgb, staged for coalesced stores
    __syncwarp();
    if (ok) {
      stage[3 * lane] = w * gr;
      stage[3 * lane + 1] = w * gg;
      stage[3 * lane + 2] = w * gb;
    }
    __syncwarp();
    float* dst = g_radiance + (ray * s + base) * 3;
    for (int e = lane; e < 3 * cnt; e += 32) dst[e] = stage[e];
    suffix += shfl_f64(sfx, 0);
  }
}

// ------------------------------------------------------------------------------------------------
// K7 / K8 for rays of at most 32*G samples (G = 2, 4, 6, 8; the reference's 64 and 192 are G = 2 and 6).  Same
// arithmetic as the chunk-walking kernels above; what changes is the memory schedule: a warp issues EVERY load of its
// ray (sigma, delta, t, the 3*S radiance floats as consecutive 128-byte rows) before the first dependent instruction,
// so ~3.8 KB per warp are in flight instead of one 640-byte chunk, and each input is read exactly once.  The forward
// keeps the samples striped over the lanes (chunk scans interleaved); the backward is the blocked kernel further down.
// ------------------------------------------------------------------------------------------------
// inclusive float64 scans of G independent 32-element chunks, interleaved so the G shuffle/add chains overlap
template <int G>
__device__ __forceinline__ void chunk_scans_up(double (&inc)[G]) {
  const int lane = lane_id();
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    double up[G];
#pragma unroll
    for (int c = 0; c < G; ++c) up[c] = shfl_up_f64(inc[c], d);
#pragma unroll
    for (int c = 0; c < G; ++c)
      if (lane >= d) inc[c] += up[c];
  }
}
// Transmittance of every sample of the ray: x = sigma*delta (:41), float64 prefix sums rounded per element like torch's
// cumsum, shifted to a true exclusive scan (:44-52).  The running sum across chunks is carry_c = carry_{c-1} + (chunk
// c-1's total), the same additions the chunk-walking kernel makes.
template <int G>
__device__ __forceinline__ void transmittance(const float (&x)[G], float (&trans)[G]) {
  const int lane = lane_id();
  double inc[G];
#pragma unroll
  for (int c = 0; c < G; ++c) inc[c] = (double)x[c];
  chunk_scans_up<G>(inc);
  double carry = 0.0;
#pragma unroll
  for (int c = 0; c < G; ++c) {
    const double tot = shfl_f64(inc[c], 31);
    const float csum = (float)(inc[c] + carry);  // torch cumsum output element (float32)
    float excl = __shfl_up_sync(0xffffffffu, csum, 1);
    if (lane == 0) excl = (float)carry;
    trans[c] = expf(-excl);
    carry += tot;
  }
}

template <int G>
__global__ void __launch_bounds__(kCompWarps * 32)
    composite_fwd_reg_kernel(const float* __restrict__ sigma, const float* __restrict__ radiance,
                             const float* __restrict__ delta, const float* __restrict__ tvals, int64_t n, int s,
                             float* __restrict__ rgb_out, float* __restrict__ w_out, float* __restrict__ depth_out,
                             float* __restrict__ opacity_out) {
  __shared__ float stage_all[kCompWarps][96 * G];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  float* stage = stage_all[warp];
  const int64_t ray = (int64_t)blockIdx.x * kCompWarps + warp;
  if (ray >= n) return;
  const float* sg_row = sigma + ray * s;
  const float* dl_row = delta + ray * s;
  const float* rad_row = radiance + ray * s * 3;
  float x[G], tv[G];
  {
    float sg[G], dl[G], rad[3 * G];
#pragma unroll
    for (int c = 0; c < G; ++c) {
      const int i = 32 * c + lane;
      const bool ok = i < s;
      sg[c] = ok ? __ldg(sg_row + i) : 0.f;
      dl[c] = ok ? __ldg(dl_row + i) : 0.f;
      tv[c] = (depth_out && ok) ? __ldg(tvals + ray * s + i) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 3 * G; ++k) {
      const int e = 32 * k + lane;
      rad[k] = e < 3 * s ? __ldg(rad_row + e) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 3 * G; ++k) stage[32 * k + lane] = rad[k];
#pragma unroll
    for (int c = 0; c < G; ++c) x[c] = __fmul_rn(sg[c], dl[c]);  // :41
  }
  __syncwarp();
  float trans[G];
  transmittance<G>(x, trans);
  float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_d = 0.f, acc_o = 0.f;
#pragma unroll
  for (int c = 0; c < G; ++c) {
    const int i = 32 * c + lane;
    const bool ok = i < s;
    const float alpha = __fsub_rn(1.0f, expf(-x[c]));       // :55
    const float w = ok ? __fmul_rn(trans[c], alpha) : 0.f;  // :58
    if (ok && w_out) w_out[ray * s + i] = w;
    acc_r = fmaf(w, stage[3 * i], acc_r);
    acc_g = fmaf(w, stage[3 * i + 1], acc_g);
    acc_b = fmaf(w, stage[3 * i + 2], acc_b);
    if (depth_out) acc_d = fmaf(w, tv[c], acc_d);
    acc_o += w;
  }
  acc_r = warp_sum(acc_r), acc_g = warp_sum(acc_g), acc_b = warp_sum(acc_b);
  if (depth_out) acc_d = warp_sum(acc_d);
  if (opacity_out) acc_o = warp_sum(acc_o);
  if (lane == 0) {
    rgb_out[3 * ray + 0] = acc_r;
    rgb_out[3 * ray + 1] = acc_g;
    rgb_out[3 * ray + 2] = acc_b;
    if (depth_out) depth_out[ray] = acc_d;
    if (opacity_out) opacity_out[ray] = acc_o;
  }
}

// ------------------------------------------------------------------------------------------------
// K8 in BLOCKED order: lane l owns the G consecutive samples [G l, G l + G) of its ray.  A striped backward (samples 32 c + lane, like the forward above) spends
// the SM's data pipe on 64-bit shuffles (two float64 scans of G chunks = ~170 shuffles per ray, ncu: L1 pipe 80 % busy
// at 61 % DRAM); here the scans are serial inside a lane plus ONE warp scan of the lane totals (~25 shuffles).  Global
// accesses stay coalesced: rows are staged through shared memory, stored striped and read blocked (and the reverse for
// the outputs) with one pad float per lane block so that the blocked accesses have an odd stride (no bank conflicts).
// ------------------------------------------------------------------------------------------------
template <int G>
struct BlkLayout {
  static constexpr int kS1 = G + 1;        // floats per lane block of a per-sample row (sigma, delta, g_w, g_sigma)
  static constexpr int kS3 = 3 * G + 1;    // floats per lane block of the radiance / g_radiance row
  static constexpr int kRow1 = 32 * kS1, kRow3 = 32 * kS3;
  static constexpr int kWarpFloats = 2 * kRow1 + kRow3;  // sigma (reused for g_sigma) | delta (also g_w) | radiance
  __device__ static __forceinline__ int pos1(int e) { return e + e / G; }          // striped element e -> padded slot
  __device__ static __forceinline__ int pos3(int f) { return f + f / (3 * G); }
};

template <int G, int MB, bool kMse>
__global__ void __launch_bounds__(kCompWarps * 32, MB)
    composite_bwd_blk_kernel(const float* __restrict__ sigma, const float* __restrict__ radiance,
                             const float* __restrict__ delta, const float* __restrict__ g_rgb,
                             const float* __restrict__ g_w_ext, int64_t n, int s, float* __restrict__ g_sigma,
                             float* __restrict__ g_radiance, const float* __restrict__ mse_target, float inv_cnt,
                             float* __restrict__ loss_accum) {
  using L = BlkLayout<G>;
  __shared__ float stage_all[kCompWarps][L::kWarpFloats];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  float* s_sg = stage_all[warp];
  float* s_dl = s_sg + L::kRow1;
  float* s_rad = s_dl + L::kRow1;
  const int64_t ray = (int64_t)blockIdx.x * kCompWarps + warp;
  if (ray >= n) return;
  const float* sg_row = sigma + ray * s;
  const float* dl_row = delta + ray * s;
  const float* rad_row = radiance + ray * s * 3;
  {
    // every load of the ray first, then the striped -> blocked hand-over through shared memory
    float sg[G], dl[G], ge[G], rad[3 * G];
#pragma unroll
    for (int c = 0; c < G; ++c) {
      const int e = 32 * c + lane;
      const bool ok = e < s;
      sg[c] = ok ? __ldg(sg_row + e) : 0.f;
      dl[c] = ok ? __ldg(dl_row + e) : 0.f;
      ge[c] = (g_w_ext && ok) ? __ldg(g_w_ext + ray * s + e) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 3 * G; ++k) {
      const int f = 32 * k + lane;
      rad[k] = f < 3 * s ? __ldg(rad_row + f) : 0.f;
    }
#pragma unroll
    for (int c = 0; c < G; ++c) {
      const int e = 32 * c + lane;
      s_sg[L::pos1(e)] = __fmul_rn(sg[c], dl[c]);  // x = sigma * delta (:41)
      s_dl[L::pos1(e)] = dl[c];
    }
#pragma unroll
    for (int k = 0; k < 3 * G; ++k) s_rad[L::pos3(32 * k + lane)] = rad[k];
    __syncwarp();
    // blocked reads
    float x[G], dlb[G], gw[G];
    // g_rgb is either given, or (training step) `g_rgb` holds the rendered colour and the MSE head is folded in here:
    // g = 2/(3N) (rgb - target), loss += mean((rgb - target)^2)   (runner_utils.py:731, train.py:180/202)
    float gr = __ldg(g_rgb + 3 * ray), gg = __ldg(g_rgb + 3 * ray + 1), gb = __ldg(g_rgb + 3 * ray + 2);
    if (kMse) {
      const float dr = gr - __ldg(mse_target + 3 * ray), dg = gg - __ldg(mse_target + 3 * ray + 1),
                  db = gb - __ldg(mse_target + 3 * ray + 2);
      gr = 2.0f * inv_cnt * dr, gg = 2.0f * inv_cnt * dg, gb = 2.0f * inv_cnt * db;
      if (lane == 0) atomicAdd(loss_accum, (dr * dr + dg * dg + db * db) * inv_cnt);
    }
#pragma unroll
    for (int k = 0; k < G; ++k) {
      x[k] = s_sg[lane * L::kS1 + k];
      dlb[k] = s_dl[lane * L::kS1 + k];
      const float r = s_rad[lane * L::kS3 + 3 * k], g = s_rad[lane * L::kS3 + 3 * k + 1], b = s_rad[lane * L::kS3 + 3 * k + 2];
      gw[k] = gr * r + gg * g + gb * b;
    }
    if (g_w_ext) {  // the external weight gradient goes through the x slots once x has been read
      __syncwarp();
#pragma unroll
      for (int c = 0; c < G; ++c) s_sg[L::pos1(32 * c + lane)] = ge[c];
      __syncwarp();
#pragma unroll
      for (int k = 0; k < G; ++k) gw[k] += s_sg[lane * L::kS1 + k];
    }
    // transmittance: float64 prefix of x inside the lane, one warp scan of the lane totals, every prefix rounded to
    // float32 like torch's cumsum output, shifted to a true exclusive scan (:44-52)
    double pre[G];
    double run = 0.0;
#pragma unroll
    for (int k = 0; k < G; ++k) {
      run += (double)x[k];
      pre[k] = run;
    }
    double off = run;  // inclusive scan of the lane totals ...
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double up = shfl_up_f64(off, d);
      if (lane >= d) off += up;
    }
    off = shfl_up_f64(off, 1);  // ... shifted: the sum of everything before this lane's first sample
    if (lane == 0) off = 0.0;
    float w[G], tnext[G];
    double sfx[G];
    float excl = (float)off;
#pragma unroll
    for (int k = 0; k < G; ++k) {
      const bool ok = G * lane + k < s;
      const float trans = expf(-excl);
      const float ex = expf(-x[k]);
      w[k] = ok ? __fmul_rn(trans, __fsub_rn(1.0f, ex)) : 0.f;
      tnext[k] = trans * ex;
      excl = (float)(off + pre[k]);
    }
    // exclusive suffix sums of g_w * w: inside the lane, then across the later lanes
    double tail = 0.0;
#pragma unroll
    for (int k = G - 1; k >= 0; --k) {
      sfx[k] = tail;
      tail += (double)gw[k] * (double)w[k];
    }
    double later = tail;  // inclusive suffix scan of the lane totals
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double dn = shfl_down_f64(later, d);
      if (lane + d < 32) later += dn;
    }
    later = shfl_down_f64(later, 1);
    if (lane == 31) later = 0.0;
    __syncwarp();  // all blocked reads are done: the staging rows now collect the outputs
#pragma unroll
    for (int k = 0; k < G; ++k) {
      const float gx = (float)((double)gw[k] * (double)tnext[k] - (sfx[k] + later));
      s_sg[lane * L::kS1 + k] = dlb[k] * gx;
      s_rad[lane * L::kS3 + 3 * k] = w[k] * gr;
      s_rad[lane * L::kS3 + 3 * k + 1] = w[k] * gg;
      s_rad[lane * L::kS3 + 3 * k + 2] = w[k] * gb;
    }
    __syncwarp();
  }
  float* gs_row = g_sigma + ray * s;
  float* gc_row = g_radiance + ray * s * 3;
#pragma unroll
  for (int c = 0; c < G; ++c) {
    const int e = 32 * c + lane;
    if (e < s) gs_row[e] = s_sg[BlkLayout<G>::pos1(e)];
  }
#pragma unroll
  for (int k = 0; k < 3 * G; ++k) {
    const int f = 32 * k + lane;
    if (f < 3 * s) gc_row[f] = s_rad[BlkLayout<G>::pos3(f)];
  }
}

template <int G>
static void launch_composite_fwd_reg(const float* sigma, const float* radiance, const float* delta, const float* t,
                                     int64_t n, int s, float* rgb, float* w, float* depth, float* opacity,
                                     cudaStream_t stream) {
  composite_fwd_reg_kernel<G><<<(unsigned)ceil_div64(n, kCompWarps), kCompWarps * 32, 0, stream>>>(
      sigma, radiance, delta, t, n, s, rgb, w, depth, opacity);
}
template <int G>
static void launch_composite_bwd_blk(const float* sigma, const float* radiance, const float* delta, const float* g_rgb,
                                     const float* g_w, int64_t n, int s, float* g_sigma, float* g_radiance,
                                     cudaStream_t stream, const float* mse_target = nullptr, float inv_cnt = 0.f,
                                     float* loss_accum = nullptr) {
  // 6 CTAs per SM (80 registers per thread) measured best for 192 samples: 945 us at the unconstrained 95 registers,
  // 866 us at 80, 875 us at 64 (640 000 rays)
  // CTAs per SM measured on 192 samples (640 000 rays): 5 -> 778 us, 6 -> 768 us, 7 (72 registers) -> 704 us, 8 -> 797 us
  constexpr int kMinBlocks = G <= 6 ? 7 : 4;
  if (mse_target != nullptr)
    composite_bwd_blk_kernel<G, kMinBlocks, true><<<(unsigned)ceil_div64(n, kCompWarps), kCompWarps * 32, 0, stream>>>(
        sigma, radiance, delta, g_rgb, g_w, n, s, g_sigma, g_radiance, mse_target, inv_cnt, loss_accum);
  else
    composite_bwd_blk_kernel<G, kMinBlocks, false><<<(unsigned)ceil_div64(n, kCompWarps), kCompWarps * 32, 0, stream>>>(
        sigma, radiance, delta, g_rgb, g_w, n, s, g_sigma, g_radiance, nullptr, 0.f, nullptr);
}

}  // namespace nerf

using namespace nerf;

extern "C" {

int nerf_posenc(const float* x_dev, int64_t m, int in_dim, int embed_level, int include_input, float* out_dev,
                int64_t ld_out, nerf_stream_t stream) {
  if (m == 0) return NERF_OK;
  NERF_CHECK_ARG(x_dev && out_dev, "nerf_posenc: null pointer");
  NERF_CHECK_ARG(m >= 0 && in_dim > 0 && embed_level >= 0 && embed_level <= 32, "nerf_posenc: bad sizes");
  const int out_dim = in_dim * (2 * embed_level + (include_input ? 1 : 0));
  NERF_CHECK_ARG(out_dim > 0 && ld_out >= out_dim, "nerf_posenc: ld_out smaller than the encoding width");
  if (m == 0) return NERF_OK;
  size_t smem = sizeof(float) * kEncRows * out_dim;
  NERF_CHECK_ARG(smem <= 200 * 1024, "nerf_posenc: encoding too wide");
  if (smem > 48 * 1024)
    NERF_CUDA(cudaFuncSetAttribute(posenc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  posenc_kernel<<<(unsigned)ceil_div64(m, kEncRows), 256, smem, as_stream(stream)>>>(x_dev, m, in_dim, embed_level,
                                                                                   include_input, out_dev, ld_out);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_composite_fwd(const float* sigma_dev, const float* radiance_dev, const float* delta_dev,
                       const float* t_dev, int64_t n, int s, float* rgb_dev, float* w_dev, float* depth_dev,
                       float* opacity_dev, nerf_stream_t stream) {
  NERF_CHECK_ARG(n >= 0 && s > 0, "nerf_composite_fwd: bad sizes");
  if (n == 0) return NERF_OK;
  NERF_CHECK_ARG(sigma_dev && radiance_dev && delta_dev && rgb_dev, "nerf_composite_fwd: null pointer");
  NERF_CHECK_ARG(depth_dev == nullptr || t_dev != nullptr, "nerf_composite_fwd: depth needs t");
  if (n == 0) return NERF_OK;
  cudaStream_t cs = as_stream(stream);
  if (s <= 64) launch_composite_fwd_reg<2>(sigma_dev, radiance_dev, delta_dev, t_dev, n, s, rgb_dev, w_dev, depth_dev, opacity_dev, cs);
  else if (s <= 128) launch_composite_fwd_reg<4>(sigma_dev, radiance_dev, delta_dev, t_dev, n, s, rgb_dev, w_dev, depth_dev, opacity_dev, cs);
  else if (s <= 192) launch_composite_fwd_reg<6>(sigma_dev, radiance_dev, delta_dev, t_dev, n, s, rgb_dev, w_dev, depth_dev, opacity_dev, cs);
  else if (s <= 256) launch_composite_fwd_reg<8>(sigma_dev, radiance_dev, delta_dev, t_dev, n, s, rgb_dev, w_dev, depth_dev, opacity_dev, cs);
  else
    composite_fwd_kernel<<<(unsigned)ceil_div64(n, kCompWarps), kCompWarps * 32, 0, cs>>>(
        sigma_dev, radiance_dev, delta_dev, t_dev, n, s, rgb_dev, w_dev, depth_dev, opacity_dev);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_composite_bwd(const float* sigma_dev, const float* radiance_dev, const float* delta_dev,
                       const float* g_rgb_dev, const float* g_w_dev, int64_t n, int s, float* g_sigma_dev,
                       float* g_radiance_dev, nerf_stream_t stream) {
  NERF_CHECK_ARG(n >= 0 && s > 0 && s <= 2048, "nerf_composite_bwd: samples per ray must be in [1,2048]");
  if (n == 0) return NERF_OK;
  NERF_CHECK_ARG(sigma_dev && radiance_dev && delta_dev && g_rgb_dev && g_sigma_dev && g_radiance_dev,
                 "nerf_composite_bwd: null pointer");
  if (n == 0) return NERF_OK;
  cudaStream_t cs = as_stream(stream);
  if (s <= 64) launch_composite_bwd_blk<2>(sigma_dev, radiance_dev, delta_dev, g_rgb_dev, g_w_dev, n, s, g_sigma_dev, g_radiance_dev, cs);
  else if (s <= 128) launch_composite_bwd_blk<4>(sigma_dev, radiance_dev, delta_dev, g_rgb_dev, g_w_dev, n, s, g_sigma_dev, g_radiance_dev, cs);
  else if (s <= 192) launch_composite_bwd_blk<6>(sigma_dev, radiance_dev, delta_dev, g_rgb_dev, g_w_dev, n, s, g_sigma_dev, g_radiance_dev, cs);
  else if (s <= 256) launch_composite_bwd_blk<8>(sigma_dev, radiance_dev, delta_dev, g_rgb_dev, g_w_dev, n, s, g_sigma_dev, g_radiance_dev, cs);
  else
    composite_bwd_kernel<<<(unsigned)ceil_div64(n, kCompWarps), kCompWarps * 32, 0, cs>>>(
        sigma_dev, radiance_dev, delta_dev, g_rgb_dev, g_w_dev, n, s, g_sigma_dev, g_radiance_dev);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

// The training step's loss head folded into the compositing backward: rgb_dev (N,3) is the rendered colour, target_dev
// (N,3) the ground truth; g_rgb = 2/(3N) (rgb - target) never goes to memory, *loss_accum_dev += mean((rgb - target)^2).
int nerf_composite_bwd_mse(const float* sigma_dev, const float* radiance_dev, const float* delta_dev, const float* rgb_dev,
                           const float* target_dev, int64_t n, int s, float* g_sigma_dev, float* g_radiance_dev,
                           float* loss_accum_dev, nerf_stream_t stream) {
  NERF_CHECK_ARG(n >= 0 && s > 0 && s <= 256, "nerf_composite_bwd_mse: samples per ray must be in [1,256]");
  if (n == 0) return NERF_OK;
  NERF_CHECK_ARG(sigma_dev && radiance_dev && delta_dev && rgb_dev && target_dev && g_sigma_dev && g_radiance_dev && loss_accum_dev,
                 "nerf_composite_bwd_mse: null pointer");
  cudaStream_t cs = as_stream(stream);
  const float inv_cnt = 1.0f / (float)(3 * n);
  if (s <= 64) launch_composite_bwd_blk<2>(sigma_dev, radiance_dev, delta_dev, rgb_dev, nullptr, n, s, g_sigma_dev, g_radiance_dev, cs, target_dev, inv_cnt, loss_accum_dev);
  else if (s <= 128) launch_composite_bwd_blk<4>(sigma_dev, radiance_dev, delta_dev, rgb_dev, nullptr, n, s, g_sigma_dev, g_radiance_dev, cs, target_dev, inv_cnt, loss_accum_dev);
  else if (s <= 192) launch_composite_bwd_blk<6>(sigma_dev, radiance_dev, delta_dev, rgb_dev, nullptr, n, s, g_sigma_dev, g_radiance_dev, cs, target_dev, inv_cnt, loss_accum_dev);
  else launch_composite_bwd_blk<8>(sigma_dev, radiance_dev, delta_dev, rgb_dev, nullptr, n, s, g_sigma_dev, g_radiance_dev, cs, target_dev, inv_cnt, loss_accum_dev);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// MSE loss head of the training step (runner_utils.py:731 nn.MSELoss, train.py:180/202):
//   loss += mean((rgb - target)^2) over N*3 elements;  g_rgb = 2/(3N) * (rgb - target)
// ------------------------------------------------------------------------------------------------
namespace nerf {
__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ rgb, const float* __restrict__ target,
                                                  int64_t cnt, float inv_cnt, float* __restrict__ g_rgb,
                                                  float* __restrict__ loss_accum) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float sq = 0.f;
  if (e < cnt) {
    float d = rgb[e] - target[e];
    g_rgb[e] = 2.0f * inv_cnt * d;
    sq = d * d;
  }
  sq = warp_sum(sq);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(loss_accum, t * inv_cnt);
  }
}
}  // namespace nerf

extern "C" int nerf_mse_loss(const float* rgb_dev, const float* target_dev, int64_t n, float* g_rgb_dev,
                             float* loss_accum_dev, nerf_stream_t stream) {
  NERF_CHECK_ARG(n >= 0, "nerf_mse_loss: negative ray count");
  if (n == 0) return NERF_OK;
  NERF_CHECK_ARG(rgb_dev && target_dev && g_rgb_dev && loss_accum_dev, "nerf_mse_loss: null pointer");
  const int64_t cnt = 3 * n;
  nerf::mse_kernel<<<(unsigned)nerf::ceil_div64(cnt, 256), 256, 0, nerf::as_stream(stream)>>>(
      rgb_dev, target_dev, cnt, 1.0f / (float)cnt, g_rgb_dev, loss_accum_dev);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}
