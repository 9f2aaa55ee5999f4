// K5/K6 in fp32 VALIDATION mode: the NeRF MLP (network/nerf.py:65-121) and its backward as a chain of
// CUDA-core SGEMMs with fused bias / ReLU / sigmoid / ReLU-mask epilogues.  This path exists to hold the
// <= 1e-3 parity gate against the reference's fp32 arithmetic (tcgen05 has no fp32-input MMA); the
// performance path is the bf16 tensor-core chain in mlp_tc.cu.
//
// Concatenations of the reference are realised by layout, not copies of activations:
//   x5 = cat[pos, h4]  (nerf.py:108)  : fc_4 writes h4 straight into x5[:, P:]
//   z  = [sigma_pre | feat | view]    : fc_8 writes its 1+F outputs into z[:, 0:1+F]; fc_9 reads z[:, 1:]
//                                       (= cat[out[:,1:], view_dir], nerf.py:116)
#include <math.h>

#include "common.cuh"

namespace nerf {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2 };

// C[m,n] (+)= epi( sum_k A(m,k) * B(n,k) )
//   A(m,k) = A[m*a_rs + k*a_ks], B(n,k) = B[n*b_rs + k*b_ks]   (one of the two strides is 1)
//   epi: + bias[n]; activation; * (mask[m*ld_mask + n] > 0); atomicAdd when split-K
struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  const float* bias;
  const float* mask;
  int64_t M;
  int N, K;
  int64_t a_rs, a_ks, b_rs, b_ks, ldc, ld_mask;
  int act;
  int64_t k_chunk;  // K range per blockIdx.z
  int atomic;
};

constexpr int TM = 128, TN = 128, TK = 16;

template <bool A_KCONTIG, bool B_KCONTIG>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 8x8 outputs each (strided by 16)
  const int64_t m0 = (int64_t)blockIdx.y * TM;
  const int n0 = blockIdx.x * TN;
  const int64_t kbeg = (int64_t)blockIdx.z * g.k_chunk;
  const int64_t kend = min((int64_t)g.K, kbeg + g.k_chunk);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += TK) {
    // tile loads: consecutive threads walk the contiguous dimension
#pragma unroll
    for (int e = tid; e < TM * TK; e += 256) {
      int r, k;
      if (A_KCONTIG) { k = e % TK; r = e / TK; } else { r = e % TM; k = e / TM; }
      int64_t gm = m0 + r, gk = k0 + k;
      As[k][r] = (gm < g.M && gk < kend) ? __ldg(g.A + gm * g.a_rs + gk * g.a_ks) : 0.f;
    }
#pragma unroll
    for (int e = tid; e < TN * TK; e += 256) {
      int r, k;
      if (B_KCONTIG) { k = e % TK; r = e / TK; } else { r = e % TN; k = e / TN; }
      int64_t gn = n0 + r, gk = k0 + k;
      Bs[k][r] = (gn < g.N && gk < kend) ? __ldg(g.B + gn * g.b_rs + gk * g.b_ks) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t gm = m0 + ty + 16 * i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int gn = n0 + tx + 16 * j;
      if (gn >= g.N) continue;
      float v = acc[i][j];
      if (g.atomic) {
        atomicAdd(g.C + gm * g.ldc + gn, v);
        continue;
      }
      if (g.bias) v += __ldg(g.bias + gn);
      if (g.act == ACT_RELU) v = fmaxf(v, 0.f);
      else if (g.act == ACT_SIGMOID) v = 1.0f / (1.0f + expf(-v));
      if (g.mask) v = (__ldg(g.mask + gm * g.ld_mask + gn) > 0.f) ? v : 0.f;
      g.C[gm * g.ldc + gn] = v;
    }
  }
}

static int launch_gemm(const GemmArgs& g, bool a_kcontig, bool b_kcontig, int splits, cudaStream_t st) {
  GemmArgs a = g;
  a.k_chunk = (splits > 1) ? ((g.K + splits - 1) / splits + TK - 1) / TK * TK : g.K;
  if (a.k_chunk <= 0) a.k_chunk = TK;
  int zs = (splits > 1) ? (int)((g.K + a.k_chunk - 1) / a.k_chunk) : 1;
  a.atomic = splits > 1;
  dim3 grid((g.N + TN - 1) / TN, (unsigned)((g.M + TM - 1) / TM), zs);
  if (a_kcontig && b_kcontig) sgemm_kernel<true, true><<<grid, 256, 0, st>>>(a);
  else if (a_kcontig && !b_kcontig) sgemm_kernel<true, false><<<grid, 256, 0, st>>>(a);
  else if (!a_kcontig && b_kcontig) sgemm_kernel<false, true><<<grid, 256, 0, st>>>(a);
  else sgemm_kernel<false, false><<<grid, 256, 0, st>>>(a);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

// y[M,N] (ldy) = act(x[M,K] (ldx) . W[N,K]^T (ldw) + b)
static int linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* b, float* y, int64_t ldy,
                      int64_t m, int n, int k, int act, cudaStream_t st) {
  GemmArgs g{};
  g.A = x, g.B = w, g.C = y, g.bias = b, g.mask = nullptr;
  g.M = m, g.N = n, g.K = k;
  g.a_rs = ldx, g.a_ks = 1, g.b_rs = ldw, g.b_ks = 1, g.ldc = ldy, g.ld_mask = 0, g.act = act;
  return launch_gemm(g, true, true, 1, st);
}

// dx[M,K'] (lddx) = (g[M,N] (ldg) . W[N, koff:koff+K'] (ldw)) * (mask > 0)
static int linear_dgrad(const float* gr, int64_t ldg, const float* w, int64_t ldw, float* dx, int64_t lddx,
                        const float* mask, int64_t ldmask, int64_t m, int n, int kout, cudaStream_t st) {
  GemmArgs g{};
  g.A = gr, g.B = w, g.C = dx, g.bias = nullptr, g.mask = mask;
  g.M = m, g.N = kout, g.K = n;           // reduction over the layer's outputs
  g.a_rs = ldg, g.a_ks = 1;               // A(m, o) = g[m*ldg + o]
  g.b_rs = 1, g.b_ks = ldw;               // B(i, o) = W[o*ldw + i]
  g.ldc = lddx, g.ld_mask = ldmask, g.act = ACT_NONE;
  return launch_gemm(g, true, false, 1, st);
}

// dW[N,K] (contiguous) = g[M,N]^T . x[M,K]; split over M with atomics into a zeroed dW
static int linear_wgrad(const float* gr, int64_t ldg, const float* x, int64_t ldx, float* dw, int64_t m, int n, int k,
                        cudaStream_t st) {
  NERF_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)n * k, st));
  GemmArgs g{};
  g.A = gr, g.B = x, g.C = dw, g.bias = nullptr, g.mask = nullptr;
  g.M = n, g.N = k, g.K = (int)m;  // reduction over rows
  g.a_rs = 1, g.a_ks = ldg;        // A(o, row) = g[row*ldg + o]
  g.b_rs = 1, g.b_ks = ldx;        // B(i, row) = x[row*ldx + i]
  g.ldc = k, g.act = ACT_NONE;
  int tiles = ((n + TM - 1) / TM) * ((k + TN - 1) / TN);
  int splits = (2 * sm_count() + tiles - 1) / tiles;
  int max_splits = (int)((m + 4 * TK - 1) / (4 * TK));
  if (splits > max_splits) splits = max_splits;
  if (splits < 2) splits = 2;  // keep the atomic path (dW is zero-initialised)
  return launch_gemm(g, false, false, splits, st);
}

// db[n] = sum_m g[m,n]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ g, int64_t ldg, int64_t m, int n,
                                                      int64_t rows_per_block, float* __restrict__ db) {
  int col = blockIdx.x * 32 + (threadIdx.x & 31);
  int sub = threadIdx.x >> 5;  // 8 row lanes
  int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = min(m, r0 + rows_per_block);
  float acc = 0.f;
  if (col < n)
    for (int64_t r = r0 + sub; r < r1; r += 8) acc += __ldg(g + r * ldg + col);
  __shared__ float red[8][33];
  red[sub][threadIdx.x & 31] = acc;
  __syncthreads();
  if (sub == 0 && col < n) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
    atomicAdd(db + col, t);
  }
}

static int bias_grad(const float* gr, int64_t ldg, float* db, int64_t m, int n, cudaStream_t st) {
  NERF_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * n, st));
  int64_t rows_per_block = 2048;
  dim3 grid((n + 31) / 32, (unsigned)((m + rows_per_block - 1) / rows_per_block));
  colsum_kernel<<<grid, 256, 0, st>>>(gr, ldg, m, n, rows_per_block, db);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

// strided 2-D copy (M x cols) of fp32
__global__ void copy2d_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd,
                              int64_t m, int cols) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m * cols) return;
  int64_t r = e / cols;
  int c = (int)(e - r * cols);
  dst[r * ldd + c] = src[r * lds + c];
}

// sigma = relu(z[:,0])  (nerf.py:115)
__global__ void sigma_kernel(const float* __restrict__ z, int64_t ldz, int64_t m, float* __restrict__ sigma) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < m) sigma[r] = fmaxf(z[r * ldz], 0.f);
}

// g_z = g_rgb * rgb * (1 - rgb)   (sigmoid backward, nerf.py:119)
__global__ void sigmoid_bwd_kernel(const float* __restrict__ g_rgb, const float* __restrict__ rgb, int64_t cnt,
                                   float* __restrict__ out) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < cnt) {
    float y = rgb[e];
    out[e] = g_rgb[e] * y * (1.0f - y);
  }
}

// g8[:,0] = g_sigma * (z[:,0] > 0)   (relu backward on the density head)
__global__ void sigma_bwd_kernel(const float* __restrict__ g_sigma, const float* __restrict__ z, int64_t ldz,
                                 int64_t m, float* __restrict__ g8, int64_t ldg8) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < m) g8[r * ldg8] = (z[r * ldz] > 0.f) ? g_sigma[r] : 0.f;
}

struct CacheLayout {
  int P, V, F, H;
  int64_t m;
  // offsets in floats
  int64_t h[4];  // outputs of fc_in, fc_1, fc_2, fc_3   (M,F)
  int64_t x5;    // (M, P+F): [pos | h4]
  int64_t h5, h6, h7;  // (M,F)
  int64_t z;     // (M, 1+F+V): [sigma_pre | feat | view]
  int64_t h9;    // (M, F/2)
  int64_t total;
  explicit CacheLayout(const nerf_mlp_dims_t& d, int64_t m_) : P(d.pos_dim), V(d.view_dim), F(d.feat_dim), H(d.feat_dim / 2), m(m_) {
    int64_t off = 0;
    for (int i = 0; i < 4; ++i) { h[i] = off; off += m * F; }
    x5 = off; off += m * (P + F);
    h5 = off; off += m * F;
    h6 = off; off += m * F;
    h7 = off; off += m * F;
    z = off; off += m * (1 + F + V);
    h9 = off; off += m * H;
    total = off;
  }
};

static bool dims_ok(const nerf_mlp_dims_t* d) {
  return d && d->pos_dim > 0 && d->view_dim > 0 && d->feat_dim > 1 && d->feat_dim % 2 == 0;
}

}  // namespace nerf

using namespace nerf;

// parameter slots in state_dict order
enum { W_IN = 0, B_IN, W_1, B_1, W_2, B_2, W_3, B_3, W_4, B_4, W_5, B_5, W_6, B_6, W_7, B_7, W_8, B_8, W_9, B_9, W_OUT, B_OUT };

extern "C" {

size_t nerf_mlp_f32_cache_floats(const nerf_mlp_dims_t* dims, int64_t m) {
  if (!dims_ok(dims) || m < 0) return 0;
  return (size_t)CacheLayout(*dims, m).total;
}

size_t nerf_mlp_f32_bwd_scratch_floats(const nerf_mlp_dims_t* dims, int64_t m) {
  if (!dims_ok(dims) || m < 0) return 0;
  // two ping-pong gradient buffers of width 1+F, plus g_z (M,3)
  return (size_t)(2 * m * (1 + dims->feat_dim) + 3 * m);
}

int nerf_mlp_f32_forward(const nerf_mlp_dims_t* dims, const float* const* params, const float* pos_dev,
                         const float* view_dev, int64_t m, float* sigma_dev, float* rgb_dev, float* cache_dev,
                         nerf_stream_t stream) {
  NERF_CHECK_ARG(dims_ok(dims), "nerf_mlp_f32_forward: bad dims");
  NERF_CHECK_ARG(params && pos_dev && view_dev && sigma_dev && rgb_dev && cache_dev, "nerf_mlp_f32_forward: null pointer");
  NERF_CHECK_ARG(m >= 0, "nerf_mlp_f32_forward: negative row count");
  if (m == 0) return NERF_OK;
  cudaStream_t st = as_stream(stream);
  CacheLayout L(*dims, m);
  const int P = L.P, V = L.V, F = L.F, H = L.H;
  float* c = cache_dev;
  int rc;
  const int T = 256;
  // x5[:, :P] = pos ; z[:, 1+F:] = view
  copy2d_kernel<<<(unsigned)ceil_div64(m * P, T), T, 0, st>>>(pos_dev, P, c + L.x5, P + F, m, P);
  copy2d_kernel<<<(unsigned)ceil_div64(m * V, T), T, 0, st>>>(view_dev, V, c + L.z + 1 + F, 1 + F + V, m, V);
  NERF_LAUNCH_CHECK();
  // nerf.py:102-106
  if ((rc = linear_fwd(pos_dev, P, params[W_IN], P, params[B_IN], c + L.h[0], F, m, F, P, ACT_RELU, st))) return rc;
  for (int i = 1; i <= 3; ++i)
    if ((rc = linear_fwd(c + L.h[i - 1], F, params[W_IN + 2 * i], F, params[B_IN + 2 * i], c + L.h[i], F, m, F, F, ACT_RELU, st))) return rc;
  if ((rc = linear_fwd(c + L.h[3], F, params[W_4], F, params[B_4], c + L.x5 + P, P + F, m, F, F, ACT_RELU, st))) return rc;
  // nerf.py:108-113
  if ((rc = linear_fwd(c + L.x5, P + F, params[W_5], P + F, params[B_5], c + L.h5, F, m, F, P + F, ACT_RELU, st))) return rc;
  if ((rc = linear_fwd(c + L.h5, F, params[W_6], F, params[B_6], c + L.h6, F, m, F, F, ACT_RELU, st))) return rc;
  if ((rc = linear_fwd(c + L.h6, F, params[W_7], F, params[B_7], c + L.h7, F, m, F, F, ACT_RELU, st))) return rc;
  if ((rc = linear_fwd(c + L.h7, F, params[W_8], F, params[B_8], c + L.z, 1 + F + V, m, 1 + F, F, ACT_NONE, st))) return rc;
  // nerf.py:115-119
  sigma_kernel<<<(unsigned)ceil_div64(m, T), T, 0, st>>>(c + L.z, 1 + F + V, m, sigma_dev);
  NERF_LAUNCH_CHECK();
  if ((rc = linear_fwd(c + L.z + 1, 1 + F + V, params[W_9], F + V, params[B_9], c + L.h9, H, m, H, F + V, ACT_RELU, st))) return rc;
  if ((rc = linear_fwd(c + L.h9, H, params[W_OUT], H, params[B_OUT], rgb_dev, 3, m, 3, H, ACT_SIGMOID, st))) return rc;
  return NERF_OK;
}

int nerf_mlp_f32_backward(const nerf_mlp_dims_t* dims, const float* const* params, const float* cache_dev,
                          const float* rgb_dev, int64_t m, const float* g_sigma_dev, const float* g_rgb_dev,
                          float* const* grads, float* scratch_dev, nerf_stream_t stream) {
  NERF_CHECK_ARG(dims_ok(dims), "nerf_mlp_f32_backward: bad dims");
  NERF_CHECK_ARG(params && cache_dev && rgb_dev && g_sigma_dev && g_rgb_dev && grads && scratch_dev,
                 "nerf_mlp_f32_backward: null pointer");
  NERF_CHECK_ARG(m > 0, "nerf_mlp_f32_backward: row count must be positive");
  cudaStream_t st = as_stream(stream);
  CacheLayout L(*dims, m);
  const int P = L.P, V = L.V, F = L.F, H = L.H;
  const float* c = cache_dev;
  const int64_t ldg = 1 + F;
  float* ga = scratch_dev;
  float* gb = scratch_dev + m * ldg;
  float* gz = scratch_dev + 2 * m * ldg;
  int rc;
  const int T = 256;
  // fc_out
  sigmoid_bwd_kernel<<<(unsigned)ceil_div64(3 * m, T), T, 0, st>>>(g_rgb_dev, rgb_dev, 3 * m, gz);
  NERF_LAUNCH_CHECK();
  if ((rc = linear_wgrad(gz, 3, c + L.h9, H, grads[W_OUT], m, 3, H, st))) return rc;
  if ((rc = bias_grad(gz, 3, grads[B_OUT], m, 3, st))) return rc;
  // g9 = (gz . W_out) * (h9 > 0)  -> ga (ld H)
  if ((rc = linear_dgrad(gz, 3, params[W_OUT], H, ga, H, c + L.h9, H, m, 3, H, st))) return rc;
  // fc_9: input z[:, 1:] (width F+V)
  if ((rc = linear_wgrad(ga, H, c + L.z + 1, 1 + F + V, grads[W_9], m, H, F + V, st))) return rc;
  if ((rc = bias_grad(ga, H, grads[B_9], m, H, st))) return rc;
  // g8[:, 1:1+F] = g9 . W_9[:, :F] (no relu on feat);  g8[:,0] = g_sigma * (sigma_pre > 0)   -> gb (ld 1+F)
  if ((rc = linear_dgrad(ga, H, params[W_9], F + V, gb + 1, ldg, nullptr, 0, m, H, F, st))) return rc;
  sigma_bwd_kernel<<<(unsigned)ceil_div64(m, T), T, 0, st>>>(g_sigma_dev, c + L.z, 1 + F + V, m, gb, ldg);
  NERF_LAUNCH_CHECK();
  // fc_8
  if ((rc = linear_wgrad(gb, ldg, c + L.h7, F, grads[W_8], m, 1 + F, F, st))) return rc;
  if ((rc = bias_grad(gb, ldg, grads[B_8], m, 1 + F, st))) return rc;
  if ((rc = linear_dgrad(gb, ldg, params[W_8], F, ga, F, c + L.h7, F, m, 1 + F, F, st))) return rc;   // g7 in ga
  // fc_7
  if ((rc = linear_wgrad(ga, F, c + L.h6, F, grads[W_7], m, F, F, st))) return rc;
  if ((rc = bias_grad(ga, F, grads[B_7], m, F, st))) return rc;
  if ((rc = linear_dgrad(ga, F, params[W_7], F, gb, F, c + L.h6, F, m, F, F, st))) return rc;         // g6 in gb
  // fc_6
  if ((rc = linear_wgrad(gb, F, c + L.h5, F, grads[W_6], m, F, F, st))) return rc;
  if ((rc = bias_grad(gb, F, grads[B_6], m, F, st))) return rc;
  if ((rc = linear_dgrad(gb, F, params[W_6], F, ga, F, c + L.h5, F, m, F, F, st))) return rc;         // g5 in ga
  // fc_5: input x5 = [pos | h4]; only the h4 columns carry gradient further
  if ((rc = linear_wgrad(ga, F, c + L.x5, P + F, grads[W_5], m, F, P + F, st))) return rc;
  if ((rc = bias_grad(ga, F, grads[B_5], m, F, st))) return rc;
  if ((rc = linear_dgrad(ga, F, params[W_5] + P, P + F, gb, F, c + L.x5 + P, P + F, m, F, F, st))) return rc;  // g4 in gb
  // fc_4 .. fc_1
  float* cur = gb;
  float* nxt = ga;
  for (int i = 4; i >= 1; --i) {
    const float* xin = c + L.h[i - 1];
    if ((rc = linear_wgrad(cur, F, xin, F, grads[W_IN + 2 * i], m, F, F, st))) return rc;
    if ((rc = bias_grad(cur, F, grads[B_IN + 2 * i], m, F, st))) return rc;
    if ((rc = linear_dgrad(cur, F, params[W_IN + 2 * i], F, nxt, F, xin, F, m, F, F, st))) return rc;
    float* t = cur; cur = nxt; nxt = t;
  }
  // fc_in: input pos = x5[:, :P]
  if ((rc = linear_wgrad(cur, F, c + L.x5, P + F, grads[W_IN], m, F, P, st))) return rc;
  if ((rc = bias_grad(cur, F, grads[B_IN], m, F, st))) return rc;
  return NERF_OK;
}

}  // extern "C"
