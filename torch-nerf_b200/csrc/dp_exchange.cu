// Data-parallel gradient exchange fused with the optimizer, over NVLink peer memory (the reference
// is single-device, runner_utils.py:431-453, so this has no reference counterpart beyond torch.optim.Adam's math).
//
// One kernel per step and rank replaces "NCCL all-reduce of the 4.77 MB flat gradient buffer, then Adam":
//
//   A  entry barrier       every rank tells every peer "my gradients are complete" (one flag word per (sender, receiver),
//                          release/acquire at system scope through peer-mapped flag pads)
//   B  reduce-scatter      rank k sums slice k of ALL ranks' gradient buffers with P2P loads over NVLink (fixed rank order:
//      + all-gather        every element is reduced by exactly one rank, so all replicas receive bit-identical sums) and
//                          stores the sum into slice k of every rank's buffer with P2P stores
//   C  barrier             "slice k has landed everywhere"
//   D  Adam                every rank updates ALL parameters from its now fully reduced local buffer (the replicas keep a
//                          full optimizer state, so checkpoints and the single-GPU path are unchanged)
//
// Per rank and step the links carry (W-1)/W x 4.77 MB in and out -- what a ring / NVLS all-reduce moves -- but there is
// one launch, no separate reduction kernel and no host round trip between collective and optimizer.  All blocks are
// resident (grid <= SM count, one small block per SM) because they spin on flags; every spin is bounded and traps.
#include "common.cuh"

namespace nerf {

constexpr int kDpThreads = 512;
constexpr int kDpMaxWorld = 16;

struct DpArgs {
  float* grad[kDpMaxWorld];       // rank r's flat gradient buffer as mapped into THIS process (peer memory for r != rank)
  uint32_t* flags[kDpMaxWorld];   // rank r's flag pad (2 * world words): [sender] entry flags, [world + sender] phase-B flags
  int rank, world;
  float* param;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t n;
  float grad_scale, beta1, beta2, step_size, inv_sqrt_bc2, eps;
  uint32_t seq;                   // 1, 2, 3 ... : the value the flags take in this step
  uint32_t* counter;              // local: blocks that finished phase B
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// loads / stores that bypass L1 (the data is written by other GPUs)
__device__ __forceinline__ float4 ld_cg4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_cg4(float* p, const float4& v) {
  asm volatile("st.global.cg.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// waits until flag word `p` has reached `seq` (flags only grow); bounded (~70 s at 1.9 GHz: ranks of a real training job
// can be seconds apart, e.g. one of them loading data), then trap: a peer that never arrives must not hang the GPU
__device__ __forceinline__ void wait_flag(const uint32_t* p, uint32_t seq) {
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(p) - seq) < 0) {
    if (clock64() - t0 > (1ll << 37)) {
      printf("nerf_b200: data-parallel exchange timed out waiting for a peer (block %d, flag %p, want %u)\n", blockIdx.x,
             (const void*)p, seq);
      __trap();
    }
    __nanosleep(64);
  }
}

__global__ void __launch_bounds__(kDpThreads, 1) dp_exchange_adam_kernel(DpArgs a) {
  const int W = a.world, R = a.rank;
  __shared__ int s_last;
  // ---- A: entry barrier.  The previous kernels of this stream (both wgrads) are complete, so this rank's gradients are
  //      final; block 0 publishes that to every peer, every block waits for all peers.
  if (blockIdx.x == 0 && threadIdx.x < W) {
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + R, a.seq);
  }
  if (threadIdx.x < W) wait_flag(a.flags[R] + threadIdx.x, a.seq);
  __syncthreads();
  // ---- B: slice R of every buffer -> sum in rank order -> slice R of every buffer
  const int64_t quads = (a.n + 3) / 4;  // the buffers are padded to a multiple of 4 floats by the caller
  const int64_t q0 = quads * R / W, q1 = quads * (R + 1) / W;
  for (int64_t q = q0 + (int64_t)blockIdx.x * kDpThreads + threadIdx.x; q < q1; q += (int64_t)gridDim.x * kDpThreads) {
    float4 v[kDpMaxWorld];
#pragma unroll
    for (int r = 0; r < kDpMaxWorld; ++r)
      if (r < W) v[r] = ld_cg4(a.grad[r] + 4 * q);  // all loads in flight before the first add
    float4 s = v[0];
#pragma unroll
    for (int r = 1; r < kDpMaxWorld; ++r)
      if (r < W) s.x += v[r].x, s.y += v[r].y, s.z += v[r].z, s.w += v[r].w;
#pragma unroll
    for (int r = 0; r < kDpMaxWorld; ++r)
      if (r < W) st_cg4(a.grad[r] + 4 * q, s);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(a.counter, 1u);
    s_last = (prev == gridDim.x - 1);
    if (s_last) *a.counter = 0;  // ready for the next step (nobody touches it again in this launch)
  }
  __syncthreads();
  if (s_last && threadIdx.x < W) {
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + W + R, a.seq);
  }
  // ---- C: every rank's slice has landed in this rank's buffer
  if (threadIdx.x < W) wait_flag(a.flags[R] + W + threadIdx.x, a.seq);
  __syncthreads();
  // ---- D: Adam over all parameters (same arithmetic as adam_kernel, csrc/optim.cu); param == null: exchange only
  if (a.param == nullptr) return;
  for (int64_t q = (int64_t)blockIdx.x * kDpThreads + threadIdx.x; q < quads; q += (int64_t)gridDim.x * kDpThreads) {
    const int64_t i4 = 4 * q;
    const float4 gv = ld_cg4(a.grad[R] + i4);
    const float gs[4] = {gv.x * a.grad_scale, gv.y * a.grad_scale, gv.z * a.grad_scale, gv.w * a.grad_scale};
    if (i4 + 3 < a.n) {
      float4 pv = *reinterpret_cast<float4*>(a.param + i4), mv = *reinterpret_cast<float4*>(a.exp_avg + i4),
             vv = *reinterpret_cast<float4*>(a.exp_avg_sq + i4);
      float* pe = &pv.x;
      float* me = &mv.x;
      float* ve = &vv.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        me[k] = me[k] + (1.f - a.beta1) * (gs[k] - me[k]);
        ve[k] = a.beta2 * ve[k] + (1.f - a.beta2) * gs[k] * gs[k];
        const float denom = sqrtf(ve[k]) * a.inv_sqrt_bc2 + a.eps;
        pe[k] -= a.step_size * (me[k] / denom);
      }
      *reinterpret_cast<float4*>(a.param + i4) = pv;
      *reinterpret_cast<float4*>(a.exp_avg + i4) = mv;
      *reinterpret_cast<float4*>(a.exp_avg_sq + i4) = vv;
    } else {
      for (int k = 0; i4 + k < a.n; ++k) {
        const float mm = a.exp_avg[i4 + k] + (1.f - a.beta1) * (gs[k] - a.exp_avg[i4 + k]);
        const float vv = a.beta2 * a.exp_avg_sq[i4 + k] + (1.f - a.beta2) * gs[k] * gs[k];
        a.exp_avg[i4 + k] = mm;
        a.exp_avg_sq[i4 + k] = vv;
        a.param[i4 + k] -= a.step_size * (mm / (sqrtf(vv) * a.inv_sqrt_bc2 + a.eps));
      }
    }
  }
}

}  // namespace nerf

extern "C" int nerf_dp_exchange_adam(float* const* grad_ptrs, uint32_t* const* flag_ptrs, int rank, int world, float* param_dev,
                                     float* exp_avg_dev, float* exp_avg_sq_dev, int64_t n, double lr, double beta1, double beta2,
                                     double eps, int64_t step, double grad_scale, uint32_t seq, uint32_t* counter_dev,
                                     nerf_stream_t stream) {
  using namespace nerf;
  NERF_CHECK_ARG(world >= 1 && world <= kDpMaxWorld && rank >= 0 && rank < world, "nerf_dp_exchange_adam: bad rank / world");
  NERF_CHECK_ARG(n > 0 && step >= 1 && seq >= 1, "nerf_dp_exchange_adam: n, step and seq must be positive");
  NERF_CHECK_ARG(grad_ptrs && flag_ptrs && counter_dev, "nerf_dp_exchange_adam: null pointer");
  NERF_CHECK_ARG(param_dev == nullptr || (exp_avg_dev && exp_avg_sq_dev), "nerf_dp_exchange_adam: null optimizer state");
  DpArgs a;
  for (int r = 0; r < world; ++r) {
    NERF_CHECK_ARG(grad_ptrs[r] && flag_ptrs[r], "nerf_dp_exchange_adam: null peer pointer");
    NERF_CHECK_ARG(reinterpret_cast<uintptr_t>(grad_ptrs[r]) % 16 == 0, "nerf_dp_exchange_adam: gradient buffers must be 16-byte aligned");
    a.grad[r] = grad_ptrs[r];
    a.flags[r] = flag_ptrs[r];
  }
  NERF_CHECK_ARG((reinterpret_cast<uintptr_t>(param_dev) | reinterpret_cast<uintptr_t>(exp_avg_dev) |
                  reinterpret_cast<uintptr_t>(exp_avg_sq_dev)) % 16 == 0,
                 "nerf_dp_exchange_adam: buffers must be 16-byte aligned");
  a.rank = rank, a.world = world;
  a.param = param_dev, a.exp_avg = exp_avg_dev, a.exp_avg_sq = exp_avg_sq_dev, a.n = n;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  a.grad_scale = (float)grad_scale, a.beta1 = (float)beta1, a.beta2 = (float)beta2;
  a.step_size = (float)(lr / bc1), a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2)), a.eps = (float)eps;
  a.seq = seq, a.counter = counter_dev;
  const int64_t quads = (n + 3) / 4;
  int64_t blocks = ceil_div64(quads, kDpThreads);
  if (blocks > sm_count()) blocks = sm_count();  // every block must be resident: they wait on each other's flags
  dp_exchange_adam_kernel<<<(unsigned)blocks, kDpThreads, 0, as_stream(stream)>>>(a);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}
