// Self test of the tensor-core building blocks: one tcgen05 tile with K-major and with MN-major operands, checked
// against a plain matmul by tests/test_gpu_tensorcore.py (validates descriptors, swizzle and the TMEM layout).
#include "common.cuh"
#include "tc_common.cuh"

namespace nerf {
using namespace tc;

// ------------------------------------------------------------------------------------------------
// self test: one UMMA tile, both operand majors
// ------------------------------------------------------------------------------------------------
// variant 0: a (128 x k) row-major, b (n x k) row-major        -> K-major images
// variant 1: a (k x 128) row-major (= A^T), b (k x n) row-major -> MN-major images (the wgrad form)
// variant 2: like 0, but the A operand is first written to TMEM with tcgen05.st (packed bf16 pairs, row = lane)
//            and the MMAs use the TS form (A from tensor memory) -- the chain kernels' operand path
__global__ void __launch_bounds__(128, 1) selftest_umma_kernel(const uint16_t* __restrict__ a,
                                                               const uint16_t* __restrict__ b, float* __restrict__ d,
                                                               int n, int k, int variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = k / 64;
  const int reps = (variant >> 4) + 1;  // variant 3 only: the whole product is accumulated `reps` times
  variant &= 15;
  const bool split = variant == 3;  // two issuing threads accumulate into one accumulator
  if (split) variant = 2;
  uint8_t* sA;
  uint8_t* sB;
  uint32_t a_blk, b_blk;  // byte stride between 64-column blocks
  if (variant == 0 || variant == 2 || variant == 3) {
    a_blk = 128 * 128;
    b_blk = n * 128;
    sA = smem;
    sB = smem + nkb * a_blk;
    for (int e = threadIdx.x; e < 128 * k; e += 128) {
      int r = e / k, c = e % k;
      *reinterpret_cast<uint16_t*>(sA + (c / 64) * a_blk + tile_off(r, c % 64)) = a[e];
    }
    for (int e = threadIdx.x; e < n * k; e += 128) {
      int r = e / k, c = e % k;
      *reinterpret_cast<uint16_t*>(sB + (c / 64) * b_blk + tile_off(r, c % 64)) = b[e];
    }
  } else {
    a_blk = k * 128;  // block = k rows x 64 M-columns
    b_blk = k * 128;
    sA = smem;
    sB = smem + 2 * a_blk;
    for (int e = threadIdx.x; e < k * 128; e += 128) {
      int r = e / 128, c = e % 128;  // r = K index, c = M index
      *reinterpret_cast<uint16_t*>(sA + (c / 64) * a_blk + tile_off(r, c % 64)) = a[e];
    }
    for (int e = threadIdx.x; e < k * n; e += 128) {
      int r = e / n, c = e % n;
      *reinterpret_cast<uint16_t*>(sB + (c / 64) * b_blk + tile_off(r, c % 64)) = b[e];
    }
  }
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bar, split ? 2 : 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (split) {  // the accumulator starts at zero: every MMA accumulates, so the issue order of the two threads is free
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0u;
    for (int c = 0; c < n / 32; ++c) tmem_st32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, z);
    tmem_st_wait();
  }
  if (variant == 2) {
    // row = TMEM lane; 64 bf16 (one k-block) = 32 packed words = 32 TMEM columns at [256 + 32*kb, +32)
    const int row = warp * 32 + lane;
    for (int kb = 0; kb < nkb; ++kb) {
      uint32_t w[32];
      for (int c = 0; c < 32; ++c) {
        const uint32_t lo = a[(size_t)row * k + kb * 64 + 2 * c], hi = a[(size_t)row * k + kb * 64 + 2 * c + 1];
        w[c] = lo | (hi << 16);
      }
      tmem_st32(tmem_base + ((uint32_t)(warp * 32) << 16) + 256 + kb * 32, w);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (split) {
    if ((warp == 0 || warp == 1) && lane == 0) {
      const uint32_t idesc = make_idesc_bf16((uint32_t)n, false, false);
      for (int rep = 0; rep < reps; ++rep) {
        for (int kk = warp; kk < k / 16; kk += 2) {
          const uint64_t db = desc_kmajor(smem_u32(sB) + (kk / 4) * b_blk + (kk % 4) * 32);
          umma_bf16_ts(tmem_base, tmem_base + 256 + 8 * kk, db, idesc, 1u);
        }
      }
      umma_commit(&bar);
    }
  } else if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16((uint32_t)n, variant == 1, variant == 1);
    for (int kk = 0; kk < k / 16; ++kk) {
      uint64_t da, db;
      if (variant == 2) {
        db = desc_kmajor(smem_u32(sB) + (kk / 4) * b_blk + (kk % 4) * 32);
        umma_bf16_ts(tmem_base, tmem_base + 256 + 8 * kk, db, idesc, kk > 0 ? 1u : 0u);
        continue;
      }
      if (variant == 0) {
        da = desc_kmajor(smem_u32(sA) + (kk / 4) * a_blk + (kk % 4) * 32);
        db = desc_kmajor(smem_u32(sB) + (kk / 4) * b_blk + (kk % 4) * 32);
      } else {
        da = desc_mnmajor(smem_u32(sA) + kk * 2048, a_blk);
        db = desc_mnmajor(smem_u32(sB) + kk * 2048, b_blk);
      }
      umma_bf16(tmem_base, da, db, idesc, kk > 0 ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < n / 32; ++c) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) d[(size_t)row * n + c * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}


// ------------------------------------------------------------------------------------------------
// self test + rate probe of the CTA-pair MMA (cta_group::2, M = 256): a (256 x k), b (n x k), d (256 x n) row-major.
// CTA r of the pair holds rows [128 r, +128) of A (shared memory, or tensor memory when ts != 0) and of D, and rows
// [n/2 r, +n/2) of B.  With iters > 0 the leader afterwards times iters x (k/16) MMAs: cycles[pair] = SM cycles.
// ------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    selftest_umma2_kernel(const uint16_t* __restrict__ a, const uint16_t* __restrict__ b, float* __restrict__ d, int n, int k,
                          int ts_mode, int iters, unsigned long long* __restrict__ cycles) {
  const int ts = ts_mode & 1, both = (ts_mode >> 1) & 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int nkb = k / 64, nh = n / 2;
  const uint32_t a_blk = 128 * 128, b_blk = (uint32_t)nh * 128;
  uint8_t* sA = smem;
  uint8_t* sB = smem + nkb * a_blk;
  for (int e = threadIdx.x; e < 128 * k; e += 128) {
    int r = e / k, c = e % k;
    *reinterpret_cast<uint16_t*>(sA + (c / 64) * a_blk + tile_off(r, c % 64)) = a[(size_t)(rank * 128 + r) * k + c];
  }
  for (int e = threadIdx.x; e < nh * k; e += 128) {
    int r = e / k, c = e % k;
    *reinterpret_cast<uint16_t*>(sB + (c / 64) * b_blk + tile_off(r, c % 64)) = b[(size_t)(rank * nh + r) * k + c];
  }
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], both ? 2 : 1);  // one multicast commit per issuing CTA
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc2(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (ts) {
    const int row = warp * 32 + lane;
    for (int kb = 0; kb < nkb; ++kb) {
      uint32_t w[32];
      for (int c = 0; c < 32; ++c) {
        const size_t off = (size_t)(rank * 128 + row) * k + kb * 64 + 2 * c;
        w[c] = (uint32_t)a[off] | ((uint32_t)a[off + 1] << 16);
      }
      tmem_st32(tmem_base + ((uint32_t)(warp * 32) << 16) + 256 + kb * 32, w);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
  }
  const uint32_t idesc = make_idesc_bf16_m256((uint32_t)n);
  auto issue_all = [&](uint32_t acc) {
    for (int kk = 0; kk < k / 16; ++kk) {
      const uint64_t db = desc_kmajor(smem_u32(sB) + (kk / 4) * b_blk + (kk % 4) * 32);
      if (ts) {
        umma2_bf16_ts(acc, tmem_base + 256 + 8 * kk, db, idesc, kk > 0 ? 1u : 0u);
      } else {
        const uint64_t da = desc_kmajor(smem_u32(sA) + (kk / 4) * a_blk + (kk % 4) * 32);
        umma2_bf16(acc, da, db, idesc, kk > 0 ? 1u : 0u);
      }
    }
  };
  if (rank == 0 && threadIdx.x == 0) {
    issue_all(tmem_base);
    umma2_commit(&bar[0], 3);
  }
  mbar_wait(&bar[0], 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < n / 32; ++c) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) d[(size_t)(rank * 128 + row) * n + c * 32 + i] = __uint_as_float(v[i]);
  }
  if (iters > 0) {
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    // both != 0: BOTH CTAs of the pair issue pair MMAs, each into its own accumulator (columns 128 * rank; n <= 128):
    // every instruction costs its issuing thread the same ~100 cycles but feeds both SMs' tensor pipes
    if (threadIdx.x == 0 && (rank == 0 || both)) {
      const uint32_t acc = tmem_base + (both ? 128u * rank : 0u);
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) issue_all(acc);
      umma2_commit(&bar[1], 3);
      mbar_wait(&bar[1], 0);
      cycles[both ? blockIdx.x : blockIdx.x / 2] = (unsigned long long)(clock64() - t0);
    } else {
      mbar_wait(&bar[1], 0);
    }
    if (both) {  // the accumulator the SECOND CTA issued into goes back to d: the caller checks it like the first
      tc_fence_after();
      for (int c = 0; c < n / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + 128 + c * 32, v);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) d[(size_t)(rank * 128 + row) * n + c * 32 + i] = __uint_as_float(v[i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// micro-benchmark: sustained tcgen05.mma issue rate of one CTA per SM (M=128, K=16 per instruction)
//   mode 0: A and B from shared memory (SS), N = n;   mode 1: A from TMEM (TS), B from shared memory
// out[block] = cycles for `iters` groups of 4 MMAs (one 64-wide k-block), measured by the issuing thread.
// ------------------------------------------------------------------------------------------------
// bg_warps extra warps stream tcgen05.ld (32x32b.x32 of the accumulator region) or tcgen05.st concurrently, to
// measure how epilogue traffic and the MMA pipe share tensor memory.  out[block] = MMA-thread cycles,
// out[gridDim.x + block] = cycles of the first background warp for its `bg_iters` accesses.
__global__ void __launch_bounds__(640, 1) mma_rate_kernel(int iters, int n, int mode, int bg_warps, int bg_iters, int bg_store,
                                                           unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[2];  // [0]: target of per-group commits, [1]: a completed barrier to probe
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&bar2[0], 1);
    mbar_init(&bar2[1], 1);
    fence_barrier_init();
    mbar_arrive(&bar2[1]);  // phase 0 of bar2[1] is complete: waiting on parity 0 returns at once
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if ((warp == 0 || (warp == 1 && (mode & 32))) && iters > 0) {
    // issue pattern of the chain kernels: the whole warp runs the loop, one elected lane issues.  mode bit 5: warp 1
    // is a second issuer working on its own accumulator (columns [128,256)) and its own completion barrier
    const bool leader = elect_one();
    // mode bit 8: M = 64 instructions (same descriptor with the M field halved; the data is not checked here)
    const uint32_t idesc = (mode & 256) ? ((make_idesc_bf16((uint32_t)n, false, false) & ~(0x1fu << 24)) | ((64u >> 4) << 24))
                                        : make_idesc_bf16((uint32_t)n, false, false);
    const uint64_t da = desc_kmajor(smem_u32(smem));
    const uint64_t db = desc_kmajor(smem_u32(smem) + 16384);
    uint64_t* done = warp == 0 ? &bar : &bar2[0];
    const uint32_t acc0 = tmem_base + ((warp == 1 && !(mode & 128)) ? 128u : 0u);  // mode bit 7: both issuers share one accumulator
    const long long t0 = clock64();
    const int m = mode & 1;
    for (int it = 0; it < iters; ++it) {
      if (mode & 4) {  // probe an already-completed barrier before every group (mode bit 6: with test_wait)
        if (mode & 64) mbar_wait_spin(&bar2[1], 0);
        else mbar_wait(&bar2[1], 0);
      }
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // mode bit 4: successive MMAs alternate between two independent accumulators (columns 0 and 128; n <= 128)
          const uint32_t acc = acc0 + (((mode & 16) && (k & 1)) ? 128u : 0u);
          if (m == 0) umma_bf16(acc, da + 2 * k, db + 2 * k, idesc, 1u);
          else umma_bf16_ts(acc, tmem_base + 256 + 8 * k, db + 2 * k, idesc, 1u);
        }
        if ((mode & 2) && warp == 0) umma_commit(&bar2[0]);  // commit after every group (nobody waits on it)
      }
      if (mode & 8) __syncwarp();
    }
    if (leader) umma_commit(done);
    __syncwarp();
    mbar_wait(done, 0);
    const long long t1 = clock64();
    if (lane == 0 && warp == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  if (warp >= 4 && warp < 4 + bg_warps) {
    // background tensor-memory traffic on columns [384,512) (not touched by the MMAs)
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 384;
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = lane + i;
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < bg_iters; ++it) {
      if (bg_store) {
        tmem_st32(taddr + (it & 3) * 32, v);
        tmem_st_wait();
      } else {
        tmem_ld32(taddr + (it & 3) * 32, v);
        tmem_ld_wait();
        acc ^= v[it & 31];
      }
    }
    const long long t1 = clock64();
    if (warp == 4 && lane == 0) out[gridDim.x + blockIdx.x] = (unsigned long long)(t1 - t0) + (acc & 0u);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// Write-bandwidth probe: every CTA streams 4 KB pieces to a contiguous range, round-robin over the grid.
//   mode 0: st.global.v4      1: st.global.cs.v4 (streaming)      2: bulk store (TMA engine) from shared memory
//   mode 3: bulk store with an L2 evict_first cache hint          4: bulk store with an L2 evict_last hint
__global__ void __launch_bounds__(256) write_bw_kernel(uint8_t* dst, size_t bytes, int mode) {
  __shared__ __align__(1024) uint8_t buf[4][4096];
  const size_t pieces = bytes / 4096;
  for (int i = threadIdx.x; i < 4 * 4096 / 16; i += blockDim.x) reinterpret_cast<uint4*>(&buf[0][0])[i] = make_uint4(i, 1, 2, 3);
  fence_proxy_async();
  __syncthreads();
  if (mode <= 1) {
    const uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
    for (size_t p = blockIdx.x; p < pieces; p += gridDim.x) {
      uint4* d = reinterpret_cast<uint4*>(dst + p * 4096) + threadIdx.x;
      if (mode == 0) *d = v;
      else asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(d), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
  } else if (threadIdx.x == 0) {
    uint64_t policy = 0;
    if (mode == 3) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    if (mode == 4) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    int k = 0;
    for (size_t p = blockIdx.x; p < pieces; p += gridDim.x, ++k) {
      if (mode == 2)
        bulk_s2g(dst + p * 4096, buf[k & 3], 4096);
      else
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst + p * 4096),
                     "r"(smem_u32(buf[k & 3])), "r"(4096), "l"(policy)
                     : "memory");
      bulk_commit();
      bulk_wait_read<3>();
    }
    bulk_wait_all<0>();
  }
}

}  // namespace nerf

using namespace nerf;

extern "C" {

int nerf_selftest_write_bw(void* dst_dev, size_t bytes, int mode, int blocks, nerf_stream_t stream) {
  NERF_CHECK_ARG(dst_dev && bytes >= 4096 && mode >= 0 && mode <= 4 && blocks > 0, "nerf_selftest_write_bw: bad arguments");
  write_bw_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<uint8_t*>(dst_dev), bytes, mode);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_selftest_umma(const uint16_t* a_dev, const uint16_t* b_dev, float* d_dev, int n, int k, int variant,
                       nerf_stream_t stream) {
  NERF_CHECK_ARG(a_dev && b_dev && d_dev, "nerf_selftest_umma: null pointer");
  NERF_CHECK_ARG(n >= 64 && n <= 256 && n % 64 == 0 && k >= 64 && k <= 256 && k % 64 == 0 && (variant >= 0 && (variant & 15) <= 3 && (variant < 16 || (variant & 15) == 3)),
                 "nerf_selftest_umma: n, k must be multiples of 64 in [64,256]; variant 0|1|2|3");
  size_t smem = (size_t)128 * k * 2 + (size_t)n * k * 2 + 1024;
  NERF_CUDA(cudaFuncSetAttribute(selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  selftest_umma_kernel<<<1, 128, smem, as_stream(stream)>>>(a_dev, b_dev, d_dev, n, k, variant);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_selftest_umma2(const uint16_t* a_dev, const uint16_t* b_dev, float* d_dev, int n, int k, int ts, int pairs, int iters,
                        unsigned long long* cycles_dev, nerf_stream_t stream) {
  NERF_CHECK_ARG(a_dev && b_dev && d_dev, "nerf_selftest_umma2: null pointer");
  NERF_CHECK_ARG(n >= 32 && n <= 256 && n % 32 == 0 && k >= 64 && k <= 256 && k % 64 == 0 && pairs >= 1 && iters >= 0 &&
                     (iters == 0 || cycles_dev) && ts >= 0 && ts <= 3 && (!(ts & 2) || (n <= 128 && iters > 0)),
                 "nerf_selftest_umma2: n multiple of 32 in [32,256], k multiple of 64 in [64,256]");
  size_t smem = (size_t)128 * k * 2 + (size_t)(n / 2) * k * 2 + 1024;
  NERF_CUDA(cudaFuncSetAttribute(selftest_umma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  selftest_umma2_kernel<<<2 * pairs, 128, smem, as_stream(stream)>>>(a_dev, b_dev, d_dev, n, k, ts, iters, cycles_dev);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_selftest_mma_rate(int blocks, int iters, int n, int mode, int bg_warps, int bg_iters, int bg_store,
                           unsigned long long* cycles_dev, nerf_stream_t stream) {
  NERF_CHECK_ARG(cycles_dev && blocks > 0 && iters >= 0 && n >= 16 && n <= 256 && n % 16 == 0 && mode >= 0 && mode < 512 &&
                     bg_warps >= 0 && bg_warps <= 16 && bg_iters >= 0,
                 "nerf_selftest_mma_rate: bad arguments");
  const size_t smem = 16384 + 32768 + 1024;
  NERF_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mma_rate_kernel<<<blocks, 128 + 32 * bg_warps, smem, as_stream(stream)>>>(iters, n, mode, bg_warps, bg_iters, bg_store,
                                                                              cycles_dev);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

}  // extern "C"
