// K4+K5 on the sm_100a tensor cores: the NeRF MLP forward (network/nerf.py:65-121) as a chain of tcgen05 BF16 MMAs
// with fp32 accumulation in TMEM; positional encoding (signal_encoder/positional_encoder.py:49-104, as applied by
// scene/primitives/cube.py:62-69) is computed in-kernel as the first layer's operand; weights are streamed from L2
// by the TMA engine (cp.async.bulk + mbarrier) through a 4-stage ring of 32 KB stages (two chunks of 128 outputs x 64 inputs).
//
// One CTA per SM, 128 sample rows per tile.  Activations live in TENSOR MEMORY between layers:
//
//   TMEM columns [0,256)    fp32 accumulator of the current layer (two N-halves of 128 columns)
//                [256,384)  A operand buffer 0  (128 rows x 256 bf16 as packed pairs)   \ ping-pong: layer l reads
//                [384,512)  A operand buffer 1                                          / buffer l&1, writes (l+1)&1
//
//   warp 0      weight loader   1-D bulk copies of pre-swizzled bf16 weight chunks
//   warps 1,10  MMA issuers     one thread each issues tcgen05.mma (M=128, N=128, K=16) in the TS form: A from TMEM, B from
//                               shared memory -- this takes the activations off the shared-memory port, which limits
//                               the SS form to one 128x256x16 MMA per 168 cycles.  A 256-wide layer is two N-halves
//                               committed separately: while the tensor pipe computes output columns [128,256) the
//                               epilogue already drains [0,128), and the next layer's first MMAs (which only need the
//                               k-blocks produced from half 0) start the moment half 1 is issued.
//   warps 2-9   epilogue        two warps per TMEM lane quarter: tcgen05.ld -> +bias, ReLU -> bf16 pairs -> tcgen05.st
//                               into the other A buffer, signalled per 64-column k-block; the same warps build the
//                               encoded inputs of the next tile (those two k-blocks stay in shared memory, SS form)
//
// In training mode the epilogue additionally writes every layer input as a tile image (through a shared-memory staging
// area and per-warp 4 KB bulk stores) plus ReLU sign-bit masks into the training cache.
//
// Tensor-core layers: fc_in, fc_1..fc_7, fc_8 rows 1..256 (features), fc_9.  The density head (fc_8 row 0,
// nerf.py:115) and fc_out + sigmoid (nerf.py:119) are fp32 dot products in the epilogues of layers 7 and 9,
// taken from the fp32 accumulators, so sigma keeps fp32 accuracy where the 1e8 last interval makes it matter.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "mlp_tc_layout.cuh"
#include "tc_common.cuh"

namespace nerf {
using namespace tc;

constexpr int kStages = 4;
constexpr int kStageBytes = 2 * kChunkBytes;  // two chunks (128 output rows x 64 K-columns each) per barrier: every
                                              // mbarrier probe costs the issuing thread ~125 cycles of dead tensor time
constexpr int kFwdThreads = 352;   // loader, MMA issuer A, 8 epilogue warps, MMA issuer B
constexpr int kMmaWarpB = 10;
constexpr int kEpiThreads = 256;
// shared memory map (bytes from the 1024-aligned base)
constexpr int kSmStage = 0;                     // 8 warps x 4 KB: staging of bf16 activations for the cache (training)
constexpr int kSmIn = 32768;                    // pe block 16 KB | de block 16 KB (A operands of the SS-form chunks)
constexpr int kSmW = 65536;                     // weight ring
constexpr int kSmC = kSmW + kStages * kStageBytes;
constexpr int kSmX = kSmC + kCFloats * 4;       // 128 x 4 floats: partial sigma / rgb exchange between warp groups
constexpr int kSmBar = kSmX + 128 * 16;
constexpr int kSmTotal = kSmBar + 256;
constexpr int kFwdSmemBytes = kSmTotal + 1024;  // + alignment slack
// tensor memory map (columns)
constexpr uint32_t kTmAcc = 0;
constexpr uint32_t kTmA = 256;                  // two A buffers of 128 columns

struct FwdArgs {
  const uint8_t* packed;
  const float* pts;     // (M,3) or null
  const float* dirs;    // (M,3) or null
  const float* ray_o;   // (N,3)
  const float* ray_d;   // (N,3)
  const float* t;       // (N,S)
  int s;
  int64_t m;
  float* sigma;
  float* rgb;
  uint8_t* cache;       // training cache or null
  unsigned long long* prof;  // optional timeline of CTA 0 (nerf_debug_set_profile_buffer)
  int prof_tiles;
};

// [v | sin(2^l v) | cos(2^l v)]_{l<L} for a 3-vector, written as bf16 into the first NCHUNK 16-byte chunks of a
// swizzled tile-image row.  Higher octaves come from angle doubling (sin 2a = 2 sin a cos a,
// cos 2a = (cos a - sin a)(cos a + sin a)); the accumulated error (~2^l ulp) is far below bf16 resolution.
template <int L, int NCHUNK>
__device__ __forceinline__ void encode_row(float x, float y, float z, uint8_t* row_ptr, int row) {
  float v[NCHUNK * 8];
  float sn[3], cs[3];
  v[0] = x, v[1] = y, v[2] = z;
  sincosf(x, &sn[0], &cs[0]);
  sincosf(y, &sn[1], &cs[1]);
  sincosf(z, &sn[2], &cs[2]);
#pragma unroll
  for (int l = 0; l < L; ++l) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      v[3 + 6 * l + c] = sn[c];
      v[6 + 6 * l + c] = cs[c];
      float s2 = 2.f * sn[c] * cs[c];
      float c2 = (cs[c] - sn[c]) * (cs[c] + sn[c]);
      sn[c] = s2, cs[c] = c2;
    }
  }
#pragma unroll
  for (int i = 3 + 6 * L; i < NCHUNK * 8; ++i) v[i] = 0.f;
#pragma unroll
  for (int j = 0; j < NCHUNK; ++j) {
    uint4 q = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                         pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
    *reinterpret_cast<uint4*>(row_ptr + ((j ^ (row & 7)) << 4)) = q;
  }
}

// 32 accumulator columns -> +bias -> (ReLU).  Returns the sign bits of the pre-activations, element i at bit
// (31 - i): one funnel shift per element (the training cache's ReLU mask, see mlp_tc_layout.cuh).
template <bool RELU>
__device__ __forceinline__ uint32_t finish_group(const uint32_t (&v)[32], const float* __restrict__ bias, float (&f)[32]) {
  uint32_t neg = 0;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 b = *reinterpret_cast<const float4*>(bias + i);
    const float t[4] = {__uint_as_float(v[i]) + b.x, __uint_as_float(v[i + 1]) + b.y, __uint_as_float(v[i + 2]) + b.z,
                        __uint_as_float(v[i + 3]) + b.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (RELU) {
        neg = __funnelshift_l(__float_as_uint(t[e]), neg, 1);
        f[i + e] = fmaxf(t[e], 0.f);
      } else {
        f[i + e] = t[e];
      }
    }
  }
  return neg;
}

// 32 fp32 values -> 16 packed bf16 pairs
__device__ __forceinline__ void pack_group(const float (&f)[32], uint32_t* w) {
#pragma unroll
  for (int j = 0; j < 16; ++j) w[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
}

// 16 packed words (32 columns) -> four 16-byte chunks of a swizzled tile-image row
__device__ __forceinline__ void store_words(const uint32_t* w, uint8_t* blk_row, int row, int chunk0) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(blk_row + (((chunk0 + j) ^ (row & 7)) << 4)) =
        make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
}

template <bool kTrain>
__global__ void __launch_bounds__(kFwdThreads, 1) mlp_fwd_kernel(FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sStage = smem + kSmStage;
  uint8_t* sIn = smem + kSmIn;
  uint8_t* sW = smem + kSmW;
  float* sC = reinterpret_cast<float*>(smem + kSmC);
  float* sX = reinterpret_cast<float*>(smem + kSmX);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmBar);
  uint64_t* full = bars;                   // [kStages]
  uint64_t* empty = bars + kStages;        // [kStages]
  uint64_t* a_ready = bars + 2 * kStages;  // [2 used] k-block pairs {0,1} and {2,3}: one completion per producing layer
  uint64_t* in_ready = a_ready + 4;        // [1]   one completion per tile
  uint64_t* acc_full = in_ready + 1;       // [2 N-halves of the accumulator]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = num_tiles(a.m);

  {
    const float* cg = reinterpret_cast<const float*>(a.packed + kPackedConstOff);
    for (int i = threadIdx.x; i < kCFloats; i += kFwdThreads) sC[i] = __ldg(cg + i);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&a_ready[i], kEpiThreads);
    mbar_init(in_ready, kEpiThreads);
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ weight loader (warp in lock step, one lane issues)
    {
      const bool leader = elect_one();
      uint32_t g = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint8_t* src = a.packed + kPackedFwdOff;
        for (int l = 0; l < kNumFwdLayers; ++l) {
          for (int nh = 0; nh < fwd_nh(l); ++nh) {
            for (int c0 = 0; c0 < fwd_nk(l); c0 += 2) {
              const uint32_t bytes = (uint32_t)min(2, fwd_nk(l) - c0) * kChunkBytes;
              const uint32_t s = g % kStages, ph = (g / kStages) & 1;
              mbar_wait(&empty[s], ph ^ 1);
              if (leader) {
                mbar_arrive_expect_tx(&full[s], bytes);
                bulk_g2s(sW + s * kStageBytes, src, bytes, &full[s]);
              }
              __syncwarp();
              src += bytes;
              ++g;
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp == kMmaWarpB) {
    // ------------------------------------------------------------------ MMA issuers (two warps)
    // tcgen05.mma issue is effectively synchronous (the issuing thread is held while its MMA executes) and every
    // mbarrier probe costs ~125-150 cycles, so a single issuer leaves the tensor pipe idle during each probe.  Warp 1
    // issues N-half 0 of every layer, warp kMmaWarpB N-half 1: while one polls its barriers the other's MMAs run.
    // Each warp runs its loop in lock step and one elected lane issues.
    {
      const int nh = (warp == 1) ? 0 : 1;
      const bool leader = elect_one();
      uint32_t g_layer = 0, a_cnt = 0, in_cnt = 0;
      constexpr uint32_t idesc = make_idesc_bf16(128, false, false);
      const uint32_t sIn_u = smem_u32(sIn), sW_u = smem_u32(sW);
      const uint32_t acc = tmem_base + kTmAcc + (uint32_t)nh * 128u;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // the encoded inputs are ready; all epilogue warps have also left the previous tile (accumulator drained)
        mbar_wait(in_ready, in_cnt & 1);
        ++in_cnt;
        for (int l = 0; l < kNumFwdLayers; ++l) {
          const int nk = fwd_nk(l);
          const uint32_t stages_per_half = (uint32_t)(nk + 1) / 2;
          uint32_t g = g_layer + (nh ? stages_per_half : 0u);  // this warp's first weight stage of the layer
          g_layer += stages_per_half * (uint32_t)fwd_nh(l);
          if (nh >= fwd_nh(l)) continue;                      // fc_9 is a single N-half
          const uint32_t a_tm = tmem_base + kTmA + (uint32_t)(l & 1) * 128u;  // this layer's A operand in TMEM
          const bool stamp = a.prof != nullptr && blockIdx.x == 0 && (int)in_cnt <= a.prof_tiles && nh == 0;
          if (stamp && leader) a.prof[(((int)in_cnt - 1) * kNumFwdLayers + l) * 8 + 0] = clock64();
          long long wait_a = 0, wait_w = 0;
          // a_ready[p] (k-block pair p) completes once per producing layer 0..8; layer l >= 1 consumes round (l - 1).
          // Every epilogue thread arrives on pair 0 before pair 1, so pair 1 complete implies pair 0 complete.
          const uint32_t a_par = (a_cnt + (uint32_t)(l - 1)) & 1;
          if (l >= 1) {
            // N-half 0 overwrites accumulator columns [0,128) and needs k-blocks {0,1}: pair 0.  N-half 1 overwrites
            // [128,256) (drained by the epilogue's second half) and reads all four k-blocks: pair 1.
            const long long w0 = stamp ? clock64() : 0;
            mbar_wait(&a_ready[nh], a_par);
            if (stamp) wait_a += clock64() - w0;
          }
          for (int kb = 0; kb < nk; ++kb) {
            int ab = -1;             // A k-block in TMEM, or -1 for the shared-memory blocks
            uint32_t a_smem = 0;
            int nsteps = 4;
            if (l == 0 || (l == 5 && kb == 0)) {
              a_smem = sIn_u;                      // encoded position
            } else if (l == 9 && kb == 4) {
              a_smem = sIn_u + kBlockBytes;        // encoded view direction (K = 32)
              nsteps = 2;
            } else {
              ab = (l == 5) ? kb - 1 : kb;
              if (nh == 0 && ab == 2) {
                const long long w0 = stamp ? clock64() : 0;
                mbar_wait(&a_ready[1], a_par);
                if (stamp) wait_a += clock64() - w0;
              }
            }
            const uint32_t s = g % kStages, ph = (g / kStages) & 1;
            if ((kb & 1) == 0) {  // first chunk of a weight stage
              const long long w1 = stamp ? clock64() : 0;
              mbar_wait(&full[s], ph);
              if (stamp) wait_w += clock64() - w1;
            }
            tc_fence_after();
            const bool stage_done = (kb & 1) || kb == nk - 1;
            if (leader) {
              const uint64_t db = desc_kmajor(sW_u + s * kStageBytes + (kb & 1) * kChunkBytes);
              if (ab >= 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_ts(acc, a_tm + (uint32_t)(ab * 32 + k * 8), db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              } else {
                const uint64_t da = desc_kmajor(a_smem);
#pragma unroll 4
                for (int k = 0; k < nsteps; ++k) umma_bf16(acc, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              }
              if (stage_done) umma_commit(&empty[s]);
            }
            __syncwarp();
            if (stage_done) ++g;
          }
          if (leader) umma_commit(&acc_full[nh]);
          __syncwarp();
          if (stamp && leader) {
            unsigned long long* pr = a.prof + (((int)in_cnt - 1) * kNumFwdLayers + l) * 8;
            pr[1] = clock64();
            pr[4] = (unsigned long long)wait_a;
            pr[5] = (unsigned long long)wait_w;
          }
        }
        a_cnt += 9;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;   // warp group: drains k-block `half` of N-half 0 and k-block `2 + half` of N-half 1
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t accn[2] = {0, 0};  // completions seen per N-half barrier
    uint8_t* st_slot = sStage + (warp - 2) * 4096;  // this warp's 32 rows x 128 B staging slice
    uint8_t* st_row = st_slot + lane * 128;
    int tile_iter = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_iter) {
      const int64_t grow = tile * kTileM + row;
      const bool stamp = a.prof != nullptr && blockIdx.x == 0 && tile_iter < a.prof_tiles && warp == 2 && lane == 0;
      uint8_t* cache_tile = kTrain ? a.cache + (size_t)tile * kCacheTileBytes : nullptr;
      uint32_t* mask_tile =
          kTrain ? reinterpret_cast<uint32_t*>(a.cache + cache_mask_offset(a.m) + (size_t)tile * kMaskTileBytes) : nullptr;
      // ---- encoded inputs (cube.py:62-69): half 0 encodes the point, half 1 the view direction
      {
        float x = 0.f, y = 0.f, z = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;
        if (grow < a.m) {
          if (a.pts != nullptr) {
            if (half == 0) {
              x = __ldg(a.pts + 3 * grow), y = __ldg(a.pts + 3 * grow + 1), z = __ldg(a.pts + 3 * grow + 2);
            } else {
              dx = __ldg(a.dirs + 3 * grow), dy = __ldg(a.dirs + 3 * grow + 1), dz = __ldg(a.dirs + 3 * grow + 2);
            }
          } else {
            const int64_t ray = grow / a.s;
            dx = __ldg(a.ray_d + 3 * ray), dy = __ldg(a.ray_d + 3 * ray + 1), dz = __ldg(a.ray_d + 3 * ray + 2);
            if (half == 0) {
              const float tt = __ldg(a.t + grow);
              // stratified_sampler.py:126: o + t*d, product and sum rounded separately
              x = __fadd_rn(__ldg(a.ray_o + 3 * ray), __fmul_rn(tt, dx));
              y = __fadd_rn(__ldg(a.ray_o + 3 * ray + 1), __fmul_rn(tt, dy));
              z = __fadd_rn(__ldg(a.ray_o + 3 * ray + 2), __fmul_rn(tt, dz));
            }
          }
        }
        if (kTrain) {
          if (lane == 0) bulk_wait_read<0>();  // the previous tile's stores (incl. the one out of the input block) are read
          __syncwarp();
        }
        if (half == 0) encode_row<10, 8>(x, y, z, sIn + row * 128, row);
        else encode_row<4, 4>(dx, dy, dz, sIn + kBlockBytes + row * 128, row);
        fence_proxy_async();
        if (kTrain) {
          __syncwarp();
          if (lane == 0) {
            const int blk = half == 0 ? kCachePe : kCacheDe;
            bulk_s2g(cache_tile + (size_t)blk * kBlockBytes + q * 4096, sIn + half * kBlockBytes + q * 4096, 4096);
            bulk_commit();
          }
        }
        mbar_arrive(in_ready);
      }
      float sigma_part = 0.f;
      float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
      for (int l = 0; l < kNumFwdLayers; ++l) {
        const uint32_t taddr = lane_addr + kTmAcc;
        const uint32_t a_next = lane_addr + kTmA + (uint32_t)((l + 1) & 1) * 128u;  // next layer's A operand
        const float* bias = sC + ((l < 8) ? kCBias + 256 * l : (l == 8 ? kCBias8 : kCBias9));
        if (l < 9) {
#pragma unroll 1
          for (int t = 0; t < 2; ++t) {
            // N-half t of the layer is complete: this warp drains its 64 columns of it = k-block kb of the next layer
            const int kb = half + 2 * t;
            mbar_wait(&acc_full[t], accn[t] & 1);
            ++accn[t];
            tc_fence_after();
            if (stamp && t == 0) a.prof[((tile_iter * kNumFwdLayers + l) * 8) + 2] = clock64();
            uint32_t v0[32], v1[32];
            tmem_ld32(taddr + kb * 64, v0);
            tmem_ld32(taddr + kb * 64 + 32, v1);
            if (kTrain) {
              if (lane == 0) bulk_wait_read<0>();  // this warp's previous bulk store out of its staging slice has been read
              __syncwarp();
            }
            tmem_ld_wait();
            float f[32];
            uint32_t w[32];
            uint32_t neg;
            if (l == 8) neg = finish_group<false>(v0, bias + kb * 64, f);
            else neg = finish_group<true>(v0, bias + kb * 64, f);
            if (l == 7) {
#pragma unroll
              for (int i = 0; i < 32; ++i) sigma_part = fmaf(f[i], sC[kCW8Row0 + kb * 64 + i], sigma_part);
            }
            if (kTrain && l < 8) mask_tile[(l * 8 + 2 * kb) * kTileM + row] = neg;
            pack_group(f, w);
            if (l == 8) neg = finish_group<false>(v1, bias + kb * 64 + 32, f);
            else neg = finish_group<true>(v1, bias + kb * 64 + 32, f);
            if (l == 7) {
#pragma unroll
              for (int i = 0; i < 32; ++i) sigma_part = fmaf(f[i], sC[kCW8Row0 + kb * 64 + 32 + i], sigma_part);
            }
            if (kTrain && l < 8) mask_tile[(l * 8 + 2 * kb + 1) * kTileM + row] = neg;
            pack_group(f, w + 16);
            tmem_st32(a_next + kb * 32, w);  // 64 bf16 = 32 packed columns of this row
            if (kTrain) {
              store_words(w, st_row, row, 0);
              store_words(w + 16, st_row, row, 4);
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                const int blk = (l < 8 ? cache_h(l) : kCacheFeat) + kb;
                bulk_s2g(cache_tile + (size_t)blk * kBlockBytes + q * 4096, st_slot, 4096);
                bulk_commit();
              }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&a_ready[t]);  // k-block pair t = {2t, 2t+1}
          }
          if (l == 7 && half == 1) sX[row * 4 + 3] = sigma_part;
          if (stamp) a.prof[((tile_iter * kNumFwdLayers + l) * 8) + 3] = clock64();
        } else {
          // fc_9 output (128 columns = one N-half): this warp group owns columns [64*half, 64*half + 64)
          mbar_wait(&acc_full[0], accn[0] & 1);
          ++accn[0];
          tc_fence_after();
          if (stamp) a.prof[((tile_iter * kNumFwdLayers + l) * 8) + 2] = clock64();
          uint32_t v0[32], v1[32];
          tmem_ld32(taddr + half * 64, v0);
          tmem_ld32(taddr + half * 64 + 32, v1);
          if (kTrain) {
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
          }
          tmem_ld_wait();
          float f[32];
          uint32_t w[32];
          uint32_t neg = finish_group<true>(v0, bias + half * 64, f);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            rgb0 = fmaf(f[i], sC[kCWout + half * 64 + i], rgb0);
            rgb1 = fmaf(f[i], sC[kCWout + 128 + half * 64 + i], rgb1);
            rgb2 = fmaf(f[i], sC[kCWout + 256 + half * 64 + i], rgb2);
          }
          if (kTrain) {
            mask_tile[(64 + 2 * half) * kTileM + row] = neg;
            pack_group(f, w);
          }
          neg = finish_group<true>(v1, bias + half * 64 + 32, f);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            rgb0 = fmaf(f[i], sC[kCWout + half * 64 + 32 + i], rgb0);
            rgb1 = fmaf(f[i], sC[kCWout + 128 + half * 64 + 32 + i], rgb1);
            rgb2 = fmaf(f[i], sC[kCWout + 256 + half * 64 + 32 + i], rgb2);
          }
          if (kTrain) {
            mask_tile[(64 + 2 * half + 1) * kTileM + row] = neg;
            pack_group(f, w + 16);
            store_words(w, st_row, row, 0);
            store_words(w + 16, st_row, row, 4);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              bulk_s2g(cache_tile + (size_t)(kCacheH9 + half) * kBlockBytes + q * 4096, st_slot, 4096);
              bulk_commit();
            }
          }
          tc_fence_before();
          if (half == 1) {
            sX[row * 4 + 0] = rgb0;
            sX[row * 4 + 1] = rgb1;
            sX[row * 4 + 2] = rgb2;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
          if (half == 0) {
            const float4 o = *reinterpret_cast<const float4*>(sX + row * 4);
            const float sp = sigma_part + o.w + sC[kCB8_0];
            if (grow < a.m) {
              a.sigma[grow] = fmaxf(sp, 0.f);                                  // nerf.py:115
              a.rgb[3 * grow] = 1.f / (1.f + __expf(-(rgb0 + o.x + sC[kCBout])));    // nerf.py:119
              a.rgb[3 * grow + 1] = 1.f / (1.f + __expf(-(rgb1 + o.y + sC[kCBout + 1])));
              a.rgb[3 * grow + 2] = 1.f / (1.f + __expf(-(rgb2 + o.z + sC[kCBout + 2])));
            }
            if (kTrain) mask_tile[kMaskSigmaWord * kTileM + row] = (grow < a.m && sp > 0.f) ? 1u : 0u;
          }
          if (stamp) a.prof[((tile_iter * kNumFwdLayers + l) * 8) + 3] = clock64();
        }
      }
    }
    if (kTrain) {
      if (lane == 0) bulk_wait_all<0>();
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

static unsigned long long* g_prof_buf = nullptr;
static int g_prof_tiles = 0;

}  // namespace nerf

using namespace nerf;

extern "C" int nerf_debug_set_profile_buffer(unsigned long long* buf_dev, int tiles) {
  g_prof_buf = buf_dev;
  g_prof_tiles = tiles;
  return NERF_OK;
}

extern "C" int nerf_mlp_bf16_forward(const void* packed_dev, const float* pts_dev, const float* dirs_dev,
                                     const float* ray_o_dev, const float* ray_d_dev, const float* t_dev, int s,
                                     int64_t m, float* sigma_dev, float* rgb_dev, void* cache_dev,
                                     nerf_stream_t stream) {
  NERF_CHECK_ARG(m >= 0, "nerf_mlp_bf16_forward: negative row count");
  if (m == 0) return NERF_OK;
  NERF_CHECK_ARG(packed_dev && sigma_dev && rgb_dev, "nerf_mlp_bf16_forward: null pointer");
  NERF_CHECK_ARG((pts_dev && dirs_dev) || (ray_o_dev && ray_d_dev && t_dev && s > 0),
                 "nerf_mlp_bf16_forward: give (pts, dirs) or (ray_o, ray_d, t, s)");
  static bool attr_set = false;
  if (!attr_set) {
    NERF_CUDA(cudaFuncSetAttribute(mlp_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
    NERF_CUDA(cudaFuncSetAttribute(mlp_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
    attr_set = true;
  }
  FwdArgs a;
  a.packed = reinterpret_cast<const uint8_t*>(packed_dev);
  a.pts = pts_dev, a.dirs = dirs_dev, a.ray_o = ray_o_dev, a.ray_d = ray_d_dev, a.t = t_dev, a.s = s, a.m = m;
  a.sigma = sigma_dev, a.rgb = rgb_dev;
  a.cache = reinterpret_cast<uint8_t*>(cache_dev);
  a.prof = g_prof_buf;
  a.prof_tiles = g_prof_tiles;
  const int64_t ntiles = num_tiles(m);
  const int grid = (int)((ntiles < sm_count()) ? ntiles : sm_count());
  if (cache_dev)
    mlp_fwd_kernel<true><<<grid, kFwdThreads, kFwdSmemBytes, as_stream(stream)>>>(a);
  else
    mlp_fwd_kernel<false><<<grid, kFwdThreads, kFwdSmemBytes, as_stream(stream)>>>(a);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}
