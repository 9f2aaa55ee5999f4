// K4+K5 on the sm_100a tensor cores: the NeRF MLP forward (network/nerf.py:65-121) as a chain of tcgen05 BF16 MMAs
// with fp32 accumulation in TMEM; positional encoding (signal_encoder/positional_encoder.py:49-104, as applied by
// scene/primitives/cube.py:62-69) is computed in-kernel as the first layer's operand; weights are streamed from L2
// by the TMA engine (cp.async.bulk + mbarrier) through a 3-stage ring.
//
// One CTA per SM, 128 sample rows per tile, activations never leave the SM (inference).  In training mode the
// kernel additionally writes every layer input as a tile image plus ReLU bit masks into the training cache.
//
//   warp 0      weight loader   1-D bulk copies of pre-swizzled bf16 weight chunks (N x 64)
//   warp 1      MMA issuer      one thread issues tcgen05.mma (M=128, N=256|128, K=16); accumulators ping-pong
//                               between TMEM columns [0,256) and [256,512) from layer to layer
//   warps 2-9   epilogue        two warps per TMEM lane quarter (column halves): tcgen05.ld -> +bias, ReLU -> bf16
//                               -> swizzled smem = next layer's A operand, signalled per 64-column k-block so the
//                               next layer's MMAs start while the rest of the accumulator is still being drained;
//                               the same warps build the encoded inputs of the next tile
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

// i-th weight chunk (= k-block of the layer's K dimension) consumed by layer l.  4-block layers follow kb_order;
// the 5-block layers put the block that does not depend on the previous epilogue first (fc_5: encoded position,
// chunk 0) or keep it last (fc_9: encoded view direction, chunk 4).
__host__ __device__ constexpr int fwd_chunk(int l, int i) {
  return fwd_nk(l) == 4 ? kb_order(i) : (l == 5 ? (i == 0 ? 0 : 1 + kb_order(i - 1)) : (l == 9 ? (i < 4 ? kb_order(i) : 4) : i));
}

constexpr int kStages = 3;
constexpr int kStageBytes = 32768;
constexpr int kFwdThreads = 320;
constexpr int kEpiThreads = 256;
// shared memory map (bytes from the 1024-aligned base)
constexpr int kSmA = 0;                         // 4 k-blocks x 16 KB : current activations (A operand)
constexpr int kSmIn = 65536;                    // pe block 16 KB | de block 16 KB
constexpr int kSmW = 98304;                     // weight ring
constexpr int kSmC = kSmW + kStages * kStageBytes;
constexpr int kSmX = kSmC + kCFloats * 4;       // 128 x 4 floats: partial sigma / rgb exchange between column halves
constexpr int kSmBar = kSmX + 128 * 16;
constexpr int kSmTotal = kSmBar + 256;
constexpr int kFwdSmemBytes = kSmTotal + 1024;  // + alignment slack

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

// 32 accumulator columns -> +bias -> (ReLU) -> bf16 -> four 16-byte chunks of the A operand row
template <bool RELU>
__device__ __forceinline__ void finish_group(const uint32_t (&v)[32], const float* __restrict__ bias, float (&f)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 b = *reinterpret_cast<const float4*>(bias + i);
    float t0 = __uint_as_float(v[i]) + b.x, t1 = __uint_as_float(v[i + 1]) + b.y;
    float t2 = __uint_as_float(v[i + 2]) + b.z, t3 = __uint_as_float(v[i + 3]) + b.w;
    if (RELU) {
      t0 = fmaxf(t0, 0.f), t1 = fmaxf(t1, 0.f), t2 = fmaxf(t2, 0.f), t3 = fmaxf(t3, 0.f);
    }
    f[i] = t0, f[i + 1] = t1, f[i + 2] = t2, f[i + 3] = t3;
  }
}

__device__ __forceinline__ uint32_t relu_mask(const float (&f)[32]) {
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) m |= (f[i] > 0.f ? 1u : 0u) << i;
  return m;
}

__device__ __forceinline__ void store_group(const float (&f)[32], uint8_t* blk_row, int row, int chunk0) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 qv = make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                          pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
    *reinterpret_cast<uint4*>(blk_row + (((chunk0 + j) ^ (row & 7)) << 4)) = qv;
  }
}

template <bool kTrain>
__global__ void __launch_bounds__(kFwdThreads, 1) mlp_fwd_kernel(FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem + kSmA;
  uint8_t* sIn = smem + kSmIn;
  uint8_t* sW = smem + kSmW;
  float* sC = reinterpret_cast<float*>(smem + kSmC);
  float* sX = reinterpret_cast<float*>(smem + kSmX);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmBar);
  uint64_t* full = bars;                   // [kStages]
  uint64_t* empty = bars + kStages;        // [kStages]
  uint64_t* a_ready = bars + 2 * kStages;  // [4]   one completion per producing layer
  uint64_t* in_ready = a_ready + 4;        // [1]   one completion per tile
  uint64_t* acc_full = in_ready + 1;       // [2]
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
    for (int i = 0; i < 4; ++i) mbar_init(&a_ready[i], 128);
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
    // ------------------------------------------------------------------ weight loader
    if (lane == 0) {
      uint32_t g = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint8_t* layer_src = a.packed + kPackedFwdOff;
        for (int l = 0; l < kNumFwdLayers; ++l) {
          const uint32_t bytes = fwd_n(l) * 128;
          const int nk = fwd_nk(l);
          for (int i = 0; i < nk; ++i) {
            const int kb = fwd_chunk(l, i);
            const uint32_t s = g % kStages, ph = (g / kStages) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&full[s], bytes);
            bulk_g2s(sW + s * kStageBytes, layer_src + (size_t)kb * bytes, bytes, &full[s]);
            ++g;
          }
          layer_src += (size_t)nk * bytes;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t g = 0, a_cnt = 0, in_cnt = 0;
      constexpr uint32_t idesc256 = make_idesc_bf16(256, false, false);
      constexpr uint32_t idesc128 = make_idesc_bf16(128, false, false);
      const uint32_t sA_u = smem_u32(sA), sIn_u = smem_u32(sIn), sW_u = smem_u32(sW);
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(in_ready, in_cnt & 1);
        ++in_cnt;
        for (int l = 0; l < kNumFwdLayers; ++l) {
          const uint32_t acc = tmem_base + (uint32_t)(l & 1) * 256u;
          const uint32_t idesc = (l == 9) ? idesc128 : idesc256;
          const int nk = fwd_nk(l);
          const bool stamp = a.prof != nullptr && blockIdx.x == 0 && (int)in_cnt <= a.prof_tiles;
          if (stamp) a.prof[(((int)in_cnt - 1) * kNumFwdLayers + l) * 8 + 0] = clock64();
          long long wait_a = 0, wait_w = 0;
          for (int i = 0; i < nk; ++i) {
            const int kb = fwd_chunk(l, i);
            uint32_t a_addr;
            int nsteps = 4;
            if (l == 0 || (l == 5 && kb == 0)) {
              a_addr = sIn_u;                      // encoded position
            } else if (l == 9 && kb == 4) {
              a_addr = sIn_u + kBlockBytes;        // encoded view direction (K = 32)
              nsteps = 2;
            } else {
              const int ab = (l == 5) ? kb - 1 : kb;
              // a_ready[ab] completes once per producing layer 0..8; layer l consumes round (l - 1)
              const long long w0 = stamp ? clock64() : 0;
              mbar_wait(&a_ready[ab], (a_cnt + (uint32_t)(l - 1)) & 1);
              if (stamp) wait_a += clock64() - w0;
              a_addr = sA_u + ab * kBlockBytes;
            }
            const uint32_t s = g % kStages, ph = (g / kStages) & 1;
            const long long w1 = stamp ? clock64() : 0;
            mbar_wait(&full[s], ph);
            if (stamp) wait_w += clock64() - w1;
            tc_fence_after();
            const uint64_t da = desc_kmajor(a_addr);
            const uint64_t db = desc_kmajor(sW_u + s * kStageBytes);
#pragma unroll 4
            for (int k = 0; k < nsteps; ++k) umma_bf16(acc, da + 2 * k, db + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty[s]);
            ++g;
          }
          umma_commit(&acc_full[l & 1]);
          if (stamp) {
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
    const int half = (warp - 2) >> 2;   // column half: k-blocks {half, half + 2}
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t accn0 = 0, accn1 = 0;
    uint8_t* a_row = sA + row * 128;
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
          if (lane == 0) bulk_wait_read<1>();
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
        if (l & 1) {
          mbar_wait(&acc_full[1], accn1 & 1);
          ++accn1;
        } else {
          mbar_wait(&acc_full[0], accn0 & 1);
          ++accn0;
        }
        tc_fence_after();
        if (stamp) a.prof[((tile_iter * kNumFwdLayers + l) * 8) + 2] = clock64();
        const uint32_t taddr = lane_addr + (uint32_t)(l & 1) * 256u;
        const float* bias = sC + ((l < 8) ? kCBias + 256 * l : (l == 8 ? kCBias8 : kCBias9));
        if (l < 9) {
#pragma unroll 1
          for (int t = 0; t < 2; ++t) {
            const int kb = half + 2 * t;
            uint32_t v0[32], v1[32];
            tmem_ld32(taddr + kb * 64, v0);
            tmem_ld32(taddr + kb * 64 + 32, v1);
            if (kTrain) {
              if (lane == 0) bulk_wait_read<1>();  // this warp's previous store out of block kb has been read
              __syncwarp();
            }
            tmem_ld_wait();
            uint8_t* blk_row = a_row + kb * kBlockBytes;
            float f[32];
            if (l == 8) finish_group<false>(v0, bias + kb * 64, f);
            else finish_group<true>(v0, bias + kb * 64, f);
            if (l == 7) {
#pragma unroll
              for (int i = 0; i < 32; ++i) sigma_part = fmaf(f[i], sC[kCW8Row0 + kb * 64 + i], sigma_part);
            }
            if (kTrain && l < 8) mask_tile[(l * 8 + 2 * kb) * kTileM + row] = relu_mask(f);
            store_group(f, blk_row, row, 0);
            if (l == 8) finish_group<false>(v1, bias + kb * 64 + 32, f);
            else finish_group<true>(v1, bias + kb * 64 + 32, f);
            if (l == 7) {
#pragma unroll
              for (int i = 0; i < 32; ++i) sigma_part = fmaf(f[i], sC[kCW8Row0 + kb * 64 + 32 + i], sigma_part);
            }
            if (kTrain && l < 8) mask_tile[(l * 8 + 2 * kb + 1) * kTileM + row] = relu_mask(f);
            store_group(f, blk_row, row, 4);
            fence_proxy_async();
            if (kTrain) {
              __syncwarp();
              if (lane == 0) {
                const int blk = (l < 8 ? cache_h(l) : kCacheFeat) + kb;
                bulk_s2g(cache_tile + (size_t)blk * kBlockBytes + q * 4096, sA + kb * kBlockBytes + q * 4096, 4096);
                bulk_commit();
              }
            }
            tc_fence_before();
            mbar_arrive(&a_ready[kb]);
          }
          if (l == 7 && half == 1) sX[row * 4 + 3] = sigma_part;
          if (stamp) a.prof[((tile_iter * kNumFwdLayers + l) * 8) + 3] = clock64();
        } else {
          // fc_9 output (128 columns): this half owns columns [64*half, 64*half + 64)
          uint32_t v0[32], v1[32];
          tmem_ld32(taddr + half * 64, v0);
          tmem_ld32(taddr + half * 64 + 32, v1);
          if (kTrain) {
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
          }
          tmem_ld_wait();
          float f[32];
          finish_group<true>(v0, bias + half * 64, f);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            rgb0 = fmaf(f[i], sC[kCWout + half * 64 + i], rgb0);
            rgb1 = fmaf(f[i], sC[kCWout + 128 + half * 64 + i], rgb1);
            rgb2 = fmaf(f[i], sC[kCWout + 256 + half * 64 + i], rgb2);
          }
          if (kTrain) {
            mask_tile[(64 + 2 * half) * kTileM + row] = relu_mask(f);
            store_group(f, a_row + half * kBlockBytes, row, 0);
          }
          finish_group<true>(v1, bias + half * 64 + 32, f);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            rgb0 = fmaf(f[i], sC[kCWout + half * 64 + 32 + i], rgb0);
            rgb1 = fmaf(f[i], sC[kCWout + 128 + half * 64 + 32 + i], rgb1);
            rgb2 = fmaf(f[i], sC[kCWout + 256 + half * 64 + 32 + i], rgb2);
          }
          if (kTrain) {
            mask_tile[(64 + 2 * half + 1) * kTileM + row] = relu_mask(f);
            store_group(f, a_row + half * kBlockBytes, row, 4);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              bulk_s2g(cache_tile + (size_t)(kCacheH9 + half) * kBlockBytes + q * 4096, sA + half * kBlockBytes + q * 4096,
                       4096);
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
