// K4+K5 (and K6) on the sm_100a tensor cores: the NeRF MLP (network/nerf.py:65-121) as a chain of tcgen05 BF16
// MMAs with fp32 accumulation in TMEM, positional encoding (signal_encoder/positional_encoder.py:49-104) computed
// in-kernel as the first layer's operand, weights streamed from L2 by the TMA engine (cp.async.bulk + mbarrier).
//
// Forward chain, one CTA per SM, 128 sample rows per tile, activations never leave the SM:
//
//   warp 0      weight loader   1-D bulk copies of pre-swizzled bf16 weight chunks (N x 64) into a 3-stage ring
//   warp 1      MMA issuer      one thread issues tcgen05.mma (M=128, N=256|128, K=16); accumulators ping-pong
//                               between TMEM columns [0,256) and [256,512) from layer to layer
//   warps 2-5   epilogue        tcgen05.ld -> +bias, ReLU -> bf16 -> swizzled smem = next layer's A operand,
//                               signalled per 64-column k-block so the next layer's MMAs start while the rest of
//                               the accumulator is still being drained; the same warps build the encoded inputs
//
// Tensor-core layers: fc_in, fc_1..fc_7, fc_8 rows 1..256 (features), fc_9.  The density head (fc_8 row 0,
// nerf.py:115) and fc_out + sigmoid (nerf.py:119) are fp32 dot products in the epilogues of layers 7 and 9,
// taken from the fp32 accumulators, so sigma keeps fp32 accuracy where the 1e8 last interval makes it matter.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace nerf {
using namespace tc;

// ------------------------------------------------------------------------------------------------
// geometry of the fixed network (pos 63 / view 27 / feat 256)
// ------------------------------------------------------------------------------------------------
constexpr int kP = 63, kV = 27, kF = 256, kH = 128;
constexpr int kTileM = 128;
constexpr int kNumTcLayers = 10;  // fc_in, fc_1..fc_7, fc_8(feat), fc_9
__host__ __device__ constexpr int layer_nk(int l) { return l == 0 ? 1 : ((l == 5 || l == 9) ? 5 : 4); }
__host__ __device__ constexpr int layer_n(int l) { return l == 9 ? kH : kF; }
constexpr size_t kFwdWeightBytes = (size_t)(1 + 16 + 5 + 12) * 32768 + 5 * 16384;

// fp32 constants block appended to the packed weights (float offsets)
constexpr int kCBias = 0;          // 8 x 256 : fc_in, fc_1..fc_7
constexpr int kCBias8 = 2048;      // 256     : fc_8 bias rows 1..256
constexpr int kCBias9 = 2304;      // 128
constexpr int kCW8Row0 = 2432;     // 256     : fc_8 weight row 0 (density head)
constexpr int kCWout = 2688;       // 3 x 128
constexpr int kCB8_0 = 3072;       // 1
constexpr int kCBout = 3073;       // 3
constexpr int kCFloats = 3080;

constexpr size_t kPackedFwdOff = 0;
constexpr size_t kPackedConstOff = kFwdWeightBytes;
constexpr size_t kPackedBytes = kPackedConstOff + sizeof(float) * kCFloats;

// ------------------------------------------------------------------------------------------------
// weight packing: fp32 (out,in) -> bf16 K-major swizzled chunks in the order the chain consumes them
// ------------------------------------------------------------------------------------------------
struct PackChunk {
  int param;      // index into the 22-pointer parameter array
  int src_row0;   // first source row
  int src_col0;   // first source column
  int ld;         // source leading dimension
  int nrows;      // destination rows (N of the MMA)
  int valid_k;    // columns copied; the rest of the 64 are zero
  uint32_t dst_off;
};
constexpr int kMaxPackChunks = 64;
__constant__ PackChunk c_pack[kMaxPackChunks];

__global__ void __launch_bounds__(256) pack_weights_kernel(const float* const* __restrict__ params_dev,
                                                            uint8_t* __restrict__ packed) {
  const PackChunk pc = c_pack[blockIdx.y];
  int u = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte unit (8 bf16) per thread
  if (u >= pc.nrows * 8) return;
  int n = u >> 3, j = u & 7;
  const float* w = params_dev[pc.param] + (size_t)(pc.src_row0 + n) * pc.ld + pc.src_col0;
  uint32_t out[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int k0 = 8 * j + 2 * e;
    float lo = (k0 < pc.valid_k) ? w[k0] : 0.f;
    float hi = (k0 + 1 < pc.valid_k) ? w[k0 + 1] : 0.f;
    out[e] = pack_bf16(lo, hi);
  }
  uint4* dst = reinterpret_cast<uint4*>(packed + pc.dst_off + n * 128 + ((j ^ (n & 7)) << 4));
  *dst = make_uint4(out[0], out[1], out[2], out[3]);
}

__global__ void pack_consts_kernel(const float* const* __restrict__ params_dev, float* __restrict__ c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kCFloats) return;
  float v = 0.f;
  if (i < kCBias8) v = params_dev[2 * (i >> 8) + 1][i & 255];             // biases of fc_in, fc_1..fc_7
  else if (i < kCBias9) v = params_dev[17][1 + (i - kCBias8)];            // fc_8.bias[1:]
  else if (i < kCW8Row0) v = params_dev[19][i - kCBias9];                 // fc_9.bias
  else if (i < kCWout) v = params_dev[16][i - kCW8Row0];                  // fc_8.weight[0, :]
  else if (i < kCB8_0) v = params_dev[20][i - kCWout];                    // fc_out.weight (3,128)
  else if (i == kCB8_0) v = params_dev[17][0];                            // fc_8.bias[0]
  else if (i < kCBout + 3) v = params_dev[21][i - kCBout];                // fc_out.bias
  c[i] = v;
}

static int build_pack_table(PackChunk* t) {
  int n = 0;
  uint32_t off = 0;
  auto add = [&](int param, int row0, int col0, int ld, int nrows, int valid) {
    t[n++] = PackChunk{param, row0, col0, ld, nrows, valid, off};
    off += (uint32_t)nrows * 128u;
  };
  add(0, 0, 0, kP, kF, kP);                                               // fc_in
  for (int l = 1; l <= 4; ++l)
    for (int kb = 0; kb < 4; ++kb) add(2 * l, 0, 64 * kb, kF, kF, 64);    // fc_1..fc_4
  add(10, 0, 0, kP + kF, kF, kP);                                         // fc_5, pos columns
  for (int kb = 0; kb < 4; ++kb) add(10, 0, kP + 64 * kb, kP + kF, kF, 64);
  for (int l = 6; l <= 7; ++l)
    for (int kb = 0; kb < 4; ++kb) add(2 * l, 0, 64 * kb, kF, kF, 64);    // fc_6, fc_7
  for (int kb = 0; kb < 4; ++kb) add(16, 1, 64 * kb, kF, kF, 64);         // fc_8 rows 1..256
  for (int kb = 0; kb < 4; ++kb) add(18, 0, 64 * kb, kF + kV, kH, 64);    // fc_9, feature columns
  add(18, 0, kF, kF + kV, kH, kV);                                        // fc_9, view columns
  return n;
}

// ------------------------------------------------------------------------------------------------
// forward chain kernel
// ------------------------------------------------------------------------------------------------
constexpr int kStages = 3;
constexpr int kStageBytes = 32768;
constexpr int kFwdThreads = 192;
// shared memory map (bytes from the 1024-aligned base)
constexpr int kSmA = 0;                         // 4 k-blocks x 16 KB : current activations (A operand)
constexpr int kSmIn = 65536;                    // pe block 16 KB | de block 16 KB
constexpr int kSmW = 98304;                     // weight ring
constexpr int kSmC = kSmW + kStages * kStageBytes;
constexpr int kSmBar = kSmC + kCFloats * 4;
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
};

// encoded inputs of one row -> bf16 tile-image rows (pe: 64 columns, de: 32 columns)
__device__ __forceinline__ void encode_row(float x, float y, float z, float dx, float dy, float dz, uint8_t* pe_row,
                                           uint8_t* de_row, int row) {
  float v[64];
  float sn[3], cs[3];
  v[0] = x, v[1] = y, v[2] = z;
  sincosf(x, &sn[0], &cs[0]);
  sincosf(y, &sn[1], &cs[1]);
  sincosf(z, &sn[2], &cs[2]);
#pragma unroll
  for (int l = 0; l < 10; ++l) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      v[3 + 6 * l + c] = sn[c];
      v[6 + 6 * l + c] = cs[c];
      // angle doubling: sin 2a = 2 sin a cos a, cos 2a = (cos a - sin a)(cos a + sin a)
      float s2 = 2.f * sn[c] * cs[c];
      float c2 = (cs[c] - sn[c]) * (cs[c] + sn[c]);
      sn[c] = s2, cs[c] = c2;
    }
  }
  v[63] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint4 q = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                         pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
    *reinterpret_cast<uint4*>(pe_row + ((j ^ (row & 7)) << 4)) = q;
  }
  float w[32];
  w[0] = dx, w[1] = dy, w[2] = dz;
  sincosf(dx, &sn[0], &cs[0]);
  sincosf(dy, &sn[1], &cs[1]);
  sincosf(dz, &sn[2], &cs[2]);
#pragma unroll
  for (int l = 0; l < 4; ++l) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      w[3 + 6 * l + c] = sn[c];
      w[6 + 6 * l + c] = cs[c];
      float s2 = 2.f * sn[c] * cs[c];
      float c2 = (cs[c] - sn[c]) * (cs[c] + sn[c]);
      sn[c] = s2, cs[c] = c2;
    }
  }
#pragma unroll
  for (int i = 27; i < 32; ++i) w[i] = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 q = make_uint4(pack_bf16(w[8 * j], w[8 * j + 1]), pack_bf16(w[8 * j + 2], w[8 * j + 3]),
                         pack_bf16(w[8 * j + 4], w[8 * j + 5]), pack_bf16(w[8 * j + 6], w[8 * j + 7]));
    *reinterpret_cast<uint4*>(de_row + ((j ^ (row & 7)) << 4)) = q;
  }
}

__global__ void __launch_bounds__(kFwdThreads, 1) mlp_fwd_kernel(FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem + kSmA;
  uint8_t* sIn = smem + kSmIn;
  uint8_t* sW = smem + kSmW;
  float* sC = reinterpret_cast<float*>(smem + kSmC);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmBar);
  uint64_t* full = bars;                  // [kStages]
  uint64_t* empty = bars + kStages;       // [kStages]
  uint64_t* a_ready = bars + 2 * kStages; // [4]
  uint64_t* in_ready = a_ready + 4;       // [1]
  uint64_t* acc_full = in_ready + 1;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = (a.m + kTileM - 1) / kTileM;

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
    mbar_init(in_ready, 128);
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
        const uint8_t* src = a.packed + kPackedFwdOff;
        for (int l = 0; l < kNumTcLayers; ++l) {
          const uint32_t bytes = layer_n(l) * 128;
          for (int kb = 0; kb < layer_nk(l); ++kb) {
            const uint32_t s = g % kStages, ph = (g / kStages) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&full[s], bytes);
            bulk_g2s(sW + s * kStageBytes, src, bytes, &full[s]);
            src += bytes;
            ++g;
          }
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
        for (int l = 0; l < kNumTcLayers; ++l) {
          const uint32_t acc = tmem_base + (uint32_t)(l & 1) * 256u;
          const uint32_t idesc = (l == 9) ? idesc128 : idesc256;
          const int nk = layer_nk(l);
          // a_ready[kb] completes once per producing layer (layers 0..8); consumer layer l>=1 reads round l-1
          for (int kb = 0; kb < nk; ++kb) {
            uint32_t a_addr;
            int nsteps = 4;
            if (l == 0) {
              a_addr = sIn_u;
            } else if (l == 5 && kb == 0) {
              a_addr = sIn_u;
            } else if (l == 9 && kb == 4) {
              a_addr = sIn_u + 16384;
              nsteps = 2;
            } else {
              const int ab = (l == 5) ? kb - 1 : kb;
              mbar_wait(&a_ready[ab], (a_cnt + (uint32_t)(l - 1)) & 1);
              a_addr = sA_u + ab * 16384;
            }
            const uint32_t s = g % kStages, ph = (g / kStages) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint64_t da = desc_kmajor(a_addr);
            const uint64_t db = desc_kmajor(sW_u + s * kStageBytes);
#pragma unroll 4
            for (int k = 0; k < nsteps; ++k) umma_bf16(acc, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty[s]);
            ++g;
          }
          umma_commit(&acc_full[l & 1]);
        }
        a_cnt += 9;  // nine a_ready rounds per tile
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (128 rows)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t accn0 = 0, accn1 = 0;
    uint8_t* a_row = sA + row * 128;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t grow = tile * kTileM + row;
      // ---- encoded inputs (cube.py:62-69): point and view direction of this row
      {
        float x = 0.f, y = 0.f, z = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;
        if (grow < a.m) {
          if (a.pts != nullptr) {
            x = __ldg(a.pts + 3 * grow), y = __ldg(a.pts + 3 * grow + 1), z = __ldg(a.pts + 3 * grow + 2);
            dx = __ldg(a.dirs + 3 * grow), dy = __ldg(a.dirs + 3 * grow + 1), dz = __ldg(a.dirs + 3 * grow + 2);
          } else {
            const int64_t ray = grow / a.s;
            const float tt = __ldg(a.t + grow);
            dx = __ldg(a.ray_d + 3 * ray), dy = __ldg(a.ray_d + 3 * ray + 1), dz = __ldg(a.ray_d + 3 * ray + 2);
            // stratified_sampler.py:126: o + t*d, product and sum rounded separately
            x = __fadd_rn(__ldg(a.ray_o + 3 * ray), __fmul_rn(tt, dx));
            y = __fadd_rn(__ldg(a.ray_o + 3 * ray + 1), __fmul_rn(tt, dy));
            z = __fadd_rn(__ldg(a.ray_o + 3 * ray + 2), __fmul_rn(tt, dz));
          }
        }
        encode_row(x, y, z, dx, dy, dz, sIn + row * 128, sIn + 16384 + row * 128, row);
        fence_proxy_async();
        mbar_arrive(in_ready);
      }
      float sigma_pre = sC[kCB8_0];
      float rgb0 = sC[kCBout], rgb1 = sC[kCBout + 1], rgb2 = sC[kCBout + 2];
      for (int l = 0; l < kNumTcLayers; ++l) {
        if (l & 1) {
          mbar_wait(&acc_full[1], accn1 & 1);
          ++accn1;
        } else {
          mbar_wait(&acc_full[0], accn0 & 1);
          ++accn0;
        }
        tc_fence_after();
        const uint32_t taddr = lane_addr + (uint32_t)(l & 1) * 256u;
        const float* bias = sC + ((l < 8) ? kCBias + 256 * l : (l == 8 ? kCBias8 : kCBias9));
        if (l < 9) {
#pragma unroll 1
          for (int c = 0; c < 8; ++c) {
            uint32_t v[32];
            tmem_ld32(taddr + c * 32, v);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float t = __uint_as_float(v[i]) + bias[c * 32 + i];
              f[i] = (l == 8) ? t : fmaxf(t, 0.f);
            }
            if (l == 7) {
#pragma unroll
              for (int i = 0; i < 32; ++i) sigma_pre = fmaf(f[i], sC[kCW8Row0 + c * 32 + i], sigma_pre);
            }
            uint8_t* blk = a_row + (c >> 1) * 16384;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 qv = make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                    pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
              const int chunk = (c & 1) * 4 + j;
              *reinterpret_cast<uint4*>(blk + ((chunk ^ (row & 7)) << 4)) = qv;
            }
            if (c & 1) {
              fence_proxy_async();
              tc_fence_before();
              mbar_arrive(&a_ready[c >> 1]);
            }
          }
        } else {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            tmem_ld32(taddr + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float h = fmaxf(__uint_as_float(v[i]) + bias[c * 32 + i], 0.f);
              rgb0 = fmaf(h, sC[kCWout + c * 32 + i], rgb0);
              rgb1 = fmaf(h, sC[kCWout + 128 + c * 32 + i], rgb1);
              rgb2 = fmaf(h, sC[kCWout + 256 + c * 32 + i], rgb2);
            }
          }
          tc_fence_before();
          if (grow < a.m) {
            a.sigma[grow] = fmaxf(sigma_pre, 0.f);  // nerf.py:115
            a.rgb[3 * grow] = 1.f / (1.f + __expf(-rgb0));  // nerf.py:119
            a.rgb[3 * grow + 1] = 1.f / (1.f + __expf(-rgb1));
            a.rgb[3 * grow + 2] = 1.f / (1.f + __expf(-rgb2));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// self test: one UMMA tile, both operand majors
// ------------------------------------------------------------------------------------------------
// variant 0: a (128 x k) row-major, b (n x k) row-major        -> K-major images
// variant 1: a (k x 128) row-major (= A^T), b (k x n) row-major -> MN-major images (the wgrad form)
__global__ void __launch_bounds__(128, 1) selftest_umma_kernel(const uint16_t* __restrict__ a,
                                                               const uint16_t* __restrict__ b, float* __restrict__ d,
                                                               int n, int k, int variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = k / 64;
  uint8_t* sA;
  uint8_t* sB;
  uint32_t a_blk, b_blk;  // byte stride between 64-column blocks
  if (variant == 0) {
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
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16((uint32_t)n, variant == 1, variant == 1);
    for (int kk = 0; kk < k / 16; ++kk) {
      uint64_t da, db;
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
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

static bool g_pack_table_ready = false;
static int g_pack_chunks = 0;

}  // namespace nerf

using namespace nerf;

extern "C" {

size_t nerf_mlp_bf16_packed_bytes(void) { return kPackedBytes; }

int nerf_mlp_bf16_pack(const float* const* params, void* packed_dev, nerf_stream_t stream) {
  NERF_CHECK_ARG(params && packed_dev, "nerf_mlp_bf16_pack: null pointer");
  cudaStream_t st = as_stream(stream);
  if (!g_pack_table_ready) {
    PackChunk table[kMaxPackChunks];
    g_pack_chunks = build_pack_table(table);
    NERF_CUDA(cudaMemcpyToSymbol(c_pack, table, sizeof(PackChunk) * g_pack_chunks));
    g_pack_table_ready = true;
  }
  // the 22 parameter pointers ride at the tail of the packed buffer's constant block? no: a small device array
  static thread_local const float** params_dev = nullptr;
  if (!params_dev) NERF_CUDA(cudaMalloc(&params_dev, sizeof(float*) * NERF_NUM_PARAM_TENSORS));
  NERF_CUDA(cudaMemcpyAsync(params_dev, params, sizeof(float*) * NERF_NUM_PARAM_TENSORS, cudaMemcpyHostToDevice, st));
  dim3 grid((kF * 8 + 255) / 256, g_pack_chunks);
  pack_weights_kernel<<<grid, 256, 0, st>>>(params_dev, reinterpret_cast<uint8_t*>(packed_dev) + kPackedFwdOff);
  NERF_LAUNCH_CHECK();
  pack_consts_kernel<<<(kCFloats + 255) / 256, 256, 0, st>>>(
      params_dev, reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(packed_dev) + kPackedConstOff));
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

size_t nerf_mlp_bf16_cache_bytes(int64_t m) {
  (void)m;
  return 0;
}

int nerf_mlp_bf16_forward(const void* packed_dev, const float* pts_dev, const float* dirs_dev,
                          const float* ray_o_dev, const float* ray_d_dev, const float* t_dev, int s, int64_t m,
                          float* sigma_dev, float* rgb_dev, void* cache_dev, nerf_stream_t stream) {
  NERF_CHECK_ARG(packed_dev && sigma_dev && rgb_dev, "nerf_mlp_bf16_forward: null pointer");
  NERF_CHECK_ARG((pts_dev && dirs_dev) || (ray_o_dev && ray_d_dev && t_dev && s > 0),
                 "nerf_mlp_bf16_forward: give (pts, dirs) or (ray_o, ray_d, t, s)");
  NERF_CHECK_ARG(m >= 0, "nerf_mlp_bf16_forward: negative row count");
  NERF_CHECK_ARG(cache_dev == nullptr, "nerf_mlp_bf16_forward: training cache not implemented yet");
  if (m == 0) return NERF_OK;
  static bool attr_set = false;
  if (!attr_set) {
    NERF_CUDA(cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
    attr_set = true;
  }
  FwdArgs a;
  a.packed = reinterpret_cast<const uint8_t*>(packed_dev);
  a.pts = pts_dev, a.dirs = dirs_dev, a.ray_o = ray_o_dev, a.ray_d = ray_d_dev, a.t = t_dev, a.s = s, a.m = m;
  a.sigma = sigma_dev, a.rgb = rgb_dev;
  const int64_t ntiles = (m + kTileM - 1) / kTileM;
  const int grid = (int)((ntiles < sm_count()) ? ntiles : sm_count());
  mlp_fwd_kernel<<<grid, kFwdThreads, kFwdSmemBytes, as_stream(stream)>>>(a);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

size_t nerf_mlp_bf16_bwd_scratch_bytes(int64_t m) {
  (void)m;
  return 0;
}

int nerf_mlp_bf16_backward(const void* packed_dev, const void* cache_dev, const float* rgb_dev, int64_t m,
                           const float* g_sigma_dev, const float* g_rgb_dev, float* const* grads,
                           void* scratch_dev, nerf_stream_t stream) {
  (void)packed_dev, (void)cache_dev, (void)rgb_dev, (void)m, (void)g_sigma_dev, (void)g_rgb_dev, (void)grads;
  (void)scratch_dev, (void)stream;
  set_error("nerf_mlp_bf16_backward: not implemented yet");
  return NERF_ERR_ARG;
}

int nerf_selftest_umma(const uint16_t* a_dev, const uint16_t* b_dev, float* d_dev, int n, int k, int variant,
                       nerf_stream_t stream) {
  NERF_CHECK_ARG(a_dev && b_dev && d_dev, "nerf_selftest_umma: null pointer");
  NERF_CHECK_ARG(n >= 64 && n <= 256 && n % 64 == 0 && k >= 64 && k <= 256 && k % 64 == 0 && (variant == 0 || variant == 1),
                 "nerf_selftest_umma: n, k must be multiples of 64 in [64,256]; variant 0|1");
  size_t smem = (size_t)128 * k * 2 + (size_t)n * k * 2 + 1024;
  NERF_CUDA(cudaFuncSetAttribute(selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  selftest_umma_kernel<<<1, 128, smem, as_stream(stream)>>>(a_dev, b_dev, d_dev, n, k, variant);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

}  // extern "C"
