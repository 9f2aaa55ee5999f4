// K4+K5 on the sm_100a tensor cores: the NeRF MLP forward (network/nerf.py:65-121) as a chain of tcgen05 BF16 MMAs
// with fp32 accumulation in TMEM; positional encoding (signal_encoder/positional_encoder.py:49-104, as applied by
// scene/primitives/cube.py:62-69) is computed in-kernel as the first layer's operand; weights are streamed from L2
// by the TMA engine (cp.async.bulk + mbarrier) through a ring of 16 KB stages (one chunk of 128 outputs x 64 inputs).
//
// One CTA per SM works on TWO 128-row tiles at a time ("slots" X and Y).  A layer's output is the next layer's
// input, so within one tile the tensor pipe has to wait for the epilogue (TMEM -> bias/ReLU -> bf16 -> TMEM) between
// layers; with two independent tiles the pipe runs tile Y's MMAs while tile X is in its epilogue, and every weight
// chunk fetched from L2 is used twice.  Activations live in TENSOR MEMORY between layers, per slot s:
//
//   TMEM columns [256 s, +128)        fp32 accumulator of one N-half (128 output columns) of the current layer
//                [256 s + 128, +128)  A operand: the layer input, 128 rows x 256 bf16 as packed pairs
//
//   warp 0      weight loader   1-D bulk copies of pre-swizzled bf16 weight chunks (shared by both slots)
//   warps 1,18  MMA issuers     one per slot; one elected lane issues tcgen05.mma (M=128, N=128, K=16) in the TS form:
//                               A from TMEM, B from shared memory -- this takes the activations off the shared-memory
//                               port, which limits the SS form.  A 256-wide layer is two N-halves through the same
//                               128 accumulator columns: half 1 starts as soon as the epilogue has pulled half 0 into
//                               registers (acc_free), and the layer input is overwritten in place once half 1 is done.
//   warps 2-17  epilogue        four warps per TMEM lane quarter, each thread owns 32 accumulator columns per event
//                               (slot, N-half): tcgen05.ld -> +bias, ReLU -> bf16 pairs; the half-0 result is held in
//                               registers and both halves are written with tcgen05.st after half 1 (a_ready).  The
//                               same warps build the encoded inputs of the next tile pair (those k-blocks stay in
//                               shared memory, SS form).
//
// In training mode the epilogue additionally writes every layer input as a tile image (two warps fill one 32-row x
// 128 B staging slice, one 4 KB bulk store per slice) plus ReLU sign-bit masks into the training cache -- after the
// barrier arrives, off the tensor pipe's critical path.
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

constexpr int kMaxStages = 10;
template <bool kTrain>
struct FwdCfg {
  static constexpr int kStages = 6;  // (ring depth 6 .. 10 measured equal)
};
constexpr int kEpiWarps = 16;
constexpr int kSlotThreads = kEpiWarps * 32 / 2;   // epilogue threads serving one tile slot
constexpr int kMmaWarpB = 2 + kEpiWarps;
constexpr int kFwdThreads = 32 * (3 + kEpiWarps);  // loader, MMA issuer X, 16 epilogue warps, MMA issuer Y
// shared memory map (bytes from the 1024-aligned base)
constexpr int kSmIn = 0;                        // pe block slot X | pe block slot Y | de block (X: chunks 0-3, Y: chunks 4-7)
constexpr int kSmC = 3 * kBlockBytes;           // fp32 constants
constexpr int kSmX = kSmC + kCFloats * 4;       // 2 slots x 128 rows x {rgb0, rgb1, rgb2, sigma} partial sums of block 1
constexpr int kSmBar = kSmX + 2 * 128 * 16;
constexpr int kSmStage = (kSmBar + 512 + 1023) / 1024 * 1024;  // training: 16 warps x 4 KB staging slices;
                                                               // inference: second set of the three input blocks
constexpr int kSmW = kSmStage + 65536;          // weight ring
constexpr int kSmTotal = kSmW + 6 * kChunkBytes;
constexpr int kFwdSmemBytes = kSmTotal + 1024;  // + alignment slack
static_assert(kFwdSmemBytes <= 232448, "shared memory budget");
// tensor memory map (columns), per slot
constexpr uint32_t kTmSlot = 256;
constexpr uint32_t kTmA = 128;

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
  int dbg_store;  // debug: 0 normal, 1 skip the training-cache stores, 2 wrap them onto 64 tiles (L2 resident)
};

// [v | sin(2^l v) | cos(2^l v)]_{l<L} for a 3-vector, written as bf16 into NCHUNK 16-byte chunks (logical chunks
// chunk0 .. chunk0 + NCHUNK - 1) of a swizzled tile-image row.  Higher octaves come from angle doubling
// (sin 2a = 2 sin a cos a, cos 2a = (cos a - sin a)(cos a + sin a)); the accumulated error (~2^l ulp) is far below
// bf16 resolution.  With kGlobal the same chunks also go to `global_row_ptr` (logical chunks 0 .. NCHUNK - 1).
template <int L, int NCHUNK, bool kGlobal>
__device__ __forceinline__ void encode_row(float x, float y, float z, uint8_t* row_ptr, int chunk0,
                                           uint8_t* global_row_ptr, int row) {
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
    *reinterpret_cast<uint4*>(row_ptr + (((chunk0 + j) ^ (row & 7)) << 4)) = q;
    if (kGlobal) *reinterpret_cast<uint4*>(global_row_ptr + ((j ^ (row & 7)) << 4)) = q;
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

// two fp32 adds in one instruction (FADD2 on sm_100)
__device__ __forceinline__ void add2(float a0, float a1, float b0, float b1, float& o0, float& o1) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(o0), "=f"(o1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// two fp32 -> packed bf16 pair (lo in the low half), ReLU folded into the conversion
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t w;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(hi), "f"(lo));
  return w;
}

// 32 accumulator columns -> +bias -> (ReLU) -> 16 packed bf16 pairs, without materialising the fp32 activations:
// FADD2 for the bias, the ReLU rides on the bf16 conversion.  Returns the sign bits like finish_group when MASK.
template <bool RELU, bool MASK>
__device__ __forceinline__ uint32_t finish_pack(const uint32_t (&v)[32], const float* __restrict__ bias, uint32_t (&w)[16]) {
  uint32_t neg = 0;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 b = *reinterpret_cast<const float4*>(bias + i);
    float t0, t1, t2, t3;
    add2(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), b.x, b.y, t0, t1);
    add2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]), b.z, b.w, t2, t3);
    if (RELU && MASK) {
      neg = __funnelshift_l(__float_as_uint(t0), neg, 1);
      neg = __funnelshift_l(__float_as_uint(t1), neg, 1);
      neg = __funnelshift_l(__float_as_uint(t2), neg, 1);
      neg = __funnelshift_l(__float_as_uint(t3), neg, 1);
    }
    w[i / 2] = RELU ? pack_bf16_relu(t0, t1) : pack_bf16(t0, t1);
    w[i / 2 + 1] = RELU ? pack_bf16_relu(t2, t3) : pack_bf16(t2, t3);
  }
  return neg;
}

// 32 fp32 values -> 16 packed bf16 pairs
__device__ __forceinline__ void pack_group(const float (&f)[32], uint32_t (&w)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) w[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
}

// 16 packed words (32 columns) -> four 16-byte chunks of a swizzled tile-image row
__device__ __forceinline__ void store_words(const uint32_t (&w)[16], uint8_t* blk_row, int row, int chunk0) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(blk_row + (((chunk0 + j) ^ (row & 7)) << 4)) =
        make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
}

template <bool kTrain>
__global__ void __launch_bounds__(kFwdThreads, 1) mlp_fwd_kernel(FwdArgs a) {
  constexpr int kStages = FwdCfg<kTrain>::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sIn = smem + kSmIn;
  uint8_t* sStage = smem + kSmStage;
  uint8_t* sW = smem + kSmW;
  float* sC = reinterpret_cast<float*>(smem + kSmC);
  float4* sX = reinterpret_cast<float4*>(smem + kSmX);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmBar);
  uint64_t* full = bars;                      // [kMaxStages] weight chunk landed
  uint64_t* empty = bars + kMaxStages;        // [kMaxStages] both slots' MMAs on the chunk have completed
  uint64_t* in_ready = bars + 2 * kMaxStages; // [2] per slot: encoded inputs written, previous tile fully drained
  uint64_t* a_ready = in_ready + 2;           // [2] per slot: layer input rewritten in TMEM (one completion per layer 0..8)
  uint64_t* acc_free = a_ready + 2;           // [2] per slot: N-half 0 pulled out of the accumulator
  uint64_t* acc_full = acc_free + 2;          // [2] per slot: the MMAs of one N-half have completed
  uint64_t* drained = acc_full + 2;           // [2] per slot (inference): fc_9's accumulator has been read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(drained + 2);
  // Inference double-buffers the encoded inputs (second set in the unused staging area): the next tile's inputs are
  // written during the current tile's last layers, so the first MMAs of a tile only wait for the accumulator.
  constexpr bool kEarlyInputs = !kTrain;
  uint8_t* const in_set[2] = {sIn, kEarlyInputs ? sStage : sIn};

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = num_tiles(a.m);
  const int64_t npairs = (ntiles + 1) / 2;

  {
    const float* cgp = reinterpret_cast<const float*>(a.packed + kPackedConstOff);
    for (int i = threadIdx.x; i < kCFloats; i += kFwdThreads) sC[i] = __ldg(cgp + i);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&in_ready[i], kSlotThreads);
      mbar_init(&a_ready[i], kSlotThreads);
      mbar_init(&acc_free[i], kSlotThreads);
      mbar_init(&acc_full[i], 1);
      mbar_init(&drained[i], kSlotThreads);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ weight loader (warp in lock step, one lane issues)
    const bool leader = elect_one();
    uint32_t g = 0;
    for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      int chunk0 = 0;  // first chunk of layer l in the packed order (fwd_layer_chunk0)
      for (int l = 0; l < kNumFwdLayers; chunk0 += fwd_nk(l) * fwd_nh(l), ++l) {
        for (int h = 0; h < fwd_nh(l); ++h) {
          for (int kb = 0; kb < fwd_nk(l); ++kb) {  // consumption order of the issuers
            const uint8_t* src = a.packed + kPackedFwdOff + (size_t)(chunk0 + kb * fwd_nh(l) + h) * kChunkBytes;
            const uint32_t s = g % kStages, ph = (g / kStages) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            if (leader) {
              mbar_arrive_expect_tx(&full[s], kChunkBytes);
              bulk_g2s(sW + s * kChunkBytes, src, kChunkBytes, &full[s]);
            }
            __syncwarp();
            ++g;
          }
        }
      }
    }
  } else if (warp == 1 || warp == kMmaWarpB) {
    // ------------------------------------------------------------------ MMA issuers (one warp per slot)
    // tcgen05.mma issue is effectively synchronous (the issuing thread is held while its MMA executes) and every
    // mbarrier probe costs ~125-150 cycles, so each slot has its own issuer: while one polls, the other's MMAs run.
    // Each warp runs its loop in lock step and one elected lane issues.
    const int slot = (warp == 1) ? 0 : 1;
    const bool leader = elect_one();
    uint32_t g = 0, n_in = 0, n_a = 0, n_free = 0;
    constexpr uint32_t idesc = make_idesc_bf16(128, false, false);
    const uint32_t sW_u = smem_u32(sW);
    const uint32_t acc = tmem_base + (uint32_t)slot * kTmSlot;
    const uint32_t a_tm = acc + kTmA;
    int iter = 0;
    for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++iter) {
      mbar_wait(&in_ready[slot], n_in & 1);
      ++n_in;
      if (kEarlyInputs && iter > 0) mbar_wait(&drained[slot], (uint32_t)(iter - 1) & 1);  // previous tile's fc_9 is out
      const uint32_t sPe_u = smem_u32(in_set[iter & 1]) + (uint32_t)slot * kBlockBytes;
      const uint32_t sDe_u = smem_u32(in_set[iter & 1]) + 2u * kBlockBytes + (uint32_t)slot * 64u;
      for (int l = 0; l < kNumFwdLayers; ++l) {
        const int nk = fwd_nk(l);
        const bool stamp = a.prof != nullptr && blockIdx.x == 0 && iter < a.prof_tiles && slot == 0;
        if (stamp && leader) a.prof[(iter * kNumFwdLayers + l) * 8 + 0] = clock64();
        long long wait_a = 0, wait_w = 0;
        for (int h = 0; h < fwd_nh(l); ++h) {
          {
            const long long w0 = stamp ? clock64() : 0;
            if (h == 0) {
              if (l >= 1) {  // the epilogue has rewritten this slot's layer input (and drained the accumulator)
                issuer_wait(&a_ready[slot], n_a & 1);
                ++n_a;
              }
            } else {         // N-half 0 is out of the accumulator
              issuer_wait(&acc_free[slot], n_free & 1);
              ++n_free;
            }
            if (stamp) wait_a += clock64() - w0;
          }
          for (int kb = 0; kb < nk; ++kb) {
            int ab = -1;             // A k-block in TMEM, or -1 for the shared-memory blocks
            uint32_t a_smem = 0;
            int nsteps = 4;
            if (l == 0 || (l == 5 && kb == 0)) {
              a_smem = sPe_u;                      // encoded position
            } else if (l == 9 && kb == 4) {
              a_smem = sDe_u;                      // encoded view direction (K = 32)
              nsteps = 2;
            } else {
              ab = (l == 5) ? kb - 1 : kb;
            }
            const uint32_t s = g % kStages, ph = (g / kStages) & 1;
            {
              const long long w1 = stamp ? clock64() : 0;
              mbar_wait(&full[s], ph);
              if (stamp) wait_w += clock64() - w1;
            }
            tc_fence_after();
            if (leader) {
              const uint64_t db = desc_kmajor(sW_u + s * kChunkBytes);
              if (ab >= 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_ts(acc, a_tm + (uint32_t)(ab * 32 + k * 8), db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              } else {
                const uint64_t da = desc_kmajor(a_smem);
#pragma unroll 4
                for (int k = 0; k < nsteps; ++k) umma_bf16(acc, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              }
              umma_commit(&empty[s]);
            }
            __syncwarp();
            ++g;
          }
          if (leader) umma_commit(&acc_full[slot]);
          __syncwarp();
        }
        if (stamp && leader) {
          unsigned long long* pr = a.prof + (iter * kNumFwdLayers + l) * 8;
          pr[1] = clock64();
          pr[4] = (unsigned long long)wait_a;
          pr[5] = (unsigned long long)wait_w;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    // 16 warps = 4 per TMEM lane quarter; two of them serve tile slot X and two slot Y.  A warp owns 32 rows x the
    // 64 columns [64 blk, +64) of each N-half of its slot's accumulator = one whole k-block (128-byte row) of the next
    // layer's operand, so its training-cache store needs nobody else: private 4 KB staging slice, one bulk store.
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int e = (warp - 2) >> 2;
    const int slot = e >> 1;            // tile slot served by this warp
    const int blk = e & 1;              // which 64-column block of an N-half
    const int row = q * 32 + lane;
    const uint32_t t_slot = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)slot * kTmSlot;
    uint8_t* st_row = sStage + (warp - 2) * 4096 + lane * 128;
    uint32_t n_full = 0;
    int iter = 0;

    // the staged 32 rows x 128 B -> HBM; `gdst` = rows [32 q, 32 q + 32) of one block
    auto bulk_out = [&](uint8_t* gdst) {
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        bulk_s2g(gdst, st_row, 4096);  // lane 0: st_row = start of the slice
        bulk_commit();
      }
    };
    auto staging_free = [&]() {
      if (lane == 0) bulk_wait_read<0>();  // the previous store out of the slice has been read
      __syncwarp();
    };

    // this thread's input row of a tile: the sample position (block 0 warps) or the view direction (block 1 warps)
    const bool is_pe = blk == 0;
    auto load_xyz = [&](int64_t tile_, float& x, float& y, float& z) {
      const int64_t grow_ = tile_ * kTileM + row;
      x = y = z = 0.f;
      if (grow_ < a.m) {
        if (a.pts != nullptr) {
          const float* src = is_pe ? a.pts : a.dirs;
          x = __ldg(src + 3 * grow_), y = __ldg(src + 3 * grow_ + 1), z = __ldg(src + 3 * grow_ + 2);
        } else {
          const int64_t ray = a.m < (int64_t(1) << 31) ? (int64_t)((uint32_t)grow_ / (uint32_t)a.s) : grow_ / a.s;
          x = __ldg(a.ray_d + 3 * ray), y = __ldg(a.ray_d + 3 * ray + 1), z = __ldg(a.ray_d + 3 * ray + 2);
          if (is_pe) {
            const float tt = __ldg(a.t + grow_);
            // stratified_sampler.py:126: o + t*d, product and sum rounded separately
            x = __fadd_rn(__ldg(a.ray_o + 3 * ray), __fmul_rn(tt, x));
            y = __fadd_rn(__ldg(a.ray_o + 3 * ray + 1), __fmul_rn(tt, y));
            z = __fadd_rn(__ldg(a.ray_o + 3 * ray + 2), __fmul_rn(tt, z));
          }
        }
      }
    };
    float x, y, z;  // this thread's input row of the tile that is encoded next
    // ---- encoded inputs (cube.py:62-69) of tile `tile_` into one input set: block 0 warps encode the point, block 1
    //      warps the view direction; then in_ready.  (Training: the pe block also goes to the cache by bulk store.)
    auto encode_inputs = [&](uint8_t* set, int64_t tile_, uint8_t* cache_tile_, bool to_cache_) {
      if (is_pe) {
        if (kTrain) staging_free();  // also covers the previous pair's store out of the pe block (same thread's groups)
        encode_row<10, 8, false>(x, y, z, set + slot * kBlockBytes + row * 128, 0, nullptr, row);
        fence_proxy_async();
        if (kTrain) {
          __syncwarp();
          if (lane == 0 && to_cache_) {
            bulk_s2g(cache_tile_ + cache_slice_off(kCachePe, q), set + slot * kBlockBytes + q * 4096, 4096);
            bulk_commit();
          }
        }
      } else {
        if (kTrain && to_cache_)
          encode_row<4, 4, true>(x, y, z, set + 2 * kBlockBytes + row * 128, 4 * slot,
                                 cache_tile_ + cache_slice_off(kCacheDe, q) + lane * 128, row);
        else
          encode_row<4, 4, false>(x, y, z, set + 2 * kBlockBytes + row * 128, 4 * slot, nullptr, row);
        fence_proxy_async();
      }
      (void)tile_;
      mbar_arrive(&in_ready[slot]);  // training: inputs written AND this thread has left the slot's previous tile
    };
    load_xyz(2 * (int64_t)blockIdx.x + slot, x, y, z);

    for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++iter) {
      const bool stamp = a.prof != nullptr && blockIdx.x == 0 && iter < a.prof_tiles && warp == 2 && lane == 0;
      const int64_t tile = 2 * pair + slot;
      const int64_t grow = tile * kTileM + row;
      const bool to_cache = kTrain && tile < ntiles;
      uint8_t* cache_tile = kTrain ? a.cache + (size_t)tile * kCacheTileBytes : nullptr;
      uint32_t* mask_row =
          kTrain ? reinterpret_cast<uint32_t*>(a.cache + cache_mask_offset(a.m) + (size_t)tile * kMaskTileBytes) + row : nullptr;
      if (!kEarlyInputs || iter == 0) encode_inputs(in_set[0], tile, cache_tile, to_cache);
      if (kEarlyInputs) load_xyz(2 * (pair + gridDim.x) + slot, x, y, z);  // consumed during layer 8 of this tile
      float sigma_part = 0.f;
      uint32_t wh[2][16];  // bf16 pairs of N-half 0, held until half 1's MMAs have stopped reading the layer input
#pragma unroll 1
      for (int l = 0; l < kNumFwdLayers - 1; ++l) {
        const float* bias = sC + ((l < 8) ? kCBias + 256 * l : kCBias8);
        // one 32-column group: accumulator values -> bf16 pairs `w` (+ ReLU sign bits); layer 7 also feeds the density head
        auto finish = [&](const uint32_t (&v)[32], int col0, uint32_t (&w)[16]) -> uint32_t {
          if (l == 7) {
            // the density head (fc_8 row 0) is an fp32 dot product with the fp32 activations
            float f[32];
            const uint32_t neg = finish_group<true>(v, bias + col0, f);
            pack_group(f, w);
#pragma unroll
            for (int i = 0; i < 32; ++i) sigma_part = fmaf(f[i], sC[kCW8Row0 + col0 + i], sigma_part);
            return neg;
          }
          if (l == 8) return finish_pack<false, false>(v, bias + col0, w);
          return finish_pack<true, kTrain>(v, bias + col0, w);
        };
        const int cblk0 = (l < 8 ? cache_h(l) : kCacheFeat) + blk;  // cache block of N-half 0's columns
        // ---- N-half 0: pull it out of the accumulator first (the issuer is waiting for that), then convert and hold it
        {
          epilogue_wait(&acc_full[slot], n_full & 1);
          ++n_full;
          tc_fence_after();
          if (stamp) a.prof[((iter * kNumFwdLayers + l) * 8) + 2] = clock64();
          uint32_t v0[32], v1[32];
          tmem_ld32(t_slot + 64 * blk, v0);
          tmem_ld32(t_slot + 64 * blk + 32, v1);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&acc_free[slot]);
          const uint32_t neg0 = finish(v0, 64 * blk, wh[0]);
          const uint32_t neg1 = finish(v1, 64 * blk + 32, wh[1]);
          if (to_cache) {
            staging_free();
            if (l < 8) {
              mask_row[(l * 8 + 2 * blk) * kTileM] = neg0;
              mask_row[(l * 8 + 2 * blk + 1) * kTileM] = neg1;
            }
            store_words(wh[0], st_row, row, 0);
            store_words(wh[1], st_row, row, 4);
            if (a.dbg_store != 1) bulk_out(cache_tile + cache_slice_off(cblk0, q));
          }
        }
        // ---- N-half 1: once it is complete nothing reads the layer input any more -> overwrite it in place, signal
        //      the issuer, and only then do the training-cache stores
        {
          epilogue_wait(&acc_full[slot], n_full & 1);
          ++n_full;
          tc_fence_after();
          uint32_t v[32], w0[16], w1[16];
          tmem_ld32(t_slot + 64 * blk, v);
          tmem_ld_wait();
          const uint32_t neg0 = finish(v, 128 + 64 * blk, w0);
          tmem_st16(t_slot + kTmA + 32 * blk, wh[0]);            // output columns [64 blk, +32)
          tmem_st16(t_slot + kTmA + 64 + 32 * blk, w0);          // output columns [128 + 64 blk, +32)
          tmem_ld32(t_slot + 64 * blk + 32, v);
          tmem_ld_wait();
          const uint32_t neg1 = finish(v, 128 + 64 * blk + 32, w1);
          tmem_st16(t_slot + kTmA + 32 * blk + 16, wh[1]);
          tmem_st16(t_slot + kTmA + 64 + 32 * blk + 16, w1);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&a_ready[slot]);
          if (to_cache) {
            staging_free();
            if (l < 8) {
              mask_row[(l * 8 + 4 + 2 * blk) * kTileM] = neg0;
              mask_row[(l * 8 + 4 + 2 * blk + 1) * kTileM] = neg1;
            }
            store_words(w0, st_row, row, 0);
            store_words(w1, st_row, row, 4);
            if (a.dbg_store != 1) bulk_out(cache_tile + cache_slice_off(cblk0 + 2, q));
          }
          if (stamp) a.prof[((iter * kNumFwdLayers + l) * 8) + 3] = clock64();
        }
      }
      if (kEarlyInputs) {
        // the next tile's encoded inputs go into the other input set while fc_9 is still in the tensor pipe
        encode_inputs(in_set[(iter + 1) & 1], 0, nullptr, false);
      } else {
        // the next tile's coordinates: in flight while fc_9 is still in the tensor pipe
        load_xyz(2 * (pair + gridDim.x) + slot, x, y, z);
      }
      // ---- fc_9 output (128 columns = one N-half): this warp owns columns [64 blk, 64 blk + 64)
      {
        const int l = kNumFwdLayers - 1;
        epilogue_wait(&acc_full[slot], n_full & 1);
        ++n_full;
        tc_fence_after();
        if (stamp) a.prof[((iter * kNumFwdLayers + l) * 8) + 2] = clock64();
        if (kTrain) staging_free();
        float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
        uint32_t v0[32], v1[32];
        tmem_ld32(t_slot + 64 * blk, v0);
        tmem_ld32(t_slot + 64 * blk + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        if (kEarlyInputs) mbar_arrive(&drained[slot]);  // the next tile's first MMAs may overwrite the accumulator
        auto head_group = [&](const uint32_t (&v)[32], int gi) {
          const int col0 = 64 * blk + 32 * gi;
          float f[32];
          const uint32_t neg = finish_group<true>(v, sC + kCBias9 + col0, f);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            rgb0 = fmaf(f[i], sC[kCWout + col0 + i], rgb0);
            rgb1 = fmaf(f[i], sC[kCWout + 128 + col0 + i], rgb1);
            rgb2 = fmaf(f[i], sC[kCWout + 256 + col0 + i], rgb2);
          }
          if (to_cache) {
            uint32_t w[16];
            mask_row[(64 + 2 * blk + gi) * kTileM] = neg;
            pack_group(f, w);
            store_words(w, st_row, row, 4 * gi);
          }
        };
        head_group(v0, 0);
        head_group(v1, 1);
        if (blk == 1) sX[slot * 128 + row] = make_float4(rgb0, rgb1, rgb2, sigma_part);
        if (to_cache) bulk_out(cache_tile + cache_slice_off(kCacheH9 + blk, q));
        named_bar_sync(1 + slot, 256);  // the slot's eight warps
        if (blk == 0) {
          const float4 o = sX[slot * 128 + row];
          const float sp = sigma_part + o.w + sC[kCB8_0];
          if (grow < a.m) {
            a.sigma[grow] = fmaxf(sp, 0.f);                                        // nerf.py:115
            a.rgb[3 * grow] = 1.f / (1.f + __expf(-(rgb0 + o.x + sC[kCBout])));    // nerf.py:119
            a.rgb[3 * grow + 1] = 1.f / (1.f + __expf(-(rgb1 + o.y + sC[kCBout + 1])));
            a.rgb[3 * grow + 2] = 1.f / (1.f + __expf(-(rgb2 + o.z + sC[kCBout + 2])));
          }
          if (to_cache) mask_row[kMaskSigmaWord * kTileM] = (grow < a.m && sp > 0.f) ? 1u : 0u;
        }
        if (stamp) a.prof[((iter * kNumFwdLayers + l) * 8) + 3] = clock64();
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
static int g_dbg_store = 0;

}  // namespace nerf

using namespace nerf;

extern "C" int nerf_debug_set_profile_buffer(unsigned long long* buf_dev, int tiles) {
  g_prof_buf = buf_dev;
  g_prof_tiles = tiles & 0xffff;
  g_dbg_store = tiles >> 16;  // debug only: store mode in the high bits
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
  {  // the shared-memory opt-in is a per-device function attribute
    constexpr int kMaxDevices = 64;
    static bool attr_set[kMaxDevices] = {};
    int dev = 0;
    NERF_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices || !attr_set[dev]) {
      NERF_CUDA(cudaFuncSetAttribute(mlp_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
      NERF_CUDA(cudaFuncSetAttribute(mlp_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
      if (dev >= 0 && dev < kMaxDevices) attr_set[dev] = true;
    }
  }
  FwdArgs a;
  a.packed = reinterpret_cast<const uint8_t*>(packed_dev);
  a.pts = pts_dev, a.dirs = dirs_dev, a.ray_o = ray_o_dev, a.ray_d = ray_d_dev, a.t = t_dev, a.s = s, a.m = m;
  a.sigma = sigma_dev, a.rgb = rgb_dev;
  a.cache = reinterpret_cast<uint8_t*>(cache_dev);
  a.prof = g_prof_buf;
  a.prof_tiles = g_prof_tiles;
  a.dbg_store = g_dbg_store;
  const int64_t npairs = (num_tiles(m) + 1) / 2;  // a CTA works on two tiles at a time
  const int grid = (int)((npairs < sm_count()) ? npairs : sm_count());
  if (cache_dev)
    mlp_fwd_kernel<true><<<grid, kFwdThreads, kFwdSmemBytes, as_stream(stream)>>>(a);
  else
    mlp_fwd_kernel<false><<<grid, kFwdThreads, kFwdSmemBytes, as_stream(stream)>>>(a);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}
