// K6 on the sm_100a tensor cores: backward of the NeRF MLP (autograd of network/nerf.py:102-119) in two kernels.
//
//  1. mlp_dgrad_kernel  -- the activation-gradient chain, same structure as the forward chain: per 128-row tile
//        G9 = (g_rgb * rgb(1-rgb)) . W_out  masked by h9>0          (CUDA cores, 3 -> 128)
//        j = 0: G8feat = G9 . W9[:, :256]                           (tcgen05, weights = transposed chunks)
//        j = 1: G7 = (G8feat . W8[1:, :] + g_sigma_pre (x) W8[0, :]) masked by h7>0
//        j = 2..8: G6 .. G0                                          (fc_5 uses only its h4 columns)
//     every G is written to the backward scratch as a tile image for wgrad; ReLU masks come from the bit words
//     the forward pass saved.  {gz, g_sigma_pre} are written per row as one fp32 float4 for the head gradients.
//  2. mlp_wgrad_kernel  -- dW_l = G_l^T . X_l as split-K tcgen05 GEMMs over the saved tile images, both operands
//     MN-major (rows = reduction index); a persistent CTA streams 32-row slices through a 5-stage bulk-copy ring,
//     keeps a (2 x 128) x N fp32 accumulator in TMEM and flushes it with fp32 atomics at segment boundaries;
//     bias gradients are column sums of the G slices taken from shared memory by otherwise idle warps, which also
//     accumulate the two fp32 heads (fc_out, and row 0 of fc_8 = density) from the X slices already in smem.
//
// HBM-bound by design: wgrad must read G and X (2 x 512 B per row and layer); the chain kernels are tensor-bound.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "mlp_tc_layout.cuh"
#include "tc_common.cuh"

namespace nerf {
using namespace tc;

// ================================================================================================
// 1. dgrad chain
// ================================================================================================
constexpr int kDgStages = 9;                     // 16 KB weight chunks (128 input features x 64 output features)
constexpr int kDgEpiWarps = 16;
constexpr int kDgEpiThreads = kDgEpiWarps * 32;
constexpr int kDgMmaWarpB = 2 + kDgEpiWarps;
constexpr int kDgThreads = 32 * (3 + kDgEpiWarps);  // loader, MMA issuer X, 16 epilogue warps, MMA issuer Y
// shared memory map
constexpr int kDgSmC = 0;                        // w8row0 (256) | wout (384)
constexpr int kDgSmBar = kDgSmC + 640 * 4;
constexpr int kDgSmStage = 4096;                 // 16 warps x 4 KB staging slices for the bulk stores
constexpr int kDgSmW = kDgSmStage + 65536;       // weight ring
constexpr int kDgSmTotal = kDgSmW + kDgStages * kChunkBytes;
constexpr int kDgSmemBytes = kDgSmTotal + 1024;
static_assert(kDgSmemBytes <= 232448, "shared memory budget");
// tensor memory map (columns), per slot: fp32 accumulator of one N-half [256 s, +128), A operand (G) [256 s + 128, +128)
constexpr uint32_t kDgTmSlot = 256;
constexpr uint32_t kDgTmA = 128;

struct DgradArgs {
  const uint8_t* packed;
  const uint8_t* cache;   // training cache (masks are read)
  const float* rgb;       // (M,3) forward output
  const float* g_sigma;   // (M)
  const float* g_rgb;     // (M,3)
  uint8_t* scratch;       // gradient tile images + gz + g_sigma_pre
  int64_t m;
  int64_t tile0, tile1;   // this launch covers tiles [tile0, tile1) of the m rows (tile0 even)
};

// `neg` holds the sign bits of the forward pre-activations, column i at bit (31 - i): set = ReLU was inactive
__device__ __forceinline__ void masked_group(const uint32_t (&v)[32], uint32_t neg, float add_scale,
                                             const float* __restrict__ add_vec, float (&f)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float t = __uint_as_float(v[i]);
    if (add_vec != nullptr) t = fmaf(add_scale, add_vec[i], t);
    f[i] = ((neg >> (31 - i)) & 1u) ? 0.f : t;
  }
}

__device__ __forceinline__ void pack_words(const float (&f)[32], uint32_t (&w)[16]) {
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

// Same pipeline as mlp_fwd_kernel: two 128-row tiles ("slots") per CTA so that one tile's MMAs run during the other's
// epilogue, one MMA issuer warp per slot, 16 epilogue warps, weights shared by both slots.
__global__ void __launch_bounds__(kDgThreads, 1) mlp_dgrad_kernel(DgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sStage = smem + kDgSmStage;
  uint8_t* sW = smem + kDgSmW;
  float* sC = reinterpret_cast<float*>(smem + kDgSmC);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDgSmBar);
  uint64_t* full = bars;                        // [kDgStages]
  uint64_t* empty = bars + kDgStages;           // [kDgStages] both slots' MMAs on the chunk have completed
  uint64_t* a_ready = bars + 2 * kDgStages;     // [2] per slot: G operand rewritten in TMEM (input stage + layers 0..7)
  uint64_t* acc_free = a_ready + 2;             // [2] per slot: N-half 0 pulled out of the accumulator
  uint64_t* acc_full = acc_free + 2;            // [2] per slot: the MMAs of one N-half have completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = a.tile1;
  const int64_t pair0 = a.tile0 / 2, npairs = (a.tile1 + 1) / 2;
  {
    const float* cgp = reinterpret_cast<const float*>(a.packed + kPackedConstOff);
    for (int i = threadIdx.x; i < 640; i += kDgThreads) sC[i] = __ldg(cgp + kCW8Row0 + i);  // w8row0 then wout (contiguous)
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kDgStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_ready[i], kDgEpiThreads / 2);   // the slot's eight epilogue warps
      mbar_init(&acc_free[i], kDgEpiThreads / 2);
      mbar_init(&acc_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const float* sW8 = sC;         // fc_8.weight[0, :]
  const float* sWout = sC + 256;  // fc_out.weight (3,128)

  if (warp == 0) {
    const bool leader = elect_one();
    uint32_t g = 0;
    for (int64_t pair = pair0 + blockIdx.x; pair < npairs; pair += gridDim.x) {
      const uint8_t* src = a.packed + kPackedBwdOff;
      for (int c = 0; c < kBwdChunks; ++c) {
        const uint32_t s = g % kDgStages, ph = (g / kDgStages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&full[s], kChunkBytes);
          bulk_g2s(sW + s * kChunkBytes, src, kChunkBytes, &full[s]);
        }
        __syncwarp();
        src += kChunkBytes;
        ++g;
      }
    }
  } else if (warp == 1 || warp == kDgMmaWarpB) {
    // one MMA issuer per slot (see mlp_tc_fwd.cu); each warp runs in lock step, one elected lane issues
    const int slot = (warp == 1) ? 0 : 1;
    const bool leader = elect_one();
    uint32_t g = 0, n_a = 0, n_free = 0;
    constexpr uint32_t idesc = make_idesc_bf16(128, false, false);
    const uint32_t sW_u = smem_u32(sW);
    const uint32_t acc = tmem_base + (uint32_t)slot * kDgTmSlot;
    const uint32_t a_tm = acc + kDgTmA;
    for (int64_t pair = pair0 + blockIdx.x; pair < npairs; pair += gridDim.x) {
      for (int j = 0; j < kNumBwdLayers; ++j) {
        const int nk = bwd_nk(j);
        for (int h = 0; h < 2; ++h) {
          if (h == 0) {  // the layer's G operand is in TMEM (and the accumulator has been drained)
            issuer_wait(&a_ready[slot], n_a & 1);
            ++n_a;
          } else {       // N-half 0 is out of the accumulator
            issuer_wait(&acc_free[slot], n_free & 1);
            ++n_free;
          }
#pragma unroll 1
          for (int kb = 0; kb < nk; ++kb) {
            const uint32_t s = g % kDgStages, ph = (g / kDgStages) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (leader) {
              const uint64_t db = desc_kmajor(sW_u + s * kChunkBytes);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ts(acc, a_tm + (uint32_t)(kb * 32 + k * 8), db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              umma_commit(&empty[s]);
            }
            __syncwarp();
            ++g;
          }
          if (leader) umma_commit(&acc_full[slot]);
          __syncwarp();
        }
      }
    }
  } else {
    // 16 epilogue warps = 4 per TMEM lane quarter, two serving slot X and two slot Y; a warp owns 32 rows x the 64
    // columns [64 blk, +64) of each N-half of its slot = one whole gradient block row (128 B): private 4 KB staging
    // slice, one bulk store per event (see mlp_tc_fwd.cu)
    const int q = warp & 3;
    const int e = (warp - 2) >> 2;
    const int slot = e >> 1, blk = e & 1;
    const int row = q * 32 + lane;
    const uint32_t t_slot = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)slot * kDgTmSlot;
    uint8_t* st_row = sStage + (warp - 2) * 4096 + lane * 128;
    uint32_t n_full = 0;
    float4* ghead_out = reinterpret_cast<float4*>(a.scratch + scratch_ghead_offset(a.m));

    auto bulk_out = [&](uint8_t* gdst) {  // the staged 32 rows x 128 B -> HBM
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

    // head gradients and the h9 ReLU masks of this thread's row of a tile (loaded one tile ahead)
    struct HeadIn {
      float gz0, gz1, gz2, gsp;
      uint32_t mk[2];
    } nx;
    auto load_heads = [&](int64_t tile_) {
      nx.gz0 = nx.gz1 = nx.gz2 = nx.gsp = 0.f;
      nx.mk[0] = nx.mk[1] = 0u;
      const int64_t grow_ = tile_ * kTileM + row;
      if (tile_ < ntiles) {
        const uint32_t* mrow =
            reinterpret_cast<const uint32_t*>(a.cache + cache_mask_offset(a.m) + (size_t)tile_ * kMaskTileBytes) + row;
        nx.mk[0] = __ldg(mrow + (64 + 2 * blk) * kTileM);
        nx.mk[1] = __ldg(mrow + (64 + 2 * blk + 1) * kTileM);
        if (grow_ < a.m) {
          const float r0 = __ldg(a.rgb + 3 * grow_), r1 = __ldg(a.rgb + 3 * grow_ + 1), r2 = __ldg(a.rgb + 3 * grow_ + 2);
          nx.gz0 = __ldg(a.g_rgb + 3 * grow_) * r0 * (1.f - r0);
          nx.gz1 = __ldg(a.g_rgb + 3 * grow_ + 1) * r1 * (1.f - r1);
          nx.gz2 = __ldg(a.g_rgb + 3 * grow_ + 2) * r2 * (1.f - r2);
          const uint32_t smask = __ldg(mrow + kMaskSigmaWord * kTileM);
          nx.gsp = (smask & 1u) ? __ldg(a.g_sigma + grow_) : 0.f;
        }
      }
    };
    load_heads(2 * (pair0 + (int64_t)blockIdx.x) + slot);

    for (int64_t pair = pair0 + blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int64_t tile = 2 * pair + slot;
      const bool valid = tile < ntiles;
      uint8_t* g_tile = a.scratch + (size_t)tile * kGradTileBytes;
      const uint32_t* mask_row =
          reinterpret_cast<const uint32_t*>(a.cache + cache_mask_offset(a.m) + (size_t)tile * kMaskTileBytes) + row;
      // ---- input stage: heads on CUDA cores: gz = g_rgb * rgb (1 - rgb) (sigmoid backward),
      //      g_sigma_pre = g_sigma * (sigma_pre > 0); G9 = (gz . W_out) masked by h9 > 0 into TMEM.
      //      (the loads were issued during the previous tile's last layers)
      const float gz0 = nx.gz0, gz1 = nx.gz1, gz2 = nx.gz2, gsp = nx.gsp;
      const uint32_t mk9[2] = {nx.mk[0], nx.mk[1]};
      if (blk == 0 && valid) ghead_out[tile * kTileM + row] = make_float4(gz0, gz1, gz2, gsp);
      staging_free();
#pragma unroll
      for (int gi = 0; gi < 2; ++gi) {
        const int col0 = 64 * blk + 32 * gi;
        const uint32_t mk = mk9[gi];
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float t = gz0 * sWout[col0 + i];
          t = fmaf(gz1, sWout[128 + col0 + i], t);
          t = fmaf(gz2, sWout[256 + col0 + i], t);
          f[i] = ((mk >> (31 - i)) & 1u) ? 0.f : t;
        }
        uint32_t w[16];
        pack_words(f, w);
        tmem_st16(t_slot + kDgTmA + 32 * blk + 16 * gi, w);  // layer 0 reads k-blocks 0, 1
        if (valid) store_words(w, st_row, row, 4 * gi);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&a_ready[slot]);  // also: this thread has drained the slot's previous tile
      if (valid) bulk_out(g_tile + grad_slice_off(kGradG9 + blk, q));

      uint32_t wh[2][16];  // bf16 pairs of N-half 0, held until half 1's MMAs have stopped reading the G operand
#pragma unroll 1
      for (int j = 0; j < kNumBwdLayers; ++j) {
        if (j == kNumBwdLayers - 2) load_heads(2 * (pair + gridDim.x) + slot);  // in flight during the last two layers
        const int slot_m = 8 - j;  // ReLU mask of the layer output this gradient flows into (j >= 1): h7 .. h0
        // sign-bit masks of this thread's four 32-column groups (0 = every column passes: layer j = 0 has no ReLU)
        uint32_t mk[4] = {0u, 0u, 0u, 0u};
        if (j >= 1 && valid) {
#pragma unroll
          for (int i = 0; i < 4; ++i) mk[i] = __ldg(mask_row + (slot_m * 8 + 4 * (i >> 1) + 2 * blk + (i & 1)) * kTileM);
        }
        auto finish = [&](const uint32_t (&v)[32], uint32_t m, int col0, uint32_t (&w)[16]) {
          float f[32];
          masked_group(v, m, gsp, j == 1 ? sW8 + col0 : nullptr, f);
          pack_words(f, w);
        };
        uint8_t* gblk0 = g_tile + grad_slice_off(2 + 4 * j + blk, q);  // gradient block of N-half 0's columns
        // ---- N-half 0: pull it out of the accumulator first (the issuer is waiting for that), then convert and hold it
        {
          epilogue_wait(&acc_full[slot], n_full & 1);
          ++n_full;
          tc_fence_after();
          uint32_t v0[32], v1[32];
          tmem_ld32(t_slot + 64 * blk, v0);
          tmem_ld32(t_slot + 64 * blk + 32, v1);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&acc_free[slot]);
          finish(v0, mk[0], 64 * blk, wh[0]);
          finish(v1, mk[1], 64 * blk + 32, wh[1]);
          if (valid) {
            staging_free();
            store_words(wh[0], st_row, row, 0);
            store_words(wh[1], st_row, row, 4);
            bulk_out(gblk0);
          }
        }
        // ---- N-half 1: once complete nothing reads the G operand any more -> overwrite it in place, signal the
        //      issuer, then store the gradient block for wgrad
        {
          epilogue_wait(&acc_full[slot], n_full & 1);
          ++n_full;
          tc_fence_after();
          uint32_t v[32], w0[16], w1[16];
          tmem_ld32(t_slot + 64 * blk, v);
          tmem_ld_wait();
          finish(v, mk[2], 128 + 64 * blk, w0);
          if (j < kNumBwdLayers - 1) {
            tmem_st16(t_slot + kDgTmA + 32 * blk, wh[0]);
            tmem_st16(t_slot + kDgTmA + 64 + 32 * blk, w0);
          }
          tmem_ld32(t_slot + 64 * blk + 32, v);
          tmem_ld_wait();
          finish(v, mk[3], 128 + 64 * blk + 32, w1);
          if (j < kNumBwdLayers - 1) {
            tmem_st16(t_slot + kDgTmA + 32 * blk + 16, wh[1]);
            tmem_st16(t_slot + kDgTmA + 64 + 32 * blk + 16, w1);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&a_ready[slot]);
          } else {
            tc_fence_before();
          }
          if (valid) {
            staging_free();
            store_words(w0, st_row, row, 0);
            store_words(w1, st_row, row, 4);
            bulk_out(gblk0 + 2 * kSliceBytes);  // block + 2 of the same slice
          }
        }
      }
    }
    if (lane == 0) bulk_wait_all<0>();
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================
// 2. wgrad
// ================================================================================================
struct WUnit {
  int g_blk0;      // first G block of the unit in the gradient tile image
  int n_gblk;      // 0 (head-only unit), 2 or 4 (64 output features each)
  int x_blk0;      // first X block in the cache tile image (contiguous blocks)
  int n_xblk;      // full 64-column X blocks
  int x_extra;     // extra narrow block (view-direction encoding, 32 columns) right after the full blocks, or -1
  int n_xload;     // X-side blocks fetched per slice, contiguous from x_blk0 (MMA blocks, then blocks only the head reads)
  int head_xblk;   // first X-side block (relative to x_blk0) the head sums read
  int param_w;     // weight gradient slot
  int w_row0;      // first output row in the weight tensor
  int w_col0;      // first input column
  int w_ld;        // weight leading dimension
  int valid_cols;  // accumulator columns that map to real weight columns
  int param_b;     // bias slot or -1
  int b_off;
  int cost;        // relative time per tile for the work partition (blocks streamed; more for CUDA-core-only units)
  int head;        // 0: none; 1: density head (d fc_8.weight[0,:], d fc_8.bias[0]) from g_sigma_pre and the unit's X blocks
                   // 2: fc_out (d fc_out.weight, d fc_out.bias) from gz and the two h9 blocks fetched after the MMA blocks
};
constexpr int kNumWUnits = 11;
__constant__ WUnit c_wunits[kNumWUnits];

// The ring holds as many stages as fit: one stage = the 32-row slices of every block a unit streams (+ 512 B of head
// gradients), so units that stream few bytes per slice get a deeper ring and keep the same number of bytes in flight.
constexpr int kWgSlice = 4096;             // 32 rows of one block
constexpr int kWgRingBytes = 188 * 1024;
constexpr int kWgMaxStages = 16;
constexpr int kWgCudaWarps = 8;            // bias / head sums and the accumulator flush
constexpr int kWgThreads = 32 * (2 + kWgCudaWarps);
constexpr int kWgSmBar = kWgRingBytes;
constexpr int kWgSmemBytes = kWgSmBar + 512 + 1024;
static_assert(kWgSmemBytes <= 232448, "shared memory budget");

struct WgradArgs {
  const uint8_t* cache;
  const uint8_t* scratch;
  ParamPtrs grads;
  int64_t m;
  int64_t tile0, tile1;      // this launch covers tiles [tile0, tile1) of the m rows
  int64_t wrap_g, wrap_x;    // timing experiment only (results unusable): read G / X of tile (t % wrap) instead of t; 0 = off
  unsigned long long* prof;  // optional per-CTA timeline (nerf_debug_set_wgrad_profile): 8 values per CTA
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct Segment {
  int unit;
  int64_t tile0, tile1;
};

// cost-balanced contiguous partition of (unit, tile) pairs over the grid
__device__ __forceinline__ int unit_cost(int u) { return c_wunits[u].cost; }

__device__ inline int build_segments(int64_t tile_first, int64_t ntiles, Segment* seg) {
  int64_t total = 0;
  for (int u = 0; u < kNumWUnits; ++u) total += ntiles * unit_cost(u);
  const int64_t lo = total * blockIdx.x / gridDim.x, hi = total * (blockIdx.x + 1) / gridDim.x;
  int n = 0;
  int64_t off = 0;
  for (int u = 0; u < kNumWUnits; ++u) {
    const int c = unit_cost(u);
    const int64_t span = ntiles * c;
    int64_t a = lo - off, b = hi - off;
    if (a < 0) a = 0;
    if (b > span) b = span;
    if (b > a) {
      const int64_t t0 = (a + c - 1) / c, t1 = (b + c - 1) / c;
      if (t1 > t0) {
        seg[n].unit = u;
        seg[n].tile0 = tile_first + t0;
        seg[n].tile1 = tile_first + (t1 < ntiles ? t1 : ntiles);
        ++n;
      }
    }
    off += span;
  }
  return n;
}

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// 8 consecutive bf16 columns (one 16-byte chunk of a tile-image row)
__device__ __forceinline__ void unpack8(const uint4& c, float (&x)[8]) {
  x[0] = bf16_lo(c.x), x[1] = bf16_hi(c.x), x[2] = bf16_lo(c.y), x[3] = bf16_hi(c.y);
  x[4] = bf16_lo(c.z), x[5] = bf16_hi(c.z), x[6] = bf16_lo(c.w), x[7] = bf16_hi(c.w);
}

// column sums of ROWS rows (r0 ..) of the 16-byte column group `cg` of a staged operand (blocks kWgSlice apart)
template <int ROWS>
__device__ __forceinline__ void col_sums(const uint8_t* blocks, int cg, int r0, float (&acc)[8]) {
  uint4 c[ROWS];
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr) {
    const int r = r0 + rr;
    c[rr] = *reinterpret_cast<const uint4*>(blocks + (cg >> 3) * 4096 + r * 128 + (((cg ^ r) & 7) << 4));
  }
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr) {
    float x[8];
    unpack8(c[rr], x);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += x[i];
  }
}

// head weight gradients: acc[j][i] += g_j[row] * x[row][8 cg + i]; NH = 1 takes g = gh[row].w (density head),
// NH = 3 takes {x, y, z} (fc_out)
template <int ROWS, int NH>
__device__ __forceinline__ void head_sums(const uint8_t* blocks, const float4* gh, int cg, int r0, float (&acc)[3][8]) {
  uint4 c[ROWS];
  float4 gv[ROWS];
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr) {
    const int r = r0 + rr;
    c[rr] = *reinterpret_cast<const uint4*>(blocks + (cg >> 3) * 4096 + r * 128 + (((cg ^ r) & 7) << 4));
    gv[rr] = gh[r];
  }
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr) {
    float x[8];
    unpack8(c[rr], x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (NH == 1) {
        acc[0][i] = fmaf(gv[rr].w, x[i], acc[0][i]);
      } else {
        acc[0][i] = fmaf(gv[rr].x, x[i], acc[0][i]);
        acc[1][i] = fmaf(gv[rr].y, x[i], acc[1][i]);
        acc[2][i] = fmaf(gv[rr].z, x[i], acc[2][i]);
      }
    }
  }
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1) mlp_wgrad_kernel(WgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgSmBar);
  uint64_t* full = bars;                      // [kWgMaxStages]
  uint64_t* empty = bars + kWgMaxStages;      // [kWgMaxStages]
  uint64_t* acc_done = bars + 2 * kWgMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);
  __shared__ Segment segs[kNumWUnits];
  __shared__ int nseg_s;
  __shared__ uint32_t stage_base[kWgMaxStages];  // completed phases of every stage's barriers (segments change the ring geometry)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = a.tile1 - a.tile0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgMaxStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1 + kWgCudaWarps);  // MMA commit + the CUDA-core warps
    }
    mbar_init(acc_done, 1);
    fence_barrier_init();
    for (int i = 0; i < kWgMaxStages; ++i) stage_base[i] = 0;
    nseg_s = build_segments(a.tile0, ntiles, segs);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nseg = nseg_s;
  if (a.prof != nullptr && threadIdx.x == 64) {
    unsigned long long* pr = a.prof + blockIdx.x * 16;
    pr[0] = global_ns();
    pr[3] = (unsigned long long)segs[0].unit;
    pr[4] = (unsigned long long)nseg;
    pr[5] = (unsigned long long)(segs[0].tile1 - segs[0].tile0);
    pr[6] = nseg > 1 ? (unsigned long long)(segs[1].tile1 - segs[1].tile0) : 0ull;
  }

  for (int si = 0; si < nseg; ++si) {
    const WUnit u = c_wunits[segs[si].unit];
    const int64_t tile0 = segs[si].tile0, tile1 = segs[si].tile1;
    const int nxb = u.n_xload;
    const uint32_t head_off = (uint32_t)(u.n_gblk + nxb) * kWgSlice;
    const uint32_t bytes = head_off + (u.head ? 512u : 0u);
    const uint32_t stage_bytes = (bytes + 1023u) & ~1023u;
    const uint32_t nst = min((uint32_t)kWgMaxStages, (uint32_t)kWgRingBytes / stage_bytes);
    uint32_t phase_bits0 = 0;  // parity of the next phase of every stage's barriers at the start of this segment
#pragma unroll
    for (int i = 0; i < kWgMaxStages; ++i) phase_bits0 |= (stage_base[i] & 1u) << i;

    if (warp == 0) {
      // ---------------------------------------------------------------- loader: 32-row slices of every block
      const bool leader = elect_one();
      uint32_t g = 0, s = 0, phase_bits = phase_bits0;  // stage cursor; bit i = parity of stage i's next phase
      long long ld_wait = 0;
      const long long ld_t0 = a.prof ? clock64() : 0;
      const uint8_t* ghead = a.scratch + scratch_ghead_offset(a.m);
      for (int64_t tile = tile0; tile < tile1; ++tile) {
        const uint8_t* gtile = a.scratch + (size_t)(a.wrap_g ? tile % a.wrap_g : tile) * kGradTileBytes;
        const uint8_t* xtile = a.cache + (size_t)(a.wrap_x ? tile % a.wrap_x : tile) * kCacheTileBytes;
        for (int sl = 0; sl < 4; ++sl) {
          const uint32_t ph = (phase_bits >> s) & 1u;
          const long long w0 = a.prof ? clock64() : 0;
          mbar_wait(&empty[s], ph ^ 1);
          if (a.prof) ld_wait += clock64() - w0;
          if (leader) {
            // slice-major tiles: the pieces of consecutive blocks of one slice are contiguous -> one copy per operand
            mbar_arrive_expect_tx(&full[s], bytes);
            uint8_t* dst = smem + s * stage_bytes;
            if (u.n_gblk) bulk_g2s(dst, gtile + grad_slice_off(u.g_blk0, sl), (uint32_t)u.n_gblk * kWgSlice, &full[s]);
            uint8_t* xdst = dst + u.n_gblk * kWgSlice;
            bulk_g2s(xdst, xtile + cache_slice_off(u.x_blk0, sl), (uint32_t)u.n_xload * kWgSlice, &full[s]);
            if (u.head) bulk_g2s(dst + head_off, ghead + ((size_t)tile * kTileM + sl * 32) * 16, 512, &full[s]);
          }
          __syncwarp();
          phase_bits ^= 1u << s;
          s = (s + 1 == nst) ? 0u : s + 1;
          ++g;
        }
      }
      if (a.prof && leader && si == 0) {
        a.prof[blockIdx.x * 16 + 8] = (unsigned long long)ld_wait;
        a.prof[blockIdx.x * 16 + 9] = (unsigned long long)(clock64() - ld_t0);
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer (MN-major operands)
      const bool leader = elect_one();
      uint32_t g = 0, s = 0, phase_bits = phase_bits0;  // stage cursor; bit i = parity of stage i's next phase
      const uint32_t sm_u = smem_u32(smem);
      const int nhalf = u.n_gblk / 2;
      const uint32_t n_main = (uint32_t)u.n_xblk * 64u;
      const uint32_t idesc_main = make_idesc_bf16(n_main, true, true);
      const uint32_t idesc_extra = make_idesc_bf16(32, true, true);
      bool first = true;
      long long mma_wait = 0;
      const long long mma_t0 = a.prof ? clock64() : 0;
      for (int64_t tile = tile0; tile < tile1; ++tile) {
        for (int sl = 0; sl < 4; ++sl) {
          const uint32_t ph = (phase_bits >> s) & 1u;
          const long long w0 = a.prof ? clock64() : 0;
          mbar_wait(&full[s], ph);
          if (a.prof) mma_wait += clock64() - w0;
          tc_fence_after();
          const uint32_t st = sm_u + s * stage_bytes;
          const uint32_t xb = st + u.n_gblk * kWgSlice;
          if (leader) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              for (int h = 0; h < nhalf; ++h) {
                const uint64_t da = desc_mnmajor(st + (2 * h) * kWgSlice + k * 2048, kWgSlice);
                const uint64_t db = desc_mnmajor(xb + k * 2048, kWgSlice);
                umma_bf16(tmem_base + h * 256, da, db, idesc_main, (first && k == 0) ? 0u : 1u);
                if (u.x_extra >= 0) {
                  const uint64_t de = desc_mnmajor(xb + u.n_xblk * kWgSlice + k * 2048, kWgSlice);
                  umma_bf16(tmem_base + h * 256 + n_main, da, de, idesc_extra, (first && k == 0) ? 0u : 1u);
                }
              }
            }
            umma_commit(&empty[s]);
          }
          first = false;
          __syncwarp();
          phase_bits ^= 1u << s;
          s = (s + 1 == nst) ? 0u : s + 1;
          ++g;
        }
      }
      if (leader) umma_commit(acc_done);
      __syncwarp();
      if (a.prof && leader && si == 0) {
        a.prof[blockIdx.x * 16 + 10] = (unsigned long long)mma_wait;
        a.prof[blockIdx.x * 16 + 11] = (unsigned long long)(clock64() - mma_t0);
      }
    } else {
      // ---------------------------------------------------------------- bias / head sums on CUDA cores + accumulator flush
      // 256 threads; thread t owns one 16-byte column group (8 bf16 columns) and a share of the slice's 32 rows
      const int t = threadIdx.x - 64;
      const int b_ncg = u.param_b >= 0 ? u.n_gblk * 8 : 0;  // column groups of the bias sums over G: 32, 16 or none
      const int h_ncg = u.head == 1 ? 32 : (u.head == 2 ? 16 : 0);  // column groups of the head sums: density over 4 X blocks, fc_out over h9
      float bacc[8], hacc[3][8];
#pragma unroll
      for (int i = 0; i < 8; ++i) bacc[i] = hacc[0][i] = hacc[1][i] = hacc[2][i] = 0.f;
      float hb = 0.f;  // head bias sums (threads 0..127: one row and component each)
      const int b_cg = b_ncg ? t % b_ncg : 0, b_rows = b_ncg / 8, b_r0 = b_ncg ? (t / b_ncg) * b_rows : 0;
      const int h_cg = h_ncg ? t % h_ncg : 0, h_rows = h_ncg / 8, h_r0 = h_ncg ? (t / h_ncg) * h_rows : 0;
      uint32_t g = 0, s = 0, phase_bits = phase_bits0;  // stage cursor; bit i = parity of stage i's next phase
      long long cc_wait = 0;
      const long long cc_t0 = a.prof ? clock64() : 0;
      for (int64_t tile = tile0; tile < tile1; ++tile) {
        for (int sl = 0; sl < 4; ++sl) {
          const uint32_t ph = (phase_bits >> s) & 1u;
          const long long w0 = a.prof ? clock64() : 0;
          mbar_wait(&full[s], ph);
          if (a.prof) cc_wait += clock64() - w0;
          const uint8_t* st = smem + s * stage_bytes;
          if (b_ncg == 32) col_sums<4>(st, b_cg, b_r0, bacc);
          else if (b_ncg == 16) col_sums<2>(st, b_cg, b_r0, bacc);
          if (u.head) {
            const float4* gh = reinterpret_cast<const float4*>(st + head_off);
            const uint8_t* xs = st + (u.n_gblk + u.head_xblk) * kWgSlice;
            if (u.head == 1) head_sums<4, 1>(xs, gh, h_cg, h_r0, hacc);  // 4 X blocks
            else head_sums<2, 3>(xs, gh, h_cg, h_r0, hacc);              // 2 X blocks
            if (t < 128) hb += reinterpret_cast<const float*>(gh)[t];  // row t/4, component t%4 of {gz0, gz1, gz2, g_sigma_pre}
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[s]);
          phase_bits ^= 1u << s;
          s = (s + 1 == nst) ? 0u : s + 1;
          ++g;
        }
      }
      if (a.prof && threadIdx.x == 64 && si == 0) {
        a.prof[blockIdx.x * 16 + 12] = (unsigned long long)cc_wait;
        a.prof[blockIdx.x * 16 + 13] = (unsigned long long)(clock64() - cc_t0);
        a.prof[blockIdx.x * 16 + 14] = (unsigned long long)g;
      }
      if (b_ncg) {
        float* db = a.grads.p[u.param_b] + u.b_off + 8 * b_cg;
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(db + i, bacc[i]);
      }
      if (u.head) {  // lanes with equal (lane & 3) hold the same head-gradient component
        hb += __shfl_xor_sync(0xffffffffu, hb, 4);
        hb += __shfl_xor_sync(0xffffffffu, hb, 8);
        hb += __shfl_xor_sync(0xffffffffu, hb, 16);
      }
      if (u.head == 1) {
        float* dw = a.grads.p[W_8] + 8 * h_cg;  // d fc_8.weight[0, :]
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(dw + i, hacc[0][i]);
        if (t < 128 && lane == 3) atomicAdd(a.grads.p[B_8], hb);  // d fc_8.bias[0] = sum g_sigma_pre
      } else if (u.head == 2) {
        float* dw = a.grads.p[W_OUT] + 8 * h_cg;
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int i = 0; i < 8; ++i) atomicAdd(dw + j * kH + i, hacc[j][i]);
        if (t < 128 && lane < 3) atomicAdd(a.grads.p[B_OUT] + lane, hb);
      }
      // flush the accumulator of this segment: two warps per TMEM lane quarter alternate over the 32-column chunks
      mbar_wait(acc_done, (uint32_t)si & 1);
      tc_fence_after();
      if (a.prof != nullptr && threadIdx.x == 64 && si == nseg - 1) a.prof[blockIdx.x * 16 + 1] = global_ns();
      const int q = warp & 3, wsel = (warp - 2) >> 2;
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      const int nhalf = u.n_gblk / 2;
      const int ncols = u.n_xblk * 64 + (u.x_extra >= 0 ? 32 : 0);
      const bool vec = ((u.w_ld & 3) == 0) && ((u.w_col0 & 3) == 0);
      float* dw = a.grads.p[u.param_w];
      int item = 0;
      for (int h = 0; h < nhalf; ++h) {
        const int out_row = u.w_row0 + h * 128 + q * 32 + lane;
        float* dst = dw + (size_t)out_row * u.w_ld + u.w_col0;
        for (int c0 = 0; c0 < ncols; c0 += 32, ++item) {
          if ((item & 1) != wsel) continue;
          uint32_t v[32];
          tmem_ld32(lane_addr + h * 256 + c0, v);
          tmem_ld_wait();
          if (vec && c0 + 32 <= u.valid_cols) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              red_add_v4(dst + c0 + i, __uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                         __uint_as_float(v[i + 3]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c0 + i < u.valid_cols) atomicAdd(dst + c0 + i, __uint_as_float(v[i]));
          }
        }
      }
      tc_fence_before();
    }
    // ---- segment boundary: the next unit uses another ring geometry.  The ring is drained here (every stage consumed
    //      and released); the barriers keep their state, only the number of phases each one has completed is recorded
    __syncthreads();
    if (si + 1 < nseg) {
      if (threadIdx.x == 0) {
        const uint32_t total = (uint32_t)(tile1 - tile0) * 4u;  // stages streamed in this segment
        for (uint32_t i = 0; i < nst; ++i)
          if (total > i) stage_base[i] += (total - i + nst - 1) / nst;
      }
      __syncthreads();
      tc_fence_after();
    }
  }
  if (a.prof != nullptr && threadIdx.x == 64) a.prof[blockIdx.x * 16 + 2] = global_ns();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

static void build_wunits(WUnit* u) {
  int n = 0;
  // cost = relative time per tile for the work partition: blocks streamed per slice, corrected by measurement
  // (tools/prof_wgrad.py) for the units whose slices are short or that carry CUDA-core head sums
  auto add = [&](int g0, int ng, int x0, int nx, int xe, int nload, int pw, int row0, int col0, int ld, int valid, int pb,
                 int boff, int head, int head_xblk, int cost) {
    u[n++] = WUnit{g0, ng, x0, nx, xe, nload, head_xblk, pw, row0, col0, ld, valid, pb, boff, cost, head};
  };
  add(grad_g(0), 4, kCachePe, 1, -1, 1, W_IN, 0, 0, kP, kP, B_IN, 0, 0, 0, 64);                      // fc_in : G0 x pe
  for (int l = 1; l <= 4; ++l) add(grad_g(l), 4, cache_h(l - 1), 4, -1, 4, 2 * l, 0, 0, kF, kF, 2 * l + 1, 0, 0, 0, 80);  // fc_1..4
  add(grad_g(5), 4, kCachePe, 1, -1, 1, W_5, 0, 0, kP + kF, kP, B_5, 0, 0, 0, 64);                   // fc_5, position columns
  add(grad_g(5), 4, cache_h(4), 4, -1, 4, W_5, 0, kP, kP + kF, kF, -1, 0, 0, 0, 80);                 // fc_5, h4 columns
  add(grad_g(6), 4, cache_h(5), 4, -1, 4, W_6, 0, 0, kF, kF, B_6, 0, 0, 0, 80);
  add(grad_g(7), 4, cache_h(6), 4, -1, 4, W_7, 0, 0, kF, kF, B_7, 0, 0, 0, 80);
  add(kGradG8, 4, cache_h(7), 4, -1, 4, W_8, 1, 0, kF, kF, B_8, 1, 1, 0, 88);                        // fc_8 rows 1..256 + density row
  // fc_9 : G9 x [feat | de], and fc_out from gz x h9 on the CUDA cores (feat, de, h9 are consecutive cache blocks)
  static_assert(kCacheDe == kCacheFeat + 4 && kCacheH9 == kCacheDe + 1, "fc_9 unit fetches feat, de, h9 with one copy");
  add(kGradG9, 2, kCacheFeat, 4, kCacheDe, 7, W_9, 0, 0, kF + kV, kF + kV, B_9, 0, 2, 5, 84);
}

__global__ void zero_grads_kernel(ParamPtrs g) {
  const int sizes[22] = {kF * kP, kF, kF * kF, kF, kF * kF, kF, kF * kF, kF, kF * kF, kF, kF * (kP + kF), kF,
                         kF * kF, kF, kF * kF, kF, (kF + 1) * kF, kF + 1, kH * (kF + kV), kH, 3 * kH, 3};
  float* p = g.p[blockIdx.y];
  const int n = sizes[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0.f;
}

static unsigned long long* g_wgrad_prof = nullptr;
static int64_t g_wgrad_wrap_g = 0, g_wgrad_wrap_x = 0;
static int g_wgrad_cost_override[kNumWUnits] = {};  // > 0: replaces the unit's cost (tuning aid, nerf_debug_set_wgrad_costs)

// Per-device one-time setup (shared-memory opt-in of the kernels, the wgrad unit table in constant memory).  A process
// that drives several GPUs must do this on each of them, hence the per-device flags.
constexpr int kMaxDevices = 64;
static bool g_bwd_ready[kMaxDevices] = {};

static int bwd_device_setup() {
  int dev = 0;
  NERF_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < kMaxDevices && g_bwd_ready[dev]) return NERF_OK;
  NERF_CUDA(cudaFuncSetAttribute(mlp_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDgSmemBytes));
  NERF_CUDA(cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes));
  WUnit table[kNumWUnits];
  build_wunits(table);
  for (int u = 0; u < kNumWUnits; ++u)
    if (g_wgrad_cost_override[u] > 0) table[u].cost = g_wgrad_cost_override[u];
  NERF_CUDA(cudaMemcpyToSymbol(c_wunits, table, sizeof(table)));
  if (dev >= 0 && dev < kMaxDevices) g_bwd_ready[dev] = true;
  return NERF_OK;
}

// phases: bit 0 zero the gradients, bit 1 dgrad chain, bit 2 wgrad; tiles [tile0, tile1); ctas = grid size (0 = one per SM)
static int launch_backward(const void* packed_dev, const void* cache_dev, const float* rgb_dev, int64_t m,
                           const float* g_sigma_dev, const float* g_rgb_dev, float* const* grads, void* scratch_dev,
                           int phases, int64_t tile0, int64_t tile1, int ctas, cudaStream_t st) {
  if (int rc = bwd_device_setup()) return rc;
  ParamPtrs gp;
  for (int i = 0; i < NERF_NUM_PARAM_TENSORS; ++i) {
    NERF_CHECK_ARG(grads[i] != nullptr, "nerf_mlp_bf16_backward: null gradient pointer");
    gp.p[i] = grads[i];
  }
  if (phases & 1) {
    zero_grads_kernel<<<dim3(32, 22), 256, 0, st>>>(gp);
    NERF_LAUNCH_CHECK();
  }
  const int sms = (ctas > 0 && ctas < sm_count()) ? ctas : sm_count();
  if (tile1 <= tile0) return NERF_OK;
  if (phases & 2) {
    DgradArgs a;
    a.packed = reinterpret_cast<const uint8_t*>(packed_dev);
    a.cache = reinterpret_cast<const uint8_t*>(cache_dev);
    a.rgb = rgb_dev, a.g_sigma = g_sigma_dev, a.g_rgb = g_rgb_dev;
    a.scratch = reinterpret_cast<uint8_t*>(scratch_dev);
    a.m = m;
    a.tile0 = tile0, a.tile1 = tile1;
    const int64_t npairs = (tile1 + 1) / 2 - tile0 / 2;  // a CTA works on two tiles at a time
    const int grid = (int)(npairs < sms ? npairs : sms);
    mlp_dgrad_kernel<<<grid, kDgThreads, kDgSmemBytes, st>>>(a);
    NERF_LAUNCH_CHECK();
  }
  if (phases & 4) {
    WgradArgs a;
    a.cache = reinterpret_cast<const uint8_t*>(cache_dev);
    a.scratch = reinterpret_cast<const uint8_t*>(scratch_dev);
    a.grads = gp;
    a.m = m;
    a.tile0 = tile0, a.tile1 = tile1;
    a.wrap_g = g_wgrad_wrap_g, a.wrap_x = g_wgrad_wrap_x;
    a.prof = g_wgrad_prof;
    mlp_wgrad_kernel<<<sms, kWgThreads, kWgSmemBytes, st>>>(a);
    NERF_LAUNCH_CHECK();
  }
  return NERF_OK;
}

}  // namespace nerf

using namespace nerf;

extern "C" int nerf_debug_set_wgrad_profile(unsigned long long* buf_dev) {
  g_wgrad_prof = buf_dev;
  return NERF_OK;
}

extern "C" int nerf_debug_set_wgrad_costs(const int* costs, int n) {
  for (int u = 0; u < kNumWUnits; ++u) g_wgrad_cost_override[u] = (costs != nullptr && u < n) ? costs[u] : 0;
  for (int d = 0; d < kMaxDevices; ++d) g_bwd_ready[d] = false;  // re-upload the unit table on the next launch
  return NERF_OK;
}

extern "C" int nerf_debug_set_wgrad_wrap(int64_t wrap_g, int64_t wrap_x) {
  g_wgrad_wrap_g = wrap_g, g_wgrad_wrap_x = wrap_x;
  return NERF_OK;
}

extern "C" int nerf_mlp_bf16_backward(const void* packed_dev, const void* cache_dev, const float* rgb_dev, int64_t m,
                                      const float* g_sigma_dev, const float* g_rgb_dev, float* const* grads,
                                      void* scratch_dev, nerf_stream_t stream) {
  NERF_CHECK_ARG(m > 0, "nerf_mlp_bf16_backward: row count must be positive");
  NERF_CHECK_ARG(packed_dev && cache_dev && rgb_dev && g_sigma_dev && g_rgb_dev && grads && scratch_dev,
                 "nerf_mlp_bf16_backward: null pointer");
  return launch_backward(packed_dev, cache_dev, rgb_dev, m, g_sigma_dev, g_rgb_dev, grads, scratch_dev, 7, 0, num_tiles(m), 0,
                         as_stream(stream));
}

extern "C" int nerf_mlp_bf16_backward_part(const void* packed_dev, const void* cache_dev, const float* rgb_dev, int64_t m,
                                           const float* g_sigma_dev, const float* g_rgb_dev, float* const* grads,
                                           void* scratch_dev, int phases, int64_t tile0, int64_t tile1, int ctas,
                                           nerf_stream_t stream) {
  NERF_CHECK_ARG(m > 0, "nerf_mlp_bf16_backward_part: row count must be positive");
  NERF_CHECK_ARG(packed_dev && cache_dev && rgb_dev && g_sigma_dev && g_rgb_dev && grads && scratch_dev,
                 "nerf_mlp_bf16_backward_part: null pointer");
  NERF_CHECK_ARG((phases & ~7) == 0, "nerf_mlp_bf16_backward_part: phases is a mask of bits 0..2");
  NERF_CHECK_ARG(tile0 >= 0 && (tile0 & 1) == 0 && tile1 <= num_tiles(m) && tile0 <= tile1,
                 "nerf_mlp_bf16_backward_part: need 0 <= tile0 <= tile1 <= tiles(m) with tile0 even");
  NERF_CHECK_ARG(ctas >= 0, "nerf_mlp_bf16_backward_part: negative CTA count");
  return launch_backward(packed_dev, cache_dev, rgb_dev, m, g_sigma_dev, g_rgb_dev, grads, scratch_dev, phases, tile0, tile1,
                         ctas, as_stream(stream));
}
