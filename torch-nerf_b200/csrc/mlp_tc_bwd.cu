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
constexpr int kDgStages = 10;                    // 16 KB weight chunks (128 input features x 64 output features)
constexpr int kDgEpiWarps = 16;
constexpr int kDgEpiThreads = kDgEpiWarps * 32;
constexpr int kDgMmaWarpB = 2 + kDgEpiWarps;
constexpr int kDgThreads = 32 * (3 + kDgEpiWarps);  // loader, MMA issuer X, 16 epilogue warps, MMA issuer Y
// shared memory map
constexpr int kDgSmC = 0;                        // w8row0 (256) | wout (384)
constexpr int kDgSmBar = kDgSmC + 640 * 4;
constexpr int kDgSmStage = 4096;                 // 8 warp pairs x 4 KB staging slices for the bulk stores
constexpr int kDgSmW = kDgSmStage + 32768;       // weight ring
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
  const int64_t ntiles = num_tiles(a.m);
  const int64_t npairs = (ntiles + 1) / 2;
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
      mbar_init(&a_ready[i], kDgEpiThreads);
      mbar_init(&acc_free[i], kDgEpiThreads);
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
    for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
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
    for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      for (int j = 0; j < kNumBwdLayers; ++j) {
        const int nk = bwd_nk(j);
        for (int h = 0; h < 2; ++h) {
          if (h == 0) {  // the layer's G operand is in TMEM (and the accumulator has been drained)
            mbar_wait(&a_ready[slot], n_a & 1);
            ++n_a;
          } else {       // N-half 0 is out of the accumulator
            mbar_wait(&acc_free[slot], n_free & 1);
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
    // 16 epilogue warps = 4 per TMEM lane quarter; per event (slot, N-half h) column group cg owns accumulator
    // columns [32 cg, +32) = gradient columns [128 h + 32 cg, +32)
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    // warps (cg, cg^1) of one quarter share a 32-row x 128 B staging slice and a named barrier; the slice
    // leaves through the TMA engine (one 4 KB bulk store), which unlike st.global (32 B/clk/SM through the LSU,
    // measured) does not hold up the epilogue warps
    const int pair_id = q * 2 + (cg >> 1);
    uint8_t* st_slice = sStage + pair_id * 4096;
    const bool pair_leader = ((cg & 1) == 0) && lane == 0;
    uint32_t n_full[2] = {0, 0};
    float4* ghead_out = reinterpret_cast<float4*>(a.scratch + scratch_ghead_offset(a.m));

    // this warp pair's 32 rows of one 64-column block -> HBM; `gdst` = the block's rows [32 q, 32 q + 32)
    auto stage_store = [&](const uint32_t (&w)[16], uint8_t* gdst) {
      if (pair_leader) bulk_wait_read<0>();  // the previous store out of the slice has been read
      __syncwarp();
      named_bar_sync(2 + pair_id, 64);
      store_words(w, st_slice + lane * 128, row, 4 * (cg & 1));
      fence_proxy_async();
      named_bar_sync(2 + pair_id, 64);
      if (pair_leader) {
        bulk_s2g(gdst, st_slice, 4096);
        bulk_commit();
      }
      __syncwarp();
    };

    for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int64_t tile0 = 2 * pair;
      float gsp[2];
      // ---- input stage per slot: heads on CUDA cores, G9 into TMEM
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int64_t tile = tile0 + s;
        const bool valid = tile < ntiles;
        const int64_t grow = tile * kTileM + row;
        const uint32_t* mask_row =
            reinterpret_cast<const uint32_t*>(a.cache + cache_mask_offset(a.m) + (size_t)tile * kMaskTileBytes) + row;
        // gz = g_rgb * rgb (1 - rgb) (sigmoid backward), g_sigma_pre = g_sigma * (sigma_pre > 0)
        float gz0 = 0.f, gz1 = 0.f, gz2 = 0.f;
        gsp[s] = 0.f;
        if (grow < a.m) {
          const float r0 = __ldg(a.rgb + 3 * grow), r1 = __ldg(a.rgb + 3 * grow + 1), r2 = __ldg(a.rgb + 3 * grow + 2);
          gz0 = __ldg(a.g_rgb + 3 * grow) * r0 * (1.f - r0);
          gz1 = __ldg(a.g_rgb + 3 * grow + 1) * r1 * (1.f - r1);
          gz2 = __ldg(a.g_rgb + 3 * grow + 2) * r2 * (1.f - r2);
          const uint32_t smask = __ldg(mask_row + kMaskSigmaWord * kTileM);
          gsp[s] = (smask & 1u) ? __ldg(a.g_sigma + grow) : 0.f;
        }
        if (cg == 0 && valid) ghead_out[tile * kTileM + row] = make_float4(gz0, gz1, gz2, gsp[s]);
        // G9 = (gz . W_out) masked by h9 > 0; this thread owns columns [32 cg, 32 cg + 32)
        const int col0 = 32 * cg;
        const uint32_t mk = valid ? __ldg(mask_row + (64 + cg) * kTileM) : 0u;
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
        tmem_st16(lane_addr + (uint32_t)s * kDgTmSlot + kDgTmA + 16 * cg, w);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&a_ready[s]);  // also: this thread has drained the slot's previous tile
        if (valid) stage_store(w, a.scratch + (size_t)tile * kGradTileBytes + (size_t)(kGradG9 + (cg >> 1)) * kBlockBytes + q * 4096);
      }
      uint32_t wh[2][16];  // bf16 pairs of N-half 0, held until half 1's MMAs have stopped reading the G operand
#pragma unroll 1
      for (int j = 0; j < kNumBwdLayers; ++j) {
        const int slot_m = 8 - j;  // ReLU mask of the layer output this gradient flows into (j >= 1): h7 .. h0
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const int64_t tile = tile0 + s;
            const bool valid = tile < ntiles;
            const int col0 = 128 * h + 32 * cg;
            const uint32_t t_slot = lane_addr + (uint32_t)s * kDgTmSlot;
            uint32_t mk = 0u;  // sign-bit mask: 0 = every column passes (layer j = 0 has no ReLU)
            if (j >= 1 && valid)
              mk = __ldg(reinterpret_cast<const uint32_t*>(a.cache + cache_mask_offset(a.m) + (size_t)tile * kMaskTileBytes) +
                         row + (slot_m * 8 + 4 * h + cg) * kTileM);
            mbar_wait(&acc_full[s], n_full[s] & 1);
            ++n_full[s];
            tc_fence_after();
            uint32_t v[32];
            tmem_ld32(t_slot + 32 * cg, v);
            tmem_ld_wait();
            if (h == 0) {
              tc_fence_before();
              mbar_arrive(&acc_free[s]);
            }
            float f[32];
            uint32_t w[16];
            masked_group(v, mk, gsp[s], j == 1 ? sW8 + col0 : nullptr, f);
            pack_words(f, w);
            if (h == 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) wh[s][i] = w[i];
            } else if (j < kNumBwdLayers - 1) {
              tmem_st16(t_slot + kDgTmA + 16 * cg, wh[s]);
              tmem_st16(t_slot + kDgTmA + 64 + 16 * cg, w);
              tmem_st_wait();
              tc_fence_before();
              mbar_arrive(&a_ready[s]);
            } else {
              tc_fence_before();
            }
            // gradient block for wgrad
            if (valid)
              stage_store(w, a.scratch + (size_t)tile * kGradTileBytes + (size_t)(2 + 4 * j + 2 * h + (cg >> 1)) * kBlockBytes + q * 4096);
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
  int x_extra;     // extra narrow block (view-direction encoding, 32 columns) or -1
  int param_w;     // weight gradient slot
  int w_row0;      // first output row in the weight tensor
  int w_col0;      // first input column
  int w_ld;        // weight leading dimension
  int valid_cols;  // accumulator columns that map to real weight columns
  int param_b;     // bias slot or -1
  int b_off;
  int cost;        // relative time per tile for the work partition (blocks streamed; more for CUDA-core-only units)
  int head;        // 0: none; 1: density head (d fc_8.weight[0,:], d fc_8.bias[0]) from g_sigma_pre and the X slices
                   // 2: fc_out (d fc_out.weight, d fc_out.bias) from gz and the X slices (X = h9, no MMA)
};
constexpr int kNumWUnits = 12;
__constant__ WUnit c_wunits[kNumWUnits];

constexpr int kWgStages = 5;
constexpr int kWgSlice = 4096;            // 32 rows of one block
constexpr int kWgHeadOff = 9 * kWgSlice;  // 32 x float4 head gradients
constexpr int kWgStageBytes = kWgHeadOff + 1024;
constexpr int kWgThreads = 192;
constexpr int kWgSmBar = kWgStages * kWgStageBytes;
constexpr int kWgSmemBytes = kWgSmBar + 256 + 1024;

struct WgradArgs {
  const uint8_t* cache;
  const uint8_t* scratch;
  ParamPtrs grads;
  int64_t m;
};

struct Segment {
  int unit;
  int64_t tile0, tile1;
};

// cost-balanced contiguous partition of (unit, tile) pairs over the grid
__device__ __forceinline__ int unit_cost(int u) { return c_wunits[u].cost; }

__device__ inline int build_segments(int64_t ntiles, Segment* seg) {
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
        seg[n].tile0 = t0;
        seg[n].tile1 = t1 < ntiles ? t1 : ntiles;
        ++n;
      }
    }
    off += span;
  }
  return n;
}

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

__global__ void __launch_bounds__(kWgThreads, 1) mlp_wgrad_kernel(WgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgSmBar);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWgStages;
  uint64_t* acc_done = bars + 2 * kWgStages;
  uint64_t* acc_free = acc_done + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 1);
  __shared__ Segment segs[kNumWUnits];
  __shared__ int nseg_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = num_tiles(a.m);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1 + 4);  // MMA commit + the four CUDA-core warps
    }
    mbar_init(acc_done, 1);
    mbar_init(acc_free, 128);
    fence_barrier_init();
    nseg_s = build_segments(ntiles, segs);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nseg = nseg_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ loader: 32-row slices of every block
    {
      const bool leader = elect_one();
      uint32_t g = 0;
      const uint8_t* ghead = a.scratch + scratch_ghead_offset(a.m);
      for (int si = 0; si < nseg; ++si) {
        const WUnit u = c_wunits[segs[si].unit];
        const int nxb = u.n_xblk + (u.x_extra >= 0 ? 1 : 0);
        const uint32_t bytes = (uint32_t)(u.n_gblk + nxb) * kWgSlice + (u.head ? 512u : 0u);
        for (int64_t tile = segs[si].tile0; tile < segs[si].tile1; ++tile) {
          const uint8_t* gsrc = a.scratch + (size_t)tile * kGradTileBytes + (size_t)u.g_blk0 * kBlockBytes;
          const uint8_t* xsrc = a.cache + (size_t)tile * kCacheTileBytes;
          for (int sl = 0; sl < 4; ++sl) {
            const uint32_t s = g % kWgStages, ph = (g / kWgStages) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            if (leader) {
              mbar_arrive_expect_tx(&full[s], bytes);
              uint8_t* dst = smem + s * kWgStageBytes;
              for (int b = 0; b < u.n_gblk; ++b)
                bulk_g2s(dst + b * kWgSlice, gsrc + (size_t)b * kBlockBytes + sl * kWgSlice, kWgSlice, &full[s]);
              uint8_t* xdst = dst + u.n_gblk * kWgSlice;
              for (int b = 0; b < u.n_xblk; ++b)
                bulk_g2s(xdst + b * kWgSlice, xsrc + (size_t)(u.x_blk0 + b) * kBlockBytes + sl * kWgSlice, kWgSlice, &full[s]);
              if (u.x_extra >= 0)
                bulk_g2s(xdst + u.n_xblk * kWgSlice, xsrc + (size_t)u.x_extra * kBlockBytes + sl * kWgSlice, kWgSlice, &full[s]);
              if (u.head) bulk_g2s(dst + kWgHeadOff, ghead + ((size_t)tile * kTileM + sl * 32) * 16, 512, &full[s]);
            }
            __syncwarp();
            ++g;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (MN-major operands)
    {
      const bool leader = elect_one();
      uint32_t g = 0;
      const uint32_t sm_u = smem_u32(smem);
      for (int si = 0; si < nseg; ++si) {
        const WUnit u = c_wunits[segs[si].unit];
        const int nhalf = u.n_gblk / 2;
        const uint32_t n_main = (uint32_t)u.n_xblk * 64u;
        const uint32_t idesc_main = make_idesc_bf16(n_main, true, true);
        const uint32_t idesc_extra = make_idesc_bf16(32, true, true);
        if (si > 0) mbar_wait(acc_free, (uint32_t)(si - 1) & 1);
        tc_fence_after();
        bool first = true;
        for (int64_t tile = segs[si].tile0; tile < segs[si].tile1; ++tile) {
          for (int sl = 0; sl < 4; ++sl) {
            const uint32_t s = g % kWgStages, ph = (g / kWgStages) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t st = sm_u + s * kWgStageBytes;
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
            ++g;
          }
        }
        if (leader) umma_commit(acc_done);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ bias / head sums on CUDA cores + accumulator flush
    const int q = warp & 3;
    const int tid = (warp - 2) * 32 + lane;  // 0..127
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t g = 0;
    for (int si = 0; si < nseg; ++si) {
      const WUnit u = c_wunits[segs[si].unit];
      const bool do_bias = u.param_b >= 0 && tid < u.n_gblk * 32;  // (units with a bias have hsplit == 1)
      // heads: 2 columns per thread; with fewer than 128 column pairs (fc_out: 64) the rows are split instead
      const int head_pairs = u.n_xblk * 32;
      const bool do_head = u.head != 0;
      const int hsplit = (u.head != 0 && head_pairs < 128) ? 128 / head_pairs : 1;  // 1 or 2
      const int hrows = 32 / hsplit;
      const int hrow0 = (tid / (128 / hsplit)) * hrows;
      const int col = 2 * (tid % (128 / hsplit));  // this thread's column pair, in G (bias) and in X (heads)
      const uint32_t coff = (uint32_t)(col >> 6) * kWgSlice + ((col & 63) & 7) * 2;
      const uint32_t chunk = (uint32_t)((col & 63) >> 3);
      float b0 = 0.f, b1 = 0.f;                                                 // bias column sums
      float h00 = 0.f, h01 = 0.f, h10 = 0.f, h11 = 0.f, h20 = 0.f, h21 = 0.f;  // head sums [j][column of the pair]
      float hb = 0.f;                                                           // head bias sums (threads 0..3)
      for (int64_t tile = segs[si].tile0; tile < segs[si].tile1; ++tile) {
        for (int sl = 0; sl < 4; ++sl) {
          const uint32_t s = g % kWgStages, ph = (g / kWgStages) & 1;
          mbar_wait(&full[s], ph);
          const uint8_t* st = smem + s * kWgStageBytes;
          if (do_bias) {
            const uint8_t* gs = st + coff;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              const uint32_t w = *reinterpret_cast<const uint32_t*>(gs + r * 128 + (((chunk ^ (uint32_t)(r & 7)) & 7u) << 4));
              b0 += bf16_lo(w);
              b1 += bf16_hi(w);
            }
          }
          if (u.head) {
            const float4* gh = reinterpret_cast<const float4*>(st + kWgHeadOff);
            if (do_head) {
              const uint8_t* xs = st + u.n_gblk * kWgSlice + coff;
#pragma unroll 8
              for (int rr = 0; rr < hrows; ++rr) {
                const int r = hrow0 + rr;
                const uint32_t w = *reinterpret_cast<const uint32_t*>(xs + r * 128 + (((chunk ^ (uint32_t)(r & 7)) & 7u) << 4));
                const float x0 = bf16_lo(w), x1 = bf16_hi(w);
                const float4 gv = gh[r];
                if (u.head == 1) {
                  h00 = fmaf(gv.w, x0, h00);
                  h01 = fmaf(gv.w, x1, h01);
                } else {
                  h00 = fmaf(gv.x, x0, h00), h01 = fmaf(gv.x, x1, h01);
                  h10 = fmaf(gv.y, x0, h10), h11 = fmaf(gv.y, x1, h11);
                  h20 = fmaf(gv.z, x0, h20), h21 = fmaf(gv.z, x1, h21);
                }
              }
            }
            if (tid < 4) {
              const float* ghf = reinterpret_cast<const float*>(gh);
#pragma unroll 8
              for (int r = 0; r < 32; ++r) hb += ghf[4 * r + tid];
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[s]);
          ++g;
        }
      }
      if (do_bias) {
        float* db = a.grads.p[u.param_b] + u.b_off;
        atomicAdd(db + col, b0);
        atomicAdd(db + col + 1, b1);
      }
      if (u.head == 1) {
        if (do_head) {
          atomicAdd(a.grads.p[W_8] + col, h00);  // d fc_8.weight[0, col]
          atomicAdd(a.grads.p[W_8] + col + 1, h01);
        }
        if (tid == 3) atomicAdd(a.grads.p[B_8], hb);  // d fc_8.bias[0] = sum g_sigma_pre
      } else if (u.head == 2) {
        if (do_head) {
          float* dw = a.grads.p[W_OUT];
          atomicAdd(dw + col, h00), atomicAdd(dw + col + 1, h01);
          atomicAdd(dw + kH + col, h10), atomicAdd(dw + kH + col + 1, h11);
          atomicAdd(dw + 2 * kH + col, h20), atomicAdd(dw + 2 * kH + col + 1, h21);
        }
        if (tid < 3) atomicAdd(a.grads.p[B_OUT] + tid, hb);
      }
      // flush the accumulator of this segment
      mbar_wait(acc_done, (uint32_t)si & 1);
      tc_fence_after();
      const int nhalf = u.n_gblk / 2;
      const int ncols = u.n_xblk * 64 + (u.x_extra >= 0 ? 32 : 0);
      float* dw = a.grads.p[u.param_w];
      for (int h = 0; h < nhalf; ++h) {
        const int out_row = u.w_row0 + h * 128 + q * 32 + lane;
        float* dst = dw + (size_t)out_row * u.w_ld + u.w_col0;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(lane_addr + h * 256 + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < u.valid_cols) atomicAdd(dst + c0 + i, __uint_as_float(v[i]));
        }
      }
      tc_fence_before();
      mbar_arrive(acc_free);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

static void build_wunits(WUnit* u) {
  int n = 0;
  auto add = [&](int g0, int ng, int x0, int nx, int xe, int pw, int row0, int col0, int ld, int valid, int pb, int boff,
                 int head) {
    const int cost = ng > 0 ? ng + nx + (xe >= 0 ? 1 : 0) : 5;  // fc_out streams 2 blocks but is CUDA-core bound
    u[n++] = WUnit{g0, ng, x0, nx, xe, pw, row0, col0, ld, valid, pb, boff, cost, head};
  };
  add(grad_g(0), 4, kCachePe, 1, -1, W_IN, 0, 0, kP, kP, B_IN, 0, 0);                      // fc_in : G0 x pe
  for (int l = 1; l <= 4; ++l) add(grad_g(l), 4, cache_h(l - 1), 4, -1, 2 * l, 0, 0, kF, kF, 2 * l + 1, 0, 0);  // fc_1..4
  add(grad_g(5), 4, kCachePe, 1, -1, W_5, 0, 0, kP + kF, kP, B_5, 0, 0);                   // fc_5, position columns
  add(grad_g(5), 4, cache_h(4), 4, -1, W_5, 0, kP, kP + kF, kF, -1, 0, 0);                 // fc_5, h4 columns
  add(grad_g(6), 4, cache_h(5), 4, -1, W_6, 0, 0, kF, kF, B_6, 0, 0);
  add(grad_g(7), 4, cache_h(6), 4, -1, W_7, 0, 0, kF, kF, B_7, 0, 0);
  add(kGradG8, 4, cache_h(7), 4, -1, W_8, 1, 0, kF, kF, B_8, 1, 1);                        // fc_8 rows 1..256 + density row
  add(kGradG9, 2, kCacheFeat, 4, kCacheDe, W_9, 0, 0, kF + kV, kF + kV, B_9, 0, 0);        // fc_9 : G9 x [feat | de]
  add(0, 0, kCacheH9, 2, -1, W_OUT, 0, 0, kH, 0, -1, 0, 2);                                // fc_out (CUDA cores only)
}

__global__ void zero_grads_kernel(ParamPtrs g) {
  const int sizes[22] = {kF * kP, kF, kF * kF, kF, kF * kF, kF, kF * kF, kF, kF * kF, kF, kF * (kP + kF), kF,
                         kF * kF, kF, kF * kF, kF, (kF + 1) * kF, kF + 1, kH * (kF + kV), kH, 3 * kH, 3};
  float* p = g.p[blockIdx.y];
  const int n = sizes[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0.f;
}

static bool g_wunits_ready = false;
static int g_bwd_phase_mask = 7;  // bit 0: zero the gradients, bit 1: dgrad chain, bit 2: wgrad (profiling aid)

}  // namespace nerf

using namespace nerf;

extern "C" int nerf_debug_set_bwd_phases(int mask) {
  g_bwd_phase_mask = mask;
  return NERF_OK;
}

extern "C" int nerf_mlp_bf16_backward(const void* packed_dev, const void* cache_dev, const float* rgb_dev, int64_t m,
                                      const float* g_sigma_dev, const float* g_rgb_dev, float* const* grads,
                                      void* scratch_dev, nerf_stream_t stream) {
  NERF_CHECK_ARG(m > 0, "nerf_mlp_bf16_backward: row count must be positive");
  NERF_CHECK_ARG(packed_dev && cache_dev && rgb_dev && g_sigma_dev && g_rgb_dev && grads && scratch_dev,
                 "nerf_mlp_bf16_backward: null pointer");
  cudaStream_t st = as_stream(stream);
  static bool attr_set = false;
  if (!attr_set) {
    NERF_CUDA(cudaFuncSetAttribute(mlp_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDgSmemBytes));
    NERF_CUDA(cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes));
    attr_set = true;
  }
  if (!g_wunits_ready) {
    WUnit table[kNumWUnits];
    build_wunits(table);
    NERF_CUDA(cudaMemcpyToSymbol(c_wunits, table, sizeof(table)));
    g_wunits_ready = true;
  }
  ParamPtrs gp;
  for (int i = 0; i < NERF_NUM_PARAM_TENSORS; ++i) {
    NERF_CHECK_ARG(grads[i] != nullptr, "nerf_mlp_bf16_backward: null gradient pointer");
    gp.p[i] = grads[i];
  }
  if (g_bwd_phase_mask & 1) {
    zero_grads_kernel<<<dim3(32, 22), 256, 0, st>>>(gp);
    NERF_LAUNCH_CHECK();
  }
  const int64_t ntiles = num_tiles(m);
  const int sms = sm_count();
  if (g_bwd_phase_mask & 2) {
    DgradArgs a;
    a.packed = reinterpret_cast<const uint8_t*>(packed_dev);
    a.cache = reinterpret_cast<const uint8_t*>(cache_dev);
    a.rgb = rgb_dev, a.g_sigma = g_sigma_dev, a.g_rgb = g_rgb_dev;
    a.scratch = reinterpret_cast<uint8_t*>(scratch_dev);
    a.m = m;
    const int64_t npairs = (ntiles + 1) / 2;  // a CTA works on two tiles at a time
    const int grid = (int)(npairs < sms ? npairs : sms);
    mlp_dgrad_kernel<<<grid, kDgThreads, kDgSmemBytes, st>>>(a);
    NERF_LAUNCH_CHECK();
  }
  if (g_bwd_phase_mask & 4) {
    WgradArgs a;
    a.cache = reinterpret_cast<const uint8_t*>(cache_dev);
    a.scratch = reinterpret_cast<const uint8_t*>(scratch_dev);
    a.grads = gp;
    a.m = m;
    mlp_wgrad_kernel<<<sms, kWgThreads, kWgSmemBytes, st>>>(a);
    NERF_LAUNCH_CHECK();
  }
  return NERF_OK;
}
