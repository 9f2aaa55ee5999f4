// Fixed geometry and memory layouts of the bf16 tensor-core MLP path (NeRF(63, 27, 256), network/nerf.py:49-59).
//
// HBM layouts ("tile images", see tc_common.cuh): everything the training step stages in HBM is stored as
// 128-row x 64-column bf16 blocks (128B-swizzled rows) so that 1-D bulk copies (TMA engine) move them and the
// very same bytes serve as K-major operands (forward / dgrad chains) and MN-major operands (wgrad).
//
//   packed weights : [forward chunks][dgrad (transposed) chunks][fp32 constants]; a chunk is one contiguous 16 KB block
//   training cache : per 128-row tile 40 activation blocks, then per tile 69 x 128 ReLU-mask words
//   bwd scratch    : per tile 38 gradient blocks, then per row float4 {gz0, gz1, gz2, g_sigma_pre}
//
// Cache and scratch tiles are stored SLICE-MAJOR: the tile is four 32-row slices, each slice holds the 4 KB
// (32 rows x 128 B) piece of every block back to back.  The chain epilogues write 4 KB pieces anyway, and wgrad
// fetches the pieces of several consecutive blocks of one slice with a single bulk copy.
#pragma once

#include <stddef.h>
#include <stdint.h>

namespace nerf {

constexpr int kP = 63, kV = 27, kF = 256, kH = 128;
constexpr int kTileM = 128;
constexpr int kBlockBytes = 16384;  // one 128 x 64 bf16 block

// Weight chunks are 128 output rows x 64 K-columns (16 KB).  A 256-wide layer is computed as two N-halves: the MMAs
// of half 0 (all K chunks) are committed on their own so the epilogue drains output columns [0,128) while the
// tensor pipe works on half 1, and the next layer's first k-blocks are ready the moment half 1 completes.
constexpr int kChunkBytes = 16384;

// ---- forward chain: tensor-core layers fc_in, fc_1..fc_7, fc_8 (feature rows), fc_9
constexpr int kNumFwdLayers = 10;
__host__ __device__ constexpr int fwd_nk(int l) { return l == 0 ? 1 : ((l == 5 || l == 9) ? 5 : 4); }
__host__ __device__ constexpr int fwd_nh(int l) { return l == 9 ? 1 : 2; }  // N-halves (fc_9 is 128 wide)
constexpr int kFwdChunks = 2 * (1 + 4 * 4 + 5 + 3 * 4) + 5;  // 73
// packed order: per layer, per K chunk, per N-half (the halves of one K chunk are adjacent: together one 256-row operand)
__host__ __device__ constexpr int fwd_layer_chunk0(int l) {
  int c = 0;
  for (int i = 0; i < l; ++i) c += fwd_nk(i) * fwd_nh(i);
  return c;
}
__host__ __device__ constexpr int fwd_chunk_index(int l, int h, int kb) { return fwd_layer_chunk0(l) + kb * fwd_nh(l) + h; }
constexpr size_t kFwdWeightBytes = (size_t)kFwdChunks * kChunkBytes;

// Order in which the chains visit the four activation k-blocks.  The epilogue's column half h owns blocks {h, h+2}
// and finishes h first, so blocks {0, 1} become ready before {2, 3}: natural order.  (Measured: visiting 0,2,1,3
// costs +7% per layer.)
__host__ __device__ constexpr int kb_order(int i) { return i; }

// ---- dgrad chain: layers j = 0..8 multiply by fc_9^T (128 -> 256), fc_8^T, fc_7^T, fc_6^T, fc_5^T (h4 columns),
//      fc_4^T .. fc_1^T; every chunk is 128 rows (input features of one N-half) x 64 (output features)
constexpr int kNumBwdLayers = 9;
__host__ __device__ constexpr int bwd_nk(int j) { return j == 0 ? 2 : 4; }
constexpr int kBwdChunks = 2 * (2 + 8 * 4);  // 68
constexpr size_t kBwdWeightBytes = (size_t)kBwdChunks * kChunkBytes;

// ---- fp32 constants block (float offsets)
constexpr int kCBias = 0;        // 8 x 256 : biases of fc_in, fc_1..fc_7
constexpr int kCBias8 = 2048;    // 256     : fc_8.bias[1:257]
constexpr int kCBias9 = 2304;    // 128
constexpr int kCW8Row0 = 2432;   // 256     : fc_8.weight[0, :]  (density head)
constexpr int kCWout = 2688;     // 3 x 128 : fc_out.weight
constexpr int kCB8_0 = 3072;     // 1
constexpr int kCBout = 3073;     // 3
constexpr int kCFloats = 3080;

constexpr size_t kPackedFwdOff = 0;
constexpr size_t kPackedBwdOff = kFwdWeightBytes;
constexpr size_t kPackedConstOff = kPackedBwdOff + kBwdWeightBytes;
constexpr size_t kPackedBytes = kPackedConstOff + sizeof(float) * kCFloats;

// ---- training cache: activation blocks per tile
constexpr int kCachePe = 0;                                  // encoded position (63 + zero column)
__host__ __device__ constexpr int cache_h(int l) { return 1 + 4 * l; }  // h_l = output of fwd layer l (l = 0..7), 4 blocks
constexpr int kCacheFeat = 33;                               // fc_8 feature output (no ReLU), 4 blocks
constexpr int kCacheDe = 37;                                 // encoded view direction (27 of 64 columns)
constexpr int kCacheH9 = 38;                                 // fc_9 output (128 columns), 2 blocks
constexpr int kCacheBlocks = 40;
constexpr size_t kCacheTileBytes = (size_t)kCacheBlocks * kBlockBytes;
constexpr int kSliceBytes = 4096;  // 32 rows x 128 B of one block
// byte offset inside a tile of rows [32 q, 32 q + 32) of block `blk` (`nblk` = blocks per tile)
__host__ __device__ constexpr size_t slice_off(int nblk, int blk, int q) { return ((size_t)q * nblk + blk) * kSliceBytes; }
__host__ __device__ constexpr size_t cache_slice_off(int blk, int q) { return ((size_t)q * kCacheBlocks + blk) * kSliceBytes; }
// ReLU masks: word (slot s, 32-column group c) of row r at ((s*8 + c)*128 + r); slots 0..7 = h0..h7, slot 8 = h9
// (4 words).  Bit (31 - i) of a word is the SIGN BIT of the pre-activation of column 32c + i (1 = negative = no
// gradient), collected with one funnel shift per element.  Word 68 bit 0 = (sigma_pre > 0).
constexpr int kMaskWords = 69;
constexpr int kMaskSigmaWord = 68;
constexpr size_t kMaskTileBytes = (size_t)kMaskWords * kTileM * 4;

__host__ __device__ inline int64_t num_tiles(int64_t m) { return (m + kTileM - 1) / kTileM; }
__host__ __device__ inline size_t cache_mask_offset(int64_t m) { return (size_t)num_tiles(m) * kCacheTileBytes; }
__host__ __device__ inline size_t cache_bytes(int64_t m) { return (size_t)num_tiles(m) * (kCacheTileBytes + kMaskTileBytes); }

// ---- backward scratch: gradient blocks per tile (G_x = dL/d(pre-activation of layer x))
constexpr int kGradG9 = 0;                                   // 2 blocks
constexpr int kGradG8 = 2;                                   // feature part of fc_8's output gradient, 4 blocks
__host__ __device__ constexpr int grad_g(int l) { return 6 + 4 * (7 - l); }  // G_l for l = 7..0
constexpr int kGradBlocks = 38;
constexpr size_t kGradTileBytes = (size_t)kGradBlocks * kBlockBytes;
__host__ __device__ constexpr size_t grad_slice_off(int blk, int q) { return ((size_t)q * kGradBlocks + blk) * kSliceBytes; }
// head gradients per row as float4 {gz0, gz1, gz2, g_sigma_pre} (gz = dL/d(fc_out pre-sigmoid))
__host__ __device__ inline size_t scratch_ghead_offset(int64_t m) { return (size_t)num_tiles(m) * kGradTileBytes; }
__host__ __device__ inline size_t scratch_bytes(int64_t m) { return scratch_ghead_offset(m) + (size_t)num_tiles(m) * kTileM * 16; }

// parameter slots in state_dict order
enum ParamSlot {
  W_IN = 0, B_IN, W_1, B_1, W_2, B_2, W_3, B_3, W_4, B_4, W_5, B_5, W_6, B_6, W_7, B_7, W_8, B_8, W_9, B_9, W_OUT, B_OUT
};

struct ParamPtrs {
  float* p[22];
};

}  // namespace nerf
