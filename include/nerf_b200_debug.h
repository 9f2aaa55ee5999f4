/* nerf_b200_debug.h -- test and profiling aids of libnerf_b200; NOT part of the drop-in boundary (include/nerf_b200.h).
 *
 *   nerf_debug_*     in-kernel timeline hooks, exported by libnerf_b200.so (the kernels read the buffers they set)
 *   nerf_selftest_*  tcgen05 building-block checks and micro-benchmarks, exported by libnerf_b200_selftest.so, a separate
 *                    library built from csrc/mlp_tc_selftest.cu that only tests/ and tools/ load
 */
#ifndef NERF_B200_DEBUG_H
#define NERF_B200_DEBUG_H

#include "nerf_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * self tests of the tensor-core building blocks (used by tests/ on the GPU box)
 * ---------------------------------------------------------------------------------------------- */
/* D(128 x n) = A(128 x k) * B(n x k)^T with one tcgen05 tile; a/b bf16 bits row-major, d fp32 row-major.
 * variant selects operand majors: 0 = K-major/K-major, 1 = MN-major A and B (wgrad form). */
int nerf_selftest_umma(const uint16_t* a_dev, const uint16_t* b_dev, float* d_dev, int n, int k, int variant,
                       nerf_stream_t stream);

/* debugging aid: when buf_dev != NULL, CTA 0 of the next nerf_mlp_bf16_forward launches records, for its first
 * `tiles` tiles and every layer, 8 values into buf_dev[(tile*10 + layer)*8 + k]: SM-clock stamps k=0 MMA layer
 * start, 1 MMA layer issued, 2 accumulator seen by the epilogue, 3 epilogue done; cycle sums k=4 MMA thread waiting
 * for activations, 5 waiting for weights.  Pass NULL to switch it off.  Bits 16+ of `tiles` select a debug store mode
 * of the training-mode forward (0 normal, 1 skip the cache block stores, 2 wrap them onto an L2-resident window): timing
 * experiments only, the results of modes 1 and 2 are unusable. */
int nerf_debug_set_profile_buffer(unsigned long long* buf_dev, int tiles);

/* debugging aid: when buf_dev != NULL every CTA of the next wgrad launches records 8 values into
 * buf_dev[cta*16 + k]: globaltimer ns k=0 start, 1 last accumulator complete, 2 end; k=3 first unit, 4 segment count,
 * 5/6 tiles of the first / second segment.  Pass NULL to switch it off. */
int nerf_debug_set_wgrad_profile(unsigned long long* buf_dev);

/* timing experiment (the gradients of such a launch are unusable): the next wgrad launches read the gradient tile images of
 * tile (t % wrap_g) and the activation tile images of tile (t % wrap_x) instead of tile t, i.e. from an L2-resident window --
 * what the kernel would run at if a producer kept that operand in L2.  0 switches a wrap off. */
int nerf_debug_set_wgrad_wrap(int64_t wrap_g, int64_t wrap_x);

/* tuning aid: replaces the relative per-tile costs of the 11 wgrad work units (fc_in, fc_1..4, fc_5 position columns, fc_5 h4
 * columns, fc_6, fc_7, fc_8, fc_9) that the static (unit, tile) partition balances by; NULL restores the built-in table. */
int nerf_debug_set_wgrad_costs(const int* costs, int n);

/* self test + rate probe of the CTA-pair MMA (tcgen05 cta_group::2, M = 256): a (256 x k), b (n x k) bf16 bits, d (256 x n)
 * fp32, all row-major; bit 0 of ts puts the A operand in tensor memory.  `pairs` clusters of two CTAs all compute the same
 * product; with iters > 0 each leader then times iters x (k/16) MMAs into cycles_dev[pair].  With bit 1 of ts (n <= 128,
 * iters > 0) BOTH CTAs of every pair issue pair MMAs, each into its own accumulator: cycles_dev[cta] for 2 x pairs
 * entries, and d receives the accumulator the second CTA issued into. */
int nerf_selftest_umma2(const uint16_t* a_dev, const uint16_t* b_dev, float* d_dev, int n, int k, int ts, int pairs, int iters,
                        unsigned long long* cycles_dev, nerf_stream_t stream);

/* micro-benchmark: `blocks` CTAs each issue `iters` x 4 tcgen05.mma (M=128, N=n, K=16) -- mode bit 0: A operand from TMEM instead of
 * shared memory; bit 1: tcgen05.commit after every group of 4; bit 2: probe a completed mbarrier before every group; bit 4: alternate between two accumulators -- while `bg_warps` extra warps each perform `bg_iters` tcgen05.ld
 * (bg_store = 0) or tcgen05.st (1) of 32 lanes x 32 columns.  cycles_dev[block] = SM cycles of the MMA thread,
 * cycles_dev[blocks + block] = cycles of one background warp.  cycles_dev holds 2*blocks entries. */
int nerf_selftest_mma_rate(int blocks, int iters, int n, int mode, int bg_warps, int bg_iters, int bg_store,
                           unsigned long long* cycles_dev, nerf_stream_t stream);

/* micro-benchmark: `blocks` CTAs stream `bytes` (multiple of 4096) to dst_dev in 4 KB pieces.  mode 0: st.global.v4,
 * 1: st.global.cs.v4, 2: bulk stores (TMA engine) from shared memory, 3 / 4: bulk stores with an L2 evict_first /
 * evict_last cache hint.  Used to find the HBM write ceiling the training-cache stores run against. */
int nerf_selftest_write_bw(void* dst_dev, size_t bytes, int mode, int blocks, nerf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NERF_B200_DEBUG_H */
