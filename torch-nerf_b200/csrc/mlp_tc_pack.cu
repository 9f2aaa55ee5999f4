// Weight packing for the tensor-core chains: fp32 (out,in) nn.Linear weights (network/nerf.py:49-59) -> bf16
// 128B-swizzled K-major chunks (N rows x 64 K-columns) in exactly the order the forward and dgrad chains stream
// them, plus the fp32 constants the epilogues need.  Re-run after every optimizer step (2.3 MB written).
#include <initializer_list>

#include "common.cuh"
#include "mlp_tc_layout.cuh"
#include "tc_common.cuh"

namespace nerf {
using namespace tc;

struct PackChunk {
  int param;      // index into the 22-pointer parameter array
  int src_row0;   // first source row
  int src_col0;   // first source column
  int ld;         // source leading dimension
  int nrows;      // destination rows (N of the MMA)
  int valid_k;    // K columns copied; the rest of the 64 are zero
  int transpose;  // 0: dst(n,k) = W[row0+n][col0+k]   1: dst(n,k) = W[row0+k][col0+n]
  uint32_t dst_off;
};
constexpr int kMaxPackChunks = 160;
__constant__ PackChunk c_pack[kMaxPackChunks];

__device__ __forceinline__ void pack_unit(const ParamPtrs& params, uint8_t* __restrict__ packed, const PackChunk& pc, int u) {
  if (u >= pc.nrows * 8) return;  // one 16-byte unit (8 bf16) per thread
  int n = u >> 3, j = u & 7;
  const float* w = params.p[pc.param];
  uint32_t out[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int k = 8 * j + 2 * e + h;
      float x = 0.f;
      if (k < pc.valid_k)
        x = pc.transpose ? w[(size_t)(pc.src_row0 + k) * pc.ld + pc.src_col0 + n]
                         : w[(size_t)(pc.src_row0 + n) * pc.ld + pc.src_col0 + k];
      v[h] = x;
    }
    out[e] = pack_bf16(v[0], v[1]);
  }
  uint4* dst = reinterpret_cast<uint4*>(packed + pc.dst_off + n * 128 + ((j ^ (n & 7)) << 4));
  *dst = make_uint4(out[0], out[1], out[2], out[3]);
}

__device__ __forceinline__ void pack_const(const ParamPtrs& params, float* __restrict__ c, int i) {
  if (i >= kCFloats) return;
  float v = 0.f;
  if (i < kCBias8) v = params.p[2 * (i >> 8) + 1][i & 255];   // biases of fc_in, fc_1..fc_7
  else if (i < kCBias9) v = params.p[B_8][1 + (i - kCBias8)];  // fc_8.bias[1:]
  else if (i < kCW8Row0) v = params.p[B_9][i - kCBias9];
  else if (i < kCWout) v = params.p[W_8][i - kCW8Row0];        // fc_8.weight[0, :]
  else if (i < kCB8_0) v = params.p[W_OUT][i - kCWout];        // fc_out.weight (3,128)
  else if (i == kCB8_0) v = params.p[B_8][0];
  else if (i < kCBout + 3) v = params.p[B_OUT][i - kCBout];
  c[i] = v;
}

__global__ void __launch_bounds__(256) pack_weights_kernel(ParamPtrs params, uint8_t* __restrict__ packed) {
  pack_unit(params, packed, c_pack[blockIdx.y], blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void pack_consts_kernel(ParamPtrs params, float* __restrict__ c) {
  pack_const(params, c, blockIdx.x * blockDim.x + threadIdx.x);
}

// Prologue of a training iteration in ONE launch: the bf16 images of BOTH networks (weights + constants) and the zero
// fill of two buffers (the flat gradient buffer, the loss accumulators).  grid = (4, chunks + 4 + kZeroRows, 2):
// blockIdx.z = network; blockIdx.y selects the role.
constexpr int kZeroRows = 32;
struct PrologueArgs {
  ParamPtrs params[2];
  uint8_t* packed[2];
  float* zero_ptr[2];
  int64_t zero_count[2];
  int n_chunks;
};
__global__ void __launch_bounds__(256) train_prologue_kernel(PrologueArgs a) {
  const int net = blockIdx.z, y = blockIdx.y;
  if (y < a.n_chunks) {
    pack_unit(a.params[net], a.packed[net], c_pack[y], blockIdx.x * blockDim.x + threadIdx.x);
  } else if (y < a.n_chunks + 4) {
    pack_const(a.params[net], reinterpret_cast<float*>(a.packed[net] + kPackedConstOff),
               ((y - a.n_chunks) * 4 + blockIdx.x) * 256 + threadIdx.x);
  } else {
    float* z = a.zero_ptr[net];
    const int64_t cnt = a.zero_count[net];
    const int64_t stride = (int64_t)kZeroRows * 4 * 256;
    for (int64_t i = ((int64_t)(y - a.n_chunks - 4) * 4 + blockIdx.x) * 256 + threadIdx.x; i < cnt; i += stride) z[i] = 0.f;
  }
}

static int build_pack_table(PackChunk* t) {
  int n = 0;
  uint32_t off = (uint32_t)kPackedFwdOff;
  auto add = [&](int param, int row0, int col0, int ld, int nrows, int valid, int transpose) {
    t[n++] = PackChunk{param, row0, col0, ld, nrows, valid, transpose, off};
    off += (uint32_t)nrows * 128u;
  };
  // ---- forward chain: per layer, per K chunk, per N-half; chunks are 128 output rows x 64 K-columns.  The two N-halves
  //      of one K chunk are adjacent, so the 32 KB pair is also ONE K-major operand of 256 rows (the inference
  //      chain's N = 256 MMAs); the training chain fetches the 16 KB halves separately (fwd_chunk_index).
  struct KCols {
    int col0, valid;
  };
  auto fwd_layer = [&](int param, int row0, int ld, int nh_count, std::initializer_list<KCols> ks) {
    for (const KCols& k : ks)
      for (int nh = 0; nh < nh_count; ++nh) add(param, row0 + 128 * nh, k.col0, ld, 128, k.valid, 0);
  };
  fwd_layer(W_IN, 0, kP, 2, {{0, kP}});
  for (int l = 1; l <= 4; ++l) fwd_layer(2 * l, 0, kF, 2, {{0, 64}, {64, 64}, {128, 64}, {192, 64}});
  fwd_layer(W_5, 0, kP + kF, 2, {{0, kP}, {kP, 64}, {kP + 64, 64}, {kP + 128, 64}, {kP + 192, 64}});  // position columns, then h4
  for (int l = 6; l <= 7; ++l) fwd_layer(2 * l, 0, kF, 2, {{0, 64}, {64, 64}, {128, 64}, {192, 64}});
  fwd_layer(W_8, 1, kF, 2, {{0, 64}, {64, 64}, {128, 64}, {192, 64}});                                   // rows 1..256
  fwd_layer(W_9, 0, kF + kV, 1, {{0, 64}, {64, 64}, {128, 64}, {192, 64}, {kF, kV}});                    // features, then view columns
  // ---- dgrad chain: dst(n = input feature, k = output feature); per layer, per N-half of the INPUT features
  off = (uint32_t)kPackedBwdOff;
  auto bwd_layer = [&](int param, int out_row0, int in_col0, int ld, int nk) {
    for (int nh = 0; nh < 2; ++nh)
      for (int kb = 0; kb < nk; ++kb) add(param, out_row0 + 64 * kb, in_col0 + 128 * nh, ld, 128, 64, 1);
  };
  bwd_layer(W_9, 0, 0, kF + kV, 2);       // fc_9^T, feature inputs only (128 outputs -> 2 K chunks)
  bwd_layer(W_8, 1, 0, kF, 4);            // fc_8^T, feature rows
  bwd_layer(W_7, 0, 0, kF, 4);
  bwd_layer(W_6, 0, 0, kF, 4);
  bwd_layer(W_5, 0, kP, kP + kF, 4);      // fc_5^T, h4 columns only
  for (int l = 4; l >= 1; --l) bwd_layer(2 * l, 0, 0, kF, 4);
  return n;
}

// the chunk table lives in constant memory, i.e. per device: uploaded once on every device the process uses
constexpr int kMaxDevices = 64;
static bool g_pack_table_ready[kMaxDevices] = {};
static int g_pack_chunks = 0;

}  // namespace nerf

using namespace nerf;

extern "C" {

size_t nerf_mlp_bf16_packed_bytes(void) { return kPackedBytes; }

int nerf_mlp_bf16_pack(const float* const* params, void* packed_dev, nerf_stream_t stream) {
  NERF_CHECK_ARG(params && packed_dev, "nerf_mlp_bf16_pack: null pointer");
  cudaStream_t st = as_stream(stream);
  int dev = 0;
  NERF_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices || !g_pack_table_ready[dev]) {
    PackChunk table[kMaxPackChunks];
    g_pack_chunks = build_pack_table(table);
    NERF_CUDA(cudaMemcpyToSymbol(c_pack, table, sizeof(PackChunk) * g_pack_chunks));
    if (dev >= 0 && dev < kMaxDevices) g_pack_table_ready[dev] = true;
  }
  ParamPtrs pp;
  for (int i = 0; i < NERF_NUM_PARAM_TENSORS; ++i) {
    NERF_CHECK_ARG(params[i] != nullptr, "nerf_mlp_bf16_pack: null parameter pointer");
    pp.p[i] = const_cast<float*>(params[i]);
  }
  dim3 grid((128 * 8 + 255) / 256, g_pack_chunks);
  pack_weights_kernel<<<grid, 256, 0, st>>>(pp, reinterpret_cast<uint8_t*>(packed_dev));
  NERF_LAUNCH_CHECK();
  pack_consts_kernel<<<(kCFloats + 255) / 256, 256, 0, st>>>(
      pp, reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(packed_dev) + kPackedConstOff));
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_train_prologue(const float* const* params_a, void* packed_a_dev, const float* const* params_b, void* packed_b_dev,
                        float* zero0_dev, int64_t zero0_count, float* zero1_dev, int64_t zero1_count, nerf_stream_t stream) {
  NERF_CHECK_ARG(params_a && packed_a_dev && params_b && packed_b_dev, "nerf_train_prologue: null pointer");
  NERF_CHECK_ARG(zero0_count >= 0 && zero1_count >= 0 && (zero0_dev || zero0_count == 0) && (zero1_dev || zero1_count == 0),
                 "nerf_train_prologue: bad zero-fill arguments");
  int dev = 0;
  NERF_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices || !g_pack_table_ready[dev]) {
    PackChunk table[kMaxPackChunks];
    g_pack_chunks = build_pack_table(table);
    NERF_CUDA(cudaMemcpyToSymbol(c_pack, table, sizeof(PackChunk) * g_pack_chunks));
    if (dev >= 0 && dev < kMaxDevices) g_pack_table_ready[dev] = true;
  }
  PrologueArgs a;
  for (int i = 0; i < NERF_NUM_PARAM_TENSORS; ++i) {
    NERF_CHECK_ARG(params_a[i] != nullptr && params_b[i] != nullptr, "nerf_train_prologue: null parameter pointer");
    a.params[0].p[i] = const_cast<float*>(params_a[i]);
    a.params[1].p[i] = const_cast<float*>(params_b[i]);
  }
  a.packed[0] = reinterpret_cast<uint8_t*>(packed_a_dev), a.packed[1] = reinterpret_cast<uint8_t*>(packed_b_dev);
  a.zero_ptr[0] = zero0_dev, a.zero_ptr[1] = zero1_dev;
  a.zero_count[0] = zero0_count, a.zero_count[1] = zero1_count;
  a.n_chunks = g_pack_chunks;
  train_prologue_kernel<<<dim3(4, g_pack_chunks + 4 + kZeroRows, 2), 256, 0, as_stream(stream)>>>(a);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

size_t nerf_mlp_bf16_cache_bytes(int64_t m) { return m > 0 ? cache_bytes(m) : 0; }
size_t nerf_mlp_bf16_bwd_scratch_bytes(int64_t m) { return m > 0 ? scratch_bytes(m) : 0; }

}  // extern "C"
