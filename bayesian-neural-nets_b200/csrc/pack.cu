// Operand staging for the bf16 tensor-core path: fp32 -> bf16 operand pairs (+ transposes), and the
// column sums over the batch that feed the bias gradients.
#include "common.cuh"

namespace lbbnn {
namespace {

// 32x32 tile per block (32x8 threads): row-major outputs straight from registers, transposed
// outputs through a padded smem tile so both directions are coalesced.
__global__ void __launch_bounds__(256) bf16_pack_kernel(const float* __restrict__ a, const float* __restrict__ b, int op,
                                                        int64_t rows, int64_t cols, __nv_bfloat16* __restrict__ o1,
                                                        __nv_bfloat16* __restrict__ o2, __nv_bfloat16* __restrict__ o1T,
                                                        __nv_bfloat16* __restrict__ o2T) {
  __shared__ float t1[32][33], t2[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + ty + i * 8, c = c0 + tx;
    float v1 = 0.f, v2 = 0.f;
    if (r < rows && c < cols) {
      v1 = a[r * cols + c];
      v2 = op == LBBNN_PACK_PAIR ? b[r * cols + c] : (op == LBBNN_PACK_SQUARE ? v1 * v1 : v1 * b[r * cols + c]);
      if (o1) o1[r * cols + c] = __float2bfloat16_rn(v1);
      if (o2) o2[r * cols + c] = __float2bfloat16_rn(v2);
    }
    t1[ty + i * 8][tx] = v1;
    t2[ty + i * 8][tx] = v2;
  }
  if (o1T == nullptr) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t c = c0 + ty + i * 8, r = r0 + tx;   // output (cols, rows): row index c, column index r
    if (c < cols && r < rows) {
      o1T[c * rows + r] = __float2bfloat16_rn(t1[tx][ty + i * 8]);
      o2T[c * rows + r] = __float2bfloat16_rn(t2[tx][ty + i * 8]);
    }
  }
}

// column sums in two deterministic stages: stage 1 = grid (cols/32, S row slices), each block sums its
// slice with 8 row-lanes per column and a fixed-order smem reduction -> partial[s][2][cols];
// stage 2 sums the S partials in order.
template <bool BF16>
__global__ void __launch_bounds__(256) colsum2_stage1(const void* __restrict__ a_, const void* __restrict__ b_, int64_t rows,
                                                      int64_t cols, int64_t rows_per_slice, float* __restrict__ partial) {
  __shared__ float s1[8][33], s2[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + tx;
  const int64_t rbeg = (int64_t)blockIdx.y * rows_per_slice, rend = min(rows, rbeg + rows_per_slice);
  float acc1 = 0.f, acc2 = 0.f;
  if (c < cols) {
#pragma unroll 4
    for (int64_t r = rbeg + ty; r < rend; r += 8) {
      if (BF16) {
        acc1 += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a_)[r * cols + c]);
        acc2 += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(b_)[r * cols + c]);
      } else {
        const float v = __ldg(reinterpret_cast<const float*>(a_) + r * cols + c);
        acc1 += v;
        acc2 += b_ ? v * __ldg(reinterpret_cast<const float*>(b_) + r * cols + c) : 0.f;
      }
    }
  }
  s1[ty][tx] = acc1;
  s2[ty][tx] = acc2;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float r1 = 0.f, r2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { r1 += s1[i][tx]; r2 += s2[i][tx]; }
    partial[((int64_t)blockIdx.y * 2 + 0) * cols + c] = r1;
    partial[((int64_t)blockIdx.y * 2 + 1) * cols + c] = r2;
  }
}

__global__ void __launch_bounds__(256) colsum2_stage2(const float* __restrict__ partial, int slices, int64_t cols,
                                                      float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over 2*cols
  if (i >= 2 * cols) return;
  const int64_t which = i / cols, c = i % cols;
  float acc = 0.f;
  for (int s = 0; s < slices; ++s) acc += partial[((int64_t)s * 2 + which) * cols + c];
  out[i] = acc;
}

int colsum_slices(int64_t rows) {
  int64_t s = rows / 128;
  return (int)(s < 1 ? 1 : (s > 64 ? 64 : s));
}

}  // namespace
}  // namespace lbbnn

using namespace lbbnn;

extern "C" int lbbnn_bf16_pack(const float* a, const float* b, int op, int64_t rows, int64_t cols, void* out1, void* out2,
                               void* out1T, void* out2T, lbbnn_stream s) {
  LBBNN_REQUIRE(a && rows > 0 && cols > 0, "bad input");
  LBBNN_REQUIRE(op == LBBNN_PACK_SQUARE || b, "second operand required");
  LBBNN_REQUIRE((out1T == nullptr) == (out2T == nullptr), "transposed outputs come in pairs");
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
  LBBNN_REQUIRE(grid.y <= 65535, "too many rows for one launch (%lld)", (long long)rows);
  bf16_pack_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(a, b, op, rows, cols, (__nv_bfloat16*)out1, (__nv_bfloat16*)out2,
                                                     (__nv_bfloat16*)out1T, (__nv_bfloat16*)out2T);
  return check_launch("bf16_pack");
}

extern "C" size_t lbbnn_colsum2_workspace_bytes(int64_t rows, int64_t cols) {
  if (rows <= 0 || cols <= 0) return 0;
  return (size_t)colsum_slices(rows) * 2 * cols * sizeof(float);
}

extern "C" int lbbnn_colsum2(const void* a, const void* b, int a_is_bf16, int64_t rows, int64_t cols, float* out,
                             void* ws, size_t ws_bytes, lbbnn_stream s) {
  LBBNN_REQUIRE(a && out && rows > 0 && cols > 0, "bad input");
  LBBNN_REQUIRE(!a_is_bf16 || b, "bf16 mode needs both tensors");
  LBBNN_REQUIRE(ws && ws_bytes >= lbbnn_colsum2_workspace_bytes(rows, cols), "colsum2 workspace too small");
  const int slices = colsum_slices(rows);
  const int64_t rps = ceil_div(rows, slices);
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)slices);
  if (a_is_bf16) colsum2_stage1<true><<<grid, 256, 0, (cudaStream_t)s>>>(a, b, rows, cols, rps, (float*)ws);
  else colsum2_stage1<false><<<grid, 256, 0, (cudaStream_t)s>>>(a, b, rows, cols, rps, (float*)ws);
  if (int rc = check_launch("colsum2_stage1")) return rc;
  colsum2_stage2<<<(unsigned)ceil_div(2 * cols, 256), 256, 0, (cudaStream_t)s>>>((const float*)ws, slices, cols, out);
  return check_launch("colsum2_stage2");
}
