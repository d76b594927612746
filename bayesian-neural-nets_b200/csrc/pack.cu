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

// one block per 32 columns; 8 row-lanes per column accumulate strided rows, then a fixed-order
// smem reduction: deterministic.
template <bool BF16>
__global__ void __launch_bounds__(256) colsum2_kernel(const void* __restrict__ a_, const void* __restrict__ b_, int64_t rows,
                                                      int64_t cols, float* __restrict__ out) {
  __shared__ float s1[8][33], s2[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + tx;
  float acc1 = 0.f, acc2 = 0.f;
  if (c < cols) {
    for (int64_t r = ty; r < rows; r += 8) {
      if (BF16) {
        acc1 += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a_)[r * cols + c]);
        acc2 += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(b_)[r * cols + c]);
      } else {
        const float v = reinterpret_cast<const float*>(a_)[r * cols + c];
        acc1 += v;
        acc2 += b_ ? v * reinterpret_cast<const float*>(b_)[r * cols + c] : 0.f;
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
    out[c] = r1;
    out[cols + c] = r2;
  }
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

extern "C" int lbbnn_colsum2(const void* a, const void* b, int a_is_bf16, int64_t rows, int64_t cols, float* out,
                             lbbnn_stream s) {
  LBBNN_REQUIRE(a && out && rows > 0 && cols > 0, "bad input");
  LBBNN_REQUIRE(!a_is_bf16 || b, "bf16 mode needs both tensors");
  const unsigned grid = (unsigned)ceil_div(cols, 32);
  if (a_is_bf16) colsum2_kernel<true><<<grid, 256, 0, (cudaStream_t)s>>>(a, b, rows, cols, out);
  else colsum2_kernel<false><<<grid, 256, 0, (cudaStream_t)s>>>(a, b, rows, cols, out);
  return check_launch("colsum2");
}
