// Error plumbing, noise export, loss head and optimizer kernels of liblbbnn.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace lbbnn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return LBBNN_ERR_CUDA;
  }
  return LBBNN_OK;
}

int sm_count() {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) == cudaSuccess &&
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
    cached = n;
  } else {
    cudaGetLastError();
    return 148;  // B200; only reached by size queries on a machine without a GPU
  }
  return cached;
}

namespace {

__global__ void philox_export_kernel(float* __restrict__ out, int64_t n, uint64_t seed, uint64_t stream, int normal) {
  const int64_t nquads = ceil_div(n, 4);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += (int64_t)gridDim.x * blockDim.x) {
    float v[4];
    if (normal) philox_normal4(seed, stream, (uint64_t)q, v);
    else philox_uniform4(seed, stream, (uint64_t)q, v);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (q * 4 + j < n) out[q * 4 + j] = v[j];
  }
}

int philox_export(float* out, int64_t n, uint64_t seed, uint64_t stream_id, int normal, lbbnn_stream s) {
  LBBNN_REQUIRE(out && n >= 0, "bad output");
  if (n == 0) return LBBNN_OK;
  int64_t blocks = ceil_div(ceil_div(n, 4), 256);
  if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
  philox_export_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(out, n, seed, stream_id, normal);
  return check_launch("philox_export");
}

// one warp per row: log_softmax(dim=1), nll(sum) and its gradient (LRT:210,223)
__global__ void logsoftmax_nll_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, int64_t B,
                                      int64_t C, float* __restrict__ logp, float* __restrict__ nll_sum,
                                      float* __restrict__ dlogits, float gscale) {
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float local = 0.f;
  for (int64_t b = warp; b < B; b += nw) {  // single block: fixed summation order
    const float* row = logits + b * C;
    float mx = -INFINITY;
    for (int64_t c = lane; c < C; c += 32) mx = fmaxf(mx, row[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    for (int64_t c = lane; c < C; c += 32) se += expf(row[c] - mx);
    se = warp_sum(se);
    const float lse = mx + logf(se);
    const int64_t t = target[b];
    for (int64_t c = lane; c < C; c += 32) {
      const float lp = row[c] - lse;
      if (logp) logp[b * C + c] = lp;
      if (dlogits) dlogits[b * C + c] = gscale * (expf(lp) - (c == t ? 1.0f : 0.0f));
      if (c == t) local -= lp;
    }
  }
  const float tot = block_sum(local, red);
  if (threadIdx.x == 0 && nll_sum) *nll_sum = tot;
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                            const int64_t* __restrict__ step_dev) {
  const double t = (double)(*step_dev + 1);
  const float bc1 = (float)(1.0 - pow((double)b1, t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, t));
  const float step_size = lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * (1.0f - b1);   // lerp form, as torch's _single_tensor_adam
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= step_size * (mi / denom);
  }
}

__global__ void counter_inc_kernel(int64_t* c) { *c += 1; }

}  // namespace
}  // namespace lbbnn

using namespace lbbnn;

extern "C" const char* lbbnn_last_error(void) { return g_err; }
extern "C" int lbbnn_abi_version(void) { return LBBNN_ABI_VERSION; }

extern "C" int lbbnn_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10;
}

extern "C" int lbbnn_philox_normal(float* out, int64_t n, uint64_t seed, uint64_t stream_id, lbbnn_stream s) {
  return philox_export(out, n, seed, stream_id, 1, s);
}
extern "C" int lbbnn_philox_uniform(float* out, int64_t n, uint64_t seed, uint64_t stream_id, lbbnn_stream s) {
  return philox_export(out, n, seed, stream_id, 0, s);
}

extern "C" int lbbnn_logsoftmax_nll_f32(const float* logits, const int64_t* target, int64_t B, int64_t C, float* logp,
                                        float* nll_sum, float* dlogits, float grad_scale, lbbnn_stream s) {
  LBBNN_REQUIRE(logits && target && B > 0 && C > 0, "bad logits/target");
  logsoftmax_nll_kernel<<<1, 1024, 0, (cudaStream_t)s>>>(logits, target, B, C, logp, nll_sum, dlogits, grad_scale);
  return check_launch("logsoftmax_nll");
}

extern "C" int lbbnn_adam_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                              float beta1, float beta2, float eps, const int64_t* step_dev, lbbnn_stream s) {
  LBBNN_REQUIRE(param && grad && exp_avg && exp_avg_sq && step_dev && n > 0, "NULL argument");
  int64_t blocks = ceil_div(n, 256 * 4);
  if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
  adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                             step_dev);
  return check_launch("adam");
}

extern "C" int lbbnn_counter_inc(int64_t* counter, lbbnn_stream s) {
  LBBNN_REQUIRE(counter, "NULL counter");
  counter_inc_kernel<<<1, 1, 0, (cudaStream_t)s>>>(counter);
  return check_launch("counter_inc");
}
