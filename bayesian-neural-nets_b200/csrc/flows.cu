// Masked coupling flows of flows2.py as fused small-MLP kernels: PropagateFlow (flows2:14-46) over
//   RNVP (flows2:188-219):  m ~ Bern(.5);  y = MLP(m z);  g = sigmoid(s(y));  x = (1-m) z g + (1-g) t(y) + m z
//   MNF/IAF (flows2:225-241): h = tanh(f(m z));           g = sigmoid(k(h));  x = (1-m)(z g + (1-g) mu(h)) + m z
// log_det = sum (1-m) log g.   One CTA per row of z runs the WHOLE stack of transforms: every Linear is a
// GEMV whose output neurons are spread over the warps (lanes stride the contiguous weight row -> coalesced,
// shuffle reduction), activations live in shared memory between the fused layers.  In the MNF layer only
// one row per stack evaluation is live, and all flow evaluations of a step depend on
// parameters and noise only -- the host batches them as rows of one launch.
// The backward kernel mirrors it: per row it writes the parameter gradients of ITS evaluation into its own
// slice of a (rows, n_params) buffer with plain coalesced stores (no atomics; rows are summed afterwards).
#include <string.h>

#include "common.cuh"

namespace lbbnn {
namespace {

constexpr int kFlowThreads = 512;
constexpr int kWarps = kFlowThreads / 32;
constexpr int kMaxH = 128;   // widest hidden layer supported (flows2 uses 75 / 100, flows_simstudy 50)

struct Lin { const float* W; const float* b; float* dW; float* db; int in, out; };   // nn.Linear: W is (out, in)

struct FlowDev {
  int kind, dim, n_transforms, n_hidden;
  Lin hidden[LBBNN_FLOW_MAX_T][LBBNN_FLOW_MAX_HIDDEN];
  Lin shift[LBBNN_FLOW_MAX_T], scale[LBBNN_FLOW_MAX_T];
  int64_t grad_row_stride;   // floats between the gradient slices of consecutive rows
  int save_stride;           // floats saved per (row, transform): zin[D] gate[D] shift[D] h[sum H]
};

__device__ __forceinline__ float act_fwd(int kind, bool last, float v) {
  if (kind == LBBNN_FLOW_IAF) return tanhf(v);
  return last ? v : (v > 0.f ? v : 0.1f * v);          // LeakyReLU(0.1); the MLP drops the last activation (flows2:185)
}
__device__ __forceinline__ float act_bwd(int kind, bool last, float h) {   // derivative from the post-activation value
  if (kind == LBBNN_FLOW_IAF) return 1.0f - h * h;
  return last ? 1.0f : (h > 0.f ? 1.0f : 0.1f);
}

__device__ __forceinline__ float mask_of(const float* __restrict__ masks, const Noise& nz, int t, int64_t r, int64_t R, int d, int D) {
  if (masks) return masks[((int64_t)t * R + r) * D + d];
  return philox_uniform1(nz.seed, nz.stream + (uint64_t)t, (uint64_t)r * (uint64_t)D + (uint64_t)d) < 0.5f ? 1.0f : 0.0f;
}

// out[j] = act(b[j] + sum_i W[j,i] v[i]) for j over the warps; v in smem (16-byte aligned).  Two output neurons per
// warp pass and float4 loads when the row length allows: 8 independent 16-byte loads in flight per lane.
__device__ __forceinline__ void gemv_warp(const Lin& L, const float* __restrict__ v, float* __restrict__ out, int kind, bool last,
                                          float* __restrict__ save) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec = (L.in % 4 == 0) && aligned16(L.W) && aligned16(v);
  for (int j = warp * 2; j < L.out; j += kWarps * 2) {
    const bool two = j + 1 < L.out;
    const float* w0 = L.W + (int64_t)j * L.in;
    const float* w1 = w0 + (two ? L.in : 0);
    float a0 = 0.f, a1 = 0.f;
    if (vec) {
      const float4* p0 = reinterpret_cast<const float4*>(w0);
      const float4* p1 = reinterpret_cast<const float4*>(w1);
      const float4* pv = reinterpret_cast<const float4*>(v);
      const int n4 = L.in >> 2;
#pragma unroll 4
      for (int i = lane; i < n4; i += 32) {
        const float4 x = pv[i], q0 = __ldg(p0 + i), q1 = __ldg(p1 + i);
        a0 = fmaf(q0.x, x.x, fmaf(q0.y, x.y, fmaf(q0.z, x.z, fmaf(q0.w, x.w, a0))));
        a1 = fmaf(q1.x, x.x, fmaf(q1.y, x.y, fmaf(q1.z, x.z, fmaf(q1.w, x.w, a1))));
      }
    } else {
#pragma unroll 4
      for (int i = lane; i < L.in; i += 32) {
        const float x = v[i];
        a0 = fmaf(__ldg(w0 + i), x, a0);
        a1 = fmaf(__ldg(w1 + i), x, a1);
      }
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) {
      const float h0 = act_fwd(kind, last, a0 + __ldg(L.b + j));
      out[j] = h0;
      if (save) save[j] = h0;
      if (two) {
        const float h1 = act_fwd(kind, last, a1 + __ldg(L.b + j + 1));
        out[j + 1] = h1;
        if (save) save[j + 1] = h1;
      }
    }
  }
}

// ---- forward -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFlowThreads) flow_fwd_kernel(const FlowDev f, const float* __restrict__ z_in,
                                                                const float* __restrict__ masks, const Noise mask_noise,
                                                                int64_t R, float* __restrict__ z_out,
                                                                float* __restrict__ logdet, float* __restrict__ save) {
  extern __shared__ __align__(16) float sm[];
  const int D = f.dim;
  float* zs = sm;             // [D] current z
  float* xm = sm + D;         // [D] m * z
  float* ms = sm + 2 * D;     // [D] this transform's mask
  float* ha = sm + 3 * D;     // [kMaxH]
  float* hb = ha + kMaxH;     // [kMaxH]
  __shared__ float red[32];
  Noise nz = mask_noise;
  nz.resolve();
  const int64_t r = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int d = tid; d < D; d += kFlowThreads) zs[d] = z_in[r * D + d];
  __syncthreads();
  float ld_total = 0.f;
  for (int t = 0; t < f.n_transforms; ++t) {
    float* sv = save ? save + ((int64_t)r * f.n_transforms + t) * f.save_stride : nullptr;
    for (int d = tid; d < D; d += kFlowThreads) {
      const float m = mask_of(masks, nz, t, r, R, d, D);
      ms[d] = m;
      xm[d] = m * zs[d];
      if (sv) sv[d] = zs[d];
    }
    __syncthreads();
    const float* v = xm;
    float* cur = ha;
    float* svh = sv ? sv + 3 * D : nullptr;
    for (int l = 0; l < f.n_hidden; ++l) {
      gemv_warp(f.hidden[t][l], v, cur, f.kind, l == f.n_hidden - 1, svh);
      if (svh) svh += f.hidden[t][l].out;
      __syncthreads();
      v = cur;
      cur = (cur == ha) ? hb : ha;
    }
    // shift / scale heads and the coupling: one output dim per thread (H <= 128 inputs: each thread streams its two
    // weight rows through L1, no shuffle reductions)
    const Lin& Ls = f.shift[t];
    const Lin& Lc = f.scale[t];
    const int H = Ls.in;
    float ld = 0.f;
    for (int d = tid; d < D; d += kFlowThreads) {
      const float* ws = Ls.W + (int64_t)d * H;
      const float* wc = Lc.W + (int64_t)d * H;
      float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f;
      int i = 0;
      for (; i + 1 < H; i += 2) {
        a1 = fmaf(__ldg(ws + i), v[i], a1);
        b1 = fmaf(__ldg(ws + i + 1), v[i + 1], b1);
        a2 = fmaf(__ldg(wc + i), v[i], a2);
        b2 = fmaf(__ldg(wc + i + 1), v[i + 1], b2);
      }
      if (i < H) {
        a1 = fmaf(__ldg(ws + i), v[i], a1);
        a2 = fmaf(__ldg(wc + i), v[i], a2);
      }
      const float sh = (a1 + b1) + __ldg(Ls.b + d), g = 1.0f / (1.0f + expf(-((a2 + b2) + __ldg(Lc.b + d))));
      const float z = zs[d], m = ms[d];
      float x;
      if (f.kind == LBBNN_FLOW_RNVP) x = (1.0f - m) * z * g + (1.0f - g) * sh + m * z;       // flows2:215
      else x = m * z + (1.0f - m) * (z * g + (1.0f - g) * sh);                               // flows2:238
      zs[d] = x;
      ld += (1.0f - m) * logf(g);
      if (sv) { sv[D + d] = g; sv[2 * D + d] = sh; }
    }
    const float tot = block_sum(ld, red);
    if (tid == 0) ld_total += tot;
    __syncthreads();
  }
  for (int d = tid; d < D; d += kFlowThreads) z_out[r * D + d] = zs[d];
  if (tid == 0) logdet[r] = ld_total;
}

// ---- backward ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFlowThreads) flow_bwd_kernel(const FlowDev f, const float* __restrict__ masks,
                                                                const Noise mask_noise, int64_t R,
                                                                const float* __restrict__ dz_out,
                                                                const float* __restrict__ dlogdet,
                                                                const float* __restrict__ save, float* __restrict__ dz_in) {
  extern __shared__ __align__(16) float sm[];
  const int D = f.dim;
  float* dz = sm;                 // [D] gradient wrt the current transform's output, then its input
  float* dsh = sm + D;            // [D]
  float* dsc = sm + 2 * D;        // [D]
  float* xm = sm + 3 * D;         // [D] m * zin (input of the conditioner net)
  float* ms = sm + 4 * D;         // [D] this transform's mask
  float* da = sm + 5 * D;         // [kMaxH] gradient wrt a hidden layer's pre-activation
  float* dh = da + kMaxH;         // [kMaxH] gradient wrt a hidden layer's output
  float* part = dh + kMaxH;       // [kWarps][kMaxH]
  Noise nz = mask_noise;
  nz.resolve();
  const int64_t r = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float dld = dlogdet ? dlogdet[r] : 0.f;
  const int64_t goff = r * f.grad_row_stride;
  for (int d = tid; d < D; d += kFlowThreads) dz[d] = dz_out ? dz_out[r * D + d] : 0.f;
  __syncthreads();
  for (int t = f.n_transforms - 1; t >= 0; --t) {
    const float* sv = save + ((int64_t)r * f.n_transforms + t) * f.save_stride;
    const float* zin = sv;
    const float* gate = sv + D;
    const float* shf = sv + 2 * D;
    const float* hsave = sv + 3 * D;
    const Lin& Ls = f.shift[t];
    const Lin& Lc = f.scale[t];
    const int H = Ls.in;
    int hoff_last = 0;
    for (int l = 0; l + 1 < f.n_hidden; ++l) hoff_last += f.hidden[t][l].out;
    const float* y = hsave + hoff_last;   // output of the conditioner net
    // coupling backward (elementwise over D)
    for (int d = tid; d < D; d += kFlowThreads) {
      const float m = mask_of(masks, nz, t, r, R, d, D);
      ms[d] = m;
      const float z = zin[d], g = gate[d], sh = shf[d], dx = dz[d];
      float dg, ds, dzd;
      if (f.kind == LBBNN_FLOW_RNVP) {
        dg = dx * ((1.0f - m) * z - sh) + dld * (1.0f - m) / g;
        ds = dx * (1.0f - g);
      } else {
        dg = dx * (1.0f - m) * (z - sh) + dld * (1.0f - m) / g;
        ds = dx * (1.0f - m) * (1.0f - g);
      }
      dzd = dx * ((1.0f - m) * g + m);
      dsh[d] = ds;
      dsc[d] = dg * g * (1.0f - g);
      dz[d] = dzd;            // direct path; the path through the conditioner is added below
      xm[d] = m * z;
      Ls.db[goff + d] = ds;
      Lc.db[goff + d] = dsc[d];
    }
    __syncthreads();
    // head weight gradients: outer products dsh x y, dsc x y
    for (int d = warp; d < D; d += kWarps) {      // rows over the warps, lanes along the (contiguous) row
      const float a = dsh[d], c = dsc[d];
      float* ps = Ls.dW + goff + (int64_t)d * H;
      float* pc = Lc.dW + goff + (int64_t)d * H;
      for (int i = lane; i < H; i += 32) {
        const float yi = y[i];
        ps[i] = a * yi;
        pc[i] = c * yi;
      }
    }
    // dy[i] = sum_d Wt[d,i] dsh[d] + Ws[d,i] dsc[d]: warps take slices of d, lanes run over i (coalesced rows)
    {
      float acc[kMaxH / 32];
#pragma unroll
      for (int k = 0; k < kMaxH / 32; ++k) acc[k] = 0.f;
      for (int d = warp; d < D; d += kWarps) {
        const float a = dsh[d], c = dsc[d];
        const float* ws = Ls.W + (int64_t)d * H;
        const float* wc = Lc.W + (int64_t)d * H;
#pragma unroll
        for (int k = 0; k < kMaxH / 32; ++k) {
          const int i = lane + 32 * k;
          if (i < H) acc[k] = fmaf(__ldg(ws + i), a, fmaf(__ldg(wc + i), c, acc[k]));
        }
      }
#pragma unroll
      for (int k = 0; k < kMaxH / 32; ++k) part[warp * kMaxH + lane + 32 * k] = acc[k];
    }
    __syncthreads();
    for (int i = tid; i < H; i += kFlowThreads) {
      float s = 0.f;
      for (int w = 0; w < kWarps; ++w) s += part[w * kMaxH + i];
      dh[i] = s;
    }
    __syncthreads();
    // conditioner net backward, last hidden layer first
    int hoff = hoff_last;
    for (int l = f.n_hidden - 1; l >= 0; --l) {
      const Lin& L = f.hidden[t][l];
      const float* hout = hsave + hoff;                                  // this layer's post-activation output
      const float* vin = (l == 0) ? xm : (hsave + hoff - f.hidden[t][l - 1].out);   // its input
      const bool vin_smem = (l == 0);
      for (int j = tid; j < L.out; j += kFlowThreads) {
        const float g = dh[j] * act_bwd(f.kind, l == f.n_hidden - 1, hout[j]);
        da[j] = g;
        L.db[goff + j] = g;
      }
      __syncthreads();
      for (int j = warp; j < L.out; j += kWarps) {
        const float a = da[j];
        float* pw = L.dW + goff + (int64_t)j * L.in;
        const float* src = vin_smem ? xm : vin;
        for (int i = lane; i < L.in; i += 32) pw[i] = a * src[i];
      }
      // gradient wrt the layer input: dv[i] = sum_j W[j,i] da[j]; threads over i (coalesced), loop over j
      if (l == 0) {
        for (int i = tid; i < L.in; i += kFlowThreads) {
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          int j = 0;
          for (; j + 3 < L.out; j += 4) {
            s0 = fmaf(__ldg(L.W + (int64_t)(j + 0) * L.in + i), da[j + 0], s0);
            s1 = fmaf(__ldg(L.W + (int64_t)(j + 1) * L.in + i), da[j + 1], s1);
            s2 = fmaf(__ldg(L.W + (int64_t)(j + 2) * L.in + i), da[j + 2], s2);
            s3 = fmaf(__ldg(L.W + (int64_t)(j + 3) * L.in + i), da[j + 3], s3);
          }
          for (; j < L.out; ++j) s0 = fmaf(__ldg(L.W + (int64_t)j * L.in + i), da[j], s0);
          dz[i] += ms[i] * ((s0 + s1) + (s2 + s3));                       // input of the net was m * z
        }
      } else {
        for (int i = tid; i < L.in; i += kFlowThreads) {
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          int j = 0;
          for (; j + 3 < L.out; j += 4) {
            s0 = fmaf(__ldg(L.W + (int64_t)(j + 0) * L.in + i), da[j + 0], s0);
            s1 = fmaf(__ldg(L.W + (int64_t)(j + 1) * L.in + i), da[j + 1], s1);
            s2 = fmaf(__ldg(L.W + (int64_t)(j + 2) * L.in + i), da[j + 2], s2);
            s3 = fmaf(__ldg(L.W + (int64_t)(j + 3) * L.in + i), da[j + 3], s3);
          }
          for (; j < L.out; ++j) s0 = fmaf(__ldg(L.W + (int64_t)j * L.in + i), da[j], s0);
          part[i] = (s0 + s1) + (s2 + s3);
        }
        __syncthreads();
        for (int i = tid; i < L.in; i += kFlowThreads) dh[i] = part[i];
        hoff -= f.hidden[t][l - 1].out;
      }
      __syncthreads();
    }
  }
  for (int d = tid; d < D; d += kFlowThreads) dz_in[r * D + d] = dz[d];
}

int to_dev(const lbbnn_flow* F, const lbbnn_flow_grads* G, FlowDev* out) {
  LBBNN_REQUIRE(F != nullptr, "flow is NULL");
  LBBNN_REQUIRE(F->kind == LBBNN_FLOW_RNVP || F->kind == LBBNN_FLOW_IAF, "unknown flow kind %d", F->kind);
  LBBNN_REQUIRE(F->dim > 0 && F->n_transforms > 0 && F->n_transforms <= LBBNN_FLOW_MAX_T, "bad flow shape");
  LBBNN_REQUIRE(F->n_hidden > 0 && F->n_hidden <= LBBNN_FLOW_MAX_HIDDEN, "bad number of hidden layers %d", F->n_hidden);
  FlowDev d;
  memset(&d, 0, sizeof(d));
  d.kind = F->kind; d.dim = F->dim; d.n_transforms = F->n_transforms; d.n_hidden = F->n_hidden;
  int hsum = 0;
  for (int t = 0; t < F->n_transforms; ++t) {
    int prev = F->dim, hs = 0;
    for (int l = 0; l < F->n_hidden; ++l) {
      const lbbnn_flow_linear& s = F->t[t].hidden[l];
      LBBNN_REQUIRE(s.W && s.b && s.in == prev && s.out > 0 && s.out <= kMaxH, "bad hidden layer %d of transform %d", l, t);
      d.hidden[t][l] = Lin{s.W, s.b, G ? G->t[t].hidden[l].dW : nullptr, G ? G->t[t].hidden[l].db : nullptr, s.in, s.out};
      prev = s.out;
      hs += s.out;
    }
    const lbbnn_flow_linear& a = F->t[t].shift;
    const lbbnn_flow_linear& c = F->t[t].scale;
    LBBNN_REQUIRE(a.W && a.b && c.W && c.b && a.in == prev && c.in == prev && a.out == F->dim && c.out == F->dim, "bad heads of transform %d", t);
    d.shift[t] = Lin{a.W, a.b, G ? G->t[t].shift.dW : nullptr, G ? G->t[t].shift.db : nullptr, a.in, a.out};
    d.scale[t] = Lin{c.W, c.b, G ? G->t[t].scale.dW : nullptr, G ? G->t[t].scale.db : nullptr, c.in, c.out};
    if (t == 0) hsum = hs;
    LBBNN_REQUIRE(hs == hsum, "all transforms of a flow must share one architecture");
  }
  d.save_stride = 3 * F->dim + hsum;
  d.grad_row_stride = G ? G->row_stride : 0;
  *out = d;
  return LBBNN_OK;
}

}  // namespace
}  // namespace lbbnn

using namespace lbbnn;

extern "C" size_t lbbnn_flow_save_floats(const lbbnn_flow* F, int64_t rows) {
  if (!F || rows <= 0) return 0;
  size_t hsum = 0;
  for (int l = 0; l < F->n_hidden; ++l) hsum += (size_t)F->t[0].hidden[l].out;
  return (size_t)rows * F->n_transforms * (3 * (size_t)F->dim + hsum);
}

extern "C" int lbbnn_flow_fwd(const lbbnn_flow* F, const float* z_in, int64_t rows, const float* masks,
                              const lbbnn_noise* mask_u, float* z_out, float* logdet, float* save, lbbnn_stream s) {
  FlowDev d;
  if (int rc = to_dev(F, nullptr, &d)) return rc;
  LBBNN_REQUIRE(z_in && z_out && logdet && rows > 0, "NULL argument");
  const size_t smem = (size_t)(3 * d.dim + 2 * kMaxH) * sizeof(float);
  if (smem > 48 * 1024) LBBNN_CUDA(cudaFuncSetAttribute(flow_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  flow_fwd_kernel<<<(unsigned)rows, kFlowThreads, smem, (cudaStream_t)s>>>(d, z_in, masks, make_noise(mask_u), rows, z_out, logdet, save);
  return check_launch("flow_fwd");
}

extern "C" int lbbnn_flow_bwd(const lbbnn_flow* F, const lbbnn_flow_grads* G, int64_t rows, const float* masks,
                              const lbbnn_noise* mask_u, const float* dz_out, const float* dlogdet, const float* save,
                              float* dz_in, lbbnn_stream s) {
  FlowDev d;
  LBBNN_REQUIRE(G != nullptr, "flow grads NULL");
  if (int rc = to_dev(F, G, &d)) return rc;
  LBBNN_REQUIRE(save && dz_in && rows > 0, "NULL argument");
  const size_t smem = (size_t)(5 * d.dim + 2 * kMaxH + kWarps * kMaxH) * sizeof(float);
  if (smem > 48 * 1024) LBBNN_CUDA(cudaFuncSetAttribute(flow_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  flow_bwd_kernel<<<(unsigned)rows, kFlowThreads, smem, (cudaStream_t)s>>>(d, masks, make_noise(mask_u), rows, dz_out, dlogdet, save, dz_in);
  return check_launch("flow_bwd");
}
