// Masked coupling flows of flows2.py as fused small-MLP kernels: PropagateFlow (flows2:14-46) over
//   RNVP (flows2:188-219):  m ~ Bern(.5);  y = MLP(m z);  g = sigmoid(s(y));  x = (1-m) z g + (1-g) t(y) + m z
//   MNF/IAF (flows2:225-241): h = tanh(f(m z));           g = sigmoid(k(h));  x = (1-m)(z g + (1-g) mu(h)) + m z
// log_det = sum (1-m) log g.   One thread-block CLUSTER (up to 8 CTAs) per row of z runs the WHOLE stack of
// transforms: every Linear is a GEMV with one warp per output neuron (lanes stride the contiguous weight row ->
// coalesced, shuffle reduction); the D-wide layers are split across the CTAs of the cluster, which exchange their
// results through distributed shared memory and a cluster barrier; activations never leave shared memory between
// the fused layers.  In the MNF layer only one row per stack evaluation is live, and all
// flow evaluations of a step depend on parameters and noise only -- the host batches them as rows of one launch.
// The backward kernel mirrors it: per row it writes the parameter gradients of ITS evaluation into its own
// slice of a (rows, n_params) buffer with plain coalesced stores (no atomics; rows are summed afterwards).
#include <cooperative_groups.h>
#include <string.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace lbbnn {
namespace {

constexpr int kMaxCluster = 8;   // portable cluster size
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

// One warp per output neuron j of a caller-chosen subset (j = first, first + step, ...): out[j] = act(b[j] + W[j,:] v).
// `bcast` > 0: the result is written into the `out` array of EVERY CTA of the cluster (distributed shared memory).
__device__ __forceinline__ void gemv_rows(const Lin& L, const float* __restrict__ v, float* __restrict__ out, int kind, bool last,
                                          float* __restrict__ save, int first, int step, cg::cluster_group& cl, int bcast) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec = (L.in % 4 == 0) && aligned16(L.W) && aligned16(v);
  for (int j = first + warp * step; j < L.out; j += kWarps * step) {
    const float* w0 = L.W + (int64_t)j * L.in;
    float a0 = 0.f, a1 = 0.f;
    if (vec) {
      const float4* p0 = reinterpret_cast<const float4*>(w0);
      const float4* pv = reinterpret_cast<const float4*>(v);
      const int n4 = L.in >> 2;
      int i = lane;
      for (; i + 32 < n4; i += 64) {
        const float4 x = pv[i], q = __ldg(p0 + i), y = pv[i + 32], u = __ldg(p0 + i + 32);
        a0 = fmaf(q.x, x.x, fmaf(q.y, x.y, fmaf(q.z, x.z, fmaf(q.w, x.w, a0))));
        a1 = fmaf(u.x, y.x, fmaf(u.y, y.y, fmaf(u.z, y.z, fmaf(u.w, y.w, a1))));
      }
      if (i < n4) {
        const float4 x = pv[i], q = __ldg(p0 + i);
        a0 = fmaf(q.x, x.x, fmaf(q.y, x.y, fmaf(q.z, x.z, fmaf(q.w, x.w, a0))));
      }
    } else {
      for (int i = lane; i < L.in; i += 32) a0 = fmaf(__ldg(w0 + i), v[i], a0);
    }
    a0 = warp_sum(a0 + a1);
    const float h = act_fwd(kind, last, a0 + __ldg(L.b + j));
    if (bcast > 0) {
      if (lane < bcast) cl.map_shared_rank(out, lane)[j] = h;
    } else if (lane == 0) {
      out[j] = h;
    }
    if (save && lane == 0) save[j] = h;
  }
}

// ---- forward -------------------------------------------------------------------------------------------
// One CLUSTER of C CTAs per row of z.  Every CTA keeps the whole z in shared memory; the two big GEMVs of a transform
// are split across the cluster -- the first hidden layer by output neuron (results broadcast into every CTA's shared
// memory through DSMEM), the shift / scale heads and the coupling by output dimension (each CTA broadcasts its slice
// of the new z) -- with one cluster barrier after each.  The small hidden layers are recomputed by every CTA.
__global__ void __launch_bounds__(kFlowThreads) flow_fwd_kernel(const FlowDev f, const float* __restrict__ z_in,
                                                                const float* __restrict__ masks, const Noise mask_noise,
                                                                int64_t R, float* __restrict__ z_out,
                                                                float* __restrict__ logdet, float* __restrict__ save) {
  extern __shared__ __align__(16) float sm[];
  cg::cluster_group cl = cg::this_cluster();
  const int C = (int)cl.num_blocks(), c = (int)cl.block_rank();
  const int D = f.dim;
  const int Dp = (D + 3) & ~3;
  float* zs = sm;              // [D] current z
  float* xm = sm + Dp;         // [D] m * z
  float* ms = sm + 2 * Dp;     // [D] this transform's mask
  float* ha = sm + 3 * Dp;     // [kMaxH]
  float* hb = ha + kMaxH;      // [kMaxH]
  float* ldp = hb + kMaxH;     // [kMaxCluster] log-det partials of the cluster (read by rank 0)
  __shared__ float red[32];
  Noise nz = mask_noise;
  nz.resolve();
  const int64_t r = blockIdx.x / C;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int per = (D + C - 1) / C, d0 = c * per, d1 = min(D, d0 + per);    // this CTA's slice of the output dims
  for (int d = tid; d < D; d += kFlowThreads) zs[d] = z_in[r * D + d];
  cl.sync();                   // every CTA of the cluster is running before any remote shared-memory access
  float ld_total = 0.f;
  for (int t = 0; t < f.n_transforms; ++t) {
    float* sv = save ? save + ((int64_t)r * f.n_transforms + t) * f.save_stride : nullptr;
    for (int d = tid; d < D; d += kFlowThreads) {
      const float m = mask_of(masks, nz, t, r, R, d, D);
      ms[d] = m;
      xm[d] = m * zs[d];
      if (sv && d >= d0 && d < d1) sv[d] = zs[d];
    }
    __syncthreads();
    float* svh = sv ? sv + 3 * D : nullptr;
    // first hidden layer: neurons c, c + C, ... of this CTA, broadcast to the whole cluster
    gemv_rows(f.hidden[t][0], xm, ha, f.kind, f.n_hidden == 1, svh, c, C, cl, C);
    cl.sync();
    const float* v = ha;
    float* cur = hb;
    if (svh) svh += f.hidden[t][0].out;
    for (int l = 1; l < f.n_hidden; ++l) {
      gemv_rows(f.hidden[t][l], v, cur, f.kind, l == f.n_hidden - 1, c == 0 ? svh : nullptr, 0, 1, cl, 0);
      if (svh) svh += f.hidden[t][l].out;
      __syncthreads();
      v = cur;
      cur = (cur == ha) ? hb : ha;
    }
    // shift / scale heads and the coupling for the dims of this CTA's slice: one warp per dim, lanes over the hidden units
    const Lin& Ls = f.shift[t];
    const Lin& Lc = f.scale[t];
    const int H = Ls.in;
    float ld = 0.f;
    for (int d = d0 + warp; d < d1; d += kWarps) {
      const float* ws = Ls.W + (int64_t)d * H;
      const float* wc = Lc.W + (int64_t)d * H;
      float a1 = 0.f, a2 = 0.f;
      for (int i = lane; i < H; i += 32) {
        const float vi = v[i];
        a1 = fmaf(__ldg(ws + i), vi, a1);
        a2 = fmaf(__ldg(wc + i), vi, a2);
      }
      a1 = warp_sum(a1);
      a2 = warp_sum(a2);
      const float sh = a1 + __ldg(Ls.b + d), g = 1.0f / (1.0f + expf(-(a2 + __ldg(Lc.b + d))));
      const float z = zs[d], m = ms[d];
      float x;
      if (f.kind == LBBNN_FLOW_RNVP) x = (1.0f - m) * z * g + (1.0f - g) * sh + m * z;       // flows2:215
      else x = m * z + (1.0f - m) * (z * g + (1.0f - g) * sh);                               // flows2:238
      if (lane < C) cl.map_shared_rank(zs, lane)[d] = x;          // the new z of this dim, into every CTA
      if (lane == 0) {
        ld += (1.0f - m) * logf(g);
        if (sv) { sv[D + d] = g; sv[2 * D + d] = sh; }
      }
    }
    const float tot = block_sum(ld, red);
    if (tid == 0) cl.map_shared_rank(ldp, 0)[c] = tot;
    cl.sync();
    if (c == 0 && tid == 0)
      for (int k = 0; k < C; ++k) ld_total += ldp[k];             // fixed order
  }
  for (int d = d0 + tid; d < d1; d += kFlowThreads) z_out[r * D + d] = zs[d];
  if (c == 0 && tid == 0) logdet[r] = ld_total;
}

// ---- backward ------------------------------------------------------------------------------------------
// Same cluster per row.  The gradient wrt z only ever travels inside a CTA's own slice of dims (coupling backward ->
// input gradient of the first hidden layer are both per-dim), so the one exchange per transform is the all-reduce of
// the (H,) gradient wrt the conditioner output, whose contraction over D is split across the cluster.  Parameter
// gradients: heads by dim slice, hidden layers by output row (row j belongs to CTA j mod C), into the row's slice of
// the (rows, n_params) buffer with plain stores.
__global__ void __launch_bounds__(kFlowThreads) flow_bwd_kernel(const FlowDev f, const float* __restrict__ masks,
                                                                const Noise mask_noise, int64_t R,
                                                                const float* __restrict__ dz_out,
                                                                const float* __restrict__ dlogdet,
                                                                const float* __restrict__ save, float* __restrict__ dz_in) {
  extern __shared__ __align__(16) float sm[];
  cg::cluster_group cl = cg::this_cluster();
  const int C = (int)cl.num_blocks(), c = (int)cl.block_rank();
  const int D = f.dim;
  const int Dp = (D + 3) & ~3;
  float* dz = sm;                 // [D] gradient wrt the current transform's output, then its input (own slice live)
  float* dsh = sm + Dp;           // [D]
  float* dsc = sm + 2 * Dp;       // [D]
  float* xm = sm + 3 * Dp;        // [D] m * zin (input of the conditioner net), all dims
  float* ms = sm + 4 * Dp;        // [D] this transform's mask, all dims
  float* da = sm + 5 * Dp;        // [kMaxH] gradient wrt a hidden layer's pre-activation
  float* dh = da + kMaxH;         // [kMaxH] gradient wrt a hidden layer's output
  float* part = dh + kMaxH;       // [kWarps][kMaxH]
  float* dyp = part + kWarps * kMaxH;   // [kMaxCluster][kMaxH] per-CTA partials of the conditioner-output gradient
  Noise nz = mask_noise;
  nz.resolve();
  const int64_t r = blockIdx.x / C;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int per = (D + C - 1) / C, d0 = c * per, d1 = min(D, d0 + per);
  const float dld = dlogdet ? dlogdet[r] : 0.f;
  const int64_t goff = r * f.grad_row_stride;
  for (int d = d0 + tid; d < d1; d += kFlowThreads) dz[d] = dz_out ? dz_out[r * D + d] : 0.f;
  cl.sync();
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
    // masks and conditioner input for all dims (the first layer's weight gradient needs them); coupling backward for
    // the dims of this CTA's slice
    for (int d = tid; d < D; d += kFlowThreads) {
      const float m = mask_of(masks, nz, t, r, R, d, D);
      const float z = zin[d];
      ms[d] = m;
      xm[d] = m * z;
      if (d >= d0 && d < d1) {
        const float g = gate[d], sh = shf[d], dx = dz[d];
        float dg, ds;
        if (f.kind == LBBNN_FLOW_RNVP) {
          dg = dx * ((1.0f - m) * z - sh) + dld * (1.0f - m) / g;
          ds = dx * (1.0f - g);
        } else {
          dg = dx * (1.0f - m) * (z - sh) + dld * (1.0f - m) / g;
          ds = dx * (1.0f - m) * (1.0f - g);
        }
        const float dc = dg * g * (1.0f - g);
        dsh[d] = ds;
        dsc[d] = dc;
        dz[d] = dx * ((1.0f - m) * g + m);     // direct path; the path through the conditioner is added below
        Ls.db[goff + d] = ds;
        Lc.db[goff + d] = dc;
      }
    }
    __syncthreads();
    // head weight gradients (outer products dsh x y, dsc x y) and the partial of dy[i] = sum_d Wt[d,i] dsh[d] + Ws[d,i] dsc[d]
    // over this CTA's dims: warps take dims, lanes run over i (coalesced rows)
    {
      float acc[kMaxH / 32];
#pragma unroll
      for (int k = 0; k < kMaxH / 32; ++k) acc[k] = 0.f;
      for (int d = d0 + warp; d < d1; d += kWarps) {
        const float a = dsh[d], cc = dsc[d];
        const float* ws = Ls.W + (int64_t)d * H;
        const float* wc = Lc.W + (int64_t)d * H;
        float* ps = Ls.dW + goff + (int64_t)d * H;
        float* pc = Lc.dW + goff + (int64_t)d * H;
#pragma unroll
        for (int k = 0; k < kMaxH / 32; ++k) {
          const int i = lane + 32 * k;
          if (i < H) {
            const float yi = y[i];
            ps[i] = a * yi;
            pc[i] = cc * yi;
            acc[k] = fmaf(__ldg(ws + i), a, fmaf(__ldg(wc + i), cc, acc[k]));
          }
        }
      }
#pragma unroll
      for (int k = 0; k < kMaxH / 32; ++k) part[warp * kMaxH + lane + 32 * k] = acc[k];
    }
    __syncthreads();
    for (int i = tid; i < H; i += kFlowThreads) {
      float s = 0.f;
      for (int w = 0; w < kWarps; ++w) s += part[w * kMaxH + i];
      for (int k = 0; k < C; ++k) cl.map_shared_rank(dyp, k)[c * kMaxH + i] = s;   // this CTA's partial, into every CTA
    }
    cl.sync();
    for (int i = tid; i < H; i += kFlowThreads) {
      float s = 0.f;
      for (int k = 0; k < C; ++k) s += dyp[k * kMaxH + i];                          // fixed order
      dh[i] = s;
    }
    __syncthreads();
    // conditioner net backward, last hidden layer first (every CTA carries the full (H,) gradients)
    int hoff = hoff_last;
    for (int l = f.n_hidden - 1; l >= 0; --l) {
      const Lin& L = f.hidden[t][l];
      const float* hout = hsave + hoff;                                  // this layer's post-activation output
      const float* vin = (l == 0) ? xm : (hsave + hoff - f.hidden[t][l - 1].out);   // its input
      for (int j = tid; j < L.out; j += kFlowThreads) {
        const float g = dh[j] * act_bwd(f.kind, l == f.n_hidden - 1, hout[j]);
        da[j] = g;
        if (c == 0) L.db[goff + j] = g;
      }
      __syncthreads();
      for (int j = c + warp * C; j < L.out; j += kWarps * C) {          // weight-gradient rows j = c (mod C)
        const float a = da[j];
        float* pw = L.dW + goff + (int64_t)j * L.in;
        for (int i = lane; i < L.in; i += 32) pw[i] = a * vin[i];
      }
      // gradient wrt the layer input: dv[i] = sum_j W[j,i] da[j]; threads over i (coalesced), loop over j
      if (l == 0) {
        for (int i = d0 + tid; i < d1; i += kFlowThreads) {               // only this CTA's dims
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
    cl.sync();     // nobody writes the next transform's partials into a CTA that is still summing this one's
  }
  for (int d = d0 + tid; d < d1; d += kFlowThreads) dz_in[r * D + d] = dz[d];
}

// cluster width for a flow of dimension D: enough dims per CTA to keep its warps busy
int cluster_for(int D) {
  int c = 1;
  while (c < kMaxCluster && D / (2 * c) >= 48) c *= 2;
  return c;
}

template <typename... Args>
int launch_cluster(void (*kernel)(Args...), int rows, int C, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(rows * C));
  cfg.blockDim = dim3(kFlowThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LBBNN_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...));
  return LBBNN_OK;
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
  LBBNN_REQUIRE(rows < (1 << 20), "too many rows");
  const size_t smem = (size_t)(3 * ((d.dim + 3) & ~3) + 2 * kMaxH + kMaxCluster) * sizeof(float);
  if (smem > 48 * 1024) LBBNN_CUDA(cudaFuncSetAttribute(flow_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (int rc = launch_cluster(flow_fwd_kernel, (int)rows, cluster_for(d.dim), smem, (cudaStream_t)s, (const FlowDev)d, z_in, masks,
                              (const Noise)make_noise(mask_u), rows, z_out, logdet, save))
    return rc;
  return check_launch("flow_fwd");
}

extern "C" int lbbnn_flow_bwd(const lbbnn_flow* F, const lbbnn_flow_grads* G, int64_t rows, const float* masks,
                              const lbbnn_noise* mask_u, const float* dz_out, const float* dlogdet, const float* save,
                              float* dz_in, lbbnn_stream s) {
  FlowDev d;
  LBBNN_REQUIRE(G != nullptr, "flow grads NULL");
  if (int rc = to_dev(F, G, &d)) return rc;
  LBBNN_REQUIRE(save && dz_in && rows > 0, "NULL argument");
  LBBNN_REQUIRE(rows < (1 << 20), "too many rows");
  const size_t smem = (size_t)(5 * ((d.dim + 3) & ~3) + 2 * kMaxH + kWarps * kMaxH + kMaxCluster * kMaxH) * sizeof(float);
  if (smem > 48 * 1024) LBBNN_CUDA(cudaFuncSetAttribute(flow_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (int rc = launch_cluster(flow_bwd_kernel, (int)rows, cluster_for(d.dim), smem, (cudaStream_t)s, (const FlowDev)d, masks,
                              (const Noise)make_noise(mask_u), rows, dz_out, dlogdet, save, dz_in))
    return rc;
  return check_launch("flow_bwd");
}
