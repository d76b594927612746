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

// Every GEMV of a flow evaluation is a LATENCY chain (a few hundred weights per thread-block, one row of z): the mappings
// below put all weight loads of a phase in flight at once (fully unrolled, predicated) so a phase costs about one L2 round
// trip instead of one per loop iteration / per warp pass (r01: 65 us forward, 113 us backward for dim 784, 2 transforms).
constexpr int kQ = 4;                 // threads per output of the narrow GEMVs
constexpr int kHK = kMaxH / kQ;       // weights per thread (in <= kMaxH)
constexpr int kGroup = kFlowThreads / kQ;   // outputs per pass of a narrow GEMV (= kMaxH)
static_assert(kGroup == kMaxH, "one pass of the narrow GEMV covers the widest hidden layer");

// dot(w0[0..n), v[0..n)) by one warp; lanes take float4 columns, 8 independent loads in flight per lane
__device__ __forceinline__ float dot_row_warp(const float* __restrict__ w0, const float* __restrict__ v, int n, bool vec, int lane) {
  float a = 0.f;
  if (vec) {
    const float4* p0 = reinterpret_cast<const float4*>(w0);
    const float4* pv = reinterpret_cast<const float4*>(v);
    const int n4 = n >> 2;
    for (int base = 0; base < n4; base += 256) {
      float4 q[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = base + lane + 32 * k;
        q[k] = i < n4 ? __ldg(p0 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = base + lane + 32 * k;
        if (i < n4) {
          const float4 x = pv[i];
          a = fmaf(q[k].x, x.x, fmaf(q[k].y, x.y, fmaf(q[k].z, x.z, fmaf(q[k].w, x.w, a))));
        }
      }
    }
  } else {
    for (int base = 0; base < n; base += 256) {
      float q[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = base + lane + 32 * k;
        q[k] = i < n ? __ldg(w0 + i) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = base + lane + 32 * k;
        if (i < n) a = fmaf(q[k], v[i], a);
      }
    }
  }
  return warp_sum(a);
}

// Wide-input layer (in = D): one warp per output neuron j of a caller-chosen subset (j = first, first + step, ...):
// out[j] = act(b[j] + W[j,:] v).  `bcast` > 0: the result is written into the `out` array of EVERY CTA of the cluster
// (distributed shared memory).
__device__ __forceinline__ void gemv_rows(const Lin& L, const float* __restrict__ v, float* __restrict__ out, int kind, bool last,
                                          float* __restrict__ save, int first, int step, cg::cluster_group& cl, int bcast) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec = (L.in % 4 == 0) && aligned16(L.W) && aligned16(v);
  for (int j = first + warp * step; j < L.out; j += kWarps * step) {
    const float a0 = dot_row_warp(L.W + (int64_t)j * L.in, v, L.in, vec, lane);
    const float h = act_fwd(kind, last, a0 + __ldg(L.b + j));
    if (bcast > 0) {
      if (lane < bcast) cl.map_shared_rank(out, lane)[j] = h;
    } else if (lane == 0) {
      out[j] = h;
    }
    if (save && lane == 0) save[j] = h;
  }
}

// Narrow layer (in, out <= kMaxH), whole layer in one pass: kQ threads per neuron, thread p of a quad takes inputs p, p + 4, ...
__device__ __forceinline__ void gemv_small(const Lin& L, const float* __restrict__ v, float* __restrict__ out, int kind, bool last,
                                           float* __restrict__ save) {
  const int j = threadIdx.x >> 2, p = threadIdx.x & 3;
  float acc = 0.f;
  if (j < L.out) {
    const float* w = L.W + (int64_t)j * L.in;
    float wr[kHK];
#pragma unroll
    for (int k = 0; k < kHK; ++k) {
      const int i = p + kQ * k;
      wr[k] = i < L.in ? __ldg(w + i) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kHK; ++k) {
      const int i = p + kQ * k;
      if (i < L.in) acc = fmaf(wr[k], v[i], acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (j < L.out && p == 0) {
    const float h = act_fwd(kind, last, acc + __ldg(L.b + j));
    out[j] = h;
    if (save) save[j] = h;
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
      gemv_small(f.hidden[t][l], v, cur, f.kind, l == f.n_hidden - 1, c == 0 ? svh : nullptr);
      if (svh) svh += f.hidden[t][l].out;
      __syncthreads();
      v = cur;
      cur = (cur == ha) ? hb : ha;
    }
    // shift / scale heads and the coupling for the dims of this CTA's slice: kQ threads per dim, each with its quarter of
    // both weight rows in flight at once
    const Lin& Ls = f.shift[t];
    const Lin& Lc = f.scale[t];
    const int H = Ls.in;
    float ld = 0.f;
    const int p = tid & 3;
    for (int base = d0; base < d1; base += kGroup) {
      const int d = base + (tid >> 2);
      const bool ok = d < d1;
      float a1 = 0.f, a2 = 0.f;
      if (ok) {
        const float* ws = Ls.W + (int64_t)d * H;
        const float* wc = Lc.W + (int64_t)d * H;
        float w1[kHK], w2[kHK];
#pragma unroll
        for (int k = 0; k < kHK; ++k) {
          const int i = p + kQ * k;
          w1[k] = i < H ? __ldg(ws + i) : 0.f;
          w2[k] = i < H ? __ldg(wc + i) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kHK; ++k) {
          const int i = p + kQ * k;
          if (i < H) {
            const float vi = v[i];
            a1 = fmaf(w1[k], vi, a1);
            a2 = fmaf(w2[k], vi, a2);
          }
        }
      }
      a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
      a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
      a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
      a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
      float x = 0.f;
      if (ok) {
        const float sh = a1 + __ldg(Ls.b + d), g = 1.0f / (1.0f + expf(-(a2 + __ldg(Lc.b + d))));
        const float z = zs[d], m = ms[d];
        if (f.kind == LBBNN_FLOW_RNVP) x = (1.0f - m) * z * g + (1.0f - g) * sh + m * z;       // flows2:215
        else x = m * z + (1.0f - m) * (z * g + (1.0f - g) * sh);                               // flows2:238
        if (p == 0) {
          ld += (1.0f - m) * logf(g);
          if (sv) { sv[D + d] = g; sv[2 * D + d] = sh; }
        }
      }
      __syncwarp();      // every thread of the quad has read zs[d] before one of them overwrites it (own CTA included)
      // the new z of this dim into every CTA (the quad shares the remote stores); no CTA reads another's dims in this phase
      if (ok)
        for (int k = p; k < C; k += kQ) cl.map_shared_rank(zs, k)[d] = x;
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
  float* part = dh + kMaxH;       // [kQ][kMaxH] partial sums of the narrow transposed GEMVs
  float* dyp = part + kQ * kMaxH; // [kMaxCluster][kMaxH] per-CTA partials of the conditioner-output gradient
  float* hall = dyp + kMaxCluster * kMaxH;   // [LBBNN_FLOW_MAX_HIDDEN][kMaxH] this transform's saved hidden activations
  Noise nz = mask_noise;
  nz.resolve();
  const int64_t r = blockIdx.x / C;
  const int tid = threadIdx.x;
  const int per = (D + C - 1) / C, d0 = c * per, d1 = min(D, d0 + per);
  const float dld = dlogdet ? dlogdet[r] : 0.f;
  const int64_t goff = r * f.grad_row_stride;
  const int gi = tid & (kMaxH - 1), gp = tid >> 7;      // narrow transposed GEMVs: output index, quarter of the reduction
  static_assert(kFlowThreads == kQ * kMaxH, "thread = (output, quarter)");
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
    // saved hidden activations of this transform -> shared memory (layer l at hall + l * kMaxH)
    {
      int off = 0;
      for (int l = 0; l < f.n_hidden; ++l) {
        const int n = f.hidden[t][l].out;
        if (tid < n) hall[l * kMaxH + tid] = hsave[off + tid];
        off += n;
      }
    }
    const float* y = hall + (f.n_hidden - 1) * kMaxH;   // output of the conditioner net
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
    // head weight gradients: the outer products dsh x y, dsc x y of this CTA's dims, one flat coalesced sweep
    {
      const int n = (d1 - d0) * H;
      float* ps = Ls.dW + goff + (int64_t)d0 * H;
      float* pc = Lc.dW + goff + (int64_t)d0 * H;
      for (int idx = tid; idx < n; idx += kFlowThreads) {
        const int dl = idx / H, i = idx - dl * H;
        const float yi = y[i];
        ps[idx] = dsh[d0 + dl] * yi;
        pc[idx] = dsc[d0 + dl] * yi;
      }
    }
    // partial of dy[i] = sum_d Wt[d,i] dsh[d] + Ws[d,i] dsc[d] over this CTA's dims: thread (i, quarter of the dims),
    // consecutive threads read consecutive i (coalesced rows), 16 dims = 32 loads in flight per thread
    {
      float acc = 0.f;
      if (gi < H) {
        for (int base = d0 + gp; base < d1; base += kQ * 16) {
          float w1[16], w2[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int d = base + kQ * k;
            w1[k] = d < d1 ? __ldg(Ls.W + (int64_t)d * H + gi) : 0.f;
            w2[k] = d < d1 ? __ldg(Lc.W + (int64_t)d * H + gi) : 0.f;
          }
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int d = base + kQ * k;
            if (d < d1) acc = fmaf(w1[k], dsh[d], fmaf(w2[k], dsc[d], acc));
          }
        }
      }
      part[gp * kMaxH + gi] = acc;
    }
    __syncthreads();
    for (int i = tid; i < H; i += kFlowThreads) {
      const float s = (part[i] + part[kMaxH + i]) + (part[2 * kMaxH + i] + part[3 * kMaxH + i]);
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
    for (int l = f.n_hidden - 1; l >= 0; --l) {
      const Lin& L = f.hidden[t][l];
      const float* hout = hall + l * kMaxH;                              // this layer's post-activation output
      const float* vin = (l == 0) ? xm : (hall + (l - 1) * kMaxH);       // its input
      for (int j = tid; j < L.out; j += kFlowThreads) {
        const float g = dh[j] * act_bwd(f.kind, l == f.n_hidden - 1, hout[j]);
        da[j] = g;
        if (c == 0) L.db[goff + j] = g;
      }
      __syncthreads();
      // weight-gradient rows j = c (mod C): outer product da[j] x vin, one flat coalesced sweep over (own rows) x in
      {
        const int rows = c < L.out ? (L.out - c + C - 1) / C : 0;
        const int n = rows * L.in;
        for (int idx = tid; idx < n; idx += kFlowThreads) {
          const int jr = idx / L.in, i = idx - jr * L.in;
          const int j = c + jr * C;
          L.dW[goff + (int64_t)j * L.in + i] = da[j] * vin[i];
        }
      }
      // gradient wrt the layer input: dv[i] = sum_j W[j,i] da[j]; thread (i, quarter of j): consecutive threads read
      // consecutive i of a weight row, all of a thread's loads in flight at once
      if (l == 0) {
        for (int base = d0; base < d1; base += kMaxH) {                    // only this CTA's dims
          const int i = base + gi;
          float acc = 0.f;
          if (i < d1) {
            float wr[kHK];
#pragma unroll
            for (int k = 0; k < kHK; ++k) {
              const int j = gp + kQ * k;
              wr[k] = j < L.out ? __ldg(L.W + (int64_t)j * L.in + i) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < kHK; ++k) {
              const int j = gp + kQ * k;
              if (j < L.out) acc = fmaf(wr[k], da[j], acc);
            }
          }
          part[gp * kMaxH + gi] = acc;
          __syncthreads();
          if (tid < kMaxH && base + tid < d1) {
            const int d = base + tid;
            dz[d] += ms[d] * ((part[tid] + part[kMaxH + tid]) + (part[2 * kMaxH + tid] + part[3 * kMaxH + tid]));   // net input was m * z
          }
          __syncthreads();
        }
      } else {
        float acc = 0.f;
        if (gi < L.in) {
          float wr[kHK];
#pragma unroll
          for (int k = 0; k < kHK; ++k) {
            const int j = gp + kQ * k;
            wr[k] = j < L.out ? __ldg(L.W + (int64_t)j * L.in + gi) : 0.f;
          }
#pragma unroll
          for (int k = 0; k < kHK; ++k) {
            const int j = gp + kQ * k;
            if (j < L.out) acc = fmaf(wr[k], da[j], acc);
          }
        }
        part[gp * kMaxH + gi] = acc;
        __syncthreads();
        for (int i = tid; i < L.in; i += kFlowThreads)
          dh[i] = (part[i] + part[kMaxH + i]) + (part[2 * kMaxH + i] + part[3 * kMaxH + i]);
        __syncthreads();
      }
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
  const size_t smem = (size_t)(5 * ((d.dim + 3) & ~3) + 2 * kMaxH + kQ * kMaxH + kMaxCluster * kMaxH + LBBNN_FLOW_MAX_HIDDEN * kMaxH) * sizeof(float);
  if (smem > 48 * 1024) LBBNN_CUDA(cudaFuncSetAttribute(flow_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (int rc = launch_cluster(flow_bwd_kernel, (int)rows, cluster_for(d.dim), smem, (cudaStream_t)s, (const FlowDev)d, masks,
                              (const Noise)make_noise(mask_u), rows, dz_out, dlogdet, save, dz_in))
    return rc;
  return check_launch("flow_bwd");
}
