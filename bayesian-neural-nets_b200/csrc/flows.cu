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

#ifdef LBBNN_FLOW_PROF   // profiles/flow_phase_prof.cu: clock64 of CTA 0 / thread 0 at the phase boundaries
__device__ long long g_flow_prof[128];
#define FLOW_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_flow_prof[i] = clock64(); } while (0)
#else
#define FLOW_STAMP(i) do { } while (0)
#endif

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
// the masks of dims [4q, 4q + 4) of row r (same values as mask_of): one Philox call when the row starts on a quad boundary
__device__ __forceinline__ void mask_quad(const float* __restrict__ masks, const Noise& nz, int t, int64_t r, int64_t R, int q, int D,
                                          float m[4]) {
  const uint64_t e0 = (uint64_t)r * (uint64_t)D + (uint64_t)(4 * q);
  if (!masks && (e0 & 3) == 0) {
    float u[4];
    philox_uniform4(nz.seed, nz.stream + (uint64_t)t, e0 >> 2, u);
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = u[k] < 0.5f ? 1.0f : 0.0f;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = 4 * q + k < D ? mask_of(masks, nz, t, r, R, 4 * q + k, D) : 0.f;
  }
}

// Every GEMV of a flow evaluation is a LATENCY chain (a few hundred weights per thread-block, one row of z): the mappings
// below put all weight loads of a phase in flight at once (fully unrolled, predicated) so a phase costs about one L2 round
// trip instead of one per loop iteration / per warp pass (r01: 65 us forward, 113 us backward for dim 784, 2 transforms).
// compiler-level fence: every load above it is issued before the first use below (ptxas otherwise interleaves the unrolled
// loads with their uses four at a time, i.e. one L2 round trip per group instead of one per phase)
#define LOADS_ISSUED() asm volatile("" ::: "memory")
constexpr int kQ = 4;                 // threads per output of the narrow GEMVs
constexpr int kHK = kMaxH / kQ;       // weights per thread (in <= kMaxH); the kernels are instantiated for HK = 20 (hidden <= 80) and kHK
constexpr int kGroup = kFlowThreads / kQ;   // outputs per pass of a narrow GEMV (= kMaxH)
static_assert(kGroup == kMaxH, "one pass of the narrow GEMV covers the widest hidden layer");

// dot(w0[0..n), v[0..n)) by one warp; lanes take float4 columns, 8 independent loads in flight per lane.  Branch-free
// (out-of-range lanes re-read the last element and multiply by zero): a branch per element would fence the loads into
// separate basic blocks and ptxas then issues them one L2 round trip at a time.
__device__ __forceinline__ float dot_row_warp(const float* __restrict__ w0, const float* __restrict__ v, int n, bool vec, int lane) {
  float a = 0.f;
  if (vec) {
    const float4* p0 = reinterpret_cast<const float4*>(w0);
    const float4* pv = reinterpret_cast<const float4*>(v);
    const int n4 = n >> 2;
    for (int base = 0; base < n4; base += 256) {
      float4 q[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) q[k] = __ldg(p0 + min(base + lane + 32 * k, n4 - 1));
      LOADS_ISSUED();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = base + lane + 32 * k;
        const float4 x = pv[min(i, n4 - 1)];
        const float s = fmaf(q[k].x, x.x, fmaf(q[k].y, x.y, fmaf(q[k].z, x.z, q[k].w * x.w)));
        a += i < n4 ? s : 0.f;
      }
    }
  } else {
    for (int base = 0; base < n; base += 256) {
      float q[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) q[k] = __ldg(w0 + min(base + lane + 32 * k, n - 1));
      LOADS_ISSUED();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = base + lane + 32 * k;
        a = fmaf(i < n ? q[k] : 0.f, v[min(i, n - 1)], a);
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

// Narrow GEMV rows (in <= 32 * HL), coalesced: each warp takes RW consecutive rows, lanes run along the contiguous weight
// row (row-per-thread-quad reads cost 8 L1 tag lookups per load instruction and made a 75 x 75 layer take 1.5 us); all
// RW * HL loads of a lane are in flight at once, then RW interleaved shuffle reductions.  sums[r] = W[row_r, :] . v on EVERY
// lane.  Rows past `nrows` redo the last row (branch-free, see dot_row_warp).
template <int RW, int HL>
__device__ __forceinline__ void warp_rows_dot(const float* __restrict__ W, int in, int row_first, int nrows,
                                              const float* __restrict__ v, float (&sums)[RW]) {
  const int lane = threadIdx.x & 31;
  float w[RW][HL];
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const float* wr = W + (int64_t)min(row_first + r, nrows - 1) * in;
#pragma unroll
    for (int k = 0; k < HL; ++k) w[r][k] = __ldg(wr + min(lane + 32 * k, in - 1));
  }
  LOADS_ISSUED();
  float x[HL];
#pragma unroll
  for (int k = 0; k < HL; ++k) x[k] = lane + 32 * k < in ? v[lane + 32 * k] : 0.f;
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < HL; ++k) a = fmaf(w[r][k], x[k], a);
    sums[r] = a;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < RW; ++r) sums[r] += __shfl_xor_sync(0xffffffffu, sums[r], o);
  }
}

// the same for two weight matrices sharing v (the shift / scale heads): both sets of loads in flight together
template <int RW, int HL>
__device__ __forceinline__ void warp_rows_dot2(const float* __restrict__ W1, const float* __restrict__ W2, int in, int row_first,
                                               int nrows, const float* __restrict__ v, float (&s1)[RW], float (&s2)[RW]) {
  const int lane = threadIdx.x & 31;
  float w1[RW][HL], w2[RW][HL];
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const int64_t ro = (int64_t)min(row_first + r, nrows - 1) * in;
#pragma unroll
    for (int k = 0; k < HL; ++k) {
      const int i = min(lane + 32 * k, in - 1);
      w1[r][k] = __ldg(W1 + ro + i);
      w2[r][k] = __ldg(W2 + ro + i);
    }
  }
  LOADS_ISSUED();
  float x[HL];
#pragma unroll
  for (int k = 0; k < HL; ++k) x[k] = lane + 32 * k < in ? v[lane + 32 * k] : 0.f;
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < HL; ++k) { a = fmaf(w1[r][k], x[k], a); b = fmaf(w2[r][k], x[k], b); }
    s1[r] = a;
    s2[r] = b;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      s1[r] += __shfl_xor_sync(0xffffffffu, s1[r], o);
      s2[r] += __shfl_xor_sync(0xffffffffu, s2[r], o);
    }
  }
}

// lane-indexed pick from a register array without dynamic indexing
template <int RW>
__device__ __forceinline__ float pick(const float (&a)[RW], int r) {
  float m = a[0];
#pragma unroll
  for (int k = 1; k < RW; ++k) m = (r == k) ? a[k] : m;
  return m;
}

// Narrow layer (in, out <= kMaxH), whole layer in one pass of the CTA's 16 warps: out[j] = act(b[j] + W[j,:] v)
template <int HK>
__device__ __forceinline__ void gemv_small(const Lin& L, const float* __restrict__ v, float* __restrict__ out, int kind, bool last,
                                           float* __restrict__ save) {
  constexpr int RW = (4 * HK + kWarps - 1) / kWarps, HL = (4 * HK + 31) / 32;     // rows per warp, loads per lane and row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float sums[RW];
  warp_rows_dot<RW, HL>(L.W, L.in, warp * RW, L.out, v, sums);
  const int j = warp * RW + lane;
  if (lane < RW && j < L.out) {
    const float h = act_fwd(kind, last, pick<RW>(sums, lane) + __ldg(L.b + j));
    out[j] = h;
    if (save) save[j] = h;
  }
}

// ---- forward -------------------------------------------------------------------------------------------
// One CLUSTER of C CTAs per row of z.  Every CTA keeps the whole z in shared memory; the two big GEMVs of a transform
// are split across the cluster -- the first hidden layer by output neuron (results broadcast into every CTA's shared
// memory through DSMEM), the shift / scale heads and the coupling by output dimension (each CTA broadcasts its slice
// of the new z) -- with one cluster barrier after each.  The small hidden layers are recomputed by every CTA.
template <int HK>
__global__ void __launch_bounds__(kFlowThreads, 1) flow_fwd_kernel(const FlowDev f, const float* __restrict__ z_in,
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
  FLOW_STAMP(0);
  cl.sync();                   // every CTA of the cluster is running before any remote shared-memory access
  FLOW_STAMP(1);
  float ld_total = 0.f;
  for (int t = 0; t < f.n_transforms; ++t) {
    float* sv = save ? save + ((int64_t)r * f.n_transforms + t) * f.save_stride : nullptr;
    for (int q = tid; 4 * q < D; q += kFlowThreads) {
      float m4[4];
      mask_quad(masks, nz, t, r, R, q, D, m4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int d = 4 * q + k;
        if (d < D) {
          ms[d] = m4[k];
          xm[d] = m4[k] * zs[d];
          if (sv && d >= d0 && d < d1) sv[d] = zs[d];
        }
      }
    }
    __syncthreads();
    FLOW_STAMP(2 + 8 * t);
    float* svh = sv ? sv + 3 * D : nullptr;
    // first hidden layer: neurons c, c + C, ... of this CTA, broadcast to the whole cluster
    gemv_rows(f.hidden[t][0], xm, ha, f.kind, f.n_hidden == 1, svh, c, C, cl, C);
    FLOW_STAMP(3 + 8 * t);
    cl.sync();
    FLOW_STAMP(4 + 8 * t);
    const float* v = ha;
    float* cur = hb;
    if (svh) svh += f.hidden[t][0].out;
    for (int l = 1; l < f.n_hidden; ++l) {
      gemv_small<HK>(f.hidden[t][l], v, cur, f.kind, l == f.n_hidden - 1, c == 0 ? svh : nullptr);
      if (svh) svh += f.hidden[t][l].out;
      __syncthreads();
      v = cur;
      cur = (cur == ha) ? hb : ha;
    }
    FLOW_STAMP(5 + 8 * t);
    // shift / scale heads and the coupling for the dims of this CTA's slice: each warp takes 8 consecutive dims per pass
    // (coalesced weight rows, all loads of both heads in flight), lane (r, q) then owns dim r and the remote stores to
    // CTAs q, q + 4
    const Lin& Ls = f.shift[t];
    const Lin& Lc = f.scale[t];
    const int H = Ls.in;
    float ld = 0.f;
    {
      constexpr int RW = 8, HL = (4 * HK + 31) / 32;
      const int r_ = lane & 7, q_ = lane >> 3;
      for (int base = d0; base < d1; base += kWarps * RW) {
        float s1[RW], s2[RW];
        const int first = base + warp * RW;
        warp_rows_dot2<RW, HL>(Ls.W, Lc.W, H, first, d1, v, s1, s2);
        const int d = first + r_;
        const bool ok = d < d1;
        float x = 0.f;
        if (ok) {
          const float sh = pick<RW>(s1, r_) + __ldg(Ls.b + d), g = 1.0f / (1.0f + expf(-(pick<RW>(s2, r_) + __ldg(Lc.b + d))));
          const float z = zs[d], m = ms[d];
          if (f.kind == LBBNN_FLOW_RNVP) x = (1.0f - m) * z * g + (1.0f - g) * sh + m * z;       // flows2:215
          else x = m * z + (1.0f - m) * (z * g + (1.0f - g) * sh);                               // flows2:238
          if (q_ == 0) {
            ld += (1.0f - m) * logf(g);
            if (sv) { sv[D + d] = g; sv[2 * D + d] = sh; }
          }
        }
        __syncwarp();      // every lane has read zs[d] before one of them overwrites it (own CTA included)
        // the new z of this dim into every CTA; no CTA reads another's dims in this phase
        if (ok)
          for (int k = q_; k < C; k += 4) cl.map_shared_rank(zs, k)[d] = x;
      }
    }
    FLOW_STAMP(6 + 8 * t);
    const float tot = block_sum(ld, red);
    if (tid == 0) cl.map_shared_rank(ldp, 0)[c] = tot;
    FLOW_STAMP(7 + 8 * t);
    cl.sync();
    FLOW_STAMP(8 + 8 * t);
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
template <int HK>
__global__ void __launch_bounds__(kFlowThreads, 1) flow_bwd_kernel(const FlowDev f, const float* __restrict__ masks,
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
  FLOW_STAMP(64);
  cl.sync();
  FLOW_STAMP(65);
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
    for (int q = tid; 4 * q < D; q += kFlowThreads) {
      float m4[4];
      mask_quad(masks, nz, t, r, R, q, D, m4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int d = 4 * q + k;
        if (d >= D) continue;
        const float m = m4[k];
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
    }
    __syncthreads();
    FLOW_STAMP(66 + 16 * t);
    // head weight gradients: the outer products dsh x y, dsc x y of this CTA's dims, one flat coalesced sweep
    if (gi < H) {
      const float yi = y[gi];
      float* ps = Ls.dW + goff + gi;
      float* pc = Lc.dW + goff + gi;
#pragma unroll 4
      for (int d = d0 + gp; d < d1; d += kQ) {
        ps[(int64_t)d * H] = dsh[d] * yi;
        pc[(int64_t)d * H] = dsc[d] * yi;
      }
    }
    FLOW_STAMP(67 + 16 * t);
    // partial of dy[i] = sum_d Wt[d,i] dsh[d] + Ws[d,i] dsc[d] over this CTA's dims: thread (i, quarter of the dims),
    // consecutive threads read consecutive i (coalesced rows), 16 dims = 32 loads in flight per thread
    {
      float acc = 0.f;
      const int ic = min(gi, H - 1);
      for (int base = d0 + gp; base < d1; base += kQ * 16) {
        float w1[16], w2[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int d = min(base + kQ * k, d1 - 1);
          w1[k] = __ldg(Ls.W + (int64_t)d * H + ic);
          w2[k] = __ldg(Lc.W + (int64_t)d * H + ic);
        }
        LOADS_ISSUED();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int d = base + kQ * k, dc = min(d, d1 - 1);
          const float t_ = fmaf(w1[k], dsh[dc], w2[k] * dsc[dc]);
          acc += d < d1 ? t_ : 0.f;
        }
      }
      part[gp * kMaxH + gi] = acc;
    }
    __syncthreads();
    for (int i = tid; i < H; i += kFlowThreads) {
      const float s = (part[i] + part[kMaxH + i]) + (part[2 * kMaxH + i] + part[3 * kMaxH + i]);
      for (int k = 0; k < C; ++k) cl.map_shared_rank(dyp, k)[c * kMaxH + i] = s;   // this CTA's partial, into every CTA
    }
    FLOW_STAMP(68 + 16 * t);
    cl.sync();
    FLOW_STAMP(69 + 16 * t);
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
      FLOW_STAMP(70 + 16 * t + 3 * (f.n_hidden - 1 - l));
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
      FLOW_STAMP(71 + 16 * t + 3 * (f.n_hidden - 1 - l));
      // gradient wrt the layer input: dv[i] = sum_j W[j,i] da[j]; thread (i, quarter of j): consecutive threads read
      // consecutive i of a weight row, all of a thread's loads in flight at once
      if (l == 0) {
        for (int base = d0; base < d1; base += kMaxH) {                    // only this CTA's dims
          const int i = min(base + gi, d1 - 1);
          float acc = 0.f;
          {
            float wr[HK];
#pragma unroll
            for (int k = 0; k < HK; ++k) wr[k] = __ldg(L.W + (int64_t)min(gp + kQ * k, L.out - 1) * L.in + i);
            LOADS_ISSUED();
#pragma unroll
            for (int k = 0; k < HK; ++k) {
              const int j = gp + kQ * k;
              acc = fmaf(j < L.out ? wr[k] : 0.f, da[min(j, L.out - 1)], acc);
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
        {
          const int ic = min(gi, L.in - 1);
          float wr[HK];
#pragma unroll
          for (int k = 0; k < HK; ++k) wr[k] = __ldg(L.W + (int64_t)min(gp + kQ * k, L.out - 1) * L.in + ic);
          LOADS_ISSUED();
#pragma unroll
          for (int k = 0; k < HK; ++k) {
            const int j = gp + kQ * k;
            acc = fmaf(j < L.out ? wr[k] : 0.f, da[min(j, L.out - 1)], acc);
          }
        }
        part[gp * kMaxH + gi] = acc;
        __syncthreads();
        for (int i = tid; i < L.in; i += kFlowThreads)
          dh[i] = (part[i] + part[kMaxH + i]) + (part[2 * kMaxH + i] + part[3 * kMaxH + i]);
        __syncthreads();
      }
    }
    FLOW_STAMP(82 + 16 * t);
    cl.sync();     // nobody writes the next transform's partials into a CTA that is still summing this one's
  }
  for (int d = d0 + tid; d < d1; d += kFlowThreads) dz_in[r * D + d] = dz[d];
}

// widest hidden layer of the conditioner nets (selects the HK instantiation)
int max_hidden(const FlowDev& d) {
  int h = 0;
  for (int t = 0; t < d.n_transforms; ++t)
    for (int l = 0; l < d.n_hidden; ++l) h = d.hidden[t][l].out > h ? d.hidden[t][l].out : h;
  return h;
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
  const bool narrow = max_hidden(d) <= 4 * 20;
  auto kernel = narrow ? flow_fwd_kernel<20> : flow_fwd_kernel<kMaxH / kQ>;
  if (smem > 48 * 1024) LBBNN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (int rc = launch_cluster(kernel, (int)rows, cluster_for(d.dim), smem, (cudaStream_t)s, (const FlowDev)d, z_in, masks,
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
  const bool narrow = max_hidden(d) <= 4 * 20;
  auto kernel = narrow ? flow_bwd_kernel<20> : flow_bwd_kernel<kMaxH / kQ>;
  if (smem > 48 * 1024) LBBNN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (int rc = launch_cluster(kernel, (int)rows, cluster_for(d.dim), smem, (cudaStream_t)s, (const FlowDev)d, masks,
                              (const Noise)make_noise(mask_u), rows, dz_out, dlogdet, save, dz_in))
    return rc;
  return check_launch("flow_bwd");
}
