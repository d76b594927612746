// Phase profile of the fused flow kernels (clock64 of CTA 0 at the phase boundaries) + their event-timed duration.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DLBBNN_FLOW_PROF \
//        profiles/flow_phase_prof.cu -o gpurun_out/flow_phase_prof && gpurun_out/flow_phase_prof
#include "../bayesian-neural-nets_b200/csrc/util.cu"
#include "../bayesian-neural-nets_b200/csrc/flows.cu"
#include <vector>
#include <cstdlib>

int main(int argc, char** argv) {
  const int D = argc > 1 ? atoi(argv[1]) : 784, T = 2, NH = 4, H = 75, R = argc > 2 ? atoi(argv[2]) : 2;
  std::vector<float*> bufs;
  auto dev = [&](size_t n, float scale) {
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = scale * ((float)rand() / RAND_MAX - 0.5f);
    float* d;
    cudaMalloc(&d, n * sizeof(float));
    cudaMemcpy(d, h.data(), n * sizeof(float), cudaMemcpyHostToDevice);
    bufs.push_back(d);
    return d;
  };
  lbbnn_flow F;
  lbbnn_flow_grads G;
  memset(&F, 0, sizeof(F));
  memset(&G, 0, sizeof(G));
  F.kind = LBBNN_FLOW_RNVP; F.dim = D; F.n_transforms = T; F.n_hidden = NH;
  size_t P = 0;
  for (int t = 0; t < T; ++t) {
    int prev = D;
    for (int l = 0; l < NH; ++l) { P += (size_t)prev * H + H; prev = H; }
    P += 2 * ((size_t)H * D + D);
  }
  float* gbuf;
  cudaMalloc(&gbuf, (size_t)R * P * sizeof(float));
  size_t off = 0;
  for (int t = 0; t < T; ++t) {
    int prev = D;
    for (int l = 0; l < NH; ++l) {
      F.t[t].hidden[l] = {dev((size_t)prev * H, 0.2f), dev(H, 0.1f), prev, H};
      G.t[t].hidden[l] = {gbuf + off, gbuf + off + (size_t)prev * H};
      off += (size_t)prev * H + H;
      prev = H;
    }
    F.t[t].shift = {dev((size_t)H * D, 0.2f), dev(D, 0.1f), H, D};
    G.t[t].shift = {gbuf + off, gbuf + off + (size_t)H * D};
    off += (size_t)H * D + D;
    F.t[t].scale = {dev((size_t)H * D, 0.2f), dev(D, 0.1f), H, D};
    G.t[t].scale = {gbuf + off, gbuf + off + (size_t)H * D};
    off += (size_t)H * D + D;
  }
  G.row_stride = (int64_t)P;
  float* z = dev((size_t)R * D, 1.0f);
  float *zo, *ld, *save, *dzi;
  cudaMalloc(&zo, (size_t)R * D * 4); cudaMalloc(&ld, R * 4); cudaMalloc(&dzi, (size_t)R * D * 4);
  cudaMalloc(&save, lbbnn_flow_save_floats(&F, R) * 4);
  float* dzo = dev((size_t)R * D, 1.0f);
  float* dld = dev(R, 1.0f);
  lbbnn_noise nz = {nullptr, 1234, 77, nullptr, 0};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int pass = 0; pass < 2; ++pass) {
    const int reps = 50;
    for (int i = 0; i < 5; ++i) {
      if (lbbnn_flow_fwd(&F, z, R, nullptr, &nz, zo, ld, save, 0)) { printf("fwd: %s\n", lbbnn_last_error()); return 1; }
      if (lbbnn_flow_bwd(&F, &G, R, nullptr, &nz, dzo, dld, save, dzi, 0)) { printf("bwd: %s\n", lbbnn_last_error()); return 1; }
    }
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) {
      if (pass == 0) lbbnn_flow_fwd(&F, z, R, nullptr, &nz, zo, ld, save, 0);
      else lbbnn_flow_bwd(&F, &G, R, nullptr, &nz, dzo, dld, save, dzi, 0);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%s: %.2f us per launch (back to back, D=%d rows=%d)\n", pass == 0 ? "flow_fwd" : "flow_bwd", ms * 1e3 / reps, D, R);
  }
  {   // do independent evaluations on different streams overlap?  (3 streams x 10 launches each vs 30 on one stream)
    cudaStream_t st[3];
    for (int i = 0; i < 3; ++i) cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
    float *zo2[3], *ld2[3], *save2[3];
    for (int i = 0; i < 3; ++i) { cudaMalloc(&zo2[i], (size_t)R * D * 4); cudaMalloc(&ld2[i], R * 4); cudaMalloc(&save2[i], lbbnn_flow_save_floats(&F, R) * 4); }
    for (int mode = 0; mode < 2; ++mode) {
      cudaDeviceSynchronize();
      cudaEventRecord(e0, st[0]);
      cudaStreamWaitEvent(st[1], e0); cudaStreamWaitEvent(st[2], e0);
      for (int i = 0; i < 30; ++i) {
        const int k = mode == 0 ? 0 : i % 3;
        lbbnn_flow_fwd(&F, z, R, nullptr, &nz, zo2[k], ld2[k], save2[k], st[k]);
      }
      cudaEvent_t j1, j2;
      cudaEventCreate(&j1); cudaEventCreate(&j2);
      cudaEventRecord(j1, st[1]); cudaEventRecord(j2, st[2]);
      cudaStreamWaitEvent(st[0], j1); cudaStreamWaitEvent(st[0], j2);
      cudaEventRecord(e1, st[0]);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("30 flow_fwd launches on %s: %.1f us total\n", mode == 0 ? "one stream" : "three streams", ms * 1e3);
    }
  }
  {   // the same three-stream pattern captured into ONE graph (what a captured training step replays)
    cudaStream_t st[3];
    for (int i = 0; i < 3; ++i) cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
    float *zo2[3], *ld2[3], *save2[3];
    for (int i = 0; i < 3; ++i) { cudaMalloc(&zo2[i], (size_t)R * D * 4); cudaMalloc(&ld2[i], R * 4); cudaMalloc(&save2[i], lbbnn_flow_save_floats(&F, R) * 4); }
    for (int mode = 0; mode < 2; ++mode) {
      cudaGraph_t g;
      cudaGraphExec_t ge;
      cudaEvent_t f0, j1, j2;
      cudaEventCreate(&f0); cudaEventCreate(&j1); cudaEventCreate(&j2);
      cudaStreamBeginCapture(st[0], cudaStreamCaptureModeThreadLocal);
      cudaEventRecord(f0, st[0]);
      cudaStreamWaitEvent(st[1], f0); cudaStreamWaitEvent(st[2], f0);
      for (int i = 0; i < 30; ++i) {
        const int k = mode == 0 ? 0 : i % 3;
        lbbnn_flow_fwd(&F, z, R, nullptr, &nz, zo2[k], ld2[k], save2[k], st[k]);
      }
      cudaEventRecord(j1, st[1]); cudaEventRecord(j2, st[2]);
      cudaStreamWaitEvent(st[0], j1); cudaStreamWaitEvent(st[0], j2);
      cudaStreamEndCapture(st[0], &g);
      cudaGraphInstantiate(&ge, g, 0);
      cudaGraphLaunch(ge, st[0]);
      cudaStreamSynchronize(st[0]);
      cudaEventRecord(e0, st[0]);
      cudaGraphLaunch(ge, st[0]);
      cudaEventRecord(e1, st[0]);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("graph of 30 flow_fwd launches, %s: %.1f us per replay\n", mode == 0 ? "one branch" : "three branches", ms * 1e3);
    }
  }
  long long prof[128];
  cudaMemcpyFromSymbol(prof, lbbnn::g_flow_prof, sizeof(prof));
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double us = 1e3 / clk_khz;
  printf("forward (CTA 0), clocks and us since the previous stamp:\n");
  const char* fn[] = {"init load", "cl.sync", "masks", "first layer", "cl.sync", "hidden 1..3", "heads", "logdet sum", "cl.sync"};
  long long prev = prof[0];
  for (int i = 1; i <= 8 + 8 * (T - 1); ++i) {
    const int k = i <= 8 ? i : i - 8;
    printf("  [%2d] %-14s %7lld clk %6.2f us\n", i, fn[k <= 1 ? k : k], prof[i] - prev, (prof[i] - prev) * us);
    prev = prof[i];
  }
  printf("  total %.2f us\n", (prof[8 + 8 * (T - 1)] - prof[0]) * us);
  printf("backward (CTA 0):\n");
  prev = prof[64];
  printf("  cl.sync %lld clk\n", prof[65] - prev);
  prev = prof[65];
  for (int t = T - 1; t >= 0; --t) {
    const int b = 66 + 16 * t;
    const char* bn[] = {"masks+coupling", "head dW sweep", "dy partial", "cl.sync", "da l3", "dW l3", "dv l3(+da l2)", "dW l2", "dv l2", "..", "..", "..", "..", "..", "..", "..", "end"};
    for (int i = 0; i <= 16; ++i) {
      if (prof[b + i] == 0) continue;
      printf("  t%d [%2d] %-16s %7lld clk %6.2f us\n", t, i, bn[i], prof[b + i] - prev, (prof[b + i] - prev) * us);
      prev = prof[b + i];
    }
  }
  printf("  total %.2f us\n", (prev - prof[64]) * us);
  return 0;
}
