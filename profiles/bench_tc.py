"""Microbenchmark of the tcgen05 dual GEMM at the wide-config layer shape (B=8192, 4096x4096)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bayesian-neural-nets_b200"))
import torch
from lbbnn import _capi as K
bf = torch.bfloat16
M, N, Kd = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (8192, 4096, 4096)))
mode = sys.argv[4] if len(sys.argv) > 4 else "all"
torch.manual_seed(0)
a1, a2 = torch.randn(M, Kd, device="cuda").to(bf), torch.randn(M, Kd, device="cuda").to(bf)
b1, b2 = torch.randn(N, Kd, device="cuda").to(bf), torch.randn(N, Kd, device="cuda").to(bf)
d1, d2 = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
flops = 2 * 2.0 * M * N * Kd
st = K.current_stream()

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps

if mode in ("all", "raw"):
    ms = timeit(lambda: K.check(K.lib.lbbnn_tc_dual_gemm_raw(K.ptr(a1, bf), K.ptr(a2, bf), K.ptr(b1, bf), K.ptr(b2, bf), M, N, Kd, K.ptr(d1), K.ptr(d2), st)))
    print(f"raw   {M}x{N}x{Kd}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s")
if mode in ("all", "fwd"):
    bmu, brho = torch.zeros(N, device="cuda"), torch.full((N,), -4.5, device="cuda")
    outs = [torch.empty(M, N, dtype=bf, device="cuda") for _ in range(2)] + [torch.empty(N, M, dtype=bf, device="cuda") for _ in range(2)]
    dsf, actf = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
    a2p = a2.abs(); b2p = (b2.abs() * 1e-3).to(bf)
    noise = K.make_noise(None, 1, 2)
    ms = timeit(lambda: K.check(K.lib.lbbnn_tc_lrt_fwd(K.ptr(a1, bf), K.ptr(a2p, bf), K.ptr(b1, bf), K.ptr(b2p, bf), M, Kd, N, K.ptr(bmu), K.ptr(brho), noise, K.FLAG_SAMPLE | K.FLAG_RELU, *[K.ptr(t, bf) for t in outs], K.ptr(dsf), None, st)))
    print(f"fwd   {M}x{N}x{Kd}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s (philox eps, relu, act/act2 + transposes + dsf)")
if mode in ("all", "mm"):
    ms = timeit(lambda: (torch.matmul(a1, b1.T), torch.matmul(a2, b2.T)))
    print(f"cublas 2x matmul bf16: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s")
