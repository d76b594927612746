"""A/B timing of the tensor-core calls of one 4096 x 4096 layer at batch 8192 on ONE box: the r01 forms (transposed K-major
operands, transposed outputs, separate update pass) against the r02 forms (operands in place, fused update, bias partial sums).
CUDA events, 10 repetitions after 3 warm-up calls, operands > L2.   python profiles/bench_wide_calls.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bayesian-neural-nets_b200"))
import torch  # noqa: E402
import lbbnn  # noqa: E402
from lbbnn import _capi as K  # noqa: E402

bf = torch.bfloat16
B, I, O = 8192, 4096, 4096
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
torch.manual_seed(0)
st = K.current_stream()
P = K.ptr
flops = 2 * 2.0 * B * I * O


def timeit(name, fn, extra=""):
    for _ in range(3):
        K.check(fn())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        K.check(fn())
    e1.record()
    e1.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{name:62s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s {extra}", flush=True)
    return us


x = torch.rand(B, I, device="cuda")
xb, x2b, xT, x2T = K.bf16_pack(x, None, K.PACK_SQUARE)
net = lbbnn.BayesianNetwork((I, O)).cuda()
l = net.layers[0]
desc = K.make_layer(l.weight_mu.data, l.weight_rho.data, l.lambdal.data, l.bias_mu.data, l.bias_rho.data)
Mb, Vb, MT, VT = (torch.empty(O, I, dtype=bf, device="cuda"), torch.empty(O, I, dtype=bf, device="cuda"),
                  torch.empty(I, O, dtype=bf, device="cuda"), torch.empty(I, O, dtype=bf, device="cuda"))
ws = torch.empty(max(1 << 20, int(K.lib.lbbnn_lrt_bf16_prologue_workspace_bytes(I, O))), dtype=torch.uint8, device="cuda")
kl = torch.zeros(1, device="cuda")
K.check(K.lib.lbbnn_lrt_bf16_prologue(desc, l.cfg.priors, l.cfg.var_mode, P(Mb, bf), P(Vb, bf), P(MT, bf), P(VT, bf), None, None,
                                      P(kl), ws.data_ptr(), ws.numel(), st))
act, act2 = torch.empty(B, O, dtype=bf, device="cuda"), torch.empty(B, O, dtype=bf, device="cuda")
actT, act2T = torch.empty(O, B, dtype=bf, device="cuda"), torch.empty(O, B, dtype=bf, device="cuda")
dsf = torch.empty(B, O, device="cuda")
noise = K.make_noise(None, 1, 2)
fl = K.FLAG_SAMPLE | K.FLAG_RELU
print(f"layer {I} -> {O}, batch {B}; {reps} repetitions per call")
timeit("fwd   r01: act, act^2 + transposes, dsf", lambda: K.lib.lbbnn_tc_lrt_fwd(
    P(xb, bf), P(x2b, bf), P(Mb, bf), P(Vb, bf), B, I, O, P(l.bias_mu.data), P(l.bias_rho.data), noise, fl, P(act, bf), P(act2, bf),
    P(actT, bf), P(act2T, bf), P(dsf), None, st))
timeit("fwd   r02: act, act^2, dsf (no transposed outputs)", lambda: K.lib.lbbnn_tc_lrt_fwd(
    P(xb, bf), P(x2b, bf), P(Mb, bf), P(Vb, bf), B, I, O, P(l.bias_mu.data), P(l.bias_rho.data), noise, fl, P(act, bf), P(act2, bf),
    None, None, P(dsf), None, st))
for dist in (os.environ.get("LBBNN_AB_PREFETCH", "").split() or []):
    os.environ["LBBNN_TC_L2_PREFETCH"] = dist
    timeit(f"fwd   r02 with TMA L2 prefetch {dist} K blocks ahead", lambda: K.lib.lbbnn_tc_lrt_fwd(
        P(xb, bf), P(x2b, bf), P(Mb, bf), P(Vb, bf), B, I, O, P(l.bias_mu.data), P(l.bias_rho.data), noise, fl, P(act, bf), P(act2, bf),
        None, None, P(dsf), None, st))
os.environ.pop("LBBNN_TC_L2_PREFETCH", None)

de = (torch.randn(B, O, device="cuda") * 0.1).to(bf)
ds = (torch.randn(B, O, device="cuda") * 0.01).to(bf)
deT, dsT = de.T.contiguous(), ds.T.contiguous()
dM, dV = torch.empty(O, I, device="cuda"), torch.empty(O, I, device="cuda")
t_raw = timeit("dW    r01: raw, K-major transposed operands", lambda: K.lib.lbbnn_tc_dual_gemm_raw(
    P(deT, bf), P(dsT, bf), P(xT, bf), P(x2T, bf), O, I, B, P(dM), P(dV), st))
timeit("dW    r02: raw_ex, operands in place (MN-major A and B)", lambda: K.lib.lbbnn_tc_dual_gemm_raw_ex(
    P(de, bf), P(ds, bf), P(xb, bf), P(x2b, bf), O, I, B, 1, 1, P(dM), P(dV), st))
for dist in (os.environ.get("LBBNN_AB_PREFETCH", "").split() or []):
    os.environ["LBBNN_TC_L2_PREFETCH"] = dist
    timeit(f"dW    r02 raw_ex with TMA L2 prefetch {dist} K blocks ahead", lambda: K.lib.lbbnn_tc_dual_gemm_raw_ex(
        P(de, bf), P(ds, bf), P(xb, bf), P(x2b, bf), O, I, B, 1, 1, P(dM), P(dV), st))
os.environ.pop("LBBNN_TC_L2_PREFETCH", None)
names = ("weight_mu", "weight_rho", "lambdal", "bias_mu", "bias_rho")
m = {k: torch.zeros_like(getattr(l, k).data) for k in names}
v = {k: torch.zeros_like(getattr(l, k).data) for k in names}
coef = torch.zeros(2, device="cuda")
step_dev = torch.ones(1, dtype=torch.int64, device="cuda")
K.check(K.lib.lbbnn_adam_prepare(P(step_dev, torch.int64), 1e-6, 0.9, 0.999, P(coef), st))
ast = K.AdamLayerState()
for j, k in enumerate(names):
    ast.exp_avg[j], ast.exp_avg_sq[j] = m[k].data_ptr(), v[k].data_ptr()
ast.coef, ast.beta1, ast.beta2, ast.eps = coef.data_ptr(), 0.9, 0.999, 1e-8
colsum = torch.zeros(2 * O, device="cuda")
t_fin = timeit("      r01: finalize_adam (chain rule + KL + Adam pass)", lambda: K.lib.lbbnn_lrt_f32_finalize_adam(
    desc, P(dM), P(dV), P(colsum), l.cfg.priors, l.cfg.var_mode, K.FLAG_SAMPLE, None, 1.0 / 600, ast, st),
    extra=f"(1.34 GB -> {1.34e9 / 1e3:.0f} MB)")
t_fused = timeit("dW    r02: dw_adam (GEMM + update in the epilogue)", lambda: K.lib.lbbnn_tc_lrt_dw_adam(
    P(de, bf), P(ds, bf), P(xb, bf), P(x2b, bf), desc, B, l.cfg.priors, l.cfg.var_mode, 1.0 / 600, ast, st))
print(f"{'':62s} r01 raw + finalize_adam = {t_raw + t_fin:.1f} us  vs fused {t_fused:.1f} us")

xin = act                                   # a relu output: the mask source of the dX epilogue
dsf_prev = torch.randn(B, I, device="cuda")
o = [torch.empty(B, I, dtype=bf, device="cuda") for _ in range(2)] + [torch.empty(I, B, dtype=bf, device="cuda") for _ in range(2)]
fx = K.FLAG_SAMPLE | K.FLAG_MASK_DX
timeit("dX    r01: M^T, V^T operands, dE/dS + transposes", lambda: K.lib.lbbnn_tc_lrt_bwd_input(
    P(de, bf), P(ds, bf), P(MT, bf), P(VT, bf), B, I, O, P(xin, bf), P(dsf_prev), fx, P(o[0], bf), P(o[1], bf), P(o[2], bf), P(o[3], bf), st))
timeit("dX    r01 kernel without transposed outputs", lambda: K.lib.lbbnn_tc_lrt_bwd_input(
    P(de, bf), P(ds, bf), P(MT, bf), P(VT, bf), B, I, O, P(xin, bf), P(dsf_prev), fx, P(o[0], bf), P(o[1], bf), None, None, st))
part = torch.empty(int(K.lib.lbbnn_tc_colsum_part_floats(B, I)), device="cuda")
timeit("dX    r02: M, V in place (MN-major B), no bias sums", lambda: K.lib.lbbnn_tc_lrt_bwd_input_mn(
    P(de, bf), P(ds, bf), P(Mb, bf), P(Vb, bf), B, I, O, P(xin, bf), P(dsf_prev), fx, P(o[0], bf), P(o[1], bf), None, st))
timeit("dX    r02: M, V in place + bias partial sums in the epilogue", lambda: K.lib.lbbnn_tc_lrt_bwd_input_mn(
    P(de, bf), P(ds, bf), P(Mb, bf), P(Vb, bf), B, I, O, P(xin, bf), P(dsf_prev), fx, P(o[0], bf), P(o[1], bf), P(part), st))
a1, a2 = torch.randn(B, I, device="cuda").to(bf), torch.randn(B, I, device="cuda").to(bf)
b1, b2 = torch.randn(O, I, device="cuda").to(bf), torch.randn(O, I, device="cuda").to(bf)
timeit("cuBLAS: 2 x torch.matmul bf16 (same FLOPs, no epilogue)", lambda: (torch.matmul(a1, b1.T), torch.matmul(a2, b2.T)) and 0)
