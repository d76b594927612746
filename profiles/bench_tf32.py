"""3xTF32 tensor-core linear layer at the MC-predictive shapes: time per launch (CUDA events, L2 flushed) against the
fp32 CUDA-core batched GEMM, accuracy of both against fp64, and the whole MCPredictor loop for several launch widths."""
import os, sys, json, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bayesian-neural-nets_b200"))
import torch, lbbnn
from lbbnn import _capi as K

torch.manual_seed(0)
dev = torch.device("cuda")
st = K.current_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=8):
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); K.check(fn()); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.mean(ts[2:])


def split(x):
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    K.check(K.lib.lbbnn_tf32_split(K.ptr(x), x.numel(), K.ptr(hi), K.ptr(lo), st))
    return hi, lo


out = {}
B = 1000
for SB in (32,):
    for (k, o, shared) in ((784, 400, True), (400, 600, False)):
        x = torch.rand((B, k) if shared else (SB, B, k), device=dev)
        w = torch.randn(SB, o, k, device=dev) * 0.1 * (torch.rand(SB, o, k, device=dev) < 0.5)
        bias = torch.rand(SB, o, device=dev)
        xh, xl = split(x); wh, wl = split(w)
        y = torch.empty(SB, B, o, device=dev)
        if shared:
            yh, yl = torch.empty(B, SB * o, device=dev), torch.empty(B, SB * o, device=dev)
            tc = lambda: K.lib.lbbnn_tc_linear_tf32x3(K.ptr(xh), K.ptr(xl), k, 0, K.ptr(wh), K.ptr(wl), K.ptr(bias), 1, B, SB * o, k,
                                                      K.FLAG_RELU, None, K.ptr(yh), K.ptr(yl), SB * o, o, st)
            simt = lambda: K.lib.lbbnn_linear_f32_batched(K.ptr(x), 0, K.ptr(w), K.ptr(bias), SB, B, k, o, K.FLAG_RELU, K.ptr(y), st)
            ref = torch.relu(torch.einsum("bk,sok->sbo", x.double(), w.double()) + bias.double()[:, None, :])
        else:
            tc = lambda: K.lib.lbbnn_tc_linear_tf32x3(K.ptr(xh), K.ptr(xl), k, B * k, K.ptr(wh), K.ptr(wl), K.ptr(bias), SB, B, o, k,
                                                      K.FLAG_RELU, K.ptr(y), None, None, o, B * o, st)
            simt = lambda: K.lib.lbbnn_linear_f32_batched(K.ptr(x), B * k, K.ptr(w), K.ptr(bias), SB, B, k, o, K.FLAG_RELU, K.ptr(y), st)
            ref = torch.relu(torch.einsum("sbk,sok->sbo", x.double(), w.double()) + bias.double()[:, None, :])
        us_tc = timed(tc)
        got = (yh + yl).view(B, SB, o).permute(1, 0, 2) if shared else y.clone()
        err_tc = ((got.double() - ref).abs().max() / ref.abs().max()).item()
        us_simt = timed(simt)
        err_simt = ((y.double() - ref).abs().max() / ref.abs().max()).item()
        fl = 2.0 * B * k * o * SB
        out[f"SB{SB}_{k}x{o}"] = {"tc_us": round(us_tc, 1), "tc_tflops": round(fl / us_tc / 1e6, 1), "tc_err": err_tc,
                                  "simt_us": round(us_simt, 1), "simt_tflops": round(fl / us_simt / 1e6, 1), "simt_err": err_simt}
        print(f"SB{SB}_{k}x{o}", out[f"SB{SB}_{k}x{o}"], flush=True)

# classifier head: fused GEMV + accumulation kernel vs batched SIMT GEMM + accumulation kernel
for SB in (16, 32):
    k, c = 600, 10
    h = torch.rand(SB, B, k, device=dev)
    w = torch.randn(SB, c, k, device=dev) * 0.1
    bias = torch.rand(SB, c, device=dev)
    logits = torch.empty(SB, B, c, device=dev)
    sl, sp = torch.zeros(B, c, dtype=torch.float64, device=dev), torch.zeros(B, c, dtype=torch.float64, device=dev)
    sl2, sp2 = torch.zeros_like(sl), torch.zeros_like(sp)
    us_g = timed(lambda: K.lib.lbbnn_linear_f32_batched(K.ptr(h), B * k, K.ptr(w), K.ptr(bias), SB, B, k, c, 0, K.ptr(logits), st))
    us_a = timed(lambda: K.lib.lbbnn_mc_accumulate_batched(K.ptr(logits), SB, B, c, sl.data_ptr(), sp.data_ptr(), None, st))
    us_f = timed(lambda: K.lib.lbbnn_mc_head_accumulate(K.ptr(h), B * k, K.ptr(w), K.ptr(bias), SB, B, k, c, sl2.data_ptr(),
                                                        sp2.data_ptr(), None, st))
    out[f"head_SB{SB}"] = {"sgemm_us": round(us_g, 1), "accumulate_us": round(us_a, 1), "fused_us": round(us_f, 1),
                           "rel_diff": ((sl2 - sl).abs().max() / sl.abs().max()).item()}
    print(f"head_SB{SB}", out[f"head_SB{SB}"], flush=True)

net = lbbnn.mf.BayesianNetwork().to(dev)
with torch.no_grad():
    for l in net.layers:
        l.lambdal.normal_(0, 2)
x = torch.rand(B, 784, device=dev)
S = 1152
for gemm, SB, lanes, fh in (("simt", 21, 1, False), ("auto", 32, 1, False), ("auto", 32, 1, True), ("auto", 32, 2, False),
                            ("auto", 32, 2, True), ("auto", 48, 2, True), ("auto", 32, 3, True), ("auto", 64, 2, True)):
    mc = lbbnn.mf.MCPredictor(net, batch=B, seed=1, samples_per_launch=SB, gemm=gemm, lanes=lanes, fused_head=fh)
    mc.run(x, S); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        mc.run(x, S)
    b.record(); b.synchronize()
    r = S * 3 / (a.elapsed_time(b) * 1e-3)
    pred = mc.result(S)["pred"]
    key = f"mc_{gemm}_SB{SB}_L{lanes}_{'fused' if fh else 'split'}head"
    out[key] = {"samples_per_s": round(r), "n_tc": mc.n_tc}
    if gemm == "simt" and lanes == 1:
        base_pred, base_logp = pred.clone(), mc.sum_logp.clone()
    else:
        out[key]["pred_equal_simt"] = bool(torch.equal(pred, base_pred))
        out[key]["rel_err_logp_vs_simt"] = ((mc.sum_logp - base_logp).abs().max() / base_logp.abs().max()).item()
    print(key, out[key], flush=True)
    del mc
print(json.dumps(out))
