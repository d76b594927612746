"""Posterior-predictive model averaging for the LRT and MNF networks, batched over Monte-Carlo samples.

Reference: the per-batch bodies of `test_ensemble` (LBBNN-GP-MF-LRT.py:239-265, LBBNN-GP-MF-MNF.py:287-318) and `outofsample`
(LBBNN-GP-MF-LRT.py:305-341): TEST_SAMPLES times `net(data, sample=True)`, a device -> host NumPy round trip per sample for
the row-normalised expit average, `outputs[0:10].mean(0)` for the ensemble prediction, `net(data, sample=False)` for the
posterior-mean prediction.

Here the samples of a launch are stacked along the batch dimension and go through the fused LRT forward kernels
(csrc/lrt_f32.cu, fp32: the 1e-5 / bit-exact-argmax parity mode) ONCE per layer:

  * the LRT weights' moments M = alpha mu, V = alpha^2 sigma^2 do not depend on the sample, so layers 2.. are ONE dual GEMM
    over (samples x batch) rows, with per-sample native noise streams (`lbbnn_lrt_f32_fwd_ex`);
  * layer 1 sees the same input for every sample: e_b = x M^T + b_mu and var_b = x^2 V^T + sigma_b^2 are computed once per
    test batch (LRT:247 "only eps changes") and every sample only adds its sqrt(var_b) eps (`lbbnn_lrt_sample_expand`);
  * MNF: every sample draws its own multiplicative z per layer (the flows run once per launch on one row per sample); z
    scales the input rows of the mean product only (MNF:197-198), which the kernel applies per row group;
  * log-softmax, the row-normalised expit and both running sums stay on the device in fp64 (`lbbnn_mc_accumulate_batched`).

Sample s of layer l draws from Philox stream  l * 2 + s * 2 n_layers  of the predictor's seed, whatever the launch width.
"""
import torch
import torch.nn.functional as F

from . import _capi as K
from .lrt import current_seed


class EnsemblePredictor:
    def __init__(self, net, batch, samples_per_launch=10, seed=None):
        K.require_device()
        self.net, self.layers, self.B = net, list(net.layers), int(batch)
        self.SB = max(1, int(samples_per_launch))
        self.seed = current_seed() if seed is None else int(seed)
        self.mnf = hasattr(self.layers[0], "z_flow")
        dev = self.layers[0].weight_mu.device
        if dev.type != "cuda":
            raise K.LbbnnError("EnsemblePredictor needs the network on a CUDA device (no CPU fallback)")
        self.device = dev
        f32 = dict(dtype=torch.float32, device=dev)
        self.sizes = sizes = [(l.in_features, l.out_features) for l in self.layers]
        B, SB, C = self.B, self.SB, sizes[-1][1]
        self.x = torch.zeros(B, sizes[0][0], **f32)
        self.x_rep = torch.zeros(SB * B, sizes[0][0], **f32) if self.mnf else None
        self.e1 = torch.zeros(B, sizes[0][1], **f32)
        self.var1 = torch.zeros(B, sizes[0][1], **f32)
        self.h = [torch.zeros(SB * B, o, **f32) for _, o in sizes]          # stacked activations; the last one = logits
        self.ws = torch.empty(max(K.lrt_workspace_bytes(SB * B, i, o) for i, o in sizes), dtype=torch.uint8, device=dev)
        f64 = dict(dtype=torch.float64, device=dev)
        self.sum_logp, self.sum_prob = torch.zeros(B, C, **f64), torch.zeros(B, C, **f64)
        self.first_logp, self.first_prob = torch.zeros(B, C, **f64), torch.zeros(B, C, **f64)
        self.kernels_per_launch = 0

    # ---- noise -------------------------------------------------------------------------------------------------------
    def _stream(self, layer, sample):
        return layer * 2 + sample * 2 * len(self.layers)

    def _group_noise_ok(self, out_features):
        return (self.B * out_features) % 4 == 0

    def _noise_for(self, li, s0, off, n, eps):
        """(lbbnn_noise, group stride, keepalive) for the n samples with global indices s0 .. of layer li (injected noise is
        indexed by `off`, the position inside this run() call)."""
        o = self.sizes[li][1]
        if eps is not None:
            e = eps[li][off:off + n].reshape(n * self.B, o).contiguous()
            return K.make_noise(e), 0, e
        stride = 2 * len(self.layers)
        if self._group_noise_ok(o):
            return K.make_noise(None, self.seed, self._stream(li, s0)), stride, None
        # odd shapes: materialise each sample's stream (what a batch-B call on that stream draws) and inject
        e = torch.cat([K.philox_normal((self.B, o), self.seed, self._stream(li, s0 + j), self.device) for j in range(n)])
        return K.make_noise(e), 0, e

    def _z_rows(self, li, off, n, z_noise):
        """MNF: the n samples' z of layer li = the flow image of one draw z0 per sample (MNF:182-187 with batch rows that are
        never used dropped).  z_noise[li] = {"eps_z": (S, in), "z_masks": [T x (S, in)]} injects the draws."""
        l = self.layers[li]
        if z_noise is not None:
            eps_z = z_noise[li]["eps_z"][off:off + n]
            masks = [m[off:off + n] for m in z_noise[li]["z_masks"]]
        else:
            eps_z, masks = torch.randn(n, l.in_features, device=self.device), None
        z0 = l.q0_mean + l.q0_log_var.exp().sqrt() * eps_z
        zs, _ = l.z_flow(z0, masks, per_row=True)
        return zs.contiguous()

    # ---- one launch: samples s0 .. s0 + n ----------------------------------------------------------------------------------
    def _launch(self, s0, off, n, eps, z_noise, first_k):
        st = K.current_stream()
        L, B = len(self.layers), self.B
        ws, wsn = self.ws.data_ptr(), self.ws.numel()
        nk = 0
        h = None
        for li, l in enumerate(self.layers):
            fi, fo = self.sizes[li]
            last = li == L - 1
            flags = K.FLAG_SAMPLE | (0 if last else K.FLAG_RELU)
            desc = K.make_layer(l.weight_mu.data, l.weight_rho.data, l.lambdal.data, l.bias_mu.data, l.bias_rho.data)
            nz, gstride, keep = self._noise_for(li, s0, off, n, eps)
            out = self.h[li][:n * B]
            if li == 0 and not self.mnf:
                # e_b, var_b were computed once for this test batch (run()); every sample adds its own sqrt(var_b) eps
                K.check(K.lib.lbbnn_lrt_sample_expand(K.ptr(self.e1), K.ptr(self.var1), B, fo, n, nz, gstride, flags, K.ptr(out), st))
                nk += 1
            else:
                xin = self.x_rep[:n * B] if li == 0 else h
                z = self._z_rows(li, off, n, z_noise) if self.mnf else None
                K.check(K.lib.lbbnn_lrt_f32_fwd_ex(desc, K.ptr(xin), n * B, nz, l.cfg.priors, l.cfg.var_mode, flags, K.ptr(out), None,
                                                   None, None, K.ptr(z, allow_none=True), B, gstride, ws, wsn, st))
                nk += 3
            h = out
        C = self.sizes[-1][1]
        K.check(K.lib.lbbnn_mc_accumulate_batched(K.ptr(h), n, B, C, self.sum_logp.data_ptr(), self.sum_prob.data_ptr(), None, st))
        nk += 1
        nf = min(n, first_k - s0)
        if nf > 0:     # the statistic over the FIRST first_k samples (outputs[0:10].mean(0), LRT:262)
            K.check(K.lib.lbbnn_mc_accumulate_batched(K.ptr(h), nf, B, C, self.first_logp.data_ptr(), self.first_prob.data_ptr(),
                                                      None, st))
            nk += 1
        self.kernels_per_launch = nk

    @torch.no_grad()
    def run(self, x, samples, first_sample=0, eps=None, z_noise=None, first_k=10):
        """Accumulate `samples` stochastic forwards (global sample indices first_sample ..) over the input batch x.
        eps: optional injected N(0,1) noise, one (S, batch, out) tensor per layer; z_noise: MNF draws (see _z_rows)."""
        st = K.current_stream()
        self.x.copy_(x.reshape(self.x.shape), non_blocking=True)
        for t in (self.sum_logp, self.sum_prob, self.first_logp, self.first_prob):
            t.zero_()
        if self.mnf:
            self.x_rep.view(self.SB, *self.x.shape).copy_(self.x.unsqueeze(0).expand(self.SB, *self.x.shape))
        else:
            l = self.layers[0]
            desc = K.make_layer(l.weight_mu.data, l.weight_rho.data, l.lambdal.data, l.bias_mu.data, l.bias_rho.data)
            K.check(K.lib.lbbnn_lrt_f32_fwd_ex(desc, K.ptr(self.x), self.B, K.make_noise(None, 0, 0), l.cfg.priors, l.cfg.var_mode,
                                               K.FLAG_SAMPLE | K.FLAG_MOMENTS, K.ptr(self.e1), K.ptr(self.var1), None, None, None, 0,
                                               0, self.ws.data_ptr(), self.ws.numel(), st))
        s = int(first_sample)
        end = s + int(samples)
        while s < end:
            n = min(self.SB, end - s)
            self._launch(s, s - int(first_sample), n, eps, z_noise, first_k)
            s += n
        self.samples = int(samples)

    # ---- the reference's statistics ------------------------------------------------------------------------------------------
    def result(self, total_samples=None, first_k=10):
        S = self.samples if total_samples is None else int(total_samples)
        k = min(S, first_k)
        probs = self.sum_prob / S
        return {"mean_logp": self.sum_logp / S, "mean_prob": probs,                                    # `mydata_means`, LRT:249-258
                "ensemble": (self.first_logp / k).argmax(1),                                             # LRT:262-263
                "entropy": -(probs * torch.log(probs)).sum(1)}                                           # LRT:325-330

    @torch.no_grad()
    def test_ensemble(self, x, samples, ensemble_first=10, eps=None, z_noise=None, mean_forward=None):
        """The per-batch statistics of `test_ensemble` (LRT:239-265 / MNF:287-318): mean_prob, ensemble (argmax of the mean
        of the first `ensemble_first` log-softmax outputs), posterior_mean (argmax of net(x, sample=False)), entropy."""
        self.run(x, samples, eps=eps, z_noise=z_noise, first_k=ensemble_first)
        out = self.result(samples, ensemble_first)
        was_training = self.net.training
        self.net.eval()
        mean_out = mean_forward(x) if mean_forward is not None else self.net(x, sample=False)           # LRT:264
        if was_training:
            self.net.train()
        out["posterior_mean"] = mean_out.argmax(1)
        return out

    @torch.no_grad()
    def outofsample(self, x, samples, eps=None, z_noise=None):
        """`outofsample` for one batch (LRT:305-341): predictive entropy of the averaged row-normalised expit over all
        samples, and the prediction from `outputs[1:TEST_SAMPLES].mean(0)` -- the reference skips sample 0 there."""
        self.run(x, samples, eps=eps, z_noise=z_noise, first_k=1)
        probs = self.sum_prob / samples
        rest = (self.sum_logp - self.first_logp) / max(samples - 1, 1)
        return {"entropy": -(probs * torch.log(probs)).sum(1), "pred": rest.argmax(1), "mean_prob": probs}


@torch.no_grad()
def mf_outofsample(net, x, samples, medimod=False, seed=None, samples_per_launch=32, predictor=None):
    """`outofsample` of the MF script for one batch (LBBNN-GP-MF.py:450-502) on the batched MC kernels (lbbnn.mf.MCPredictor):
    `samples` stochastic forwards with fresh weights; medimod=True fixes the inclusion masks to the median-probability model
    [alpha > 0.5] (MF:462-465) instead of drawing gamma ~ Bernoulli(alpha).  Returns the predictive entropy of the averaged
    row-normalised expit (MF:478-494), the prediction from outputs[1:TEST_SAMPLES].mean(0) (MF:496-498) and the predictor."""
    from . import mf
    mc = predictor or mf.MCPredictor(net, batch=x.shape[0], seed=seed, samples_per_launch=min(samples_per_launch, max(samples, 1)))
    masks = mf.median_probability_masks(net) if medimod else None
    mc.run(x, 1, first_sample=0, masks=masks)                    # sample 0 alone: the reference leaves it out of `output`
    l0, p0 = mc.sum_logp.clone(), mc.sum_prob.clone()
    if samples > 1:
        mc.run(x, samples - 1, first_sample=1, masks=masks)
        lr, pr = mc.sum_logp, mc.sum_prob
    else:
        lr, pr = torch.zeros_like(l0), torch.zeros_like(p0)
    probs = (p0 + pr) / samples
    return {"entropy": -(probs * torch.log(probs)).sum(1), "pred": (lr / max(samples - 1, 1)).argmax(1), "mean_prob": probs,
            "predictor": mc}
