"""Drop-in MNF layer and network (reference LBBNN-GP-MF-MNF.py:133-260; sim-study parametrisation
LBBNN-GP-MF-MNFsim_study.py:145-251 through the prior / init keyword arguments).

What runs where: the z flows and the auxiliary r flow -> fused flow kernels (flows.py); the activation
path mm(x*z, M^T), mm(x^2, V^T), eps/sqrt -> the fused LRT kernels with z folded into the weight prologue
(MNF:197-198: mm(x*z, M^T) == mm(x, (M*z)^T) because z is a single (in,) vector); the KL over all weights
with (mu*z2 - mu_p)^2 (MNF:230-233) and the W_mean/W_var moments -> prologue/finalize kernels; log q0, the two
auxiliary GEMVs r0_c @ W^T (MNF:216-217), tanh, the outer-product means and log r_b -> one fused forward and one
fused backward kernel (csrc/mnf_aux.cu).  Only z0 = q0_mean + sqrt(exp(q0_log_var)) eps and the final scalar sum
stay in torch.

Reference quirks kept: one z (the last batch row's) is broadcast over the batch, so only
that row is pushed through the flow; the KL branch draws its own z (self.z becomes (1,in)); z_b[-1] in
log r_b is the last ELEMENT of the flowed vector; log pi (not log 2 pi) in both Gaussians.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _capi as K
from .flows import PropagateFlow
from .lrt import BernoulliView, GaussianView, LayerConfig, _LRTFunction, current_seed, _layer_ids


class _MomentsKL(torch.autograd.Function):
    """(M0 = alpha mu, V = sigma^2 alpha^2, kl(weights with mu*z_kl, biases)) and their gradients."""

    @staticmethod
    def forward(ctx, wmu, wrho, lam, bmu, brho, z_kl, cfg):
        K.require_device()
        ps = [t.contiguous() for t in (wmu, wrho, lam, bmu, brho)]
        zc = z_kl.contiguous()
        layer = K.make_layer(*ps, None, zc)
        M, V = torch.empty_like(ps[0]), torch.empty_like(ps[0])
        kl = torch.zeros((), dtype=torch.float32, device=wmu.device)
        ws = K.workspace(1 << 20, wmu.device)
        K.check(K.lib.lbbnn_lrt_f32_prologue(layer, cfg.priors, cfg.var_mode, K.FLAG_SAMPLE, K.ptr(M), K.ptr(V), K.ptr(kl),
                                             ws.data_ptr(), ws.numel(), K.current_stream()))
        ctx.save_for_backward(*ps, zc)
        ctx.cfg = cfg
        return M, V, kl

    @staticmethod
    def backward(ctx, dM, dV, dkl):
        wmu, wrho, lam, bmu, brho, z = ctx.saved_tensors
        cfg = ctx.cfg
        layer = K.make_layer(wmu, wrho, lam, bmu, brho, None, z)
        dM = torch.zeros_like(wmu) if dM is None else dM.contiguous()
        dV = torch.zeros_like(wmu) if dV is None else dV.contiguous()
        colsum = torch.zeros(2 * wmu.shape[0], dtype=torch.float32, device=wmu.device)
        grads = [torch.empty_like(t) for t in (wmu, wrho, lam, bmu, brho)]
        dz = torch.zeros_like(z)
        use_kl = dkl is not None
        g = K.LayerGrads(*[K.ptr(t) for t in grads], None, K.ptr(dz))
        K.check(K.lib.lbbnn_lrt_f32_finalize(layer, K.ptr(dM), K.ptr(dV), K.ptr(colsum), cfg.priors, cfg.var_mode,
                                             K.FLAG_SAMPLE, K.ptr(dkl.contiguous().float()) if use_kl else None,
                                             1.0 if use_kl else 0.0, g, K.current_stream()))
        return (*grads, dz, None)


class _AuxKL(torch.autograd.Function):
    """log_q0 - log_rb of the KL branch (MNF:212-227 minus the flows) in one forward and one backward kernel
    (csrc/mnf_aux.cu).  Inputs: q0_mean, q0_log_var, z0 (the KL row's pre-flow draw), r0_c, r0_b1, r0_b2, z2 (its flow
    image), M0 = alpha mu, V, eps_r, z_b = r_flow(z2)."""

    @staticmethod
    def forward(ctx, q0_mean, q0_log_var, z0, r0_c, r0_b1, r0_b2, z2, M0, V, eps_r, z_b, ticket):
        K.require_device()
        ts = [t.contiguous() for t in (q0_mean, q0_log_var, z0, r0_c, r0_b1, r0_b2, z2, M0, V, eps_r, z_b)]
        O, D = ts[7].shape
        dev = ts[7].device
        out = torch.empty(3, dtype=torch.float32, device=dev)
        save = torch.empty(int(K.lib.lbbnn_mnf_aux_save_floats(O)), dtype=torch.float32, device=dev)
        aux = K.MnfAux(D, O, *[K.ptr(t) for t in ts])
        K.check(K.lib.lbbnn_mnf_aux_kl_fwd(aux, K.ptr(out), K.ptr(save), ticket.data_ptr(), K.current_stream()))
        ctx.save_for_backward(*ts, save)
        ctx.shapes = [t.shape for t in (q0_mean, q0_log_var, z0, r0_c, r0_b1, r0_b2, z2, z_b)]
        return out[0]

    @staticmethod
    def backward(ctx, g):
        *ts, save = ctx.saved_tensors
        O, D = ts[7].shape
        dev = ts[7].device
        vec = torch.empty(8, D, dtype=torch.float32, device=dev)
        dM0, dV = torch.empty_like(ts[7]), torch.empty_like(ts[8])
        aux = K.MnfAux(D, O, *[K.ptr(t) for t in ts])
        grads = K.MnfAuxGrads(*[vec[i].data_ptr() for i in range(8)], K.ptr(dM0), K.ptr(dV))
        K.check(K.lib.lbbnn_mnf_aux_kl_bwd(aux, K.ptr(save), K.ptr(g.contiguous().float()), grads, K.current_stream()))
        v = [vec[i].view(ctx.shapes[i]) for i in range(8)]
        #      q0_mean q0_lv z0    r0_c  b1    b2    z2    M0   V   eps_r z_b  ticket
        return v[0], v[1], v[2], v[3], v[4], v[5], v[6], dM0, dV, None, v[7], None


class BayesianLinear(nn.Module):
    """MNF layer, drop-in for LBBNN-GP-MF-MNF.py:133-239: ctor `(in_features, out_features, num_transforms)`.
    `noise=` on forward injects the draws of SURVEY.md §3.2 (see tests/cases.py:mnf_noise)."""

    def __init__(self, in_features, out_features, num_transforms=2, *, device=None, mu_prior=0.0, sigma_prior=1.0,
                 alpha_prior=0.05, bias_mu_prior=0.0, bias_sigma_prior=1.0, mu_init=0.01, lambda_init=(0.0, 1.0),
                 z_flow_type="RNVP", r_flow_type="RNVP", h_sizes=(75, 75, 75, 75)):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        E = torch.empty
        self.weight_mu = nn.Parameter(E(out_features, in_features).uniform_(-mu_init, mu_init))
        self.weight_rho = nn.Parameter(E(out_features, in_features).uniform_(-5, -4))
        self.lambdal = nn.Parameter(E(out_features, in_features).uniform_(*lambda_init))
        E(out_features, in_features).uniform_(0.999, 0.9999)            # the reference's alpha_q draw (MNF:149)
        self.bias_mu = nn.Parameter(E(out_features).uniform_(-0.2, 0.2))
        self.bias_rho = nn.Parameter(E(out_features).uniform_(-5, -4))
        self.q0_mean = nn.Parameter(0.1 * torch.randn(in_features))
        self.q0_log_var = nn.Parameter(-9 + 0.1 * torch.randn(in_features))
        self.r0_c = nn.Parameter(0.1 * torch.randn(in_features))
        self.r0_b1 = nn.Parameter(0.1 * torch.randn(in_features))
        self.r0_b2 = nn.Parameter(0.1 * torch.randn(in_features))
        self.z_flow = PropagateFlow(z_flow_type, in_features, num_transforms, h_sizes)
        self.r_flow = PropagateFlow(r_flow_type, in_features, num_transforms, h_sizes)
        self.cfg = LayerConfig(mu_prior, sigma_prior, alpha_prior, bias_mu_prior, bias_sigma_prior)
        self.weight = GaussianView(self.weight_mu, self.weight_rho)
        self.bias = GaussianView(self.bias_mu, self.bias_rho)
        self.gamma = BernoulliView(self)
        self.kl = 0
        self.z = 0
        self._uid = next(_layer_ids)
        self._calls = 0
        self.last_noise_key = None
        self._aux_ticket = None
        if device is not None:
            self.to(device)

    @property
    def alpha_q(self):
        return 1 / (1 + torch.exp(-self.lambdal.detach()))

    def _z0(self, eps):
        return self.q0_mean + self.q0_log_var.exp().sqrt() * eps       # MNF:183-185

    def _prepare(self, want_kl, nz):
        """Everything of forward() that does not depend on the input batch: the z flow on the live rows, and (training /
        calculate_log_probs) the whole KL branch with the auxiliary r flow.  Returns (z_k, kl or 0)."""
        dev = self.weight_mu.device
        D = self.in_features
        inj = "eps_z" in nz
        # every z-flow evaluation of this call as rows of ONE launch: [activation row (last batch row), KL row]
        eps_rows = [nz["eps_z"][-1:] if inj else torch.randn(1, D, device=dev)]
        mask_rows = [[m[-1:] for m in nz["z_masks"]]] if inj else None
        if want_kl:
            eps_rows.append(nz["eps_z2"] if inj else torch.randn(1, D, device=dev))
            if inj:
                mask_rows.append(list(nz["z_masks2"]))
        z0 = self._z0(torch.cat(eps_rows, 0))
        masks = [torch.cat([r[t] for r in mask_rows], 0) for t in range(len(mask_rows[0]))] if inj else None
        # per-row log-determinants: the rows are independent evaluations (the reference calls sample_z() once per row set,
        # MNF:194 and MNF:210), so the IAF kind's "sum over everything" (flows2:241) must not mix them
        zs, logdets = self.z_flow(z0, masks, per_row=True)
        z_k = zs[0]
        if not want_kl:
            self.z = z0[:1]
            return z_k, 0
        self.z = z0[1:2]                                                # sample_z() overwrites self.z with (1,in)
        z2 = zs[1]
        log_det_q = logdets[1]                                          # the KL row's own log-det (sample_z() with batch 1)
        M0, V, kl_wb = _MomentsKL.apply(self.weight_mu, self.weight_rho, self.lambdal, self.bias_mu, self.bias_rho, z2, self.cfg)
        eps_r = nz["eps_r"] if "eps_r" in nz else torch.randn(self.out_features, device=dev)
        z_b, log_det_r = self.r_flow(z2, nz.get("r_masks"))
        if self._aux_ticket is None or self._aux_ticket.device != dev:
            self._aux_ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        # log_q0 (MNF:212-214) - log_rb (MNF:216-227): the GEMVs r0_c @ W_mean^T, r0_c^2 @ W_var^T with W_mean = z2 mu alpha
        # (z2 folded into the vector side), tanh, the outer-product means and the two Gaussian sums, fused
        aux = _AuxKL.apply(self.q0_mean, self.q0_log_var, self.z, self.r0_c, self.r0_b1, self.r0_b2, z2, M0, V, eps_r, z_b,
                           self._aux_ticket)
        return z_k, kl_wb + aux - log_det_q - log_det_r

    def _activation(self, input, z_k, sample_branch, nz):
        self._calls += 1
        self.last_noise_key = (current_seed(), (self._uid << 40) | self._calls)
        act, _ = _LRTFunction.apply(input, self.weight_mu, self.weight_rho, self.lambdal, self.bias_mu, self.bias_rho, z_k,
                                    nz.get("eps"), self.cfg, sample_branch, False, self.last_noise_key)
        return act

    def forward(self, input, sample=False, calculate_log_probs=False, noise=None):
        nz = noise or {}
        z_k, self.kl = self._prepare(self.training or calculate_log_probs, nz)
        return self._activation(input, z_k, self.training or sample, nz)


class BayesianNetwork(nn.Module):
    """784-400-600-10 MNF network, drop-in for LBBNN-GP-MF-MNF.py:244-260."""

    def __init__(self, sizes=(28 * 28, 400, 600, 10), num_transforms=2, **layer_kwargs):
        super().__init__()
        self.sizes = tuple(sizes)
        for n, (i, o) in enumerate(zip(sizes[:-1], sizes[1:]), 1):
            setattr(self, f"l{n}", BayesianLinear(i, o, num_transforms=num_transforms, **layer_kwargs))
        self._names = [f"l{n}" for n in range(1, len(sizes))]
        self._streams = None

    @property
    def layers(self):
        return [getattr(self, n) for n in self._names]

    def forward(self, x, sample=False, noises=None, calculate_log_probs=False):
        """The flows and the KL branch of a layer depend on parameters and noise only, never on the activations: they
        are issued for all layers first, each on its own CUDA stream (concurrent GEMV chains instead of one after the
        other; inside a captured graph they become parallel branches), then the LRT layers run on the caller's stream."""
        x = x.view(-1, self.sizes[0])
        ls = self.layers
        nzs = [(None if noises is None else noises[i]) or {} for i in range(len(ls))]
        cur = torch.cuda.current_stream()
        if self._streams is None or self._streams[0].device != x.device:
            self._streams = [torch.cuda.Stream(device=x.device) for _ in ls]
        prepared = []
        for l, s, nz in zip(ls, self._streams, nzs):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                z_k, kl = l._prepare(l.training or calculate_log_probs, nz)
            prepared.append((z_k, kl))
        for i, (l, s, nz) in enumerate(zip(ls, self._streams, nzs)):
            cur.wait_stream(s)
            z_k, l.kl = prepared[i]
            z_k.record_stream(cur)
            if torch.is_tensor(l.kl):
                l.kl.record_stream(cur)
            x = l._activation(x, z_k, l.training or sample, nz)
            x = F.relu(x) if i < len(ls) - 1 else F.log_softmax(x, dim=1)
        return x

    def kl(self):
        return sum(l.kl for l in self.layers)
