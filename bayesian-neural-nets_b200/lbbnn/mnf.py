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
from .flows import PropagateFlow, _build_flow
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


_N_LAYER_PARAMS = 10      # weight_mu, weight_rho, lambdal, bias_mu, bias_rho, q0_mean, q0_log_var, r0_c, r0_b1, r0_b2
_zero_colsums = {}


def _zeros_const(n, dev):
    """A never-written zero buffer (the weight-KL finalize's `colsum` input: the KL branch has no activation gradient)."""
    key = (dev.index, n)
    t = _zero_colsums.get(key)
    if t is None:
        if torch.cuda.is_current_stream_capturing():
            raise K.LbbnnError("constant buffer would be created during CUDA-graph capture; run one eager step first")
        t = torch.zeros(n, dtype=torch.float32, device=dev)
        _zero_colsums[key] = t
    return t


def _flow_grad_table(flow, params, gbuf):
    """lbbnn_flow_grads over a (rows, P) buffer laid out like `params` (PropagateFlow._params() order); returns it with
    the per-parameter offsets."""
    sizes = [p.numel() for p in params]
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + n)
    grads = K.FlowGrads()
    grads.row_stride = gbuf.shape[1]
    base = gbuf.data_ptr()
    i = 0
    for t in range(flow.n_transforms):
        for l in range(flow.n_hidden):
            grads.t[t].hidden[l].dW, grads.t[t].hidden[l].db = base + 4 * offs[i], base + 4 * offs[i + 1]
            i += 2
        grads.t[t].shift.dW, grads.t[t].shift.db = base + 4 * offs[i], base + 4 * offs[i + 1]
        grads.t[t].scale.dW, grads.t[t].scale.db = base + 4 * offs[i + 2], base + 4 * offs[i + 3]
        i += 4
    return grads, offs


class _Prepare(torch.autograd.Function):
    """Everything of an MNF layer call that does not depend on the input batch, as ONE autograd node (MNF:183-194 for the
    activation row and the whole KL branch MNF:208-235): the z0 draw, the z flow on [activation row, KL row], the weight
    moments + weight/bias KL with the KL row's z, the auxiliary r flow, log q0 - log r_b, and the combination into the
    layer's kl.  Forward = 6 launches, backward = 9 (eager formulation with one autograd node per piece: ~45 per layer,
    most of them elementwise glue and gradient accumulation).  Returns (z_k, kl, z0 row the reference leaves in self.z)."""

    @staticmethod
    def forward(ctx, meta, *params):
        K.require_device()
        layer, want_kl, nz = meta
        ps = [p.contiguous() for p in params]
        wmu, wrho, lam, bmu, brho, q0m, q0lv, r0c, rb1, rb2 = ps[:_N_LAYER_PARAMS]
        nzp = len(layer.z_flow._params())
        zp, rp = ps[_N_LAYER_PARAMS:_N_LAYER_PARAMS + nzp], ps[_N_LAYER_PARAMS + nzp:]
        dev, D, O = wmu.device, layer.in_features, layer.out_features
        f32 = dict(dtype=torch.float32, device=dev)
        st = K.current_stream()
        R = 2 if want_kl else 1
        inj = "eps_z" in nz
        # ---- z0 rows [activation row (the last batch row's draw), KL row] + eps_r
        layer._prep_calls += 1
        skey = (layer._uid << 40) | (1 << 39) | (layer._prep_calls << 2)
        if inj:
            eps_rows = torch.cat([nz["eps_z"][-1:], nz["eps_z2"]], 0) if want_kl else nz["eps_z"][-1:].contiguous()
            noise_z = K.make_noise(eps_rows.contiguous())
        else:
            noise_z = K.make_noise(None, current_seed(), skey)
        eps, z0 = torch.empty(R, D, **f32), torch.empty(R, D, **f32)
        eps_r = torch.empty(O, **f32) if want_kl else None
        noise_r = K.make_noise(nz["eps_r"].contiguous()) if "eps_r" in nz else K.make_noise(None, current_seed(), skey + 1)
        K.check(K.lib.lbbnn_mnf_draw(K.ptr(q0m), K.ptr(q0lv), noise_z, R, D, K.ptr(eps), K.ptr(z0), noise_r, O,
                                     K.ptr(eps_r, allow_none=True), st))
        # ---- z flow on both rows (per-row log-determinants: the rows are independent sample_z() calls, MNF:194 / MNF:210)
        zf, rf = layer.z_flow, layer.r_flow
        masks_z = None
        if inj:
            rows = [[m[-1:] for m in nz["z_masks"]]] + ([list(nz["z_masks2"])] if want_kl else [])
            masks_z = torch.stack([torch.cat([r[t].reshape(1, D) for r in rows], 0) for t in range(len(rows[0]))]).contiguous()
        key_z = zf._next_key()
        flow_z = _build_flow(zf.kind, D, len(zf.transforms), zf.n_hidden, zp)
        zs, ld = torch.empty(R, D, **f32), torch.empty(R, **f32)
        save_z = torch.empty(int(K.lib.lbbnn_flow_save_floats(flow_z, R)), **f32)
        K.check(K.lib.lbbnn_flow_fwd(flow_z, K.ptr(z0), R, K.ptr(masks_z, allow_none=True), K.make_noise(None, *key_z),
                                     K.ptr(zs), K.ptr(ld), K.ptr(save_z), st))
        ctx.meta = (layer, want_kl, key_z)
        if not want_kl:
            ctx.save_for_backward(*ps, eps, masks_z, save_z)
            z_row, kl0 = z0[:1], z0.new_zeros(())
            ctx.mark_non_differentiable(z_row, kl0)
            return zs[0], kl0, z_row
        z2 = zs[1]
        # ---- weight moments + weight / bias KL with the KL row's z (MNF:211, 228-234)
        M0, V = torch.empty_like(wmu), torch.empty_like(wmu)
        kls = torch.empty(4, **f32)                       # [kl_wb, kl, unused, unused]
        out3 = torch.empty(3, **f32)
        ws = K.workspace(1 << 20, dev)
        lay = K.make_layer(wmu, wrho, lam, bmu, brho, None, z2)
        K.check(K.lib.lbbnn_lrt_f32_prologue(lay, layer.cfg.priors, layer.cfg.var_mode, K.FLAG_SAMPLE, K.ptr(M0), K.ptr(V),
                                             kls.data_ptr(), ws.data_ptr(), ws.numel(), st))
        # ---- auxiliary r flow on z2 (MNF:222-223)
        masks_r = None
        if "r_masks" in nz:
            masks_r = torch.stack([m.reshape(1, D) for m in nz["r_masks"]]).contiguous()
        key_r = rf._next_key()
        flow_r = _build_flow(rf.kind, D, len(rf.transforms), rf.n_hidden, rp)
        z_b, ld_r = torch.empty(1, D, **f32), torch.empty(1, **f32)
        save_r = torch.empty(int(K.lib.lbbnn_flow_save_floats(flow_r, 1)), **f32)
        K.check(K.lib.lbbnn_flow_fwd(flow_r, z2.data_ptr(), 1, K.ptr(masks_r, allow_none=True), K.make_noise(None, *key_r),
                                     K.ptr(z_b), K.ptr(ld_r), K.ptr(save_r), st))
        # ---- log q0 - log r_b (MNF:212-227) and the layer's kl (MNF:235)
        if layer._aux_ticket is None or layer._aux_ticket.device != dev:
            layer._aux_ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        save_a = torch.empty(int(K.lib.lbbnn_mnf_aux_save_floats(O)), **f32)
        aux = K.MnfAux(D, O, K.ptr(q0m), K.ptr(q0lv), z0[1].data_ptr(), K.ptr(r0c), K.ptr(rb1), K.ptr(rb2), z2.data_ptr(),
                       K.ptr(M0), K.ptr(V), K.ptr(eps_r), K.ptr(z_b))
        K.check(K.lib.lbbnn_mnf_aux_kl_fwd(aux, K.ptr(out3), K.ptr(save_a), layer._aux_ticket.data_ptr(), st))
        K.check(K.lib.lbbnn_mnf_kl_combine(kls.data_ptr(), out3.data_ptr(), ld[1:].data_ptr(), ld_r.data_ptr(),
                                           kls[1:].data_ptr(), st))
        ctx.meta = (layer, want_kl, key_z, key_r)
        ctx.save_for_backward(*ps, eps, masks_z, save_z, z0, zs, M0, V, eps_r, z_b, masks_r, save_r, save_a)
        z_row = z0[1:2]
        ctx.mark_non_differentiable(z_row)
        return zs[0], kls[1], z_row

    @staticmethod
    def backward(ctx, d_zk, d_kl, _d_z0):
        layer, want_kl, key_z = ctx.meta[:3]
        zf, rf = layer.z_flow, layer.r_flow
        nzp = len(zf._params())
        saved = ctx.saved_tensors
        n_par = _N_LAYER_PARAMS + nzp + len(rf._params())
        ps = saved[:n_par]
        wmu, wrho, lam, bmu, brho, q0m, q0lv, r0c, rb1, rb2 = ps[:_N_LAYER_PARAMS]
        zp, rp = ps[_N_LAYER_PARAMS:_N_LAYER_PARAMS + nzp], ps[_N_LAYER_PARAMS + nzp:]
        dev, D, O = wmu.device, layer.in_features, layer.out_features
        f32 = dict(dtype=torch.float32, device=dev)
        st = K.current_stream()
        dz_k = d_zk.contiguous() if d_zk is not None else None
        flow_z = _build_flow(zf.kind, D, len(zf.transforms), zf.n_hidden, zp)
        Pz = sum(p.numel() for p in zp)
        d_mean, d_lv = torch.empty(D, **f32), torch.empty(D, **f32)
        if not want_kl:
            eps, masks_z, save_z = saved[n_par:]
            rows = dz_k.reshape(1, D) if dz_k is not None else torch.zeros(1, D, **f32)
            gz = torch.empty(1, Pz, **f32)
            gt, offs = _flow_grad_table(flow_z, zp, gz)
            dz0 = torch.empty(1, D, **f32)
            K.check(K.lib.lbbnn_flow_bwd(flow_z, gt, 1, K.ptr(masks_z, allow_none=True), K.make_noise(None, *key_z), K.ptr(rows),
                                         None, K.ptr(save_z), K.ptr(dz0), st))
            K.check(K.lib.lbbnn_mnf_draw_bwd(K.ptr(q0lv), K.ptr(eps), K.ptr(dz0), 1, D, -1, None, None, None, K.ptr(d_mean),
                                             K.ptr(d_lv), st))
            gzp = [gz[0, offs[j]:offs[j + 1]].view_as(zp[j]) for j in range(len(zp))]
            return (None, None, None, None, None, None, d_mean, d_lv, None, None, None, *gzp, *([None] * len(rp)))
        key_r = ctx.meta[3]
        eps, masks_z, save_z, z0, zs, M0, V, eps_r, z_b, masks_r, save_r, save_a = saved[n_par:]
        z2 = zs[1]
        g = d_kl.reshape(1).contiguous().float() if d_kl is not None else torch.zeros(1, **f32)
        dld = torch.empty(2, **f32)
        K.check(K.lib.lbbnn_mnf_bwd_rows(K.ptr(g), K.ptr(dld), None, None, None, D, None, st))
        # auxiliary term: every direct gradient + the rank-one dM0, dV
        vec = torch.empty(8, D, **f32)
        dM0, dV = torch.empty_like(M0), torch.empty_like(V)
        aux = K.MnfAux(D, O, K.ptr(q0m), K.ptr(q0lv), z0[1].data_ptr(), K.ptr(r0c), K.ptr(rb1), K.ptr(rb2), z2.data_ptr(),
                       K.ptr(M0), K.ptr(V), K.ptr(eps_r), K.ptr(z_b))
        ag = K.MnfAuxGrads(*[vec[i].data_ptr() for i in range(8)], K.ptr(dM0), K.ptr(dV))
        K.check(K.lib.lbbnn_mnf_aux_kl_bwd(aux, K.ptr(save_a), K.ptr(g), ag, st))
        # r flow: d z_b from the auxiliary term, d log_det_r = -g
        flow_r = _build_flow(rf.kind, D, len(rf.transforms), rf.n_hidden, rp)
        gr = torch.empty(1, sum(p.numel() for p in rp), **f32)
        grt, roffs = _flow_grad_table(flow_r, rp, gr)
        dzr = torch.empty(1, D, **f32)
        K.check(K.lib.lbbnn_flow_bwd(flow_r, grt, 1, K.ptr(masks_r, allow_none=True), K.make_noise(None, *key_r),
                                     vec[7].data_ptr(), dld[1:].data_ptr(), K.ptr(save_r), K.ptr(dzr), st))
        # z flow's output gradient rows: [d z_k ; d z2 = auxiliary + r flow (+ the weight KL, accumulated by finalize)]
        rows = torch.empty(2, D, **f32)
        K.check(K.lib.lbbnn_mnf_bwd_rows(None, None, K.ptr(dz_k, allow_none=True), vec[6].data_ptr(), K.ptr(dzr), D, K.ptr(rows), st))
        grads = [torch.empty_like(t) for t in (wmu, wrho, lam, bmu, brho)]
        lay = K.make_layer(wmu, wrho, lam, bmu, brho, None, z2)
        lg = K.LayerGrads(*[K.ptr(t) for t in grads], None, rows[1].data_ptr())
        K.check(K.lib.lbbnn_lrt_f32_finalize(lay, K.ptr(dM0), K.ptr(dV), K.ptr(_zeros_const(2 * O, dev)), layer.cfg.priors,
                                             layer.cfg.var_mode, K.FLAG_SAMPLE, K.ptr(g), 1.0, lg, st))
        gz = torch.empty(2, Pz, **f32)
        gt, offs = _flow_grad_table(flow_z, zp, gz)
        dz0 = torch.empty(2, D, **f32)
        K.check(K.lib.lbbnn_flow_bwd(flow_z, gt, 2, K.ptr(masks_z, allow_none=True), K.make_noise(None, *key_z), K.ptr(rows),
                                     K.ptr(dld), K.ptr(save_z), K.ptr(dz0), st))
        gzs = gz[0] + gz[1]
        K.check(K.lib.lbbnn_mnf_draw_bwd(K.ptr(q0lv), K.ptr(eps), K.ptr(dz0), 2, D, 1, vec[0].data_ptr(), vec[1].data_ptr(),
                                         vec[2].data_ptr(), K.ptr(d_mean), K.ptr(d_lv), st))
        gzp = [gzs[offs[j]:offs[j + 1]].view_as(zp[j]) for j in range(len(zp))]
        grp = [gr[0, roffs[j]:roffs[j + 1]].view_as(rp[j]) for j in range(len(rp))]
        return (None, *grads, d_mean, d_lv, vec[3], vec[4], vec[5], *gzp, *grp)


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
        self._prep_calls = 0
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
        calculate_log_probs) the whole KL branch with the auxiliary r flow -- one fused autograd node (_Prepare).
        Returns (z_k, kl or 0)."""
        params = [self.weight_mu, self.weight_rho, self.lambdal, self.bias_mu, self.bias_rho, self.q0_mean, self.q0_log_var,
                  self.r0_c, self.r0_b1, self.r0_b2, *self.z_flow._params(), *self.r_flow._params()]
        z_k, kl, z_row = _Prepare.apply((self, bool(want_kl), nz), *params)
        self.z = z_row                                                  # sample_z() leaves its last draw in self.z
        return z_k, (kl if want_kl else 0)

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
