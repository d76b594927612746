"""Drop-in `PropagateFlow` (reference flows2.py:14-46, flows_simstudy.py) for the two transforms on the hot
path: 'RNVP' (flows2:188-219) and 'MNF' (the IAF-style masked flow, flows2:225-241).  Sub-module and
parameter names match the reference (`transforms.{i}.network.{0,2,4,6}`, `.t`, `.s` / `.f`, `.g`, `.k`), so
state_dicts interchange.  forward/backward run in the fused flow kernels of liblbbnn (csrc/flows.cu).
The other flow types of flows2.py (Planar/Radial/Sylvester/Householder) are never selected by the reference
scripts (Z_FLOW_TYPE = R_FLOW_TYPE = 'RNVP', MNF:46-47) and are out of scope."""
import itertools

import torch
import torch.nn as nn

from . import _capi as K
from .lrt import current_seed

_flow_ids = itertools.count(1)


class RNVP(nn.Module):
    def __init__(self, dim, h_sizes=(75, 75, 75, 75)):
        super().__init__()
        sizes = [dim] + list(h_sizes)
        layers = []
        for a, b in zip(sizes[:-1], sizes[1:]):
            layers += [nn.Linear(a, b), nn.LeakyReLU(0.1)]
        self.network = nn.Sequential(*layers[:-1])      # last activation dropped (flows2:185)
        self.t = nn.Linear(sizes[-1], dim)
        self.s = nn.Linear(sizes[-1], dim)

    def linears(self):
        return [m for m in self.network if isinstance(m, nn.Linear)], self.t, self.s


class IAF(nn.Module):
    """flows2.MNF."""

    def __init__(self, dim, hidden=100):
        super().__init__()
        self.f = nn.Linear(dim, hidden)
        self.g = nn.Linear(hidden, dim)
        self.k = nn.Linear(hidden, dim)

    def linears(self):
        return [self.f], self.g, self.k


class _FlowFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, masks, kind, n_hidden, key, *params):
        K.require_device()
        z = z.contiguous()
        R, D = z.shape
        per_t = 2 * (n_hidden + 2)
        T = len(params) // per_t
        params = [p.contiguous() for p in params]
        flow = _build_flow(kind, D, T, n_hidden, params)
        zo, ld = torch.empty_like(z), torch.empty(R, dtype=torch.float32, device=z.device)
        need_bwd = any(ctx.needs_input_grad)
        save = (torch.empty(int(K.lib.lbbnn_flow_save_floats(flow, R)), dtype=torch.float32, device=z.device)
                if need_bwd else None)
        masks = masks.contiguous() if masks is not None else None
        noise = K.make_noise(None, key[0], key[1])
        K.check(K.lib.lbbnn_flow_fwd(flow, K.ptr(z), R, K.ptr(masks, allow_none=True), noise, K.ptr(zo), K.ptr(ld),
                                     K.ptr(save, allow_none=True), K.current_stream()))
        ctx.save_for_backward(save, masks, *params)
        ctx.meta = (kind, D, T, n_hidden, key, R)
        return zo, ld

    @staticmethod
    def backward(ctx, dz, dld):
        save, masks, *params = ctx.saved_tensors
        kind, D, T, n_hidden, key, R = ctx.meta
        flow = _build_flow(kind, D, T, n_hidden, params)
        sizes = [p.numel() for p in params]
        P = sum(sizes)
        gbuf = torch.empty(R, P, dtype=torch.float32, device=save.device)
        grads = K.FlowGrads()
        grads.row_stride = P
        offs = list(itertools.accumulate([0] + sizes))
        base = gbuf.data_ptr()
        i = 0
        for t in range(T):
            for l in range(n_hidden):
                grads.t[t].hidden[l].dW, grads.t[t].hidden[l].db = base + 4 * offs[i], base + 4 * offs[i + 1]
                i += 2
            grads.t[t].shift.dW, grads.t[t].shift.db = base + 4 * offs[i], base + 4 * offs[i + 1]
            grads.t[t].scale.dW, grads.t[t].scale.db = base + 4 * offs[i + 2], base + 4 * offs[i + 3]
            i += 4
        dzi = torch.empty(R, D, dtype=torch.float32, device=save.device)
        noise = K.make_noise(None, key[0], key[1])
        dz_c = dz.contiguous() if dz is not None else None
        dld_c = dld.contiguous() if dld is not None else None
        K.check(K.lib.lbbnn_flow_bwd(flow, grads, R, K.ptr(masks, allow_none=True), noise, K.ptr(dz_c, allow_none=True),
                                     K.ptr(dld_c, allow_none=True), K.ptr(save), K.ptr(dzi), K.current_stream()))
        g = gbuf.sum(0) if R > 1 else gbuf[0]
        pg = [g[offs[j]:offs[j + 1]].view_as(params[j]) for j in range(len(params))]
        return (dzi, None, None, None, None, *pg)


def _build_flow(kind, D, T, n_hidden, params):
    flow = K.Flow()
    flow.kind, flow.dim, flow.n_transforms, flow.n_hidden = kind, D, T, n_hidden
    i = 0

    def lin(dst, w, b):
        dst.W, dst.b, dst.in_, dst.out = K.ptr(w), K.ptr(b), w.shape[1], w.shape[0]
    for t in range(T):
        for l in range(n_hidden):
            lin(flow.t[t].hidden[l], params[i], params[i + 1])
            i += 2
        lin(flow.t[t].shift, params[i], params[i + 1])
        lin(flow.t[t].scale, params[i + 2], params[i + 3])
        i += 4
    return flow


class PropagateFlow(nn.Module):
    def __init__(self, transform, dim, num_transforms, h_sizes=(75, 75, 75, 75), hidden=100):
        super().__init__()
        if transform == "RNVP":
            self.transforms = nn.ModuleList([RNVP(dim, h_sizes) for _ in range(num_transforms)])
            self.kind, self.n_hidden = K.FLOW_RNVP, len(h_sizes)
        elif transform == "MNF":
            self.transforms = nn.ModuleList([IAF(dim, hidden) for _ in range(num_transforms)])
            self.kind, self.n_hidden = K.FLOW_IAF, 1
        else:
            raise NotImplementedError(f"flow type {transform!r}: only 'RNVP' and 'MNF' are on the reference's hot path")
        if num_transforms > K.FLOW_MAX_T or self.n_hidden > K.FLOW_MAX_HIDDEN:
            raise ValueError("too many transforms / hidden layers for the fused kernel")
        self.dim = dim
        self._uid = next(_flow_ids)
        self._calls = 0
        self.last_noise_key = None

    def _params(self):
        out = []
        for tr in self.transforms:
            hidden, a, b = tr.linears()
            for m in hidden + [a, b]:
                out += [m.weight, m.bias]
        return out

    def _next_key(self):
        """Philox (seed, stream) of the next evaluation's native masks (transform t draws from stream + t)."""
        self._calls += 1
        self.last_noise_key = (current_seed(), (self._uid << 44) | (self._calls << 8))
        return self.last_noise_key

    def forward(self, z, masks=None, per_row=False):
        """z: (B, dim) or (dim,).  masks: optional injected list/tensor of {0,1} masks, one per transform, each
        shaped like z.  Returns (z_out, logdet) with the reference's shapes: logdet (B,) / scalar for RNVP, a
        scalar summed over everything for the 'MNF' kind (flows2:241).  per_row=True returns the (B,) per-row
        log-determinants for either kind (a caller that batched independent flow evaluations into one launch
        takes the rows apart itself)."""
        one_d = z.dim() == 1
        z2 = z.reshape(1, -1) if one_d else z
        if masks is not None:
            masks = torch.stack([m.reshape(z2.shape) for m in masks]) if not torch.is_tensor(masks) else masks.reshape(-1, *z2.shape)
        zo, ld = _FlowFunction.apply(z2, masks, self.kind, self.n_hidden, self._next_key(), *self._params())
        if per_row:
            return (zo[0], ld[0]) if one_d else (zo, ld)
        if self.kind == K.FLOW_IAF:
            return (zo[0] if one_d else zo), ld.sum()
        return (zo[0], ld[0]) if one_d else (zo, ld)
