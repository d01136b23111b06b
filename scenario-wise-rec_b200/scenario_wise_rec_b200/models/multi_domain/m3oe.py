"""M3oE (reference: scenario_wise_rec/models/multi_domain/m3oe.py:8-198).

state_dict keys: ``embedding.*``, ``_weight_{exp_d,exp_t,bal_d,bal_t}.deep_weights``, ``skip_conn.domain_specific.*``,
``shared_weight/bias``, ``slot_weight/bias.<d>``, ``star_mlp.domain_specific.*``, ``expert.<e>.domain_specific.*``,
``domain_expert.<d>.domain_specific.*``, ``gate.<d>.0.*``, ``tower.<d>.{0,1,3}.*``.
``_weight_exp_t`` / ``_weight_bal_t`` are never used by the forward (their ``.grad`` stays ``None``).
There is no BatchNorm: every normalisation is a row-local LayerNorm, so the stack has no batch coupling.
"""
import numpy as np
import torch
from torch import nn

from ... import _native as N
from ...basic.layers import EmbeddingLayer
from ._base import MultiDomainModel


class Weights(nn.Module):
    """Scalar mixing weight (m3oe.py:8-42); only softmax_type 3 (sigmoid) is reachable in the reference."""

    def __init__(self, weight_shape, tau, tau_step, initial_deep, softmax_type=2):
        super().__init__()
        assert isinstance(weight_shape, (int, list))
        norm = weight_shape[-1] if isinstance(weight_shape, list) else weight_shape
        if initial_deep is None:
            initial_deep = np.ones(weight_shape, dtype=np.float32) / norm
        else:
            initial_deep = np.ones(weight_shape, dtype=np.float32) * initial_deep
        self.deep_weights = nn.Parameter(torch.from_numpy(initial_deep), requires_grad=True)
        self.softmax_type = softmax_type
        self.tau = tau
        self.tau_step = tau_step


class Mlp_N(nn.Module):
    """[Linear -> LayerNorm -> ReLU] per consecutive pair of ``fcn_dim`` (m3oe.py:45-68); parameter container."""

    def __init__(self, fcn_dim):
        super().__init__()
        self.fcn_dim = fcn_dim
        self.n = len(fcn_dim)
        self.domain_specific = nn.ModuleList()
        for i in range(self.n - 1):
            self.domain_specific.append(nn.Linear(fcn_dim[i], fcn_dim[i + 1]))
            self.domain_specific.append(nn.LayerNorm(fcn_dim[i + 1]))
            self.domain_specific.append(nn.ReLU())

    def layers(self):
        return [(self.domain_specific[3 * i], self.domain_specific[3 * i + 1]) for i in range(self.n - 1)]


class M3oE(MultiDomainModel):
    def __init__(self, features, domain_num, fcn_dims, expert_num, exp_d, exp_t, bal_d, bal_t, tau=1, task_num=1,
                 tau_step=0.00005, softmax_type=3, device="cpu"):
        super().__init__()
        if softmax_type != 3:
            raise NotImplementedError("only softmax_type=3 is reachable in the reference (the others assert 0)")
        self.features = features
        self.input_dim = sum(fea.embed_dim for fea in features)
        self.layer_num = len(fcn_dims) + 1
        self.fcn_dim = [self.input_dim] + list(fcn_dims)
        self.domain_num = domain_num
        self.task_num = task_num
        self.expert_num = expert_num
        self.embedding = EmbeddingLayer(features)
        self.device = device
        self._weight_exp_d = Weights(1, tau, tau_step, exp_d, softmax_type)
        self._weight_exp_t = Weights(1, tau, tau_step, exp_t, softmax_type)
        self._weight_bal_d = Weights(1, tau, tau_step, bal_d, softmax_type)
        self._weight_bal_t = Weights(1, tau, tau_step, bal_t, softmax_type)
        assert len(self.fcn_dim) > 3, "too few layers assigned, must larger than 3. Star owns 3 layers, mmoe owns the rest."
        self.star_dim = self.fcn_dim[:3]
        self.fcn_dim = self.fcn_dim[3:]
        if len(self.fcn_dim) < 2:
            raise NotImplementedError("M3oE needs at least one expert layer (len(fcn_dims) >= 4)")
        self.skip_conn = Mlp_N([self.star_dim[0], self.star_dim[2]])
        self.shared_weight = nn.Parameter(torch.empty(self.star_dim[0], self.star_dim[1]))
        self.shared_bias = nn.Parameter(torch.zeros(self.star_dim[1]))
        self.slot_weight = nn.ParameterList(
            [nn.Parameter(torch.empty(self.star_dim[0], self.star_dim[1])) for _ in range(domain_num)])
        self.slot_bias = nn.ParameterList([nn.Parameter(torch.zeros(self.star_dim[1])) for _ in range(domain_num)])
        self.star_mlp = Mlp_N([self.star_dim[1], self.star_dim[2]])
        torch.nn.init.xavier_uniform_(self.shared_weight.data)
        for m in self.slot_weight:
            torch.nn.init.xavier_uniform_(m.data)
        self.expert = nn.ModuleList(Mlp_N(self.fcn_dim) for _ in range(expert_num))
        self.domain_expert = nn.ModuleList(Mlp_N(self.fcn_dim) for _ in range(domain_num))
        self.gate = nn.ModuleList(
            nn.Sequential(nn.Linear(self.fcn_dim[0], expert_num), nn.Softmax(dim=1)) for _ in range(domain_num))
        self.tower = nn.ModuleList(
            nn.Sequential(nn.Linear(self.fcn_dim[-1], self.fcn_dim[-1]), nn.LayerNorm(self.fcn_dim[-1]), nn.ReLU(),
                          nn.Linear(self.fcn_dim[-1], 1)) for _ in range(domain_num))

    @staticmethod
    def _ln(y, ln):
        return (y, ln.weight, ln.bias, ln.eps, N.ACT_RELU)

    def _lower(self, b, col_dtypes):
        D = self.domain_num
        dom = self._dom_dtype(col_dtypes)
        x = self.embedding.lower(b, self.features, col_dtypes)
        (sk_lin, sk_ln), = self.skip_conn.layers()
        # skip-connection Linear and the D STAR-style slot layers read the same input: one grouped launch
        outs = b.fc([dict(src=x, W=sk_lin.weight, b=sk_lin.bias)] +
                    [dict(src=x, W=self.slot_weight[d], W2=self.shared_weight, b=self.slot_bias[d], b2=self.shared_bias,
                          layout=N.W_KN) for d in range(D)])
        skip = b.layernorm([self._ln(outs[0], sk_ln)])[0]
        emb0 = b.select(outs[1:], dom)
        (sm_lin, sm_ln), = self.star_mlp.layers()
        y1 = b.fc([dict(src=emb0, W=sm_lin.weight, b=sm_lin.bias)])[0]
        emb = b.ew(N.EW_ADD, [(b.layernorm([self._ln(y1, sm_ln)])[0], skip)])[0]
        mods = list(self.expert) + list(self.domain_expert)
        depth = len(mods[0].layers())
        # first expert layer + the gates (which see emb.detach(), m3oe.py:151) in one launch
        groups = [dict(src=emb, W=m.layers()[0][0].weight, b=m.layers()[0][0].bias) for m in mods]
        groups += [dict(src=emb, W=self.gate[d][0].weight, b=self.gate[d][0].bias, detach=True) for d in range(D)]
        outs = b.fc(groups)
        ys, gate_logits = outs[:len(mods)], outs[len(mods):]
        cur = b.layernorm([self._ln(y, m.layers()[0][1]) for y, m in zip(ys, mods)])
        for i in range(1, depth):
            ys = b.fc([dict(src=a, W=m.layers()[i][0].weight, b=m.layers()[i][0].bias) for a, m in zip(cur, mods)])
            cur = b.layernorm([self._ln(y, m.layers()[i][1]) for y, m in zip(ys, mods)])
        fea, dfea = cur[:self.expert_num], cur[self.expert_num:]
        pooled = b.pool([(gate_logits[d], fea) for d in range(D)])
        b.mix(dfea, pooled, self._weight_exp_d.deep_weights, self._weight_bal_d.deep_weights)
        ts = b.fc([dict(src=pooled[d], W=self.tower[d][0].weight, b=self.tower[d][0].bias) for d in range(D)])
        tl = b.layernorm([self._ln(ts[d], self.tower[d][1]) for d in range(D)])
        b.head([(tl[d], self.tower[d][3].weight, self.tower[d][3].bias) for d in range(D)], dom, sig_before_select=True)
