"""Progressive Layered Extraction (reference: scenario_wise_rec/models/multi_domain/ple.py:13-136).

state_dict keys: ``embedding.*``, ``cgc_layers.<l>.experts_specific.<i>.mlp.*``,
``cgc_layers.<l>.experts_shared.<j>.mlp.*``, ``cgc_layers.<l>.gates_specific.<d>.mlp.*``,
``cgc_layers.<l>.gate_shared.mlp.*`` (levels before the last), ``towers.<d>.mlp.*``.
Every CGC level lowers to: first layer of all experts + all gates (grouped by the activation they
read) -> remaining expert layers -> one pooling kernel for the D (+1) gates.
"""
from torch import nn

from ...basic.layers import MLP, EmbeddingLayer, bn_norm
from ... import _native as N
from ._base import MultiDomainModel


class CGC(nn.Module):
    """Customized Gate Control level (reference ple.py:67-136); a parameter container here."""

    def __init__(self, cur_level, n_level, domain_num, n_expert_specific, n_expert_shared, input_dims, expert_params):
        super().__init__()
        self.cur_level, self.n_level, self.domain_num = cur_level, n_level, domain_num
        self.n_expert_specific, self.n_expert_shared = n_expert_specific, n_expert_shared
        self.n_expert_all = n_expert_specific * domain_num + n_expert_shared
        input_dims = input_dims if cur_level == 1 else expert_params["dims"][-1]
        self.experts_specific = nn.ModuleList(
            MLP(input_dims, output_layer=False, **expert_params) for _ in range(domain_num * n_expert_specific))
        self.experts_shared = nn.ModuleList(
            MLP(input_dims, output_layer=False, **expert_params) for _ in range(n_expert_shared))
        self.gates_specific = nn.ModuleList(
            MLP(input_dims, **{"dims": [n_expert_specific + n_expert_shared], "activation": "softmax",
                               "output_layer": False}) for _ in range(domain_num))
        if cur_level < n_level:
            self.gate_shared = MLP(input_dims, **{"dims": [self.n_expert_all], "activation": "softmax",
                                                  "output_layer": False})

    def lower(self, b, xs):
        D, ns = self.domain_num, self.n_expert_specific
        experts = list(self.experts_specific) + list(self.experts_shared)
        srcs = [xs[i // ns] for i in range(D * ns)] + [xs[-1]] * self.n_expert_shared
        gates = list(self.gates_specific)
        gsrcs = [xs[d] for d in range(D)]
        if self.cur_level < self.n_level:
            gates.append(self.gate_shared)
            gsrcs.append(xs[-1])
        for m in experts:
            m.check_dropout()
        groups = []
        for m, a in zip(experts, srcs):
            lin, bn = m.hidden()[0]
            groups.append(dict(src=a, W=lin.weight, b=lin.bias, norm=bn_norm(bn), act=m.act_code()))
        for g, a in zip(gates, gsrcs):
            lin, bn = g.hidden()[0]
            groups.append(dict(src=a, W=lin.weight, b=lin.bias, norm=bn_norm(bn), act=N.ACT_NONE))
        outs = b.fc(groups)
        cur, gate_acts = outs[:len(experts)], outs[len(experts):]
        for lvl in range(1, len(experts[0].dims)):
            groups = []
            for m, a in zip(experts, cur):
                lin, bn = m.hidden()[lvl]
                groups.append(dict(src=a, W=lin.weight, b=lin.bias, norm=bn_norm(bn), act=m.act_code()))
            cur = b.fc(groups)
        spec, shared = cur[:D * ns], cur[D * ns:]
        pools = [(gate_acts[d], spec[d * ns:(d + 1) * ns] + shared) for d in range(D)]
        if self.cur_level < self.n_level:
            pools.append((gate_acts[D], spec + shared))
        return b.pool(pools)


class PLE(MultiDomainModel):
    def __init__(self, features, domain_num, n_level, n_expert_specific, n_expert_shared, expert_params, tower_params):
        super().__init__()
        self.features = features
        self.domain_num = domain_num
        self.n_level = n_level
        self.input_dims = sum(fea.embed_dim for fea in features)
        self.embedding = EmbeddingLayer(features)
        self.cgc_layers = nn.ModuleList(
            CGC(i + 1, n_level, domain_num, n_expert_specific, n_expert_shared, self.input_dims, expert_params)
            for i in range(n_level))
        self.towers = nn.ModuleList(
            MLP(expert_params["dims"][-1], output_layer=True, **tower_params) for _ in range(domain_num))

    def _lower(self, b, col_dtypes):
        from ...basic.layers import lower_mlps
        x = self.embedding.lower(b, self.features, col_dtypes)
        xs = [x] * (self.domain_num + 1)
        for cgc in self.cgc_layers:
            xs = cgc.lower(b, xs)
        tops = lower_mlps(b, list(self.towers), xs[:self.domain_num])
        heads = []
        for t, a in zip(self.towers, tops):
            lin = t.out_linear()
            heads.append((a, lin.weight, lin.bias))
        b.head(heads, self._dom_dtype(col_dtypes), sig_before_select=True)
