"""Multi-gate Mixture-of-Experts (reference: scenario_wise_rec/models/multi_domain/mmoe.py:6-56).

state_dict keys: ``embedding.*``, ``experts.<e>.mlp.*``, ``gates.<d>.mlp.{0,1}.*``, ``towers.<d>.mlp.*``.
Device program: K1 gather -> level 1 of all experts AND all gates in one grouped launch (they share
the input) -> remaining expert levels -> gate BatchNorm+softmax+pooling (one kernel for all domains)
-> towers -> head.
"""
from torch import nn

from ...basic.layers import MLP, EmbeddingLayer, bn_norm, lower_mlps
from ... import _native as N
from ._base import MultiDomainModel


class MMOE(MultiDomainModel):
    def __init__(self, features, domain_num, n_expert, expert_params, tower_params):
        super().__init__()
        self.features = features
        self.domain_num = domain_num
        self.n_expert = n_expert
        self.embedding = EmbeddingLayer(features)
        self.input_dims = sum(fea.embed_dim for fea in features)
        self.experts = nn.ModuleList(MLP(self.input_dims, output_layer=False, **expert_params) for _ in range(n_expert))
        self.gates = nn.ModuleList(
            MLP(self.input_dims, output_layer=False, **{"dims": [n_expert], "activation": "softmax"})
            for _ in range(domain_num))
        self.towers = nn.ModuleList(MLP(expert_params["dims"][-1], **tower_params) for _ in range(domain_num))

    def _lower(self, b, col_dtypes):
        x = self.embedding.lower(b, self.features, col_dtypes)
        experts, gates = list(self.experts), list(self.gates)
        for m in experts:
            m.check_dropout()
        # level 1: experts and gates read the same activation -> one launch
        groups = []
        for m in experts:
            lin, bn = m.hidden()[0]
            groups.append(dict(src=x, W=lin.weight, b=lin.bias, norm=bn_norm(bn), act=m.act_code()))
        for g in gates:
            lin, bn = g.hidden()[0]
            groups.append(dict(src=x, W=lin.weight, b=lin.bias, norm=bn_norm(bn), act=N.ACT_NONE))
        outs = b.fc(groups)
        cur, gate_acts = outs[:len(experts)], outs[len(experts):]
        for lvl in range(1, len(experts[0].dims)):
            groups = []
            for m, a in zip(experts, cur):
                lin, bn = m.hidden()[lvl]
                groups.append(dict(src=a, W=lin.weight, b=lin.bias, norm=bn_norm(bn), act=m.act_code()))
            cur = b.fc(groups)
        pooled = b.pool([(g, cur) for g in gate_acts])
        tops = lower_mlps(b, list(self.towers), pooled)
        heads = []
        for t, a in zip(self.towers, tops):
            lin = t.out_linear()
            if lin is None:
                raise NotImplementedError("towers need output_layer=True (the reference default)")
            heads.append((a, lin.weight, lin.bias))
        b.head(heads, self._dom_dtype(col_dtypes), sig_before_select=True)
