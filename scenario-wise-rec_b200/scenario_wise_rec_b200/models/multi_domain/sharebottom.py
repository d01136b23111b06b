"""Shared-Bottom (reference: scenario_wise_rec/models/multi_domain/sharebottom.py:6-50).

state_dict keys: ``embedding.embed_dict.<name>.weight``, ``bottom_mlp.mlp.*``, ``towers.<d>.mlp.*``.
Device program: K1 gather -> bottom layers -> all D towers level by level (one grouped launch per
level, every tower on every row so the BatchNorm statistics cover the full batch as in the
reference) -> fused Linear(.,1) + sigmoid + domain mask-select.
"""
from torch import nn

from ...basic.layers import MLP, EmbeddingLayer, lower_mlps
from ._base import MultiDomainModel


class SharedBottom(MultiDomainModel):
    def __init__(self, features, domain_num, bottom_params, tower_params):
        super().__init__()
        self.features = features
        self.embedding = EmbeddingLayer(features)
        self.bottom_dims = sum(fea.embed_dim for fea in features)
        self.domain_num = domain_num
        self.bottom_mlp = MLP(self.bottom_dims, **{**bottom_params, **{"output_layer": False}})
        self.towers = nn.ModuleList(MLP(bottom_params["dims"][-1], **tower_params) for _ in range(domain_num))

    def _lower(self, b, col_dtypes):
        x = self.embedding.lower(b, self.features, col_dtypes)
        h = lower_mlps(b, [self.bottom_mlp], [x])[0]
        tops = lower_mlps(b, list(self.towers), [h] * self.domain_num)
        heads = []
        for t, a in zip(self.towers, tops):
            lin = t.out_linear()
            if lin is None:
                raise NotImplementedError("towers need output_layer=True (the reference default)")
            heads.append((a, lin.weight, lin.bias))
        b.head(heads, self._dom_dtype(col_dtypes), sig_before_select=True)
