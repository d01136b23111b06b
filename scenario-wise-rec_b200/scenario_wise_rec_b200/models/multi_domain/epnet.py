"""EPNet (reference: scenario_wise_rec/models/multi_domain/epnet.py:6-32).

state_dict keys: ``sce_embedding.*``, ``agn_embedding.*``, ``gatenu.network.{0,2}.*``, ``mlp.mlp.0.*``.
Quirk kept: ``MLP(agn_dims, fcn_dims)`` passes ``fcn_dims`` as ``output_layer`` (epnet.py:21), so "mlp" is a
single ``Linear(agn_dims, 1)``; there is no domain mask.  ``gate_input = cat(sce_x, agn_x.detach())``: the gate
net back-propagates into the scenario tables only; the agnostic tables get their gradient through
``agn_x * gate``.
"""
from torch import nn

from ... import _native as N
from ...basic.layers import MLP, EmbeddingLayer, GateNU
from ._base import MultiDomainModel


class EPNet(MultiDomainModel):
    def __init__(self, sce_features, agn_features, fcn_dims):
        super().__init__()
        self.sce_features = sce_features
        self.agn_features = agn_features
        self.sce_embedding = EmbeddingLayer(sce_features)
        self.agn_embedding = EmbeddingLayer(agn_features)
        self.sce_dims = sum(fea.embed_dim for fea in sce_features)
        self.agn_dims = sum(fea.embed_dim for fea in agn_features)
        self.dims = self.sce_dims + self.agn_dims
        self.gatenu = GateNU(self.dims, self.agn_dims)
        self.mlp = MLP(self.agn_dims, fcn_dims)        # == Linear(agn_dims, 1), see module docstring
        self.sigmoid = nn.Sigmoid()

    def _feature_lists(self):
        return [self.sce_features, self.agn_features]

    def _columns(self):
        cols = super()._columns()
        return cols[:-1]            # EPNet never reads domain_indicator

    def _lower(self, b, col_dtypes):
        ss, sd = self.sce_embedding.split_sharded(b, self.sce_features, col_dtypes)
        ags, agd = self.agn_embedding.split_sharded(b, self.agn_features, col_dtypes)
        x = b.gather_parts([(ss, sd, True), (ags, agd, True)], col_dtypes)
        x.grad_cols = self.sce_dims                      # the gate net sees agn_x.detach()
        agn = b.subview(x, self.sce_dims, self.agn_dims)
        g = self.gatenu
        hid = b.fc([dict(src=x, W=g.network[0].weight, b=g.network[0].bias, act=N.ACT_RELU)])[0]
        gate = b.fc([dict(src=hid, W=g.network[2].weight, b=g.network[2].bias, act=N.ACT_SIGMOID)])[0]
        prod = b.ew(N.EW_MUL, [(agn, gate)], scale=g.gemma)[0]
        lin = self.mlp.out_linear()
        b.head([(prod, lin.weight, lin.bias)], None, sig_before_select=N.HEAD_NO_SELECT)
