"""PPNet (reference: scenario_wise_rec/models/multi_domain/ppnet.py:8-67).

state_dict keys: ``id_embedding.*``, ``agn_embedding.*``, ``domain_tower.<d>.gate_layers.<l>.network.{0,2}.*``,
``domain_tower.<d>.mlp_layers.<l>.mlp.{0,1}.*``, ``domain_tower.<d>.final_layer.*``.
``gate_input = cat(id_x, agn_x.detach())`` feeds both the gate nets and the towers (ppnet.py:23,54), so the
agnostic tables never receive a gradient (their ``.grad`` stays ``None``, as in the reference).
Device program: two K1 gathers into one buffer -> every GateNU hidden layer of every domain and level in ONE
grouped launch (they all read gate_input) -> one launch for the GateNU output layers -> per level: grouped
tower layer (D domains) + fused ``relu(bn(.)) * 2*sigmoid(.)`` product -> head.
"""
from torch import nn

from ... import _native as N
from ...basic.layers import MLP, EmbeddingLayer, GateNU, bn_norm
from ._base import MultiDomainModel


class PPTowerBlock(nn.Module):
    def __init__(self, input_dim, fcn_dims):
        super().__init__()
        self.input_dim = input_dim
        self.dims = [input_dim] + list(fcn_dims)
        self.gate_layers = nn.ModuleList()
        self.mlp_layers = nn.ModuleList()
        for i in range(len(self.dims) - 1):
            self.mlp_layers.append(MLP(input_dim=self.dims[i], dims=[self.dims[i + 1]], output_layer=False))
            self.gate_layers.append(GateNU(self.dims[0], self.dims[i + 1]))
        self.final_layer = nn.Linear(self.dims[-1], 1)


class PPNet(MultiDomainModel):
    def __init__(self, id_features, agn_features, domain_num, fcn_dims):
        super().__init__()
        self.id_features = id_features
        self.agn_features = agn_features
        self.domain_num = domain_num
        self.id_embedding = EmbeddingLayer(id_features)
        self.agn_embedding = EmbeddingLayer(agn_features)
        self.id_dims = sum(fea.embed_dim for fea in id_features)
        self.agn_dims = sum(fea.embed_dim for fea in agn_features)
        self.input_dims = self.id_dims + self.agn_dims
        self.domain_tower = nn.ModuleList(PPTowerBlock(self.input_dims, fcn_dims) for _ in range(domain_num))

    def _feature_lists(self):
        return [self.id_features, self.agn_features]

    def _lower(self, b, col_dtypes):
        D = self.domain_num
        ids, idd = self.id_embedding.split_sharded(b, self.id_features, col_dtypes)
        ags, agd = self.agn_embedding.split_sharded(b, self.agn_features, col_dtypes)
        x = b.gather_parts([(ids, idd, True), (ags, agd, False)], col_dtypes)
        x.grad_cols = self.id_dims
        L = len(self.domain_tower[0].mlp_layers)
        gates = [(d, l, self.domain_tower[d].gate_layers[l]) for l in range(L) for d in range(D)]
        hid = b.fc([dict(src=x, W=g.network[0].weight, b=g.network[0].bias, act=N.ACT_RELU) for _d, _l, g in gates])
        gout = b.fc([dict(src=h, W=g.network[2].weight, b=g.network[2].bias, act=N.ACT_SIGMOID)
                     for h, (_d, _l, g) in zip(hid, gates)])
        gate_of = {(d, l): a for a, (d, l, _g) in zip(gout, gates)}
        cur = [x] * D
        for l in range(L):
            groups = []
            for d in range(D):
                m = self.domain_tower[d].mlp_layers[l]
                m.check_dropout()
                lin, bn = m.hidden()[0]
                groups.append(dict(src=cur[d], W=lin.weight, b=lin.bias, norm=bn_norm(bn), act=m.act_code()))
            hs = b.fc(groups)
            gemma = self.domain_tower[0].gate_layers[l].gemma
            cur = b.ew(N.EW_MUL, [(hs[d], gate_of[(d, l)]) for d in range(D)], scale=gemma)
        b.head([(cur[d], self.domain_tower[d].final_layer.weight, self.domain_tower[d].final_layer.bias) for d in range(D)],
               self._dom_dtype(col_dtypes), sig_before_select=True)
