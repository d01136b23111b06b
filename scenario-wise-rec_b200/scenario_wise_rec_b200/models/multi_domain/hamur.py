"""HAMUR (reference: scenario_wise_rec/models/multi_domain/hamur.py:9-378).

state_dict keys: ``embedding.*``, ``layer_list.<d>.<i>.*`` (Linear / BatchNorm1d / ReLU triples + the final
Linear), ``u.<i>``, ``v.<i>``, ``b_list.<i>``, ``hyper_net.<i>.*``, ``gamma1/bias1`` (``gamma2/bias2`` for Large).

The reference materialises per-sample adapter weights ``W[b] = U H_b V`` ([B, m, 32] tensors, hamur.py:177,186).
Here the adapter is re-associated as ``((h U) H_b) V``: two small grouped FCs around a per-sample k x k
mat-vec (BMV kernel), so nothing of size [B, m, n] ever reaches HBM.  The shared hyper-network is evaluated
once per step instead of once per domain (its output is identical each time); its BatchNorm running
statistics are updated ``domain_num`` times to match the reference buffers.  Domain norm uses the unbiased
batch variance (``tmp_out.var(dim=0)``) in train AND eval mode, as in the reference.
"""
import torch
from torch import nn
from torch.nn.parameter import Parameter

from ... import _native as N
from ...basic.layers import EmbeddingLayer, bn_norm
from ...program import Norm
from ._base import MultiDomainModel


class _Hamur(MultiDomainModel):
    n_backbone = 2
    cells = {1: 0}          # backbone layer index after which adapter cell c is applied

    def __init__(self, features, domain_num, fcn_dims, hyper_dims, k):
        super().__init__()
        if len(fcn_dims) != self.n_backbone:
            raise ValueError(f"{type(self).__name__} needs {self.n_backbone} fcn_dims")
        self.features = features
        self.input_dim = sum(fea.embed_dim for fea in features)
        self.layer_num = len(fcn_dims) + 1
        self.fcn_dim = [self.input_dim] + list(fcn_dims)
        self.domain_num = domain_num
        self.embedding = EmbeddingLayer(features)
        self.layer_list = nn.ModuleList()
        for _d in range(domain_num):
            ds = nn.ModuleList()
            for i in range(self.n_backbone):
                ds.append(nn.Linear(self.fcn_dim[i], self.fcn_dim[i + 1]))
                ds.append(nn.BatchNorm1d(self.fcn_dim[i + 1]))
                ds.append(nn.ReLU())
            ds.append(nn.Linear(self.fcn_dim[self.n_backbone], 1))
            self.layer_list.append(ds)
        self.k = k
        self.u = nn.ParameterList()
        self.v = nn.ParameterList()
        widths = [self.fcn_dim[l + 1] for l in sorted(self.cells)]
        for m in widths:
            self.u.append(Parameter(torch.ones((m, k))))
            self.u.append(Parameter(torch.ones((32, k))))
        for m in widths:
            self.v.append(Parameter(torch.ones((k, 32))))
            self.v.append(Parameter(torch.ones((k, m))))
        hyper_dims += [k * k]           # the reference mutates the caller's list too (hamur.py:77,288)
        input_dim = self.input_dim
        layers = []
        for i_dim in hyper_dims:
            layers += [nn.Linear(input_dim, i_dim), nn.BatchNorm1d(i_dim), nn.ReLU(), nn.Dropout(p=0)]
            input_dim = i_dim
        self.hyper_net = nn.Sequential(*layers)
        self.b_list = nn.ParameterList()
        for m in widths:
            self.b_list.append(Parameter(torch.zeros((32))))
            self.b_list.append(Parameter(torch.zeros((m))))
        for c, m in enumerate(widths):
            setattr(self, f"gamma{c + 1}", nn.Parameter(torch.ones(m)))
            setattr(self, f"bias{c + 1}", nn.Parameter(torch.zeros(m)))
        self.eps = 1e-5

    def _adapter(self, b, hs, H, c):
        """One adapter cell for every domain (hamur.py:174-197 / :343-367)."""
        D = self.domain_num
        u0, u1, v0, v1 = self.u[2 * c], self.u[2 * c + 1], self.v[2 * c], self.v[2 * c + 1]
        b0, b1 = self.b_list[2 * c], self.b_list[2 * c + 1]
        p0 = b.fc([dict(src=hs[d], W=u0, layout=N.W_KN) for d in range(D)])
        q0 = b.bmv(p0, H, self.k)
        t1 = b.fc([dict(src=q0[d], W=v0, b=b0, layout=N.W_KN, act=N.ACT_SIGMOID) for d in range(D)])
        p1 = b.fc([dict(src=t1[d], W=u1, layout=N.W_KN) for d in range(D)])
        q1 = b.bmv(p1, H, self.k)
        gamma, bias = getattr(self, f"gamma{c + 1}"), getattr(self, f"bias{c + 1}")
        t2 = b.fc([dict(src=q1[d], W=v1, b=b1, layout=N.W_KN,
                        norm=Norm(gamma=gamma, beta=bias, eps=self.eps, always_batch=True, unbiased=True))
                   for d in range(D)])
        return b.ew(N.EW_ADD, [(t2[d], hs[d]) for d in range(D)])

    def _lower(self, b, col_dtypes):
        D = self.domain_num
        x = self.embedding.lower(b, self.features, col_dtypes)
        h = x
        for i in range(len(self.hyper_net) // 4):
            lin, bn = self.hyper_net[4 * i], self.hyper_net[4 * i + 1]
            norm = bn_norm(bn)
            norm.repeat = D             # hyper_net(domain_input) runs once per domain in the reference
            h = b.fc([dict(src=h, W=lin.weight, b=lin.bias, norm=norm, act=N.ACT_RELU)])[0]
        H = b.ew(N.EW_COPY, [(h, None)])[0]
        cur = [x] * D
        for l in range(self.n_backbone):
            cur = b.fc([dict(src=cur[d], W=self.layer_list[d][3 * l].weight, b=self.layer_list[d][3 * l].bias,
                             norm=bn_norm(self.layer_list[d][3 * l + 1]), act=N.ACT_RELU) for d in range(D)])
            if l in self.cells:
                cur = self._adapter(b, cur, H, self.cells[l])
        fin = [self.layer_list[d][3 * self.n_backbone] for d in range(D)]
        b.head([(cur[d], fin[d].weight, fin[d].bias) for d in range(D)], self._dom_dtype(col_dtypes), sig_before_select=True)


class HamurSmall(_Hamur):
    """2-layer backbone with one adapter cell (hamur.py:247-378)."""
    n_backbone = 2
    cells = {1: 0}


class HamurLarge(_Hamur):
    """7-layer backbone with adapter cells after layers 6 and 7 (hamur.py:9-244)."""
    n_backbone = 7
    cells = {5: 0, 6: 1}
