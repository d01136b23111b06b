"""STAR (reference: scenario_wise_rec/models/multi_domain/star.py:10-118).

state_dict keys: ``embedding.*``, ``dn_share_gamma/bias``, ``auxnet.mlp.*``, ``share_parm_w/b.<l>``,
``domain_specific_dn_gamma/bias.<d>``, ``domain_specific_w/b.<d>.<l>`` (weights stored [K, N]),
``domain_specific_bn.<d>.<l>.*``.
Device program: K1 gather -> column moments of the embedding (once; every domain normalises the WHOLE
batch, star.py:95-100) -> per layer ONE grouped launch over the D domains whose weight staging forms
``W_share (.) W_d`` and ``b_share + b_d`` inline -> head ``sigmoid(select_d(relu(bn(.))) + aux)``.
"""
import torch
from torch import nn
from torch.nn import init
from torch.nn.parameter import Parameter

from ... import _native as N
from ...basic.layers import MLP, EmbeddingLayer, bn_norm, lower_mlps
from ...program import Norm
from ._base import MultiDomainModel


class Star(MultiDomainModel):
    def __init__(self, features, num_domains, fcn_dims, aux_dims):
        super().__init__()
        self.features = features
        self.input_dim = sum(fea.embed_dim for fea in features)
        self.layer_num = len(fcn_dims) + 1
        self.fcn_dim = [self.input_dim] + list(fcn_dims) + [1]
        self.num_domains = num_domains
        self.aux_dims = aux_dims
        self.embedding = EmbeddingLayer(features)
        self.dn_share_gamma = Parameter(torch.ones(self.input_dim))
        self.dn_share_bias = Parameter(torch.zeros(self.input_dim))
        self.eps = 1e-6
        self.auxnet = MLP(self.input_dim, dims=self.aux_dims)
        self.share_parm_w = nn.ParameterList()
        self.share_parm_b = nn.ParameterList()
        for i in range(self.layer_num):
            self.share_parm_w.append(Parameter(torch.empty((self.fcn_dim[i], self.fcn_dim[i + 1]))))
            self.share_parm_b.append(Parameter(torch.empty(self.fcn_dim[i + 1])))
        self.domain_specific_dn_gamma = nn.ParameterList()
        self.domain_specific_dn_bias = nn.ParameterList()
        self.domain_specific_w = nn.ParameterList()
        self.domain_specific_b = nn.ParameterList()
        self.domain_specific_bn = nn.ModuleList()
        for _d in range(num_domains):
            self.domain_specific_dn_gamma.append(Parameter(torch.ones(self.input_dim)))
            self.domain_specific_dn_bias.append(Parameter(torch.zeros(self.input_dim)))
            lay_w, lay_b, lay_bn = nn.ParameterList(), nn.ParameterList(), nn.ModuleList()
            for i in range(self.layer_num):
                lay_w.append(Parameter(torch.empty((self.fcn_dim[i], self.fcn_dim[i + 1]))))
                lay_b.append(Parameter(torch.empty(self.fcn_dim[i + 1])))
                lay_bn.append(nn.BatchNorm1d(self.fcn_dim[i + 1]))
            self.domain_specific_w.append(lay_w)
            self.domain_specific_b.append(lay_b)
            self.domain_specific_bn.append(lay_bn)
        self.reset_parameters()

    def reset_parameters(self):
        # same initialisers, same order as star.py:69-76 (same RNG stream for a seed)
        with torch.no_grad():
            for i in range(len(self.share_parm_w)):
                init.kaiming_uniform_(self.share_parm_w[i])
                init.uniform_(self.share_parm_b[i], 0, 1)
            for d in range(len(self.domain_specific_w)):
                for i in range(len(self.domain_specific_w[d])):
                    init.kaiming_uniform_(self.domain_specific_w[d][i])
                    init.uniform_(self.domain_specific_b[d][i], 0, 1)

    def _lower(self, b, col_dtypes):
        D = self.num_domains
        x = self.embedding.lower(b, self.features, col_dtypes)
        aux_h = lower_mlps(b, [self.auxnet], [x])[0]
        lin = self.auxnet.out_linear()
        aux = b.fc([dict(src=aux_h, W=lin.weight, b=lin.bias)])[0]
        cur = b.colstats(x, [Norm(gamma=self.dn_share_gamma, gamma2=self.domain_specific_dn_gamma[d],
                                  beta=self.dn_share_bias, beta2=self.domain_specific_dn_bias[d],
                                  eps=self.eps, always_batch=True) for d in range(D)])
        for l in range(self.layer_num):
            cur = b.fc([dict(src=cur[d], W=self.share_parm_w[l], W2=self.domain_specific_w[d][l],
                             b=self.share_parm_b[l], b2=self.domain_specific_b[d][l], layout=N.W_KN,
                             norm=bn_norm(self.domain_specific_bn[d][l]), act=N.ACT_RELU) for d in range(D)])
        b.head([(a, None, None) for a in cur], self._dom_dtype(col_dtypes), sig_before_select=False, add=aux)
