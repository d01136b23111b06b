"""Building blocks with the reference's constructor signatures and state_dict keys
(reference: scenario_wise_rec/basic/layers.py).

The modules are parameter containers plus *lowering helpers*: a model (or the module's
own ``forward``) describes its computation to a :class:`ProgramBuilder` through
``lower(...)``; the arithmetic then runs in the CUDA kernels of libswr_b200.so.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import nn

from .. import _native as N
from ..fused import FusedModule
from ..program import Act, Norm, ProgramBuilder
from .features import DenseFeature, SparseFeature

_ACT_CODES = {"relu": N.ACT_RELU, "sigmoid": N.ACT_SIGMOID, "leakyrelu": N.ACT_LEAKY}


def activation_layer(act_name):
    """reference basic/activation.py:28-54 (dice / prelu are outside the accelerated path)."""
    if isinstance(act_name, str):
        name = act_name.lower()
        if name == "sigmoid":
            return nn.Sigmoid()
        if name == "relu":
            return nn.ReLU(inplace=True)
        if name == "softmax":
            return nn.Softmax(dim=1)
        if name == "leakyrelu":
            return nn.LeakyReLU(0.1)
        raise NotImplementedError(f"activation {act_name!r} is not supported by the fused kernels")
    if isinstance(act_name, type) and issubclass(act_name, nn.Module):
        return act_name()
    raise NotImplementedError


def bn_norm(bn: nn.BatchNorm1d) -> Norm:
    if bn.momentum is None or abs(bn.momentum - 0.1) > 1e-12 or not bn.track_running_stats or not bn.affine:
        raise NotImplementedError("only default BatchNorm1d (affine, momentum 0.1, running stats) is supported")
    return Norm(gamma=bn.weight, beta=bn.bias, rmean=bn.running_mean, rvar=bn.running_var,
                nbt=bn.num_batches_tracked, eps=bn.eps)


class EmbeddingLayer(FusedModule):
    """reference basic/layers.py:27-114.  ``embed_dict[name]`` holds the ``nn.Embedding`` of every
    sparse feature (shared through the feature object's cache); ``forward`` is the fused gather K1:
    sparse rows in list order followed by the dense scalars."""

    def __init__(self, features):
        super().__init__()
        self.features = features
        self.embed_dict = nn.ModuleDict()
        self.n_dense = 0
        for fea in features:
            if fea.name in self.embed_dict:
                continue
            if isinstance(fea, SparseFeature) and fea.shared_with is None:
                self.embed_dict[fea.name] = fea.get_embedding_layer()
            elif isinstance(fea, DenseFeature):
                self.n_dense += 1
        self._active: Optional[list] = None
        self._squeeze = True

    # ---- lowering ---------------------------------------------------------------------------
    def split(self, features):
        sparse, dense = [], []
        for fea in features:
            if isinstance(fea, SparseFeature):
                tab = self.embed_dict[fea.name if fea.shared_with is None else fea.shared_with].weight
                sparse.append((fea.name, tab))
            elif isinstance(fea, DenseFeature):
                dense.append(fea.name)
            else:
                raise NotImplementedError(f"feature type {type(fea).__name__} is outside the accelerated path")
        return sparse, dense

    def split_sharded(self, b: ProgramBuilder, features, col_dtypes):
        """``split`` with row-sharded tables replaced by their virtual table / virtual index column (parallel.py)."""
        sparse, dense = self.split(features)
        feas = [f for f in features if isinstance(f, SparseFeature)]
        out = []
        for (name, tab), fea in zip(sparse, feas):
            if getattr(fea, "shard", None) is not None and b.p2p:
                out.append((name, b.peer_table(fea, tab)))      # K1 / K2 address the owners' shards directly (NVLink)
            elif getattr(fea, "shard", None) is not None:
                f = b.virtual_field(fea, tab)
                col_dtypes[f.vcol] = torch.int64
                out.append((f.vcol, f.virt))
            else:
                out.append((name, tab))
        return out, dense

    def lower(self, b: ProgramBuilder, features, col_dtypes) -> Act:
        sparse, dense = self.split_sharded(b, features, col_dtypes)
        return b.gather(sparse, dense, col_dtypes)

    @staticmethod
    def columns(features) -> List[str]:
        return [f.name for f in features]

    # ---- standalone use ---------------------------------------------------------------------
    def _columns(self):
        return self.columns(self._active)

    def _lower(self, b, col_dtypes):
        b.output(self.lower(b, self._active, col_dtypes))

    def forward(self, x, features, squeeze_dim=False):
        sparse, dense = self.split(features)
        if not sparse and not dense:
            raise ValueError("The input features can note be empty")
        if not squeeze_dim:
            if not sparse:
                raise ValueError("If keep the original shape:[batch_size, num_features, embed_dim], expected "
                                 "SparseFeatures in feature list, got %s" % (features,))
            dims = {int(t.shape[1]) for _, t in sparse}
            if len(dims) != 1:
                raise RuntimeError("squeeze_dim=False needs one embed_dim for all sparse features")
            active = [f for f in features if isinstance(f, SparseFeature)]
        else:
            active = list(features)
        if self._active is None or [f.name for f in self._active] != [f.name for f in active]:
            self._active, self._programs = active, {}
        out = self._run(x)
        if not squeeze_dim:
            return out.view(out.shape[0], len(sparse), -1)
        return out


class MLP(nn.Module):
    """reference basic/layers.py:231-264: ``[Linear -> BatchNorm1d -> act -> Dropout] * len(dims)``
    (+ ``Linear(., 1)`` if ``output_layer``), stored as ``self.mlp`` with the same indices."""

    def __init__(self, input_dim, output_layer=True, dims=None, dropout=0, activation="relu"):
        super().__init__()
        dims = [] if dims is None else list(dims)
        layers = []
        self.input_dim, self.dims, self.output_layer = input_dim, dims, bool(output_layer)
        self.activation, self.dropout = activation, dropout
        for d in dims:
            layers += [nn.Linear(input_dim, d), nn.BatchNorm1d(d), activation_layer(activation), nn.Dropout(p=dropout)]
            input_dim = d
        if output_layer:
            layers.append(nn.Linear(input_dim, 1))
        self.mlp = nn.Sequential(*layers)

    def hidden(self):
        """[(Linear, BatchNorm1d)] of the hidden layers."""
        return [(self.mlp[4 * i], self.mlp[4 * i + 1]) for i in range(len(self.dims))]

    def out_linear(self) -> Optional[nn.Linear]:
        return self.mlp[4 * len(self.dims)] if self.output_layer else None

    def act_code(self) -> int:
        name = self.activation.lower() if isinstance(self.activation, str) else None
        if name not in _ACT_CODES:
            raise NotImplementedError(f"MLP activation {self.activation!r} is not supported by the fused kernels")
        return _ACT_CODES[name]

    def check_dropout(self):
        if self.dropout and self.training:
            raise NotImplementedError("dropout > 0 in train mode cannot reproduce the reference RNG stream "
                                      "inside the fused kernels; every hot-path config uses dropout=0")

    def forward(self, x):
        raise RuntimeError("MLP is lowered by its owning model (fused device program); it has no standalone "
                           "eager forward and there is no CPU fallback")


def lower_mlps(b: ProgramBuilder, mlps: Sequence[MLP], srcs: Sequence[Act], softmax_gate: bool = False) -> List[Act]:
    """Lower several same-depth MLPs level by level, one grouped launch per level.  Returns the lazy
    activation after the last hidden layer of each MLP (output Linear layers are left to the caller)."""
    depth = {len(m.dims) for m in mlps}
    if len(depth) != 1:
        raise ValueError("lower_mlps: MLPs must have the same depth")
    cur = list(srcs)
    for m in mlps:
        m.check_dropout()
    for lvl in range(depth.pop()):
        groups = []
        for m, a in zip(mlps, cur):
            lin, bn = m.hidden()[lvl]
            act = N.ACT_NONE if (softmax_gate and m.activation == "softmax") else m.act_code()
            groups.append(dict(src=a, W=lin.weight, b=lin.bias, norm=bn_norm(bn), act=act))
        cur = b.fc(groups)
    return cur


class GateNU(nn.Module):
    """reference basic/layers.py:307-320: ``Linear -> ReLU -> Linear -> Sigmoid`` times ``gemma``."""

    def __init__(self, input_dim, output_dim, hidden_dim=None, gemma=2.0):
        super().__init__()
        hidden_dim = output_dim if hidden_dim is None else hidden_dim
        self.gemma = gemma
        self.network = nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.ReLU(),
                                     nn.Linear(hidden_dim, output_dim), nn.Sigmoid())

    def forward(self, inputs):
        raise RuntimeError("GateNU is lowered by its owning model (fused device program)")
