"""fp32 torch-CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Every function is a *functional* restatement: it takes a reference-keyed ``state``
dict (``name -> tensor``, the same keys as the reference module's ``state_dict()``)
plus the batch dict ``x`` and returns the model output.  Tensors in ``state`` that
have ``requires_grad=True`` receive gradients through plain torch autograd, which
gives the gradient oracle.  BatchNorm running-stat updates (train mode) are
collected into ``bn_out`` so buffer parity can be checked too.

Feature lists are plain tuples ``(name, kind, vocab, dim)`` with kind in
{"sparse", "dense"} (dense: vocab=0, dim=1) -- the oracle does not depend on the
product package.

All ``file:line`` citations are relative to ``/root/reference/scenario_wise_rec``.
Pinned against the reference itself by ``oracle/make_golden.py`` ->
``tests/golden/*.npz`` and ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import torch

BN_EPS = 1e-5        # torch.nn.BatchNorm1d default used at basic/layers.py:255
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------- #
# building blocks
# --------------------------------------------------------------------------- #
def embedding_layer(x, feats, state, prefix):
    """basic/layers.py:64-114 with squeeze_dim=True.

    Sparse lookups ``W_f[x[f].long()]`` are concatenated in list order among the
    sparse features, THEN the dense scalars (``.float()``) in list order
    (layers.py:104) -- regardless of how sparse/dense interleave in ``feats``.
    """
    sparse, dense = [], []
    for name, kind, _vocab, _dim in feats:
        if kind == "sparse":
            w = state[f"{prefix}.embed_dict.{name}.weight"]
            idx = x[name].long()
            if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= w.shape[0]):
                raise IndexError("index out of range in self")   # torch CPU behaviour
            sparse.append(w[idx])
        else:
            dense.append(x[name].float().unsqueeze(1))
    if not sparse and not dense:
        raise ValueError("The input features can note be empty")   # layers.py:107
    parts = []
    if sparse:
        parts.append(torch.cat(sparse, dim=1))
    if dense:
        parts.append(torch.cat(dense, dim=1))
    return torch.cat(parts, dim=1) if len(parts) > 1 else parts[0]


def batch_norm(y, state, prefix, training, bn_out, eps=BN_EPS):
    """torch.nn.BatchNorm1d semantics (basic/layers.py:255, star.py:57, hamur.py:28).

    train: batch mean / biased variance normalise; running stats get the unbiased
    variance with momentum 0.1; num_batches_tracked += 1.  eval: running stats.
    """
    g, b = state[f"{prefix}.weight"], state[f"{prefix}.bias"]
    rm, rv = state[f"{prefix}.running_mean"], state[f"{prefix}.running_var"]
    if training:
        n = y.shape[0]
        if n <= 1:
            raise ValueError("Expected more than 1 value per channel when training")
        mean = y.mean(dim=0)
        var = ((y - mean) ** 2).mean(dim=0)
        if bn_out is not None:
            # a shared BN called several times per step (HAMUR hyper_net) keeps updating
            prev_rm = bn_out.get(f"{prefix}.running_mean", rm)
            prev_rv = bn_out.get(f"{prefix}.running_var", rv)
            prev_n = bn_out.get(f"{prefix}.num_batches_tracked", state[f"{prefix}.num_batches_tracked"])
            bn_out[f"{prefix}.running_mean"] = ((1 - BN_MOMENTUM) * prev_rm + BN_MOMENTUM * mean).detach()
            bn_out[f"{prefix}.running_var"] = ((1 - BN_MOMENTUM) * prev_rv + BN_MOMENTUM * var * n / (n - 1)).detach()
            bn_out[f"{prefix}.num_batches_tracked"] = prev_n + 1
    else:
        mean, var = rm, rv
    return (y - mean) / torch.sqrt(var + eps) * g + b


def linear(h, state, prefix):
    return h @ state[f"{prefix}.weight"].t() + state[f"{prefix}.bias"]


def _act(h, name):
    """basic/activation.py:38-49 (dice / prelu are out of scope)."""
    name = name.lower()
    if name == "relu":
        return torch.relu(h)
    if name == "sigmoid":
        return torch.sigmoid(h)
    if name == "softmax":
        return torch.softmax(h, dim=1)
    if name == "leakyrelu":
        return torch.nn.functional.leaky_relu(h, 0.1)
    raise NotImplementedError(name)


def mlp(h, state, prefix, dims, output_layer=True, activation="relu", training=True, bn_out=None):
    """basic/layers.py:231-264: [Linear -> BatchNorm1d -> act -> Dropout(0)] per dim (+ Linear(.,1))."""
    i = 0
    for _ in dims or []:
        h = linear(h, state, f"{prefix}.mlp.{i}")
        h = batch_norm(h, state, f"{prefix}.mlp.{i + 1}", training, bn_out)
        h = _act(h, activation)
        i += 4
    if output_layer:
        h = linear(h, state, f"{prefix}.mlp.{i}")
    return h


def gate_nu(h, state, prefix, gemma=2.0):
    """basic/layers.py:307-320: Linear -> ReLU -> Linear -> Sigmoid, times gemma."""
    h = torch.relu(linear(h, state, f"{prefix}.network.0"))
    return torch.sigmoid(linear(h, state, f"{prefix}.network.2")) * gemma


def mask_select(outs, domain_id):
    """The mask-select idiom (base_example.py:61-77; every model): rows whose id is
    outside [0, D) yield 0."""
    final = torch.zeros_like(outs[0])
    for d, o in enumerate(outs):
        m = (domain_id == d)
        final = torch.where(m.unsqueeze(1) if o.dim() == 2 else m, o, final)
    return final


# --------------------------------------------------------------------------- #
# models
# --------------------------------------------------------------------------- #
def shared_bottom(x, feats, state, domain_num, bottom_dims, tower_dims, training=True, bn_out=None,
                  bottom_act="relu", tower_act="relu"):
    """models/multi_domain/sharebottom.py:28-50."""
    dom = x["domain_indicator"]
    h = embedding_layer(x, feats, state, "embedding")
    h = mlp(h, state, "bottom_mlp", bottom_dims, False, bottom_act, training, bn_out)
    ys = [torch.sigmoid(mlp(h, state, f"towers.{d}", tower_dims, True, tower_act, training, bn_out))
          for d in range(domain_num)]
    return mask_select(ys, dom).squeeze(1)


def mmoe(x, feats, state, domain_num, n_expert, expert_dims, tower_dims, training=True, bn_out=None):
    """models/multi_domain/mmoe.py:33-56."""
    dom = x["domain_indicator"]
    e = embedding_layer(x, feats, state, "embedding")
    experts = torch.stack([mlp(e, state, f"experts.{i}", expert_dims, False, "relu", training, bn_out)
                           for i in range(n_expert)], dim=1)                       # [B, nE, H]
    ys = []
    for d in range(domain_num):
        g = mlp(e, state, f"gates.{d}", [n_expert], False, "softmax", training, bn_out)  # mmoe.py:26-30
        pooled = (g.unsqueeze(-1) * experts).sum(dim=1)                               # mmoe.py:48-49
        ys.append(torch.sigmoid(mlp(pooled, state, f"towers.{d}", tower_dims, True, "relu", training, bn_out)))
    return mask_select(ys, dom).squeeze(1)


def ple(x, feats, state, domain_num, n_level, n_spec, n_shared, expert_dims, tower_dims,
        training=True, bn_out=None):
    """models/multi_domain/ple.py:41-64 (PLE) and :107-136 (CGC)."""
    dom = x["domain_indicator"]
    e = embedding_layer(x, feats, state, "embedding")
    xs = [e] * (domain_num + 1)
    for lvl in range(n_level):
        p = f"cgc_layers.{lvl}"
        spec = []
        for d in range(domain_num):
            for j in range(n_spec):
                spec.append(mlp(xs[d], state, f"{p}.experts_specific.{d * n_spec + j}", expert_dims, False,
                                "relu", training, bn_out))
        shared = [mlp(xs[-1], state, f"{p}.experts_shared.{j}", expert_dims, False, "relu", training, bn_out)
                  for j in range(n_shared)]
        outs = []
        for d in range(domain_num):
            g = mlp(xs[d], state, f"{p}.gates_specific.{d}", [n_spec + n_shared], False, "softmax", training, bn_out)
            cur = torch.stack(spec[d * n_spec:(d + 1) * n_spec] + shared, dim=1)
            outs.append((g.unsqueeze(-1) * cur).sum(dim=1))
        if lvl + 1 < n_level:                                                          # ple.py:127-134
            g = mlp(xs[-1], state, f"{p}.gate_shared", [n_spec * domain_num + n_shared], False, "softmax",
                    training, bn_out)
            cur = torch.stack(spec + shared, dim=1)
            outs.append((g.unsqueeze(-1) * cur).sum(dim=1))
        xs = outs
    ys = [torch.sigmoid(mlp(xs[d], state, f"towers.{d}", tower_dims, True, "relu", training, bn_out))
          for d in range(domain_num)]
    return mask_select(ys, dom).squeeze(1)


def star(x, feats, state, num_domains, fcn_dims, aux_dims, training=True, bn_out=None):
    """models/multi_domain/star.py:78-118.  Weights are stored [K, N]; the partitioned
    norm uses the WHOLE batch (biased var, eps 1e-6) for every domain (star.py:92-100)."""
    dom = x["domain_indicator"]
    emb = embedding_layer(x, feats, state, "embedding")
    aux = mlp(emb, state, "auxnet", aux_dims, True, "relu", training, bn_out)
    n_layer = len(fcn_dims) + 1
    outs = []
    for d in range(num_domains):
        mean = emb.mean(dim=0)
        var = ((emb - mean) ** 2).mean(dim=0)
        h = (emb - mean) / torch.sqrt(var + 1e-6)
        h = ((state["dn_share_gamma"] * state[f"domain_specific_dn_gamma.{d}"]) * h
             + state["dn_share_bias"] + state[f"domain_specific_dn_bias.{d}"])
        for l in range(n_layer):
            w = state[f"share_parm_w.{l}"] * state[f"domain_specific_w.{d}.{l}"]
            b = state[f"share_parm_b.{l}"] + state[f"domain_specific_b.{d}.{l}"]
            h = h @ w + b
            h = batch_norm(h, state, f"domain_specific_bn.{d}.{l}", training, bn_out)
            h = torch.relu(h)
        outs.append(h)
    final = mask_select(outs, dom)
    return torch.sigmoid(final + aux).squeeze(1)


def ppnet(x, id_feats, agn_feats, state, domain_num, fcn_dims, training=True, bn_out=None):
    """models/multi_domain/ppnet.py:47-67 and PPTowerBlock :21-29.  ``hidden`` starts
    from gate_input = cat(id_x, agn_x.detach()), so agn tables never get a gradient."""
    dom = x["domain_indicator"]
    id_x = embedding_layer(x, id_feats, state, "id_embedding")
    agn_x = embedding_layer(x, agn_feats, state, "agn_embedding")
    gate_in = torch.cat((id_x, agn_x.detach()), dim=1)
    outs = []
    for d in range(domain_num):
        p = f"domain_tower.{d}"
        h = gate_in
        for l, n in enumerate(fcn_dims):
            g = gate_nu(gate_in, state, f"{p}.gate_layers.{l}")
            h = mlp(h, state, f"{p}.mlp_layers.{l}", [n], False, "relu", training, bn_out)
            h = h * g
        outs.append(torch.sigmoid(linear(h, state, f"{p}.final_layer")))
    return mask_select(outs, dom).squeeze(1)


def epnet(x, sce_feats, agn_feats, state, training=True, bn_out=None):
    """models/multi_domain/epnet.py:25-32.  ``self.mlp = MLP(agn_dims, fcn_dims)``
    passes fcn_dims as ``output_layer`` so the "mlp" is a single Linear(agn_dims, 1)
    (epnet.py:21); there is no domain mask."""
    sce_x = embedding_layer(x, sce_feats, state, "sce_embedding")
    agn_x = embedding_layer(x, agn_feats, state, "agn_embedding")
    gate = gate_nu(torch.cat((sce_x, agn_x.detach()), dim=1), state, "gatenu")
    out = linear(agn_x * gate, state, "mlp.mlp.0")
    return torch.sigmoid(out).squeeze()


def _layer_norm(h, state, prefix, eps=1e-5):
    mean = h.mean(dim=1, keepdim=True)
    var = ((h - mean) ** 2).mean(dim=1, keepdim=True)
    return (h - mean) / torch.sqrt(var + eps) * state[f"{prefix}.weight"] + state[f"{prefix}.bias"]


def _mlp_n(h, state, prefix, n_layers):
    """m3oe.py:45-68: [Linear -> LayerNorm -> ReLU] per layer."""
    for i in range(n_layers):
        h = linear(h, state, f"{prefix}.domain_specific.{3 * i}")
        h = _layer_norm(h, state, f"{prefix}.domain_specific.{3 * i + 1}")
        h = torch.relu(h)
    return h


def m3oe(x, feats, state, domain_num, fcn_dims, expert_num, training=True, bn_out=None):
    """models/multi_domain/m3oe.py:135-198 (softmax_type 3: sigmoid scalar weights)."""
    dom = x["domain_indicator"]
    emb_in = embedding_layer(x, feats, state, "embedding")
    n_rest = len(fcn_dims) - 3          # fcn_dim[3:] has len(fcn_dims)-2 entries -> that many minus one Linear layers
    skip = _mlp_n(emb_in, state, "skip_conn", 1)
    outs = []
    for d in range(domain_num):
        w = state[f"slot_weight.{d}"] * state["shared_weight"]
        outs.append(emb_in @ w + state[f"slot_bias.{d}"] + state["shared_bias"])
    emb = mask_select(outs, dom)
    emb = _mlp_n(emb, state, "star_mlp", 1) + skip
    gates = [torch.softmax(linear(emb.detach(), state, f"gate.{d}.0"), dim=1) for d in range(domain_num)]
    fea = torch.stack([_mlp_n(emb, state, f"expert.{i}", n_rest) for i in range(expert_num)], dim=1)
    dfea = [_mlp_n(emb, state, f"domain_expert.{d}", n_rest) for d in range(domain_num)]
    w_bal = torch.sigmoid(state["_weight_bal_d.deep_weights"])
    w_exp = torch.sigmoid(state["_weight_exp_d.deep_weights"])
    ys = []
    for d in range(domain_num):
        mix = w_bal * dfea[d]
        for j in range(domain_num):
            if j != d:
                mix = mix + (1 - w_bal) / (domain_num - 1) * dfea[j]
        fused = (gates[d].unsqueeze(-1) * fea).sum(dim=1) + w_exp * mix
        p = f"tower.{d}"
        t = linear(fused, state, f"{p}.0")
        t = torch.relu(_layer_norm(t, state, f"{p}.1"))
        ys.append(torch.sigmoid(linear(t, state, f"{p}.3").squeeze(1)))
    return mask_select(ys, dom)


def _hyper_net(emb, state, n_hyper, training, bn_out):
    """hamur.py:76-86 / :287-297: [Linear -> BN -> ReLU -> Dropout(0)] per hyper dim (+ k*k)."""
    h = emb
    for i in range(n_hyper):
        h = linear(h, state, f"hyper_net.{4 * i}")
        h = batch_norm(h, state, f"hyper_net.{4 * i + 1}", training, bn_out)
        h = torch.relu(h)
    return h


def _adapter(h, hyper, state, cell, gamma, bias, eps=1e-5):
    """One HAMUR adapter cell (hamur.py:174-197 / :343-367), evaluated exactly as the
    reference does (materialised per-sample weights); domain norm uses the UNBIASED
    batch variance (``tmp_out.var(dim=0)``)."""
    u0, v0 = state[f"u.{2 * cell}"], state[f"v.{2 * cell}"]
    u1, v1 = state[f"u.{2 * cell + 1}"], state[f"v.{2 * cell + 1}"]
    w1 = torch.einsum("mi,bij,jn->bmn", u0, hyper, v0)
    t = torch.einsum("bf,bfj->bj", h, w1) + state[f"b_list.{2 * cell}"]
    t = torch.sigmoid(t)
    w2 = torch.einsum("mi,bij,jn->bmn", u1, hyper, v1)
    t = torch.einsum("bf,bfj->bj", t, w2) + state[f"b_list.{2 * cell + 1}"]
    mean = t.mean(dim=0)
    var = t.var(dim=0)
    return state[gamma] * ((t - mean) / torch.sqrt(var + eps)) + state[bias] + h


def hamur(x, feats, state, domain_num, n_backbone, n_hyper, k, training=True, bn_out=None):
    """HamurSmall (hamur.py:308-378, n_backbone=2, one adapter cell after layer 2) and
    HamurLarge (hamur.py:101-244, n_backbone=7, cells after layers 6 and 7).  The shared
    hyper_net is evaluated once per domain, so its BN buffers update D times per step."""
    dom = x["domain_indicator"]
    emb = embedding_layer(x, feats, state, "embedding")
    outs = []
    for d in range(domain_num):
        hyper = _hyper_net(emb, state, n_hyper, training, bn_out).reshape(emb.shape[0], k, k)
        h = emb
        p = f"layer_list.{d}"
        for l in range(n_backbone):
            h = linear(h, state, f"{p}.{3 * l}")
            h = batch_norm(h, state, f"{p}.{3 * l + 1}", training, bn_out)
            h = torch.relu(h)
            if n_backbone == 2 and l == 1:
                h = _adapter(h, hyper, state, 0, "gamma1", "bias1")
            if n_backbone == 7 and l == 5:
                h = _adapter(h, hyper, state, 0, "gamma1", "bias1")
            if n_backbone == 7 and l == 6:
                h = _adapter(h, hyper, state, 1, "gamma2", "bias2")
        outs.append(torch.sigmoid(linear(h, state, f"{p}.{3 * n_backbone}")))
    return mask_select(outs, dom).squeeze(1)


# --------------------------------------------------------------------------- #
# dispatch used by tests / bench
# --------------------------------------------------------------------------- #
def forward(model_name, x, state, cfg, training=True, bn_out=None):
    """Run the restated forward of ``model_name`` with config dict ``cfg`` (the same
    keys make_golden.py stores next to each golden vector)."""
    f = cfg.get("features")
    if model_name == "SharedBottom":
        return shared_bottom(x, f, state, cfg["domain_num"], cfg["bottom_dims"], cfg["tower_dims"], training, bn_out)
    if model_name == "MMOE":
        return mmoe(x, f, state, cfg["domain_num"], cfg["n_expert"], cfg["expert_dims"], cfg["tower_dims"],
                    training, bn_out)
    if model_name == "PLE":
        return ple(x, f, state, cfg["domain_num"], cfg["n_level"], cfg["n_expert_specific"],
                   cfg["n_expert_shared"], cfg["expert_dims"], cfg["tower_dims"], training, bn_out)
    if model_name == "Star":
        return star(x, f, state, cfg["domain_num"], cfg["fcn_dims"], cfg["aux_dims"], training, bn_out)
    if model_name == "PPNet":
        return ppnet(x, cfg["id_features"], cfg["agn_features"], state, cfg["domain_num"], cfg["fcn_dims"],
                     training, bn_out)
    if model_name == "EPNet":
        return epnet(x, cfg["sce_features"], cfg["agn_features"], state, training, bn_out)
    if model_name == "M3oE":
        return m3oe(x, f, state, cfg["domain_num"], cfg["fcn_dims"], cfg["expert_num"], training, bn_out)
    if model_name in ("HamurSmall", "HamurLarge"):
        return hamur(x, f, state, cfg["domain_num"], 2 if model_name == "HamurSmall" else 7,
                     len(cfg["hyper_dims"]) + 1, cfg["k"], training, bn_out)
    raise KeyError(model_name)


def bce_loss(y_pred, y):
    """torch.nn.BCELoss (trainers/ctr_trainer.py:56,70): mean of -[y log p + (1-y) log(1-p)],
    logs clamped at -100 like torch."""
    lp = torch.clamp(torch.log(y_pred), min=-100.0)
    l1p = torch.clamp(torch.log(1 - y_pred), min=-100.0)
    return -(y * lp + (1 - y) * l1p).mean()
