"""Generate tests/golden/*.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (the GPU box has no /root/reference):

    python oracle/make_golden.py            # writes tests/golden/<case>.npz

For every hot-path model it builds the reference
``scenario_wise_rec.models.multi_domain.<Model>`` from ``/root/reference``, randomises
its state (so BatchNorm / embeddings are numerically non-trivial), runs one train-mode
forward + ``BCELoss`` + backward and one eval-mode forward on a seeded synthetic batch,
and stores inputs, the state before, output, loss, every parameter gradient (with the
``grad is None`` pattern), the BN buffers after the train forward and the eval output.
The reference has no golden vectors of its own; these files are
the pin for ``oracle/ref_models.py`` and, transitively, for the CUDA path.
"""
from __future__ import annotations

import copy
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SWR_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _features(spec):
    from scenario_wise_rec.basic.features import DenseFeature, SparseFeature
    out = []
    for name, kind, vocab, dim in spec:
        out.append(SparseFeature(name, vocab_size=vocab, embed_dim=dim) if kind == "sparse" else DenseFeature(name))
    return out


def _batch(spec, B, D, gen, dtype_zoo=False):
    x = {}
    zoo_i = [torch.int64, torch.int32, torch.int16, torch.int8]
    zoo_f = [torch.float32, torch.float16, torch.float64]
    si = di = 0
    for name, kind, vocab, _dim in spec:
        if kind == "sparse":
            v = torch.randint(0, vocab, (B,), generator=gen)
            if dtype_zoo:
                dt = zoo_i[si % 4]
                if dt == torch.int8 and vocab > 127:
                    dt = torch.int16
                v = v.to(dt)
            si += 1
        else:
            v = torch.rand(B, generator=gen)
            if dtype_zoo:
                v = v.to(zoo_f[di % 3])
            di += 1
        x[name] = v
    dom = torch.randint(0, D, (B,), generator=gen)
    dom[0] = D + 1          # a row whose domain id is outside [0, D): output must be 0
    x["domain_indicator"] = dom
    y = (torch.rand(B, generator=gen) < 0.3).float()
    return x, y


def _randomise(model, gen):
    """Make every tensor numerically non-trivial but well-conditioned."""
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "embed_dict" in name:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.5)
            elif name.endswith("deep_weights"):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.5 + 0.3)
            elif p.dim() == 1:
                p.add_(torch.randn(p.shape, generator=gen) * 0.1)
            elif name.startswith(("u.", "v.")):      # HAMUR u/v are all-ones; keep scale sane
                p.copy_(torch.randn(p.shape, generator=gen) * 0.3)
        for name, b in model.named_buffers():
            if name.endswith("running_mean"):
                b.copy_(torch.randn(b.shape, generator=gen) * 0.1)
            elif name.endswith("running_var"):
                b.copy_(torch.rand(b.shape, generator=gen) + 0.5)
            elif name.endswith("num_batches_tracked"):
                b.fill_(3)


SP = lambda n, v, d=8: (n, "sparse", v, d)   # noqa: E731
DE = lambda n: (n, "dense", 0, 1)            # noqa: E731

FEATS_A = [DE("d0"), SP("s0", 50), SP("s1", 7), DE("d1"), SP("s2", 300), SP("s3", 13), SP("s4", 2)]
FEATS_SPARSE = [SP("s0", 50, 16), SP("s1", 7, 16), SP("s2", 200, 16)]


def cases():
    c = []
    c.append(("sharedbottom_small", "SharedBottom", dict(features=FEATS_A, domain_num=3, bottom_dims=[32], tower_dims=[8]), 64, False))
    c.append(("sharedbottom_zoo", "SharedBottom", dict(features=FEATS_A, domain_num=3, bottom_dims=[24, 12], tower_dims=[8, 4]), 37, True))
    c.append(("mmoe_small", "MMOE", dict(features=FEATS_A, domain_num=3, n_expert=4, expert_dims=[32, 16, 8], tower_dims=[16]), 64, False))
    c.append(("ple_small", "PLE", dict(features=FEATS_A, domain_num=3, n_level=1, n_expert_specific=2, n_expert_shared=2,
                                       expert_dims=[16, 8], tower_dims=[8]), 64, False))
    c.append(("ple_2level", "PLE", dict(features=FEATS_A, domain_num=2, n_level=2, n_expert_specific=1, n_expert_shared=2,
                                        expert_dims=[16, 8], tower_dims=[8]), 48, False))
    c.append(("star_small", "Star", dict(features=FEATS_A, domain_num=3, fcn_dims=[32, 16, 8], aux_dims=[16]), 64, False))
    sce = [SP("scene", 3, 8)]
    c.append(("ppnet_small", "PPNet", dict(id_features=[SP("s0", 50), SP("s2", 300)],
                                           agn_features=[SP("s1", 7), SP("s3", 13), DE("d0"), DE("d1")] + sce,
                                           domain_num=3, fcn_dims=[32, 16, 8]), 64, False))
    c.append(("epnet_small", "EPNet", dict(sce_features=sce, agn_features=[SP("s0", 50), SP("s1", 7), SP("s3", 13), DE("d0")],
                                           fcn_dims=[32, 16, 8], domain_num=3), 64, False))
    c.append(("m3oe_small", "M3oE", dict(features=FEATS_A, domain_num=3, fcn_dims=[32, 16, 16, 8], expert_num=4), 64, False))
    c.append(("hamursmall_small", "HamurSmall", dict(features=FEATS_SPARSE, domain_num=4, fcn_dims=[32, 16], hyper_dims=[8], k=5), 64, False))
    c.append(("hamurlarge_small", "HamurLarge", dict(features=FEATS_A, domain_num=2, fcn_dims=[32, 32, 16, 16, 16, 16, 8],
                                                     hyper_dims=[8], k=5), 48, False))
    return c


def build_reference(model_name, cfg):
    import scenario_wise_rec.models.multi_domain as M
    f = lambda key: _features(cfg[key])   # noqa: E731   fresh Feature objects: they cache nn.Embedding
    D = cfg.get("domain_num")
    if model_name == "SharedBottom":
        return M.SharedBottom(f("features"), D, bottom_params={"dims": list(cfg["bottom_dims"])},
                              tower_params={"dims": list(cfg["tower_dims"])})
    if model_name == "MMOE":
        return M.MMOE(f("features"), D, n_expert=cfg["n_expert"], expert_params={"dims": list(cfg["expert_dims"])},
                      tower_params={"dims": list(cfg["tower_dims"])})
    if model_name == "PLE":
        return M.PLE(f("features"), D, n_level=cfg["n_level"], n_expert_specific=cfg["n_expert_specific"],
                     n_expert_shared=cfg["n_expert_shared"], expert_params={"dims": list(cfg["expert_dims"])},
                     tower_params={"dims": list(cfg["tower_dims"])})
    if model_name == "Star":
        return M.Star(f("features"), D, fcn_dims=list(cfg["fcn_dims"]), aux_dims=list(cfg["aux_dims"]))
    if model_name == "PPNet":
        return M.PPNet(id_features=f("id_features"), agn_features=f("agn_features"), domain_num=D,
                       fcn_dims=list(cfg["fcn_dims"]))
    if model_name == "EPNet":
        return M.EPNet(sce_features=f("sce_features"), agn_features=f("agn_features"), fcn_dims=list(cfg["fcn_dims"]))
    if model_name == "M3oE":
        return M.M3oE(f("features"), D, fcn_dims=list(cfg["fcn_dims"]), expert_num=cfg["expert_num"],
                      exp_d=1, exp_t=1, bal_d=1, bal_t=1, device="cpu")
    if model_name == "HamurSmall":
        return M.HamurSmall(f("features"), D, fcn_dims=list(cfg["fcn_dims"]), hyper_dims=list(cfg["hyper_dims"]), k=cfg["k"])
    if model_name == "HamurLarge":
        return M.HamurLarge(f("features"), D, fcn_dims=list(cfg["fcn_dims"]), hyper_dims=list(cfg["hyper_dims"]), k=cfg["k"])
    raise KeyError(model_name)


def all_feature_specs(cfg):
    seen, out = set(), []
    for key in ("features", "id_features", "agn_features", "sce_features"):
        for s in cfg.get(key, []):
            if s[0] not in seen:
                seen.add(s[0])
                out.append(s)
    return out


def run_case(name, model_name, cfg, B, zoo, seed):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    model = build_reference(model_name, cfg)
    _randomise(model, gen)
    x, y = _batch(all_feature_specs(cfg), B, cfg["domain_num"], gen, zoo)
    if model_name == "EPNet":
        x["domain_indicator"] = x["domain_indicator"].clamp(max=cfg["domain_num"] - 1)
    state0 = copy.deepcopy(model.state_dict())

    model.train()
    out = model(x)
    loss = torch.nn.BCELoss()(out, y)
    model.zero_grad()
    loss.backward()
    grads = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in model.named_parameters()}
    state1 = copy.deepcopy(model.state_dict())

    model.load_state_dict(state0)
    model.eval()
    with torch.no_grad():
        out_eval = model(x)

    blob = {"cfg": np.array(json.dumps(cfg)), "model": np.array(model_name), "B": np.array(B)}
    for k, v in x.items():
        blob[f"x/{k}"] = v.numpy()
    blob["y"] = y.numpy()
    for k, v in state0.items():
        blob[f"state0/{k}"] = v.numpy()
    for k, v in state1.items():
        if "running_" in k or "num_batches" in k:
            blob[f"state1/{k}"] = v.numpy()
    blob["out_train"] = out.detach().numpy()
    blob["out_eval"] = out_eval.numpy()
    blob["loss"] = loss.detach().numpy()
    none = []
    for k, g in grads.items():
        if g is None:
            none.append(k)
        else:
            blob[f"grad/{k}"] = g.numpy()
    blob["grad_none"] = np.array(json.dumps(none))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **blob)
    print(f"{name}: B={B} out[:3]={out[:3].tolist()} loss={float(loss):.6f} params={len(grads)} none={len(none)}")


def main():
    sys.path.insert(0, REF)
    for i, (name, model_name, cfg, B, zoo) in enumerate(cases()):
        run_case(name, model_name, cfg, B, zoo, seed=100 + i)


if __name__ == "__main__":
    main()
