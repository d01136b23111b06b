"""Build product models (scenario_wise_rec_b200) from the config dicts stored in the goldens."""
import torch

from scenario_wise_rec_b200.basic.features import DenseFeature, SparseFeature
import scenario_wise_rec_b200.models.multi_domain as M


def features(spec, shard_min_rows=0):
    """Fresh feature objects (they cache their nn.Embedding).  shard_min_rows > 0 (one process per GPU): tables with at
    least that many rows are row-sharded over the default process group (parallel.shard_features)."""
    feats = [SparseFeature(n, vocab_size=v, embed_dim=d) if k == "sparse" else DenseFeature(n) for n, k, v, d in spec]
    if shard_min_rows > 0:
        from scenario_wise_rec_b200 import parallel
        parallel.shard_features(feats, min_rows=shard_min_rows)
    return feats


def build(model_name, cfg, shard_min_rows=0):
    f = lambda key: features(cfg[key], shard_min_rows)   # noqa: E731
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


def supported(model_name):
    return hasattr(M, model_name)


def check_against_golden(model, g, device="cpu", out_tol=1e-4, grad_rtol=2e-4):
    """Train fwd/bwd + BN buffers + eval fwd of ``model`` (already holding g.state0) vs the golden.
    Tolerances: outputs 1e-4 abs (north_star); gradients absolute, scaled by max|grad| of the tensor."""
    x = {k: v.to(device) for k, v in g.x.items()}
    y = g.y.to(device)
    model.train()
    out = model(x)
    assert out.shape == g.out_train.shape
    err = float((out.detach().cpu() - g.out_train).abs().max())
    assert err <= out_tol, f"train output err {err}"
    loss = torch.nn.BCELoss()(out, y)
    model.zero_grad()
    loss.backward()
    assert abs(float(loss.detach()) - g.loss) <= 1e-4
    params = dict(model.named_parameters())
    assert set(params) == set(g.param_names()), set(params) ^ set(g.param_names())
    worst = ("", 0.0)
    for k, p in params.items():
        if k in g.grad_none:
            assert p.grad is None, f"{k} should have no gradient"
            continue
        assert p.grad is not None, f"{k} has no gradient"
        ref = g.grads[k]
        scale = max(float(ref.abs().max()), 1e-3)
        e = float((p.grad.detach().cpu() - ref).abs().max()) / scale
        if e > worst[1]:
            worst = (k, e)
    assert worst[1] <= grad_rtol, f"gradient {worst[0]} rel-to-max err {worst[1]}"
    st = model.state_dict()
    for k, v in g.state1.items():
        torch.testing.assert_close(st[k].cpu().to(v.dtype), v, atol=2e-5, rtol=1e-4, msg=lambda m, k=k: f"{k}: {m}")
    model.load_state_dict(g.state0)
    model.eval()
    with torch.no_grad():
        out_eval = model(x)
    err = float((out_eval.cpu() - g.out_eval).abs().max())
    assert err <= out_tol, f"eval output err {err}"
    return worst
