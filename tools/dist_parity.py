"""N-process parity check used by bench.py (before timing, every N >= 2) and by tests/dist_p2p_check.py: on a reference-generated
golden case (tests/golden/mmoe_small.npz) the fused training step with ROW-SHARDED tables (peer-memory K1 / K2, one
barrier + one all-reduce per step) must follow the same trajectory as the fused step with REPLICATED tables, each rank
training on its slice of the golden batch.  Returns the largest loss and state differences seen."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "scenario-wise-rec_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def run(dev, steps=4, min_rows=40):
    from golden_util import Golden
    import model_factory
    from scenario_wise_rec_b200 import parallel
    from scenario_wise_rec_b200.basic.features import SparseFeature
    from scenario_wise_rec_b200.trainers import CTRTrainer
    import scenario_wise_rec_b200.models.multi_domain as M
    rank, world = dist.get_rank(), dist.get_world_size()
    g = Golden("mmoe_small")
    per = g.B // world
    sl = slice(rank * per, (rank + 1) * per)
    x, y = {k: v[sl] for k, v in g.x.items()}, g.y[sl]
    noisy = {k for k, v in g.grads.items() if float(v.abs().max()) < 1e-6} | {k for k in g.state0 if k.endswith("running_mean")}
    opt = {"lr": 1e-2, "weight_decay": 1e-4}

    def build(shard):
        feats = model_factory.features(g.cfg["features"])
        names = parallel.shard_features(feats, min_rows=min_rows) if shard else []
        mm = M.MMOE(feats, g.cfg["domain_num"], n_expert=g.cfg["n_expert"], expert_params={"dims": list(g.cfg["expert_dims"])},
                    tower_params={"dims": list(g.cfg["tower_dims"])})
        return mm, feats, names
    rep, _, _ = build(False)
    rep.load_state_dict(g.state0)
    sh, feats, names = build(True)
    info = next(f.shard for f in feats if isinstance(f, SparseFeature) and f.shard is not None)
    st = dict(g.state0)
    for n in names:
        st[f"embedding.embed_dict.{n}.weight"] = parallel.shard_of(g.state0[f"embedding.embed_dict.{n}.weight"], info)
    sh.load_state_dict(st)
    ts = []
    for mm in (rep, sh):
        tt = CTRTrainer(mm, "parity", optimizer_params=opt, device=str(dev))
        tt.enable_data_parallel()
        mm.train()
        ts.append(tt)
    dloss = 0.0
    for _ in range(steps):
        la, lb = ts[0].train_step(x, y).item(), ts[1].train_step(x, y).item()
        dloss = max(dloss, abs(la - lb))
    torch.cuda.synchronize()
    fs = next(iter(ts[1]._steps.values()))
    sa, sb = rep.state_dict(), sh.state_dict()
    derr = 0.0
    for k in sa:
        if k in noisy or not sa[k].dtype.is_floating_point:
            continue
        want = sa[k]
        if any(k == f"embedding.embed_dict.{n}.weight" for n in names):
            want = parallel.shard_of(want, info)
        derr = max(derr, float((sb[k] - want).abs().max()))
    t = torch.tensor([dloss, derr], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"case": "mmoe_small golden, %d steps, batch split over %d ranks" % (steps, world), "sharded_tables": names,
            "peer_memory_path": bool(fs.p2p), "captured_in_graph": fs.graph is not None,
            "max_loss_diff_vs_replicated": float(t[0]), "max_state_diff_vs_replicated": float(t[1]),
            "ok": bool(float(t[0]) < 2e-5 and float(t[1]) < 5e-5)}
