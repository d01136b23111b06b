"""2-GPU check of the peer-memory path of row-sharded tables (fused step): torchrun --nproc-per-node 2 tests/dist_p2p_check.py"""
import os, sys, faulthandler
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "scenario-wise-rec_b200"), HERE, os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch, torch.distributed as dist
faulthandler.dump_traceback_later(100, exit=True)
from golden_util import Golden
import model_factory
from scenario_wise_rec_b200 import parallel
from scenario_wise_rec_b200.basic.features import SparseFeature
from scenario_wise_rec_b200.trainers import CTRTrainer
import scenario_wise_rec_b200.models.multi_domain as M


def log(*a):
    print(f"[r{os.environ['RANK']}]", *a, file=sys.stderr, flush=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    opt = {"lr": 1e-2, "weight_decay": 1e-4}
    g = Golden("mmoe_small")
    half = g.B // world
    sl = slice(rank * half, (rank + 1) * half)
    x, y = {k: v[sl] for k, v in g.x.items()}, g.y[sl]
    noisy = {k for k, v in g.grads.items() if float(v.abs().max()) < 1e-6} | {k for k in g.state0 if k.endswith("running_mean")}

    def build(shard):
        feats = model_factory.features(g.cfg["features"])
        names = parallel.shard_features(feats, min_rows=40) if shard else []
        mm = M.MMOE(feats, g.cfg["domain_num"], n_expert=g.cfg["n_expert"], expert_params={"dims": list(g.cfg["expert_dims"])},
                    tower_params={"dims": list(g.cfg["tower_dims"])})
        return mm, feats, names
    rep, _, _ = build(False); rep.load_state_dict(g.state0)
    sh, feats, names = build(True)
    info = next(f.shard for f in feats if isinstance(f, SparseFeature) and f.shard is not None)
    st = dict(g.state0)
    for n in names:
        st[f"embedding.embed_dict.{n}.weight"] = parallel.shard_of(g.state0[f"embedding.embed_dict.{n}.weight"], info)
    sh.load_state_dict(st)
    log("models built; sharded:", names)
    ts = []
    for mm in (rep, sh):
        tt = CTRTrainer(mm, "dp", optimizer_params=opt, device=str(dev)); tt.enable_data_parallel(); mm.train(); ts.append(tt)
    for step in range(5):
        la = ts[0].train_step(x, y).item()
        log("step", step, "replicated loss", la)
        lb = ts[1].train_step(x, y).item()
        log("step", step, "sharded loss", lb)
        assert abs(la - lb) < 2e-5, (step, la, lb)
    torch.cuda.synchronize()
    fs = next(iter(ts[1]._steps.values()))
    assert fs.p2p and fs.graph is not None, (fs.p2p, fs.graph)
    sa, sb = rep.state_dict(), sh.state_dict()
    for k in sa:
        want = sa[k]
        if any(k == f"embedding.embed_dict.{n}.weight" for n in names):
            want = parallel.shard_of(want, info)
        tol = dict(atol=0.12, rtol=0) if k in noisy else dict(atol=2e-5, rtol=2e-4)
        torch.testing.assert_close(sb[k], want, **tol, msg=lambda s, k=k: f"{k}: {s}")
    dist.barrier()
    if rank == 0:
        print("DIST_P2P_CHECK_OK world", world)
    os._exit(0)


if __name__ == "__main__":
    main()
