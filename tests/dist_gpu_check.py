"""Multi-GPU checks, one process per GPU:  torchrun --nproc-per-node 2 tests/dist_gpu_check.py
(wrapped by tests/test_gpu_dist.py).  NCCL, real kernels, fused (CUDA graph) and generic paths."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "scenario-wise-rec_b200"), HERE, os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

from golden_util import Golden
import model_factory
from scenario_wise_rec_b200 import parallel
from scenario_wise_rec_b200.basic.features import SparseFeature
from scenario_wise_rec_b200.trainers import CTRTrainer
import scenario_wise_rec_b200.models.multi_domain as M


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    opt = {"lr": 1e-2, "weight_decay": 1e-4}

    # 1. data parallel (fused step, gradient exchange inside the graph) == one process on the global batch (no batch
    #    coupling), with the table gradients all-reduced densely and exchanged as rows (SWR_DP_SPARSE)
    g = Golden("m3oe_small")
    half = g.B // world
    sl = slice(rank * half, (rank + 1) * half)
    for sparse in ("0", "1"):
        os.environ["SWR_DP_SPARSE"] = sparse
        m = model_factory.build(g.model, g.cfg); m.load_state_dict(g.state0)
        t = CTRTrainer(m, "dp", optimizer_params=opt, device=str(dev)); t.enable_data_parallel(); m.train()
        ref = model_factory.build(g.model, g.cfg); ref.load_state_dict(g.state0)
        tr = CTRTrainer(ref, "single", optimizer_params=opt, device=str(dev)); ref.train()
        for _ in range(5):
            t.train_step({k: v[sl] for k, v in g.x.items()}, g.y[sl])
            tr.train_step(g.x, g.y)
        torch.cuda.synchronize()
        fs = next(iter(t._steps.values()))
        assert fs.graph is not None, "DP step was not captured in a CUDA graph"
        assert (fs.sparse_sync is not None) == (sparse == "1"), (sparse, fs.sparse_sync is not None)
        for k, v in ref.state_dict().items():
            torch.testing.assert_close(m.state_dict()[k], v, atol=5e-6, rtol=2e-4, msg=lambda s, k=k: f"dp sparse={sparse} {k}: {s}")
    os.environ.pop("SWR_DP_SPARSE", None)

    # 2. row-sharded tables == replicated tables, fused and generic paths
    g = Golden("mmoe_small")
    half = g.B // world
    sl = slice(rank * half, (rank + 1) * half)
    x, y = {k: v[sl] for k, v in g.x.items()}, g.y[sl]
    noisy = {k for k, v in g.grads.items() if float(v.abs().max()) < 1e-6} | {k for k in g.state0 if k.endswith("running_mean")}
    for fused in (True, False):
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
        ts = []
        for mm in (rep, sh):
            tt = CTRTrainer(mm, "dp", optimizer_params=opt, device=str(dev), fused=fused); tt.enable_data_parallel(); mm.train(); ts.append(tt)
        for step in range(5):
            la, lb = ts[0].train_step(x, y).item(), ts[1].train_step(x, y).item()
            assert abs(la - lb) < 2e-5, (fused, step, la, lb)
        torch.cuda.synchronize()
        if fused:
            assert next(iter(ts[1]._steps.values())).graph is not None, "sharded step was not captured in a CUDA graph"
        sa, sb = rep.state_dict(), sh.state_dict()
        for k in sa:
            want = sa[k]
            if any(k == f"embedding.embed_dict.{n}.weight" for n in names):
                want = parallel.shard_of(want, info)
            tol = dict(atol=0.12, rtol=0) if k in noisy else dict(atol=2e-5, rtol=2e-4)
            torch.testing.assert_close(sb[k], want, **tol, msg=lambda s, k=k: f"shard fused={fused} {k}: {s}")
        rep.eval(), sh.eval()
        with torch.no_grad():
            xd = {k: v.to(dev) for k, v in x.items()}
            torch.testing.assert_close(sh(xd), rep(xd), atol=1e-5, rtol=1e-4)
    dist.barrier()
    if rank == 0:
        print("DIST_GPU_CHECK_OK world", world, flush=True)
    # captured step graphs hold NCCL work on the communicator: destroy_process_group() would wait on it forever
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
