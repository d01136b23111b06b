"""World-size-2 tests of the multi-GPU host logic on CPU (gloo): data-parallel gradient averaging and the
row-sharded embedding exchange (all_gather of indices, all_to_all of looked-up rows and of their gradients).
The device programs run on the CPU interpreter (oracle/ops_ref.RefRunner) and the local K1/K2 calls on their
torch stand-ins -- checker code injected by the test, exactly like tests/test_program_cpu.py."""
import os
import socket
import sys
import traceback

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(rank, world, port):
    for p in (ROOT, os.path.join(ROOT, "scenario-wise-rec_b200"), HERE, os.path.join(ROOT, "tools")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ops_ref
    from scenario_wise_rec_b200 import parallel
    from scenario_wise_rec_b200.fused import FusedModule
    FusedModule._runner_factory = ops_ref.RefRunner
    parallel.local_gather, parallel.local_scatter = ops_ref.ref_local_gather, ops_ref.ref_local_scatter


def _worker(fn, rank, world, port, q):
    try:
        _setup(rank, world, port)
        fn(rank, world)
        q.put((rank, None))
    except Exception:
        q.put((rank, traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def _spawn(fn, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(fn, r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    errs = [f"rank {r}:\n{e}" for r, e in res if e]
    assert not errs, "\n".join(errs)


# ---- data parallel == one process on the global batch (model without batch coupling) ---------------------
def _dp_equals_global_batch(rank, world):
    from golden_util import Golden
    import model_factory
    from scenario_wise_rec_b200.trainers import CTRTrainer
    g = Golden("m3oe_small")             # LayerNorm only: no batch statistics, so DP == global batch exactly
    half = g.B // world
    sl = slice(rank * half, (rank + 1) * half)
    m = model_factory.build(g.model, g.cfg)
    m.load_state_dict(g.state0)
    t = CTRTrainer(m, "dp", optimizer_params={"lr": 1e-2, "weight_decay": 1e-4}, device="cpu", fused=False)
    t.enable_data_parallel()
    m.train()
    ref = model_factory.build(g.model, g.cfg)
    ref.load_state_dict(g.state0)
    tr = CTRTrainer(ref, "single", optimizer_params={"lr": 1e-2, "weight_decay": 1e-4}, device="cpu", fused=False)
    ref.train()
    for _ in range(3):
        t.train_step({k: v[sl] for k, v in g.x.items()}, g.y[sl])
        tr.train_step(g.x, g.y)
    sa, sb = m.state_dict(), ref.state_dict()
    for k in sb:
        torch.testing.assert_close(sa[k], sb[k], atol=2e-6, rtol=1e-4, msg=lambda s, k=k: f"{k}: {s}")


def test_data_parallel_equals_global_batch():
    _spawn(_dp_equals_global_batch)


# ---- row-sharded table == replicated table under the same data-parallel split ------------------------------
def _sharded_equals_replicated(rank, world):
    from golden_util import Golden
    import model_factory
    from scenario_wise_rec_b200 import parallel
    from scenario_wise_rec_b200.basic.features import SparseFeature
    from scenario_wise_rec_b200.trainers import CTRTrainer
    import scenario_wise_rec_b200.models.multi_domain as M
    g = Golden("mmoe_small")
    half = g.B // world
    sl = slice(rank * half, (rank + 1) * half)
    x, y = {k: v[sl] for k, v in g.x.items()}, g.y[sl]
    cfg = g.cfg

    def build(shard):
        feats = model_factory.features(cfg["features"])
        names = parallel.shard_features(feats, min_rows=40) if shard else []
        m = M.MMOE(feats, cfg["domain_num"], n_expert=cfg["n_expert"], expert_params={"dims": list(cfg["expert_dims"])},
                   tower_params={"dims": list(cfg["tower_dims"])})
        return m, feats, names

    rep, _, _ = build(False)
    rep.load_state_dict(g.state0)
    sh, feats, names = build(True)
    assert sorted(names) == ["s0", "s2"]                      # vocab 50 and 300
    st = dict(g.state0)
    info = next(f.shard for f in feats if isinstance(f, SparseFeature) and f.shard is not None)
    for n in names:
        k = f"embedding.embed_dict.{n}.weight"
        st[k] = parallel.shard_of(g.state0[k], info)
        assert sh.state_dict()[k].shape == st[k].shape
    sh.load_state_dict(st)
    trainers = []
    for m in (rep, sh):
        t = CTRTrainer(m, "dp", optimizer_params={"lr": 1e-2, "weight_decay": 1e-4}, device="cpu", fused=False)
        t.enable_data_parallel()
        m.train()
        trainers.append(t)
    for step in range(3):
        la = trainers[0].train_step(x, y).item()
        lb = trainers[1].train_step(x, y).item()
        assert abs(la - lb) < 1e-6, (step, la, lb)
    sa, sb = rep.state_dict(), sh.state_dict()
    for k in sa:
        want = sa[k]
        if any(k == f"embedding.embed_dict.{n}.weight" for n in names):
            want = parallel.shard_of(want, info)
            # padding rows of the shard (beyond the vocabulary) only see weight decay of zeros: stay zero
        torch.testing.assert_close(sb[k], want, atol=2e-6, rtol=1e-4, msg=lambda s, k=k: f"{k}: {s}")
    # eval forward through the exchange (no autograd)
    rep.eval(), sh.eval()
    with torch.no_grad():
        torch.testing.assert_close(sh(x), rep(x), atol=1e-6, rtol=1e-5)
    # out-of-range index on a sharded field
    bad = dict(x)
    bad["s2"] = bad["s2"].clone()
    bad["s2"][0] = 300
    with torch.no_grad():
        sh(bad)
    with pytest.raises(IndexError):
        sh.check_indices()


def test_sharded_table_equals_replicated():
    _spawn(_sharded_equals_replicated)


# ---- checkpoints of row-sharded models keep the reference's format ---------------------
def _sharded_checkpoint_roundtrip(rank, world):
    from golden_util import Golden
    import model_factory
    from scenario_wise_rec_b200 import parallel
    import scenario_wise_rec_b200.models.multi_domain as M
    g = Golden("mmoe_small")
    cfg = g.cfg

    def build(shard):
        feats = model_factory.features(cfg["features"])
        names = parallel.shard_features(feats, min_rows=40) if shard else []
        m = M.MMOE(feats, cfg["domain_num"], n_expert=cfg["n_expert"], expert_params={"dims": list(cfg["expert_dims"])},
                   tower_params={"dims": list(cfg["tower_dims"])})
        return m, names

    sh, names = build(True)
    assert names
    # reference-format state (full tables) -> sharded model: every rank keeps its rows
    parallel.load_full_state_dict(sh, g.state0)
    for n in names:
        k = f"embedding.embed_dict.{n}.weight"
        full = g.state0[k]
        mine = full[rank::world]
        assert torch.equal(sh.state_dict()[k][:mine.shape[0]], mine)
    # sharded model -> reference-format state: identical to what was loaded, on every rank, and loadable
    # (strict) into an unsharded model
    sd = parallel.full_state_dict(sh)
    assert set(sd) == set(g.state0)
    for k, v in g.state0.items():
        assert sd[k].shape == v.shape and torch.equal(sd[k], v), k
    rep, _ = build(False)
    rep.load_state_dict(sd, strict=True)


def test_sharded_checkpoint_roundtrip():
    _spawn(_sharded_checkpoint_roundtrip)
