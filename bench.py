#!/usr/bin/env python
"""bench.py -- CTR training samples/sec of the Scenario-Wise-Rec hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one synthetic
batch exactly as the reference trainer drives it (scenario_wise_rec/trainers/ctr_trainer.py:67-73):
EmbeddingLayer gather -> expert / gate / tower stack -> BCELoss -> zero_grad -> backward (incl.
embedding-gradient scatter) -> Adam(lr 1e-3, weight_decay 1e-5).

* ``value``     samples/s, inputs resident in HBM, CUDA-event timed, max over ranks
* ``e2e``       the same through ``CTRTrainer.train_step`` with pinned HOST batches: the H2D copy
                of every feature column and a D2H read of the loss are inside the timed region
* ``roofline``  the dominant kernel of the step, timed live with CUDA events on the launch stream
                (swr_profile_begin/end of the C ABI), against MEASURED_PEAKS.json
* ``cpu_baseline`` the CPU oracle port (oracle/ref_models.py + torch autograd + torch Adam) on this
                box's host cores, bounded sample (rank 0, N=1 only)
* ``--impl reference``: only the CPU path (the reference is pure PyTorch; its restatement
  oracle/ref_models.py issues the same ATen ops), all host threads, same JSON keys.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "scenario-wise-rec_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import workloads  # noqa: E402

METRIC = "CTR training samples/sec (bs=4096 per GPU)"
UNIT = "samples/s"
NB_ROTATE = 8        # distinct batches, rotated (different rows touched every step)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]), bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons polled in-process through NVML every ~2 ms on a thread, from the start of the
    warm-up to the end of the timed regions (a 14 ms timed region is invisible to ``nvidia-smi -lms 100``).
    ``mark()`` brackets the timed regions; ``summary()`` reports the samples inside them (and the total)."""

    def __init__(self, index):
        self.index, self.rows, self.marks = index, [], []
        self._stop = threading.Event()
        self.t = None
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical(index))
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    @staticmethod
    def _physical(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def __enter__(self):
        if self.h is not None:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        return self

    def _poll(self):
        nv = self.nv
        bits = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.perf_counter(), sm, tuple(n for n, b in bits if r & b)))
            except Exception:
                pass
            time.sleep(0.002)

    def mark(self):
        self.marks.append(time.perf_counter())

    def __exit__(self, *a):
        self._stop.set()
        if self.t is not None:
            self.t.join(timeout=2)

    def summary(self):
        spans = list(zip(self.marks[0::2], self.marks[1::2]))
        inside = [r for r in self.rows if any(a <= r[0] <= b for a, b in spans)] if spans else []
        use = inside if inside else self.rows
        if not use:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted({n for r in use for n in r[2]})
        return {"sm_mhz": float(np.median([r[1] for r in use])), "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(use), "samples_in_timed_region": len(inside), "samples_total": len(self.rows),
                "how": "NVML polled every ~2 ms from warm-up to the end of the timed regions"}


# --------------------------------------------------------------------------------------------
# CPU path (oracle port): the reference's own algorithm on the host cores
# --------------------------------------------------------------------------------------------
class CpuReference:
    """fwd + BCELoss + zero_grad + backward + Adam of the restated reference forward (oracle/ref_models.py),
    torch CPU with every host thread.  Used for ``cpu_baseline`` and ``--impl reference`` only."""

    def __init__(self, case, seed=0):
        from oracle import ref_models
        import model_factory
        self.rm = ref_models
        self.model_name, self.cfg, self.B = workloads.CASES[case]
        torch.manual_seed(seed)
        # parameter shapes / init come from the product's parameter containers (plain torch modules on CPU);
        # no product compute is involved: the forward below is the oracle's.
        m = model_factory.build(self.model_name, self.cfg)
        self.state = {}
        for k, v in m.state_dict().items():
            v = v.detach().clone()
            if v.dtype.is_floating_point and "running_" not in k:
                v.requires_grad_(True)
            self.state[k] = v
        self.params = [v for v in self.state.values() if v.requires_grad]
        self.opt = torch.optim.Adam(self.params, lr=1e-3, weight_decay=1e-5)
        self.cores = torch.get_num_threads()

    def step(self, x, y):
        bn_out = {}
        out = self.rm.forward(self.model_name, x, self.state, self.cfg, training=True, bn_out=bn_out)
        loss = torch.nn.functional.binary_cross_entropy(out, y)
        for p in self.params:
            p.grad = None
        loss.backward()
        self.opt.step()
        with torch.no_grad():
            for k, v in bn_out.items():
                self.state[k] = v
        return float(loss.detach())

    def time(self, steps, warmup, budget_s=None):
        feats = workloads.all_feature_specs(self.cfg)
        batches = [workloads.make_batch(feats, self.B, self.cfg["domain_num"], seed=100 + i) for i in range(4)]
        for i in range(warmup):
            self.step(*batches[i % 4])
        t0 = time.perf_counter()
        n = 0
        for i in range(steps):
            self.step(*batches[i % 4])
            n += 1
            if budget_s is not None and time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
        return n, dt


def base_config(args, world, parallelism=None):
    """``config`` of the JSON line: identical keys (and, for the same flags, values) in both arms."""
    model_name, cfg, B = workloads.CASES[args.workload]
    feats = workloads.all_feature_specs(cfg)
    if parallelism is None:
        parallelism = "single" if world == 1 else (f"dp{world}" if args.shard_min_rows <= 0 else f"dp{world}+rowshard")
    return {"workload": args.workload, "model": model_name, "batch_per_gpu": B, "global_batch": B * world, "parallelism": parallelism,
            "shard_min_rows": args.shard_min_rows if world > 1 else 0,
            "l2": "working set per step (tables + dense grads + Adam moments, %.0f MB) exceeds the 126 MB L2; %d rotating batches"
                  % (4 * workloads.table_bytes(feats) / 1e6, NB_ROTATE),
            "optimizer": "Adam(lr=1e-3, weight_decay=1e-5)"}


def reference_arm(args, device, steps, warmup, budget_s=None):
    """The unmodified reference (baseline/_ref, tools/ref_arm.py) through its own CTRTrainer.train_one_epoch on
    ``device``; falls back to the oracle port on the CPU when baseline/_ref did not travel.  -> dict for the line."""
    import ref_arm
    model_name, cfg, B = workloads.CASES[args.workload]
    feats = workloads.all_feature_specs(cfg)
    batches = [workloads.make_batch(feats, B, cfg["domain_num"], seed=100 + i, pin=(device != "cpu")) for i in range(4)]
    if ref_arm.available():
        arm = ref_arm.ReferenceArm(model_name, cfg, device=device)
        n, dt = arm.time(batches, steps, warmup, budget_s=budget_s)
        kind, cores = "reference", arm.cores
        path = "baseline/_ref scenario_wise_rec (unmodified): models.multi_domain.%s + trainers.CTRTrainer.train_one_epoch" % model_name
    else:
        if device != "cpu":
            return None
        ref = CpuReference(args.workload)
        n, dt = ref.time(steps, warmup, budget_s=budget_s)
        kind, cores, path = "port", ref.cores, "oracle/ref_models.py (baseline/_ref missing)"
    return {"value": n * B / dt, "unit": UNIT, "cores": cores, "kind": kind, "steps": n, "ms_per_step": dt / n * 1e3, "device": device,
            "path": path, "sample": f"{n} full train steps of {args.workload} (B={B}) after {warmup} warm-up, {dt:.1f} s"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    # bounded: the whole run must end within a few minutes whatever --steps says
    r = reference_arm(args, "cpu", args.steps, min(args.warmup, 3), budget_s=args.ref_budget)
    val = r["value"]
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"], "warmup": min(args.warmup, 3),
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": base_config(args, world),
        "reference_path": r["path"],
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# --------------------------------------------------------------------------------------------
# op costs (algorithmic bytes / flops) from the step's records
# --------------------------------------------------------------------------------------------
FAMILY = {"fc_fwd": "fc_fwd", "fc_dgrad": "fc_dgrad", "fc_wgrad": "fc_wgrad", "fc_presplit": "fc_presplit", "gather": "gather", "scatter": "scatter",
          "adam": "optimizer", "adam_rows": "optimizer", "adam_flush": "optimizer", "zero": "optimizer"}


def op_table(recs, B, feats):
    """header index -> dict(name, family, bound, work) for every op of a record list.  work = algorithmic bytes (HBM-bound
    ops) or fp32 FLOPs (FC ops: 2 * B * K * N per group; not multiplied by the 3x of the operand split)."""
    from scenario_wise_rec_b200 import _native as N
    names = {getattr(N, k): k[3:].lower() for k in dir(N) if k.startswith("OP_")}
    i64 = lambda r: (int(r["i"][0]) & 0xFFFFFFFF) | (int(r["i"][1]) << 32)      # noqa: E731
    out = {}
    i = 0
    while i < len(recs):
        h = recs[i]
        kind, ns = int(h["kind"]), int(h["n_sub"])
        subs = recs[i + 1:i + 1 + ns]
        base = names.get(kind, str(kind))
        d = {"name": base, "family": FAMILY.get(base, base), "bound": "hbm", "work": None}
        if kind in (N.OP_FC_FWD, N.OP_FC_DGRAD, N.OP_FC_WGRAD):
            fl = sum(2.0 * B * int(r["i"][13] if int(r["i"][13]) > 0 else r["i"][1]) * int(r["i"][5]) for r in subs)
            d.update(bound="tensor", work=fl, name=f"{base}[{ns}g K{int(subs[0]['i'][1])} N{int(subs[0]['i'][5])}]")
        elif kind == N.OP_FC_PRESPLIT:
            d["work"] = float(sum(4.0 * int(r["i"][1]) * int(r["i"][5]) * 5 for r in subs))      # read W, write 2 planes x 2 orientations
        elif kind == N.OP_GATHER:
            d["work"] = float(B * workloads.gather_bytes_per_sample(feats))
        elif kind == N.OP_SCATTER:
            d["work"] = float(B * workloads.scatter_bytes_per_sample(feats))
        elif kind == N.OP_ZERO:
            d["work"] = float(i64(h))
            d["name"] = f"zero[{i64(h) / 1e6:.1f}MB]"
        elif kind == N.OP_ADAM:
            d["work"] = 28.0 * i64(h)          # read p, g, m, v; write p, m, v (fp32)
            d["name"] = f"adam[{i64(h) / 1e6:.2f}M]"
        elif kind == N.OP_ADAM_ROWS:
            # per looked-up row: read p, g, m, v and write p, m, v (+ zero g) = 32 B per element, + index and bookkeeping
            d["work"] = float(sum(B * (32.0 * int(r["i"][5]) + 16.0) for r in subs))
            d["name"] = "adam_rows[%s]" % ("catch-up" if int(h["i"][1]) == 0 else "update")
        elif kind == N.OP_ADAM_FLUSH:
            d["work"] = float(sum(24.0 * i64(r) * int(r["i"][5]) for r in subs))
        out[i] = d
        i += 1 + ns
    return out


def saturated_gather_scatter(feats, dev, pk, lookups_log2=18, min_table_bytes=1.5e9):
    """K1 / K2 through the raw C ABI at a size that fills the machine: 2^18 batch rows x every sparse field of the workload
    (~6 M lookups per launch), on tables scaled up until they hold >= 1.5 GB (an order of magnitude beyond the 126 MB L2, so
    the rows really come from HBM), fresh random indices per launch, CUDA events on the launch stream."""
    import ctypes
    from scenario_wise_rec_b200 import _native as N
    sp = [(v, d) for _, k, v, d in feats if k == "sparse"]
    nd = sum(1 for _, k, _, _ in feats if k == "dense")
    E = sp[0][1]
    if any(d != E for _, d in sp):
        return None
    scale = max(1, int(np.ceil(min_table_bytes / sum(v * d * 4 for v, d in sp))))
    sp = [(v * scale if v >= 1000 else v, d) for v, d in sp]        # tiny fields stay tiny (they live in cache anyway)
    tbytes = sum(v * d * 4 for v, d in sp)
    Bs = 1 << lookups_log2
    g = torch.Generator(device=dev).manual_seed(17)
    tables = [torch.randn(v, E, device=dev, generator=g) for v, _ in sp]
    gtabs = [torch.zeros(v, E, device=dev) for v, _ in sp]
    sets = [[torch.randint(0, v, (Bs,), device=dev, generator=g) for v, _ in sp] for _ in range(4)]
    dense = [torch.rand(Bs, device=dev, generator=g) for _ in range(nd)]
    IN = len(sp) * E + nd
    out = torch.empty(Bs, IN, device=dev)
    oob = torch.zeros(2, dtype=torch.int32, device=dev)
    arr = lambda ts: (ctypes.c_void_p * max(len(ts), 1))(*[t.data_ptr() for t in ts])      # noqa: E731
    vocab = (ctypes.c_int64 * len(sp))(*[v for v, _ in sp])
    idt = (ctypes.c_int32 * len(sp))(*[N.DT_I64] * len(sp))
    ddt = (ctypes.c_int32 * max(nd, 1))(*[N.DT_F32] * nd)
    st = torch.cuda.current_stream().cuda_stream
    L = N.lib()

    def gather(i):
        N.check(L.swr_embedding_gather_fwd(arr(tables), vocab, arr(sets[i % 4]), idt, arr(dense) if nd else None, ddt, out.data_ptr(),
                                           IN, Bs, len(sp), E, nd, oob.data_ptr(), st), "gather")

    def scatter(i):
        N.check(L.swr_embedding_scatter_bwd(out.data_ptr(), IN, Bs, arr(sets[i % 4]), idt, arr(gtabs), vocab, len(sp), E, st), "scatter")

    res = {"table_bytes": tbytes, "note": "tables scaled x%d to %.2f GB (L2 is 126 MB)" % (scale, tbytes / 1e9)}
    for name, fn, per in (("gather", gather, workloads.gather_bytes_per_sample(feats)), ("scatter", scatter, workloads.scatter_bytes_per_sample(feats))):
        for i in range(3):
            fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(10):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gbps = per * Bs / ms / 1e6
        res[name] = {"rows": Bs, "lookups": Bs * len(sp), "ms": ms, "GBps": gbps, "frac_hbm": gbps / pk["hbm"],
                     "traffic": ncu_traffic(name + "_saturated")}
    del tables, gtabs, sets, out
    torch.cuda.empty_cache()
    return res


def ncu_traffic(op_name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the op's kernel, from the committed ncu --set full
    capture (profiles/ncu_traffic.json, written by tools/summarize_profiles.py); None if that op was not captured."""
    f = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(f):
        return None
    try:
        return json.load(open(f)).get(op_name)
    except Exception:
        return None


def measure_scopes(trainer, fs, model, devb, host, B, dev, iters=100):
    """SURVEY.md 8d scopes beside the full train step: (ii) forward + BCELoss + backward without the optimizer, (iii) the
    evaluation forward (eval-mode program: running statistics, no gradient buffers).  Each scope is captured in a CUDA graph
    and replayed on device-resident batches, CUDA-event timed."""
    from scenario_wise_rec_b200 import _native as N
    out = {}
    side = torch.cuda.Stream(dev)

    def timed_graph(body, prep=None):
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            for _ in range(3):
                body()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            body()
        torch.cuda.synchronize()
        for _ in range(5):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(iters):
            if prep is not None:
                prep(i)
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    # (ii) forward + loss + backward: the step's own records minus the optimizer ops
    recs = fs.scope_records("fwd_bwd")
    st = lambda: torch.cuda.current_stream(dev).cuda_stream      # noqa: E731

    def prep(i):
        N.memcpy_async(fs.dev_stage.data_ptr(), devb[i % len(devb)].buf.data_ptr(), fs.stage_bytes, st())
    ms = timed_graph(lambda: N.program_run(recs, fs.ptrs, st()), prep)
    fs.restore_after_scope()
    out["fwd_bwd"] = {"ms": ms, "samples_per_s": B / ms * 1e3}
    # (iii) evaluation forward
    was = model.training
    model.eval()
    xd = {k: v.to(dev) for k, v in host[0][0].items()}
    with torch.no_grad():
        runner = model._runner(xd)
        runner._bind_inputs(xd)
        ms = timed_graph(lambda: N.program_run(runner.prog.recs_fwd, runner.ptrs, st()))
    model.train(was)
    out["eval_fwd"] = {"ms": ms, "samples_per_s": B / ms * 1e3, "launches": int(runner.prog.n_launch_fwd)}
    return out


# --------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def emit(line: dict):
    """The one JSON line, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries write to stdout behind Python's back (NCCL prints its version banner there): keep file descriptor 1
    # for the JSON line alone and send everything else to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=workloads.DEFAULT_CASE, choices=sorted(workloads.CASES))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-saturated", action="store_true", help="skip the saturated K1 / K2 measurement (1.5 GB of tables)")
    ap.add_argument("--no-scopes", action="store_true", help="skip the fwd+bwd / eval-forward scopes")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--ref-budget", type=float, default=60.0, help="seconds of CPU work for --impl reference")
    ap.add_argument("--eager-budget", type=float, default=6.0, help="seconds for the eager-CUDA run of the unmodified reference")
    ap.add_argument("--shard-min-rows", type=int, default=30000,
                    help="with more than one GPU, row-shard every table with at least this many rows (0: replicate all tables)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if os.environ.get("SWR_BENCH_WATCHDOG"):     # development aid: dump every thread's Python stack if the run stalls
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["SWR_BENCH_WATCHDOG"]), repeat=False, exit=True)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch.distributed as dist
    import model_factory
    from scenario_wise_rec_b200 import _native as N
    from scenario_wise_rec_b200.trainers import CTRTrainer

    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N.check(N.lib().swr_device_check(), "swr_device_check")

    model_name, cfg, B = workloads.CASES[args.workload]
    feats = workloads.all_feature_specs(cfg)
    torch.manual_seed(0)                     # identical initial weights on every rank
    model = model_factory.build(model_name, cfg, shard_min_rows=args.shard_min_rows if world > 1 else 0)
    trainer = CTRTrainer(model, "synthetic", optimizer_params={"lr": 1e-3, "weight_decay": 1e-5}, device=str(dev))
    if world > 1:
        trainer.enable_data_parallel()
    model.train()

    NB = NB_ROTATE
    host = [workloads.make_batch(feats, B, cfg["domain_num"], seed=1000 * rank + i, pin=True) for i in range(NB)]
    fs = trainer.packer(host[0][0])          # the fused step (one CUDA graph) for this batch shape
    devb = [fs.pack(x, y, device=dev) for x, y in host]      # device-resident batches in the staging layout
    h2d = fs.stage_bytes + fs.SC

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(run, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0 = N.launch_count()
        t0 = time.perf_counter()
        e0.record()
        run(steps)
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1])
        return ms, wall, N.launch_count() - l0

    def run_resident(steps):                 # inputs already in HBM: one D2D copy + one graph launch per step
        for i in range(steps):
            trainer.train_step(devb[i % NB])

    def run_e2e(steps):                      # the public API on pinned host batches (H2D + loss D2H inside)
        trainer.train_one_epoch([host[i % NB] for i in range(steps)])

    with ClockSampler(local) as clk:
        run_resident(args.warmup)
        run_e2e(max(args.warmup, 3))
        clk.mark()
        ms, wall, _ = timed(run_resident, args.steps)
        ms_e2e, wall_e2e, _ = timed(run_e2e, args.steps)
        clk.mark()
    model.check_indices()
    value = world * B * args.steps / (ms * 1e-3)
    e2e = world * B * args.steps / (max(ms_e2e, wall_e2e) * 1e-3)
    launches = fs.n_launch * args.steps      # kernels + memsets inside each replayed graph (counted from the records)

    # live per-op device times: the same step run eagerly (no graph) with CUDA events on the launch stream
    # around every op of the record list (swr_profile_begin/end of the C ABI)
    P = min(args.steps, 20)
    fs.use_graph = False
    run_resident(2)
    torch.cuda.synchronize()
    N.profile_begin()
    run_resident(P)
    kinds, recs, opms = N.profile_end()
    fs.use_graph = True
    tabs = [(fs.recs_a, op_table(fs.recs_a, B, feats))]
    if fs.recs_b is not None:
        tabs.append((fs.recs_b, op_table(fs.recs_b, B, feats)))
    if getattr(fs, "_flush_recs", None) is not None:
        tabs.append((fs._flush_recs, op_table(fs._flush_recs, B, feats)))
    agg = {}
    for k, r, t in zip(kinds.tolist(), recs.tolist(), opms.tolist()):
        d = None
        for ti, (rl, tb) in enumerate(tabs):
            if r in tb and r < len(rl) and int(rl[r]["kind"]) == k:
                d, key = tb[r], (ti, r)
                break
        if d is None:
            d, key = {"name": str(k), "family": str(k), "bound": "hbm", "work": None}, (-1, k)
        agg.setdefault(key, [d, []])[1].append(t)
    pk = peaks()
    # per op: mean device time per STEP (an op that runs every 32nd step, the lazy-Adam flush, is amortised)
    ops = [{"op": d["name"], "family": d["family"], "ms": float(np.sum(ts)) / P, "bound": d["bound"], "work": d["work"],
            "runs_per_step": len(ts) / P} for d, ts in agg.values()]
    ops.sort(key=lambda o: -o["ms"])
    step_ms = max(sum(o["ms"] for o in ops), 1e-9)
    fams = {}
    for o in ops:
        f = fams.setdefault(o["family"], {"ms": 0.0, "work": 0.0, "bound": o["bound"], "launches": 0.0, "ops": []})
        f["ms"] += o["ms"]
        f["work"] += (o["work"] or 0.0) * o["runs_per_step"]
        f["launches"] += o["runs_per_step"]
        f["ops"].append(o["op"])

    def roofline_of(name, f):
        if f["bound"] == "tensor":
            ach = f["work"] / (f["ms"] * 1e-3) / 1e12
            r = {"bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"],
                 "note": "algorithmic fp32 FLOPs (2 B K N per layer) over the summed device time of every launch of the kernel; the "
                         "3xTF32 arithmetic issues 3x that on a pipe with half the bf16 rate, so 1/6 of the peak is this arithmetic's ceiling"}
        else:
            ach = f["work"] / (f["ms"] * 1e-3) / 1e9
            r = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"]}
        r.update(kernel=name, launches_per_step=f["launches"], kernel_ms_per_step=f["ms"], share_of_step=f["ms"] / step_ms,
                 peak_source=pk["source"], traffic=ncu_traffic(name), ops=f["ops"][:8])
        return r
    # the dominant kernel FUNCTION of the hot path (gather / FC stack / scatter), summed over its launches; the optimizer
    # (a section-8f op) is reported beside it
    hot = {k: v for k, v in fams.items() if k in ("fc_fwd", "fc_dgrad", "fc_wgrad", "gather", "scatter") and v["work"]}
    top_name = max(hot, key=lambda k: hot[k]["ms"])
    roof = roofline_of(top_name, hot[top_name])
    roof_by_kernel = {k: {kk: vv for kk, vv in roofline_of(k, v).items() if kk in ("bound", "achieved", "peak", "unit", "frac", "kernel_ms_per_step",
                                                                                  "launches_per_step", "share_of_step", "traffic")}
                      for k, v in hot.items()}
    roof_opt = roofline_of("optimizer", fams["optimizer"]) if "optimizer" in fams and fams["optimizer"]["work"] else None
    sat = saturated_gather_scatter(feats, dev, pk) if (world == 1 and not args.no_saturated) else None
    gather = next((o for o in ops if o["op"] == "gather"), None)
    scatter = next((o for o in ops if o["op"] == "scatter"), None)
    scopes = measure_scopes(trainer, fs, model, devb, host, B, dev) if (world == 1 and not args.no_scopes) else None
    parity = None
    if world > 1:
        import dist_parity
        parity = dist_parity.run(dev)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": base_config(args, world),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 256, "ms_per_step": max(ms_e2e, wall_e2e) / args.steps,
                "api": "CTRTrainer.train_one_epoch(list of pinned host (x_dict, y) batches)"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
        "roofline": roof,
        "roofline_by_kernel": roof_by_kernel,
        "roofline_optimizer": roof_opt,
        "scopes": scopes,
        "dist_parity": parity,
        "ops_ms": [{"op": o["op"], "ms": round(o["ms"], 5)} for o in ops[:16]],
        "ops_ms_total": round(step_ms, 5),
        "gather": None if not gather else {"ms": gather["ms"], "GBps": gather["work"] / gather["ms"] / 1e6, "frac_hbm": gather["work"] / gather["ms"] / 1e6 / pk["hbm"]},
        "scatter": None if not scatter else {"ms": scatter["ms"], "GBps": scatter["work"] / scatter["ms"] / 1e6, "frac_hbm": scatter["work"] / scatter["ms"] / 1e6 / pk["hbm"]},
        "gather_scatter_saturated": sat,
        "fc_arithmetic": {0: "fp32 FFMA", 1: "tcgen05 3xTF32 (all FC layers)", 2: "tcgen05 3xTF32 (wide FC layers) + fp32 FFMA (narrow)"}[N.get_fc_mode()],
        "wall_ms_per_step": wall / args.steps,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the reference itself (unmodified, baseline/_ref) on this box: eager PyTorch on cuda:0 -- the bar SURVEY.md 8d
        # sets -- and on the host cores; both through its own CTRTrainer.train_one_epoch with host batches
        eager = reference_arm(args, str(dev), 200, 3, budget_s=args.eager_budget)
        if eager is not None:
            line["eager_cuda_baseline"] = {k: eager[k] for k in ("value", "unit", "kind", "ms_per_step", "device", "path", "sample")}
        torch.set_num_threads(os.cpu_count() or 1)
        cpu = reference_arm(args, "cpu", 10_000, 2, budget_s=args.cpu_budget)
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        emit(line)
    if world > 1:
        # the captured step graphs hold NCCL work on the communicator: destroy_process_group() would wait on it
        # forever.  Everybody is past the last collective once the barrier returns; leave without the teardown.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
