"""One CUDA graph per training step.

``CTRTrainer.train_step`` (reference loop: trainers/ctr_trainer.py:67-73) maps onto

    host   pack the batch columns into one pinned staging buffer          (no per-column H2D copies)
    H2D    ONE cudaMemcpyAsync of the packed batch + 64 B of per-step scalars
    graph  forward program -> BCELoss fwd/grad -> zero gradient arenas -> backward program
           (-> NCCL all-reduce of the arenas under data parallel) -> Adam over the flat arenas
           -> D2H copy of the loss ring
    host   the loss of step k is read from the pinned ring when the caller asks for it

Parameters stay ordinary ``nn.Parameter`` objects under the reference's names; their storage is
re-pointed into flat arenas (one for the tower / expert / gate weights, one for the embedding
tables) laid out like the program's gradient arenas, so the optimizer is a single streaming
kernel per arena.  ``optimizer.state`` holds views of the flat moment arenas: ``state_dict()`` /
``load_state_dict()`` of model and optimizer keep working.  Semantics are torch.optim.Adam's
(dense gradients, L2 weight decay on every row of every table, every step).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import os

import numpy as np
import torch

from .. import _native as N
from ..program import CudaRunner, Program, ProgramBuilder

_NP = {torch.int8: np.int8, torch.int16: np.int16, torch.int32: np.int32, torch.int64: np.int64, torch.uint8: np.uint8,
       torch.float16: np.float16, torch.float32: np.float32, torch.float64: np.float64, torch.bool: np.bool_}
RING = 64          # loss ring / per-step scalar ring
NSTAGE = 4         # rotating pinned staging buffers


def adam_supported(opt) -> bool:
    if type(opt) is not torch.optim.Adam or len(opt.param_groups) != 1:
        return False
    g = opt.param_groups[0]
    return not (g.get("amsgrad") or g.get("maximize") or g.get("differentiable") or g.get("capturable"))


def _align(n: int, a: int = 16) -> int:
    return (n + a - 1) // a * a


class FlatArenas:
    """Flat parameter / gradient / Adam-moment storage shared by every program of one model.

    Arenas: ``dense`` (tower / expert / gate weights), ``emb`` (replicated tables), ``shard`` (this rank's rows of
    row-sharded tables) hold parameters, gradients and both moments; ``virt`` holds only the gradients of the
    virtual tables of the embedding exchange (parallel.py), which are per-step buffers, not parameters."""

    def __init__(self, prog: Program, optimizer, device):
        self.device = device
        entries = [(p, a, off, n) for p, (a, off, n) in zip(prog.params, prog.param_arena)]
        self.size = {k: max(v, 4) for k, v in prog.arena_size.items()}
        # shards of the NCCL-exchange path (virtual tables) follow the shard tables the program itself addresses (peer path)
        off = prog.arena_size.get("shard", 0)
        self.shards = []
        for f in prog.virtual_fields:
            if f.shard.requires_grad and all(f.shard is not q for q, *_ in self.shards):
                self.shards.append((f.shard, "shard", off, f.shard.numel()))
                off += _align(f.shard.numel(), 4)
        self.size["shard"] = max(off, 4)
        self.layout = [(id(p), a, o, n) for p, a, o, n in entries]
        z = lambda k: torch.zeros(self.size[k], dtype=torch.float32, device=device)      # noqa: E731
        # dense + emb gradients are contiguous: data-parallel replicas sum them with ONE all-reduce
        self.g_repl = torch.zeros(self.size["dense"] + self.size["emb"], dtype=torch.float32, device=device)
        self.g = {k: z(k) for k in self.size if k not in ("dense", "emb")}
        self.g["dense"] = self.g_repl[:self.size["dense"]]
        self.g["emb"] = self.g_repl[self.size["dense"]:]
        self.opt_arenas = [k for k in self.size if k != "virt"]
        self.p = {k: z(k) for k in self.opt_arenas}
        # peer path: the shard parameters and their gradients live in memory every rank of the node can address
        self.peer = None
        self.peer_tables = {}
        if prog.peer_tabs:
            from ..parallel import PeerArena
            self.peer = {"p": PeerArena(self.size["shard"], device), "g": PeerArena(self.size["shard"], device)}
            self.p["shard"], self.g["shard"] = self.peer["p"].tensor, self.peer["g"].tensor
            for p_, a, o, n in entries:
                if a == "shard":
                    self.peer_tables[("peers_p", id(p_))] = self.peer["p"].peer_table(o)
                    self.peer_tables[("peers_g", id(p_))] = self.peer["g"].peer_table(o)
            # sharded tables the program only reads (no gradient reaches them: PPNet's agnostic tables, ppnet.py:54) still
            # have to be addressable by every rank; they get their own arena, outside the range the optimizer sweeps
            frozen = [t for t, _v, _w in prog.peer_tabs if ("peers_p", id(t)) not in self.peer_tables]
            if frozen:
                total = sum(_align(t.numel(), 4) for t in frozen)
                self.peer["f"] = PeerArena(max(total, 4), device)
                o = 0
                with torch.no_grad():
                    for t in frozen:
                        n = t.numel()
                        view = self.peer["f"].tensor[o:o + n]
                        view.copy_(t.data.reshape(-1))
                        t.data = view.view(t.shape)
                        self.peer_tables[("peers_p", id(t))] = self.peer["f"].peer_table(o)
                        o += _align(n, 4)
                import torch.distributed as _dist
                _dist.barrier()
        self.m = {k: z(k) for k in self.opt_arenas}
        self.v = {k: z(k) for k in self.opt_arenas}
        self.params = []
        self._entries = [e for e in entries + self.shards if e[1] != "virt"]
        step = 0
        with torch.no_grad():
            for p, a, off, n in entries + self.shards:
                if a == "virt":
                    continue
                if p.device != device or p.dtype != torch.float32:
                    raise RuntimeError("parameters must be float32 on the trainer's device")
                flat = self.p[a][off:off + n]
                flat.copy_(p.data.reshape(-1))
                p.data = flat.view(p.shape)
                p.grad = self.g[a][off:off + n].view(p.shape)
                st = optimizer.state.get(p)
                if st and "exp_avg" in st:
                    self.m[a][off:off + n].copy_(st["exp_avg"].reshape(-1))
                    self.v[a][off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                    step = max(step, int(st["step"]))
                optimizer.state[p] = {"step": torch.tensor(float(step)), "exp_avg": self.m[a][off:off + n].view(p.shape),
                                      "exp_avg_sq": self.v[a][off:off + n].view(p.shape)}
                self.params.append(p)
        self.step = step
        self.optimizer = optimizer
        self._ptr0 = [p.data_ptr() for p in self.params[:1] + self.params[-1:]]
        # row-lazy Adam over the replicated tables (single process; SWR_LAZY_ADAM=0 keeps the dense sweep)
        self.lazy = None
        import torch.distributed as dist
        single = not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)
        tables = [(p, off, int(p.shape[0]), int(p.shape[1])) for p, a, off, n in entries if a == "emb" and p.dim() == 2]
        if single and tables and os.environ.get("SWR_LAZY_ADAM", "1") != "0" and len(tables) == sum(1 for _p, a, *_ in entries if a == "emb"):
            self.lazy = LazyTables(self, tables)
        # peer mode: the shard arena is swept densely while that is cheap (cfg2 on 8 GPUs: 87 MB per step) and updated
        # row-lazily once the sweep would dominate (the 85 M-row item table: 4.8 GB per step and rank)
        shard_tabs = [(p, off, int(p.shape[0]), int(p.shape[1])) for p, a, off, n in entries if a == "shard" and p.dim() == 2]
        min_elems = int(float(os.environ.get("SWR_LAZY_SHARD_MIN", "32e6")))
        if (self.peer is not None and shard_tabs and os.environ.get("SWR_LAZY_ADAM", "1") != "0" and not self.shards
                and sum(v * e for _p, _o, v, e in shard_tabs) >= min_elems):
            self.lazy = LazyTables(self, shard_tabs, arena="shard")

    def shard_grad(self, shard_param) -> torch.Tensor:
        for p, _a, off, n in self.shards:
            if p is shard_param:
                return self.g["shard"][off:off + n].view(p.shape)
        raise KeyError("not a sharded table of this model")

    def valid(self) -> bool:
        """False after ``model.to()`` / ``load`` replaced parameter storage behind our back."""
        return [p.data_ptr() for p in self.params[:1] + self.params[-1:]] == self._ptr0

    def matches(self, prog: Program) -> bool:
        cur = [(id(p), a, off, n) for p, (a, off, n) in zip(prog.params, prog.param_arena)]
        strip = lambda lst: [e for e in lst if e[1] != "virt"]      # noqa: E731  (virtual tables are per batch size)
        return strip(cur) == strip(self.layout)

    def reload_optimizer_state(self):
        """After ``optimizer.load_state_dict()``: torch replaced the state tensors by fresh copies; copy them back into the
        flat moment arenas and re-point ``optimizer.state`` at the arena views (the fused Adam kernels read the arenas)."""
        step = 0
        with torch.no_grad():
            layout = {id(p): (a, off, n) for p, a, off, n in self._entries}
            for p in self.params:
                a, off, n = layout[id(p)]
                st = self.optimizer.state.get(p)
                if st and "exp_avg" in st and st["exp_avg"].data_ptr() != self.m[a][off:off + n].data_ptr():
                    self.m[a][off:off + n].copy_(st["exp_avg"].reshape(-1))
                    self.v[a][off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                if st and "step" in st:
                    step = max(step, int(st["step"]))
                self.optimizer.state[p] = {"step": torch.tensor(float(step)), "exp_avg": self.m[a][off:off + n].view(p.shape),
                                           "exp_avg_sq": self.v[a][off:off + n].view(p.shape)}
        self.step = step
        if self.lazy is not None:       # every row is as current as the loaded state says
            self.lazy.last.fill_(step)
            self.lazy.base = step + 1
            self.lazy.dirty = False

    def flush_lazy(self):
        """Bring every table row to the current step (no-op when nothing is postponed)."""
        if self.lazy is not None and self.lazy.dirty and self.flush_fn is not None:
            self.flush_fn()

    flush_fn = None

    def publish_step(self):
        """Write the step counter into ``optimizer.state`` (kept lazily: one Python loop per epoch, not per step)."""
        self.flush_lazy()
        for p in self.params:
            self.optimizer.state[p]["step"].fill_(float(self.step))


class LazyTables:
    """State of the row-lazy Adam over the embedding tables of the ``emb`` arena (csrc/swr_train.cu).

    ``torch.optim.Adam`` on dense table gradients (the reference, ctr_trainer.py:50-52,73) updates every row of every
    table at every step; a row outside the batch has a zero gradient and still moves (weight decay, decaying moments).
    Those zero-gradient updates depend only on the row itself and on the step's scalars, so they are postponed and
    replayed -- with the dense kernel's own arithmetic, bit for bit -- when the row is next looked up, every
    ``flush_every`` steps, and whenever something outside the fused step reads the parameters (state_dict, evaluation,
    the end of ``train_one_epoch``).  Per step the optimizer then touches the batch's rows instead of sweeping
    O(vocab) bytes, and the dense gradient arena needs no memset (updated rows are zeroed as they are consumed)."""

    HIST_CAP = 1 << 20          # steps of scalars kept on the device (16 B each); a flush re-bases the table

    def __init__(self, flat: "FlatArenas", tables, arena: str = "emb"):
        dev = flat.device
        self.flat = flat
        self.arena = arena                        # "emb" (replicated tables, single process) or "shard" (this rank's shards, peer mode)
        self.tables = tables                      # [(param, offset, rows, E)] inside that arena
        rows = sum(v for _p, _o, v, _e in tables)
        self.row_off = {}
        o = 0
        for p_, _off, v, _e in tables:
            self.row_off[id(p_)] = o
            o += v
        self.base = flat.step + 1                 # history index 0 = the first step taken from now on
        self.last = torch.full((max(rows, 1),), flat.step, dtype=torch.int32, device=dev)
        self.claim = torch.zeros(max(rows, 1), dtype=torch.int32, device=dev)
        self.hist = torch.zeros(4 * self.HIST_CAP, dtype=torch.float32, device=dev)
        self.flush_every = max(1, int(os.environ.get("SWR_LAZY_FLUSH", "64")))      # measured at cfg2: 8 -> 0.637, 32 -> 0.610, 64 -> 0.606 ms / step
        self.dirty = False
        self._flush_recs = None
        self._flush_ptrs = None

    def field_rec(self, ptrs_base: dict, table_param, idx_slot: int, dtype_code: int):
        """(pointer list, ints) of one lookup column of ``table_param`` for an ADAM_ROWS sub-record."""
        for p_, off, v, e in self.tables:
            if p_ is table_param:
                ro = self.row_off[id(p_)]
                return ([ptrs_base["p"] + 4 * off, ptrs_base["g"] + 4 * off, ptrs_base["m"] + 4 * off, ptrs_base["v"] + 4 * off,
                         self.last.data_ptr() + 4 * ro, self.claim.data_ptr() + 4 * ro], v, e)
        raise KeyError("table is not in the emb arena")


class LossHandle:
    """The loss of one fused step; ``item()`` waits for that step only."""

    def __init__(self, owner: "FusedTrainStep", k: int):
        self.owner, self.k = owner, k
        self._v: Optional[float] = None

    def item(self) -> float:
        if self._v is None:
            self._v = self.owner._read_loss(self.k)
        return self._v

    __float__ = item

    def tensor(self) -> torch.Tensor:
        return torch.tensor(self.item())


class PackedBatch:
    """A batch already laid out in the staging format (device or pinned host): one copy per step."""

    def __init__(self, buf: torch.Tensor, B: int, key):
        self.buf, self.B, self.key = buf, B, key


class FusedTrainStep:
    def __init__(self, model, optimizer, x: Dict[str, torch.Tensor], device: torch.device, flat: Optional[FlatArenas] = None,
                 grad_sync=None, use_graph: bool = True):
        N.lib()
        self.model, self.optimizer, self.device = model, optimizer, device
        cols = model._columns()
        self.cols = cols
        self.B = int(x[cols[0]].shape[0])
        self.dts = {c: x[c].dtype for c in cols}
        self.key = (self.B, tuple(self.dts.values()))
        b = ProgramBuilder(self.B, True)
        b.exchange = model._get_exchange()
        import torch.distributed as dist
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        # row-sharded tables: K1 / K2 address the owners' shards over NVLink peer pointers (SWR_SHARD_P2P=0: NCCL exchange)
        b.p2p = bool(multi and b.exchange is not None and os.environ.get("SWR_SHARD_P2P", "1") != "0")
        model._lower(b, dict(self.dts))
        prog = b.finish()
        self.exchange = b.exchange if prog.virtual_fields else None
        if prog.out_slot < 0:
            raise RuntimeError("the model's program has no head output")
        self.prog = prog
        if flat is None:
            flat = FlatArenas(prog, optimizer, device)
            model._programs = {}            # cached runners hold the old parameter pointers
        elif not flat.matches(prog):
            raise RuntimeError("gradient-arena layout changed between programs of one model")
        self.flat = flat
        self.grad_sync = grad_sync
        self.runner = CudaRunner(prog, device)

        # ---- staging layout: [columns..., labels f32] ; scalars live in their own small ring
        self.layout = {}
        off = 0
        vfields = {f.vcol: f for f in prog.virtual_fields}
        for name in prog.inputs:
            if name == "__grad_out__" or name in vfields:
                continue
            dt = self.dts[name]
            nbytes = self.B * torch.empty((), dtype=dt).element_size()
            self.layout[name] = (off, nbytes, _NP[dt])
            off = _align(off + nbytes)
        # the index column of a row-sharded field is not an input of the device program (the program reads the
        # exchanged rows through the field's virtual table), but the exchange itself needs it on the device
        for f in prog.virtual_fields:
            if f.name not in self.layout:
                dt = self.dts[f.name]
                nbytes = self.B * torch.empty((), dtype=dt).element_size()
                self.layout[f.name] = (off, nbytes, _NP[dt])
                off = _align(off + nbytes)
        self.y_off = off
        off = _align(off + 4 * self.B)
        self.stage_bytes = off
        self.dev_stage = torch.zeros(off, dtype=torch.uint8, device=device)
        self.host_stage = [torch.zeros(off, dtype=torch.uint8).pin_memory() for _ in range(NSTAGE)]
        self.host_np = [t.numpy() for t in self.host_stage]
        self.stage_ev = [torch.cuda.Event() for _ in range(NSTAGE)]
        # per-step scalars: 8 floats of Adam hyper-parameters + 4 ints of control
        self.SC = 48
        self.dev_scal = torch.zeros(self.SC, dtype=torch.uint8, device=device)
        self.host_scal = torch.zeros(RING * self.SC, dtype=torch.uint8).pin_memory()
        self.host_scal_np = self.host_scal.numpy()
        self.loss_ring_dev = torch.zeros(RING, dtype=torch.float32, device=device)
        self.loss_ring_host = torch.zeros(RING, dtype=torch.float32).pin_memory()
        self.done_ev = [torch.cuda.Event() for _ in range(RING)]
        self.gout = torch.zeros(self.B, dtype=torch.float32, device=device)
        self.k = 0
        self._launched = [-1] * RING
        self._hold = [None] * RING          # caller buffers an in-flight async copy still reads

        # gradients of the virtual tables are per batch size: this step owns them
        self.g = dict(flat.g)
        self.g["virt"] = torch.zeros(max(prog.arena_size.get("virt", 0), 4), dtype=torch.float32, device=device)
        gsize = {a: self.g[a].numel() for a in self.g}
        # ---- slot table: the program's slots + trainer extras
        ns = len(prog.slot_desc)
        X = {"label": ns, "loss": ns + 1, "ctrl": ns + 2, "hyper": ns + 3}
        nxt = ns + 4
        self.arena_slots = {}
        for a in flat.size:
            self.arena_slots[a] = (nxt, nxt + 1, nxt + 2, nxt + 3)       # p, g, m, v
            nxt += 4
        ptrs = np.zeros(nxt, dtype=np.uint64)
        ptrs[:ns] = self.runner.ptrs
        base = self.dev_stage.data_ptr()
        for name, slot in prog.inputs.items():
            if name == "__grad_out__":
                ptrs[slot] = self.gout.data_ptr()
            elif name in vfields:
                ptrs[slot] = vfields[name].vidx.data_ptr()
            else:
                ptrs[slot] = base + self.layout[name][0]
        for i, d in enumerate(prog.slot_desc):
            if d[0] == "grad":
                ptrs[i] = self.g[d[1]].data_ptr() + 4 * d[2]
            elif d[0] == "special" and isinstance(d[1], tuple):      # device table of every rank's shard address
                ptrs[i] = flat.peer_tables[(d[1][0], id(d[1][1]))].data_ptr()
        ptrs[X["label"]] = base + self.y_off
        ptrs[X["loss"]] = self.loss_ring_dev.data_ptr()
        ptrs[X["hyper"]] = self.dev_scal.data_ptr()
        ptrs[X["ctrl"]] = self.dev_scal.data_ptr() + 32
        for a, (sp, sg, sm, sv) in self.arena_slots.items():
            ptrs[sg] = self.g[a].data_ptr()
            if a in flat.opt_arenas:
                ptrs[sp], ptrs[sm], ptrs[sv] = flat.p[a].data_ptr(), flat.m[a].data_ptr(), flat.v[a].data_ptr()
        self.ptrs = ptrs

        # ---- record lists
        def rec(kind, ints=(), slots=(), floats=()):
            r = np.zeros((), dtype=N.REC_DTYPE)
            r["kind"] = kind
            r["s"][:] = -1
            for i, v in enumerate(ints):
                r["i"][i] = v
            for i, v in enumerate(slots):
                r["s"][i] = v
            for i, v in enumerate(floats):
                r["f"][i] = v
            return r

        split = ProgramBuilder._split64
        # Data-parallel table gradients as rows instead of a dense all-reduce (see _setup_sparse_sync): every rank then
        # scales its loss gradient by 1 / world, so that cross-rank *sums* are the global-batch mean.
        self.p2p = flat.peer is not None
        self.sparse_sync = None if self.p2p else self._setup_sparse_sync(prog, ptrs)
        gscale = 1.0 / self.sparse_sync["world"] if self.sparse_sync else 1.0
        if self.p2p:
            # every rank's loss gradient is scaled by 1 / world: the rows added into the owners' shards and the summed
            # replicated gradients are then those of the global-batch mean
            gscale = 1.0 / dist.get_world_size()
            self._bar = torch.zeros(1, dtype=torch.float32, device=device)
        bce = rec(N.OP_BCE, [self.B, N.DT_F32, RING], [prog.out_slot, X["label"], prog.gout_slot, X["loss"], X["ctrl"]], [gscale])
        lazy = None
        if flat.lazy is not None:
            if flat.lazy.arena == "emb" and grad_sync is None and self.exchange is None and not self.p2p:
                lazy = flat.lazy
            elif flat.lazy.arena == "shard" and self.p2p:
                lazy = flat.lazy
        self.lazy = lazy
        pre, post = [], []
        self.g_stage = None
        if lazy is not None:
            if lazy.arena == "emb":
                self.g["emb"].zero_()        # invariant of the lazy update: the dense table-gradient arena is all zero between steps
            extra = []          # extra slot pointers appended behind the trainer slots
            ar = lazy.arena
            base = {"p": flat.p[ar].data_ptr(), "g": self.g[ar].data_ptr(), "m": flat.m[ar].data_ptr(), "v": flat.v[ar].data_ptr()}
            by_grad_slot = {}
            for p_, (a, off_, n_) in zip(prog.params, prog.param_arena):
                if a == ar:
                    for i, d in enumerate(prog.slot_desc):
                        if d[0] == "grad" and d[1] == ar and d[2] == off_:
                            by_grad_slot[i] = p_
            # one sub-record per index column that reads a lazily updated table (taken from the K2 scatter records); in
            # peer mode once per rank: the owner of a row replays / updates it for every rank's lookups, which it sees in
            # the all-gathered staging buffers
            world = dist.get_world_size() if self.p2p else 1
            rank = dist.get_rank() if self.p2p else 0
            if self.p2p:
                self.g_stage = torch.zeros(world, self.stage_bytes, dtype=torch.uint8, device=device)
            input_name = {slot: name for name, slot in prog.inputs.items()}
            subs = []
            recs_b_ = prog.recs_bwd
            i = 0
            while i < len(recs_b_):
                ns = int(recs_b_[i]["n_sub"])
                if int(recs_b_[i]["kind"]) == N.OP_SCATTER:
                    for r in recs_b_[i + 1:i + 1 + ns]:
                        tab = by_grad_slot.get(int(r["s"][0]))
                        if tab is None:
                            continue             # a table of another arena (small replicated tables in peer mode)
                        plist, rows_, E = lazy.field_rec(base, tab, int(r["s"][1]), int(r["i"][2]))
                        sl = []
                        for ptr in plist:
                            extra.append(ptr)
                            sl.append(len(ptrs) + len(extra) - 1)
                        vocab_full = (int(r["i"][0]) & 0xFFFFFFFF) | (int(r["i"][1]) << 32)
                        for q in range(world):
                            if self.p2p:         # rank q's copy of the index column inside the gathered staging buffers
                                extra.append(self.g_stage[q].data_ptr() + self.layout[input_name[int(r["s"][1])]][0])
                                idx_slot = len(ptrs) + len(extra) - 1
                            else:
                                idx_slot = int(r["s"][1])
                            sub = rec(N.OP_GROUP, [*split(vocab_full if self.p2p else rows_), int(r["i"][2]), 0, 0, E, world if self.p2p else 0, rank],
                                      sl + [idx_slot])
                            subs.append(sub)
                i += 1 + ns
            hist_slot = len(ptrs) + len(extra)
            extra.append(lazy.hist.data_ptr())
            ptrs = np.concatenate([ptrs, np.array(extra, dtype=np.uint64)])
            self.ptrs = ptrs
            hdr = lambda phase: rec(N.OP_ADAM_ROWS, [self.B, phase], [X["hyper"], X["ctrl"], hist_slot])      # noqa: E731
            for h in (hdr(0), hdr(1)):
                h["n_sub"] = len(subs)
            pre = [hdr(0)] + subs
            post = [hdr(1)] + subs
            pre[0]["n_sub"] = post[0]["n_sub"] = len(subs)
            # flush: every distinct table once, over its local rows
            fsubs, seen = [], set()
            rows_of = {base["p"] + 4 * off: v for _p, off, v, _e in lazy.tables}
            for sub in subs:
                key = int(ptrs[int(sub["s"][0])])
                if key not in seen:
                    seen.add(key)
                    fs_ = sub.copy()
                    fs_["i"][0], fs_["i"][1] = split(rows_of[key])
                    fs_["i"][6] = 0
                    fsubs.append(fs_)
            fh = rec(N.OP_ADAM_FLUSH, [self.B, 0], [X["hyper"], X["ctrl"], hist_slot])
            fh["n_sub"] = len(fsubs)
            self._flush_recs = stack_ = np.stack([fh] + fsubs).astype(N.REC_DTYPE)
            flat.flush_fn = self._flush_lazy
        # peer mode: a shard's gradient is written by every rank, so it is zeroed by its own Adam sweep (zero_grad flag)
        # at the end of the step, before the barrier that opens the next one -- not in the middle of this one
        lazy_arena = lazy.arena if lazy is not None else None
        zeros = [rec(N.OP_ZERO, split(4 * gsize[a]), [self.arena_slots[a][1]]) for a in flat.size
                 if a != lazy_arena and not (self.p2p and a == "shard")]
        adams = [rec(N.OP_ADAM, [*split(flat.size[a]), 1 if (self.p2p and a == "shard") else 0], [*self.arena_slots[a], X["hyper"]])
                 for a in flat.opt_arenas if (flat.size[a] > 4 or a == "dense") and a != lazy_arena]
        stack = lambda lst: np.stack(lst).astype(N.REC_DTYPE)      # noqa: E731
        # peer mode: the catch-up runs between the all-gather of the index columns and the barrier that opens the gather
        self.recs_pre = stack(pre) if (pre and self.p2p) else None
        parts = ([stack(pre)] if (pre and not self.p2p) else []) + [prog.recs_fwd, stack([bce] + zeros), prog.recs_bwd]
        self.recs_a = np.concatenate(parts)
        self.recs_b = stack(adams + post)
        if grad_sync is None and self.exchange is None and not self.p2p:
            self.recs_a, self.recs_b = np.concatenate([self.recs_a, self.recs_b]), None
        # typed views of the staged index columns of row-sharded fields + the gradient views the exchange routes
        self._vf = []
        for f in prog.virtual_fields:
            o, nb, _npdt = self.layout[f.name]
            idx = self.dev_stage[o:o + nb].view(self.dts[f.name])
            gv = None
            for p_, (a, off_, n_) in zip(prog.params, prog.param_arena):
                if p_ is f.virt:
                    gv = self.g["virt"][off_:off_ + n_].view(f.virt.shape)
            self._vf.append((f, idx, gv, flat.shard_grad(f.shard) if f.shard.requires_grad else None))
        count = lambda recs: 0 if recs is None else sum(1 for r in recs if int(r["kind"]) != N.OP_GROUP)      # noqa: E731
        self.n_launch = count(self.recs_a) + count(self.recs_b) + count(self.recs_pre)
        self.graph = None
        self.use_graph = use_graph
        self._warm = 0

    # ---- data-parallel table gradients as rows -------------------------------------------------------------------
    def _setup_sparse_sync(self, prog, ptrs):
        """Replicated tables under data parallelism: the dense gradient of a table is non-zero on at most B rows, so
        instead of all-reducing the whole table-gradient arena (100 MB per step at cfg2) every rank all-gathers the
        gradient of the embedding-layer output [B, IN] and the staged index columns (7 MB) and replays the K2 scatter
        of the backward program on each peer's block.  Same sums as the dense all-reduce, in a different order.
        Returns None when it does not apply (single process, row-sharded fields, tiny tables, or SWR_DP_SPARSE=0)."""
        import torch.distributed as dist
        if self.grad_sync is None or prog.virtual_fields or not (dist.is_available() and dist.is_initialized()):
            return None
        mode = os.environ.get("SWR_DP_SPARSE", "auto")
        if mode == "0":
            return None
        world, rank = dist.get_world_size(), dist.get_rank()
        recs = prog.recs_bwd
        blocks = []          # (header index, number of sub-records)
        i = 0
        while i < len(recs):
            ns = int(recs[i]["n_sub"])
            if int(recs[i]["kind"]) == N.OP_SCATTER:
                blocks.append((i, ns))
            i += 1 + ns
        if world < 2 or not blocks or "emb" not in self.flat.g:
            return None
        emb_bytes = 4 * self.flat.g["emb"].numel()
        dz_views, scat = [], []
        for hi, ns in blocks:
            d = prog.slot_desc[int(recs[hi]["s"][0])]
            if d[0] != "ws32":
                return None
            dz_views.append((int(recs[hi]["s"][0]), self.runner.ws32[d[1]:d[1] + d[2]]))
            scat.append(recs[hi:hi + 1 + ns])
        gathered_bytes = world * (sum(4 * v.numel() for _, v in dz_views) + self.stage_bytes)
        if mode != "1" and emb_bytes < 2 * gathered_bytes:
            return None       # small tables: the dense all-reduce moves less
        dev = self.device
        g_stage = torch.zeros(world, self.stage_bytes, dtype=torch.uint8, device=dev)
        g_dz = [torch.zeros(world, v.numel(), dtype=torch.float32, device=dev) for _, v in dz_views]
        input_slots = {slot: name for name, slot in prog.inputs.items() if name in self.layout}
        peer_ptrs = []
        for r in range(world):
            if r == rank:
                continue
            pr = ptrs.copy()
            for (slot, _v), buf in zip(dz_views, g_dz):
                pr[slot] = buf[r].data_ptr()
            for slot, name in input_slots.items():
                pr[slot] = g_stage[r].data_ptr() + self.layout[name][0]
            peer_ptrs.append(pr)
        return {"world": world, "rank": rank, "recs": np.concatenate(scat).astype(N.REC_DTYPE), "dz": dz_views, "g_dz": g_dz,
                "g_stage": g_stage, "peer_ptrs": peer_ptrs}

    def _sparse_grad_sync(self, stream):
        import torch.distributed as dist
        ss = self.sparse_sync
        dist.all_gather_into_tensor(ss["g_stage"].view(-1), self.dev_stage)
        for (_slot, v), buf in zip(ss["dz"], ss["g_dz"]):
            dist.all_gather_into_tensor(buf.view(-1), v)
        for pr in ss["peer_ptrs"]:                       # K2 scatter of each peer's rows into the local table gradients
            N.program_run(ss["recs"], pr, stream)
        for name, a in self.flat.g.items():              # everything else (tower / expert / gate weights): summed densely
            if name == "dense" and a.numel() > 1:
                dist.all_reduce(a, op=dist.ReduceOp.SUM)

    # ---- device work of one step (graph-capturable: no syncs, no allocations) ------------------------
    def _body(self):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        import torch.distributed as dist
        if self.p2p:
            if self.recs_pre is not None:
                # row-lazy shards: every rank sees every rank's index columns; the owner of a row replays its postponed
                # updates before anybody reads it
                dist.all_gather_into_tensor(self.g_stage.view(-1), self.dev_stage)
                N.program_run(self.recs_pre, self.ptrs, stream)
            # every owner has finished the previous step's optimizer work on its shards (and this step's catch-up) before
            # any rank reads or adds rows in them
            dist.all_reduce(self._bar, op=dist.ReduceOp.SUM)
        for f, idx, _gv, _gs in self._vf:              # embedding exchange, forward half (NCCL over NVLink)
            self.exchange.lookup(f, idx)
        N.program_run(self.recs_a, self.ptrs, stream)
        if self.recs_b is not None:
            for f, _idx, gv, gs in self._vf:           # backward half: route virtual-table gradients to their owners
                if gv is not None and gs is not None:
                    self.exchange.route_grad(f, gv, gs)
            if self.p2p:
                # ONE all-reduce sums the replicated gradients (tower / expert / gate weights + small tables); it is also the
                # barrier behind which every rank's K2 has finished adding rows into the owners' shard gradients
                dist.all_reduce(self.flat.g_repl, op=dist.ReduceOp.SUM)
            elif self.sparse_sync is not None:
                self._sparse_grad_sync(stream)
            elif self.grad_sync is not None:
                self.grad_sync(self.flat.g)
            N.program_run(self.recs_b, self.ptrs, stream)
        N.memcpy_async(self.loss_ring_host.data_ptr(), self.loss_ring_dev.data_ptr(), 4 * RING, stream)

    def _launch(self):
        if self.use_graph and self.graph is None and self._warm >= 2:
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.device)
            with torch.cuda.graph(g, stream=self._side_stream()):
                self._body()
            self.graph = g
            torch.cuda.synchronize(self.device)
            # the capture did not execute: fall through and replay it for this step
        if self.graph is not None and self.use_graph:
            self.graph.replay()
        else:
            self._body()
            self._warm += 1

    def _side_stream(self):
        if not hasattr(self, "_ss"):
            self._ss = torch.cuda.Stream(self.device)
        return self._ss

    # ---- host side -------------------------------------------------------------------------------------
    def pack(self, x: Dict[str, torch.Tensor], y: torch.Tensor, device=None) -> PackedBatch:
        """Lay a batch out in the staging format once (bench / data pipeline); ``device`` = target of the buffer."""
        buf = torch.zeros(self.stage_bytes, dtype=torch.uint8)
        self._pack_into(buf.numpy(), x, y)
        buf = buf.to(device) if device is not None and torch.device(device).type == "cuda" else buf.pin_memory()
        return PackedBatch(buf, self.B, self.key)

    def _pack_into(self, dst: np.ndarray, x, y):
        for name, (off, nbytes, npdt) in self.layout.items():
            t = x[name]
            if t.dtype != self.dts[name] or t.shape[0] != self.B:
                raise ValueError(f"column {name!r}: dtype/shape changed ({t.dtype}, {tuple(t.shape)})")
            dst[off:off + nbytes].view(npdt)[:] = t.numpy()
        dst[self.y_off:self.y_off + 4 * self.B].view(np.float32)[:] = y.numpy()

    def _scalars(self):
        g = self.optimizer.param_groups[0]
        b1, b2 = g["betas"]
        self.flat.step += 1
        t = self.flat.step
        bc1, bc2 = 1.0 - b1 ** t, 1.0 - b2 ** t
        slot = self.k % RING
        o = slot * self.SC
        self.host_scal_np[o:o + 32].view(np.float32)[:6] = (g["lr"] / bc1, b1, b2, g["eps"], g["weight_decay"], 1.0 / math.sqrt(bc2))
        ctrl = self.host_scal_np[o + 32:o + 48].view(np.int32)
        ctrl[0] = slot
        if self.flat.lazy is not None:
            ctrl[1], ctrl[2] = t, self.flat.lazy.base
        return o

    def step(self, x, y=None) -> LossHandle:
        stream = torch.cuda.current_stream(self.device).cuda_stream
        slot = self.k % RING
        if self._launched[slot] >= 0:
            self.done_ev[slot].synchronize()            # ring entry (scalars + loss) of step k - RING is free again
        if isinstance(x, PackedBatch):
            if x.key != self.key:
                raise ValueError("packed batch belongs to another program")
            N.memcpy_async(self.dev_stage.data_ptr(), x.buf.data_ptr(), self.stage_bytes, stream)
            self._hold[slot] = x.buf        # torch's caching allocators do not see this copy: keep the source alive
        else:
            first = x[self.cols[0]]
            if first.device.type == "cuda":
                base = self.dev_stage.data_ptr()
                for name, (off, nbytes, _dt) in self.layout.items():
                    t = x[name].contiguous()
                    N.memcpy_async(base + off, t.data_ptr(), nbytes, stream)
                yy = y.float().contiguous()
                N.memcpy_async(base + self.y_off, yy.data_ptr(), 4 * self.B, stream)
                self._hold[slot] = (x, yy)
            else:
                i = self.k % NSTAGE
                self.stage_ev[i].synchronize()
                self._pack_into(self.host_np[i], x, y.float() if y.dtype != torch.float32 else y)
                N.memcpy_async(self.dev_stage.data_ptr(), self.host_stage[i].data_ptr(), self.stage_bytes, stream)
                self.stage_ev[i].record()
        o = self._scalars()
        N.memcpy_async(self.dev_scal.data_ptr(), self.host_scal.data_ptr() + o, self.SC, stream)
        self._launch()
        lz = self.flat.lazy
        if self.lazy is not None:
            lz.dirty = True
            if self.flat.step % lz.flush_every == 0 or self.flat.step - lz.base + 2 >= lz.HIST_CAP:
                self._flush_lazy()
                if self.flat.step - lz.base + 2 >= lz.HIST_CAP:     # history table full: every row is current, start it over
                    lz.base = self.flat.step + 1
        self.done_ev[slot].record()
        self._launched[slot] = self.k
        h = LossHandle(self, self.k)
        self.k += 1
        return h

    def scope_records(self, scope: str):
        """Record list of a sub-scope of the step for measurements (bench.py): ``fwd_bwd`` = forward program, BCELoss,
        gradient zero-fill, backward program -- no optimizer op.  Call ``restore_after_scope`` afterwards."""
        if scope != "fwd_bwd":
            raise KeyError(scope)
        opt_kinds = (N.OP_ADAM, N.OP_ADAM_ROWS, N.OP_ADAM_FLUSH)
        keep, i = [], 0
        recs = self.recs_a
        while i < len(recs):
            ns = int(recs[i]["n_sub"])
            if int(recs[i]["kind"]) not in opt_kinds:
                keep.extend(range(i, i + 1 + ns))
            i += 1 + ns
        return np.ascontiguousarray(recs[keep])

    def restore_after_scope(self):
        """The optimizer-free scope leaves gradients behind; the lazy table update relies on an all-zero table-gradient arena."""
        for a in self.g.values():
            a.zero_()
        torch.cuda.synchronize(self.device)

    def _flush_lazy(self):
        """Replay every postponed row update up to the current step (device work on the current stream)."""
        lz = self.flat.lazy
        stream = torch.cuda.current_stream(self.device).cuda_stream
        N.program_run(self._flush_recs, self.ptrs, stream)
        lz.dirty = False

    def _read_loss(self, k: int) -> float:
        slot = k % RING
        if self._launched[slot] != k:
            raise RuntimeError("loss handle expired: more than %d steps were launched since" % RING)
        self.done_ev[slot].synchronize()
        self.runner.check_indices()
        if self.exchange is not None:
            self.exchange.check_indices()
        return float(self.loss_ring_host[slot])
