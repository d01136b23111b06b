"""Host-side lowering of a model to the device program executed by ``swr_program_run``.

A model's forward is described once per (batch size, train/eval, column dtypes) with a
:class:`ProgramBuilder`: ``gather`` -> ``fc`` layers (grouped: all experts / gates /
towers of a level in one launch) -> ``pool`` -> ``head``.  The builder then emits

* the forward record list,
* the backward record list (reverse-mode at op granularity; the per-op backward math is
  hand-derived and lives in the kernels, see DESIGN.md "Backward"),
* a slot table: every device pointer a record refers to, by index.  Parameters, buffers
  and workspace slots are static; feature columns and gradient arenas are per call.

Activations are *lazy*: an :class:`Act` is a raw buffer plus the BatchNorm / activation
its consumers apply while loading it (reference: ``Linear -> BatchNorm1d -> act`` of
basic/layers.py:253-258 collapses into the consumer of the Linear output).

The builder is backend-agnostic.  The product executes programs with
:class:`CudaRunner` (C-ABI, no fallback); tests execute the same records with the torch
CPU interpreter in ``oracle/ops_ref.py`` to check the host logic without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as N


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


@dataclass
class Norm:
    """Normalisation applied lazily to a raw activation (BatchNorm1d or STAR's partitioned norm)."""
    gamma: Optional[torch.Tensor] = None
    beta: Optional[torch.Tensor] = None
    rmean: Optional[torch.Tensor] = None
    rvar: Optional[torch.Tensor] = None
    nbt: Optional[torch.Tensor] = None
    eps: float = 1e-5
    gamma2: Optional[torch.Tensor] = None
    beta2: Optional[torch.Tensor] = None
    always_batch: bool = False      # STAR / HAMUR: batch statistics in eval mode too (star.py:95-100, hamur.py:192-195)
    unbiased: bool = False          # HAMUR domain norm: torch.var default (divide by B - 1)
    repeat: int = 1                 # running-stat updates per step (HAMUR evaluates its shared hyper-net once per domain)


class Act:
    """A lazily normalised / activated [B, n] activation living in the workspace."""

    def __init__(self, raw: int, ld: int, n: int, norm: Optional[Norm] = None, act: int = N.ACT_NONE):
        self.raw, self.ld, self.n, self.norm, self.act = raw, ld, n, norm, act
        self.stats = -1         # f64 [n][2] forward moments (batch mode)
        self.dz = -1            # gradient buffer
        self.dstats = -1        # f64 [n][2] backward sums
        self.mode = N.NORM_NONE
        self.needs_grad = False
        self.grad_written = False
        self.nslots: Dict[str, int] = {}
        self.grad_cols: Optional[int] = None   # only the first grad_cols columns receive an FC data gradient
        self.parent: Optional["Act"] = None    # column sub-view of another activation
        self.col = 0


@dataclass
class _FcGroup:
    src: Act
    out: Act
    W: torch.Tensor
    b: Optional[torch.Tensor]
    W2: Optional[torch.Tensor] = None
    b2: Optional[torch.Tensor] = None
    layout: int = N.W_NK
    e_act: int = N.ACT_NONE
    e_scale: float = 1.0
    detach: bool = False            # no data gradient flows back into src (x.detach() in the reference)
    img_f: int = -1                 # workspace slots of the presplit weight images the tcgen05 kernels read through TMA:
    img_d: int = -1                 # [2][N][K32] (forward) and [2][K][N32] (data gradient), hi / lo TF32 planes


@dataclass
class Program:
    B: int
    training: bool
    recs_fwd: np.ndarray
    recs_bwd: np.ndarray
    slot_desc: List[tuple]                     # per slot: ("static", tensor) | ("input", name) | ("ws32"|"ws64", off, n) | ("grad", arena, off, n) | ("special", name)
    ws32: int
    ws64: int
    inputs: Dict[str, int]
    params: List[torch.nn.Parameter]           # parameters that receive a gradient, in arena order
    param_arena: List[Tuple[str, int, int]]    # (arena, offset, numel) per entry of ``params``
    arena_size: Dict[str, int]
    out_slot: int = -1                         # head output [B] (-1 when the program has no head)
    gout_slot: int = -1
    oob_slot: int = -1
    outputs: List[Tuple[int, int, int, int]] = field(default_factory=list)   # plain [B, n] outputs: (raw slot, ld, n, dz slot)
    labels: Dict[int, str] = field(default_factory=dict)
    virtual_fields: list = field(default_factory=list)   # parallel._Field of every row-sharded table the program reads
    peer_tabs: list = field(default_factory=list)        # (shard parameter, full vocabulary, world) of every table read through peer pointers
    n_launch_fwd: int = 0
    n_launch_bwd: int = 0


class ProgramBuilder:
    def __init__(self, B: int, training: bool):
        if training and B <= 1:
            # torch.nn.BatchNorm1d raises this for a batch of one in train mode
            raise ValueError("Expected more than 1 value per channel when training")
        self.B, self.training = int(B), bool(training)
        self.slot_desc: List[tuple] = []
        self._static: Dict[int, int] = {}
        self.inputs: Dict[str, int] = {}
        self.ws32 = 0
        self.ws64 = 0
        self.params: List[torch.nn.Parameter] = []
        self._param_slot: Dict[int, int] = {}
        self.param_arena: List[Tuple[str, int, int]] = []
        self.arena_size = {"dense": 0, "emb": 0, "virt": 0, "shard": 0}
        self.exchange = None                 # parallel.ShardedExchange of the owning model (row-sharded tables)
        self.p2p = False                     # row-sharded tables are read / updated in their owners' memory (peer pointers)
        self.peer_tabs: Dict[int, tuple] = {}  # id(shard parameter) -> (parameter, full vocabulary, world)
        self.virtual_fields: list = []
        self.tape: List[tuple] = []
        self.norm_acts: List[Act] = []
        self.out_slot = self.gout_slot = self.oob_slot = -1
        self._stats_slots: List[int] = []
        self._outputs: List[Act] = []
        self.labels: Dict[int, str] = {}     # debug names of workspace slots (tests compare them one by one)
        self._n_acts = 0

    # ---- slots ------------------------------------------------------------------------
    def _new_slot(self, desc) -> int:
        self.slot_desc.append(desc)
        return len(self.slot_desc) - 1

    def static(self, t: Optional[torch.Tensor]) -> int:
        if t is None:
            return -1
        key = id(t)
        if key not in self._static:
            self._static[key] = self._new_slot(("static", t))
        return self._static[key]

    def input(self, name: str) -> int:
        if name not in self.inputs:
            self.inputs[name] = self._new_slot(("input", name))
        return self.inputs[name]

    def ws(self, nelem: int, f64: bool = False) -> int:
        if f64:
            off = self.ws64
            self.ws64 += _round_up(nelem, 2)
            return self._new_slot(("ws64", off, nelem))
        off = self.ws32
        self.ws32 += _round_up(nelem, 64)          # 256-byte aligned buffers
        return self._new_slot(("ws32", off, nelem))

    def grad(self, p: Optional[torch.Tensor], arena: str = "dense") -> int:
        """Slot of the gradient of parameter ``p`` inside the per-call gradient arena (-1 if frozen)."""
        if p is None or not p.requires_grad:
            return -1
        key = id(p)
        if key not in self._param_slot:
            off = self.arena_size[arena]
            self.arena_size[arena] += _round_up(p.numel(), 4)      # keep every gradient 16-byte aligned
            self.params.append(p)
            self.param_arena.append((arena, off, p.numel()))
            self._param_slot[key] = self._new_slot(("grad", arena, off, p.numel()))
        return self._param_slot[key]

    def _peer_slot(self, kind: str, tab) -> int:
        key = (kind, id(tab))
        if key not in self._static:
            self._static[key] = self._new_slot(("special", (kind, tab)))
        return self._static[key]

    def virtual_field(self, fea, shard_param):
        """The virtual table of a row-sharded feature for this batch size (see parallel.py)."""
        if self.exchange is None:
            raise RuntimeError(f"feature {fea.name!r} is row-sharded but the model has no ShardedExchange")
        f = self.exchange.field(fea, shard_param, self.B)
        if f not in self.virtual_fields:
            self.virtual_fields.append(f)
        return f

    def peer_table(self, fea, shard_param):
        """Register a row-sharded table that the gather reads (and the scatter updates) through peer pointers."""
        self.peer_tabs[id(shard_param)] = (shard_param, int(fea.vocab_size), int(fea.shard.world))
        return shard_param

    # ---- activations ------------------------------------------------------------------
    def new_act(self, n: int, norm: Optional[Norm] = None, act: int = N.ACT_NONE, raw: Optional[int] = None,
                ld: Optional[int] = None) -> Act:
        ld = _round_up(n, 4) if ld is None else ld
        a = Act(self.ws(self.B * ld) if raw is None else raw, ld, n, norm, act)
        a.name = f"a{self._n_acts}[{n}]"
        self._n_acts += 1
        self.labels.setdefault(a.raw, a.name + ".raw")
        if norm is not None:
            has_running = norm.rmean is not None and norm.rvar is not None
            a.mode = N.NORM_BATCH if (self.training or norm.always_batch or not has_running) else N.NORM_RUNNING
            if a.mode == N.NORM_BATCH:
                a.stats = self.ws(2 * n, f64=True)
                self._stats_slots.append(a.stats)
                self.labels[a.stats] = a.name + ".stats"
            a.nslots = {k: self.static(getattr(norm, k)) for k in ("gamma", "gamma2", "beta", "beta2", "rmean", "rvar")}
            self.norm_acts.append(a)
        return a

    def _ensure_grad(self, a: Act):
        if a.dz < 0:
            a.dz = self.ws(self.B * a.ld)
            a.needs_grad = True
            self.labels[a.dz] = a.name + ".dz"

    # ---- forward ops ------------------------------------------------------------------
    def gather(self, sparse: Sequence[Tuple[str, torch.Tensor]], dense: Sequence[str],
               col_dtypes: Dict[str, torch.dtype]) -> Act:
        """sparse: (column name, table parameter [vocab, E]); dense: column names."""
        return self.gather_parts([(sparse, dense, True)], col_dtypes)

    def gather_parts(self, parts, col_dtypes: Dict[str, torch.dtype]) -> Act:
        """Several EmbeddingLayer outputs concatenated into one activation (``torch.cat`` of
        ppnet.py:54 / epnet.py:28).  parts: (sparse, dense, differentiable); a part with
        differentiable=False is ``.detach()``-ed in the reference: its tables get no gradient."""
        placed, col = [], 0
        for sparse, dense, diff in parts:
            n = sum(int(t.shape[1]) for _, t in sparse) + len(dense)
            if n == 0:
                raise ValueError("The input features can note be empty")
            placed.append((col, list(sparse), list(dense), bool(diff)))
            col += n
        x = self.new_act(col)
        if self.oob_slot < 0:
            self.oob_slot = self._new_slot(("special", "oob"))
        self.tape.append(("gather", x, placed, dict(col_dtypes)))
        return x

    def subview(self, src: Act, col: int, n: int) -> Act:
        """Columns [col, col + n) of a plain activation (shares storage and gradient buffer)."""
        if src.norm is not None or src.act != N.ACT_NONE:
            raise NotImplementedError("subview of a lazy activation")
        d = self.slot_desc[src.raw]
        raw = self._new_slot(("ws32", d[1] + col, d[2] - col))
        a = self.new_act(n, raw=raw, ld=src.ld)
        a.parent, a.col = src, col
        return a

    def colstats(self, src: Act, norms: Sequence[Norm]) -> List[Act]:
        """Views of ``src`` normalised with its own whole-batch statistics, one per Norm (STAR
        partitioned norm: every domain scales/shifts the same normalised input, star.py:95-100)."""
        views = []
        for norm in norms:
            a = self.new_act(src.n, norm, N.ACT_NONE, raw=src.raw, ld=src.ld)
            if views:                       # share the statistics of the first view
                self._stats_slots.remove(a.stats)
                a.stats = views[0].stats
            a.base = src
            views.append(a)
        self.tape.append(("colstats", src, views))
        return views

    def fc(self, groups: Sequence[dict]) -> List[Act]:
        """Each group: dict(src=Act, W=, b=, [W2=, b2=, layout=], norm=Norm|None, act=code, [e_act=, e_scale=])."""
        gs = []
        for g in groups:
            W = g["W"]
            layout = g.get("layout", N.W_NK)
            n_out = int(W.shape[0] if layout == N.W_NK else W.shape[1])
            k_in = int(W.shape[1] if layout == N.W_NK else W.shape[0])
            if k_in != g["src"].n:
                raise ValueError(f"fc: weight expects {k_in} inputs, activation has {g['src'].n}")
            out = self.new_act(n_out, g.get("norm"), g.get("act", N.ACT_NONE))
            gs.append(_FcGroup(g["src"], out, W, g.get("b"), g.get("W2"), g.get("b2"), layout,
                               g.get("e_act", N.ACT_NONE), float(g.get("e_scale", 1.0)), bool(g.get("detach", False))))
        self.tape.append(("fc", gs))
        return [g.out for g in gs]

    def pool(self, gates: Sequence[Tuple[Act, Sequence[Act]]]) -> List[Act]:
        """gates: (gate logits Act [B, nE] (BatchNorm, softmax applied by the op), expert Acts)."""
        experts: List[Act] = []
        entries = []
        H = gates[0][1][0].n
        for gate, exps in gates:
            if gate.n != len(exps) or len(exps) > 16:
                raise ValueError("pool: gate width must equal its expert count (<= 16)")
            idx = []
            for e in exps:
                if e.n != H:
                    raise ValueError("pool: experts must share their width")
                if e not in experts:
                    experts.append(e)
                idx.append(experts.index(e))
            out = self.new_act(H)
            probs = self.ws(self.B * gate.n)
            self.labels[probs] = out.name + ".probs"
            entries.append((gate, idx, out, probs))
        self.tape.append(("pool", entries, experts, H))
        return [e[2] for e in entries]

    def ew(self, mode: int, pairs, scale: float = 1.0) -> List[Act]:
        """Element-wise op per pair (A, C): MUL -> value(A)*value(C)*scale, ADD -> value(A)+value(C),
        COPY (C None) -> value(A) materialised.  Returns plain activations."""
        entries = []
        for A, C in pairs:
            if mode != N.EW_COPY and A.n != C.n:
                raise ValueError("ew: operand widths differ")
            entries.append((A, C, self.new_act(A.n)))
        self.tape.append(("ew", mode, float(scale), entries))
        return [e[2] for e in entries]

    def select(self, ys: Sequence[Act], dom_dtype: torch.dtype) -> Act:
        """out[b] = value(ys[domain_indicator[b]])[b] (zeros for ids outside [0, D))."""
        out = self.new_act(ys[0].n)
        self.tape.append(("select", list(ys), out, dom_dtype))
        return out

    def layernorm(self, groups) -> List[Act]:
        """groups: (y plain FC output, gamma, beta, eps, act code) -> act(LayerNorm(y)) plain."""
        entries = []
        for y, gamma, beta, eps, act in groups:
            if y.norm is not None or y.act != N.ACT_NONE:
                raise NotImplementedError("layernorm input must be a plain activation")
            if y.n > 512:
                raise NotImplementedError("LayerNorm wider than 512 is not supported by the fused kernel")
            out = self.new_act(y.n)
            rs = self.ws(2 * self.B)
            self.labels[rs] = out.name + ".rowstats"
            entries.append((y, gamma, beta, float(eps), int(act), out, rs))
        self.tape.append(("ln", entries))
        return [e[5] for e in entries]

    def mix(self, xs: Sequence[Act], pooled: Sequence[Act], w_exp: torch.Tensor, w_bal: torch.Tensor):
        """M3oE (m3oe.py:168-187): pooled[d] += sig(w_exp) * (sig(w_bal) * xs[d] + (1 - sig(w_bal)) / (D - 1) * sum_{j != d} xs[j])."""
        for a in list(xs) + list(pooled):
            if a.norm is not None or a.act != N.ACT_NONE:
                raise NotImplementedError("mix operands must be plain activations")
        self.tape.append(("mix", list(xs), list(pooled), w_exp, w_bal))

    def bmv(self, ps: Sequence[Act], H: Act, k: int) -> List[Act]:
        """HAMUR: q[b, :] = p[b, :] @ H[b].reshape(k, k) for every p in ps (they share H)."""
        if H.norm is not None or H.act != N.ACT_NONE or H.n != k * k:
            raise NotImplementedError("bmv: H must be a plain [B, k*k] activation")
        entries = [(p_, self.new_act(k)) for p_ in ps]
        self.tape.append(("bmv", entries, H, int(k)))
        return [e[1] for e in entries]

    def head(self, domains: Sequence[Tuple[Act, Optional[torch.Tensor], Optional[torch.Tensor]]],
             dom_dtype: torch.dtype, sig_before_select=True, add: Optional[Act] = None) -> int:
        """sig_before_select: True/1 select(sigmoid(v_d)); False/0 sigmoid(select(v_d) + add);
        2 (N.HEAD_NO_SELECT) sigmoid(v_0) for every row (EPNet has no domain mask)."""
        self.out_slot = self.ws(self.B)
        self.labels[self.out_slot] = "head.out"
        self.gout_slot = self.input("__grad_out__")
        self.tape.append(("head", list(domains), dom_dtype, sig_before_select, add))
        return self.out_slot

    def output(self, a: Act) -> Act:
        """Expose a plain activation as a tensor returned by the program (its gradient comes back in)."""
        if a.norm is not None or a.act != N.ACT_NONE:
            raise NotImplementedError("only plain activations can be program outputs")
        self._outputs.append(a)
        return a

    # ---- record emission ----------------------------------------------------------------
    @staticmethod
    def _rec(kind: int, n_sub: int = 0) -> np.ndarray:
        r = np.zeros((), dtype=N.REC_DTYPE)
        r["kind"], r["n_sub"] = kind, n_sub
        r["s"][:] = -1
        return r

    def _put_act(self, r: np.ndarray, a: Act, sb: int, ib: int, fb: int):
        s = r["s"]
        s[sb + 0] = a.raw
        s[sb + 1] = a.stats
        if a.norm is not None:
            ns = a.nslots
            s[sb + 2], s[sb + 3] = ns["rmean"], ns["rvar"]
            s[sb + 4], s[sb + 5], s[sb + 6], s[sb + 7] = ns["gamma"], ns["gamma2"], ns["beta"], ns["beta2"]
            r["f"][fb + 0] = a.norm.eps
        r["f"][fb + 1] = (self.B / (self.B - 1.0)) if (a.norm is not None and a.norm.unbiased and self.B > 1) else 1.0
        s[sb + 8], s[sb + 9] = a.dz, a.dstats
        r["i"][ib + 0], r["i"][ib + 1], r["i"][ib + 2], r["i"][ib + 3] = a.ld, a.n, a.mode, a.act

    def _fc_rec(self, g: _FcGroup, flags: int = 0) -> np.ndarray:
        r = self._rec(N.OP_GROUP)
        self._put_act(r, g.src, 0, 0, 0)
        self._put_act(r, g.out, 12, 4, 2)
        s = r["s"]
        s[24], s[25], s[26], s[27] = self.static(g.W), self.static(g.W2), self.static(g.b), self.static(g.b2)
        s[28], s[29], s[30], s[31] = self.grad(g.W), self.grad(g.W2), self.grad(g.b), self.grad(g.b2)
        s[10], s[11] = g.img_f, g.img_d
        r["i"][8], r["i"][9], r["i"][10], r["i"][11] = g.layout, int(g.W.stride(0)), g.e_act, flags
        r["i"][13] = g.src.n           # full input width (a dgrad record may narrow i[1] to the columns that get a gradient)
        r["f"][4] = g.e_scale
        return r

    def _hdr(self, kind: int, n_sub: int, **kw) -> np.ndarray:
        r = self._rec(kind, n_sub)
        r["i"][0] = self.B
        r["f"][0] = 1.0 / self.B
        for k, v in kw.items():
            if k[0] == "i":
                r["i"][int(k[1:])] = v
            elif k[0] == "f":
                r["f"][int(k[1:])] = v
            else:
                r["s"][int(k[1:])] = v
        return r

    @staticmethod
    def _split64(v: int) -> Tuple[int, int]:
        lo, hi = v & 0xFFFFFFFF, (v >> 32) & 0xFFFFFFFF
        return (lo - (1 << 32) if lo >= (1 << 31) else lo), (hi - (1 << 32) if hi >= (1 << 31) else hi)

    def _zero_rec(self, slot: int, nbytes: int) -> np.ndarray:
        lo, hi = self._split64(nbytes)
        return self._hdr(N.OP_ZERO, 0, s0=slot, i0=lo, i1=hi)

    def _plain_rec(self, slots: Sequence[int], ints: Sequence[int] = (), floats: Sequence[float] = ()) -> np.ndarray:
        r = self._rec(N.OP_GROUP)
        for i, v in enumerate(slots):
            r["s"][i] = v
        for i, v in enumerate(ints):
            r["i"][i] = v
        for i, v in enumerate(floats):
            r["f"][i] = v
        return r

    def _act_rec(self, a: Act) -> np.ndarray:
        r = self._rec(N.OP_GROUP)
        self._put_act(r, a, 0, 0, 0)
        return r

    def _ew_rec(self, mode: int, scale: float, A: Act, C: Optional[Act], out: Act, flags: int = 0) -> np.ndarray:
        r = self._rec(N.OP_GROUP)
        self._put_act(r, A, 0, 0, 0)
        if C is not None:
            self._put_act(r, C, 12, 4, 2)
        r["s"][24], r["s"][25] = out.raw, out.dz
        r["i"][8], r["i"][9], r["i"][10] = out.ld, mode, flags
        r["f"][4] = scale
        return r

    def _ln_rec(self, e) -> np.ndarray:
        y, gamma, beta, eps, act, out, rs = e
        return self._plain_rec([y.raw, y.dz, self.static(gamma), self.static(beta), self.grad(gamma), self.grad(beta),
                                out.raw, out.dz, rs], [y.ld, y.n, out.ld, act], [eps])

    def _mark_needs_grad(self):
        """Which activations need a gradient buffer: everything downstream of a trainable input."""
        for op in self.tape:
            kind = op[0]
            if kind == "gather":
                x, parts = op[1], op[2]
                if any(diff and any(t.requires_grad for _, t in sparse) for _c, sparse, _d, diff in parts):
                    self._ensure_grad(x)
            elif kind == "colstats":
                src, views = op[1], op[2]
                for v in views:
                    if src.needs_grad or any(self.grad(getattr(v.norm, k)) >= 0 for k in ("gamma", "gamma2", "beta", "beta2")):
                        self._ensure_grad(v)
            elif kind == "fc":
                for g in op[1]:
                    self._ensure_grad(g.out)
            elif kind == "pool":
                for gate, idx, out, probs in op[1]:
                    self._ensure_grad(out)
            elif kind == "ew":
                for A, C, out in op[3]:
                    if A.needs_grad or (C is not None and C.needs_grad):
                        self._ensure_grad(out)
            elif kind == "select":
                if any(y.needs_grad for y in op[1]):
                    self._ensure_grad(op[2])
            elif kind == "ln":
                for e in op[1]:
                    self._ensure_grad(e[5])
            elif kind == "bmv":
                for p_, q in op[1]:
                    if p_.needs_grad or op[2].needs_grad:
                        self._ensure_grad(q)
        # column sub-views share their parent's gradient buffer
        for a in self._subviews():
            if a.parent.needs_grad and a.dz < 0:
                d = self.slot_desc[a.parent.dz]
                a.dz = self._new_slot(("ws32", d[1] + a.col, d[2] - a.col))
                a.needs_grad = True

    def _subviews(self) -> List[Act]:
        seen, out = set(), []

        def visit(a):
            if a is not None and a.parent is not None and id(a) not in seen:
                seen.add(id(a))
                out.append(a)
        for op in self.tape:
            kind = op[0]
            if kind == "fc":
                for g in op[1]:
                    visit(g.src)
            elif kind == "ew":
                for A, C, _o in op[3]:
                    visit(A), visit(C)
            elif kind == "head":
                for a, _w, _b in op[1]:
                    visit(a)
        return out

    def finish(self) -> Program:
        B = self.B
        fwd: List[np.ndarray] = []
        bwd_blocks: List[list] = []
        # sub-views must see their parent's gradient buffer before consumers are marked: two passes
        self._mark_needs_grad()
        self._mark_needs_grad()
        for a in self._outputs:
            if a.needs_grad:
                if a.grad_written:
                    raise NotImplementedError("an output activation that is also consumed inside the program")
                a.grad_written = True        # the caller's gradient is copied into dz before the backward runs
        for a in self.norm_acts:
            if a.needs_grad:
                a.dstats = self.ws(2 * a.n, f64=True)
                self.labels[a.dstats] = a.name + ".dstats"
        dstat_slots = [a.dstats for a in self.norm_acts if a.dstats >= 0]
        mix_red: Dict[int, int] = {}
        for ti, op in enumerate(self.tape):
            if op[0] == "mix":
                mix_red[ti] = self.ws(2, f64=True)
                dstat_slots.append(mix_red[ti])

        # weight images for the tensor-core FC kernels: written by ONE presplit launch at the start of every forward
        # (weights change between passes: optimizer steps, load_state_dict), read by the FC ops of both passes
        fc_groups = [g for op in self.tape if op[0] == "fc" for g in op[1]]
        for g in fc_groups:
            n_out, k_in = g.out.n, g.src.n
            g.img_f = self.ws(2 * n_out * _round_up(k_in, 32))
            if self.training and g.src.needs_grad and not g.detach:
                g.img_d = self.ws(2 * k_in * _round_up(n_out, 32))
        if fc_groups:
            fwd.append(self._hdr(N.OP_FC_PRESPLIT, len(fc_groups)))
            fwd.extend(self._fc_rec(g) for g in fc_groups)

        # forward statistics: every f64 buffer allocated while building is a forward statistic -> one span
        f64_fwd = [self.slot_desc[s] for s in self._stats_slots]
        if f64_fwd:
            lo = min(d[1] for d in f64_fwd)
            hi = max(d[1] + _round_up(d[2], 2) for d in f64_fwd)
            first = min(self._stats_slots, key=lambda s: self.slot_desc[s][1])
            fwd.append(self._zero_rec(first, (hi - lo) * 8))

        for ti, op in enumerate(self.tape):
            kind = op[0]
            if kind == "gather":
                x, parts, dts = op[1], op[2], op[3]
                srecs = []
                for col0, sparse, dense, diff in parts:
                    subs = []
                    col = col0
                    for name, tab in sparse:
                        r = self._rec(N.OP_GROUP)
                        vocab, E = int(tab.shape[0]), int(tab.shape[1])
                        r["s"][0], r["s"][1] = self.static(tab), self.input(name)
                        peer = self.peer_tabs.get(id(tab))
                        if peer is not None:      # vocabulary = the whole table; rows live at peers[row % world]
                            vocab = peer[1]
                            r["i"][6] = peer[2]
                            r["s"][2] = self._peer_slot("peers_p", tab)
                        r["i"][0], r["i"][1] = self._split64(vocab)
                        r["i"][2], r["i"][3], r["i"][4], r["i"][5] = N.torch_dtype_code(dts[name]), col, 0, E
                        subs.append(r)
                        if tab.requires_grad and diff:
                            g = r.copy()
                            is_virt = any(tab is f.virt for f in self.virtual_fields)
                            g["s"][0] = self.grad(tab, "virt" if is_virt else ("shard" if peer is not None else "emb"))
                            if peer is not None:
                                g["s"][2] = self._peer_slot("peers_g", tab)
                            srecs.append(g)
                        col += E
                    for name in dense:
                        r = self._rec(N.OP_GROUP)
                        r["s"][0] = self.input(name)
                        r["i"][2], r["i"][3], r["i"][4] = N.torch_dtype_code(dts[name]), col, 1
                        subs.append(r)
                        col += 1
                    fwd.append(self._hdr(N.OP_GATHER, len(subs), i1=len(sparse), i2=len(dense), i4=x.ld, i5=col0,
                                         s0=x.raw, s1=self.oob_slot))
                    fwd.extend(subs)
                if srecs and x.needs_grad:
                    bwd_blocks.append([self._hdr(N.OP_SCATTER, len(srecs), i1=len(srecs), i4=x.ld, s0=x.dz)] + srecs)
            elif kind == "colstats":
                src, views = op[1], op[2]
                fwd.append(self._hdr(N.OP_COLSTATS, 0, i1=src.n, i2=src.ld, s0=src.raw, s1=views[0].stats))
                if src.needs_grad:
                    bwd_blocks.append([("sumgrad", src, views)])
            elif kind == "fc":
                gs = op[1]
                fwd.append(self._hdr(N.OP_FC_FWD, len(gs)))
                fwd.extend(self._fc_rec(g) for g in gs)
                blk: list = [("wgrad", gs)]
                # one fan-in dgrad per distinct input activation that needs a gradient
                by_src: Dict[int, List[_FcGroup]] = {}
                for g in gs:
                    if g.src.needs_grad and not g.detach:
                        by_src.setdefault(id(g.src), []).append(g)
                if by_src:
                    blk.append(("dgrad", list(by_src.values())))
                bwd_blocks.append(blk)
            elif kind == "pool":
                entries, experts, H = op[1], op[2], op[3]
                subs = []
                for gate, idx, out, probs in entries:
                    r = self._rec(N.OP_GROUP)
                    self._put_act(r, gate, 0, 0, 0)
                    self._put_act(r, out, 12, 4, 2)
                    r["s"][24] = probs
                    r["i"][8] = len(idx)
                    r["i"][16:16 + len(idx)] = idx
                    subs.append(r)
                for e in experts:
                    subs.append(self._act_rec(e))
                fwd.append(self._hdr(N.OP_POOL_FWD, len(subs), i1=H, i2=len(entries), i3=len(experts)))
                fwd.extend(subs)
                bwd_blocks.append([("pool_bwd", entries, experts, H)])
            elif kind == "ew":
                mode, scale, entries = op[1], op[2], op[3]
                fwd.append(self._hdr(N.OP_EW_FWD, len(entries)))
                fwd.extend(self._ew_rec(mode, scale, A, C, out) for A, C, out in entries)
                bwd_blocks.append([("ew_bwd", mode, scale, entries)])
            elif kind == "select":
                ys, out, dom_dtype = op[1], op[2], op[3]
                fwd.append(self._hdr(N.OP_SELECT_FWD, len(ys), i1=out.n, i2=out.ld, i3=N.torch_dtype_code(dom_dtype),
                                     s0=self.input("domain_indicator"), s1=out.raw, s2=-1))
                fwd.extend(self._act_rec(y) for y in ys)
                bwd_blocks.append([("select_bwd", ys, out, dom_dtype)])
            elif kind == "ln":
                entries = op[1]
                fwd.append(self._hdr(N.OP_LN_FWD, len(entries)))
                fwd.extend(self._ln_rec(e) for e in entries)
                bwd_blocks.append([("ln_bwd", entries)])
            elif kind == "mix":
                xs, pooled, w_exp, w_bal = op[1], op[2], op[3], op[4]
                subs = [self._plain_rec([x.raw, -1, o.raw, -1]) for x, o in zip(xs, pooled)]
                fwd.append(self._hdr(N.OP_MIX_FWD, len(subs), i1=len(xs), i2=xs[0].n, i3=xs[0].ld, i4=pooled[0].ld,
                                     s0=self.static(w_exp), s1=self.static(w_bal), s2=-1, s3=-1, s4=-1))
                fwd.extend(subs)
                bwd_blocks.append([("mix_bwd", xs, pooled, w_exp, w_bal, mix_red[ti])])
            elif kind == "bmv":
                entries, H, k = op[1], op[2], op[3]
                fwd.append(self._hdr(N.OP_BMV_FWD, len(entries), i1=k, i2=H.ld, i3=0, s0=H.raw, s1=-1))
                fwd.extend(self._plain_rec([p_.raw, -1, q.raw, -1], [p_.ld, q.ld]) for p_, q in entries)
                bwd_blocks.append([("bmv_bwd", entries, H, k)])
            elif kind == "head":
                domains, dom_dtype, sbs, add = op[1], op[2], op[3], op[4]
                subs = []
                for a, w, b in domains:
                    r = self._act_rec(a)
                    r["s"][24], r["s"][26] = self.static(w), self.static(b)
                    subs.append(r)
                no_sel = int(sbs) == N.HEAD_NO_SELECT
                hdr = self._hdr(N.OP_HEAD_FWD, len(subs), i1=len(subs), i2=int(sbs),
                                i3=N.DT_I64 if no_sel else N.torch_dtype_code(dom_dtype),
                                s0=-1 if no_sel else self.input("domain_indicator"), s1=self.out_slot, s2=-1,
                                s3=(add.raw if add is not None else -1), s4=-1, i4=(add.ld if add is not None else 1))
                fwd.append(hdr)
                fwd.extend(subs)
                bwd_blocks.append([("head_bwd", domains, dom_dtype, sbs, add)])

        if self.training:
            bn = [a for a in self.norm_acts if a.mode == N.NORM_BATCH and a.norm.rmean is not None]
            if bn:
                subs = []
                for a in bn:
                    r = self._act_rec(a)
                    r["s"][24] = self.static(a.norm.nbt)
                    r["i"][8] = int(a.norm.repeat)
                    subs.append(r)
                fwd.append(self._hdr(N.OP_BN_UPDATE, len(subs), f4=0.1))
                fwd.extend(subs)

        # ---- backward: reverse the blocks, materialise deferred records now that every dz slot exists
        bwd: List[np.ndarray] = []
        if dstat_slots:
            descs = [self.slot_desc[s] for s in dstat_slots]
            lo = min(d[1] for d in descs)
            hi = max(d[1] + _round_up(d[2], 2) for d in descs)
            first = min(dstat_slots, key=lambda s: self.slot_desc[s][1])
            bwd.append(self._zero_rec(first, (hi - lo) * 8))

        def first_write(a: Act) -> bool:
            """True if this is the first gradient written into ``a`` (later writers accumulate)."""
            root = a.parent if a.parent is not None else a
            acc = a.grad_written
            a.grad_written = True
            if a.parent is not None:
                root.partial_written = True
            return not acc

        for blk in reversed(bwd_blocks):
            for item in blk:
                if isinstance(item, np.ndarray):
                    bwd.append(item)
                    continue
                tag = item[0]
                if tag == "head_bwd":
                    _, domains, dom_dtype, sbs, add = item
                    subs = []
                    for a, w, b in domains:
                        if not first_write(a):
                            raise NotImplementedError("head: a tower output consumed twice")
                        r = self._act_rec(a)
                        r["s"][24], r["s"][26] = self.static(w), self.static(b)
                        r["s"][28], r["s"][30] = self.grad(w), self.grad(b)
                        subs.append(r)
                    dadd = -1
                    if add is not None and add.needs_grad:
                        if not first_write(add):
                            raise NotImplementedError("head: additive term consumed twice")
                        dadd = add.dz
                    no_sel = int(sbs) == N.HEAD_NO_SELECT
                    bwd.append(self._hdr(N.OP_HEAD_BWD, len(subs), i1=len(subs), i2=int(sbs),
                                         i3=N.DT_I64 if no_sel else N.torch_dtype_code(dom_dtype),
                                         s0=-1 if no_sel else self.input("domain_indicator"), s1=self.out_slot, s2=self.gout_slot,
                                         s3=(add.raw if add is not None else -1), s4=dadd, i4=(add.ld if add is not None else 1)))
                    bwd.extend(subs)
                elif tag == "pool_bwd":
                    _, entries, experts, H = item
                    subs = []
                    for gate, idx, out, probs in entries:
                        if not first_write(gate):
                            raise NotImplementedError("pool: gate logits consumed twice")
                        r = self._rec(N.OP_GROUP)
                        self._put_act(r, gate, 0, 0, 0)
                        self._put_act(r, out, 12, 4, 2)
                        r["s"][24] = probs
                        r["i"][8] = len(idx)
                        r["i"][16:16 + len(idx)] = idx
                        subs.append(r)
                    for e in experts:
                        if not first_write(e):
                            raise NotImplementedError("pool: expert output consumed by two pooling ops")
                        subs.append(self._act_rec(e))
                    bwd.append(self._hdr(N.OP_POOL_BWD, len(subs), i1=H, i2=len(entries), i3=len(experts)))
                    bwd.extend(subs)
                elif tag == "wgrad":
                    gs = item[1]
                    bwd.append(self._hdr(N.OP_FC_WGRAD, len(gs)))
                    bwd.extend(self._fc_rec(g) for g in gs)
                elif tag == "dgrad":
                    # ONE launch for every destination of this level: groups sorted by destination (i[12]), each
                    # destination sums its own fan-in
                    lists = item[1]
                    bwd.append(self._hdr(N.OP_FC_DGRAD, sum(len(lst) for lst in lists), i1=len(lists)))
                    for di, lst in enumerate(lists):
                        src = lst[0].src
                        flags = 1 | (0 if first_write(src) else 2)
                        for g in lst:
                            r = self._fc_rec(g, flags)
                            r["i"][12] = di
                            if src.grad_cols is not None:
                                r["i"][1] = src.grad_cols        # narrower destination: only these columns get dA
                            bwd.append(r)
                elif tag == "sumgrad":
                    _, src, views = item
                    live = [v for v in views if v.needs_grad]
                    acc = 0 if first_write(src) else 1
                    bwd.append(self._hdr(N.OP_SUMGRAD, 1 + len(live), i1=acc))
                    bwd.append(self._act_rec(src))
                    bwd.extend(self._act_rec(v) for v in live)
                elif tag == "ew_bwd":
                    _, mode, scale, entries = item
                    subs = []
                    for A, C, out in entries:
                        if not out.needs_grad:
                            continue
                        flags = 0
                        if A.needs_grad:
                            flags |= 1 | (0 if first_write(A) else 4)
                        if C is not None and C.needs_grad:
                            flags |= 2 | (0 if first_write(C) else 8)
                        subs.append(self._ew_rec(mode, scale, A, C, out, flags))
                    if subs:
                        bwd.append(self._hdr(N.OP_EW_BWD, len(subs)))
                        bwd.extend(subs)
                elif tag == "select_bwd":
                    _, ys, out, dom_dtype = item
                    if out.needs_grad:
                        for y in ys:
                            if not first_write(y):
                                raise NotImplementedError("select: an input consumed twice")
                        bwd.append(self._hdr(N.OP_SELECT_BWD, len(ys), i1=out.n, i2=out.ld, i3=N.torch_dtype_code(dom_dtype),
                                             s0=self.input("domain_indicator"), s1=out.raw, s2=out.dz))
                        bwd.extend(self._act_rec(y) for y in ys)
                elif tag == "ln_bwd":
                    entries = item[1]
                    for e in entries:
                        if not first_write(e[0]):
                            raise NotImplementedError("layernorm: input consumed twice")
                    bwd.append(self._hdr(N.OP_LN_BWD, len(entries)))
                    bwd.extend(self._ln_rec(e) for e in entries)
                elif tag == "mix_bwd":
                    _, xs, pooled, w_exp, w_bal, red = item
                    for x in xs:
                        if not first_write(x):
                            raise NotImplementedError("mix: a domain expert output consumed twice")
                    subs = [self._plain_rec([x.raw, x.dz, o.raw, o.dz]) for x, o in zip(xs, pooled)]
                    bwd.append(self._hdr(N.OP_MIX_BWD, len(subs), i1=len(xs), i2=xs[0].n, i3=xs[0].ld, i4=pooled[0].ld,
                                         s0=self.static(w_exp), s1=self.static(w_bal), s2=self.grad(w_exp), s3=self.grad(w_bal), s4=red))
                    bwd.extend(subs)
                elif tag == "bmv_bwd":
                    _, entries, H, k = item
                    live = [(p_, q) for p_, q in entries if q.needs_grad]
                    if live:
                        for p_, _q in live:
                            if p_.needs_grad and not first_write(p_):
                                raise NotImplementedError("bmv: an input vector consumed twice")
                        acc = 0
                        if H.needs_grad:
                            acc = 0 if first_write(H) else 1
                        bwd.append(self._hdr(N.OP_BMV_BWD, len(live), i1=k, i2=H.ld, i3=acc, s0=H.raw, s1=(H.dz if H.needs_grad else -1)))
                        bwd.extend(self._plain_rec([p_.raw, p_.dz if p_.needs_grad else -1, q.raw, q.dz], [p_.ld, q.ld]) for p_, q in live)
        pg = [a for a in self.norm_acts if a.needs_grad and a.dstats >= 0 and
              (self.grad(a.norm.gamma) >= 0 or self.grad(a.norm.beta) >= 0 or
               self.grad(a.norm.gamma2) >= 0 or self.grad(a.norm.beta2) >= 0)]
        if pg:
            subs = []
            for a in pg:
                r = self._act_rec(a)
                r["s"][24], r["s"][25] = self.grad(a.norm.gamma), self.grad(a.norm.gamma2)
                r["s"][26], r["s"][27] = self.grad(a.norm.beta), self.grad(a.norm.beta2)
                subs.append(r)
            bwd.append(self._hdr(N.OP_BN_PGRAD, len(subs)))
            bwd.extend(subs)
        # the scatter blocks were emitted in place (they are np records inside bwd_blocks) -- they run after every
        # writer of x.dz because the gather is the first tape entry of its activation.

        def stack(lst):
            return np.stack(lst).astype(N.REC_DTYPE) if lst else np.zeros((0,), dtype=N.REC_DTYPE)

        def launches(lst):
            return sum(1 for r in lst if int(r["kind"]) != N.OP_GROUP)

        return Program(B=B, training=self.training, recs_fwd=stack(fwd), recs_bwd=stack(bwd),
                       slot_desc=self.slot_desc, ws32=self.ws32, ws64=self.ws64, inputs=self.inputs,
                       params=self.params, param_arena=self.param_arena, arena_size=dict(self.arena_size),
                       out_slot=self.out_slot, gout_slot=self.gout_slot, oob_slot=self.oob_slot,
                       outputs=[(a.raw, a.ld, a.n, a.dz if a.needs_grad else -1) for a in self._outputs],
                       labels=dict(self.labels), virtual_fields=list(self.virtual_fields), peer_tabs=list(self.peer_tabs.values()),
                       n_launch_fwd=launches(fwd), n_launch_bwd=launches(bwd))


# ------------------------------------------------------------------------------------------
# execution on the device (the product path)
# ------------------------------------------------------------------------------------------
class CudaRunner:
    """Owns the workspace of one Program on one CUDA device and runs it through the C ABI."""

    def __init__(self, prog: Program, device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("scenario_wise_rec_b200 runs on CUDA devices only (no CPU fallback); "
                               f"got device {device}")
        N.lib()     # raises if the extension is not built
        self.prog, self.device = prog, device
        self.ws32 = torch.zeros(max(prog.ws32, 1), dtype=torch.float32, device=device)
        self.ws64 = torch.zeros(max(prog.ws64, 1), dtype=torch.float64, device=device)
        # out-of-range index flag: mapped pinned host memory, readable without a device sync
        self.oob = torch.zeros(2, dtype=torch.int32).pin_memory()
        self.ptrs = np.zeros(len(prog.slot_desc), dtype=np.uint64)
        self._grad_slots = {k: ([], []) for k in prog.arena_size}
        b32, b64 = self.ws32.data_ptr(), self.ws64.data_ptr()
        for i, d in enumerate(prog.slot_desc):
            if d[0] == "static":
                t = d[1]
                if t.device != device:
                    raise RuntimeError(f"parameter on {t.device}, program on {device}: call model.to(device) first")
                if t.dtype not in (torch.float32, torch.int64) or not t.is_contiguous():
                    raise RuntimeError("parameters must be contiguous float32")
                self.ptrs[i] = t.data_ptr()
            elif d[0] == "ws32":
                self.ptrs[i] = b32 + 4 * d[1]
            elif d[0] == "ws64":
                self.ptrs[i] = b64 + 8 * d[1]
            elif d[0] == "grad":
                self._grad_slots[d[1]][0].append(i)
                self._grad_slots[d[1]][1].append(4 * d[2])
            elif d[0] == "special" and d[1] == "oob":
                self.ptrs[i] = self.oob.data_ptr()
        self._grad_idx = {k: (np.array(v[0], dtype=np.int64), np.array(v[1], dtype=np.uint64)) for k, v in self._grad_slots.items()}
        self.generation = 0
        self._keep = None

    def _bind_inputs(self, x: Dict[str, torch.Tensor]):
        keep = []
        for name, slot in self.prog.inputs.items():
            if name == "__grad_out__":
                continue
            t = x[name]              # KeyError for a missing feature column, like the reference
            if t.device != self.device:
                t = t.to(self.device, non_blocking=True)
            if not t.is_contiguous():
                t = t.contiguous()
            if t.dim() != 1 or t.shape[0] != self.prog.B:
                raise ValueError(f"feature column {name!r} has shape {tuple(t.shape)}, expected ({self.prog.B},)")
            keep.append(t)
            self.ptrs[slot] = t.data_ptr()
        self._keep = keep

    def check_indices(self):
        """Raise IndexError if a gather of an earlier forward saw an out-of-range index (call after a sync)."""
        if int(self.oob[0]) != 0:
            f = int(self.oob[1])
            self.oob.zero_()
            raise IndexError(f"index out of range in self (sparse field #{f})")

    def _view(self, slot: int, ld: int, n: int) -> torch.Tensor:
        o = self.prog.slot_desc[slot]
        return self.ws32[o[1]:o[1] + self.prog.B * ld].view(self.prog.B, ld)[:, :n]

    def forward(self, x: Dict[str, torch.Tensor]) -> Tuple[torch.Tensor, ...]:
        """Returns (head output [B],) if the program has a head, followed by the plain outputs [B, n]."""
        self.check_indices()
        self._bind_inputs(x)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        N.program_run(self.prog.recs_fwd, self.ptrs, stream)
        self.generation += 1
        outs = []
        if self.prog.out_slot >= 0:
            o = self.prog.slot_desc[self.prog.out_slot]
            outs.append(self.ws32[o[1]:o[1] + self.prog.B].clone())
        for raw, ld, n, _dz in self.prog.outputs:
            outs.append(self._view(raw, ld, n).clone())
        return tuple(outs)

    def backward(self, gouts: Sequence[Optional[torch.Tensor]]) -> List[torch.Tensor]:
        prog = self.prog
        gouts = list(gouts)
        arenas = {k: torch.zeros(max(v, 1), dtype=torch.float32, device=self.device) for k, v in prog.arena_size.items()}
        for k, (idx, off) in self._grad_idx.items():
            if idx.size:
                self.ptrs[idx] = np.uint64(arenas[k].data_ptr()) + off
        if prog.out_slot >= 0:
            g = gouts.pop(0)
            g = torch.zeros(prog.B, device=self.device) if g is None else g.contiguous()
            self.ptrs[prog.gout_slot] = g.data_ptr()
        for (raw, ld, n, dz), g in zip(prog.outputs, gouts):
            if dz >= 0:
                v = self._view(dz, ld, n)
                v.zero_() if g is None else v.copy_(g)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        N.program_run(prog.recs_bwd, self.ptrs, stream)
        if getattr(self, "grad_sync", None) is not None:
            self.grad_sync(arenas)
        return [arenas[a][off:off + n].view(p.shape) for p, (a, off, n) in zip(prog.params, prog.param_arena)]
