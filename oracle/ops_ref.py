"""torch-CPU interpreter of the device program records.  TEST INFRASTRUCTURE ONLY.

``RefRunner`` executes the *same* ``swr_rec_t`` records the C-ABI executor
(scenario-wise-rec_b200/csrc/swr_exec.cu) decodes, with the per-op forward and the
hand-derived backward written out in plain torch (no autograd).  It is used

* on CPU (``-m "not gpu"``): model program  ->  RefRunner  ==  reference golden vectors,
  which checks the host-side lowering and the backward derivations without a GPU;
* on the GPU box: every CUDA kernel is compared against the op it mirrors here.

It is never imported by the product package.

Record layouts follow include/swr_b200.h / DESIGN.md "Program records".
"""
from __future__ import annotations

import numpy as np
import torch

OP_ZERO, OP_GATHER, OP_SCATTER, OP_COLSTATS = 1, 2, 3, 4
OP_FC_FWD, OP_FC_DGRAD, OP_FC_WGRAD = 5, 6, 7
OP_POOL_FWD, OP_POOL_BWD, OP_HEAD_FWD, OP_HEAD_BWD = 8, 9, 10, 11
OP_BN_UPDATE, OP_BN_PGRAD, OP_GROUP = 12, 13, 100
OP_EW_FWD, OP_EW_BWD, OP_SUMGRAD, OP_SELECT_FWD, OP_SELECT_BWD = 14, 15, 16, 17, 18
OP_LN_FWD, OP_LN_BWD, OP_MIX_FWD, OP_MIX_BWD, OP_BMV_FWD, OP_BMV_BWD = 19, 20, 21, 22, 23, 24
EW_MUL, EW_ADD, EW_COPY = 0, 1, 2
NORM_NONE, NORM_BATCH, NORM_RUNNING = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_LEAKY = 0, 1, 2, 3


def act_fwd(z, act):
    if act == ACT_RELU:
        return torch.relu(z)
    if act == ACT_SIGMOID:
        return torch.sigmoid(z)
    if act == ACT_LEAKY:
        return torch.where(z > 0, z, 0.1 * z)
    return z


def act_grad(z, act):
    if act == ACT_RELU:
        return (z > 0).to(z.dtype)
    if act == ACT_SIGMOID:
        a = torch.sigmoid(z)
        return a * (1 - a)
    if act == ACT_LEAKY:
        return torch.where(z > 0, torch.ones_like(z), torch.full_like(z, 0.1))
    return torch.ones_like(z)


class _Act:
    """Decoded ActRef: tensors are [B, n] views into the slot storage."""

    def __init__(self, run, rec, sb, ib, fb):
        s, i, f = rec["s"], rec["i"], rec["f"]
        self.ld, self.n, self.mode, self.act = int(i[ib]), int(i[ib + 1]), int(i[ib + 2]), int(i[ib + 3])
        self.eps = float(f[fb])
        self.var_scale = float(f[fb + 1])
        B = run.prog.B
        # strided view: a column sub-view starts inside a row, so its last row is shorter than ld
        v2 = lambda slot: None if slot < 0 else torch.as_strided(run.slot(slot), (B, self.n), (self.ld, 1))   # noqa: E731
        v1 = lambda slot: None if slot < 0 else run.slot(slot).reshape(-1)[:self.n]                           # noqa: E731
        self.raw = v2(int(s[sb]))
        self.stats = None if s[sb + 1] < 0 else run.slot(int(s[sb + 1]))[:2 * self.n].view(self.n, 2)
        self.rmean, self.rvar = v1(int(s[sb + 2])), v1(int(s[sb + 3]))
        self.gamma, self.gamma2 = v1(int(s[sb + 4])), v1(int(s[sb + 5]))
        self.beta, self.beta2 = v1(int(s[sb + 6])), v1(int(s[sb + 7]))
        self.dz = v2(int(s[sb + 8]))
        self.dstats = None if s[sb + 9] < 0 else run.slot(int(s[sb + 9]))[:2 * self.n].view(self.n, 2)
        self.B = B

    # per-column coefficients (csrc/swr_common.cuh col_coef)
    def coef(self):
        n = self.n
        if self.mode == NORM_NONE:
            z = torch.zeros(n)
            return z, torch.ones(n), z.clone(), torch.ones(n)
        if self.mode == NORM_BATCH:
            mu = self.stats[:, 0] / self.B
            var = (self.stats[:, 1] / self.B - mu * mu).clamp_min(0.0) * self.var_scale
        else:
            mu, var = self.rmean.double(), self.rvar.double()
        r = (1.0 / torch.sqrt(var + self.eps)).float()
        g = torch.ones(n)
        if self.gamma is not None:
            g = g * self.gamma
        if self.gamma2 is not None:
            g = g * self.gamma2
        b = torch.zeros(n)
        if self.beta is not None:
            b = b + self.beta
        if self.beta2 is not None:
            b = b + self.beta2
        return mu.float(), g * r, b, r

    def z(self):
        mu, s, b, _ = self.coef()
        return (self.raw - mu) * s + b

    def value(self):
        return act_fwd(self.z(), self.act)

    def xhat(self):
        mu, _, _, r = self.coef()
        return (self.raw - mu) * r

    def write_grad(self, dA, accumulate=False):
        """stage 1: dz = dA * act'(z); accumulate the column sums used by stage 2 / d gamma, d beta."""
        plain = self.mode == NORM_NONE and self.act == ACT_NONE
        dz = dA if plain else dA * act_grad(self.z(), self.act)
        if self.mode != NORM_NONE and self.dstats is not None:
            self.dstats[:, 0] += dz.double().sum(0)
            self.dstats[:, 1] += (dz.double() * self.xhat().double()).sum(0)
        if self.dz is not None:
            if accumulate:
                self.dz += dz
            else:
                self.dz.copy_(dz)

    def dy(self):
        """stage 2: gradient wrt the raw tensor from dz (csrc/swr_common.cuh dy_coef)."""
        if self.mode == NORM_NONE:
            return self.dz
        mu, s, _, r = self.coef()
        if self.mode == NORM_RUNNING:
            return self.dz * s
        S1, S2 = self.dstats[:, 0], self.dstats[:, 1]
        sd, rd, mud, ib = s.double(), r.double(), mu.double(), 1.0 / self.B
        ibv = ib * self.var_scale
        c1 = (-sd * rd * S2 * ibv).float()
        c2 = (-sd * S1 * ib + sd * rd * S2 * mud * ibv).float()
        return s * self.dz + c1 * self.raw + c2


class RefRunner:
    """Drop-in for program.CudaRunner on CPU tensors."""

    def __init__(self, prog):
        self.prog = prog
        self.ws32 = torch.zeros(max(prog.ws32, 1), dtype=torch.float32)
        self.ws64 = torch.zeros(max(prog.ws64, 1), dtype=torch.float64)
        self.oob = torch.zeros(2, dtype=torch.int32)
        self.tensors = [None] * len(prog.slot_desc)
        for i, d in enumerate(prog.slot_desc):
            if d[0] == "static":
                self.tensors[i] = d[1].detach()
            elif d[0] == "ws32":
                self.tensors[i] = self.ws32[d[1]:d[1] + d[2]]
            elif d[0] == "ws64":
                self.tensors[i] = self.ws64[d[1]:d[1] + d[2]]
            elif d[0] == "special":
                self.tensors[i] = self.oob
        self.generation = 0

    def slot(self, i):
        t = self.tensors[i]
        assert t is not None, f"slot {i} ({self.prog.slot_desc[i][:2]}) is unbound"
        return t

    # ---- public API (mirrors CudaRunner) ------------------------------------------------
    def check_indices(self):
        if int(self.oob[0]) != 0:
            f = int(self.oob[1])
            self.oob.zero_()
            raise IndexError(f"index out of range in self (sparse field #{f})")

    def forward(self, x):
        for name, slot in self.prog.inputs.items():
            if name != "__grad_out__":
                self.tensors[slot] = x[name]
        self._run(self.prog.recs_fwd)
        self.generation += 1
        outs = []
        if self.prog.out_slot >= 0:
            o = self.prog.slot_desc[self.prog.out_slot]
            outs.append(self.ws32[o[1]:o[1] + self.prog.B].clone())
        for raw, ld, n, _dz in self.prog.outputs:
            outs.append(self._view(raw, ld, n).clone())
        return tuple(outs)

    def _view(self, slot, ld, n):
        B = self.prog.B
        return self.slot(slot)[:B * ld].view(B, ld)[:, :n]

    def backward(self, gouts):
        prog = self.prog
        gouts = list(gouts)
        arenas = {k: torch.zeros(max(v, 1), dtype=torch.float32) for k, v in prog.arena_size.items()}
        for i, d in enumerate(prog.slot_desc):
            if d[0] == "grad":
                self.tensors[i] = arenas[d[1]][d[2]:d[2] + d[3]]
        if prog.out_slot >= 0:
            g = gouts.pop(0)
            self.tensors[prog.gout_slot] = torch.zeros(prog.B) if g is None else g.contiguous()
        for (raw, ld, n, dz), g in zip(prog.outputs, gouts):
            if dz >= 0:
                v = self._view(dz, ld, n)
                v.zero_() if g is None else v.copy_(g)
        self._run(prog.recs_bwd)
        if getattr(self, "grad_sync", None) is not None:      # data parallel: mirrors CudaRunner.backward
            self.grad_sync(arenas)
        return [arenas[a][off:off + n].view(p.shape) for p, (a, off, n) in zip(prog.params, prog.param_arena)]

    # ---- interpreter ------------------------------------------------------------------------
    def _run(self, recs):
        i = 0
        while i < len(recs):
            h = recs[i]
            subs = recs[i + 1:i + 1 + int(h["n_sub"])]
            assert int(h["kind"]) != OP_GROUP and all(int(r["kind"]) == OP_GROUP for r in subs)
            getattr(self, "_op_%d" % int(h["kind"]))(h, subs)
            i += 1 + int(h["n_sub"])

    @staticmethod
    def _i64(lo, hi):
        return (int(lo) & 0xFFFFFFFF) | (int(hi) << 32)

    def _op_1(self, h, subs):       # ZERO
        t = self.slot(int(h["s"][0]))
        nbytes = self._i64(h["i"][0], h["i"][1])
        base = self.ws64 if t.dtype == torch.float64 else self.ws32
        off = t.storage_offset()
        base[off:off + nbytes // t.element_size()].zero_()

    def _op_2(self, h, subs):       # GATHER
        B, ld = int(h["i"][0]), int(h["i"][4])
        out = self.slot(int(h["s"][0]))[:B * ld].view(B, ld)      # sub-records carry absolute columns
        for f, r in enumerate(subs):
            col = int(r["i"][3])
            if int(r["i"][4]) == 0:
                tab = self.slot(int(r["s"][0]))
                idx = self.slot(int(r["s"][1])).long()
                vocab, E = self._i64(r["i"][0], r["i"][1]), int(r["i"][5])
                bad = (idx < 0) | (idx >= vocab)
                if bool(bad.any()):
                    self.oob[0], self.oob[1] = 1, f
                rows = tab[idx.clamp(0, vocab - 1)]
                rows = torch.where(bad.unsqueeze(1), torch.zeros_like(rows), rows)
                out[:, col:col + E] = rows
            else:
                out[:, col] = self.slot(int(r["s"][0])).float()

    def _op_3(self, h, subs):       # SCATTER
        B, ld = int(h["i"][0]), int(h["i"][4])
        g = self.slot(int(h["s"][0]))[:B * ld].view(B, ld)
        for r in subs:
            vocab, E, col = self._i64(r["i"][0], r["i"][1]), int(r["i"][5]), int(r["i"][3])
            gt = self.slot(int(r["s"][0]))[:vocab * E].view(vocab, E)
            idx = self.slot(int(r["s"][1])).long()
            ok = (idx >= 0) & (idx < vocab)
            gt.index_add_(0, idx[ok], g[ok, col:col + E])

    def _op_4(self, h, subs):       # COLSTATS
        B, n, ld = int(h["i"][0]), int(h["i"][1]), int(h["i"][2])
        x = self.slot(int(h["s"][0]))[:B * ld].view(B, ld)[:, :n].double()
        st = self.slot(int(h["s"][1]))[:2 * n].view(n, 2)
        st[:, 0] += x.sum(0)
        st[:, 1] += (x * x).sum(0)

    # -- fully connected ---------------------------------------------------------------------
    def _weff(self, r):
        s = r["s"]
        W = self.slot(int(s[24]))
        if s[25] >= 0:
            W = W * self.slot(int(s[25]))
        if int(r["i"][8]) == 1:      # KN -> [N, K]
            W = W.t()
        return W

    def _op_5(self, h, subs):       # FC_FWD
        for r in subs:
            A, Y = _Act(self, r, 0, 0, 0), _Act(self, r, 12, 4, 2)
            s = r["s"]
            y = A.value() @ self._weff(r).t()
            if s[26] >= 0:
                y = y + self.slot(int(s[26]))
            if s[27] >= 0:
                y = y + self.slot(int(s[27]))
            if int(r["i"][10]) != ACT_NONE:
                y = act_fwd(y, int(r["i"][10])) * float(r["f"][4])
            Y.raw.copy_(y)
            if Y.mode == NORM_BATCH:
                Y.stats[:, 0] += y.double().sum(0)
                Y.stats[:, 1] += (y.double() ** 2).sum(0)

    def _op_6(self, h, subs):       # FC_DGRAD (fan-in per destination; i[12] = destination index)
        by_dst = {}
        for r in subs:
            by_dst.setdefault(int(r["i"][12]), []).append(r)
        for recs in by_dst.values():
            dA = None
            for r in recs:
                Y = _Act(self, r, 12, 4, 2)
                t = Y.dy() @ self._weff(r)
                dA = t if dA is None else dA + t
            D = _Act(self, recs[0], 0, 0, 0)
            D.write_grad(dA[:, :D.n], accumulate=bool(int(recs[0]["i"][11]) & 2))    # D.n < K: detached trailing columns

    def _op_7(self, h, subs):       # FC_WGRAD
        for r in subs:
            A, Y = _Act(self, r, 0, 0, 0), _Act(self, r, 12, 4, 2)
            s = r["s"]
            dy = Y.dy()
            dW = dy.t() @ A.value()                     # [N, K]
            kn = int(r["i"][8]) == 1
            if kn:
                dW = dW.t()
            W, W2 = self.slot(int(s[24])), (self.slot(int(s[25])) if s[25] >= 0 else None)
            if W2 is not None:
                if s[28] >= 0:
                    self.slot(int(s[28])).view(W.shape).add_(dW * W2)
                if s[29] >= 0:
                    self.slot(int(s[29])).view(W.shape).add_(dW * W)
            elif s[28] >= 0:
                self.slot(int(s[28])).view(W.shape).add_(dW)
            db = dy.sum(0)
            if s[30] >= 0:
                self.slot(int(s[30])).add_(db)
            if s[31] >= 0:
                self.slot(int(s[31])).add_(db)

    # -- pooling -------------------------------------------------------------------------------
    def _pool_decode(self, h, subs):
        ng, ne = int(h["i"][2]), int(h["i"][3])
        gates = []
        for r in subs[:ng]:
            nE = int(r["i"][8])
            gate, out = _Act(self, r, 0, 0, 0), _Act(self, r, 12, 4, 2)
            probs = self.slot(int(r["s"][24]))[:self.prog.B * nE].view(self.prog.B, nE)
            gates.append((gate, out, probs, [int(v) for v in r["i"][16:16 + nE]]))
        experts = [_Act(self, r, 0, 0, 0) for r in subs[ng:ng + ne]]
        return gates, experts

    def _op_8(self, h, subs):       # POOL_FWD
        gates, experts = self._pool_decode(h, subs)
        vals = [e.value() for e in experts]
        for gate, out, probs, idx in gates:
            p = torch.softmax(gate.z(), dim=1)
            probs.copy_(p)
            out.raw.copy_(sum(p[:, e:e + 1] * vals[u] for e, u in enumerate(idx)))

    def _op_9(self, h, subs):       # POOL_BWD
        gates, experts = self._pool_decode(h, subs)
        vals = [e.value() for e in experts]
        dA = [torch.zeros_like(v) for v in vals]
        for gate, out, probs, idx in gates:
            dP = out.dz
            dp = torch.stack([(dP * vals[u]).sum(1) for u in idx], dim=1)
            dot = (probs * dp).sum(1, keepdim=True)
            gate.write_grad(probs * (dp - dot))          # gate act is NONE: dz = d logits
            for e, u in enumerate(idx):
                dA[u] += probs[:, e:e + 1] * dP
        for e, d in zip(experts, dA):
            e.write_grad(d)

    # -- head -----------------------------------------------------------------------------------
    def _op_10(self, h, subs):      # HEAD_FWD
        B = self.prog.B
        dom = self.slot(int(h["s"][0])).long() if h["s"][0] >= 0 else None
        out = self.slot(int(h["s"][1]))[:B]
        mode = int(h["i"][2])
        sbs = mode != 0
        if mode == 2:
            dom = torch.zeros(B, dtype=torch.long)
        v = torch.zeros(B)
        sel_any = torch.zeros(B, dtype=torch.bool)
        for d, r in enumerate(subs):
            A = _Act(self, r, 0, 0, 0)
            a = A.value()
            if r["s"][24] >= 0:
                vd = a @ self.slot(int(r["s"][24])).reshape(-1)
                if r["s"][26] >= 0:
                    vd = vd + self.slot(int(r["s"][26])).reshape(-1)[0]
            else:
                vd = a[:, 0]
            m = dom == d
            v = torch.where(m, vd, v)
            sel_any |= m
        if sbs:
            out.copy_(torch.where(sel_any, torch.sigmoid(v), torch.zeros(B)))
        else:
            add = self._plain(h["s"][3], max(int(h["i"][4]), 1), 1)[:, 0] if h["s"][3] >= 0 else 0.0
            out.copy_(torch.sigmoid(v + add))

    def _op_11(self, h, subs):      # HEAD_BWD
        B = self.prog.B
        dom = self.slot(int(h["s"][0])).long() if h["s"][0] >= 0 else None
        y = self.slot(int(h["s"][1]))[:B]
        g = self.slot(int(h["s"][2]))[:B]
        mode = int(h["i"][2])
        sbs = mode != 0
        if mode == 2:
            dom = torch.zeros(B, dtype=torch.long)
        dsig = g * y * (1 - y)
        if not sbs and h["s"][4] >= 0:
            self._plain(h["s"][4], max(int(h["i"][4]), 1), 1)[:, 0].copy_(dsig)
        for d, r in enumerate(subs):
            A = _Act(self, r, 0, 0, 0)
            dv = torch.where(dom == d, dsig, torch.zeros(B))
            if r["s"][24] >= 0:
                w = self.slot(int(r["s"][24])).reshape(-1)
                if r["s"][28] >= 0:
                    self.slot(int(r["s"][28])).reshape(-1).add_(dv @ A.value())
                if r["s"][30] >= 0:
                    self.slot(int(r["s"][30])).reshape(-1)[0] += dv.sum()
                dA = dv.unsqueeze(1) * w.unsqueeze(0)
            else:
                dA = dv.unsqueeze(1)
            A.write_grad(dA)

    # -- batch-norm bookkeeping ----------------------------------------------------------------------
    def _op_12(self, h, subs):      # BN_UPDATE
        B, m = self.prog.B, float(h["f"][4])
        for r in subs:
            A = _Act(self, r, 0, 0, 0)
            mu = A.stats[:, 0] / B
            var = (A.stats[:, 1] / B - mu * mu).clamp_min(0.0) * (B / (B - 1) if B > 1 else 1.0)
            rep = max(int(r["i"][8]), 1)
            for _ in range(rep):
                A.rmean.copy_((1 - m) * A.rmean + m * mu.float())
                A.rvar.copy_((1 - m) * A.rvar + m * var.float())
            if r["s"][24] >= 0:
                self.slot(int(r["s"][24])).add_(rep)

    def _op_13(self, h, subs):      # BN_PGRAD
        for r in subs:
            A = _Act(self, r, 0, 0, 0)
            s1, s2 = A.dstats[:, 0].float(), A.dstats[:, 1].float()
            g1 = A.gamma if A.gamma is not None else torch.ones(A.n)
            g2 = A.gamma2 if A.gamma2 is not None else torch.ones(A.n)
            s = r["s"]
            if s[24] >= 0:
                self.slot(int(s[24])).add_(s2 * g2)
            if s[25] >= 0:
                self.slot(int(s[25])).add_(s2 * g1)
            if s[26] >= 0:
                self.slot(int(s[26])).add_(s1)
            if s[27] >= 0:
                self.slot(int(s[27])).add_(s1)

    # -- element-wise glue -----------------------------------------------------------------------------
    def _plain(self, slot, ld, n):
        return None if slot < 0 else torch.as_strided(self.slot(int(slot)), (self.prog.B, n), (ld, 1))

    def _ew_decode(self, r):
        mode = int(r["i"][9])
        A = _Act(self, r, 0, 0, 0)
        C = _Act(self, r, 12, 4, 2) if mode != EW_COPY else None
        ld = int(r["i"][8])
        return mode, A, C, self._plain(r["s"][24], ld, A.n), self._plain(r["s"][25], ld, A.n), float(r["f"][4]), int(r["i"][10])

    def _op_14(self, h, subs):      # EW_FWD
        for r in subs:
            mode, A, C, out, _dout, scale, _fl = self._ew_decode(r)
            if mode == EW_MUL:
                out.copy_(A.value() * C.value() * scale)
            elif mode == EW_ADD:
                out.copy_(A.value() + C.value())
            else:
                out.copy_(A.value())

    def _op_15(self, h, subs):      # EW_BWD
        for r in subs:
            mode, A, C, _out, dout, scale, fl = self._ew_decode(r)
            if mode == EW_MUL:
                dA, dC = dout * C.value() * scale, dout * A.value() * scale
            else:
                dA, dC = dout, dout
            if fl & 1:
                A.write_grad(dA, accumulate=bool(fl & 4))
            if C is not None and fl & 2:
                C.write_grad(dC, accumulate=bool(fl & 8))

    def _op_16(self, h, subs):      # SUMGRAD
        dst = _Act(self, subs[0], 0, 0, 0)
        total = sum(_Act(self, r, 0, 0, 0).dy() for r in subs[1:])
        if int(h["i"][1]):
            dst.dz += total
        else:
            dst.dz.copy_(total)

    def _select_decode(self, h, subs):
        n, ld = int(h["i"][1]), int(h["i"][2])
        dom = self.slot(int(h["s"][0])).long()
        return [_Act(self, r, 0, 0, 0) for r in subs], dom, self._plain(h["s"][1], ld, n), self._plain(h["s"][2], ld, n)

    def _op_17(self, h, subs):      # SELECT_FWD
        ys, dom, out, _ = self._select_decode(h, subs)
        out.zero_()
        for d, y in enumerate(ys):
            out.copy_(torch.where((dom == d).unsqueeze(1), y.value(), out))

    def _op_18(self, h, subs):      # SELECT_BWD
        ys, dom, _out, dout = self._select_decode(h, subs)
        for d, y in enumerate(ys):
            y.write_grad(torch.where((dom == d).unsqueeze(1), dout, torch.zeros_like(dout)))

    def _ln_decode(self, r):
        s, i = r["s"], r["i"]
        ld_y, n, ld_o, act = int(i[0]), int(i[1]), int(i[2]), int(i[3])
        # gradient slots are unbound during the forward pass
        vec = lambda k: None if (s[k] < 0 or self.tensors[int(s[k])] is None) else self.slot(int(s[k])).reshape(-1)[:n]      # noqa: E731
        rs = self.slot(int(s[8]))[:2 * self.prog.B].view(self.prog.B, 2)
        return (self._plain(s[0], ld_y, n), self._plain(s[1], ld_y, n), vec(2), vec(3), vec(4), vec(5),
                self._plain(s[6], ld_o, n), self._plain(s[7], ld_o, n), rs, act, float(r["f"][0]))

    def _op_19(self, h, subs):      # LN_FWD
        for r in subs:
            y, _dy, gamma, beta, _dg, _db, out, _dout, rs, act, eps = self._ln_decode(r)
            mean = y.mean(dim=1)
            rstd = 1.0 / torch.sqrt(((y - mean.unsqueeze(1)) ** 2).mean(dim=1) + eps)
            rs[:, 0], rs[:, 1] = mean, rstd
            out.copy_(act_fwd((y - mean.unsqueeze(1)) * rstd.unsqueeze(1) * gamma + beta, act))

    def _op_20(self, h, subs):      # LN_BWD
        for r in subs:
            y, dy, gamma, beta, dgam, dbet, _out, dout, rs, act, _eps = self._ln_decode(r)
            xh = (y - rs[:, 0:1]) * rs[:, 1:2]
            dz = dout * act_grad(xh * gamma + beta, act)
            if dgam is not None:
                dgam.add_((dz * xh).sum(0))
            if dbet is not None:
                dbet.add_(dz.sum(0))
            g = dz * gamma
            if dy is not None:
                dy.copy_(rs[:, 1:2] * (g - g.mean(dim=1, keepdim=True) - xh * (g * xh).mean(dim=1, keepdim=True)))

    def _mix_decode(self, h, subs):
        D, H, ldx, ldo = int(h["i"][1]), int(h["i"][2]), int(h["i"][3]), int(h["i"][4])
        X = [self._plain(r["s"][0], ldx, H) for r in subs]
        dX = [self._plain(r["s"][1], ldx, H) for r in subs]
        O = [self._plain(r["s"][2], ldo, H) for r in subs]
        dO = [self._plain(r["s"][3], ldo, H) for r in subs]
        se = torch.sigmoid(self.slot(int(h["s"][0])).reshape(-1)[0])
        sb = torch.sigmoid(self.slot(int(h["s"][1])).reshape(-1)[0])
        cc = (1 - sb) / (D - 1) if D > 1 else torch.zeros(())
        return D, X, dX, O, dO, se, sb, cc

    def _op_21(self, h, subs):      # MIX_FWD
        D, X, _dX, O, _dO, se, sb, cc = self._mix_decode(h, subs)
        tot = sum(X)
        for d in range(D):
            O[d] += se * (sb * X[d] + cc * (tot - X[d]))

    def _op_22(self, h, subs):      # MIX_BWD
        D, X, dX, O, dO, se, sb, cc = self._mix_decode(h, subs)
        sx, sg = sum(X), sum(dO)
        p1 = sum((dO[d].double() * X[d].double()).sum() for d in range(D))
        p2 = sum((dO[d].double() * (sx - X[d]).double()).sum() for d in range(D))
        for d in range(D):
            if dX[d] is not None:
                dX[d].copy_(se * (sb * dO[d] + cc * (sg - dO[d])))
        dse = sb.double() * p1 + cc.double() * p2
        dsb = se.double() * (p1 - (p2 / (D - 1) if D > 1 else 0.0))
        if h["s"][2] >= 0:
            self.slot(int(h["s"][2])).reshape(-1)[0] += float(dse * se * (1 - se))
        if h["s"][3] >= 0:
            self.slot(int(h["s"][3])).reshape(-1)[0] += float(dsb * sb * (1 - sb))

    def _bmv_decode(self, h, subs):
        k, ldh = int(h["i"][1]), int(h["i"][2])
        B = self.prog.B
        H = self._plain(h["s"][0], ldh, k * k).reshape(B, k, k)
        dH = self._plain(h["s"][1], ldh, k * k)
        gs = []
        for r in subs:
            ldp, ldq = int(r["i"][0]), int(r["i"][1])
            gs.append((self._plain(r["s"][0], ldp, k), self._plain(r["s"][1], ldp, k),
                       self._plain(r["s"][2], ldq, k), self._plain(r["s"][3], ldq, k)))
        return k, H, dH, gs

    def _op_23(self, h, subs):      # BMV_FWD
        _k, H, _dH, gs = self._bmv_decode(h, subs)
        for p_, _dp, q, _dq in gs:
            q.copy_(torch.einsum("bi,bij->bj", p_, H))

    def _op_24(self, h, subs):      # BMV_BWD
        k, H, dH, gs = self._bmv_decode(h, subs)
        tot = None
        for p_, dp, _q, dq in gs:
            if dp is not None:
                dp.copy_(torch.einsum("bj,bij->bi", dq, H))
            t = torch.einsum("bi,bj->bij", p_, dq).reshape(self.prog.B, k * k)
            tot = t if tot is None else tot + t
        if dH is not None:
            if int(h["i"][3]):
                dH += tot
            else:
                dH.copy_(tot)

    # -- trainer-side ops (csrc/swr_train.cu) ---------------------------------------------------------------
    def _op_27(self, h, subs):      # FC_PRESPLIT: weight images of the tcgen05 kernels; the interpreter reads the weights themselves
        for r in subs:
            assert int(r["s"][10]) >= 0 and int(r["s"][24]) >= 0

    def _op_25(self, h, subs):      # BCE
        B, ring = int(h["i"][0]), max(int(h["i"][2]), 1)
        p = self.slot(int(h["s"][0])).reshape(-1)[:B]
        y = self.slot(int(h["s"][1])).reshape(-1)[:B].float()
        lp, l1p = torch.log(p).clamp_min(-100.0), torch.log(1 - p).clamp_min(-100.0)
        if h["s"][2] >= 0:
            self.slot(int(h["s"][2])).reshape(-1)[:B].copy_((p - y) / ((1 - p) * p).clamp_min(1e-12) / B)
        if h["s"][3] >= 0:
            idx = int(self.slot(int(h["s"][4])).reshape(-1)[0]) % ring if h["s"][4] >= 0 else 0
            self.slot(int(h["s"][3])).reshape(-1)[idx] = -(y * lp + (1 - y) * l1p).double().mean().float()

    def _op_26(self, h, subs):      # ADAM
        n = self._i64(h["i"][0], h["i"][1])
        p, g, m, v = (self.slot(int(h["s"][k])).reshape(-1)[:n] for k in range(4))
        step_size, b1, b2, eps, wd, ibc2 = (float(t) for t in self.slot(int(h["s"][4])).reshape(-1)[:6])
        gg = g + wd * p
        m.copy_(m + (1 - b1) * (gg - m))
        v.copy_(b2 * v + (1 - b2) * gg * gg)
        p.sub_(step_size * m / (v.sqrt() * ibc2 + eps))
        if int(h["i"][2]):
            g.zero_()


# ---- CPU stand-ins for parallel.local_gather / local_scatter (K1 / K2 on one field), used by the gloo tests ----
def ref_local_gather(table, idx, out):
    ok = (idx >= 0) & (idx < table.shape[0])
    out.zero_()
    out[ok] = table[idx[ok]]


def ref_local_scatter(grad_rows, idx, gtable):
    ok = (idx >= 0) & (idx < gtable.shape[0])
    gtable.index_add_(0, idx[ok], grad_rows[ok])
