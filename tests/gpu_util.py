"""Helpers for the -m gpu parity tests: run one Program on the CUDA path (C ABI) and on the
torch-CPU interpreter (oracle/ops_ref.RefRunner) and compare every workspace buffer."""
import copy

import numpy as np
import torch

from oracle.ops_ref import RefRunner
from scenario_wise_rec_b200.program import CudaRunner


def clone_model_to(model, device):
    m = copy.deepcopy(model)
    return m.to(device)


def compare_slots(cuda_runner: CudaRunner, ref_runner: RefRunner, which, atol=1e-5, rtol=1e-4, report=None):
    """Compare workspace slots whose label ends with one of ``which`` suffixes.  Returns list of failures."""
    prog = cuda_runner.prog
    fails = []
    ws32 = cuda_runner.ws32.cpu()
    ws64 = cuda_runner.ws64.cpu()
    for slot, label in sorted(prog.labels.items()):
        if not label.endswith(tuple(which)):
            continue
        d = prog.slot_desc[slot]
        if d[0] == "ws32":
            a, b = ws32[d[1]:d[1] + d[2]], ref_runner.ws32[d[1]:d[1] + d[2]]
        elif d[0] == "ws64":
            a, b = ws64[d[1]:d[1] + d[2]].float(), ref_runner.ws64[d[1]:d[1] + d[2]].float()
        else:
            continue
        scale = max(float(b.abs().max()), 1e-6)
        err = float((a - b).abs().max())
        ok = err <= atol + rtol * scale and bool(torch.isfinite(a).all())
        if report is not None:
            report.append((label, err, scale, ok))
        if not ok:
            fails.append((label, err, scale))
    return fails


def make_batch(feats, B, D, seed, device="cpu", zipf=False):
    g = torch.Generator().manual_seed(seed)
    x = {}
    for name, kind, vocab, _dim in feats:
        if kind == "sparse":
            if zipf:
                r = torch.rand(B, generator=g)
                x[name] = (vocab * r ** 3).long().clamp_(0, vocab - 1)
            else:
                x[name] = torch.randint(0, vocab, (B,), generator=g)
        else:
            x[name] = torch.rand(B, generator=g)
    x["domain_indicator"] = torch.randint(0, D, (B,), generator=g)
    y = (torch.rand(B, generator=g) < 0.3).float()
    return {k: v.to(device) for k, v in x.items()}, y.to(device)


def randomise(model, seed):
    """Non-trivial but well-conditioned state (mirrors oracle/make_golden.py::_randomise)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "embed_dict" in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
            elif name.endswith("deep_weights"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.5 + 0.3)
            elif p.dim() == 1:
                p.add_(torch.randn(p.shape, generator=g) * 0.1)
            elif name.startswith(("u.", "v.")):      # HAMUR u/v are all-ones; keep the adapter scale sane
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)
        for name, b in model.named_buffers():
            if name.endswith("running_mean"):
                b.copy_(torch.randn(b.shape, generator=g) * 0.1)
            elif name.endswith("running_var"):
                b.copy_(torch.rand(b.shape, generator=g) + 0.5)
