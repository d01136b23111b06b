"""Development aid: parameter-gradient errors of one BASELINE case under each FC arithmetic mode, vs the float64 oracle.
usage: python tests/tc_debug.py <case> [repeats]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scenario-wise-rec_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
import workloads, model_factory, gpu_util
from oracle import ref_models
from scenario_wise_rec_b200 import _native as N

case = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
model_name, cfg, B = workloads.CASES[case]
x, y = gpu_util.make_batch(workloads.all_feature_specs(cfg), B, cfg["domain_num"], seed=5, zipf=True)


def oracle(state, dt):
    st = {}
    for k, v in state.items():
        v = v.clone()
        if v.dtype.is_floating_point:
            v = v.to(dt)
            if "running_" not in k:
                v.requires_grad_(True)
        st[k] = v
    xx = {k: (v.to(dt) if v.dtype.is_floating_point else v) for k, v in x.items()}
    out = ref_models.forward(model_name, xx, st, cfg, training=True, bn_out={})
    ref_models.bce_loss(out, y.to(dt)).backward()
    return out.detach(), {k: v.grad for k, v in st.items() if v.requires_grad}


g64 = None
for mode, label in ((N.FC_SIMT, "ffma"), (N.FC_TC, "tc-all"), (N.FC_AUTO, "auto")):
    for rep in range(reps):
        N.set_fc_mode(mode)
        torch.manual_seed(7)
        model = model_factory.build(model_name, cfg)
        gpu_util.randomise(model, 11)
        state = {k: v.clone() for k, v in model.state_dict().items()}
        if g64 is None:
            ref, g64 = oracle(state, torch.float64)
        model.to("cuda:0").train()
        out = model({k: v.to("cuda:0") for k, v in x.items()})
        torch.nn.BCELoss()(out, y.to("cuda:0")).backward()
        torch.cuda.synchronize()
        errs = []
        for k, p in model.named_parameters():
            t = g64[k]
            if t is None:
                continue
            scale = max(float(t.abs().max()), 1e-3)
            errs.append((float((p.grad.cpu().double() - t).abs().max()) / scale, k))
        errs.sort(reverse=True)
        oerr = float((out.detach().cpu().double() - ref).abs().max())
        print(f"[{label} #{rep}] out err {oerr:.2e}; worst grads: " + ", ".join(f"{k}={e:.1e}" for e, k in errs[:6]))
