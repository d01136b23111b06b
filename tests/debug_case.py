"""Development aid: run one BASELINE-shaped case on the CUDA path and on the CPU interpreter (same records),
report every workspace buffer / parameter gradient that differs.  usage: python tests/debug_case.py <case> [B]"""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scenario-wise-rec_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
import workloads, model_factory, gpu_util
from oracle.ops_ref import RefRunner
from scenario_wise_rec_b200.program import ProgramBuilder

case = sys.argv[1]
model_name, cfg, B = workloads.CASES[case]
B = int(sys.argv[2]) if len(sys.argv) > 2 else B
torch.manual_seed(7)
model = model_factory.build(model_name, cfg)
gpu_util.randomise(model, 11)
model.train()
ref = copy.deepcopy(model)
x, y = gpu_util.make_batch(workloads.all_feature_specs(cfg), B, cfg["domain_num"], seed=5, zipf=True)
cols = ref._columns()
b = ProgramBuilder(B, True); ref._lower(b, {c: x[c].dtype for c in cols}); rr = RefRunner(b.finish())
model.to("cuda:0")
xg = {k: v.to("cuda:0") for k, v in x.items()}
rc = model._runner(xg)
oc = rc.forward(xg); orf = rr.forward(x)
torch.cuda.synchronize()
rep = []
gpu_util.compare_slots(rc, rr, [".raw", ".stats", ".probs", "head.out", ".rowstats"], report=rep)
pr = orf[0].clamp(1e-6, 1 - 1e-6)
gout = (pr - y) / (pr * (1 - pr)) / B          # d BCELoss / d prediction
gc = rc.backward([gout.to("cuda:0")] + [None] * (len(oc) - 1)); gr = rr.backward([gout] + [None] * (len(orf) - 1))
torch.cuda.synchronize()
gpu_util.compare_slots(rc, rr, [".dz", ".dstats"], report=rep)
rep.sort(key=lambda t: -(t[1] / t[2]))
print("worst buffers (label, abs err, scale):")
for lab, err, scale, ok in rep[:12]:
    print(f"  {lab:28s} {err:.3e} {scale:.3e} rel {err / scale:.2e}")
names = {id(p): n for n, p in ref.named_parameters()}
res = []
for p, a, b_ in zip(rr.prog.params, gc, gr):
    scale = max(float(b_.abs().max()), 1e-9)
    res.append((float((a.cpu() - b_).abs().max()) / scale, names.get(id(p), "?"), scale))
res.sort(reverse=True)
print("worst parameter gradients (rel to max, name, max|g|):")
for r in res[:10]:
    print("  %.3e %s %.3e" % r)

# element-level detail for the worst buffers
ws32 = rc.ws32.cpu()
for lab, err, scale, ok in rep[:4]:
    slot = [s for s, l in rc.prog.labels.items() if l == lab][0]
    d = rc.prog.slot_desc[slot]
    if d[0] != "ws32":
        continue
    a, b_ = ws32[d[1]:d[1] + d[2]], rr.ws32[d[1]:d[1] + d[2]]
    diff = (a - b_).abs()
    i = int(diff.argmax())
    n_bad = int((diff > 1e-3 * scale).sum())
    ld = [int(x_) for x_ in lab.split("[")[1].split("]")[0:1]][0]
    ld = (ld + 3) // 4 * 4
    rows = sorted(set((diff > 1e-3 * scale).nonzero().flatten().div(ld, rounding_mode="floor").tolist()))
    print("   rows affected:", rows[:20], "n rows", len(rows))
    print(f"{lab}: argmax flat {i}  cuda {float(a[i]):.6e} ref {float(b_[i]):.6e}  elements with err > 1e-3*scale: {n_bad} of {a.numel()}")
