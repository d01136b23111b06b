"""Quick device-time probe of the cfg2 MMoE train step (development aid, not the bench)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scenario-wise-rec_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
import model_factory, gpu_util
from workloads import CASES as BASELINE_CASES
from oracle import ref_models
from scenario_wise_rec_b200 import _native as N

dev = "cuda:0"
case = sys.argv[1] if len(sys.argv) > 1 else "cfg2_mmoe_aliccp_b4096"
model_name, cfg, B = BASELINE_CASES[case]
torch.manual_seed(0)
model = model_factory.build(model_name, cfg).to(dev)
x, y = gpu_util.make_batch(cfg["features"], B, cfg["domain_num"], seed=1, device=dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5, fused=True)
crit = torch.nn.BCELoss()

def timeit(fn, n=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n

def fwd_eval():
    model.eval()
    with torch.no_grad(): model(x)
def fwd_bwd():
    model.train(); model.zero_grad(); crit(model(x), y).backward()
def step():
    model.train(); out = model(x); loss = crit(out, y); model.zero_grad(); loss.backward(); opt.step()

for name, fn in [("eval fwd", fwd_eval), ("fwd+bwd", fwd_bwd), ("train step", step)]:
    l0 = N.launch_count(); d, w = timeit(fn); l1 = N.launch_count()
    print(f"[ours] {case} {name}: device {d:.3f} ms  wall {w:.3f} ms  -> {B / d * 1e3:,.0f} samples/s  launches/iter {(l1 - l0) / 35:.1f}")

# stock eager PyTorch on the same GPU (the restated reference forward = same ATen ops the reference issues)
state = {k: v.detach().clone() for k, v in model.state_dict().items()}
leaf = {k: (v.requires_grad_(True) if v.dtype.is_floating_point and "running_" not in k else v) for k, v in state.items()}
params = [v for v in leaf.values() if v.requires_grad]
opt2 = torch.optim.Adam(params, lr=1e-3, weight_decay=1e-5)
def eager_step():
    out = ref_models.forward(model_name, x, leaf, cfg, training=True, bn_out={})
    loss = crit(out, y)
    for p in params: p.grad = None
    loss.backward(); opt2.step()
d, w = timeit(eager_step, n=10, warm=3)
print(f"[eager torch cuda] {case} train step: device {d:.3f} ms wall {w:.3f} ms -> {B / max(d, w) * 1e3:,.0f} samples/s")
