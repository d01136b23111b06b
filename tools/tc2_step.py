"""Development aid: one eager training step of a BASELINE case with SWR_TC_DEBUG=1 (per-role clock stamps of CTA 0 of
every tcgen05 forward / data-gradient / weight-gradient launch go to stderr)."""
import os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scenario-wise-rec_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
import workloads, model_factory, gpu_util
case = sys.argv[1] if len(sys.argv) > 1 else "cfg2_mmoe_aliccp_b4096"
model_name, cfg, B = workloads.CASES[case]
x, y = gpu_util.make_batch(workloads.all_feature_specs(cfg), B, cfg["domain_num"], seed=5)
torch.manual_seed(7)
model = model_factory.build(model_name, cfg).to("cuda:0").train()
xg = {k: v.to("cuda:0") for k, v in x.items()}
for i in range(3):
    sys.stderr.write(f"--- pass {i}\n")
    out = model(xg)
    torch.cuda.synchronize()
    sys.stderr.write(f"--- backward {i}\n")
    out.sum().backward()
    torch.cuda.synchronize()
