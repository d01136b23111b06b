"""Development aid: per-phase cycle counts of the tcgen05 forward kernel (SWR_TC_DEBUG=1) on one BASELINE case.
usage: SWR_TC_DEBUG=1 python tools/tc_phase_timing.py <case>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scenario-wise-rec_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
import workloads, model_factory, gpu_util
case = sys.argv[1]
model_name, cfg, B = workloads.CASES[case]
x, y = gpu_util.make_batch(workloads.all_feature_specs(cfg), B, cfg["domain_num"], seed=5, zipf=True)
torch.manual_seed(7)
model = model_factory.build(model_name, cfg)
gpu_util.randomise(model, 11)
model.to("cuda:0").train()
xg = {k: v.to("cuda:0") for k, v in x.items()}
for i in range(3):
    print(f"--- forward {i}", file=sys.stderr)
    out = model(xg)
    torch.cuda.synchronize()
