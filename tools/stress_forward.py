import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scenario-wise-rec_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
import workloads, model_factory, gpu_util
from scenario_wise_rec_b200 import _native as N
N.set_fc_mode(N.FC_TC)
case = "cfg5a_hamursmall_mind_b16384"
model_name, cfg, B = workloads.CASES[case]
x, y = gpu_util.make_batch(workloads.all_feature_specs(cfg), B, cfg["domain_num"], seed=5, zipf=True)
torch.manual_seed(7)
model = model_factory.build(model_name, cfg)
gpu_util.randomise(model, 11)
model.to("cuda:0")
xg = {k: v.to("cuda:0") for k, v in x.items()}
for mode in ("eval", "train"):
    model.train(mode == "train")
    outs = []
    with torch.no_grad():
        for i in range(int(os.environ.get("STRESS_RUNS", "8"))):
            outs.append(model(xg).clone())
            torch.cuda.synchronize()
    ref = torch.stack(outs).median(0).values
    bad = [(i, float((o - ref).abs().max())) for i, o in enumerate(outs) if float((o - ref).abs().max()) > 1e-5]
    print(mode, "runs off the median:", bad)
