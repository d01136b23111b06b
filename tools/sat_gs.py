"""Saturated K1 / K2 alone (for an ncu --set full capture of exactly these launches): 2^18 rows x the cfg2 fields on 1.6 GB of tables."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scenario-wise-rec_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
import bench, workloads
model_name, cfg, B = workloads.CASES[workloads.DEFAULT_CASE]
print(bench.saturated_gather_scatter(workloads.all_feature_specs(cfg), torch.device("cuda", 0), bench.peaks()))
