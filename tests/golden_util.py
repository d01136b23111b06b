"""Helpers to load the golden fixtures written by oracle/make_golden.py."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def _tuples(cfg):
    for key in ("features", "id_features", "agn_features", "sce_features"):
        if key in cfg:
            cfg[key] = [tuple(f) for f in cfg[key]]
    return cfg


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
        self.name = name
        self.model = str(z["model"])
        self.cfg = _tuples(json.loads(str(z["cfg"])))
        self.B = int(z["B"])
        self.x = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("x/")}
        self.y = torch.from_numpy(z["y"])
        self.state0 = {k[7:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("state0/")}
        self.state1 = {k[7:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("state1/")}
        self.grads = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad/")}
        self.grad_none = set(json.loads(str(z["grad_none"])))
        self.out_train = torch.from_numpy(z["out_train"])
        self.out_eval = torch.from_numpy(z["out_eval"])
        self.loss = float(z["loss"])

    def param_names(self):
        return [k for k in self.state0 if "running_" not in k and "num_batches" not in k]

    def leaf_state(self):
        """state0 with float params as autograd leaves."""
        st = {}
        for k, v in self.state0.items():
            v = v.clone()
            if v.dtype.is_floating_point and "running_" not in k:
                v.requires_grad_(True)
            st[k] = v
        return st
