"""The reference arm of bench.py: the UNMODIFIED reference package (installed from /root/reference into the
git-ignored ``baseline/_ref`` with ``pip install --target``) driven through its own public
API -- ``scenario_wise_rec.models.multi_domain.<Model>`` + ``scenario_wise_rec.trainers.CTRTrainer.train_one_epoch``
(trainers/ctr_trainer.py:62-77) -- on the host cores (``device="cpu"``) or, unchanged, on ``cuda:0`` (stock eager
PyTorch: the "same box" bar of SURVEY.md 8d).  Nothing of this repo's models, kernels or engine is on that path;
only the synthetic workload definitions (tools/workloads.py) are shared so both arms see the same tensors.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def available():
    return os.path.isdir(os.path.join(REF_DIR, "scenario_wise_rec"))


def _import_reference():
    if not available():
        raise RuntimeError("baseline/_ref is missing: python -m pip install --no-index --no-build-isolation --no-deps "
                           "--target baseline/_ref /root/reference (from a writable copy)")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import scenario_wise_rec  # noqa: F401
    return scenario_wise_rec


def _features(spec):
    from scenario_wise_rec.basic.features import DenseFeature, SparseFeature
    return [SparseFeature(n, vocab_size=v, embed_dim=d) if k == "sparse" else DenseFeature(n) for n, k, v, d in spec]


def build_model(model_name, cfg, device="cpu"):
    """The reference model for a workload config (constructor kwargs as in the reference's scripts/)."""
    _import_reference()
    import scenario_wise_rec.models.multi_domain as M
    f = lambda key: _features(cfg[key])   # noqa: E731   fresh Feature objects: they cache their nn.Embedding
    D = cfg.get("domain_num")
    if model_name == "SharedBottom":
        return M.SharedBottom(f("features"), D, bottom_params={"dims": list(cfg["bottom_dims"])}, tower_params={"dims": list(cfg["tower_dims"])})
    if model_name == "MMOE":
        return M.MMOE(f("features"), D, n_expert=cfg["n_expert"], expert_params={"dims": list(cfg["expert_dims"])},
                      tower_params={"dims": list(cfg["tower_dims"])})
    if model_name == "PLE":
        return M.PLE(f("features"), D, n_level=cfg["n_level"], n_expert_specific=cfg["n_expert_specific"],
                     n_expert_shared=cfg["n_expert_shared"], expert_params={"dims": list(cfg["expert_dims"])},
                     tower_params={"dims": list(cfg["tower_dims"])})
    if model_name == "Star":
        return M.Star(f("features"), D, fcn_dims=list(cfg["fcn_dims"]), aux_dims=list(cfg["aux_dims"]))
    if model_name == "PPNet":
        return M.PPNet(id_features=f("id_features"), agn_features=f("agn_features"), domain_num=D, fcn_dims=list(cfg["fcn_dims"]))
    if model_name == "EPNet":
        return M.EPNet(sce_features=f("sce_features"), agn_features=f("agn_features"), fcn_dims=list(cfg["fcn_dims"]))
    if model_name == "M3oE":
        return M.M3oE(f("features"), D, fcn_dims=list(cfg["fcn_dims"]), expert_num=cfg["expert_num"], exp_d=1, exp_t=1, bal_d=1, bal_t=1,
                      device=device)
    if model_name == "HamurSmall":
        return M.HamurSmall(f("features"), D, fcn_dims=list(cfg["fcn_dims"]), hyper_dims=list(cfg["hyper_dims"]), k=cfg["k"])
    if model_name == "HamurLarge":
        return M.HamurLarge(f("features"), D, fcn_dims=list(cfg["fcn_dims"]), hyper_dims=list(cfg["hyper_dims"]), k=cfg["k"])
    raise KeyError(model_name)


class ReferenceArm:
    """Reference model + reference CTRTrainer on ``device``; ``time()`` runs the stock ``train_one_epoch``."""

    def __init__(self, model_name, cfg, device="cpu", seed=0):
        _import_reference()
        from scenario_wise_rec.trainers import CTRTrainer
        torch.manual_seed(seed)
        self.device = torch.device(device)
        self.model = build_model(model_name, cfg, device=str(device))
        with contextlib.redirect_stdout(io.StringIO()):
            self.trainer = CTRTrainer(self.model, "synthetic", optimizer_params={"lr": 1e-3, "weight_decay": 1e-5}, device=str(device))
        self.cores = torch.get_num_threads()

    def epoch(self, batches):
        """One pass of the reference's own loop over a list of (x_dict, y) host batches (its tqdm bar goes to stderr)."""
        self.trainer.train_one_epoch(batches)

    def time(self, batches, steps, warmup, budget_s=None):
        """-> (steps done, seconds).  CUDA: events on the current stream around the whole epoch; CPU: perf_counter."""
        cuda = self.device.type == "cuda"
        self.epoch([batches[i % len(batches)] for i in range(warmup)])
        if budget_s is not None:            # bounded sample: calibrate on one step
            t0 = time.perf_counter()
            self.epoch([batches[0]])
            if cuda:
                torch.cuda.synchronize()
            one = time.perf_counter() - t0
            steps = max(1, min(steps, int(budget_s / max(one, 1e-6))))
        seq = [batches[i % len(batches)] for i in range(steps)]
        if cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record()
            self.epoch(seq)
            e1.record()
            torch.cuda.synchronize()
            return steps, max(e0.elapsed_time(e1) * 1e-3, time.perf_counter() - t0)
        t0 = time.perf_counter()
        self.epoch(seq)
        return steps, time.perf_counter() - t0
