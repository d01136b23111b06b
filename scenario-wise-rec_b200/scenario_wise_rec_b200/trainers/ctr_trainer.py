"""CTRTrainer with the reference's constructor and methods
(reference: scenario_wise_rec/trainers/ctr_trainer.py:10-165).

``train_one_epoch`` keeps the reference loop -- batch to device, ``model(x_dict)``,
``BCELoss``, ``zero_grad``, ``backward``, ``optimizer.step`` (ctr_trainer.py:62-77) -- and
the evaluation methods return the same values.  The per-batch work is delegated to
:meth:`train_step`, which callers (bench.py) may also drive directly with one
``(x_dict, y)`` batch of host or device tensors.

When the model is one of this package's device-program models on a CUDA device and the
optimizer is plain ``torch.optim.Adam`` (the trainer default), ``train_step`` runs the whole
step -- forward, BCELoss, backward, Adam -- as ONE CUDA graph fed by one packed H2D copy
(:mod:`.fused_step`); same arithmetic, same ``state_dict`` / ``optimizer.state`` contents.
Any other combination takes the reference loop verbatim.  ``fused=False`` forces the latter.
"""
from __future__ import annotations

import os
import time

import torch

from ..basic.callback import EarlyStopper


def _progress(it, desc):
    try:
        import tqdm
        return tqdm.tqdm(it, desc=desc, smoothing=0, mininterval=1.0)
    except ImportError:          # tqdm is a reference dependency; degrade to the bare iterator
        return it


class CTRTrainer(object):
    def __init__(self, model, data_set_type, optimizer_fn=torch.optim.Adam, optimizer_params=None, scheduler_fn=None,
                 scheduler_params=None, n_epoch=10, earlystop_patience=10, device="cpu", gpus=None, model_path="./",
                 fused=True):
        self.model = model
        self.fused = fused
        self._steps = {}
        self._flat = None
        self._grad_sync = None
        self.data_set_type = data_set_type
        gpus = [] if gpus is None else gpus
        self.gpus = gpus
        if len(gpus) > 1:
            raise NotImplementedError("single-process nn.DataParallel is replaced by one process per GPU "
                                      "(torch.distributed / NCCL); launch with torchrun, see INTEGRATION.md")
        self.device = torch.device(device)
        self.model.to(self.device)
        if optimizer_params is None:
            optimizer_params = {"lr": 1e-3, "weight_decay": 1e-5}
        self.optimizer = optimizer_fn(self.model.parameters(), **optimizer_params)
        self.scheduler = None
        if scheduler_fn is not None:
            self.scheduler = scheduler_fn(self.optimizer, **scheduler_params)
        self.criterion = torch.nn.BCELoss()
        self.n_epoch = n_epoch
        self.early_stopper = EarlyStopper(patience=earlystop_patience)
        self.model_path = model_path

    def enable_data_parallel(self, group=None):
        """One process per GPU (torchrun): average the gradient arenas of every backward over the
        process group with one NCCL all-reduce per arena (replaces nn.DataParallel, ctr_trainer.py:45-47
        of the reference; like it, BatchNorm statistics stay per replica)."""
        import torch.distributed as dist

        def sync(arenas):
            for name, a in arenas.items():
                # "virt" / "shard": gradients of row-sharded tables are routed by the embedding exchange instead
                if name in ("dense", "emb") and a.numel() > 1:
                    dist.all_reduce(a, op=dist.ReduceOp.AVG, group=group)

        self.model._grad_sync = sync
        self.model._programs = {}
        self._grad_sync = sync
        self._steps = {}

    @staticmethod
    def evaluate_fn(targets, predicts):
        from sklearn.metrics import roc_auc_score
        return roc_auc_score(targets, predicts)

    # ---- one batch -----------------------------------------------------------------------------
    def _fused_step_for(self, x_dict):
        from ..fused import FusedModule
        from .fused_step import FlatArenas, FusedTrainStep, PackedBatch, adam_supported
        if not (self.fused and isinstance(self.model, FusedModule) and self.device.type == "cuda" and self.model.training
                and adam_supported(self.optimizer) and type(self.criterion) is torch.nn.BCELoss
                and self.criterion.reduction == "mean" and self.criterion.weight is None):
            return None
        if self._flat is not None and not self._flat.valid():
            self._flat, self._steps = None, {}           # parameter storage was replaced (model.to / .float ...)
        if isinstance(x_dict, PackedBatch):
            key = x_dict.key
        else:
            cols = self.model._columns()
            key = (int(x_dict[cols[0]].shape[0]), tuple(x_dict[c].dtype for c in cols))
        fs = self._steps.get(key)
        if fs is None:
            if isinstance(x_dict, PackedBatch):
                raise ValueError("no fused step was built for this packed batch; call trainer.packer(x_dict) first")
            fs = FusedTrainStep(self.model, self.optimizer, x_dict, self.device, flat=self._flat, grad_sync=self._grad_sync)
            if self._flat is None:
                self._flat = fs.flat
                self.optimizer.register_state_dict_pre_hook(lambda _opt: self._flat is not None and self._flat.publish_step())
                self.optimizer.register_load_state_dict_post_hook(lambda _opt: self._flat is not None and self._flat.reload_optimizer_state())
                if fs.flat.lazy is not None and not getattr(self.model, "_swr_lazy_hooks", False):
                    # row-lazy Adam: whatever reads the tables outside the fused step first gets every row replayed to
                    # the current step (generic forward = evaluation / prediction, state_dict = checkpoints, EarlyStopper)
                    def flush(*_a, **_k):
                        if self._flat is not None:
                            self._flat.flush_lazy()
                    self.model.register_forward_pre_hook(flush)
                    self.model.register_state_dict_pre_hook(flush)
                    self.model._swr_lazy_hooks = True
            self._steps[key] = fs
        return fs

    def packer(self, x_dict):
        """The fused step object for batches shaped like ``x_dict``: ``packer(x).pack(x, y, device)`` lays a batch out
        in the staging format once, so ``train_step(packed)`` is a single copy + one graph launch."""
        fs = self._fused_step_for(x_dict)
        if fs is None:
            raise RuntimeError("the fused training step is not available for this model / optimizer / device")
        return fs

    def train_step(self, x_dict, y=None):
        """ctr_trainer.py:67-73 for one batch; returns the loss (``.item()`` reads it; device work is not synchronised)."""
        fs = self._fused_step_for(x_dict)
        if fs is not None:
            return fs.step(x_dict, y)
        x_dict = {k: v.to(self.device, non_blocking=True) for k, v in x_dict.items()}
        y = y.to(self.device, non_blocking=True)
        y_pred = self.model(x_dict)
        loss = self.criterion(y_pred, y.float())
        self.model.zero_grad()
        loss.backward()
        self.optimizer.step()
        return loss

    def train_one_epoch(self, data_loader, log_interval=10):
        self.model.train()
        total_loss = 0
        tk0 = _progress(data_loader, "train")
        pending = []
        for i, (x_dict, y) in enumerate(tk0):
            pending.append(self.train_step(x_dict, y))
            if (i + 1) % log_interval == 0:
                # the reference adds loss.item() every step (a host sync per batch); the same running mean is
                # formed here once per log interval so the device never waits for the host in between
                total_loss = sum(l.item() for l in pending)
                pending = []
                if hasattr(tk0, "set_postfix"):
                    tk0.set_postfix(loss=total_loss / log_interval)
                total_loss = 0
        for l in pending:
            l.item()
        if self._flat is not None:
            self._flat.publish_step()
        if hasattr(self.model, "check_indices"):
            self.model.check_indices()

    def fit(self, train_dataloader, val_dataloader=None):
        for epoch_i in range(self.n_epoch):
            print("epoch:", epoch_i)
            self.train_one_epoch(train_dataloader)
            if self.scheduler is not None:
                if epoch_i % self.scheduler.step_size == 0:
                    print("Current lr : {}".format(self.optimizer.state_dict()["param_groups"][0]["lr"]))
                self.scheduler.step()
            if val_dataloader:
                auc, logloss = self.evaluate(self.model, val_dataloader)
                auc = self._agree(auc)         # one process per GPU: every rank must take the same decision
                print(f"epoch:{epoch_i} | val auc: {auc} | val logloss: {logloss}")
                if self.early_stopper.stop_training(auc, self.model.state_dict()):
                    print(f"validation: best auc: {self.early_stopper.best_auc}")
                    self.model.load_state_dict(self.early_stopper.best_weights)
                    break
        time_now = time.strftime("%m_%d_%H_%M", time.localtime(int(round(time.time() * 1000)) / 1000))
        name = self.model.__class__.__name__ + "_" + self.data_set_type + "_" + time_now + ".pth"
        # same file as the reference writes (ctr_trainer.py:94-97).  With row-sharded tables the shards are gathered
        # back into full [vocab, E] tables under the reference's keys first, and one rank writes the file.
        from .. import parallel
        import torch.distributed as dist
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        state = parallel.full_state_dict(self.model) if parallel._sharded_params(self.model) else self.model.state_dict()
        if not multi or dist.get_rank() == 0:      # replicas are identical: one writer
            torch.save(state, os.path.join(self.model_path, name))
        if multi:
            dist.barrier()

    @staticmethod
    def _agree(value):
        """Under one process per GPU every rank validates on its own shard of the data; early stopping (and with it the
        number of collectives each rank still issues) must not depend on the rank: the decision uses the mean over ranks."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return value
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0]) / dist.get_world_size()

    # ---- evaluation ------------------------------------------------------------------------------
    def _predict_batches(self, model, data_loader, desc):
        model.eval()
        with torch.no_grad():
            for x_dict, y in _progress(data_loader, desc):
                x_dict = {k: v.to(self.device, non_blocking=True) for k, v in x_dict.items()}
                yield x_dict, y, model(x_dict)
        if hasattr(model, "check_indices"):
            model.check_indices()

    # The reference calls ``.tolist()`` on the labels and the predictions of every batch (ctr_trainer.py:109-110,
    # 130-131, 164): one host sync per batch.  Here the predictions stay on the device until the loader is exhausted
    # and come back in ONE copy; the metric functions then receive the same Python lists as in the reference.
    @staticmethod
    def _to_list(chunks):
        if not chunks:
            return []
        return torch.cat([c.reshape(-1) for c in chunks]).cpu().tolist()

    def _device_metrics(self) -> bool:
        """On a CUDA device, with the default metric (sklearn's roc_auc_score, like the reference), AUC and logloss are
        evaluated on the device (trainers/metrics.py) and only the scalars come back; a user-supplied ``evaluate_fn``
        keeps receiving Python lists."""
        return self.device.type == "cuda" and self.evaluate_fn is CTRTrainer.evaluate_fn

    def evaluate(self, model, data_loader, mode="val"):
        from sklearn.metrics import log_loss
        targets, predicts = [], []
        for _x, y, y_pred in self._predict_batches(model, data_loader, "validation"):
            targets.append(y)
            predicts.append(y_pred)
        if targets and self._device_metrics():
            from .metrics import binary_auc, binary_logloss
            y = torch.cat([c.reshape(-1) for c in targets]).to(self.device, non_blocking=True)
            p = torch.cat([c.reshape(-1) for c in predicts])
            return binary_auc(y, p), binary_logloss(y, p)
        targets, predicts = self._to_list(targets), self._to_list(predicts)
        return self.evaluate_fn(targets, predicts), log_loss(targets, predicts)

    def evaluate_multi_domain_loss(self, model, data_loader, domain_num):
        from sklearn.metrics import log_loss
        t_chunks, p_chunks, d_chunks = [], [], []
        for x_dict, y, y_pred in self._predict_batches(model, data_loader, "validation"):
            t_chunks.append(y)
            p_chunks.append(y_pred)
            d_chunks.append(x_dict["domain_indicator"])
        if not p_chunks:
            return [None] * domain_num, [None] * domain_num, None, None
        if self._device_metrics():
            # per-domain masks, AUC and logloss on the device (ctr_trainer.py:113-152 of the reference builds D pairs of
            # Python lists); a domain without samples reports None like the reference
            from .metrics import binary_auc, binary_logloss
            y = torch.cat([c.reshape(-1) for c in t_chunks]).to(self.device, non_blocking=True)
            p = torch.cat([c.reshape(-1) for c in p_chunks])
            dom = torch.cat([c.reshape(-1) for c in d_chunks]).to(self.device)
            logloss_d, auc_d = [], []
            for d in range(domain_num):
                m = dom == d
                if bool(m.any()):
                    logloss_d.append(binary_logloss(y[m], p[m]))
                    auc_d.append(binary_auc(y[m], p[m]))
                else:
                    logloss_d.append(None)
                    auc_d.append(None)
            return logloss_d, auc_d, binary_logloss(y, p), binary_auc(y, p)
        y = torch.cat([c.reshape(-1) for c in t_chunks]).cpu()
        y_pred = torch.cat([c.reshape(-1) for c in p_chunks]).cpu()
        dom = torch.cat([c.reshape(-1) for c in d_chunks]).cpu()
        t_all, p_all = y.tolist(), y_pred.tolist()
        t_dom, p_dom = [], []
        for d in range(domain_num):
            m = dom == d
            t_dom.append(y[m].tolist())
            p_dom.append(y_pred[m].tolist())
        logloss_d = [log_loss(t_dom[d], p_dom[d]) if t_dom[d] else None for d in range(domain_num)]
        auc_d = [self.evaluate_fn(t_dom[d], p_dom[d]) if t_dom[d] else None for d in range(domain_num)]
        total_logloss = log_loss(t_all, p_all) if p_all else None
        total_auc = self.evaluate_fn(t_all, p_all) if p_all else None
        return logloss_d, auc_d, total_logloss, total_auc

    def predict(self, model, data_loader):
        predicts = []
        for _x, _y, y_pred in self._predict_batches(model, data_loader, "predict"):
            predicts.append(y_pred)
        return self._to_list(predicts)
