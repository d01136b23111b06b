"""Glue between ``torch.nn.Module`` / autograd and the device program.

``FusedModule`` is the base of every drop-in module: subclasses keep their parameters
as ordinary ``nn.Parameter`` / buffers under the reference's names and implement
``_lower(builder, x)`` which describes the forward with a :class:`ProgramBuilder`.
``forward`` looks up (or builds) the program for the current (batch size, train/eval,
column dtypes), runs it on the module's CUDA device and hooks the backward program into
autograd through ``_ProgramFn``.
"""
from __future__ import annotations

from typing import Dict

import torch
from torch import nn
from torch.autograd.function import once_differentiable

from .program import CudaRunner, Program, ProgramBuilder


class _ProgramFn(torch.autograd.Function):
    """forward = forward records, backward = backward records of one Program."""

    @staticmethod
    def forward(ctx, runner, x, *params):
        outs = runner.forward(x)
        ctx.runner = runner
        ctx.generation = runner.generation
        return outs

    @staticmethod
    @once_differentiable
    def backward(ctx, *gout):
        runner = ctx.runner
        if ctx.generation != runner.generation:
            raise RuntimeError("the activation workspace of this forward was overwritten by a later forward "
                               "of the same module; call backward() before running the module again")
        grads = runner.backward(gout)
        return (None, None, *grads)


class FusedModule(nn.Module):
    """nn.Module whose forward is one device program (see module docstring)."""

    #: runner class; tests swap in oracle.ops_ref.RefRunner to check the lowering on CPU
    _runner_factory = None
    #: optional callable(dict arena name -> flat gradient tensor) run after every backward (data parallel)
    _grad_sync = None

    def __init__(self):
        super().__init__()
        self._programs: Dict[tuple, object] = {}

    # parameters moved / cast (``.to``, ``.cuda``) -> every cached pointer is stale
    def _apply(self, fn, *args, **kwargs):
        self._programs = {}
        return super()._apply(fn, *args, **kwargs)

    def _device(self) -> torch.device:
        for p in self.parameters():
            return p.device
        raise RuntimeError("module has no parameters")

    def _columns(self):
        """Names of the feature columns the program reads (besides ``domain_indicator``)."""
        raise NotImplementedError

    def _lower(self, b: ProgramBuilder, col_dtypes) -> None:
        raise NotImplementedError

    def _runner(self, x):
        cols = self._columns()
        first = x[cols[0]]                      # KeyError on a missing feature, like the reference
        B = int(first.shape[0])
        dts = {c: x[c].dtype for c in cols}
        key = (B, self.training, tuple(dts.values()))
        r = self._programs.get(key)
        if r is None:
            b = ProgramBuilder(B, self.training)
            b.exchange = self._get_exchange()
            self._lower(b, dts)
            prog = b.finish()
            factory = type(self)._runner_factory
            r = factory(prog) if factory is not None else CudaRunner(prog, self._device())
            r.grad_sync = self._grad_sync
            self._programs[key] = r
        return r

    def _get_exchange(self):
        """The ShardedExchange of this model (created on first use; None when no table is row-sharded)."""
        ex = self.__dict__.get("_exchange")
        if ex is None and any(getattr(f, "shard", None) is not None for fl in self._all_feature_lists() for f in fl):
            from .parallel import ShardedExchange
            ex = self.__dict__["_exchange"] = ShardedExchange()
        return ex

    def _all_feature_lists(self):
        if hasattr(self, "_feature_lists"):
            return self._feature_lists()
        return [getattr(self, "_active", None) or getattr(self, "features", [])]

    def _run(self, x):
        r = self._runner(x)
        prog: Program = r.prog
        want_grad = torch.is_grad_enabled() and any(p.requires_grad for p in prog.params)
        params = prog.params
        if prog.virtual_fields:
            # row-sharded tables: exchange looked-up rows over NVLink, then read them through virtual tables
            from .parallel import ShardedLookup
            ex, x, subst = self._get_exchange(), dict(x), {}
            for f in prog.virtual_fields:
                if want_grad and f.shard.requires_grad:
                    subst[id(f.virt)] = ShardedLookup.apply(f.shard, x[f.name], ex, f)
                else:
                    ex.lookup(f, x[f.name])
                x[f.vcol] = f.vidx
            params = [subst.get(id(p), p) for p in prog.params]
        if want_grad:
            outs = _ProgramFn.apply(r, x, *params)
        else:
            outs = r.forward(x)
        return outs[0] if len(outs) == 1 else outs

    def check_indices(self):
        """Synchronise and raise ``IndexError`` if any embedding index was out of range (the
        reference raises on CPU / device-asserts on CUDA; here the flag is also checked at the
        start of every following forward)."""
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        for r in self._programs.values():
            r.check_indices()
        if self.__dict__.get("_exchange") is not None:
            self.__dict__["_exchange"].check_indices()
