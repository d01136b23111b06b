"""Shared plumbing of the multi-domain models."""
from __future__ import annotations

import torch

from ...fused import FusedModule


class MultiDomainModel(FusedModule):
    """forward(x: dict[str, Tensor[B]]) -> Tensor[B] in (0, 1); ``x`` must hold every feature
    column plus ``"domain_indicator"``; extra keys are ignored (as in the reference)."""

    def _feature_lists(self):
        return [self.features]

    def _columns(self):
        cols, seen = [], set()
        for feats in self._feature_lists():
            for f in feats:
                if f.name not in seen:
                    seen.add(f.name)
                    cols.append(f.name)
        return cols + ["domain_indicator"]

    def forward(self, x):
        return self._run(x)

    @staticmethod
    def _dom_dtype(col_dtypes) -> torch.dtype:
        return col_dtypes["domain_indicator"]
