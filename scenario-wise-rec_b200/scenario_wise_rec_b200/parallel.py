"""Row-sharded embedding tables across one process per GPU (no reference counterpart:
the reference keeps every table whole on one device, basic/features.py:76-79).

Layout: row ``r`` of a sharded table lives on rank ``r % R`` at local row ``r // R``; every rank keeps
``ceil(V / R)`` rows.  Small tables stay replicated (their gradients are all-reduced with the dense weights).

Per step and sharded field, for a local batch of B rows (shapes are static, nothing syncs with the host, so the
whole exchange is CUDA-graph capturable):

  forward   all_gather(idx)                       [R, B] int64   every rank sees every request
            local = idx // R where idx % R == me, else -1
            rows  = K1 gather(shard, local)       [R, B, E]      zero rows for requests owned elsewhere
            recv  = all_to_all(rows)              [R, B, E]      block o = rows served by owner o      (NVLink)
            the device program reads field f from the *virtual table* recv.view(R*B, E) at
            virtual index (idx % R) * B + b                                       (fused K1 gather as usual)
  backward  the program's K2 scatter writes the dense gradient of the virtual table = the per-owner send blocks
            grecv = all_to_all(gvirt)             [R, B, E]      block q = gradients from requester q  (NVLink)
            K2 scatter(grecv / R, local) into the shard gradient           (mean over the global batch)

Bytes per rank and direction: R*B*E*4 (8 x 4096 x 16 x 4 = 2 MB at cfg4) -- latency-bound, not bandwidth-bound.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _native as N


class ShardInfo:
    def __init__(self, rank: int, world: int, group=None):
        self.rank, self.world, self.group = rank, world, group

    def local_rows(self, vocab: int) -> int:
        return (vocab + self.world - 1) // self.world


def shard_features(features, min_rows: int = 1_000_000, group=None) -> List[str]:
    """Mark every SparseFeature with ``vocab_size >= min_rows`` as row-sharded over ``group``.  Call before the
    model is built (the feature object creates its shard-sized ``nn.Embedding``).  Returns the sharded names."""
    from .basic.features import SparseFeature
    info = ShardInfo(dist.get_rank(group), dist.get_world_size(group), group)
    names = []
    for f in features:
        if isinstance(f, SparseFeature) and f.vocab_size >= min_rows and f.shared_with is None:
            if hasattr(f, "embed"):
                raise RuntimeError(f"feature {f.name!r} already owns a full table; shard before building the model")
            f.shard = info
            names.append(f.name)
    by_name = {f.name: f for f in features if isinstance(f, SparseFeature)}
    for f in features:          # a feature that shares a sharded table reads the same shards
        if isinstance(f, SparseFeature) and f.shared_with is not None:
            owner = by_name.get(f.shared_with)
            if owner is not None and getattr(owner, "shard", None) is not None:
                f.shard = owner.shard
    return names


def shard_of(full: torch.Tensor, info: ShardInfo) -> torch.Tensor:
    """Rows of ``full`` [V, E] owned by this rank, padded with zeros to ceil(V / R) rows."""
    part = full[info.rank::info.world]
    out = full.new_zeros(info.local_rows(full.shape[0]), full.shape[1])
    out[:part.shape[0]] = part
    return out


# ---- checkpoints in the reference's format (ctr_trainer.py:90-97, basic/callback.py:27) ----------
def _sharded_params(model) -> Dict[str, "ShardInfo"]:
    """state_dict key of every row-sharded table of ``model`` -> (ShardInfo, vocab_size)."""
    from .basic.features import SparseFeature
    out = {}
    for feats in model._feature_lists():
        for f in feats:
            if isinstance(f, SparseFeature) and getattr(f, "shard", None) is not None:
                for key, prm in model.named_parameters():
                    if prm is f.get_embedding_layer().weight:
                        out[key] = (f.shard, f.vocab_size)
    return out


def full_state_dict(model) -> Dict[str, torch.Tensor]:
    """``model.state_dict()`` with every row-sharded table re-assembled to its full ``[vocab, E]`` shape under the
    reference's key (collective: every rank of the shard group must call it; every rank gets the full dict, so a
    ``.pth`` written by rank 0 loads into the reference model with ``strict=True``)."""
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    for key, (info, vocab) in _sharded_params(model).items():
        local = sd[key].contiguous()
        parts = [torch.empty_like(local) for _ in range(info.world)]
        dist.all_gather(parts, local, group=info.group)
        full = local.new_empty(vocab, local.shape[1])
        for r, part in enumerate(parts):
            n = (vocab - r + info.world - 1) // info.world      # rows r, r + R, r + 2R, ...
            full[r::info.world] = part[:n]
        sd[key] = full
    return sd


def load_full_state_dict(model, state: Dict[str, torch.Tensor], strict: bool = True):
    """Load a reference-format state dict (full tables) into a model whose large tables are row-sharded: every rank
    keeps rows ``rank, rank + R, ...`` of each sharded table (no communication)."""
    state = dict(state)
    for key, (info, vocab) in _sharded_params(model).items():
        if key in state and state[key].shape[0] == vocab:
            state[key] = shard_of(state[key], info)
    return model.load_state_dict(state, strict=strict)


# ---- peer memory: shard arenas every rank of the node can address (NVLink / NVSwitch) -------------------------------
class PeerArena:
    """A float32 device buffer of ``numel`` elements allocated outside torch's caching allocator (swr_peer_alloc) whose
    CUDA IPC handle is exchanged over ``group``: ``ptrs[r]`` is the address of rank r's buffer in THIS process (own
    rank: the local pointer).  ``tensor`` views the local buffer.  Same size on every rank."""

    def __init__(self, numel: int, device: torch.device, group=None):
        L = N.lib()
        self.numel, self.device, self.group = int(numel), device, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        ptr = ctypes.c_void_p()
        with torch.cuda.device(device):
            N.check(L.swr_peer_alloc(4 * self.numel, ctypes.byref(ptr)), "swr_peer_alloc")
            handle = (ctypes.c_ubyte * 64)()
            N.check(L.swr_peer_handle(ptr, handle), "swr_peer_handle")
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
            allh = torch.empty(self.world, 64, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(allh.view(-1), mine, group=group)
            allh = allh.cpu()
            self.local_ptr = int(ptr.value)
            self.ptrs = []
            for r in range(self.world):
                if r == self.rank:
                    self.ptrs.append(self.local_ptr)
                    continue
                buf = (ctypes.c_ubyte * 64)(*allh[r].tolist())
                q = ctypes.c_void_p()
                N.check(L.swr_peer_open(buf, ctypes.byref(q)), "swr_peer_open")
                self.ptrs.append(int(q.value))
        self.tensor = torch.as_tensor(_CudaArray(self.local_ptr, self.numel, self), device=device)
        dist.barrier(group=group)

    def peer_table(self, offset: int) -> torch.Tensor:
        """Device int64 [world]: address of element ``offset`` of every rank's buffer (kernel argument of K1 / K2)."""
        return torch.tensor([q + 4 * int(offset) for q in self.ptrs], dtype=torch.int64, device=self.device)

    def __del__(self):
        try:
            L = N.lib()
            for r, q in enumerate(self.ptrs):
                if r != self.rank:
                    L.swr_peer_close(ctypes.c_void_p(q))
            L.swr_peer_free(ctypes.c_void_p(self.local_ptr))
        except Exception:
            pass


class _CudaArray:
    """__cuda_array_interface__ view of a raw device allocation (keeps its owner alive)."""

    def __init__(self, ptr: int, numel: int, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}


# ---- local row gather / scatter through the C ABI (K1 / K2 on one field) ---------------------------------
def local_gather(table: torch.Tensor, idx: torch.Tensor, out: torch.Tensor):
    """out[i] = table[idx[i]] (zero row when idx[i] is outside [0, rows)); no out-of-range flag."""
    n, E = idx.numel(), table.shape[1]
    vp = ctypes.c_void_p
    tabs, idxs = (vp * 1)(table.data_ptr()), (vp * 1)(idx.data_ptr())
    vocab, idt = (ctypes.c_int64 * 1)(table.shape[0]), (ctypes.c_int32 * 1)(N.torch_dtype_code(idx.dtype))
    st = N.lib().swr_embedding_gather_fwd(tabs, vocab, idxs, idt, None, (ctypes.c_int32 * 1)(0), out.data_ptr(), E, n, 1, E, 0,
                                          None, torch.cuda.current_stream(table.device).cuda_stream)
    N.check(st, "swr_embedding_gather_fwd")


def local_scatter(grad_rows: torch.Tensor, idx: torch.Tensor, gtable: torch.Tensor):
    """gtable[idx[i]] += grad_rows[i] (entries with idx outside [0, rows) are skipped)."""
    n, E = idx.numel(), gtable.shape[1]
    vp = ctypes.c_void_p
    gts, idxs = (vp * 1)(gtable.data_ptr()), (vp * 1)(idx.data_ptr())
    vocab, idt = (ctypes.c_int64 * 1)(gtable.shape[0]), (ctypes.c_int32 * 1)(N.torch_dtype_code(idx.dtype))
    st = N.lib().swr_embedding_scatter_bwd(grad_rows.data_ptr(), E, n, idxs, idt, gts, vocab, 1, E,
                                           torch.cuda.current_stream(gtable.device).cuda_stream)
    N.check(st, "swr_embedding_scatter_bwd")


class _Field:
    def __init__(self, fea, shard_param: torch.nn.Parameter, B: int):
        info: ShardInfo = fea.shard
        R, E = info.world, int(shard_param.shape[1])
        dev = shard_param.device
        self.name, self.vocab, self.info, self.shard, self.B, self.E = fea.name, fea.vocab_size, info, shard_param, B, E
        self.vcol = "__vidx__" + fea.name
        # the virtual table the device program reads (static address; rewritten by every exchange)
        self.virt = torch.zeros(R * B, E, device=dev, requires_grad=True)
        self.rows = torch.zeros(R * B, E, device=dev)
        self.local = torch.full((R * B,), -1, dtype=torch.int64, device=dev)
        self.idx_all = torch.zeros(R * B, dtype=torch.int64, device=dev)
        self.vidx = torch.zeros(B, dtype=torch.int64, device=dev)
        self.grecv = torch.zeros(R * B, E, device=dev)
        self.arange = torch.arange(B, dtype=torch.int64, device=dev)
        self.bad = torch.zeros((), dtype=torch.bool, device=dev)


class ShardedExchange:
    """Owns the per-(field, batch size) buffers and runs the collectives around one model's device program."""

    def __init__(self):
        self.fields: Dict[Tuple[str, int], _Field] = {}

    def field(self, fea, shard_param, B: int) -> _Field:
        key = (fea.name, int(B))
        f = self.fields.get(key)
        if f is None or f.shard is not shard_param or f.shard.device != f.virt.device:
            f = self.fields[key] = _Field(fea, shard_param, int(B))
        return f

    def fields_for(self, B: int) -> List[_Field]:
        return [f for (n, b), f in self.fields.items() if b == int(B)]

    # ---- forward half -----------------------------------------------------------------------------------
    def lookup(self, f: _Field, idx: torch.Tensor):
        """Fill f.virt / f.vidx for the local index column ``idx`` [B] (any integer dtype)."""
        info = f.info
        R, B = info.world, f.B
        idx = idx.to(device=f.virt.device, dtype=torch.int64)
        f.bad.logical_or_(((idx < 0) | (idx >= f.vocab)).any())
        dist.all_gather_into_tensor(f.idx_all, idx.contiguous(), group=info.group)
        mine = (f.idx_all % R) == info.rank
        torch.where(mine, torch.div(f.idx_all, R, rounding_mode="floor"), torch.full_like(f.idx_all, -1), out=f.local)
        local_gather(f.shard.detach(), f.local, f.rows)
        dist.all_to_all_single(f.virt.detach(), f.rows, group=info.group)
        torch.add(torch.remainder(idx, R) * B, f.arange, out=f.vidx)

    # ---- backward half ------------------------------------------------------------------------------------
    def route_grad(self, f: _Field, gvirt: torch.Tensor, gshard: torch.Tensor):
        """gvirt [R*B, E]: dense gradient of the virtual table; accumulates this rank's rows into gshard."""
        info = f.info
        dist.all_to_all_single(f.grecv, gvirt.contiguous(), group=info.group)
        f.grecv.mul_(1.0 / info.world)          # per-rank losses are local means: average over the global batch
        local_scatter(f.grecv, f.local, gshard)

    def check_indices(self):
        for f in self.fields.values():
            if bool(f.bad):
                f.bad.zero_()
                raise IndexError(f"index out of range in self (sharded field {f.name!r})")


class ShardedLookup(torch.autograd.Function):
    """Autograd node of one sharded field on the generic (non-graph) path: shard parameter -> virtual table."""

    @staticmethod
    def forward(ctx, shard, idx, exchange: ShardedExchange, f: _Field):
        exchange.lookup(f, idx)
        ctx.exchange, ctx.f = exchange, f
        return f.virt.detach().view_as(f.virt)

    @staticmethod
    def backward(ctx, gvirt):
        f = ctx.f
        gshard = torch.zeros_like(f.shard)       # dense-gradient semantics of nn.Embedding, on the shard
        ctx.exchange.route_grad(f, gvirt, gshard)
        return gshard, None, None, None
