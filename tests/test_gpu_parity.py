"""-m gpu: the CUDA path (through the C ABI) against the oracle.

* golden fixtures (outputs / gradients / BN buffers generated from the unmodified reference),
* op-by-op: every workspace buffer of the device program vs the torch-CPU interpreter,
* BASELINE.json shapes (cfg1..cfg3) vs oracle/ref_models.py run live on the host,
* size-independent properties at large sizes (gather bit-exactness, scatter conservation).
Tolerances: gather / indices bit-exact; outputs 1e-4 abs fp32 (north_star); gradients absolute,
relative to max|grad| of the tensor (a bias feeding BatchNorm has an analytically zero gradient).
"""
import copy
import ctypes
import os

import numpy as np
import pytest
import torch

from golden_util import Golden, golden_names
import gpu_util
import model_factory
from oracle import ref_models
from oracle.ops_ref import RefRunner
from scenario_wise_rec_b200 import _native as N
from scenario_wise_rec_b200.program import CudaRunner

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture
def fc_mode(request):
    """FC arithmetic for one test (swr_set_fc_mode): FC_SIMT = fp32 FFMA, FC_AUTO = default (tcgen05 on wide layers),
    FC_TC = tcgen05 on every layer whose batch reaches 64 rows (odd widths, unaligned weights, partial k-blocks)."""
    prev = N.set_fc_mode(request.param)
    yield request.param
    N.set_fc_mode(prev)


FC_MODES = pytest.mark.parametrize("fc_mode", [N.FC_AUTO, N.FC_TC], ids=["auto", "tcgen05-everywhere"], indirect=True)


def test_library_loads_on_device():
    assert N.lib().swr_abi_version() == N.ABI_VERSION
    assert N.lib().swr_device_check() == 0, N.last_error()


@FC_MODES
@pytest.mark.parametrize("name", golden_names())
def test_golden(name, fc_mode):
    g = Golden(name)
    if not model_factory.supported(g.model):
        pytest.skip(f"{g.model} not lowered yet")
    model = model_factory.build(g.model, g.cfg)
    model.load_state_dict(g.state0)
    model.to(DEV)
    before = N.launch_count()
    model_factory.check_against_golden(model, g, device=DEV)
    assert N.launch_count() > before, "no kernel of libswr_b200.so was launched"


@FC_MODES
@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("training", [True, False])
def test_every_buffer_matches_interpreter(name, training, fc_mode):
    """Runs the same records on both executors and compares every activation, statistic,
    gradient buffer and parameter gradient."""
    g = Golden(name)
    if not model_factory.supported(g.model):
        pytest.skip(f"{g.model} not lowered yet")
    model = model_factory.build(g.model, g.cfg)
    model.load_state_dict(g.state0)
    model.train(training)
    model_ref = copy.deepcopy(model)
    model.to(DEV)
    xg = {k: v.to(DEV) for k, v in g.x.items()}
    rc = model._runner(xg)
    rr = RefRunner(_prog(model_ref, g.x))
    oc = rc.forward(xg)
    orf = rr.forward(g.x)
    torch.cuda.synchronize()
    report = []
    fails = gpu_util.compare_slots(rc, rr, [".raw", ".stats", ".probs", "head.out"], report=report)
    assert not fails, f"forward buffers differ: {fails[:5]}"
    gout = torch.linspace(-1, 1, g.B)
    gc = rc.backward([gout.to(DEV)] + [None] * (len(oc) - 1))
    gr = rr.backward([gout] + [None] * (len(orf) - 1))
    torch.cuda.synchronize()
    fails = gpu_util.compare_slots(rc, rr, [".dz", ".dstats"], atol=1e-5, rtol=2e-4)
    assert not fails, f"backward buffers differ: {fails[:5]}"
    gmax = max(float(b.abs().max()) for b in gr)
    for p, a, b in zip(rc.prog.params, gc, gr):
        scale = max(float(b.abs().max()), 1e-3)
        err = float((a.cpu() - b).abs().max())
        # a bias feeding a norm has an analytically zero gradient (sum of cancelling terms): absolute floor
        # relative to the largest gradient of the model
        assert err <= 2e-4 * scale + 1e-5 * gmax + 1e-6, f"param grad {tuple(p.shape)} err {err} scale {scale}"


def _prog(model, x):
    from scenario_wise_rec_b200.program import ProgramBuilder
    cols = model._columns()
    b = ProgramBuilder(int(x[cols[0]].shape[0]), model.training)
    model._lower(b, {c: x[c].dtype for c in cols})
    return b.finish()


# --------------------------------------------------------------------------------------------
# BASELINE.json shapes against the oracle run live on the host cores
# --------------------------------------------------------------------------------------------
import workloads

BASELINE_CASES = dict(workloads.CASES)


def _oracle(model_name, cfg, x, y, state, dt):
    st = {}
    for k, v in state.items():
        v = v.clone()
        if v.dtype.is_floating_point:
            v = v.to(dt)
            if "running_" not in k:
                v.requires_grad_(True)
        st[k] = v
    xx = {k: (v.to(dt) if v.dtype.is_floating_point else v) for k, v in x.items()}
    bn_out = {}
    out = ref_models.forward(model_name, xx, st, cfg, training=True, bn_out=bn_out)
    ref_models.bce_loss(out, y.to(dt)).backward()
    return out.detach(), {k: v.grad for k, v in st.items() if v.requires_grad}, bn_out


@pytest.mark.parametrize("fc_mode", [N.FC_SIMT, N.FC_AUTO], ids=["ffma", "tcgen05"], indirect=True)
@pytest.mark.parametrize("case", sorted(BASELINE_CASES))
def test_baseline_shapes_vs_oracle(case, fc_mode):
    """BASELINE.json shapes, once per FC arithmetic (swr_set_fc_mode).  Outputs: <= 1e-4 abs against the fp32 CPU oracle (north_star).  Gradients: judged
    against the oracle evaluated in float64, with the fp32 CPU oracle's own distance to it as the noise floor --
    deep BatchNorm stacks amplify fp32 rounding (STAR cfg4a: the CPU reference itself is 1e-2 of max|g| away from
    float64), and a ReLU whose pre-activation lands within an ulp of 0 flips its derivative for one row (measure
    zero, seen once in 16384 rows on M3oE), which a max-norm alone cannot tell from a bug; such a tensor must
    still agree in the relative L2 norm with a vanishing fraction of outlier elements.

    tcgen05 mode (the default): operands are split 3xTF32 (~2^-21 per product) and the tensor core accumulates with
    round-toward-zero instead of round-to-nearest: a truncation of about 1e-8 * K relative, *coherent in the sign of
    the products* (tools/tc_probe.cu measures it).  Outputs stay inside 1e-4, but a gradient that sums many
    same-signed per-sample terms which nearly cancel over the batch (30 % positive labels here) keeps the bias
    while the sum shrinks: up to ~5e-3 of max|g| on such components (the fp32 CPU reference itself is 1e-3 from
    float64 there).  Its per-element error (~1e-6 relative, ten times fp32's) also moves a few more pre-activations
    across 0 than the fp32 paths do (measured on cfg2: rows 495, 717, 1613 of 4096, tests/debug_case.py); each
    flipped row changes a weight gradient by that sample's own contribution, ~1/sqrt(B) of max|g|, spread over
    the whole tensor.  The tensor-core bound is therefore max-err <= max(1e-2, 8 x fp32 noise, 2/sqrt(B)) and
    relative L2 <= max(1e-2, 4 x fp32 noise, 2/sqrt(B)); a tiling / indexing bug is O(1) in both, and the small golden cases, the
    buffer-by-buffer comparison against the interpreter and the 1e-4 output bound stay as tight as for FFMA."""
    model_name, cfg, B = BASELINE_CASES[case]
    torch.manual_seed(7)
    model = model_factory.build(model_name, cfg)
    gpu_util.randomise(model, 11)
    state = {k: v.clone() for k, v in model.state_dict().items()}
    x, y = gpu_util.make_batch(workloads.all_feature_specs(cfg), B, cfg["domain_num"], seed=5, zipf=True)
    ref, g32, bn_out = _oracle(model_name, cfg, x, y, state, torch.float32)
    _r64, g64, _ = _oracle(model_name, cfg, x, y, state, torch.float64)
    # CUDA path
    model.to(DEV).train()
    out = model({k: v.to(DEV) for k, v in x.items()})
    torch.nn.BCELoss()(out, y.to(DEV)).backward()
    err = float((out.detach().cpu() - ref).abs().max())
    assert err <= 1e-4, f"output err {err}"
    for k, p in model.named_parameters():
        t = g64[k]
        assert (t is None) == (p.grad is None), k
        if t is None:
            continue
        ours = p.grad.cpu().double()
        scale = max(float(t.abs().max()), 1e-3)
        e_ours = float((ours - t).abs().max()) / scale
        e_ref = float((g32[k].double() - t).abs().max()) / scale
        if e_ours <= max(5e-4, 4 * e_ref):
            continue
        nrm = max(float(t.norm()), 1e-12)
        l2_ours, l2_ref = float((ours - t).norm()) / nrm, float((g32[k].double() - t).norm()) / nrm
        outliers = float(((ours - t).abs() > 5e-4 * scale).double().mean())
        flip = 2.0 / B ** 0.5
        # SWR_TEST_TC_STRICT=1 judges the tensor-core mode by the FFMA criteria below: with the accumulator flushes of
        # swr_fc_tc.cu 9 of the 11 cases pass them (round 1: 5 of 8); the two that do not are ReLU'(0) flips (a rank-1
        # change of one layer's weight gradient: many small outliers, rel-L2 7.5e-3)
        if not os.environ.get("SWR_TEST_TC_STRICT") and fc_mode != N.FC_SIMT and e_ours <= max(1e-2, 8 * e_ref, flip) and l2_ours <= max(1e-2, 4 * l2_ref, flip):
            continue
        # one flipped row perturbs one row / column of a weight gradient or one row of a table whose gradient
        # lives on <= B rows: rel-L2 up to ~1/sqrt(B) * O(1); a tiling / indexing bug shows up as O(1) instead
        assert l2_ours <= max(1e-2, 4 * l2_ref) and outliers <= 2e-2, \
            f"gradient {k}: max-err {e_ours:.2e} (cpu fp32 noise {e_ref:.2e}), rel-L2 {l2_ours:.2e} (noise {l2_ref:.2e}), outliers {outliers:.2e}"
    sd = model.state_dict()
    for k, v in bn_out.items():
        torch.testing.assert_close(sd[k].cpu().to(v.dtype), v, atol=2e-5, rtol=1e-4)
    # eval mode
    model.load_state_dict(state)
    model.eval()
    with torch.no_grad():
        oe = model({k: v.to(DEV) for k, v in x.items()})
        re = ref_models.forward(model_name, x, state, cfg, training=False)
    assert float((oe.cpu() - re).abs().max()) <= 1e-4


# --------------------------------------------------------------------------------------------
# K1 / K2 through the raw C ABI
# --------------------------------------------------------------------------------------------
def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _gather_abi(tables, idx, dense, B, E):
    out = torch.empty(B, len(tables) * E + len(dense), device=DEV)
    vocab = (ctypes.c_int64 * len(tables))(*[t.shape[0] for t in tables])
    idt = (ctypes.c_int32 * len(idx))(*[N.torch_dtype_code(t.dtype) for t in idx])
    ddt = (ctypes.c_int32 * max(len(dense), 1))(*[N.torch_dtype_code(t.dtype) for t in dense])
    oob = torch.zeros(2, dtype=torch.int32, device=DEV)
    st = N.lib().swr_embedding_gather_fwd(_ptr_array(tables), vocab, _ptr_array(idx), idt, _ptr_array(dense) if dense else None,
                                          ddt, out.data_ptr(), out.shape[1], B, len(tables), E, len(dense), oob.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream)
    N.check(st, "gather")
    return out, oob


@pytest.mark.parametrize("B,E", [(1, 16), (37, 16), (4096, 16), (1000, 64), (513, 6)])
def test_gather_bit_exact(B, E):
    g = torch.Generator().manual_seed(B + E)
    vocabs = [7, 300, 5000, 3, 12345]
    tables = [torch.randn(v, E, generator=g).to(DEV) for v in vocabs]
    dts = [torch.int64, torch.int32, torch.int16, torch.int8, torch.int64]
    idx = [torch.randint(0, min(v, 127 if dt == torch.int8 else v), (B,), generator=g).to(dt).to(DEV) for v, dt in zip(vocabs, dts)]
    dense = [torch.rand(B, generator=g).to(DEV), torch.rand(B, generator=g).half().to(DEV), torch.rand(B, generator=g).double().to(DEV)]
    out, oob = _gather_abi(tables, idx, dense, B, E)
    ref = torch.cat([t[i.long()] for t, i in zip(tables, idx)] + [d.float().unsqueeze(1) for d in dense], dim=1)
    assert torch.equal(out, ref)          # a gather is a copy: bit-exact
    assert int(oob[0]) == 0


def test_gather_out_of_range_flag():
    tables = [torch.randn(10, 16, device=DEV), torch.randn(5, 16, device=DEV)]
    idx = [torch.tensor([1, 2, 3], device=DEV), torch.tensor([0, 5, -1], device=DEV)]
    out, oob = _gather_abi(tables, idx, [], 3, 16)
    assert int(oob[0]) == 1 and int(oob[1]) == 1
    assert torch.equal(out[1, 16:], torch.zeros(16, device=DEV))
    assert torch.equal(out[0, 16:], tables[1][0])


@pytest.mark.parametrize("B,E,vocabs", [(4096, 16, [3, 40, 5000, 200000]), (8192, 64, [20, 748000]), (100, 6, [9, 50])])
def test_scatter_matches_index_add(B, E, vocabs):
    g = torch.Generator().manual_seed(B)
    idx = [(v * torch.rand(B, generator=g) ** 3).long().clamp_(0, v - 1).to(DEV) for v in vocabs]   # skewed: hot rows
    grad = torch.randn(B, len(vocabs) * E, generator=g).to(DEV)
    gt = [torch.zeros(v, E, device=DEV) for v in vocabs]
    vocab = (ctypes.c_int64 * len(vocabs))(*vocabs)
    idt = (ctypes.c_int32 * len(idx))(*[N.DT_I64] * len(idx))
    st = N.lib().swr_embedding_scatter_bwd(grad.data_ptr(), grad.shape[1], B, _ptr_array(idx), idt, _ptr_array(gt), vocab,
                                           len(vocabs), E, torch.cuda.current_stream().cuda_stream)
    N.check(st, "scatter")
    for f, v in enumerate(vocabs):
        ref = torch.zeros(v, E, dtype=torch.float64, device=DEV).index_add_(0, idx[f], grad[:, f * E:(f + 1) * E].double())
        # fp32 atomics reorder the sums: tolerance ~ eps * count * |g|
        torch.testing.assert_close(gt[f].double(), ref, atol=2e-4, rtol=1e-5)
    # conservation (size-independent property): total mass is preserved per field
    for f in range(len(vocabs)):
        assert abs(float(gt[f].double().sum() - grad[:, f * E:(f + 1) * E].double().sum())) < 1e-2


def test_gather_scatter_large_roundtrip():
    """2^22 lookups: gather is bit-exact; scatter of the gathered rows' all-ones gradient counts occurrences."""
    B, E, V = 1 << 22, 16, 1 << 20
    g = torch.Generator().manual_seed(3)
    table = torch.randn(V, E, generator=g).to(DEV)
    idx = torch.randint(0, V, (B,), generator=g).to(DEV)
    out, _ = _gather_abi([table], [idx], [], B, E)
    assert torch.equal(out, table[idx])
    gt = torch.zeros(V, E, device=DEV)
    ones = torch.ones(B, E, device=DEV)
    vocab = (ctypes.c_int64 * 1)(V)
    idt = (ctypes.c_int32 * 1)(N.DT_I64)
    N.check(N.lib().swr_embedding_scatter_bwd(ones.data_ptr(), E, B, _ptr_array([idx]), idt, _ptr_array([gt]), vocab, 1, E,
                                               torch.cuda.current_stream().cuda_stream), "scatter")
    counts = torch.bincount(idx, minlength=V).float()
    assert torch.equal(gt[:, 0], counts) and torch.equal(gt[:, E - 1], counts)   # small integers: exact in fp32


def test_embedding_layer_standalone():
    from scenario_wise_rec_b200.basic.layers import EmbeddingLayer
    feats = model_factory.features([("a", "sparse", 11, 8), ("d", "dense", 0, 1), ("b", "sparse", 5, 8)])
    emb = EmbeddingLayer(feats).to(DEV)
    x = {"a": torch.tensor([0, 10, 3], device=DEV), "b": torch.tensor([4, 4, 1], device=DEV), "d": torch.tensor([.5, .25, 1.], device=DEV)}
    out = emb(x, feats, squeeze_dim=True)
    ref = torch.cat([emb.embed_dict["a"].weight[x["a"]], emb.embed_dict["b"].weight[x["b"]], x["d"].unsqueeze(1)], 1)
    assert torch.equal(out.detach(), ref.detach())
    out.sum().backward()
    ga = torch.zeros(11, 8, device=DEV).index_add_(0, x["a"], torch.ones(3, 8, device=DEV))
    assert torch.equal(emb.embed_dict["a"].weight.grad, ga)
    out3 = emb(x, feats, squeeze_dim=False)
    assert out3.shape == (3, 2, 8)
    x["a"] = torch.tensor([0, 11, 3], device=DEV)
    emb(x, feats, squeeze_dim=True)
    with pytest.raises(IndexError):
        emb.check_indices()


# --------------------------------------------------------------------------------------------
# run-to-run repeatability of the persistent tcgen05 kernels with many tiles per CTA
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("training", [False, True])
def test_tensor_core_forward_is_repeatable(training):
    """HamurSmall at B = 16384 with every layer on the tcgen05 kernels: 10+ tiles per persistent CTA, tiles of mixed cost
    dealt out through the permutation table.  Eight forwards of the same batch must agree (the only run-to-run freedom is
    the order of the fp64 statistics atomics, ~1e-7 on the output).  This is the check that exposed a shared-memory tile
    cache as non-deterministic during round 2 (profiles/r02_fc_tc2_notes.md); a race in the role handshakes shows up
    here as differences of 1e-2 and more."""
    prev = N.set_fc_mode(N.FC_TC)
    try:
        model_name, cfg, B = BASELINE_CASES["cfg5a_hamursmall_mind_b16384"]
        x, _y = gpu_util.make_batch(workloads.all_feature_specs(cfg), B, cfg["domain_num"], seed=5, zipf=True)
        torch.manual_seed(7)
        model = model_factory.build(model_name, cfg)
        gpu_util.randomise(model, 11)
        model.to(DEV).train(training)
        xg = {k: v.to(DEV) for k, v in x.items()}
        with torch.no_grad():
            outs = [model(xg).clone() for _ in range(8)]
        torch.cuda.synchronize()
        ref = torch.stack(outs).median(0).values
        worst = max(float((o - ref).abs().max()) for o in outs)
        assert worst <= 1e-5, worst
    finally:
        N.set_fc_mode(prev)
