"""Pins oracle/ref_models.py to outputs of the unmodified reference (tests/golden)."""
import pytest
import torch

from golden_util import Golden, golden_names
from oracle import ref_models


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference(name):
    g = Golden(name)
    st = g.leaf_state()
    bn_out = {}
    out = ref_models.forward(g.model, g.x, st, g.cfg, training=True, bn_out=bn_out)
    assert out.shape == g.out_train.shape
    # tolerance: fp32 reassociation only (same ATen kernels underneath)
    torch.testing.assert_close(out.detach(), g.out_train, atol=2e-6, rtol=1e-5)
    loss = ref_models.bce_loss(out, g.y)
    assert abs(float(loss.detach()) - g.loss) < 1e-5
    loss.backward()
    for k in g.param_names():
        if k in g.grad_none:
            assert st[k].grad is None, k
        else:
            assert st[k].grad is not None, k
            gmax = float(g.grads[k].abs().max()) + 1e-12
            err = float((st[k].grad - g.grads[k]).abs().max())
            assert err <= 1e-5 * max(gmax, 1.0) + 2e-7, (k, err, gmax)
    for k, v in g.state1.items():
        assert k in bn_out, k
        torch.testing.assert_close(bn_out[k].to(v.dtype), v, atol=1e-6, rtol=1e-5)
    with torch.no_grad():
        out_eval = ref_models.forward(g.model, g.x, g.state0, g.cfg, training=False)
    torch.testing.assert_close(out_eval, g.out_eval, atol=2e-6, rtol=1e-5)


def test_embedding_out_of_range_raises():
    g = Golden("sharedbottom_small")
    x = dict(g.x)
    x["s1"] = x["s1"].clone()
    x["s1"][3] = 7          # vocab is 7
    with pytest.raises(IndexError):
        ref_models.forward(g.model, x, g.state0, g.cfg, training=False)
