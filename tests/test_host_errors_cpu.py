"""Error behaviour of the host mirror at the drop-in boundary, checked without a GPU on the CPU interpreter
of the device programs (checker code injected by the test, as in tests/test_program_cpu.py):

* out-of-range index          -> IndexError   (nn.Embedding on CPU, basic/layers.py:70)
* batch of one in train mode  -> ValueError   (nn.BatchNorm1d: "Expected more than 1 value per channel when training")
* missing feature column      -> KeyError     (x[fea.name], basic/layers.py:70)
* empty feature lists         -> ValueError   (basic/layers.py:107,112)
* extra dict keys             -> ignored
* domain id outside [0, D)    -> output 0 for that row (mask-select idiom, base_example.py:61-77)
"""
import pytest
import torch

from golden_util import Golden
import model_factory
from oracle.ops_ref import RefRunner
from scenario_wise_rec_b200.fused import FusedModule
import scenario_wise_rec_b200.models.multi_domain as M


@pytest.fixture(autouse=True)
def ref_backend(monkeypatch):
    monkeypatch.setattr(FusedModule, "_runner_factory", RefRunner)


def _model(name="mmoe_small"):
    g = Golden(name)
    torch.manual_seed(0)
    m = model_factory.build(g.model, g.cfg)
    m.load_state_dict(g.state0)
    return g, m


def test_out_of_range_index_raises_index_error():
    g, m = _model()
    m.eval()
    x = {k: v.clone() for k, v in g.x.items()}
    col = next(k for k in x if k.startswith("s") and not x[k].dtype.is_floating_point)
    x[col][0] = 10 ** 6
    with pytest.raises(IndexError):
        with torch.no_grad():
            m(x)
            m.check_indices()


def test_batch_of_one_in_train_mode_raises_value_error():
    g, m = _model()
    m.train()
    x = {k: v[:1] for k, v in g.x.items()}
    with pytest.raises(ValueError, match="more than 1 value per channel"):
        m(x)
    m.eval()
    with torch.no_grad():
        out = m(x)          # eval mode uses the running statistics: a single row is fine
    assert out.shape == (1,)


def test_missing_column_raises_key_error():
    g, m = _model()
    m.eval()
    x = dict(g.x)
    col = next(k for k in x if k != "domain_indicator")
    del x[col]
    with pytest.raises(KeyError):
        with torch.no_grad():
            m(x)


def test_extra_keys_are_ignored():
    g, m = _model()
    m.eval()
    x = dict(g.x)
    x["not_a_feature"] = torch.zeros(g.B)
    with torch.no_grad():
        a, b = m(x), m(g.x)
    assert torch.equal(a, b)


def test_empty_feature_list_raises_value_error():
    g, _ = _model()
    with pytest.raises(ValueError):
        m = M.MMOE([], g.cfg["domain_num"], n_expert=g.cfg["n_expert"], expert_params={"dims": list(g.cfg["expert_dims"])},
                   tower_params={"dims": list(g.cfg["tower_dims"])})
        m.eval()
        with torch.no_grad():
            m({"domain_indicator": torch.zeros(4, dtype=torch.int64)})


def test_domain_id_outside_range_yields_zero():
    g, m = _model()
    m.eval()
    x = {k: v.clone() for k, v in g.x.items()}
    x["domain_indicator"][:3] = g.cfg["domain_num"] + 2
    with torch.no_grad():
        out = m(x)
    assert torch.all(out[:3] == 0) and torch.all(out[3:] > 0)
