"""Host-side lowering checked without a GPU: product model -> device program records ->
oracle/ops_ref.RefRunner (torch CPU interpreter of the records) == reference golden vectors."""
import pytest
import torch

from golden_util import Golden, golden_names
import model_factory
from oracle.ops_ref import RefRunner
from scenario_wise_rec_b200.fused import FusedModule


@pytest.fixture()
def ref_backend(monkeypatch):
    monkeypatch.setattr(FusedModule, "_runner_factory", RefRunner)


@pytest.mark.parametrize("name", golden_names())
def test_program_matches_reference(name, ref_backend):
    g = Golden(name)
    if not model_factory.supported(g.model):
        pytest.skip(f"{g.model} not lowered yet")
    torch.manual_seed(0)
    model = model_factory.build(g.model, g.cfg)
    missing, unexpected = model.load_state_dict(g.state0, strict=True)
    assert not missing and not unexpected
    model_factory.check_against_golden(model, g)


def test_product_refuses_cpu():
    """Without the injected checker the product path must refuse to run on CPU tensors."""
    g = Golden("sharedbottom_small")
    model = model_factory.build(g.model, g.cfg)
    with pytest.raises(RuntimeError):
        model(g.x)
