"""CTRTrainer.evaluate / evaluate_multi_domain_loss / predict keep the predictions on the device until the loader is
exhausted (one copy instead of a ``.tolist()`` sync per batch, reference ctr_trainer.py:99-171) and must hand the metric
functions exactly what the reference loop hands them.  Checked on the CPU interpreter of the device programs."""
import pytest
import torch
from sklearn.metrics import log_loss, roc_auc_score

from golden_util import Golden
import model_factory
from oracle.ops_ref import RefRunner
from scenario_wise_rec_b200.fused import FusedModule
from scenario_wise_rec_b200.trainers import CTRTrainer


@pytest.fixture(autouse=True)
def ref_backend(monkeypatch):
    monkeypatch.setattr(FusedModule, "_runner_factory", RefRunner)


def _loader(g, bs):
    return [({k: v[i:i + bs] for k, v in g.x.items()}, g.y[i:i + bs]) for i in range(0, g.B, bs)]


def _reference_loop(model, loader, domain_num):
    """The reference's evaluation loops, verbatim in structure (per-batch tolist)."""
    model.eval()
    t, p = [], []
    td, pd = [[] for _ in range(domain_num)], [[] for _ in range(domain_num)]
    with torch.no_grad():
        for x, y in loader:
            yp = model(x)
            t.extend(y.tolist()); p.extend(yp.tolist())
            for d in range(domain_num):
                m = x["domain_indicator"] == d
                td[d].extend(y[m].tolist()); pd[d].extend(yp[m].tolist())
    return t, p, td, pd


def test_evaluate_matches_reference_loop():
    g = Golden("mmoe_small")
    torch.manual_seed(0)
    model = model_factory.build(g.model, g.cfg)
    model.load_state_dict(g.state0)
    tr = CTRTrainer(model, "t", device="cpu", fused=False)
    D = g.cfg["domain_num"]
    y = (torch.arange(g.B) % 3 == 0).float()          # both classes in every domain
    gx = Golden("mmoe_small")
    gx.y = y
    loader = _loader(gx, 17)                           # ragged last batch
    t, p, td, pd = _reference_loop(model, loader, D)
    auc, ll = tr.evaluate(model, loader)
    assert auc == roc_auc_score(t, p) and ll == log_loss(t, p)
    ll_d, auc_d, ll_all, auc_all = tr.evaluate_multi_domain_loss(model, loader, D)
    assert ll_all == log_loss(t, p) and auc_all == roc_auc_score(t, p)
    for d in range(D):
        assert ll_d[d] == log_loss(td[d], pd[d]) and auc_d[d] == roc_auc_score(td[d], pd[d])
    assert tr.predict(model, loader) == p
    # a domain with no rows yields None, an empty loader yields Nones
    ll_d, auc_d, _, _ = tr.evaluate_multi_domain_loss(model, loader, D + 1)
    assert ll_d[D] is None and auc_d[D] is None
    assert tr.evaluate_multi_domain_loss(model, [], D) == ([None] * D, [None] * D, None, None)
    assert tr.predict(model, []) == []


def test_device_metric_definitions_match_sklearn():
    """trainers/metrics.py (what CTRTrainer.evaluate uses on a CUDA device) against sklearn on labels / scores with ties,
    exact 0 and 1 probabilities and a skewed class balance; and sklearn's error behaviour for a single class."""
    from scenario_wise_rec_b200.trainers.metrics import binary_auc, binary_logloss
    g = torch.Generator().manual_seed(3)
    for n in (5, 257, 20000):
        y = (torch.rand(n, generator=g) < 0.2).float()
        y[0], y[1] = 1.0, 0.0
        p = torch.rand(n, generator=g)
        p[::7] = p[0]                      # ties
        p[2], p[3] = 0.0, 1.0              # clipped by log_loss
        assert abs(binary_auc(y, p) - roc_auc_score(y.tolist(), p.tolist())) <= 1e-12
        ll = log_loss(y.tolist(), p.tolist())
        assert abs(binary_logloss(y, p) - ll) <= 1e-12 * max(1.0, abs(ll))
    ones = torch.ones(8)
    with pytest.raises(ValueError):
        binary_auc(ones, torch.rand(8))
    with pytest.raises(ValueError):
        binary_logloss(ones, torch.rand(8))


def test_custom_evaluate_fn_keeps_the_list_path():
    """A user-supplied metric keeps receiving Python lists (the device metrics replace only the default AUC)."""
    from scenario_wise_rec_b200.trainers import CTRTrainer
    assert CTRTrainer.evaluate_fn is CTRTrainer.evaluate_fn
    t = CTRTrainer.__new__(CTRTrainer)
    t.device = torch.device("cpu")
    assert not t._device_metrics()
    t.device = torch.device("cuda", 0)
    assert t._device_metrics()
    t.evaluate_fn = lambda a, b: 0.5
    assert not t._device_metrics()
