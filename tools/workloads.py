"""Synthetic workloads shaped like BASELINE.json's configs.  Shared by
bench.py, the -m gpu parity tests and the development probes.  Feature specs are plain tuples
``(name, kind, vocab, dim)`` understood by tests/model_factory.py (product models) and by
oracle/ref_models.py (CPU checker).  Data is synthetic: there is no network for the datasets.
"""
from __future__ import annotations

import torch

# per-field vocabularies used by the survey probe: two large fields (~238 k users,
# ~467 k items, README.md:138 of the reference), several 50-260 k fields, ~10 tiny ones; 1.56 M rows.
ALI_VOCAB = [238635, 98, 14, 3, 8, 4, 4, 3, 5, 467298, 6929, 263942, 80232, 106399, 5888, 104830, 51878, 37148,
             3, 5853, 105622, 53843, 31858]


def ali_ccp_features(scale=1, embed_dim=16, item_rows=None):
    """item_rows: size of the item table (field s9); BASELINE.json configs[3] quotes 85 M rows (README.md:138 of the
    reference), row-sharded across 8 GPUs."""
    f = [(f"D{i}", "dense", 0, 1) for i in range(8)]
    f += [(f"s{i}", "sparse", (item_rows if (item_rows and i == 9) else max(2, v // scale)), embed_dim) for i, v in enumerate(ALI_VOCAB)]
    return f


def kuairand_features(big=4_000_000):
    vocab = [big, 1000] + [50 + 45 * i for i in range(30)]
    return [(f"s{i}", "sparse", v, 16) for i, v in enumerate(vocab)] + [(f"D{i}", "dense", 0, 1) for i in range(4)]


def ml1m_features():
    return [("user_id", "sparse", 6041, 16), ("movie_id", "sparse", 3953, 16), ("gender", "sparse", 3, 16),
            ("age", "sparse", 8, 16), ("occupation", "sparse", 22, 16), ("zip", "sparse", 3440, 16), ("d0", "dense", 0, 1)]


def mind_features(scale=1):
    return [("user", "sparse", 748_000 // scale, 64), ("item", "sparse", 20_000 // scale, 64), ("cat", "sparse", 300, 64)]


def _ali_split(item_rows=None):
    """PPNet / EPNet feature split of scripts/run_ali_ccp_ctr_ranking_multi_domain.py:155-158 of the
    reference: id = user + item fields, scenario = field '301' (s18 here, vocab 3), agnostic = the rest."""
    feats = ali_ccp_features(item_rows=item_rows)
    by = {f[0]: f for f in feats}
    idf = [by["s0"], by["s9"]]
    sce = [by["s18"]]
    agn = [f for f in feats if f[0] not in ("s0", "s9", "s18")]
    return idf, sce, agn


ITEM_85M = 85_000_000


def _cases():
    idf, sce, agn = _ali_split()
    idf85, sce85, agn85 = _ali_split(ITEM_85M)
    deep = [256, 128, 64, 32, 16, 8]
    return {
        # BASELINE.json configs[3]: the 85 M-row item table (5.44 GB fp32), row-sharded when run on several GPUs
        "cfg4a_star_aliccp85m_b4096": ("Star", dict(features=ali_ccp_features(item_rows=ITEM_85M), domain_num=3, fcn_dims=deep, aux_dims=[16]), 4096),
        "cfg4b_ppnet_aliccp85m_b4096": ("PPNet", dict(id_features=idf85, agn_features=agn85 + sce85, domain_num=3, fcn_dims=deep), 4096),
        "cfg2_mmoe_aliccp85m_b4096": ("MMOE", dict(features=ali_ccp_features(item_rows=ITEM_85M), domain_num=3, n_expert=4, expert_dims=deep,
                                                   tower_dims=[16]), 4096),
        # name: (model, cfg, B)
        "cfg1_sharedbottom_ml1m_b256": ("SharedBottom", dict(features=ml1m_features(), domain_num=3, bottom_dims=[128],
                                                             tower_dims=[8]), 256),
        "cfg2_mmoe_aliccp_b4096": ("MMOE", dict(features=ali_ccp_features(), domain_num=3, n_expert=4, expert_dims=deep,
                                                tower_dims=[16]), 4096),
        "cfg3_ple_kuairand_b8192": ("PLE", dict(features=kuairand_features(), domain_num=5, n_level=1, n_expert_specific=2,
                                                n_expert_shared=2, expert_dims=[64, 32], tower_dims=[16]), 8192),
        "cfg4a_star_aliccp_b4096": ("Star", dict(features=ali_ccp_features(), domain_num=3, fcn_dims=deep, aux_dims=[16]), 4096),
        "cfg4b_ppnet_aliccp_b4096": ("PPNet", dict(id_features=idf, agn_features=agn + sce, domain_num=3, fcn_dims=deep), 4096),
        "cfg4c_epnet_aliccp_b4096": ("EPNet", dict(sce_features=sce, agn_features=agn, domain_num=3, fcn_dims=deep), 4096),
        "cfg5a_hamursmall_mind_b16384": ("HamurSmall", dict(features=mind_features(), domain_num=4, fcn_dims=[256, 128],
                                                            hyper_dims=[64], k=35), 16384),
        # HamurLarge at the Ali-CCP shape (hamur.py:101-244: seven backbone layers, two adapter cells, k = 65)
        "cfg5c_hamurlarge_aliccp_b2048": ("HamurLarge", dict(features=ali_ccp_features(), domain_num=3, fcn_dims=[256, 256, 128, 128, 64, 64, 32],
                                                             hyper_dims=[64], k=65), 2048),
        "cfg5b_m3oe_mind_b16384": ("M3oE", dict(features=mind_features(), domain_num=4, fcn_dims=[128, 64, 64, 32],
                                                expert_num=4), 16384),
    }


CASES = _cases()
DEFAULT_CASE = "cfg2_mmoe_aliccp_b4096"


def all_feature_specs(cfg):
    seen, out = set(), []
    for key in ("features", "id_features", "agn_features", "sce_features"):
        for s in cfg.get(key, []):
            if s[0] not in seen:
                seen.add(s[0])
                out.append(s)
    return out


def make_batch(feats, B, D, seed, device="cpu", zipf=False, pin=False, pos_rate=0.3):
    """Seeded synthetic batch: int64 index columns (uniform, or skewed towards hot rows with
    ``zipf``), fp32 U(0,1) dense columns, uniform ``domain_indicator`` and Bernoulli labels."""
    g = torch.Generator().manual_seed(seed)
    x = {}
    for name, kind, vocab, _dim in feats:
        if kind == "sparse":
            if zipf:
                r = torch.rand(B, generator=g)
                x[name] = (vocab * r ** 3).long().clamp_(0, vocab - 1)
            else:
                x[name] = torch.randint(0, vocab, (B,), generator=g)
        else:
            x[name] = torch.rand(B, generator=g)
    x["domain_indicator"] = torch.randint(0, D, (B,), generator=g)
    y = (torch.rand(B, generator=g) < pos_rate).float()
    if pin:
        x = {k: v.pin_memory() for k, v in x.items()}
        y = y.pin_memory()
    return {k: v.to(device) for k, v in x.items()}, y.to(device)


def gather_bytes_per_sample(feats):
    """F_s*(E*4 + 8) + F_d*4 + IN*4 (index + row read + dense read + output write)."""
    fs = [(v, d) for _, k, v, d in feats if k == "sparse"]
    fd = sum(1 for _, k, _, _ in feats if k == "dense")
    IN = sum(d for _, d in fs) + fd
    return sum(d * 4 + 8 for _, d in fs) + fd * 4 + IN * 4


def scatter_bytes_per_sample(feats):
    """IN_sparse*4 grad read + F_s*8 idx + 2*F_s*E*4 read-modify-write of the rows."""
    fs = [d for _, k, _, d in feats if k == "sparse"]
    return sum(fs) * 4 + len(fs) * 8 + 2 * sum(fs) * 4


def table_bytes(feats):
    return sum(v * d * 4 for _, k, v, d in feats if k == "sparse")
