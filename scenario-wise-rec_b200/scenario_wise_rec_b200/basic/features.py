"""Feature config objects (reference: scenario_wise_rec/basic/features.py:49-95).

``SparseFeature`` owns -- and caches on itself -- the ``nn.Embedding`` of its table, so
two models / embedding layers built from the same feature objects share one table, as in
the reference (features.py:76-79).  Sequence features are outside the hot path in scope
(no reference script or model uses them).
"""
from .initializers import RandomNormal


def get_auto_embedding_dim(num_classes):
    """floor(6 * num_classes ** 0.26) -- reference utils/data.py:65-75 (the exponent in the
    code is 0.26, the docstring there says 1/4)."""
    import math
    return int(math.floor(6 * math.pow(num_classes, 0.26)))


class SparseFeature(object):
    def __init__(self, name, vocab_size, embed_dim=None, shared_with=None, padding_idx=None,
                 initializer=RandomNormal(0, 0.0001)):
        self.name = name
        self.vocab_size = int(vocab_size)
        self.embed_dim = get_auto_embedding_dim(vocab_size) if embed_dim is None else int(embed_dim)
        self.shared_with = shared_with
        self.padding_idx = padding_idx
        self.initializer = initializer
        self.shard = None           # parallel.ShardInfo when the table is row-sharded across ranks (parallel.shard_features)

    def __repr__(self):
        return f"<SparseFeature {self.name} with Embedding shape ({self.vocab_size}, {self.embed_dim})>"

    def get_embedding_layer(self):
        if not hasattr(self, "embed"):
            rows = self.vocab_size if self.shard is None else self.shard.local_rows(self.vocab_size)
            self.embed = self.initializer(rows, self.embed_dim)
        return self.embed


class DenseFeature(object):
    def __init__(self, name):
        self.name = name
        self.embed_dim = 1

    def __repr__(self):
        return f"<DenseFeature {self.name}>"


class SequenceFeature(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("SequenceFeature pooling is outside the accelerated hot path "
                                  "(unused by every reference model and script)")
