"""Embedding initialisers (reference: scenario_wise_rec/basic/initializers.py:4-92).

Each initialiser is a callable ``(vocab_size, embed_dim) -> torch.nn.Embedding``.  They
consume the torch RNG exactly like the reference (``nn.Embedding`` construction draws
N(0,1) first, then the initialiser overwrites), so a fixed seed yields identical tables.
"""
import torch


class _Init:
    def _make(self, vocab_size, embed_dim):
        return torch.nn.Embedding(vocab_size, embed_dim)


class RandomNormal(_Init):
    """N(mean, std) -- the default of SparseFeature is RandomNormal(0, 1e-4) (features.py:62)."""

    def __init__(self, mean=0.0, std=1.0):
        self.mean, self.std = mean, std

    def __call__(self, vocab_size, embed_dim):
        emb = self._make(vocab_size, embed_dim)
        torch.nn.init.normal_(emb.weight, self.mean, self.std)
        return emb


class RandomUniform(_Init):
    def __init__(self, minval=0.0, maxval=1.0):
        self.minval, self.maxval = minval, maxval

    def __call__(self, vocab_size, embed_dim):
        emb = self._make(vocab_size, embed_dim)
        torch.nn.init.uniform_(emb.weight, self.minval, self.maxval)
        return emb


class XavierNormal(_Init):
    def __init__(self, gain=1.0):
        self.gain = gain

    def __call__(self, vocab_size, embed_dim):
        emb = self._make(vocab_size, embed_dim)
        torch.nn.init.xavier_normal_(emb.weight, self.gain)
        return emb


class XavierUniform(_Init):
    def __init__(self, gain=1.0):
        self.gain = gain

    def __call__(self, vocab_size, embed_dim):
        emb = self._make(vocab_size, embed_dim)
        torch.nn.init.xavier_uniform_(emb.weight, self.gain)
        return emb


class Pretrained(_Init):
    """Embedding from a given 2-D weight; ``freeze=True`` keeps it out of training."""

    def __init__(self, embedding_weight, freeze=True):
        self.embedding_weight = torch.as_tensor(embedding_weight, dtype=torch.float32)
        self.freeze = freeze

    def __call__(self, vocab_size, embed_dim):
        if (vocab_size, embed_dim) != tuple(self.embedding_weight.shape):
            raise AssertionError("pretrained weight shape does not match (vocab_size, embed_dim)")
        return torch.nn.Embedding.from_pretrained(self.embedding_weight, freeze=self.freeze)
