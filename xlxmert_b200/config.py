"""Static dimensions of the X-LXMERT hot path.

Values follow the reference's canonical run: HF ``LxmertConfig`` defaults
(``transformers/models/lxmert/configuration_lxmert.py:85-87`` → 9 language / 5 relational /
5 cross-modality layers, hidden 768, 12 heads, intermediate 3072), 2048-d grid features
(``x-lxmert/src/param.py:138-140``: 8×8 grid = 64 cells, ≤ 20 text tokens) and 10 000 visual
clusters (``x-lxmert/src/param.py:167``).
"""
from dataclasses import dataclass, asdict


@dataclass(frozen=True)
class LxmertDims:
    hidden: int = 768
    heads: int = 12
    intermediate: int = 3072
    feat_dim: int = 2048
    pos_dim: int = 4
    l_layers: int = 9
    r_layers: int = 5
    x_layers: int = 5
    vocab: int = 30522
    max_pos: int = 512
    type_vocab: int = 2
    num_clusters: int = 10000
    ln_eps: float = 1e-12
    #: nn.Dropout probabilities of a TRAINING-mode forward (HF config hidden_dropout_prob / attention_probs_dropout_prob,
    #: both 0.1 in the reference's run — ``dims_from_hf_config`` copies them).  Dims built by hand default to 0: the
    #: parity tests and the bench compare against the reference at p = 0.
    hidden_dropout: float = 0.0
    attention_dropout: float = 0.0

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    def asdict(self):
        return asdict(self)


DEFAULT_DIMS = LxmertDims()

#: a tiny configuration with the same structure, used by fast CPU tests of host logic
TINY_DIMS = LxmertDims(hidden=128, heads=2, intermediate=256, feat_dim=192, l_layers=2, r_layers=1,
                       x_layers=2, vocab=512, max_pos=64, num_clusters=320)
