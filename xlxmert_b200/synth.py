"""Seeded synthetic inputs for the hot path (SURVEY.md §8d "Synthetic inputs").

Shapes and value ranges mirror what the reference's data layer hands to the model
(``x-lxmert/src/pretrain/lxmert_data.py:497-652`` → ``lxmert_pretrain.py:143-225``): token ids
``[B, L]`` int64 with a PAD tail, cluster ids ``[B, 64]`` int64 into a ``[C, 2048]`` centroid table,
``box_position(8)`` grid boxes and a per-sample visual mask with 1…64 masked cells
(``lxmert_data.py:414-419``).
"""
from __future__ import annotations

import numpy as np
import torch

from .config import LxmertDims


def box_position(grid_size: int = 8) -> np.ndarray:
    """Normalised (x0, y0, x1, y1) of each cell of a ``grid_size``² grid, row-major.

    Same values as the reference helper ``x-lxmert/src/utils.py:75-85`` (cell (i, j) →
    ``(j/g, i/g, (j+1)/g, (i+1)/g)``, float32), computed without the double loop.
    """
    g = grid_size
    ii, jj = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    boxes = np.stack([jj / g, ii / g, (jj + 1) / g, (ii + 1) / g], axis=-1)
    return boxes.reshape(g * g, 4).astype(np.float32)


def centroid_table(d: LxmertDims, seed: int = 1234, scale: float = 0.05) -> torch.Tensor:
    """Non-negative centroid table (post-ReLU fc6-like features), ``scale·|N(0,1)|``."""
    g = torch.Generator().manual_seed(seed)
    return scale * torch.randn(d.num_clusters, d.feat_dim, generator=g).abs()


def make_batch(d: LxmertDims, B: int, L: int = 20, V: int = 64, seed: int = 0,
               pad_half: bool = True) -> dict:
    """One synthetic pre-training batch (CPU tensors)."""
    g = torch.Generator().manual_seed(seed)
    lo = min(1000, d.vocab // 2)
    hi = min(30000, d.vocab)
    ids = torch.randint(lo, hi, (B, L), generator=g, dtype=torch.int64)
    cls_id, sep_id = min(101, d.vocab - 2), min(102, d.vocab - 1)
    ids[:, 0] = cls_id
    n_tok = torch.full((B,), L, dtype=torch.int64)
    if pad_half and L >= 8:
        r = torch.randint(min(8, L), L + 1, (B,), generator=g)
        half = torch.arange(B) % 2 == 1
        n_tok = torch.where(half, r, n_tok)
    pos = torch.arange(L).unsqueeze(0)
    ids = torch.where(pos < n_tok.unsqueeze(1), ids, torch.zeros_like(ids))
    ids[torch.arange(B), n_tok - 1] = sep_id
    attention_mask = ids > 0

    cluster_ids = torch.randint(0, d.num_clusters, (B, V), generator=g, dtype=torch.int64)
    grid = int(round(V ** 0.5))
    if grid * grid == V:
        vpos = torch.from_numpy(box_position(grid))
    else:  # 36-box mode: arbitrary normalised boxes
        vpos = torch.rand(V, 4, generator=g)
    visual_pos = vpos.unsqueeze(0).expand(B, -1, -1).contiguous()

    # visual mask: n ~ U{1..V} cells per sample without replacement (lxmert_data.py:414-419)
    n_mask = torch.randint(1, V + 1, (B,), generator=g)
    scores = torch.rand(B, V, generator=g)
    rank = scores.argsort(dim=1).argsort(dim=1)
    vis_mask = rank < n_mask.unsqueeze(1)
    obj_labels = torch.where(vis_mask, cluster_ids, torch.full_like(cluster_ids, -100))

    # MLM labels: 15 % of the interior tokens (ignore value -100)
    interior = attention_mask.clone()
    interior[:, 0] = False
    interior[torch.arange(B), n_tok - 1] = False
    pick = (torch.rand(B, L, generator=g) < 0.15) & interior
    word_labels = torch.where(pick, ids, torch.full_like(ids, -100))
    masked_ids = torch.where(pick, torch.full_like(ids, min(103, d.vocab - 1)), ids)
    matched_labels = torch.randint(0, 2, (B,), generator=g, dtype=torch.int64)

    return dict(input_ids=ids, masked_input_ids=masked_ids, attention_mask=attention_mask,
                token_type_ids=torch.zeros_like(ids), cluster_ids=cluster_ids,
                visual_pos=visual_pos, vis_mask=vis_mask, obj_labels=obj_labels,
                word_labels=word_labels, matched_labels=matched_labels)


def visual_feats_from(table: torch.Tensor, cluster_ids: torch.Tensor) -> torch.Tensor:
    """``vis_emb(cluster_ids)`` (x-lxmert/src/lxrt/modeling.py:185-186)."""
    return table[cluster_ids]


def feat_qa_targets(d: LxmertDims, B: int, seed: int, num_answers: int, V: int = 64):
    """Synthetic ``feat_labels`` [B, V, feat_dim] (the grid features the ``feat`` loss regresses to,
    lxmert_pretrain.py:177-179) and ``qa_labels`` [B] (lxmert_pretrain.py:184-189); unit-variance targets put the
    SmoothL1 residuals on both sides of its |x| = 1 knee."""
    g = torch.Generator().manual_seed(seed + 77)
    feat_labels = torch.randn(B, V, d.feat_dim, generator=g)
    qa_labels = torch.randint(0, num_answers, (B,), generator=g)
    return feat_labels, qa_labels
