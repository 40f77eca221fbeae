"""Input path: ``B200PretrainInputs`` against a torch restatement of the reference's
``Trainer.forward`` argument building (x-lxmert/src/pretrain/lxmert_pretrain.py:143-225).  Integer / byte work:
every tensor must be bit-identical.  The CPU half checks the packed layout contract of the C ABI."""
import pytest
import torch

from xlxmert_b200 import synth
from xlxmert_b200.config import DEFAULT_DIMS as D


def collate_batch(B, L=20, V=64, seed=0):
    """A batch dict with the keys / dtypes ``collate_fn`` produces (lxmert_data.py:497-652)."""
    b = synth.make_batch(D, B, L, V, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    other = b["input_ids"][torch.randperm(B, generator=g)]
    return dict(word_id=b["input_ids"], masked_word_id=b["masked_input_ids"], other_word_id=other,
                word_label=b["word_labels"], box_position=b["visual_pos"], vis_mask=b["vis_mask"],
                matched_label=b["matched_labels"], cluster_id=b["cluster_ids"])


def reference_trainer_forward_args(batch, task, device):
    """lxmert_pretrain.py:143-225 restated line by line with torch ops (clustering mode, --visualLosses obj)."""
    cluster_ids = batch["cluster_id"].to(device)                           # :147-149
    visual_pos = batch["box_position"].to(device)                          # :153
    vis_mask = batch["vis_mask"].to(device).bool()                         # :155
    label_dict = {}
    if task == "word_mask":
        label_dict["word_labels"] = batch["word_label"].to(device)         # :158-159
    elif task == "vis_mask":
        obj_labels = cluster_ids.detach().clone()                          # :163 (clone: the reference aliases and
        obj_labels[~vis_mask] = -100                                       #  :165 clobbers cluster_ids — see note)
        label_dict["obj_labels"] = obj_labels
    elif task == "matched":
        label_dict["matched_labels"] = batch["matched_label"].to(device)   # :181-183
    key = {"word_mask": "masked_word_id", "matched": "other_word_id", "vis_mask": "word_id"}[task]   # :193-198
    word_id = batch[key].to(device)
    return dict(input_ids=word_id, visual_pos=visual_pos, attention_mask=word_id > 0, cluster_ids=cluster_ids,
                vis_mask=vis_mask, label_dict=label_dict, task=task)


def test_packed_layout_contract():
    from xlxmert_b200.inputs import packed_layout
    for B, L, V in [(1, 1, 1), (2, 20, 64), (256, 20, 64), (7, 13, 36)]:
        offs, total = packed_layout(B, L, V)
        sizes = [B * L * 8, B * L * 8, B * 8, B * V * 8, B * V, B * V * 16, B * 8]
        assert offs[0] == 0 and all(o % 256 == 0 for o in offs)
        for o, n, nxt in zip(offs, sizes, offs[1:] + [total]):
            assert o + n <= nxt                      # sections do not overlap
        assert len(offs) == 7 and total % 256 == 0 and total < sum(sizes) + 7 * 256
    with pytest.raises(Exception):
        packed_layout(0, 20, 64)


@pytest.mark.gpu
@pytest.mark.parametrize("B,L,V", [(2, 20, 64), (256, 20, 64), (5, 13, 36), (1, 1, 1)])
def test_unpacked_tensors_equal_reference_statements(B, L, V):
    from xlxmert_b200.inputs import B200PretrainInputs, TASKS
    dev = torch.device("cuda", 0)
    inputs = B200PretrainInputs(dev)
    for step, task in enumerate(TASKS * 2):          # six steps: every slot is reused, every task staged twice
        batch = collate_batch(B, L, V, seed=step)
        inputs.stage(batch, task)
        kw = inputs.kwargs()
        ref = reference_trainer_forward_args(batch, task, dev)
        torch.cuda.synchronize()
        for k in ("input_ids", "visual_pos", "attention_mask", "cluster_ids", "vis_mask"):
            assert kw[k].dtype == ref[k].dtype and kw[k].shape == ref[k].shape, (task, k)
            assert torch.equal(kw[k], ref[k]), (task, k)
        assert set(kw["label_dict"]) == set(ref["label_dict"])
        for k, v in ref["label_dict"].items():
            assert kw["label_dict"][k].dtype == v.dtype and torch.equal(kw["label_dict"][k], v), (task, k)
        add = kw["attention_mask"]._xlx_additive
        want = (1.0 - ref["attention_mask"].float()) * torch.finfo(torch.float32).min           # HF:766-774
        assert torch.equal(add, want)
        assert kw["task"] == task and kw["visual_feats"] is None and kw["visual_attention_mask"] is None
        inputs.done()
    from xlxmert_b200.inputs import packed_layout
    assert inputs.h2d_bytes == packed_layout(B, L, V)[1]        # ONE host→device copy of this many bytes per step


@pytest.mark.gpu
def test_staged_inputs_drive_the_model_like_separate_copies():
    """Same loss (bit-identical) whether the step's arguments come from B200PretrainInputs or from the reference's
    separate .to(device) statements; two batches staged ahead (double buffering) stay distinct."""
    from test_pretrain_parity import build_model
    from xlxmert_b200.inputs import B200PretrainInputs
    model, _ = build_model(0)
    model.eval()
    dev = torch.device("cuda", 0)
    inputs = B200PretrainInputs(dev, depth=2)
    b0, b1 = collate_batch(4, seed=10), collate_batch(4, seed=11)
    inputs.stage(b0, "vis_mask")
    inputs.stage(b1, "word_mask")
    with pytest.raises(RuntimeError):
        inputs.stage(b0, "matched")                  # both slots hold staged batches
    with torch.no_grad():
        for batch, task in ((b0, "vis_mask"), (b1, "word_mask")):
            got = model(**inputs.kwargs())["total_loss"]
            inputs.done()
            want = model(**reference_trainer_forward_args(batch, task, dev))["total_loss"]
            assert torch.equal(got, want), task


@pytest.mark.gpu
@pytest.mark.parametrize("pinned", [False, True])
def test_qa_and_feature_labels_equal_reference_statements(pinned):
    """--taskQA and --visualLosses obj,feat (lxmert_pretrain.py:177-189): ``qa_labels`` with the matched rule applied,
    ``feat_labels`` = the batch's grid features, bit-identical; the extra copy is counted in ``h2d_bytes``."""
    from xlxmert_b200.inputs import B200PretrainInputs, TASKS, packed_layout
    B, L, V, F = 5, 20, 64, D.feat_dim
    dev = torch.device("cuda", 0)
    inputs = B200PretrainInputs(dev, qa_labels=True, feat_labels=True)
    for step, task in enumerate(TASKS * 2):
        batch = collate_batch(B, L, V, seed=20 + step)
        g = torch.Generator().manual_seed(step)
        batch["qa_label"] = torch.randint(0, 9500, (B,), generator=g)
        batch["vis_feats"] = torch.randn(B, V, F, generator=g)
        if pinned:
            batch["vis_feats"] = batch["vis_feats"].pin_memory()
        inputs.stage(batch, task)
        kw = inputs.kwargs()
        torch.cuda.synchronize()
        qa = batch["qa_label"].clone().to(dev)                             # :185
        if task == "matched":
            qa.masked_fill_(batch["matched_label"].to(dev) == 0, -100)     # :186-188
        assert kw["label_dict"]["qa_labels"].dtype == torch.int64 and torch.equal(kw["label_dict"]["qa_labels"], qa)
        if task == "vis_mask":
            assert torch.equal(kw["label_dict"]["feat_labels"], batch["vis_feats"].to(dev))     # :177-179
            assert inputs.h2d_bytes == packed_layout(B, L, V)[1] + B * V * F * 4
        else:
            assert "feat_labels" not in kw["label_dict"]
            assert inputs.h2d_bytes == packed_layout(B, L, V)[1]
        ref = reference_trainer_forward_args(batch, task, dev)
        for k, v in ref["label_dict"].items():
            assert torch.equal(kw["label_dict"][k], v), (task, k)
        inputs.done()
