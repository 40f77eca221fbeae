#!/usr/bin/env python
"""Benchmark of the X-LXMERT hot path on B200 (contract: DESIGN.md §Measurement).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path

Headline (`value`, `e2e`): the FULL pre-training step of BASELINE.json configs[2] — the reference's
`lxmert_pretrain.py:295-364` loop body: task round-robin vis_mask / word_mask / matched (`MASK_MODALITY[step % 3]`),
XLxmertForPretraining forward (centroid gather, embeddings, 9L/5R/5X encoder, pooler, task head + loss), backward,
data-parallel gradient exchange (N > 1), global-norm clip + AdamW — batch 256 per GPU, 20 text tokens, 8x8x2048 grid,
synthetic data, random-init weights.  `value`: inputs resident in HBM.  `e2e`: the same step fed from a HOST batch dict
through B200PretrainInputs (one packed pinned H2D per step, prefetched one step ahead) with the loss read back.
Sub-records of the same JSON line (N = 1): `encoder_only` (configs[1]), `generator_b128` (configs[3]),
`sampler_nar4_b32` (configs[4]), `gpu_eager_reference` (the reference's PyTorch modules run eagerly on the same GPU).
Prints ONE JSON line on stdout.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pretrain_samples_per_sec"
UNIT = "samples/s"
BATCH = 256
L_TOK, V_GRID = 20, 64
TASKS = ("vis_mask", "word_mask", "matched")            # MASK_MODALITY order, lxmert_pretrain.py:794-800
ENC_GFLOP_FWD_BWD = 46.17                               # per sample, SURVEY.md §8(d) C2 (algorithmic, 2·M·N·K)
STEP_GFLOP = {"vis_mask": 54.9, "word_mask": 49.1, "matched": 47.2}   # per sample fwd+bwd, SURVEY.md §8(d) C3
GEN_GFLOP = 27.755                                      # per image, SURVEY.md §8(a) a18
NAR4_GFLOP = 100.9                                      # per image, 4 steps + decode, SURVEY.md §8(d) C5
PREWARM_STEPS = 12     # untimed steps before the W warm-up steps (allocator + clocks steady state)
CPU_SAMPLE_B = 32      # bounded sample per CPU step: large enough that the fixed optimiser cost (≈ 0.5 s) does not dominate
# measured on this pool's B200s by the driver (BASELINE.md §2 keeps a copy of MEASURED_PEAKS.json)
PEAKS_COPY = {"hbm_gbs": 6532.9, "bf16_tflops": 1627.7, "bf16_tflops_sustained": 1358.9}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            out = {k: float(d[k]) for k in PEAKS_COPY if k in d}
            if len(out) == len(PEAKS_COPY):
                return out, "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return dict(PEAKS_COPY), "measured (BASELINE.md copy of MEASURED_PEAKS.json)"


def committed_profile(name):
    """A small JSON record under profiles/ produced from an ncu capture of this round (never measured live here)."""
    path = os.path.join(ROOT, "profiles", name)
    try:
        return json.load(open(path))
    except Exception:
        return None


def workload_config(n_gpus, B=BATCH, extra=None):
    c = {"workload": "full X-LXMERT pre-training step (BASELINE.json configs[2]): task round-robin vis_mask/word_mask/"
                     "matched, XLxmertForPretraining fwd (cluster-id gather, embeddings, 9L/5R/5X encoder, pooler, task "
                     "head + loss) + bwd + gradient all-reduce (N>1) + global-norm clip + AdamW; batch 256 per GPU, 20 "
                     "text tokens + 8x8x2048 grid",
         "batch_per_gpu": B, "global_batch": B * n_gpus, "text_tokens": L_TOK, "grid_cells": V_GRID,
         "parallelism": f"dp{n_gpus}", "tasks": list(TASKS),
         "l2_policy": "working set (~14 GB of saved activations + 2.5 GB of weights/moments per step) is far larger "
                      "than the 126 MB L2"}
    if extra:
        c.update(extra)
    return c


# ---------------------------------------------------------------------------------------------------
# The reference's own PyTorch path for the step (used by the CPU legs and, on the GPU, by `gpu_eager_reference`):
# HF LxmertModel is the class x-lxmert/src/lxrt/modeling.py:5,80 instantiates as self.bert; the two heads are the
# torch.nn modules of lxrt/modeling.py:8-53 (LxmertVisualObjHead, cluster mode) and HF:583-665 (LM + matched head);
# the optimiser is AdamW(eps 1e-6) + clip_grad_norm_(1.0) (lxmert_pretrain.py:110-141,343-364).  /root/reference does
# not exist on the GPU box, so the wrapper class itself cannot be imported there — these are its building blocks.
# ---------------------------------------------------------------------------------------------------

def build_torch_reference_step(B, device, seed=0):
    import torch
    from torch import nn
    from transformers import LxmertConfig, LxmertModel
    from xlxmert_b200 import synth
    from xlxmert_b200.config import DEFAULT_DIMS as D
    torch.manual_seed(seed)
    cfg = LxmertConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    bert = LxmertModel(cfg)
    H, F, C, Vc = D.hidden, D.feat_dim, D.num_clusters, D.vocab

    class Transform(nn.Module):                          # LxmertPredictionHeadTransform, HF:583-594
        def __init__(self):
            super().__init__()
            self.dense, self.act, self.LayerNorm = nn.Linear(H, H), nn.GELU(), nn.LayerNorm(H, eps=1e-12)

        def forward(self, x):
            return self.LayerNorm(self.act(self.dense(x)))

    class Ref(nn.Module):
        def __init__(self):
            super().__init__()
            self.bert = bert
            self.obj_transform, self.linear_feat, self.out_cluster = Transform(), nn.Linear(H, F), nn.Linear(F, C)
            self.lm_transform = Transform()
            self.decoder = nn.Linear(H, Vc, bias=False)
            self.decoder.weight = bert.embeddings.word_embeddings.weight        # tied, lxrt/modeling.py:86
            self.lm_bias = nn.Parameter(torch.zeros(Vc))
            self.seq_relationship = nn.Linear(H, 2)
            self.mask_feat = nn.Parameter(torch.zeros(F))
            self.vis_emb = nn.Embedding.from_pretrained(synth.centroid_table(D), freeze=True)
            self.out_cluster.weight = self.vis_emb.weight                       # frozen centroid table, :146-151
            self.ce = nn.CrossEntropyLoss()

        def forward(self, b, task):
            feats = self.vis_emb(b["cluster_ids"])
            if task == "vis_mask":
                feats = torch.where(b["vis_mask"].unsqueeze(-1), self.mask_feat.view(1, 1, -1), feats)
            ids = b["masked_input_ids"] if task == "word_mask" else b["input_ids"]
            out = self.bert(input_ids=ids, visual_feats=feats, visual_pos=b["visual_pos"], attention_mask=ids > 0,
                            return_dict=True)
            if task == "vis_mask":
                logits = self.out_cluster(self.linear_feat(self.obj_transform(out.vision_output)))
                return self.ce(logits.view(-1, C), b["obj_labels"].view(-1))
            if task == "word_mask":
                scores = self.decoder(self.lm_transform(out.language_output)) + self.lm_bias
                return self.ce(scores.view(-1, Vc), b["word_labels"].view(-1))
            return self.ce(self.seq_relationship(out.pooled_output).view(-1, 2), b["matched_labels"].view(-1))

    model = Ref().to(device).train()
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, eps=1e-6, weight_decay=0.01)
    batch = {k: v.to(device) for k, v in synth.make_batch(D, B, L_TOK, V_GRID, seed=seed).items()}
    counter = [0]

    def step(autocast_dtype=None):
        task = TASKS[counter[0] % 3]
        counter[0] += 1
        for p in params:
            p.grad = None
        if autocast_dtype is not None:
            with torch.autocast(device_type=torch.device(device).type, dtype=autocast_dtype):
                loss = model(batch, task)
        else:
            loss = model(batch, task)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        return loss
    return step


def time_cpu(B: int, steps: int, warmup: int):
    """The reference's step on the host cores (all threads), bounded sample of B samples per step."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    step = build_torch_reference_step(B, "cpu")
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return dict(value=B / dt, unit=UNIT, cores=torch.get_num_threads(), kind="reference",
                sample=f"full pre-training step (fwd + bwd + clip + AdamW, tasks round-robin) on B={B} samples/step of "
                       f"the {BATCH}-sample batch, {steps} timed steps, fp32; HF LxmertModel + the torch.nn heads of "
                       "lxrt/modeling.py"), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    cb, dt = time_cpu(CPU_SAMPLE_B, max(1, args.steps), max(1, min(args.warmup, 3)))
    return {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus, extra={
                "reference_step": f"bounded sample, B={CPU_SAMPLE_B} per step on host cores"}),
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------

def build_pretraining_model(dev, passes, seed=0):
    """B200XLxmertForPretraining with seeded random-init weights of the reference architecture + the centroid table."""
    import torch
    from xlxmert_b200 import params as P, synth
    from xlxmert_b200.config import DEFAULT_DIMS as D
    from xlxmert_b200.pretraining import B200XLxmertForPretraining
    d = D
    cls_specs = [("predictions.transform.dense.weight", (d.hidden, d.hidden)),
                 ("predictions.transform.dense.bias", (d.hidden,)),
                 ("predictions.transform.LayerNorm.weight", (d.hidden,)),
                 ("predictions.transform.LayerNorm.bias", (d.hidden,)),
                 ("predictions.bias", (d.vocab,)),
                 ("seq_relationship.weight", (2, d.hidden)), ("seq_relationship.bias", (2,))]
    full = {"bert." + k: v for k, v in P.init_state_dict(P.model_param_specs(d), seed=seed).items()}
    full.update({"obj_predict_head." + k: v for k, v in P.init_state_dict(P.objhead_param_specs(d), seed=seed + 1).items()
                 if k != "out_cluster.weight"})
    full.update({"cls." + k: v for k, v in P.init_state_dict(cls_specs, seed=seed + 2).items()})
    table = synth.centroid_table(d)
    model = B200XLxmertForPretraining(d, num_clusters=d.num_clusters, passes=passes)
    model.set_visual_embedding(table.clone())
    missing, unexpected = model.load_state_dict(full, strict=False)
    assert not unexpected, unexpected
    assert all(any(s in k for s in ("vis_emb", "out_cluster.weight", "decoder.weight", "mask_feat")) for k in missing), missing
    model = model.to(dev).train()
    model.set_visual_embedding(table.to(dev))
    return model, table.to(dev)


def run_ours(args):
    import ctypes as C
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    from xlxmert_b200 import _lib, synth
    from xlxmert_b200.config import DEFAULT_DIMS as D
    from xlxmert_b200.inputs import B200PretrainInputs
    from xlxmert_b200.optim import B200AdamW, lxmert_param_groups
    from xlxmert_b200.parallel import allreduce_gradients, enable_overlapped_gradient_sync
    lib = _lib.load()

    B = args.batch
    passes = args.passes
    model, table = build_pretraining_model(dev, passes)
    enc = model.bert.encoder
    params = [p for p in model.parameters() if p.requires_grad]
    opt = B200AdamW(lxmert_param_groups(model, 0.01), lr=1e-4)        # lxmert_pretrain.py:110-141 (eps 1e-6, two groups)
    if world > 1 and not args.no_overlap:
        # stage-wise all-reduce inside the backward; XLX_DEFER_SYNC_WAIT=1 also lets the rest of the backward overlap the
        # last range's reduce (opt-in until measured on several GPUs)
        enable_overlapped_gradient_sync(model, defer_wait=bool(int(os.environ.get("XLX_DEFER_SYNC_WAIT", "0"))))

    # ---- the step's inputs, as collate_fn hands them over (lxmert_data.py:497-652), one batch per task slot
    host = synth.make_batch(D, B, L_TOK, V_GRID, seed=rank)
    g = torch.Generator().manual_seed(1000 + rank)
    host_batch = dict(word_id=host["input_ids"], masked_word_id=host["masked_input_ids"],
                      other_word_id=host["input_ids"][torch.randperm(B, generator=g)], word_label=host["word_labels"],
                      box_position=host["visual_pos"], vis_mask=host["vis_mask"], matched_label=host["matched_labels"],
                      cluster_id=host["cluster_ids"])
    host_batch = {k: v.contiguous().pin_memory() for k, v in host_batch.items()}

    # device-resident keyword arguments per task (for `value`)
    stager = B200PretrainInputs(dev, depth=3)
    resident = {}
    for task in TASKS:
        stager.stage(host_batch, task)
    for task in TASKS:
        resident[task] = stager.kwargs()
        stager.done()
    torch.cuda.synchronize()
    sync = world > 1
    counter = [0]

    def train_step(kw, do_sync=True):
        out = model(**kw)
        out["total_loss"].backward()
        if sync and do_sync:
            allreduce_gradients(model)           # encoder arena already reduced stage-wise; one flat buffer for the rest
        opt.step(max_grad_norm=1.0)              # fused global-norm clip + AdamW (lxmert_pretrain.py:343-364)
        for p in params:
            p.grad = None
        return out["total_loss"]

    def step_resident():
        task = TASKS[counter[0] % 3]
        counter[0] += 1
        train_step(resident[task])

    # ---- end to end: host batch dict in, loss on the host out, inputs prefetched one step ahead
    inputs = B200PretrainInputs(dev, depth=2)
    # the loss of every step is read on the host, one step late: step i enqueues its device→host copy and then waits
    # for step i−1's, so the host keeps one step of launches ahead of the device (a training loop's logging does not
    # need the loss before the next step is issued); the timed region ends with a full synchronize
    loss_h = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [None, None]
    losses_read = []
    e2e_marks = []
    e2e_state = {"i": 0, "primed": False}

    def step_e2e():
        i = e2e_state["i"]
        if not e2e_state["primed"]:
            inputs.stage(host_batch, TASKS[i % 3])
            e2e_state["primed"] = True
        kw = inputs.kwargs()
        inputs.stage(host_batch, TASKS[(i + 1) % 3])          # next step's H2D + unpack overlap this step's compute
        loss = train_step(kw)
        inputs.done()
        slot = i & 1
        loss_h[slot].copy_(loss.detach(), non_blocking=True)
        loss_ev[slot] = torch.cuda.Event()
        loss_ev[slot].record()
        prev = loss_ev[slot ^ 1]
        if prev is not None:
            prev.synchronize()                                # step i−1 is complete: its loss is on the host
            losses_read.append(float(loss_h[slot ^ 1]))
        e2e_state["i"] = i + 1
        e2e_marks.append(time.perf_counter())

    spread = {}

    def timed(fn, steps, warmup, tag=None):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.xlx_launch_count()
        gc.collect()
        gc.disable()          # a generation-2 collection inside a step that syncs with the host shows up as a 100 ms stall
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps - 1)] if tag else []
        try:
            e0.record()
            for i in range(steps):
                fn()
                if tag and i < steps - 1:
                    marks[i].record()          # per-step spread (diagnostic only; the metric uses e0 → e1)
            e1.record()
            torch.cuda.synchronize()
        finally:
            gc.enable()
        if tag and steps > 1:
            ev = [e0] + marks + [e1]
            per = sorted(a.elapsed_time(b) for a, b in zip(ev[:-1], ev[1:]))
            spread[tag] = {"min": round(per[0], 3), "median": round(per[len(per) // 2], 3), "max": round(per[-1], 3)}
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.xlx_launch_count() - n0
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, launches

    if args.ncu:       # profiler mode: W warm-up + K plain steps, nothing printed that looks like a bench value
        for _ in range(args.warmup + args.steps):
            step_resident()
        torch.cuda.synchronize()
        return {"ncu_mode": True, "launches_per_step": int(lib.xlx_launch_count()) // (args.warmup + args.steps)}

    warm = max(args.warmup, 3)
    # untimed pre-warm before the contract's W warm-up steps: the caching allocator reaches its steady state and the
    # part settles at its power-capped clocks (the first second of tensor work on an idle B200 runs at other clocks than
    # the rest); reported as config.prewarm_steps
    for _ in range(PREWARM_STEPS):
        step_resident()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    counter[0] = 0
    ms_step, launches = timed(step_resident, args.steps, warm, tag="resident")
    clocks = sampler.stop() if rank == 0 else None
    e2e_warm = 5
    ms_e2e, _ = timed(step_e2e, max(3, args.steps), e2e_warm)
    e2e_steps = [round((b - a) * 1e3, 2) for a, b in zip(e2e_marks[e2e_warm - 1:], e2e_marks[e2e_warm:])]

    # data-parallel correctness (N > 1), checked BEFORE the exchange is switched off below: after all these steps every
    # replica must hold bit-identical weights (same initial weights, same reduced gradients, deterministic optimiser)
    in_sync = None
    if world > 1:
        try:
            chk = torch.stack([p.detach().double().sum() for p in params]).sum().reshape(1)
            both = torch.cat([chk, -chk])                 # one MAX reduction gives max and −min
            dist.all_reduce(both, op=dist.ReduceOp.MAX)
            in_sync = bool((both[0] == -both[1]).item())
        except Exception as ex:      # a diagnostic never takes the bench line down
            in_sync = f"check failed: {ex!r}"

    # exposed gradient-exchange time (N > 1): the same step with the exchange switched off
    exposed = None
    if world > 1:
        groups = [(e, e.grad_sync_group) for e in [enc]]
        for e, _ in groups:
            e.grad_sync_group = None
        counter[0] = 0
        ms_nosync, _ = timed(lambda: (train_step(resident[TASKS[counter[0] % 3]], do_sync=False),
                                      counter.__setitem__(0, counter[0] + 1)), args.steps, 2)
        for e, gsg in groups:
            e.grad_sync_group = gsg
        exposed = {"ms_per_step_without_exchange": ms_nosync, "exposed_exchange_ms": ms_step - ms_nosync,
                   "note": "same step, same ranks, gradient all-reduce switched off (weights diverge; timing only)"}

    # ---- roofline of the dominant kernel (tcgen05 GEMM): events around every launch, outside the timed region,
    #      one step per task
    lib.xlx_profile_gemm_begin()
    counter[0] = 0
    for _ in range(3):
        step_resident()
    tot_ms, tot_fl, n = C.c_double(), C.c_double(), C.c_int64()
    _lib.check("xlx_profile_gemm_end", lib.xlx_profile_gemm_end(C.byref(tot_ms), C.byref(tot_fl), C.byref(n)))
    peaks, peak_src = measured_peaks()
    achieved = tot_fl.value / (tot_ms.value * 1e-3) / 1e12 if tot_ms.value > 0 else 0.0
    peak = peaks["bf16_tflops_sustained"]
    traffic = committed_profile("r02_gemm_dram_traffic.json") if (passes == 3 and B == BATCH) else None
    roofline = {"bound": "tensor", "kernel": "xlx::gemm_kernel<32,…> (tcgen05 bf16x3 GEMM: every Linear fwd/dgrad/wgrad "
                                             "of encoder and heads)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": None if not traffic else traffic.get("mean_dram_bytes_per_launch"),
                "traffic_source": None if not traffic else traffic.get("source"),
                "peak_source": peak_src + ", sustained bf16 (kernel timed inside a long step)",
                "launches_per_step": n.value / 3.0, "avg_launch_us": tot_ms.value * 1e3 / max(n.value, 1),
                "gemm_share_of_step": (tot_ms.value / 3) / ms_step,
                "executed_tensor_tflops": achieved * (3 if passes == 3 else 1),
                "note": "achieved counts ALGORITHMIC FLOPs (2MNK) of the launches actually made (heads run on labelled "
                        "rows only); bf16x3 executes 3 MMAs per algorithmic MAC, so frac is capped at 1/3 by "
                        "construction in the fp32-parity mode. The per-launch timing pass keeps every kernel on one "
                        "stream (events between launches), so these durations exclude the two-stream / dependent-launch "
                        "overlap the timed step enjoys"}
    xattn = committed_profile("r02_cross_attention_ncu.json")

    line = None
    if rank == 0:
        mean_gflop = sum(STEP_GFLOP.values()) / 3
        value = B * world / (ms_step * 1e-3)
        e2e_value = B * world / (ms_e2e * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None,
                "dtype": "bf16x3 (split-bf16 tensor-core GEMMs, fp32 accumulate; fp32 elsewhere)" if passes == 3 else "bf16",
                "data": "synthetic",
                "config": workload_config(world, B, extra={"passes": passes, "prewarm_steps": PREWARM_STEPS,
                                                           "step_tflop_algorithmic": mean_gflop * B / 1e3}),
                "step_ms_spread": spread.get("resident"),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": inputs.h2d_bytes, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e, "host_ms_each_step": e2e_steps,
                        "api": "B200PretrainInputs.stage(host batch dict, task) -> B200XLxmertForPretraining(**kwargs) "
                               "-> total_loss.backward() -> allreduce_gradients -> B200AdamW.step(max_grad_norm=1); "
                               "pinned host batch in (one packed copy, prefetched one step ahead), every step's loss "
                               "scalar read on the host (one step late)",
                        "losses_read_on_host": len(losses_read), "last_loss": losses_read[-1] if losses_read else None},
                "gpu_launches": int(launches),
                "clocks": clocks, "roofline": roofline,
                "step_algorithmic_tflops": mean_gflop * B / 1e3 / (ms_step * 1e-3)}
        if xattn:
            line["cross_attention"] = xattn
        if exposed:
            line["gradient_exchange"] = exposed
    if world > 1:
        if rank == 0:
            line["replicas_in_sync"] = in_sync
        dist.barrier()

    # ---- sub-records (single GPU only: they describe one device)
    if world == 1 and rank == 0 and not args.no_extra:
        del stager, inputs
        extra = {}
        try:
            extra["encoder_only"] = bench_encoder_only(model, table, dev, B, passes, lib)
        except Exception as ex:      # a sub-record never takes the headline down
            extra["encoder_only"] = {"error": repr(ex)}
        try:
            extra["train_mode_dropout_0.1"] = bench_dropout_step(model, step_resident, counter, B)
        except Exception as ex:
            extra["train_mode_dropout_0.1"] = {"error": repr(ex)}
        del model, opt, params, resident
        torch.cuda.empty_cache()
        for name, fn in (("generator_b128", bench_generator), ("sampler_nar4_b32", bench_sampler),
                         ("kmeans_assign", bench_kmeans), ("gpu_eager_reference", bench_gpu_eager)):
            try:
                extra[name] = fn(dev, peaks)
            except Exception as ex:
                extra[name] = {"error": repr(ex)}
            torch.cuda.empty_cache()
        ge = extra.get("gpu_eager_reference", {})
        if "fp32" in ge:
            ge["ours_over_eager_fp32"] = value / ge["fp32"]["samples_per_s"]
            ge["ours_over_eager_tf32"] = value / ge["tf32"]["samples_per_s"]
            ge["ours_over_eager_autocast_bf16"] = value / ge["autocast_bf16"]["samples_per_s"]
        line.update(extra)
    if world == 1 and rank == 0 and not args.no_cpu:
        cb, _ = time_cpu(CPU_SAMPLE_B, 3, 1)
        line["cpu_baseline"] = cb
    if world > 1:
        dist.destroy_process_group()
    return line


def _timed_simple(fn, iters, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_dropout_step(model, step_resident, counter, B):
    """The same resident step with the reference's training-mode dropout (nn.Dropout(0.1) at every HF site, generated
    inside the kernels from a counter-based Philox stream).  The headline and both reference arms run with p = 0 so that
    they compute the same numbers; this record says what the dropout the reference trains with costs here."""
    from dataclasses import replace
    mods = [m for m in model.modules() if hasattr(m, "dims") and hasattr(m.dims, "hidden_dropout")]
    old = [m.dims for m in mods]
    try:
        for m in mods:
            m.dims = replace(m.dims, hidden_dropout=0.1, attention_dropout=0.1)
        counter[0] = 0
        ms = _timed_simple(step_resident, 6, 3)
    finally:
        for m, d in zip(mods, old):
            m.dims = d
    return {"workload": "full pre-training step, B=256, hidden_dropout = attention_dropout = 0.1 (train mode)",
            "ms_per_step": ms, "samples_per_s": B / ms * 1e3}


def bench_encoder_only(model, table, dev, B, passes, lib):
    """BASELINE.json configs[1]: 9L/5R/5X encoder fwd+bwd alone (last round's headline), device-resident inputs."""
    import torch
    from xlxmert_b200 import synth
    from xlxmert_b200.config import DEFAULT_DIMS as D
    enc = model.bert.encoder
    batch = synth.make_batch(D, B, L_TOK, V_GRID, seed=0)
    g = torch.Generator().manual_seed(100)
    with torch.no_grad():
        emb_d = model.bert.embeddings(batch["input_ids"].to(dev)).detach()
        feats_d = table[batch["cluster_ids"].to(dev)].contiguous()
    pos_d = batch["visual_pos"].to(dev)
    ext = ((1.0 - batch["attention_mask"].to(dev)[:, None, None, :].float()) * torch.finfo(torch.float32).min).contiguous()
    g_lang = (torch.randn(B, L_TOK, D.hidden, generator=g) / (B * L_TOK)).to(dev)
    g_vis = (torch.randn(B, V_GRID, D.hidden, generator=g) / (B * V_GRID)).to(dev)

    def step():
        emb = emb_d.requires_grad_(True)
        emb.grad = None
        (vs, _), (ls, _), _ = enc(emb, ext, feats_d, pos_d)
        torch.autograd.backward([ls[-1], vs[-1]], [g_lang, g_vis])
        for p in enc.parameters():
            p.grad = None
    ms = _timed_simple(step, 10, 3)
    return {"workload": "BASELINE.json configs[1]: encoder fwd+bwd, B=256, weights re-split every step",
            "ms_per_step": ms, "samples_per_s": B / ms * 1e3, "algorithmic_tflops": ENC_GFLOP_FWD_BWD * B / ms,
            "passes": passes}


def bench_generator(dev, peaks):
    """BASELINE.json configs[3]: Generator forward, 64x2048 grid -> 256x256 RGB, batch 128."""
    import torch
    from xlxmert_b200 import params as P, synth
    from xlxmert_b200.config import DEFAULT_DIMS as D
    from xlxmert_b200.generator import B200Generator
    G = B200Generator()
    G.load_state_dict(P.init_generator_state_dict(seed=0), strict=True)
    G = G.to(dev).eval()
    Bg = 128
    ids = torch.randint(0, D.num_clusters, (Bg, 64), device=dev)
    code = synth.centroid_table(D).to(dev)[ids]
    ms = _timed_simple(lambda: G(code.view(Bg, 8, 8, 2048), train=False), 8, 3)
    tf = GEN_GFLOP * Bg / ms
    return {"workload": "BASELINE.json configs[3]: Generator.forward B=128 (eval, noise off), device-resident codes",
            "ms": ms, "images_per_s": Bg / ms * 1e3, "algorithmic_tflops": tf,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": tf / peaks["bf16_tflops_sustained"],
                         "note": "whole forward (convolutions + glue kernels); bf16x3 caps the convolutions at 1/3"}}


def bench_sampler(dev, peaks):
    """BASELINE.json configs[4]: text -> image, 4-step NAR mask-predict + GAN decode, batch 32, images land on the host."""
    import torch
    from xlxmert_b200 import params as P, synth
    from xlxmert_b200.config import DEFAULT_DIMS as D
    from xlxmert_b200.generator import B200Generator
    from xlxmert_b200.sampler import B200ImggenModel
    pre, table = build_pretraining_model(dev, 3)
    m = B200ImggenModel(D, num_clusters=D.num_clusters)
    m.set_visual_embedding(table.clone())
    m.load_state_dict({k: v for k, v in pre.state_dict().items() if not k.startswith("cls.")}, strict=False)
    del pre
    G = B200Generator()
    G.load_state_dict(P.init_generator_state_dict(seed=0), strict=True)
    m.set_image_generator(G)
    m = m.to(dev)
    m.set_visual_embedding(table)
    Bs = 32
    tok = synth.make_batch(D, Bs, L_TOK, V_GRID, seed=3)["input_ids"].to(dev)
    ms = _timed_simple(lambda: m.sample_image_NAR(tok, n_steps=4), 8, 3)
    tf = NAR4_GFLOP * Bs / ms
    return {"workload": "BASELINE.json configs[4]: sample_image_NAR(n_steps=4) + generator decode, B=32, token ids on the "
                        "device in, images on the host out",
            "ms": ms, "images_per_s": Bs / ms * 1e3, "algorithmic_tflops_reference_work": tf,
            "note": "FLOPs counted as the reference does the work (language layers recomputed every step)"}


def bench_kmeans(dev, peaks):
    """nearest-centroid assignment of grid features (run_kmeans.py:124-143, faiss IndexFlatL2.search
    with k = 1): 10 000 centroids x 2048, features resident on the device and streamed from host memory."""
    import numpy as np
    import torch
    from xlxmert_b200 import synth
    from xlxmert_b200.config import DEFAULT_DIMS as D
    from xlxmert_b200.kmeans import B200IndexFlatL2
    cent = synth.centroid_table(D).float()
    index = B200IndexFlatL2(D.feat_dim)
    index.add(cent)
    n = 32768
    g = torch.Generator().manual_seed(0)
    x = (cent[torch.randint(0, cent.shape[0], (n,), generator=g)] + 0.3 * torch.randn(n, D.feat_dim, generator=g))
    xd = x.to(dev)
    ms = _timed_simple(lambda: index.search(xd, 1), 5, 2)
    xh = np.ascontiguousarray(np.concatenate([x.numpy()] * 2))        # 65 536 rows = 512 MiB of fp32 features on the host
    index.search(xh[:n], 1)
    t0 = time.perf_counter()
    index.search(xh, 1)
    dt = time.perf_counter() - t0
    return {"workload": "B200IndexFlatL2.search(x, 1): 10 000 centroids x 2048 (run_kmeans.py:124-143)",
            "device_resident": {"rows": n, "ms": ms, "rows_per_s": n / ms * 1e3,
                                "algorithmic_tflops": 2.0 * n * D.feat_dim * cent.shape[0] / ms / 1e9},
            "host_streamed": {"rows": int(xh.shape[0]), "ms": dt * 1e3, "rows_per_s": xh.shape[0] / dt,
                              "h2d_GBps": xh.nbytes / dt / 1e9}}


def bench_gpu_eager(dev, peaks):
    """north_star's comparison: the reference's PyTorch modules run eagerly on the SAME GPU for the SAME step."""
    import torch
    out = {"workload": "same full pre-training step, B=256: HF LxmertModel + torch.nn heads + clip_grad_norm_ + "
                       "torch.optim.AdamW, eager PyTorch on this GPU (cuBLAS/cuDNN), 6 timed steps after 3 warm-up"}
    old = torch.backends.cuda.matmul.allow_tf32
    try:
        step = build_torch_reference_step(BATCH, dev)
        torch.backends.cuda.matmul.allow_tf32 = False
        ms = _timed_simple(step, 6, 3)
        out["fp32"] = {"ms_per_step": ms, "samples_per_s": BATCH / ms * 1e3,
                       "note": "true fp32 matmuls: the precision class this repo's default mode matches"}
        torch.backends.cuda.matmul.allow_tf32 = True
        ms = _timed_simple(step, 6, 3)
        out["tf32"] = {"ms_per_step": ms, "samples_per_s": BATCH / ms * 1e3, "note": "outside the 1e-3 parity bar"}
        ms = _timed_simple(lambda: step(torch.bfloat16), 6, 3)
        out["autocast_bf16"] = {"ms_per_step": ms, "samples_per_s": BATCH / ms * 1e3,
                                "note": "what pretrain.bash's mixed-precision flag asks for; outside the parity bar"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--passes", type=int, default=3, choices=[1, 3])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the sub-records (encoder only, generator, sampler, GPU eager)")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: one all-reduce after the backward instead of stage-wise")
    ap.add_argument("--ncu", action="store_true", help="profiler mode: run warmup+steps resident steps and exit")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: everything libraries print while we run (build messages, NCCL's
    # version banner, …) is routed to stderr at the file-descriptor level
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        line = run_reference(args) if args.impl == "reference" else run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
