#!/usr/bin/env python
"""Benchmark of the X-LXMERT hot path on B200 (see the contract in DESIGN.md §Measurement).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (HF LxmertEncoder)

Workload (BASELINE.json configs[1]): 9L/5R/5X LXMERT encoder forward + backward, batch 256 per GPU, 20 text
tokens + 8×8 grid of 2048-d features, synthetic data, random-init weights.  One "step" = one forward + backward
over one batch (plus, for N > 1, one NCCL all-reduce of the flat gradient arena).  Prints ONE JSON line.
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
ENC_GFLOP_FWD = 15.389          # per sample, SURVEY.md §8(a) a11 (algorithmic, 2·M·N·K)
ENC_GFLOP_FWD_BWD = 46.17
# measured on this pool's B200s by the driver (BASELINE.md §2 keeps a copy of MEASURED_PEAKS.json)
PEAKS_COPY = {"hbm_gbs": 6532.9, "bf16_tflops": 1627.7, "bf16_tflops_sustained": 1358.9}
# mean DRAM bytes per tcgen05 GEMM launch of this workload (ncu, profiles/r01_gemm_dram_traffic_step.csv)
GEMM_DRAM_BYTES_PER_LAUNCH = 145.4e6
PEAKS_FALLBACK = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


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


# ---------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the code the reference actually runs for this path is HF's LxmertEncoder
# (x-lxmert/src/lxrt/modeling.py:5,80); it is installed in this image, so it is timed directly.
# ---------------------------------------------------------------------------------------------------

def cpu_encoder_runner(B: int):
    import torch
    from xlxmert_b200 import params as P, synth
    from xlxmert_b200.config import DEFAULT_DIMS as D
    torch.set_num_threads(os.cpu_count() or 1)
    sd = P.init_state_dict(P.model_param_specs(D), seed=0)
    batch = synth.make_batch(D, B, L_TOK, V_GRID, seed=0)
    g = torch.Generator().manual_seed(1)
    emb = torch.randn(B, L_TOK, D.hidden, generator=g)
    feats = torch.randn(B, V_GRID, D.feat_dim, generator=g).abs()
    pos = batch["visual_pos"]
    mask = (1.0 - batch["attention_mask"][:, None, None, :].float()) * torch.finfo(torch.float32).min
    kind = "reference"
    try:
        from transformers import LxmertConfig
        from transformers.models.lxmert.modeling_lxmert import LxmertEncoder
        cfg = LxmertConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
        enc = LxmertEncoder(cfg)
        enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")})
        enc.train()

        def step():
            enc.zero_grad(set_to_none=True)
            (vs, _), (ls, _), _ = enc(emb, mask, feats, pos)
            (ls[-1].sum() + vs[-1].sum()).backward()
    except Exception:
        kind = "port"
        from oracle import lxrt_oracle as O
        sde = {k: v.clone().requires_grad_(True) for k, v in O.sub(sd, "encoder").items()}

        def step():
            for v in sde.values():
                v.grad = None
            ls, vs = O.encoder(sde, emb, mask, feats, pos, None, heads=D.heads, n_l=D.l_layers, n_r=D.r_layers,
                               n_x=D.x_layers)
            (ls[-1].sum() + vs[-1].sum()).backward()
    return step, kind, torch.get_num_threads()


def time_cpu(B: int, steps: int, warmup: int):
    step, kind, threads = cpu_encoder_runner(B)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return dict(value=B / dt, unit=UNIT, cores=threads, kind=kind,
                sample=f"encoder fwd+bwd on B={B} samples/step (of the {BATCH}-sample batch), {steps} timed steps, fp32"), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    Bs = 8
    cb, dt = time_cpu(Bs, max(1, args.steps), max(1, min(args.warmup, 2)))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus, extra={"reference_step": f"bounded sample, B={Bs} per step on host cores"}),
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    return line


def workload_config(n_gpus, extra=None):
    c = {"workload": "9L/5R/5X LXMERT encoder fwd+bwd, batch 256 per GPU, 20 text tokens + 8x8x2048 grid feats "
                     "(BASELINE.json configs[1])",
         "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus, "text_tokens": L_TOK, "grid_cells": V_GRID,
         "parallelism": f"dp{n_gpus}", "l2_policy": "working set (~14 GB of saved activations per step) is far larger than the 126 MB L2"}
    if extra:
        c.update(extra)
    return c


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
    from xlxmert_b200 import _lib, params as P, synth
    from xlxmert_b200.config import DEFAULT_DIMS as D
    from xlxmert_b200.encoder import B200LxmertEncoder
    from xlxmert_b200.lxmert import B200LxmertModel
    lib = _lib.load()

    B = args.batch
    passes = args.passes
    torch.manual_seed(0)
    sd = P.init_state_dict(P.model_param_specs(D), seed=0)
    model = B200LxmertModel(D, passes=passes)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).train()
    enc: B200LxmertEncoder = model.encoder
    table = synth.centroid_table(D).to(dev)                      # frozen vis_emb table (lxrt/modeling.py:140-151)
    batch = synth.make_batch(D, B, L_TOK, V_GRID, seed=rank)

    # ---- device-resident inputs for `value`
    g = torch.Generator().manual_seed(100 + rank)
    ids_d = batch["input_ids"].to(dev)
    mask_d = batch["attention_mask"].to(dev)
    with torch.no_grad():
        emb_d = model.embeddings(ids_d).detach()
        feats_d = table[batch["cluster_ids"].to(dev)].contiguous()
    pos_d = batch["visual_pos"].to(dev)
    ext_mask = ((1.0 - mask_d[:, None, None, :].float()) * torch.finfo(torch.float32).min).contiguous()
    g_lang = (torch.randn(B, L_TOK, D.hidden, generator=g) / (B * L_TOK)).to(dev)
    g_vis = (torch.randn(B, V_GRID, D.hidden, generator=g) / (B * V_GRID)).to(dev)

    from xlxmert_b200.parallel import allreduce_gradients, enable_overlapped_gradient_sync
    if world > 1 and not args.no_overlap:
        enable_overlapped_gradient_sync(model)      # stage-wise all-reduce inside the backward

    def allreduce_grads():
        if world > 1 and not enc.arena_reduced:
            arena = enc.last_grad_arena
            dist.all_reduce(arena)
            arena.mul_(1.0 / world)

    def step_resident():
        enc.invalidate_prepared()        # weights change every optimiser step in training: re-split them
        emb = emb_d.requires_grad_(True)
        emb.grad = None
        (vs, _), (ls, _), _ = enc(emb, ext_mask, feats_d, pos_d)
        torch.autograd.backward([ls[-1], vs[-1]], [g_lang, g_vis])
        allreduce_grads()
        for p in enc.parameters():
            p.grad = None

    # ---- end-to-end through the public module API with HOST inputs
    ids_h = batch["input_ids"].pin_memory()
    mask_h = batch["attention_mask"].pin_memory()
    cids_h = batch["cluster_ids"].pin_memory()
    loss_h = torch.empty((), dtype=torch.float32).pin_memory()
    h2d = ids_h.numel() * 8 + mask_h.numel() * 1 + cids_h.numel() * 8
    d2h = 4

    e2e_marks = []

    def step_e2e():
        enc.invalidate_prepared()
        ids = ids_h.to(dev, non_blocking=True)
        am = mask_h.to(dev, non_blocking=True)
        cids = cids_h.to(dev, non_blocking=True)
        feats = table[cids]                                       # vis_emb(cluster_ids), lxrt/modeling.py:185-186
        out = model(input_ids=ids, visual_feats=feats, visual_pos=pos_d, attention_mask=am)
        loss = (out[0] * g_lang).sum() + (out[1] * g_vis).sum() + out[2].mean()
        loss.backward()
        allreduce_gradients(model)                                # encoder arena + one flat buffer for the rest
        loss_h.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()                 # the step's result is read on the host
        for p in model.parameters():
            p.grad = None
        e2e_marks.append(time.perf_counter())
        return float(loss_h)

    def timed(fn, steps, warmup):
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
        try:
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
        finally:
            gc.enable()
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

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, launches = timed(step_resident, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _ = timed(step_e2e, max(3, args.steps), 5)
    e2e_steps = [round((b - a) * 1e3, 2) for a, b in zip(e2e_marks[4:], e2e_marks[5:])]   # host clock, timed steps only

    # ---- roofline of the dominant kernel (tcgen05 GEMM), events around every launch, outside the timed region
    lib.xlx_profile_gemm_begin()
    for _ in range(2):
        step_resident()
    tot_ms, tot_fl, n = C.c_double(), C.c_double(), C.c_int64()
    _lib.check("xlx_profile_gemm_end", lib.xlx_profile_gemm_end(C.byref(tot_ms), C.byref(tot_fl), C.byref(n)))
    peaks, peak_src = measured_peaks()
    achieved = tot_fl.value / (tot_ms.value * 1e-3) / 1e12 if tot_ms.value > 0 else 0.0
    peak = peaks["bf16_tflops_sustained"]
    roofline = {"bound": "tensor", "kernel": "xlx::gemm_kernel<32,…> (tcgen05 bf16x3 GEMM, all Linear fwd/dgrad/wgrad)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": GEMM_DRAM_BYTES_PER_LAUNCH if (passes == 3 and B == BATCH) else None,
                "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the 319 GEMM launches of "
                                  "one step (profiles/r01_gemm_dram_traffic_step.csv)",
                "peak_source": peak_src + ", sustained bf16 (kernel timed inside a long step)",
                "launches_per_step": n.value // 2, "avg_launch_us": tot_ms.value * 1e3 / max(n.value, 1),
                "gemm_share_of_step": (tot_ms.value / 2) / ms_step,
                "executed_tensor_tflops": achieved * (3 if passes == 3 else 1),
                "note": "achieved counts ALGORITHMIC FLOPs (2MNK); bf16x3 executes 3 MMAs per algorithmic MAC, "
                        "so frac is capped at 1/3 by construction in the fp32-parity mode. The per-launch timing pass "
                        "keeps every kernel on one stream (events between launches), so these durations exclude the "
                        "two-stream / dependent-launch overlap the timed step enjoys"}

    if rank == 0:
        value = B * world / (ms_step * 1e-3)
        e2e_value = B * world / (ms_e2e * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16x3 (split-bf16 tensor-core GEMMs, fp32 accumulate; fp32 elsewhere)"
                if passes == 3 else "bf16", "data": "synthetic",
                "config": workload_config(world, extra={"passes": passes, "batch_per_gpu": B,
                                                        "step_tflop_algorithmic": ENC_GFLOP_FWD_BWD * B / 1e3}),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e, "host_ms_each_step": e2e_steps,
                        "api": "B200LxmertModel.forward(input_ids, visual_feats=vis_emb(cluster_ids), visual_pos, "
                               "attention_mask) + loss.backward(), pinned host ids in, loss scalar out"},
                "gpu_launches": int(launches),
                "clocks": clocks, "roofline": roofline,
                "step_algorithmic_tflops": ENC_GFLOP_FWD_BWD * B / 1e3 / (ms_step * 1e-3)}
        if world == 1 and not args.no_cpu:
            cb, _ = time_cpu(8, 2, 1)
            line["cpu_baseline"] = cb
    else:
        line = None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--passes", type=int, default=3, choices=[1, 3])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
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
