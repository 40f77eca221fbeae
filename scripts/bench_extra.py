#!/usr/bin/env python
"""Secondary measurements on one B200 (not the driver's bench line): BASELINE.json configs[2..4] —
full pre-training step per task (B=256), generator forward (B=128), NAR sampling + decode (B=32).
CUDA-event timed, 3 warm-up + N timed iterations, device-resident inputs.  Prints one JSON object."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from xlxmert_b200 import params as P, synth  # noqa: E402
from xlxmert_b200.config import DEFAULT_DIMS as D  # noqa: E402


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / iters
    ev = e0.elapsed_time(e1) / iters
    if abs(wall - ev) > 0.05 * ev:
        print(f"[timed] wall {wall:.2f} ms vs events {ev:.2f} ms", file=sys.stderr)
    return ev


def main():
    import __graft_entry__ as entry
    entry.build()
    from test_pretrain_parity import build_model
    from xlxmert_b200.generator import B200Generator
    from xlxmert_b200.sampler import B200ImggenModel
    out = {}
    model, table = build_model(0)
    model.train()
    B = 256
    batch = {k: v.cuda() for k, v in synth.make_batch(D, B, 20, 64, seed=0).items()}
    labels = dict(word_labels=batch["word_labels"], obj_labels=batch["obj_labels"], matched_labels=batch["matched_labels"])

    def step(task):
        def f():
            ids = batch["masked_input_ids"] if task == "word_mask" else batch["input_ids"]
            for p in model.parameters():
                p.grad = None
            o = model(input_ids=ids, visual_pos=batch["visual_pos"], attention_mask=batch["attention_mask"],
                      cluster_ids=batch["cluster_ids"], vis_mask=batch["vis_mask"], label_dict=labels, task=task)
            o["total_loss"].backward()
        return f
    flop = {"vis_mask": 54.9, "word_mask": 49.1, "matched": 47.2}
    for task in ("vis_mask", "word_mask", "matched"):
        ms = timed(step(task), 5)
        out[f"pretrain_{task}_B{B}"] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3,
                                        "algorithmic_tflops": flop[task] * B / ms,
                                        "note": "FLOPs counted as the reference does them (heads on all rows)"}
    # optimiser half of the step (lxmert_pretrain.py:343-364): fused clip + HF-AdamW over all trainable parameters
    from xlxmert_b200.optim import B200AdamW, lxmert_param_groups
    step("vis_mask")()                                   # leaves gradients behind
    opt = B200AdamW(lxmert_param_groups(model, 0.01), lr=1e-4)
    n_par = sum(p.numel() for p in model.parameters() if p.grad is not None)
    ms = timed(lambda: opt.step(max_grad_norm=1.0), 10)
    out["fused_clip_adamw"] = {"ms": ms, "params_with_grad": n_par, "algorithmic_bytes": 32 * n_par,
                               "achieved_GBps": 32 * n_par / ms / 1e6,
                               "note": "28 B/param AdamW (read g,p,m,v; write p,m,v) + 4 B/param for the norm pass"}
    # whole training iteration as lxmert_pretrain.py:343-364 runs it: forward, backward, clip, AdamW (weights really
    # change every step, so the encoder re-splits them)
    fstep = step("vis_mask")

    def full_iteration():
        fstep()
        opt.step(max_grad_norm=1.0)
    ms = timed(full_iteration, 5)
    out["pretrain_vis_mask_B256_with_optimizer"] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3}
    fstep()
    ref_params = [p for p in model.parameters() if p.grad is not None]
    topt = torch.optim.AdamW(ref_params, lr=1e-4, eps=1e-6, weight_decay=0.01)

    def torch_step():
        torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
        topt.step()
    ms = timed(torch_step, 10)
    out["torch_clip_plus_AdamW_foreach"] = {"ms": ms}
    del model, opt, topt
    torch.cuda.empty_cache()

    G = B200Generator()
    G.load_state_dict(P.init_generator_state_dict(seed=0), strict=True)
    G = G.cuda().eval()
    Bg = 128
    ids = torch.randint(0, D.num_clusters, (Bg, 64), device="cuda")
    code = table.cuda()[ids]
    ms = timed(lambda: G(code.view(Bg, 8, 8, 2048), train=False), 5)
    out[f"generator_fwd_B{Bg}"] = {"ms": ms, "images_per_s": Bg / ms * 1e3, "algorithmic_tflops": 27.755 * Bg / ms}

    pre, table = build_model(0)
    m = B200ImggenModel(D, num_clusters=D.num_clusters)
    m.set_visual_embedding(table.clone())
    m.load_state_dict({k: v for k, v in pre.state_dict().items() if not k.startswith("cls.")}, strict=False)
    m.set_image_generator(G)
    m = m.cuda()
    Bs = 32
    tok = synth.make_batch(D, Bs, 20, 64, seed=3)["input_ids"].cuda()
    ms = timed(lambda: m.sample_image_NAR(tok, n_steps=4), 6, warm=3)
    out[f"sample_NAR4_plus_decode_B{Bs}"] = {"ms": ms, "images_per_s": Bs / ms * 1e3,
                                             "algorithmic_tflops": 100.9 * Bs / ms}
    ms = timed(lambda: m.sample_image_NAR(tok, n_steps=4, cuda_graph=True), 6, warm=3)
    out[f"sample_NAR4_plus_decode_B{Bs}_cuda_graph"] = {"ms": ms, "images_per_s": Bs / ms * 1e3}
    ms = timed(lambda: m.sample_image_NAR(tok, n_steps=4, cache_language=False), 6, warm=3)
    out[f"sample_NAR4_plus_decode_B{Bs}_no_language_cache"] = {"ms": ms, "images_per_s": Bs / ms * 1e3}
    del m, pre, G
    torch.cuda.empty_cache()

    # nearest-centroid assignment (run_kmeans.py:124-143): 10 000 centroids × 2048, device-resident and host-streamed
    import time
    import numpy as np
    from oracle import kmeans_oracle as KO
    from xlxmert_b200.kmeans import B200IndexFlatL2
    cent = table.float()
    index = B200IndexFlatL2(2048)
    index.add(cent)
    Nk = 32768
    xg = (cent[torch.randint(0, cent.shape[0], (Nk,))] + 0.3 * torch.randn(Nk, 2048)).cuda()
    ms = timed(lambda: index.search(xg, 1), 5)
    out["kmeans_assign_device_N32768"] = {"ms": ms, "rows_per_s": Nk / ms * 1e3,
                                          "algorithmic_tflops": 2 * Nk * 2048 * cent.shape[0] / ms / 1e9}
    xh = xg.cpu().numpy()
    xh = np.concatenate([xh] * 4)                      # 131 072 rows = 1 GiB of fp32 features from host memory
    index.search(xh[:Nk], 1)
    t0 = time.perf_counter()
    index.search(xh, 1)
    dt = time.perf_counter() - t0
    out["kmeans_assign_host_streamed_N131072"] = {"ms": dt * 1e3, "rows_per_s": len(xh) / dt,
                                                  "h2d_GBps": xh.nbytes / dt / 1e9}
    t0 = time.perf_counter()
    KO.search_l2_fp32(xh[:4096], cent.numpy())
    dt = time.perf_counter() - t0
    out["kmeans_assign_cpu_oracle_N4096"] = {"ms": dt * 1e3, "rows_per_s": 4096 / dt,
                                             "note": "numpy fp32 restatement of faiss IndexFlatL2 on the host cores"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
