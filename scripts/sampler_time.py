#!/usr/bin/env python
"""NAR sampling (4 steps + decode) timing at B=32 (BASELINE.json configs[4]); optional argument: cuda_graph."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
r = bench.bench_sampler(torch.device("cuda", 0), bench.measured_peaks()[0])
print(json.dumps({"ms": r["ms"], "img_per_s": r["images_per_s"], "env": {k: v for k, v in os.environ.items() if k.startswith("XLX_")}}))
