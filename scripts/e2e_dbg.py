import sys, os, json, time
sys.argv = ["bench.py", "--steps", "12", "--warmup", "3", "--no-cpu"]
sys.path.insert(0, "/root/repo")
import torch
import bench
# monkeypatch timed to print per-step durations
src = open("/root/repo/bench.py").read()
src = src.replace("""        e0.record()
        for _ in range(steps):
            fn()
        e1.record()""", """        e0.record()
        evs = []
        for _ in range(steps):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); a.record(); fn(); b.record(); evs.append((a, b, time.perf_counter() - t0))
        e1.record()
        torch.cuda.synchronize()
        print("per-step", fn.__name__, [round(a.elapsed_time(b), 1) for a, b, _ in evs], "host", [round(h * 1e3, 1) for _, _, h in evs], "alloc_retries", torch.cuda.memory_stats().get("num_alloc_retries"), "reserved_GB", round(torch.cuda.memory_reserved() / 2**30, 1), file=sys.stderr)""")
exec(compile(src, "bench_dbg", "exec"))
