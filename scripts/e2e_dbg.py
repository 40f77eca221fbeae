"""Diagnostic: per-segment host times of bench.py's end-to-end step (forward call, backward call, sync)."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.argv = ["bench.py", "--steps", "12", "--warmup", "3", "--no-cpu"]
sys.path.insert(0, ROOT)
import torch
src = open(os.path.join(ROOT, "bench.py")).read()
src = src.replace("""        ids = ids_h.to(dev, non_blocking=True)""", """        _t = [time.perf_counter()]
        ids = ids_h.to(dev, non_blocking=True)""")
src = src.replace("""        loss = (out[0] * g_lang).sum() + (out[1] * g_vis).sum() + out[2].mean()
        loss.backward()
        allreduce_gradients(model)""", """        _t.append(time.perf_counter())
        loss = (out[0] * g_lang).sum() + (out[1] * g_vis).sum() + out[2].mean()
        loss.backward()
        _t.append(time.perf_counter())
        allreduce_gradients(model)""")
src = src.replace("""        torch.cuda.current_stream().synchronize()                 # the step's result is read on the host
""", """        torch.cuda.current_stream().synchronize()                 # the step's result is read on the host
        _t.append(time.perf_counter())
        SEG.append([round((b - a) * 1e3, 1) for a, b in zip(_t, _t[1:])] + [torch.cuda.memory_stats().get("num_device_alloc", 0), torch.cuda.memory_stats().get("num_device_free", 0)])
""")
src = src.replace("def main():", "SEG = []\ndef main():", 1)
torch.cuda.memory._record_memory_history(max_entries=200000)
exec(compile(src, "bench_dbg", "exec"))
snap = torch.cuda.memory._snapshot()
for tr in snap["device_traces"]:
    evs = [e for e in tr if e["action"] in ("segment_alloc", "segment_free")]
    for e in evs[-14:]:
        frames = [f"{os.path.basename(f['filename'])}:{f['line']}:{f['name']}" for f in e.get("frames", [])
                  if "site-packages" not in f["filename"]][:4]
        print("SEG", e["action"], round(e["size"] / 2**20, 1), "MiB", frames, file=sys.stderr)
for s in SEG:
    print("fwd_call, bwd_call, sync(ms), n_cudaMalloc, n_cudaFree:", s, file=sys.stderr)
