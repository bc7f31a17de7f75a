#!/usr/bin/env python
"""Kernel timeline of REPLAYED training steps (CUDA graphs, the bench's own path): start / duration /
stream of every kernel of a few consecutive steps, from CUPTI activity records (torch.profiler), so the
critical path and the idle gaps of the step can be read off.  Unlike an ncu launch list the kernels run
concurrently and warm, exactly as in the timed region.

    python tools/step_timeline.py [config] [gas|incagg] > timeline.txt
Output: one line per kernel of the middle profiled step (start relative to the step's first kernel, us),
then per-stream busy time and the step's span."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import incagg_gnn_b200  # noqa: F401
from incagg_gnn_b200.train import build, GraphedTrainer
from torch.profiler import profile, ProfilerActivity

config = sys.argv[1] if len(sys.argv) > 1 else "C3"
vr = len(sys.argv) > 2 and sys.argv[2] == "incagg"
STEPS = 5
dev = torch.device("cuda:0")
run = build(config, device=dev, seed=0, overrides=dict(VR_update=vr))
model, loader, opt, conf = run["model"], run["train_loader"], run["optimizer"], run["conf"]
model.train()
tr = GraphedTrainer(model, loader, opt, VR_update=vr, grad_norm=conf["grad_norm"],
                    pipeline_collate=loader.fixed_batches)
groups = loader._batches_of_epoch()
tr.warmup(groups[0])
for ids in groups[:24]:
    tr.capture(ids)
torch.cuda.synchronize()
tr.run(list(groups[:12]))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.run(list(groups[12:12 + STEPS]))
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"]
      if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
ev.sort(key=lambda e: e["ts"])
# steps are separated by the Adam kernel (last kernel of a step on the main chain)
ends = [i for i, e in enumerate(ev) if "adam" in e["name"]]
if len(ends) < 3:
    raise SystemExit(f"could not find step boundaries ({len(ends)} adam kernels)")
mid = len(ends) // 2
t_prev_end = ev[ends[mid - 1]]["ts"] + ev[ends[mid - 1]]["dur"]
t_end = ev[ends[mid]]["ts"] + ev[ends[mid]]["dur"]
step = [e for e in ev if t_prev_end - 400 <= e["ts"] < t_end]
t0 = t_prev_end
print(f"# step span (adam end -> adam end): {t_end - t_prev_end:.1f} us; kernels listed from 400 us before")
print("# start_us  dur_us  stream  kernel")
busy = {}
for e in step:
    name = e["name"].split("(")[0][:80]
    s = e["args"].get("stream", -1)
    print(f"{e['ts'] - t0:9.1f} {e['dur']:7.1f} {s:>6}  {name}")
    if e["ts"] >= t_prev_end:
        busy[s] = busy.get(s, 0.) + e["dur"]
print("# busy time per stream inside the step:", {k: round(v, 1) for k, v in sorted(busy.items())})
spans = [ev[ends[i]]["ts"] + ev[ends[i]]["dur"] - ev[ends[i - 1]]["ts"] - ev[ends[i - 1]]["dur"]
         for i in range(1, len(ends))]
print("# all step spans:", [round(x, 1) for x in spans])
