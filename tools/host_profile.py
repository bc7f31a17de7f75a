#!/usr/bin/env python
"""Host-side profile of the training step (cProfile) — where the Python time of a step goes.
    python tools/host_profile.py [--mode gas|incagg] [--steps 60]"""
import argparse, cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="gas")
ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--config", default="C3")
args = ap.parse_args()
import incagg_gnn_b200
from incagg_gnn_b200.train import build, mini_train, mini_test
vr = args.mode == "incagg"
run = build(args.config, device="cuda", overrides=dict(VR_update=vr))
model = run["model"]
mini_test(model, run["eval_loader"], VR_update=vr)
mini_train(model, run["train_loader"], run["criterion"], run["optimizer"], 10, VR_update=vr)
torch.cuda.synchronize()
t0 = time.perf_counter()
mini_train(model, run["train_loader"], run["criterion"], run["optimizer"], args.steps, VR_update=vr)
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"{args.steps} steps: host issue {t_issue*1e3/args.steps:.2f} ms/step, incl. GPU drain {t_all*1e3/args.steps:.2f} ms/step")
pr = cProfile.Profile()
pr.enable()
mini_train(model, run["train_loader"], run["criterion"], run["optimizer"], args.steps, VR_update=vr)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45)
print(s.getvalue()[:9000])
