"""icoFoam (PISO) step on the device, timed: lid-driven cavity n^3, momentum predictor
+ 2 PISO correctors per step, fixed inner iterations (tolerance 0) so that the work per step is the same everywhere.
  python tools/bench_icofoam.py --n 128 [--steps 3]   -> one JSON line"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

capi = importlib.import_module("rapidcfd-dev_b200.capi")
ico = importlib.import_module("rapidcfd-dev_b200.icofoam")
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=128)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--p-iters", type=int, default=50)
ap.add_argument("--u-iters", type=int, default=3)
args = ap.parse_args()
ctx = capi.Context(0)
m, dev = ico.cavity(capi, ctx, torch, args.n)
uc = dict(tolerance=0.0, relTol=0.0, maxIter=args.u_iters - 1)
pc = dict(tolerance=0.0, relTol=0.0, maxIter=args.p_iters - 1)
for _ in range(2):
    dev.step(nCorr=2, UControls=uc, pControls=pc)
torch.cuda.synchronize()
l0 = ctx.launches
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    perfs, cont = dev.step(nCorr=2, UControls=uc, pControls=pc)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
N = args.n ** 3
# solver iterations per step: 3 components x u_iters (PBiCG) + 2 x p_iters (PCG)
its = 3 * args.u_iters + 2 * args.p_iters
print(json.dumps({"workload": f"icoFoam cavity {args.n}^3, momentumPredictor + 2 PISO correctors, PBiCG/DILU x{args.u_iters} per "
                              f"component, PCG/DIC x{args.p_iters} per corrector", "ms_per_step": ms,
                  "Mcell_steps_per_s": N / (ms * 1e-3) / 1e6, "solver_iterations_per_step": its,
                  "launches_per_step": (ctx.launches - l0) / args.steps,
                  "continuity": [float(cont[-1][0]), float(cont[-1][1])]}))
dev.close()
ctx.close()
