#!/usr/bin/env python
"""icoFoam on the B200 core: `python rapidcfd-dev_b200/apps/icoFoam.py -case <dir> [-device 0] [-steps N]`.
Reads the case as the reference's application does (rapidcfd-dev_b200/icofoam.py: run_case), runs the PISO time
loop with every field on the device and writes U, p of the last step.  Needs a CUDA device (no CPU fallback)."""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-case", default=".")
    ap.add_argument("-device", type=int, default=0)
    ap.add_argument("-steps", type=int, default=None, help="stop after this many time steps")
    args = ap.parse_args()
    import torch
    capi = importlib.import_module("rapidcfd-dev_b200.capi")
    ico = importlib.import_module("rapidcfd-dev_b200.icofoam")
    ctx = capi.Context(args.device)
    case, _ = ico.run_case(capi, ctx, torch, args.case, maxSteps=args.steps)
    case.close()
    ctx.close()


if __name__ == "__main__":
    main()
