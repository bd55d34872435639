#!/usr/bin/env python
"""icoFoam on the B200 core: `python rapidcfd-dev_b200/apps/icoFoam.py -case <dir> [-device 0] [-steps N]`, or
`torchrun --nproc-per-node N rapidcfd-dev_b200/apps/icoFoam.py -case <dir> -parallel` on a decomposed case.
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
    ap.add_argument("-parallel", action="store_true", help="under torchrun: rank r runs processor<r>/ on GPU LOCAL_RANK")
    args = ap.parse_args()
    import numpy as np
    import torch
    capi = importlib.import_module("rapidcfd-dev_b200.capi")
    ico = importlib.import_module("rapidcfd-dev_b200.icofoam")
    if not args.parallel:
        ctx = capi.Context(args.device)
        case, _ = ico.run_case(capi, ctx, torch, args.case, maxSteps=args.steps)
    else:
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ctx = capi.Context(local)
        ctx.comm_init_from_torch()

        def allsum(v):
            t = torch.from_numpy(np.array(v, dtype=np.float64)).to(ctx.device)
            dist.all_reduce(t)
            return t.cpu().numpy()
        rank = dist.get_rank()
        case, _ = ico.run_case(capi, ctx, torch, args.case, log=print if rank == 0 else (lambda s: None),
                               maxSteps=args.steps, rank=rank, allsum=allsum)
    case.close()
    ctx.close()


if __name__ == "__main__":
    main()
