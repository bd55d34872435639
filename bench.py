#!/usr/bin/env python
"""bench.py -- Mcell-iters/s of the PCG pressure solve on the synthetic 256^3 hex cavity
(BASELINE.json metric / configs[1]) + Amul SpMV achieved HBM GB/s against the roofline.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # CPU restatement (oracle), all host threads

A "step" is one pressure solve: PCG + DIC(=AINV in the reference) with a fixed number of
inner iterations (tolerance 0, so every implementation does identical work).
value  = nCells(global) * iterations * steps / time, inputs resident in HBM (caller order);
e2e    = same through b200ldu_solve_host with pinned HOST psi/source (H2D + solve + D2H timed).
N > 1: brick decomposition, one rank per GPU (torchrun), halo over NCCL, weak in work/GPU? No:
the global mesh is fixed (256^3) => "strong" scaling.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full`
    capture (profiles/r01_ncu_full_engine_raw.csv): the Amul kernel, and one steady-state fused
    PCG iteration (last PcgAinvOp launch + last PcgAmulOp launch of the capture).  None if absent."""
    import csv
    p = os.path.join(ROOT, "profiles", "r01_ncu_full_engine_raw.csv")
    out = {"amul": None, "pcg_iteration": None}
    try:
        rows = list(csv.reader(open(p)))
        hdr, units = rows[0], rows[1]
        kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        last = {}
        for r in rows[2:]:
            byts = float(r[rd]) * mult[units[rd]] + float(r[wr]) * mult[units[wr]]
            for key in ("AmulOp<0>", "PcgAinvOp", "PcgAmulOp"):
                if key in r[kn]:
                    last[key] = byts
        out["amul"] = last.get("AmulOp<0>")
        if "PcgAinvOp" in last and "PcgAmulOp" in last:
            out["pcg_iteration"] = last["PcgAinvOp"] + last["PcgAmulOp"]
    except Exception:  # noqa: BLE001
        pass
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_case(meshmod, n, nRanks, rank):
    mesh = meshmod.hex_mesh(n) if nRanks == 1 else meshmod.decompose(n, nRanks, rank)
    coef = meshmod.pressure_laplacian(mesh)
    b = meshmod.cell_field_global(mesh, 9)  # zero-mean-ish random RHS (timing run)
    return mesh, coef, b


def host_threads(orc):
    """Threads for the CPU arm: one per physical core the process may run on.  The row sweeps are
    memory-bound; on the GPU boxes (2-way SMT) running one thread per logical CPU measured 3x
    slower (31 vs 95 Mcell-iters/s at 256^3), so SMT siblings are left idle."""
    n = orc.max_threads()
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    try:
        import psutil
        phys = psutil.cpu_count(logical=False)
        if phys:
            n = min(n, phys)
    except Exception:  # noqa: BLE001
        pass
    return max(1, n)


def set_omp_threads(n):
    """libgomp is shared by the oracle port and the OpenMP build of the reference code: set its team size."""
    import ctypes
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass


def cpu_pcg(meshmod, orc, mesh, coef, nT):
    """The CPU PCG used by the reference arm and the cpu_baseline leg.  Where the reference's own solver
    sources were compiled here (oracle/_ref/libref_solvers_omp.so: PCG.C, AINVPreconditioner.C,
    lduMatrixATmul.C ... on thrust's OpenMP host back end, oracle/ref_harness/) that code runs, with the
    reference's default favourSpeedOverMemory 2 (etc/controlDict:66) -- kind "reference"; otherwise the
    oracle's OpenMP port -- kind "port".  Returns (run(iters, b) -> nIterations, kind, description)."""
    try:
        from oracle import ref_ldu
        if ref_ldu.omp_available():
            os_, ls, lo = ref_ldu.ldu_arrays(mesh.nCells, mesh.lower, mesh.upper)
            fixed = (mesh.nCells, np.ascontiguousarray(mesh.lower, np.int32), np.ascontiguousarray(mesh.upper, np.int32),
                     os_, ls, lo, np.ascontiguousarray(coef["diag"]), np.ascontiguousarray(coef["upper"]), None)
            set_omp_threads(nT)
            z = np.zeros(mesh.nCells)

            def run_ref(iters, b):
                _, p = ref_ldu.solve("PCG", "DIC", *fixed, z, b, tolerance=0.0, maxIter=iters - 1, favourSpeed=2,
                                     omp=True)
                return p["nIterations"]
            run_ref(1, np.ones(mesh.nCells))   # loads the library, starts the OpenMP team
            return run_ref, "reference", ("reference PCG.C + AINVPreconditioner.C + lduMatrixATmul.C compiled for the "
                                           "host, thrust OpenMP back end")
    except Exception as e:  # noqa: BLE001 -- any problem with the optional library: use the port
        sys.stderr.write(f"reference-code CPU arm unavailable ({e}); using the oracle port\n")
    oa = orc.Addr(mesh.nCells, mesh.lower, mesh.upper)
    om = orc.Matrix(oa, coef["diag"], coef["upper"], None)

    def run_port(iters, b):
        _, perf = om.pcg_omp("DIC", np.zeros(mesh.nCells), b, nThreads=nT, tolerance=0.0, maxIter=iters - 1)
        return perf.nIterations
    run_port._keep = (oa, om)
    return run_port, "port", "oracle OpenMP rows (oracle/ldu_oracle_omp.c)"


def cpu_baseline_leg(meshmod, mesh, coef, b, n, ci):
    """cpu_baseline object of the JSON line: `ci` PCG iterations of the CPU arm (cpu_pcg), the oracle port
    beside it when the reference code ran, and stock OpenFOAM's serial DIC-PCG for context."""
    from oracle import ldu_oracle as orc
    nT = host_threads(orc)
    run, kind, what = cpu_pcg(meshmod, orc, mesh, coef, nT)
    t0 = time.perf_counter()
    nit = run(ci, b)
    cdt = time.perf_counter() - t0
    assert nit == ci
    cpu = {"value": mesh.nCells * ci / cdt / 1e6, "unit": "Mcell-iters/s", "cores": nT, "kind": kind,
           "sample": f"{n}^3 cells x {ci} PCG(AINV) iterations, {what}, {cdt:.1f} s"}
    oa = orc.Addr(mesh.nCells, mesh.lower, mesh.upper)
    om = orc.Matrix(oa, coef["diag"], coef["upper"], None)
    if kind == "reference":   # the oracle's own OpenMP port beside it
        t0 = time.perf_counter()
        om.pcg_omp("DIC", np.zeros(mesh.nCells), b, nThreads=nT, tolerance=0.0, maxIter=ci - 1)
        pdt = time.perf_counter() - t0
        cpu["port"] = {"value": mesh.nCells * ci / pdt / 1e6, "unit": "Mcell-iters/s", "cores": nT,
                       "sample": f"oracle OpenMP rows, {pdt:.1f} s"}
    # stock CPU OpenFOAM numerics for context (true DIC + face-loop Amul, one core = one rank)
    si = max(2, min(8, ci // 2))
    t0 = time.perf_counter()
    om.pcg_stock_dic(np.zeros(mesh.nCells), b, tolerance=0.0, maxIter=si - 1)
    sdt = time.perf_counter() - t0
    cpu["stock_dic_serial"] = {"value": mesh.nCells * si / sdt / 1e6, "unit": "Mcell-iters/s", "cores": 1,
                               "sample": f"{n}^3 cells x {si} PCG(true DIC) iterations, serial, {sdt:.1f} s"}
    return cpu


def run_reference(args, rank, world):
    """CPU arm on all physical host cores: the reference's own PCG loop where it compiled (oracle/_ref), else
    the oracle's OpenMP port -- RapidCFD numerics either way (AINV for DIC)."""
    if rank != 0:
        return
    meshmod = importlib.import_module("rapidcfd-dev_b200.mesh")
    from oracle import ldu_oracle as orc
    n = args.n
    mesh, coef, b = build_case(meshmod, n, 1, 0)
    nT = host_threads(orc)
    run, kind, what = cpu_pcg(meshmod, orc, mesh, coef, nT)
    iters = args.ref_iters
    for _ in range(args.warmup):
        run(iters, b)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        nit = run(iters, b)
    dt = time.perf_counter() - t0
    assert nit == iters
    val = mesh.nCells * iters * args.steps / dt / 1e6
    line = {"impl": "reference", "metric": "Mcell-iters/sec (PCG pressure solve, 256^3 hex cavity)",
            "value": val, "unit": "Mcell-iters/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"icoFoam cavity {n}^3 hex, PCG+DIC(AINV) pressure solve", "n": n,
                       "iterations_per_step": iters, "preconditioner": "DIC->AINV"},
            "cpu_baseline": {"value": val, "unit": "Mcell-iters/s", "cores": nT, "kind": kind,
                             "sample": f"{n}^3 cells x {iters} PCG iterations per step; {what}"},
            "e2e": {"value": val, "unit": "Mcell-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--iters", type=int, default=50, help="PCG iterations per step")
    ap.add_argument("--ref-iters", type=int, default=50,
                    help="PCG iterations per step of the CPU arm (same as --iters: per-solve set-up amortised alike)")
    ap.add_argument("--cpu-baseline-iters", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    capi = importlib.import_module("rapidcfd-dev_b200.capi")
    meshmod = importlib.import_module("rapidcfd-dev_b200.mesh")
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    n, iters = args.n, args.iters
    mesh, coef, b = build_case(meshmod, n, world, rank)
    nGlobal = n ** 3

    ctx = capi.Context(local)
    if world > 1:
        ctx.comm_init_from_torch()
    dev = ctx.device
    t0 = time.perf_counter()
    addr = capi.mesh_to_device(ctx, mesh)
    t_layout = time.perf_counter() - t0
    mat = capi.LduMatrix(addr)
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    diag, upper = tt(coef["diag"]), tt(coef["upper"])
    bou = tt(coef["bou"]) if len(coef["bou"]) else None
    mat.set(diag, upper, None, bou, bou)
    src = tt(b)
    psi = torch.zeros(mesh.nCells, dtype=torch.float64, device=dev)
    kw = dict(tolerance=0.0, maxIter=iters - 1)
    info = addr.info()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident arm ----------------
    for _ in range(args.warmup):
        psi.zero_()
        perf, _ = mat.solve("PCG", "DIC", psi, src, **kw)
    assert perf.nIterations == iters, perf.nIterations
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        psi.zero_()
        perf, _ = mat.solve("PCG", "DIC", psi, src, **kw)
    ev1.record()
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = ctx.launches - l0
    value = nGlobal * iters * args.steps / (ms * 1e-3) / 1e6

    # ---------------- Amul kernel alone (roofline) ----------------
    vl = addr.vec_len
    xb = torch.zeros(vl, dtype=torch.float64, device=dev)
    yb = torch.zeros(vl, dtype=torch.float64, device=dev)
    capi.check(capi.lib().b200ldu_to_banded(addr.h, capi._dp(src), capi._dp(xb)))
    nA = 20
    for _ in range(3):
        capi.check(capi.lib().b200ldu_amul_banded(mat.h, capi._dp(xb), capi._dp(yb)))
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(nA):
        capi.check(capi.lib().b200ldu_amul_banded(mat.h, capi._dp(xb), capi._dp(yb)))
    a1.record()
    torch.cuda.synchronize()
    amul_ms = a0.elapsed_time(a1) / nA
    N, F = mesh.nCells, mesh.nFaces
    amul_bytes = 24 * N + 16 * F  # psi, diag, Apsi + upper, owner, neighbour
    amul_gbs = amul_bytes / (amul_ms * 1e-3) / 1e9
    clocks = sampler.stop() if rank == 0 else None
    peak, peak_src = peaks()

    # ---------------- end-to-end arm: host buffers through the C ABI ----------------
    psi_h = torch.zeros(mesh.nCells, dtype=torch.float64).pin_memory()
    src_h = torch.from_numpy(b.copy()).pin_memory()
    psi_np, src_np = psi_h.numpy(), src_h.numpy()
    for _ in range(2):
        psi_np[:] = 0
        mat.solve_host("PCG", "DIC", psi_np, src_np, **kw)
    barrier()
    # each step = one b200ldu_solve_host call (H2D psi+source, solve, D2H psi), bracketed by CUDA
    # events; the harness's host-side reset of the initial guess between steps is not solver work
    e2e_local = 0.0
    for _ in range(args.steps):
        psi_np[:] = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        mat.solve_host("PCG", "DIC", psi_np, src_np, **kw)
        e1.record()
        torch.cuda.synchronize()
        e2e_local += e0.elapsed_time(e1)
    barrier()
    e2e_ms = max_over_ranks(e2e_local)
    e2e_val = nGlobal * iters * args.steps / (e2e_ms * 1e-3) / 1e6

    # ---------------- CPU baseline (rank 0, N=1 only, bounded sample) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg(meshmod, mesh, coef, b, n, args.cpu_baseline_iters)

    if rank == 0:
        pcg_bytes = (160 * N + 32 * F)  # per iteration and rank, reference op list with AINV
        it_ms = ms / (iters * args.steps)
        pcg_gbs = pcg_bytes / (it_ms * 1e-3) / 1e9
        pcg_min_gbs = (104 * N + 32 * F) / (it_ms * 1e-3) / 1e9
        traffic = ncu_traffic()
        line = {
            "metric": "Mcell-iters/sec (PCG pressure solve, 256^3 hex cavity)", "value": value,
            "unit": "Mcell-iters/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"icoFoam cavity {n}^3 hex, PCG+DIC(AINV) pressure solve", "n": n,
                       "iterations_per_step": iters, "preconditioner": "DIC->AINV", "decomposition":
                       "x".join(map(str, meshmod.brick_split(world))), "l2": "inputs >> L2 (1.2 GB per SpMV)",
                       "band_rows": info["bandRows"], "bands": info["nBands"], "layout_build_s": round(t_layout, 2)},
            # Dominant kernels of the timed region: the two engine kernels of one fused PCG iteration
            # (PcgAinvOp + PcgAmulOp = 94% of the step in profiles/r01_launches_bench_n256.csv); one
            # "launch" below = one iteration, timed live as step time / iterations.  `achieved` uses
            # SURVEY 8(d)'s PCG-iteration figure 160N+32F (the reference's unfused op list) -- the fused
            # kernels need at least 104N+32F, reported next to it, as is the ncu DRAM traffic.
            "roofline": {"bound": "hbm",
                         "kernel": "engine_kernel<PcgAinvOp> + engine_kernel<PcgAmulOp> (one fused PCG iteration)",
                         "achieved": pcg_gbs, "peak": peak, "unit": "GB/s", "frac": pcg_gbs / peak,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": pcg_bytes,
                         "ms_per_launch": it_ms, "traffic": traffic["pcg_iteration"],
                         "fused": "achieved counts the reference op list's bytes (160N+32F); minimum for the fused pair 104N+32F",
                         "achieved_min_bytes": pcg_min_gbs, "frac_min_bytes": pcg_min_gbs / peak,
                         "dram_gbs": (traffic["pcg_iteration"] / (it_ms * 1e-3) / 1e9) if traffic["pcg_iteration"] and world == 1 else None,
                         "amul": {"kernel": "engine_kernel<AmulOp<0>> (Amul SpMV, banded)", "achieved": amul_gbs,
                                  "frac": amul_gbs / peak, "algorithmic_bytes_per_launch": amul_bytes,
                                  "ms_per_launch": amul_ms, "traffic": traffic["amul"] if world == 1 else None}},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "Mcell-iters/s", "h2d_bytes_per_step": 16 * nGlobal,
                    "d2h_bytes_per_step": 8 * nGlobal, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches), "clocks": clocks,
            "solver_line": perf.line("p"),
        }
        print(json.dumps(line), flush=True)
    mat.close()
    addr.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
