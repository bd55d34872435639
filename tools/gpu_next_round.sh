# First GPU call of the next round (1 GPU, ~6 min of box time): everything written after the round-1 GPU
# budget ran out gets its first run on a B200, and the kernels that only have CUDA-event numbers get ncu captures.
#   gpurun --timeout 900 -- 'bash tools/gpu_next_round.sh'
# then, on 2 GPUs (the glue over processor patches):
#   gpurun --gpus 2 --timeout 600 -- 'python -m pytest tests/test_zzz_fvm_gpu.py tests/test_gpu_multi.py -m gpu -q -k "multi_gpu" 2>&1 | tail -15'
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02a_tests.log
# row a17 (csrc/fvmatrix.cu) has never run on a GPU: its own log, verbose, and under compute-sanitizer once
python -m pytest tests/test_zzz_fvm_gpu.py -m gpu -v 2>&1 | tail -30 > gpurun_out/r02a_fvm_tests.log
compute-sanitizer --tool memcheck python -m pytest tests/test_zzz_fvm_gpu.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02a_fvm_memcheck.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE-OK')" > gpurun_out/r02a_smoke.log 2>&1
python bench.py --impl reference > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err
python bench.py > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err
# engine variants (Jacobi, residual, AINV, asymmetric Amul/Tmul), face sums, GAMG transfer kernels
ncu --set full --clock-control none --import-source on \
    -k regex:"JacobiOp|ResidualOp|AinvOp|OffDiagOp|surface_integrate|gauss_grad|interpolate_linear|laplacian_fill|convection_fill|grad_linear|flux_linear|faceH_kernel" \
    -c 24 -f -o gpurun_out/r02a_kernels python bench_kernels.py --n 128 --reps 1 > gpurun_out/r02a_ncu_kernels.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"restrict_kernel|prolong_kernel|agg_diag_kernel|agg_faces_kernel|dense_apply" \
    -c 16 -f -o gpurun_out/r02a_gamg python bench_kernels.py --n 64 --reps 1 > gpurun_out/r02a_ncu_gamg.log 2>&1
# section 8(f) rank 2: a 64^3 cavity, 5 steps, log + wall time
python - > gpurun_out/r02a_icofoam.log 2>&1 <<'PY'
import importlib, sys, time, tempfile, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
from test_icofoam_case_cpu import write_cavity
ff = importlib.import_module("rapidcfd-dev_b200.foamfile"); meshmod = importlib.import_module("rapidcfd-dev_b200.mesh")
capi = importlib.import_module("rapidcfd-dev_b200.capi"); ico = importlib.import_module("rapidcfd-dev_b200.icofoam")
root = os.path.join(tempfile.mkdtemp(), "cavity")
write_cavity(ff, meshmod, root, 64, steps=5)
ctx = capi.Context(0)
t0 = time.perf_counter(); case, hist = ico.run_case(capi, ctx, torch, root); torch.cuda.synchronize()
print("icoFoam 64^3, 5 steps:", time.perf_counter() - t0, "s wall (incl. case reading and layout build)")
PY
cat gpurun_out/r02a_tests.log gpurun_out/r02a_fvm_tests.log gpurun_out/r02a_smoke.log | tail -40
cut -c1-300 gpurun_out/r02a_bench_ref.json gpurun_out/r02a_bench_n1.json
