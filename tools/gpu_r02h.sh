# round 2, call H (1 GPU): final evidence run -- tests, smoke, bench (+ reference arm), kernel table, ncu launch list,
# ncu --set full of the engine kernels (256^3) and of the fv / fvMatrix / GAMG kernels (128^3), icoFoam step timing
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02h_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE-OK')" > gpurun_out/r02h_smoke.log 2>&1
timeout 900 python bench.py --impl reference > gpurun_out/r02h_bench_ref.json 2> gpurun_out/r02h_bench_ref.err
timeout 1200 python bench.py > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err
timeout 300 python bench.py --n 128 --no-cpu-baseline --no-secondary > gpurun_out/r02h_bench_n128.json 2>/dev/null
B200LDU_PCG_DEFERRED=0 timeout 300 python bench.py --n 128 --no-cpu-baseline --no-secondary --no-parity > gpurun_out/r02h_bench_n128_4launch.json 2>/dev/null
B200LDU_PCG_DEFERRED=0 timeout 300 python bench.py --no-cpu-baseline --no-secondary --no-parity > gpurun_out/r02h_bench_n1_4launch.json 2>/dev/null
for br in 768 1024 1216 2368; do B200LDU_BAND_ROWS=$br timeout 300 python bench.py --n 128 --no-cpu-baseline --no-secondary --no-parity > gpurun_out/r02h_bench_n128_band$br.json 2>/dev/null; done
timeout 600 python bench_kernels.py > gpurun_out/r02h_kernels.json 2> gpurun_out/r02h_kernels.txt
timeout 300 python tools/bench_icofoam.py --n 128 > gpurun_out/r02h_icofoam_n128.json 2> gpurun_out/r02h_icofoam.err
timeout 300 python tools/bench_icofoam.py --n 192 > gpurun_out/r02h_icofoam_n192.json 2>> gpurun_out/r02h_icofoam.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-parity > gpurun_out/r02h_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"engine_kernel|fill_val|fill_diag" -s 30 -c 8 -f -o gpurun_out/r02h_engine python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-parity > gpurun_out/r02h_ncu_engine.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"JacobiOp|ResidualOp|AinvOp|OffDiagOp|CoeffSumOp|surface_integrate|gauss_grad|interpolate_linear|laplacian_upper|convection_faces|neg_sum_diag|grad_linear|flux_linear|faceH_kernel|H_kernel|relax_kernel|A_kernel|flux_internal|boundary_|residual_source|restrict|prolong|agg_|dense_apply|row_sum|limiter|limited_weights|binary_kernel" -c 60 -f -o gpurun_out/r02h_kernels python bench_kernels.py --n 128 --reps 1 > gpurun_out/r02h_ncu_kernels.log 2>&1
tail -5 gpurun_out/r02h_tests.log; tail -2 gpurun_out/r02h_smoke.log
for f in gpurun_out/r02h_bench_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:d.get(k) for k in ("value","value_solver_only","gpu_launches")}, (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("cores"))
except Exception as e: print("ERR", sys.argv[1], e)
PY
done
cat gpurun_out/r02h_icofoam_n128.json gpurun_out/r02h_icofoam_n192.json; tail -3 gpurun_out/r02h_icofoam.err
head -30 gpurun_out/r02h_kernels.txt
ls -la gpurun_out/*.ncu-rep
