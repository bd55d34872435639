# round 2, call I (1 GPU): final evidence run -- tests, smoke, bench (+ reference arm), kernel table, ncu launch list,
# ncu --set full of the engine kernels (256^3), icoFoam step timing.  .ncu-rep files go to /tmp (gpurun_out is capped at 64 MiB)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02i_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE-OK')" > gpurun_out/r02i_smoke.log 2>&1
timeout 900 python bench.py --impl reference > gpurun_out/r02i_bench_ref.json 2> gpurun_out/r02i_bench_ref.err
timeout 1200 python bench.py > gpurun_out/r02i_bench_n1.json 2> gpurun_out/r02i_bench_n1.err
timeout 300 python bench.py --n 128 --no-cpu-baseline --no-secondary > gpurun_out/r02i_bench_n128.json 2>/dev/null
B200LDU_BAND_ROWS=2048 timeout 300 python bench.py --no-cpu-baseline --no-secondary --no-parity > gpurun_out/r02i_bench_n1_band2048.json 2>/dev/null
timeout 300 python tools/bench_engine.py > gpurun_out/r02i_engine.jsonl 2> gpurun_out/r02i_engine.err
timeout 600 python bench_kernels.py > gpurun_out/r02i_kernels.json 2> gpurun_out/r02i_kernels.txt
timeout 300 python tools/bench_icofoam.py --n 128 > gpurun_out/r02i_icofoam_n128.json 2> gpurun_out/r02i_icofoam.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02i_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-parity > gpurun_out/r02i_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"engine_kernel|fill_val|fill_diag" -s 30 -c 6 -f -o /tmp/r02i_engine python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-parity > gpurun_out/r02i_ncu_engine.log 2>&1
ncu -i /tmp/r02i_engine.ncu-rep --page raw --csv > gpurun_out/r02i_ncu_full_engine_raw.csv 2>/dev/null
ncu -i /tmp/r02i_engine.ncu-rep --page details > gpurun_out/r02i_ncu_full_engine_details.txt 2>/dev/null
tail -5 gpurun_out/r02i_tests.log; tail -2 gpurun_out/r02i_smoke.log
for f in gpurun_out/r02i_bench_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:d.get(k) for k in ("value","value_solver_only","gpu_launches")}, (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("cores"))
except Exception as e: print("ERR", sys.argv[1], e)
PY
done
cat gpurun_out/r02i_engine.jsonl gpurun_out/r02i_icofoam_n128.json; tail -3 gpurun_out/r02i_icofoam.err
head -40 gpurun_out/r02i_kernels.txt
du -sh gpurun_out
