# round 2, call J (1 GPU): MULES on the device; band-size rule at the per-rank size of a 4-way split (161^3 ~ 4.2 M cells)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mules.py tests/test_gpu_limiters.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r02j_tests.log
for br in 0 1024 2048; do
  if [ $br = 0 ]; then unset B200LDU_BAND_ROWS; else export B200LDU_BAND_ROWS=$br; fi
  timeout 300 python bench.py --n 161 --no-cpu-baseline --no-secondary --no-parity > gpurun_out/r02j_bench_n161_band$br.json 2>/dev/null
done
unset B200LDU_BAND_ROWS
timeout 300 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/r02j_bench_n1.json 2>/dev/null
cat gpurun_out/r02j_tests.log
for f in gpurun_out/r02j_bench_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:d.get(k) for k in ("value","value_solver_only")}, d.get("layout"))
except Exception as e: print("ERR", sys.argv[1], e)
PY
done
