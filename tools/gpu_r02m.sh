# round 2, call M (1 GPU): the face-sum / fvMatrix / MULES kernels with hoisted load scheduling -- parity on the device, then the
# table; a short bench for the NVML clock sampler
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02m_tests.log
timeout 400 python bench_kernels.py > gpurun_out/r02m_kernels.json 2> gpurun_out/r02m_kernels.txt
timeout 200 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/r02m_bench_n1.json 2> gpurun_out/r02m_bench_n1.err
cat gpurun_out/r02m_tests.log; cat gpurun_out/r02m_kernels.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02m_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","value_solver_only","clocks","layout")}, d["e2e"]["value"], d["roofline"]["traffic"], d["roofline"].get("traffic_source"))
PY
tail -3 gpurun_out/r02m_bench_n1.err
