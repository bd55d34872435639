set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r01f_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE-OK')" > gpurun_out/r01f_smoke.log 2>&1
python bench.py > gpurun_out/r01f_bench_n1.json 2> gpurun_out/r01f_bench_n1.err
B200LDU_TAIL=0 python bench.py --no-cpu-baseline > gpurun_out/r01f_bench_n1_notail.json 2>/dev/null
python bench.py --n 128 --no-cpu-baseline > gpurun_out/r01f_bench_n128.json 2>/dev/null
B200LDU_TAIL=0 python bench.py --n 128 --no-cpu-baseline > gpurun_out/r01f_bench_n128_notail.json 2>/dev/null
python bench_kernels.py > gpurun_out/r01f_kernels.json 2> gpurun_out/r01f_kernels.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r01f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01f_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:engine_kernel -s 24 -c 4 -f -o gpurun_out/r01f_engine python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01f_ncu_full.log 2>&1
cat gpurun_out/r01f_tests.log gpurun_out/r01f_smoke.log | tail -12
cat gpurun_out/r01f_bench_n1.json gpurun_out/r01f_bench_n1_notail.json gpurun_out/r01f_bench_n128.json gpurun_out/r01f_bench_n128_notail.json | cut -c1-400
