# final 1-GPU evidence run: tests, smoke, bench (+ reference arm), ncu launch list, ncu full capture
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r01z_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE-OK')" > gpurun_out/r01z_smoke.log 2>&1
python bench.py > gpurun_out/r01z_bench_n1.json 2> gpurun_out/r01z_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01z_bench_ref.json 2> gpurun_out/r01z_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r01z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01z_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:engine_kernel -c 10 -f -o gpurun_out/r01z_engine python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01z_ncu_full.log 2>&1
cat gpurun_out/r01z_tests.log gpurun_out/r01z_smoke.log | tail -14
cut -c1-300 gpurun_out/r01z_bench_n1.json gpurun_out/r01z_bench_ref.json
