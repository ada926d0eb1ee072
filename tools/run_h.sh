mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2h_pytest.log 2>&1
python bench.py --steps 30 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
python bench.py --steps 30 --warmup 5 --workload cfg2_fwd --no-cpu-baseline > gpurun_out/r2h_bench_fwd.json 2> gpurun_out/r2h_bench_fwd.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err
