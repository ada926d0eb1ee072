mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_data_gpu.py tests/test_train_gpu.py -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --train-graph-probe > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
UNIVTG_PDL=0 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras --train-graph-probe > gpurun_out/r2i_bench_pdl0.json 2> gpurun_out/r2i_bench_pdl0.err
