mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2k_pytest.log 2>&1
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2k_bench_new.json 2> gpurun_out/r2k_bench_new.err
UNIVTG_EPI_COLSUM=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2k_bench_epicolsum.json 2> gpurun_out/r2k_bench_epicolsum.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras --static-loss-scale > gpurun_out/r2k_bench_static.json 2> gpurun_out/r2k_bench_static.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2k_bench_new2.json 2> gpurun_out/r2k_bench_new2.err
python bench.py --steps 30 --warmup 5 --workload cfg2_fwd --no-cpu-baseline > gpurun_out/r2k_bench_fwd.json 2> gpurun_out/r2k_bench_fwd.err
