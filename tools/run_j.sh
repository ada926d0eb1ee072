mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err
UNIVTG_DDP_SM_RESERVE=0 NCCL_MAX_CTAS=32 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras > gpurun_out/r2j_bench_n2_noreserve.json 2> gpurun_out/r2j_bench_n2_noreserve.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --no-overlap > gpurun_out/r2j_bench_n2_nooverlap.json 2> gpurun_out/r2j_bench_n2_nooverlap.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --train-graph-probe > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err
