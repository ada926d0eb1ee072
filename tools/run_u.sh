mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/r2u_bench_n8.json 2> gpurun_out/r2u_bench_n8.err
timeout 200 python bench.py --steps 40 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_bench_n1.err
python - <<'PY'
import json
for f in ["n8","n1"]:
    try:
        j=json.loads(open(f"gpurun_out/r2u_bench_{f}.json").read().strip().splitlines()[-1]); print(f, j["ms_per_step"], j["value"], j["e2e"]["value"], j.get("grad_sync_check"), j.get("cfg4_train",{}).get("ms_per_step"))
    except Exception as e: print(f,"ERR",e)
PY
