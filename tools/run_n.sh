mkdir -p gpurun_out
port=29540
for c in 8 24 32; do
port=$((port+1))
NCCL_MAX_CTAS=$c UNIVTG_DDP_SM_RESERVE=$c timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --steps 30 --warmup 5 --no-extras > gpurun_out/r2n_n8_cta$c.json 2> gpurun_out/r2n_n8_cta$c.err
done
port=$((port+1))
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --steps 30 --warmup 5 --no-extras > gpurun_out/r2n_n8_cta16.json 2> gpurun_out/r2n_n8_cta16.err
grep -m5 -i "nvls\|Connected all\|channels" gpurun_out/r2n_n8_cta16.err | cut -c1-200
python - <<'PY'
import json
for c in [8,16,24,32]:
    try:
        j=json.loads(open(f"gpurun_out/r2n_n8_cta{c}.json").read().strip().splitlines()[-1]); print(c, j["ms_per_step"], j["value"])
    except Exception as e: print(c,"ERR",e)
PY
