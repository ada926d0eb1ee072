mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2g_bench_new.json 2> gpurun_out/r2g_bench_new.err
UNIVTG_MERGE_BWD=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2g_bench_nomerge.json 2> gpurun_out/r2g_bench_nomerge.err
UNIVTG_LNB_WARP=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2g_bench_oldlnb.json 2> gpurun_out/r2g_bench_oldlnb.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2g_bench_new2.json 2> gpurun_out/r2g_bench_new2.err
UNIVTG_PDL=0 python tools/timeline.py train cfg2 > gpurun_out/r2g_timeline_train_pdl0.txt 2>&1
