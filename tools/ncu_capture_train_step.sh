#!/bin/bash
# ncu captures of one cfg3_train step (run on the GPU box: gpurun -- bash tools/ncu_capture_train_step.sh).  Numbers printed by bench.py
# under ncu are never bench values; the reports are summarised on the box because gpurun_out/ is limited to 64 MiB.
mkdir -p gpurun_out
T=/tmp/ncu_r2
mkdir -p $T
ncu --set full --clock-control none --import-source on -k regex:'attention|layernorm|conv_head_final|adamw|sumsq|colsum16|attn_delta' -s 135 -c 45 -f -o $T/rows python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2f_ncu_rows.log 2>&1
python tools/ncu_summary.py $T/rows.ncu-rep gpurun_out/r2f_rowkernels_ncu_full_summary.csv >> gpurun_out/r2f_ncu_rows.log 2>&1
ncu -i $T/rows.ncu-rep --page source --csv -k regex:attention_bwd > gpurun_out/r2f_attention_bwd_source.csv 2>/dev/null
ncu -i $T/rows.ncu-rep --page source --csv -k regex:layernorm_bwd_vec > gpurun_out/r2f_layernorm_bwd_source.csv 2>/dev/null
ls -la $T >> gpurun_out/r2f_ncu_rows.log
rm -f $T/rows.ncu-rep
ncu --set full --clock-control none -k regex:gemm_tcgen05 -s 174 -c 58 -f -o $T/gemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2f_ncu_gemm.log 2>&1
python tools/ncu_summary.py $T/gemm.ncu-rep gpurun_out/r2f_gemm_ncu_full_summary.csv >> gpurun_out/r2f_ncu_gemm.log 2>&1
rm -f $T/gemm.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/r2f_launches_cfg3_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2f_ncu_launches.log 2>&1
du -sh gpurun_out >> gpurun_out/r2f_ncu_rows.log
