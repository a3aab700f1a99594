#!/bin/bash
# One GPU-box call for the records of the training path (profiles/r02_train/): train-step lines at 1 / 2 / 4 frames per
# step in the ragged path (fp32-grade, plain TF32, bf16 autocast) and the padded cross-check path, a torch.profiler table,
# and an ncu launch list of one step (tools/train_kernel_table.py turns it into the per-kernel table).
# usage (from the repo root, on the GPU box): bash tools/run_train_records.sh TAG
tag=${1:-r02}
mkdir -p gpurun_out
for b in 1 2 4; do
  timeout 90 python benchmarks/train_step.py --steps 8 --warmup 3 --batch $b > gpurun_out/${tag}_train_ragged_fp32_b$b.json 2>/dev/null
done
timeout 90 python benchmarks/train_step.py --steps 8 --warmup 3 --batch 2 --tf32 > gpurun_out/${tag}_train_ragged_tf32_b2.json 2>/dev/null
timeout 90 python benchmarks/train_step.py --steps 8 --warmup 3 --batch 2 --amp > gpurun_out/${tag}_train_ragged_amp_b2.json 2>/dev/null
timeout 90 python benchmarks/train_step.py --steps 4 --warmup 3 --passes 1 --batch 1 --path padded > gpurun_out/${tag}_train_padded_fp32_b1.json 2>/dev/null
timeout 90 python benchmarks/train_step.py --steps 4 --warmup 3 --passes 1 --batch 1 --profile gpurun_out/${tag}_train_torch_profile.txt > /dev/null 2>&1
timeout 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/${tag}_train_launches.csv python benchmarks/train_step.py --steps 1 --warmup 2 --passes 1 > /dev/null 2>&1
cat gpurun_out/${tag}_train_*.json | cut -c1-100
