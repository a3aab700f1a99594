#!/bin/bash
# One GPU-box call that produces the round's evidence files under gpurun_out/ (copied to profiles/ afterwards):
# full GPU test suite, default bench run, ncu launch lists of one frame in the tensor-core modes
# (tools/kernel_traffic.py turns them into the per-kernel table), ncu --set full captures of the two dominant kernels.
# usage (from the repo root, on the GPU box): bash tools/run_round_records.sh TAG
tag=${1:-r02}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 300 gpurun_out/${tag}_bench.err
for p in bf16x3 tf32x3 tf32 bf16; do
  timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches_$p.csv python tools/profile_forward.py --iters 3 --precision $p > /dev/null 2>&1
done
for p in bf16x3 tf32x3; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_tca_tile -s 3 -c 1 -f -o gpurun_out/${tag}_tile_$p \
    python tools/profile_forward.py --iters 2 --precision $p > /dev/null 2>&1
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_ffn_tc -s 4 -c 1 -f -o gpurun_out/${tag}_ffn_$p \
    python tools/profile_forward.py --iters 2 --precision $p > /dev/null 2>&1
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_tcc_tile -s 1 -c 1 -f -o gpurun_out/${tag}_tcc_bf16x3 \
  python tools/profile_forward.py --iters 2 --precision bf16x3 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_block_geometry -s 1 -c 1 -f -o gpurun_out/${tag}_geo \
  python tools/profile_forward.py --iters 2 --precision bf16x3 > /dev/null 2>&1
ls -la gpurun_out/${tag}_*
