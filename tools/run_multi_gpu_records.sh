set -x
mkdir -p gpurun_out/n8
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $T --nproc-per-node 8 --master-port 29601 bench.py --gpus 8 --no-modes > gpurun_out/n8/r02_bench_n8.json 2> gpurun_out/n8/bench_n8.err
timeout 300 $T --nproc-per-node 4 --master-port 29602 bench.py --gpus 4 --no-modes > gpurun_out/n8/r02_bench_n4.json 2> gpurun_out/n8/bench_n4.err
timeout 300 $T --nproc-per-node 8 --master-port 29603 benchmarks/train_step.py --amp --steps 8 > gpurun_out/n8/r02_train_step_8gpu_amp.json 2> gpurun_out/n8/train8.err
timeout 200 $T --nproc-per-node 2 --master-port 29604 benchmarks/train_step.py --amp --steps 8 > gpurun_out/n8/r02_train_step_2gpu_amp.json 2> gpurun_out/n8/train2.err
for n in 2 4 8; do timeout 200 $T --nproc-per-node $n --master-port 2961$n benchmarks/shard_frame.py --steps 8 > gpurun_out/n8/r02_shard_frame_${n}gpu.json 2> gpurun_out/n8/shard$n.err; done
lscpu | grep -E "^CPU\(s\)|Socket|NUMA" > gpurun_out/n8/host.txt; nvidia-smi topo -m >> gpurun_out/n8/host.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/n8/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    if "e2e" in d: print(f, d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "pts", d["e2e_points"]["ms_per_step"], d["host_link"])
    else: print(f, json.dumps(d)[:400])
PY
