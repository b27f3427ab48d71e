#!/bin/bash
# Multi-GPU evidence on one box: bash tools/multi_gpu_runs.sh <N> <tag>   (bench lines of the four reference configs at N GPUs, LSH sweep at N)
N=$1; T=${2:-r2}; O=gpurun_out; P=29600
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P++)) "$@"; }
for cfg in bucket-size-64-18-06 huggingface-lsh depth-3-15-06 baseline; do
  run bench.py --gpus $N --steps 10 --warmup 3 --config $cfg --no-cpu-baseline > $O/${T}_bench_${cfg}_n$N.json 2> $O/${T}_bench_${cfg}_n$N.err; echo "$cfg N=$N rc=$?"
  python -c "
import json;d=json.load(open('$O/${T}_bench_${cfg}_n$N.json'));print(d['n_gpus'],round(d['ms_per_step'],3),round(d['value']),d['clocks']['sm_mhz'],d['clocks']['samples'],d['roofline']['frac'])"
done
run tools/sweep_lsh.py > $O/${T}_lsh_sweep_n$N.jsonl 2> $O/${T}_lsh_sweep_n$N.err; echo "sweep N=$N rc=$?"; tail -2 $O/${T}_lsh_sweep_n$N.jsonl | cut -c1-300
