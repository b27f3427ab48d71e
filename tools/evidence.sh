#!/bin/bash
# Round evidence in one GPU call: tests, bench, ncu launch list, ncu --set full of the attention / HBM-bound kernels, sanitizer logs.
# usage (on the GPU box): bash tools/evidence.sh <tag>     outputs: gpurun_out/<tag>_*
T=${1:-r2}; O=gpurun_out
python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${T}_pytest.log
python bench.py --steps 10 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "reference rc=$?"
# launch list of one eager step (ncu serialises and runs cold: compare shares, not absolutes)
python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline > $O/${T}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline > $O/${T}_ncu_launch.log 2>&1; echo "launch list rc=$?"
# ncu --set full: attention kernels (3rd iteration), then every other kernel of the library once at the default-workload shape
python tools/run_attn.py > $O/${T}_run_attn.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"lsh_attn_fwd|lsh_attn_bwd|lsh_merge|lsh_hash_mma" -s 8 -c 4 -o $O/${T}_attn python tools/run_attn.py > $O/${T}_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
python tools/run_all_kernels.py full 3 > $O/${T}_run_all.txt 2>&1 && \
ncu --set full --clock-control none -k regex:"lsh_sort|lsh_hash_kernel|lsh_sumsq|lsh_delta|lsh_grad_reduce|layernorm|cast_bf16" -s 14 -c 9 -o $O/${T}_hbm python tools/run_all_kernels.py full 3 > $O/${T}_ncu_hbm.log 2>&1; echo "ncu hbm rc=$?"
# race / sync checks at tiny shapes (every kernel once)
python tools/run_all_kernels.py tiny > $O/${T}_tiny.txt 2>&1 && {
  timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/run_all_kernels.py tiny > $O/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"
  timeout 900 compute-sanitizer --tool synccheck python tools/run_all_kernels.py tiny > $O/${T}_synccheck.log 2>&1; echo "synccheck rc=$?"
  timeout 900 compute-sanitizer --tool memcheck python tools/run_all_kernels.py tiny > $O/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"
}
tail -3 $O/${T}_racecheck.log $O/${T}_synccheck.log $O/${T}_memcheck.log
