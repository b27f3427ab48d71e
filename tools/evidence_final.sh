#!/bin/bash
# Final evidence of the round in one GPU call: tests, bench (+ reference arm), launch list of an eager step, ncu --set full of the cross-attention kernels.
T=${1:-r2c}; O=gpurun_out
python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${T}_pytest.log; tail -2 $O/${T}_pytest.log
python bench.py --steps 20 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "reference rc=$?"
python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline > $O/${T}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline > $O/${T}_ncu_launch.log 2>&1; echo "launch list rc=$?"
python tools/time_xattn.py > $O/${T}_time_xattn.txt 2>&1 && \
ncu --set full --clock-control none -k regex:xattn -s 12 -c 1 -f -o $O/${T}_xattn_fwd python tools/time_xattn.py x > $O/${T}_ncu_xf.log 2>&1; echo "ncu xattn fwd rc=$?"
ncu --set full --clock-control none -k regex:xattn_bwd -s 6 -c 1 -f -o $O/${T}_xattn_bwd python tools/time_xattn.py x > $O/${T}_ncu_xb.log 2>&1; echo "ncu xattn bwd rc=$?"
cat $O/${T}_time_xattn.txt
