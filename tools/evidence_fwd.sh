T=r2b; O=gpurun_out
python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline > $O/${T}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline > $O/${T}_ncu_launch.log 2>&1; echo "launch list rc=$?"
python tools/run_attn.py > $O/${T}_run_attn.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"lsh_attn_fwd|lsh_attn_bwd|lsh_merge|lsh_hash_mma" -s 8 -c 4 -f -o $O/${T}_attn python tools/run_attn.py > $O/${T}_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
python tools/ab_fwd.py > $O/${T}_ab_fwd.txt 2>&1; echo "ab rc=$?"
cat $O/${T}_run_attn.txt
