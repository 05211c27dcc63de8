#!/bin/bash
# Round-end evidence: bench line, ncu launch list of the bench command, full ncu captures of the hot
# kernels, same-GPU comparison with the reference's CUDA code.  Run on the GPU box from the repo root:
#   gpurun --timeout 1500 -- 'bash scripts/profile_round.sh r1'
R=${1:-r1}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${R}_bench_line.json 2> $O/${R}_bench.err || { echo "bench failed"; tail -5 $O/${R}_bench.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${R}_bench_launches_raw.csv \
    python bench.py --steps 5 --warmup 3 --no-attack --no-cpu-baseline > $O/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sampler_ -s 3 -c 3 -f -o $O/${R}_sampler_full \
    python scripts/run_sampler_once.py 8 2 > $O/ncu_sampler.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:allpairs_tc|lookup_fwd" -s 3 -c 2 -f -o $O/${R}_raft_full \
    python scripts/run_raft_once.py 4 2 > $O/ncu_raft.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:lookup_bwd|altcorr_fwd|altcorr_bwd|volgrad_tc" -s 2 -c 8 -f -o $O/${R}_raft_aux_full \
    python scripts/run_raft_aux_once.py > $O/ncu_raft_aux.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:lookup_convc1_kernel" -s 1 -c 1 -f -o $O/${R}_lookup_convc1_full \
    python scripts/run_lookup_convc1_once.py > $O/ncu_lc1.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:merge_grad|sampler_fwd" -s 2 -c 2 -f -o $O/${R}_merge_full \
    python scripts/run_merge_once.py > $O/ncu_merge.log 2>&1
python scripts/sweep_cfg5.py > $O/sweep_cfg5.log 2>&1; tail -2 $O/sweep_cfg5.log
ls -la $O | tail -12
