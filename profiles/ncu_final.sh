#!/bin/bash
# The round's ncu evidence for the final code (run through gpurun, one GPU): launch list of the bench command, then one
# `--set full` capture of the launches of one rna2dna train step at batch 4096 (eager launches: profiles/run_steps.py).
set -x
timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu --no-also > gpurun_out/ncu_final_plain.json 2> gpurun_out/ncu_final_plain.err || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final2.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-also > gpurun_out/ncu_ll2.log 2>&1
echo launch list exit $?
timeout 60 python profiles/run_steps.py rna2dna 4096 6 || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:gemm_tc|adamw|bn_act|bn_bwd|ingest|latent_' --launch-skip 51 --launch-count 18 -f \
  -o gpurun_out/prof_step_final2 python profiles/run_steps.py rna2dna 4096 6 > gpurun_out/ncu_step_final2.log 2>&1
echo full exit $?
ls -la gpurun_out/prof_step_final2.ncu-rep
