#!/bin/bash
# ncu evidence of round 2 (one gpurun call): launch list of a reduced training step + --set full captures per kernel
set -u
O=gpurun_out
mkdir -p $O
python tools/profile_step.py 16384 2 > $O/r02_step_plain.log 2>&1 || { echo "plain step failed"; tail -5 $O/r02_step_plain.log; exit 1; }
python tools/rns_one.py > $O/r02_rns_one_plain.log 2>&1 || { echo "plain rns_one failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file $O/r02_launches_n16384.csv \
    python tools/profile_step.py 16384 2 > $O/r02_launches.log 2>&1
for k in rns_gemm_kernel crt_kernel residue_kc_kernel residue_mc_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $O/r02_$k python tools/rns_one.py > $O/r02_ncu_$k.log 2>&1
done
for k in gram_kernel grad_sweep_kernel project_fwd_kernel project_bwd_kernel col_reduce_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $O/r02_$k python tools/profile_step.py 16384 2 > $O/r02_ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:trsv_step_kernel -s 40 -c 1 -f -o $O/r02_trsv_step_kernel python tools/profile_step.py 16384 2 > $O/r02_ncu_trsv.log 2>&1
ls -la $O/*.ncu-rep | awk '{print $5, $9}'
