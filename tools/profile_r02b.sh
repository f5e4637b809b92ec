#!/bin/bash
# ncu evidence of round 2, final kernels (one gpurun call): launch list of a reduced training step + prediction,
# --set full captures of the kernels that changed late in the round.
set -u
O=gpurun_out
mkdir -p $O
python tools/profile_step.py 16384 2 > $O/r02b_step_plain.log 2>&1 || { echo "plain step failed"; tail -5 $O/r02b_step_plain.log; exit 1; }
python tools/tri_one.py 16384 12 > $O/r02b_tri_plain.log 2>&1 || { echo "plain tri_one failed"; tail -5 $O/r02b_tri_plain.log; exit 1; }
python tools/gram_bench.py 16384 21 2 1 > $O/r02b_gram_plain.log 2>&1 || { echo "plain gram_bench failed"; exit 1; }
cat $O/r02b_tri_plain.log $O/r02b_gram_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file $O/r02b_launches_n16384.csv \
    python tools/profile_step.py 16384 2 > $O/r02b_launches.log 2>&1
# the triangular L^T L product: first rns_gemm launch of tri_one.py (one launch set, k-range per tile)
ncu --set full --clock-control none --import-source on -k regex:rns_gemm_kernel -c 1 -f -o $O/r02b_rns_gemm_tri python tools/tri_one.py 16384 12 > $O/r02b_ncu_rns_tri.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:residue_mc_kernel -c 1 -f -o $O/r02b_residue_mc_tri python tools/tri_one.py 16384 12 > $O/r02b_ncu_resmc_tri.log 2>&1
for k in gram_kernel grad_sweep2_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $O/r02b_$k python tools/gram_bench.py 16384 21 2 1 > $O/r02b_ncu_$k.log 2>&1
done
ls -la $O/r02b_*.ncu-rep | awk '{print $5, $9}'
