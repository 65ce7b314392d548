#!/bin/bash
# ncu evidence for the C3 pass: (1) every launch of the second pass with its device time, (2) one full capture of each
# hot kernel.  All inside one gpurun call (ncu first runs the command once without ncu).
set -x
python profiles/prof_path.py c3 2 > gpurun_out/ncu_plain.log 2>&1 || exit 1
cat gpurun_out/ncu_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r02_launches_c3.csv python profiles/prof_path.py c3 2 > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
export ADOBO_B200_IRLB=no_graph
for K in w_step_t v_step_t rotate_ knn_scores_tc csr_gene_moments csr_libsize_scale_log; do
  S=1; C=1
  case $K in w_step_t|v_step_t) S=700; C=2;; rotate_) S=160; C=1;; esac
  ncu --set full --clock-control none --import-source on -k regex:^$K -s $S -c $C -f -o gpurun_out/r02_$K python profiles/prof_path.py c3 2 > gpurun_out/ncu_$K.log 2>&1
  tail -1 gpurun_out/ncu_$K.log
done
ls -la gpurun_out/
