#!/bin/bash
# one call: GPU tests, default bench (C3, N=1) + reference arm, launch list, full captures of the hot kernels
TAG=${1:-r02f}
python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_gputest.log 2>&1
echo "gputest rc=$?"; tail -2 gpurun_out/${TAG}_gputest.log
python bench.py > gpurun_out/${TAG}_bench_c3_n1.json 2> gpurun_out/${TAG}_bench_c3_n1.err
echo "bench rc=$?"; head -c 1500 gpurun_out/${TAG}_bench_c3_n1.json; tail -3 gpurun_out/${TAG}_bench_c3_n1.err
python bench.py --impl reference > gpurun_out/${TAG}_bench_c3_reference_arm.json 2> gpurun_out/${TAG}_ref.err
echo "ref rc=$?"; head -c 600 gpurun_out/${TAG}_bench_c3_reference_arm.json
python profiles/prof_path.py c3 2 > gpurun_out/ncu_plain.log 2>&1 || exit 1
tail -5 gpurun_out/ncu_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/${TAG}_launches_c3.csv python profiles/prof_path.py c3 2 > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
export ADOBO_B200_IRLB=no_graph
for K in w_step_t v_step_t rotate_mma knn_scores_tc gene_moments_tiles csr_libsize_scale_log; do
  S=1; C=1
  case $K in w_step_t|v_step_t) S=700; C=2;; rotate_mma) S=60; C=1;; esac
  ncu --set full --clock-control none --import-source on -k regex:^$K -s $S -c $C -f -o gpurun_out/${TAG}_$K python profiles/prof_path.py c3 2 > gpurun_out/ncu_$K.log 2>&1
  tail -1 gpurun_out/ncu_$K.log
done
ls -la gpurun_out/
