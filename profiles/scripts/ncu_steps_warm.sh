export ADOBO_B200_IRLB=no_graph
for K in w_step_t v_step_t; do
  ncu --set full --clock-control none --cache-control none --import-source on -k regex:^$K -s 700 -c 2 -f -o gpurun_out/r02e_$K python profiles/prof_path.py c3 2 > gpurun_out/ncu_$K.log 2>&1
  tail -1 gpurun_out/ncu_$K.log
done
