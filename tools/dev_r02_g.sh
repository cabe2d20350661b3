for v in 0 1; do
HSDDP_SWEEP_KIND=$v ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv \
  --log-file gpurun_out/r02g_launches_kind$v.csv python tools/profile_case.py 16384 config3 2 > gpurun_out/ncu_r02g_$v.log 2>&1
tail -1 gpurun_out/ncu_r02g_$v.log | cut -c1-200
done
