# ncu evidence for profiles/: (1) SpeedOfLight + memory sections for every launch of one step, (2) --set full of the
# kernels the rooflines are quoted on.  usage: gpurun --timeout 1200 -- 'bash tools/ncu_round.sh [full]'
O=gpurun_out/ncu; mkdir -p $O
timeout 500 ncu --profile-from-start off --clock-control none --section SpeedOfLight --section MemoryWorkloadAnalysis \
    --section LaunchStats --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --csv --log-file $O/step_sol.csv python tools/ncu_step.py > $O/step_sol.log 2>&1
if [ "$1" = "full" ]; then
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_umma_kernel --launch-skip 2 --launch-count 1 \
    -o $O/conv1_2_fprop -f python tools/ncu_step.py > $O/full1.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:noise_packed_poisson --launch-count 1 \
    -o $O/noise_Pg -f python tools/ncu_step.py > $O/full2.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wgrad_conv_kernel --launch-skip 10 --launch-count 1 \
    -o $O/wgrad_mid -f python tools/ncu_step.py > $O/full3.log 2>&1
fi
ls -la $O; tail -2 $O/step_sol.log
