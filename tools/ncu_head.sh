# --set full captures of the kernels that are furthest from their floor: usage: gpurun --timeout 900 -- 'bash tools/ncu_head.sh'
O=gpurun_out/ncu2; mkdir -p $O
for k in head_kernel first_conv_fprop_kernel first_conv_wgrad_kernel maxpool_bwd_kernel; do
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k --launch-count 1 \
    -o $O/$k -f python tools/ncu_step.py > $O/$k.log 2>&1
done
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_umma_kernel --launch-skip 0 --launch-count 1 \
    -o $O/conv1_2_fprop -f python tools/ncu_step.py > $O/conv1_2.log 2>&1
ls -la $O
