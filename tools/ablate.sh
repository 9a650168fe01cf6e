mkdir -p gpurun_out/abl
for d in 0 1 8 16 9 17 24 25; do ELD_CONV_DBG=$d python tools/profile_layers.py 8 > gpurun_out/abl/layers_dbg$d.txt 2>&1; done
