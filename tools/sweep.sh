mkdir -p gpurun_out/sweep
for c in 1 2 3 4; do ELD_WGRAD_CPS=$c python tools/profile_layers.py 8 > gpurun_out/sweep/cps$c.txt 2>&1; done
ELD_WGRAD_CPS=4 ELD_WGRAD_NOBIAS=1 python tools/profile_layers.py 8 > gpurun_out/sweep/cps4_nobias.txt 2>&1
ELD_WGRAD_CPS=1 ELD_WGRAD_NOBIAS=1 python tools/profile_layers.py 8 > gpurun_out/sweep/cps1_nobias.txt 2>&1
for l in conv9_2 conv9_1 conv8_2 conv8_1 conv7_2 conv3_1 conv2_2 conv2_1 conv1_2 conv1_1; do echo "$l.wgrad: $(for f in cps1 cps2 cps3 cps4 cps1_nobias cps4_nobias; do grep "^$l.wgrad " gpurun_out/sweep/$f.txt | awk '{printf "%s ", $2}'; done)"; done
for f in cps1 cps2 cps3 cps4 cps1_nobias cps4_nobias; do grep "^  wgrad" gpurun_out/sweep/$f.txt; done
