timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ba_step -s 3 -c 1 -o gpurun_out/k_ba_step_r1b -f python tools/prof_ba.py 1 > /dev/null 2>&1
ls -la gpurun_out/k_ba_step_r1b.ncu-rep
