GF_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_track -s 4 -c 1 -o gpurun_out/k_track_r1 -f python tools/prof_fe.py 8 > /dev/null 2>&1
GF_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_fe_r1f.csv python tools/prof_fe.py 12 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/launches_fe_r1f.csv
ls -la gpurun_out/k_track_r1.ncu-rep
