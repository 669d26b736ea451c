timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_ba_r1g.csv python tools/prof_ba.py 2 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/launches_ba_r1g.csv
GF_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_fe_r1g.csv python tools/prof_fe.py 12 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/launches_fe_r1g.csv
