timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ba_r1f.csv python tools/prof_ba.py 4 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/launches_ba_r1f.csv
