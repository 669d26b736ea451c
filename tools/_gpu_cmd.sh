timeout 400 python -m pytest tests/test_fe_gpu.py -x -q 2>&1 | tail -5
timeout 120 python tools/prof_fe_phases.py 4 2>&1 | tail -7
timeout 250 python bench.py --steps 300 --warmup 20 2>&1 | tail -1 > gpurun_out/bench11.json; python -c "
import json; d=json.load(open('gpurun_out/bench11.json')); print(d['value'], d['e2e']['value'], d['device_ms_per_step'], d['device_ms_per_step_e2e'], d['stage_ms'])"
