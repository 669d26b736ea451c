timeout 400 python -m pytest tests/test_fe_gpu.py -x -q 2>&1 | tail -3
timeout 120 python tools/prof_fe_phases.py 5 2>&1 | tail -4
timeout 250 python bench.py --steps 300 --warmup 20 2>/dev/null | tail -1 > gpurun_out/bench13.json; python -c "
import json; d=json.load(open('gpurun_out/bench13.json')); print(d['value'], d['e2e']['value'], d['device_ms_per_step'], d['stage_ms'], d['ba']['value'], d['ba'].get('concurrent_streams'))"
