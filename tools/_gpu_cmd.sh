timeout 600 python -m pytest tests/test_fe_gpu.py -x -q 2>&1 | tail -3
timeout 400 python bench.py 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench15.json; python -c "
import json; d=json.load(open('gpurun_out/bench15.json')); print(d['value'], d['e2e']['value'], d['device_ms_per_step'], d['device_ms_per_step_e2e'], d['cpu_baseline']['value']); print(d['stage_ms']); print(d['ba']['marginalize_old'])"
