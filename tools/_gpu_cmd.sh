timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 python bench.py 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench12.json; python -c "
import json; d=json.load(open('gpurun_out/bench12.json')); print(d['value'], d['e2e']['value'], d['device_ms_per_step'], d['cpu_baseline'], d['ba'], d['clocks'])"
