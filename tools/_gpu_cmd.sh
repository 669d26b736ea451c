timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench18.json; python -c "
import json; d=json.load(open('gpurun_out/bench18.json')); print(d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['ba']['value'], d['ba']['cpu_baseline']['value'], d['ba']['marginalize_old'], d['ba']['concurrent_streams']['value'])"
