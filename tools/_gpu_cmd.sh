timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 2>gpurun_out/scale2.err | tail -1 > gpurun_out/bench_n2.json
python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print({k:d[k] for k in ('value','n_gpus','ms_per_step','device_ms_per_step','scaling')}, d['e2e'])"
tail -3 gpurun_out/scale2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
