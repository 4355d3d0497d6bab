python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_n8.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n8.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['parallelism'])"
