#!/bin/bash
# round 2, job 5 (2 GPUs): the driver's multi-GPU launch line of bench.py + the gloo/NCCL rollout tool
mkdir -p gpurun_out
N=${1:-2}
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r2j5_bench_n$N.json 2> gpurun_out/r2j5_bench_n$N.err
tail -4 gpurun_out/r2j5_bench_n$N.err
python - <<PY
import json
try:
    d = [json.loads(l) for l in open('gpurun_out/r2j5_bench_n$N.json') if l.startswith('{')][0]
    print('N', d['n_gpus'], 'value %.3f M' % (d['value']/1e6), 'steady %.3f M' % (d['steady_state']['value']/1e6), 'e2e %.3f M' % (d['e2e']['value']/1e6))
    for k, r in d.get('secondary', {}).items():
        if isinstance(r, dict): print(k, 'value %.3f M steady %.3f M e2e %.3f M' % (r['value']/1e6, r['steady_state']['value']/1e6, r['e2e']['value']/1e6))
    print('sweep', d['sweep']['ms_per_pass'], d['sweep']['value'])
    print('rollout', json.dumps(d['rollout']))
    print('gate', d['parity_gate']['ok'])
except Exception as ex:
    print('ERR', ex)
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/multi_gpu_rollout.py 4096 tower > gpurun_out/r2j5_rollout_n$N.json 2> gpurun_out/r2j5_rollout_n$N.err
tail -1 gpurun_out/r2j5_rollout_n$N.json
