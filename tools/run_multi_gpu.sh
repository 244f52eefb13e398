#!/bin/bash
# the driver's multi-GPU launch line of bench.py at N GPUs (all GPUs of the box)
mkdir -p gpurun_out
N=${1:-8}
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r2e_bench_n$N.json 2> gpurun_out/r2e_bench_n$N.err
tail -4 gpurun_out/r2e_bench_n$N.err
python - <<PY
import json
try:
    d = [json.loads(l) for l in open('gpurun_out/r2e_bench_n$N.json') if l.startswith('{')][0]
    print('N', d['n_gpus'], 'value %.3f M' % (d['value']/1e6), 'steady %.3f M' % (d['steady_state']['value']/1e6), 'e2e %.3f M' % (d['e2e']['value']/1e6), 'clocks', d['clocks'])
    for k, r in d.get('secondary', {}).items():
        if isinstance(r, dict): print(k, 'value %.3f M steady %.3f M e2e %.3f M' % (r['value']/1e6, r['steady_state']['value']/1e6, r['e2e']['value']/1e6))
    print('sweep', d['sweep']['ms_per_pass'], d['sweep']['value'])
    print('rollout %.3f M per gpu %.3f M' % (d['rollout']['value']/1e6, d['rollout']['per_gpu']/1e6), d['rollout'].get('replay_gather'))
    print('gate', d['parity_gate']['ok'])
except Exception as ex:
    print('ERR', ex)
PY
