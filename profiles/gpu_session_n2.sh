#!/bin/bash
# 2-GPU session: weak scaling line, shared-map line (NCCL map-delta broadcast every step), strong-scaling re-render, NCCL check
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.txt
run() { tag=$1; shift; ( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 "$@" ) > gpurun_out/n2_$tag.json 2> gpurun_out/n2_$tag.err; echo "$tag exit $?"; tail -2 gpurun_out/n2_$tag.err | head -1; }
PORT=29511 run weak --steps 6 --warmup 3
PORT=29512 run share --steps 6 --warmup 3 --share-map
PORT=29513 run rerender --config rerender --steps 4 --warmup 3
( time timeout 200 python bench.py --config rerender --steps 4 --warmup 3 --no-cpu-baseline ) > gpurun_out/n2_rerender_n1.json 2> gpurun_out/n2_rerender_n1.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tests/multigpu_check.py ) > gpurun_out/n2_multigpu_check.log 2>&1
echo "multigpu_check exit $?"; tail -4 gpurun_out/n2_multigpu_check.log
python - <<'PY'
import json
for t in ('weak','share','rerender','rerender_n1'):
    try:
        d = json.loads([l for l in open(f'gpurun_out/n2_{t}.json').read().splitlines() if l.startswith('{')][-1])
        print(t, 'N', d['n_gpus'], round(d['ms_per_step'], 2), 'ms/step', round(d['value']/1e6,2), 'M/s', d['timing']['per_rank_ms_per_step'], d['config']['parallelism'], d.get('rerender'))
    except Exception as e:
        print(t, 'no line', e)
PY
