#!/bin/bash
# the driver's scaling sequence: bench.py at N = 1, 2, 4, 8 on one box (tools/scale_run.sh outdir [steps] [warmup])
out=$1; steps=${2:-5}; warm=${3:-3}
mkdir -p $out
python bench.py --gpus 1 --steps $steps --warmup $warm --no-configs > $out/bench_n1.json 2> $out/bench_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps $steps --warmup $warm > $out/bench_n$n.json 2> $out/bench_n$n.err
done
python - <<PY
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.loads(open('$out/bench_n%d.json'%n).read().strip().splitlines()[-1])
        base=base or d['value']
        w=d.get('weak_scaling') or {}
        print('N=%d %s value %.0f (eff %.3f) e2e %.0f ms/step %.2f weak %.0f' % (n, d['scaling'], d['value'], d['value']/(n*base), d['e2e']['value'], d['ms_per_step'], w.get('value',0)))
    except Exception as e:
        print('N=%d FAILED %r'%(n,e)); print(open('$out/bench_n%d.err'%n).read()[-800:])
PY
