#!/bin/bash
# run the quick bench for several library variants (built with PSFR_LIB_TAG): tools/variants.sh outdir tag1 tag2 ...
out=$1; shift
mkdir -p $out
for tag in "$@"; do
  if [ "$tag" = "default" ]; then unset PSFR_LIB_TAG; else export PSFR_LIB_TAG=$tag; fi
  python bench.py --steps 3 --warmup 2 --no-cpu --no-configs --draws 2048 > $out/bench_$tag.json 2> $out/bench_$tag.err
  python - <<PY
import json
try:
    d=json.load(open('$out/bench_$tag.json'))
    print('$tag', 'value %.0f  row kernel %.3f ms  allfp64 %.0f' % (d['value'], d['roofline']['avg_launch_ms'], d.get('value_allfp64') or 0))
except Exception as e:
    print('$tag', 'FAILED', e); print(open('$out/bench_$tag.err').read()[-500:])
PY
done
