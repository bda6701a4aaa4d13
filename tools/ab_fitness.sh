#!/bin/bash
# A/B runs of the headline step: per-kernel ms under switches (one line per variant).  Usage: tools/ab_fitness.sh "ENV=.. ENV=.." ...
mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v"
  env $v python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --no-extras ${BENCH_ARGS:-} 2>gpurun_out/ab_last.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches'], {k: round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items() if v})
"
done
