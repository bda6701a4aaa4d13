# sweep of the exact-NN grid parameters (h = factor * sqrt(area / N), hx = xfactor * h); prints value, ms/step, e2e, launches, per-kernel ms
for xf in ${XFACTORS:-0.34}; do for f in ${FACTORS:-2.0}; do echo "factor $f xfactor $xf"; B2R_NN_XCELL_FACTOR=$xf B2R_NN_CELL_FACTOR=$f python bench.py --steps 20 --no-cpu-baseline ${BENCH_ARGS:-} 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches'], {k: round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items() if v})
"; done; done
