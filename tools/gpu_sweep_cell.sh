# sweep of the exact-NN grid cell size factor (h = factor * sqrt(area / N)); prints value, ms/step, e2e, launches, per-kernel ms
for f in ${FACTORS:-2.0 2.5 3.0 4.0}; do echo "factor $f"; B2R_NN_CELL_FACTOR=$f python bench.py --steps 5 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel_ms_per_step'])
"; done
