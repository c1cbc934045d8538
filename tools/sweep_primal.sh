for v in s28m3 s28m4 s64m3 s64m4 s16m3 s16m4; do
  echo "== $v"; AVBD_PRIMAL_VARIANT=$v timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print(d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['dual']['avg_launch_ms'], d['stage_ms'])
"
done
