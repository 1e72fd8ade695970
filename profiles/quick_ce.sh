# quick A/B (not a benchmark of record): warp-aggregated tally atomics
python -m pytest tests/test_gpu_eigen.py tests/test_gpu_ce_transport.py -x -q 2>&1 | tail -2
run() { python bench.py --deck $1 --no-extras --no-cpu-baseline --steps $3 --warmup 3 --inactive 4 --pop $2 $4 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 $2 $4: %.3e n/s  %.2f ms/step  seg/s %.3e' % (d['value'], d['ms_per_step'], d['segments_per_s']))
    else: print(l.rstrip())
"; }
run c5g7 100000 20
run c5g7 1000000 8
run inf 1000000 8
run slab 1000000 8
run c5g7_3d 1250000 6
run ce_pin 1000000 4
