# quick A/B (not a benchmark of record): continuous-energy history kernel, parity first, then per-cycle times
python decks/gen_decks.py > /dev/null 2>&1
python -m pytest tests/test_gpu_ce_transport.py tests/test_gpu_fixed_source.py -x -q 2>&1 | tail -2
run() { python bench.py --deck $1 --no-extras --no-cpu-baseline --steps $3 --warmup 3 --inactive 4 --pop $2 $4 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 $2 $4: %.3e n/s  %.2f ms/step  seg/s %.3e' % (d['value'], d['ms_per_step'], d['segments_per_s']))
    else: print(l.rstrip())
"; }
run ce_pin 100000 10
run ce_pin 1000000 4
run ce_asm 200000 6
run ce_asm 1250000 4
