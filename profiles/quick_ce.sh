# quick A/B of the CE history kernel configurations (not a benchmark of record)
python -m pytest tests/test_gpu_ce_transport.py -x -q 2>&1 | tail -2
for cfg in sync512 sync256 async; do
SB_CE_KERNEL=$cfg python bench.py --deck ce_pin --no-extras --no-cpu-baseline --steps 8 --warmup 3 --inactive 4 ${CE_POP:+--pop $CE_POP} 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('CE $cfg: %.3e n/s  %.2f ms/step  seg/s %.3e  kernel %.2f ms' % (d['value'], d['ms_per_step'], d['segments_per_s'], d['roofline']['kernel_ms_per_launch']))
    else: print(l.rstrip())
"
done
