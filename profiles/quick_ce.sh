# quick A/B (not a benchmark of record): more warps per SM at fewer registers per thread
for cfg in sync768 sync1024; do
for deck in ce_pin; do
for pop in 100000 1000000; do
SB_CE_KERNEL=$cfg python bench.py --deck $deck --no-extras --no-cpu-baseline --steps 4 --warmup 3 --inactive 3 --pop $pop 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$cfg $deck $pop: %.3e n/s  %.2f ms/step  seg/s %.3e  kernel %.2f ms longest %d k %.5f' % (d['value'], d['ms_per_step'], d['segments_per_s'], d['roofline']['kernel_ms_per_launch'], d['longest_history_segments'], d['keff']))
    else: print(l.rstrip())
"
done; done; done
