# quick checks (not a benchmark of record)
python -m pytest tests/test_gpu_ce_transport.py -x -q -k assembly 2>&1 | tail -3
for pop in 100000 1000000; do
python bench.py --deck ce_asm --no-extras --no-cpu-baseline --steps 4 --warmup 3 --inactive 3 --pop $pop 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('CE asm $pop: %.3e n/s  %.2f ms/step  seg/s %.3e  kernel %.2f ms longest %d k %.5f' % (d['value'], d['ms_per_step'], d['segments_per_s'], d['roofline']['kernel_ms_per_launch'], d['longest_history_segments'], d['keff']))
    else: print(l.rstrip())
"
done
