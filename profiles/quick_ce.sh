# quick checks (not a benchmark of record)
python -m pytest tests/test_gpu_distributed.py -x -q 2>&1 | tail -3
CE_POP=1000000
SB_CE_KERNEL=sync512 python bench.py --deck ce_pin --no-extras --no-cpu-baseline --steps 4 --warmup 3 --inactive 3 --pop 1000000 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('CE 1e6: %.3e n/s  %.2f ms/step  seg/s %.3e  kernel %.2f ms longest %d' % (d['value'], d['ms_per_step'], d['segments_per_s'], d['roofline']['kernel_ms_per_launch'], d['longest_history_segments']))
    else: print(l.rstrip())
"
