# quick A/B of the history kernel configurations (not a benchmark of record)
python -m pytest tests/test_gpu_ce_transport.py -x -q 2>&1 | tail -2
python -m pytest tests/test_gpu_eigen.py -x -q -k "bit_exact" 2>&1 | tail -2
for cfg in sync512; do
SB_CE_KERNEL=$cfg python bench.py --deck ce_pin --no-extras --no-cpu-baseline --steps 6 --warmup 3 --inactive 4 ${CE_POP:+--pop $CE_POP} 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('CE $cfg: %.3e n/s  %.2f ms/step  seg/s %.3e  kernel %.2f ms longest %d' % (d['value'], d['ms_per_step'], d['segments_per_s'], d['roofline']['kernel_ms_per_launch'], d['longest_history_segments']))
    else: print(l.rstrip())
"
done
for tr in ST HT; do
SB_FORCE_TRACK_KERNEL=1 python bench.py --deck c5g7 --tracking $tr --no-extras --no-cpu-baseline --steps 8 --warmup 3 --inactive 4 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('C5G7 $tr: %.3e n/s  %.2f ms/step  seg/s %.3e' % (d['value'], d['ms_per_step'], d['segments_per_s']))
    else: print(l.rstrip())
"
done
