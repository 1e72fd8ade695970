# BASELINE.json configs at their per-GPU populations, one GPU (values for the table in DESIGN.md section 7)
run() { python bench.py --deck $1 --no-extras --steps $3 --warmup 3 --inactive $4 --pop $2 --cpu-seconds 8 $5 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); c=d.get('cpu_baseline',{})
        print(json.dumps({'deck':'$1','pop':$2,'tracking':d['config']['tracking'],'neutrons_per_s':d['value'],'ms_per_cycle':d['ms_per_step'],'segments_per_s':d['segments_per_s'],'e2e':d['e2e']['value'],'keff':d['keff'],'keff_std':d['keff_std'],'cpu_keff':c.get('keff'),'keff_delta_sigma':c.get('keff_delta_in_combined_sigma'),'cpu_neutrons_per_s':c.get('value'),'cpu_cores':c.get('cores'),'cpu_pop':c.get('sample')}))
    else: print(l.rstrip())
"; }
run c5g7 100000 20 8
run inf 1000000 10 5
run slab 1000000 10 5
run ce_pin 1000000 5 3
run c5g7_3d 1250000 6 4
run ce_asm 1250000 4 3
run can 200000 10 5
