bash profiles/capture.sh r02 100000
ncu --set full --clock-control none -k regex:k_ce_lookup -s 3 -c 1 -f -o gpurun_out/prof_celarge_r02c python profiles/run_ce_lookup_large.py 10000000 300 > gpurun_out/prof_celarge_r02c.log 2>&1
ncu --set full --clock-control none -k regex:"k_ce_lookup|k_ce_sort|k_ce_bin|k_ce_" -s 12 -c 6 -f -o gpurun_out/prof_celarge_sorted_r02c python profiles/run_ce_lookup_large.py 10000000 300 engine-sort > gpurun_out/prof_celarge_sorted_r02c.log 2>&1
python bench.py > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err
python bench.py --impl reference > gpurun_out/bench_ref_r02c.json 2>/dev/null
bash profiles/configs_sweep.sh > gpurun_out/configs_r02c.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r02c.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/launches_bench_r02c.log 2>&1
tail -c 600 gpurun_out/bench_r02c.json; echo; tail -3 gpurun_out/configs_r02c.jsonl | cut -c1-300
