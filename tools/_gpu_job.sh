timeout 900 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2_t11.log; tail -12 gpurun_out/r2_t11.log
for w in c5 c3 c2 c4; do
  MCB200_DEBUG_COUNTERS=1 timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_${w}_b11.json 2> gpurun_out/r2_${w}_b11.err
  grep "mcut_b200\] pairs" gpurun_out/r2_${w}_b11.err | tail -1
done
python -c "
import json
for w in ('c5','c3','c2','c4'):
    try:
        d=json.load(open('gpurun_out/r2_%s_b11.json'%w)); print(w, d['ms_per_step'], d['e2e']['ms_per_step'], {k:round(v['ms_per_launch']*1000,1) for k,v in d['kernels'].items()})
    except Exception as e: print(w, 'ERR', e)
"
