timeout 900 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -6 > gpurun_out/r2_t13.log; tail -4 gpurun_out/r2_t13.log
for w in c2 c5 c3; do
  python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_${w}_b13.json 2> gpurun_out/r2_${w}_b13.err
  MCB200_MORTON_SORT_BITS=16 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_${w}_b13_m16.json 2>/dev/null
done
python bench.py --workload c4batch --lanes 16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c4batch_b13.json 2> gpurun_out/r2_c4batch_b13.err
python bench.py --workload c3batch --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c3batch_b13.json 2> gpurun_out/r2_c3batch_b13.err
python -c "
import json
for w in ('c2_b13','c2_b13_m16','c5_b13','c5_b13_m16','c3_b13','c3_b13_m16','c4batch_b13','c3batch_b13'):
    try:
        d=json.load(open('gpurun_out/r2_%s.json'%w)); print(w, d['ms_per_step'], d['value'], d['e2e']['value'], d.get('kernels',{}).get('k_traverse',{}).get('ms_per_launch'), d['config'].get('n_node_tests'))
    except Exception as e: print(w, 'ERR', e)
"
