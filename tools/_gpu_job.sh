timeout 40 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_c4.json 2> gpurun_out/r2_final_c4.err
python -c "
import json
d=json.load(open('gpurun_out/r2_final_c4.json')); print('c4', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['steps'])"
