python bench.py --no-variants --no-cpu --no-kernels --no-e2e --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', round(d['value']/1e6,2), 'ms', round(d['ms_per_step'],2), 'launches', d['gpu_launches'], {k:round(v['ms'],2) for k,v in d['kernel_classes'].items()})
"
