# round 2, GPU call L: full GPU test-suite, then the bench line
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py > gpurun_out/bench_r02l.json 2> gpurun_out/bench_r02l.err || tail -20 gpurun_out/bench_r02l.err
cat gpurun_out/bench_r02l.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e',d['e2e']['value']); print('roofline',{k:d['roofline'][k] for k in ('achieved','frac','traffic')}, d['roofline']['executed']['frac'])
print('latency',d['latency']); print('parity',{k:v for k,v in d['parity'].items() if k!='what'}); print('cpu',d['cpu_baseline']['value'],d['cpu_baseline'].get('value_fma_build'),d['cpu_baseline'].get('value_march_native'))
"
