# final state: whole GPU suite, smoke, both bench arms with the driver's flags
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ap_ref.json 2> gpurun_out/bench_ap_ref.err; echo "ref rc $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_ap_ref.json').read().strip().splitlines()[-1]); print('ref', d['value'], d['config']==None, d.get('impl'))"
python bench.py > gpurun_out/bench_ap.json 2> gpurun_out/bench_ap.err; echo "rc $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_ap.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print('e2e',d['e2e']['value']); print('roofline', d['roofline']['frac'], d['roofline']['executed']['frac'], d['roofline']['traffic']); print('lat',d['latency']['p50_ms']); print('parity',{k:v for k,v in d['parity'].items() if k not in ('what','variants')}); print('cpu', d['cpu_baseline']['value'])
r=json.loads(open('gpurun_out/bench_ap_ref.json').read().strip().splitlines()[-1]); print('same config', r['config']==d['config'])
"
