# final code, 2 GPUs, the driver's own command line (default flags) for both arms
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_r02al_ref_n2.json 2> gpurun_out/bench_r02al_ref_n2.err; echo "ref rc $?"; tail -c 600 gpurun_out/bench_r02al_ref_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r02al_n2.json 2> gpurun_out/bench_r02al_n2.err
echo "rc $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_r02al_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print('e2e',d['e2e']['value']); print('strong',{k:v for k,v in d['strong'].items() if k!='note'}); print('lat',d['latency']['p50_ms']); print('parity',{k:v for k,v in d['parity'].items() if k not in ('what','variants')}); print('cpu', d['cpu_baseline']['value'])
"
wc -l gpurun_out/bench_r02al_n2.json; grep -c "NCCL INFO" gpurun_out/bench_r02al_n2.err
