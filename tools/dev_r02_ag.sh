# final code, 8 GPUs: the bench line at N = 8 (weak + strong blocks), NCCL init lines in stderr
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 3 --warmup 3 --no-latency --no-cpu-baseline --no-generic > gpurun_out/bench_r02ax_n8.json 2> gpurun_out/bench_r02ax_n8.err
echo "rc $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_r02ax_n8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print('e2e',d['e2e']['value']); print('strong',{k:v for k,v in d['strong'].items() if k!='note'}); print('conv',d['convergence'])
"
grep -c "NCCL INFO" gpurun_out/bench_r02ax_n8.err; grep -E "nranks" gpurun_out/bench_r02ax_n8.err | head -2 | cut -c1-200
