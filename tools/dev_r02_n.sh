# round 2, GPU call N (2 GPUs): the multi-rank bench path (weak + strong blocks, NCCL statistics gather)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r02n_n2.json 2> gpurun_out/bench_r02n_n2.err
echo "rc $?"; tail -3 gpurun_out/bench_r02n_n2.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/bench_r02n_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print('e2e',d['e2e']['value']); print('strong',d['strong']); print('conv',d['convergence'])
"
grep -c "NCCL INFO" gpurun_out/bench_r02n_n2.err; grep -E "nranks|comm 0x" gpurun_out/bench_r02n_n2.err | head -4 | cut -c1-220
