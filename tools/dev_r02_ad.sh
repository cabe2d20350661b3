# out-of-line rollout / cost block functions (instruction footprint of the persistent and latency kernels)
for i in 1 2; do python tools/profile_case.py 16384 config3 2 | tail -1; done
python tools/profile_case.py 2048 config3 2 | tail -1
python tools/profile_case.py 4096 config3 2 | tail -1
python tools/profile_case.py 1 config2 3 | tail -1
python tools/profile_case.py 32 config4 2 | tail -1
python tools/profile_case.py 296 config2 2 | tail -1
