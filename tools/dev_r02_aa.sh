# batched loads in the flat passes of rollout / cost / LQ / update_nominal; terminal records on the fourth warp
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for i in 1 2; do python tools/profile_case.py 16384 config3 2 | tail -1; done
python tools/profile_case.py 2048 config3 2 | tail -1
python tools/profile_case.py 1 config2 3 | tail -1
