# round 2, GPU call M: concurrent line search (clusters of four blocks per problem) in the latency kernel
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "not full_size and not at_scale" 2>&1 | tail -6
for v in 0 1 0 1; do echo "cluster line search $v"; HSDDP_CLUSTER_LS=$v python tools/profile_case.py 1 config1 6 | tail -2; done
for v in 0 1; do echo "cluster line search $v, 64 problems config3"; HSDDP_CLUSTER_LS=$v python tools/profile_case.py 64 config3 4 | tail -1; done
for v in 0 1; do echo "cluster line search $v, 32 problems config4"; HSDDP_CLUSTER_LS=$v python tools/profile_case.py 32 config4 4 | tail -1; done
