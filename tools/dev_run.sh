# Developer loop on a GPU box: parity diagnostics + throughput probes (prints only; assertions live in tests/)
for m in ${MODES:-1 2}; do
export HSDDP_SOLVE_MODE=$m
echo "=== solve mode $m"
python tools/gpu_dev_check.py 2>&1 | grep -E "^\s+p[0-9]|throughput|Error|error|Traceback" 
python tools/profile_case.py 4096 config2 2
python tools/profile_case.py 4096 config3 2
python tools/profile_case.py 16384 config3 1
python tools/profile_case.py 1 config2 2
done
