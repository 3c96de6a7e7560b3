# cluster-split LU inverse, sweep stepper, C client, blocked kernel child test
mkdir -p gpurun_out
SLA_LU_CLUSTER=1 timeout 120 python profiles/inv_bench.py 256 64 2>&1 | tail -3
timeout 120 python profiles/inv_bench.py 256 64 2>&1 | tail -3
SLA_LU_CLUSTER=4 timeout 120 python profiles/inv_bench.py 256 32 2>&1 | tail -3
SLA_LU_CLUSTER=1 timeout 200 python profiles/inv_bench.py 1024 32 2>&1 | tail -3
timeout 200 python profiles/inv_bench.py 1024 32 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_abi.py -m gpu -x -q -k "lu or inv or ldiv or rdiv or sweep or c_client or info or blk" 2>&1 | tail -15
