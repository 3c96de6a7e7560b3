mkdir -p gpurun_out
SLA_LDR_ALGO=mc timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ldr_random or ldr_graded" 2>&1 | grep -v "^$" | tail -40 | cut -c1-300
SLA_LDR_ALGO=mc SLA_LDR_MC_CS=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ldr_random or ldr_graded" 2>&1 | tail -5 | cut -c1-300
