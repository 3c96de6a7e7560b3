for t in 0 6462 6463 6484 6482; do
  SLA_GEMM_TILE=$t timeout 120 python profiles/gemm_bench.py 256 256 c s 2>&1 | head -1
  SLA_GEMM_TILE=$t timeout 120 python profiles/gemm_bench.py 256 128 c 2>&1 | head -1
done
