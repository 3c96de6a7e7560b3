for v in 0 1; do
  SLA_REG_SPLITQ_SMALL=$v timeout 60 python profiles/ldr_bench_graded.py 64 1 20 2>&1 | tail -1 | sed "s/^/split=$v /"
  SLA_REG_SPLITQ_SMALL=$v timeout 60 python profiles/ldr_bench_graded.py 64 64 20 2>&1 | tail -1 | sed "s/^/split=$v /"
  SLA_REG_SPLITQ_SMALL=$v timeout 60 python profiles/ldr_bench_graded.py 128 32 40 2>&1 | tail -1 | sed "s/^/split=$v /"
  SLA_REG_SPLITQ_SMALL=$v timeout 100 python bench.py --config C1 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('split=$v C1', d['ms_per_step'], d['value'], d['gpu_launches'])"
done
SLA_REG_SPLITQ_SMALL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "ldr_random or ldr_graded or tiny or chain_C1 or test_ref_ or identity" 2>&1 | tail -3
