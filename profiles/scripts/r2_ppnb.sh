# panel mode: panel width vs accuracy (C1-size sweep, C2 chains) and speed
for nb in 8 16 32; do SLA_PP_NB=$nb SLA_LDR_PIVOT=panel python profiles/sweep_err.py 2>&1 | tail -4 | sed "s/^/NB=$nb /"; done
for nb in 32 16 8; do
  SLA_PP_NB=$nb timeout 300 python bench.py --config C2 --steps 5 --ldr-pivoting panel 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2 panel NB=$nb', d['ms_per_step'], d['parity'])"
done
for nb in 32 16; do
  SLA_PP_NB=$nb timeout 300 python bench.py --config C5 --steps 2 --chains 128 --ldr-pivoting panel 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C5(128 chains) panel NB=$nb', d['ms_per_step'], d['parity'])"
done
timeout 300 python bench.py --config C5 --steps 2 --chains 128 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C5(128 chains) strict', d['ms_per_step'], d['parity'])"
