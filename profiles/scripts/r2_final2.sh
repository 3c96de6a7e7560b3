# second final pass (after the pipelined LU inverse): full GPU test-suite, C2 / C1 / C4 / sweep lines, reference arm
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2g_pytest.log
tail -4 gpurun_out/r2g_pytest.log
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/r2g_$name.json 2> gpurun_out/r2g_$name.err; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/r2g_$name.json').read().strip().splitlines()[-1])
    print('$name', d.get('value'), d.get('unit'), d.get('ms_per_step'), 'e2e', (d.get('e2e') or {}).get('value'), 'parity', d.get('parity'), 'roofline', (d.get('roofline') or {}).get('achieved'), (d.get('roofline') or {}).get('frac'), 'clocks', d.get('clocks'))
except Exception as e:
    print('$name FAILED', e)
P
}
run C2 python bench.py
run C2_reference python bench.py --impl reference
run C1 python bench.py --config C1
run C4 python bench.py --config C4 --steps 2
run sweep_C2 python bench.py --workload sweep --config C2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
