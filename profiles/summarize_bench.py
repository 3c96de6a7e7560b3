"""Markdown table of recorded bench lines:  python profiles/summarize_bench.py profiles/bench_r2_final_*.json"""
import json, sys
print("| record | value | ms per step | e2e | step TFLOP/s | dominant-kernel roofline | LDR kernel | parity (max rel err G, |dlogdet|, sign mismatches) | SM clock |")
print("|---|---|---|---|---|---|---|---|---|")
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f"| {f} | unreadable: {e} |"); continue
    if d.get("impl") == "reference":
        print(f"| `{f}` | {d['value']:.1f} {d['unit']} (CPU arm, {d['cpu_baseline']['cores']} threads) | {d['ms_per_step']:.1f} | - | - | - | - | - | - |"); continue
    r = d.get("roofline") or {}
    p = d.get("parity") or {}
    par = f"{p.get('max_rel_err_G', float('nan')):.1e}, {p.get('max_abs_err_logdet', float('nan')):.1e}, {p.get('sign_mismatch')}" if p else "-"
    c = d.get("clocks") or {}
    print(f"| `{f}` | {d['value']:.1f} {d['unit']} | {d['ms_per_step']:.2f} | {(d.get('e2e') or {}).get('value', float('nan')):.1f} | "
          f"{r.get('step_tflops', float('nan')):.1f} | {r.get('achieved', float('nan')):.2f} / {r.get('peak', float('nan')):.2f} {r.get('unit','')} = {r.get('frac', float('nan')):.3f} | "
          f"{d.get('ldr_kernel') or p.get('ldr_kernel')} | {par} | {c.get('sm_mhz')} MHz {c.get('reasons')} |")
