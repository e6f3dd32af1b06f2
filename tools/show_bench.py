"""Print the key numbers of one or more bench.py JSON lines (files or stdin)."""
import json, sys
for path in (sys.argv[1:] or ["-"]):
    txt = sys.stdin.read() if path == "-" else open(path).read()
    for ln in txt.splitlines():
        ln = ln.strip()
        if not ln.startswith("{"):
            continue
        d = json.loads(ln)
        r = d.get("roofline", {})
        print(f"{path}: n_gpus={d.get('n_gpus')} value={d.get('value', 0):.0f} {d.get('unit')} ms/step={d.get('ms_per_step', 0):.4f} "
              f"e2e={d.get('e2e', {}).get('value', 0):.0f} kernel_ms={r.get('kernel_ms', 0):.4f} roofline_frac={r.get('frac', 0):.3f} "
              f"clocks={d.get('clocks', {}).get('sm_mhz')} launches={d.get('gpu_launches')}")
        for k, v in (d.get("extras") or {}).items():
            if isinstance(v, dict):
                print("   ", k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a in ("ms", "GBps", "frac_hbm", "TOPS", "us", "kernel_only_TOPS")})
