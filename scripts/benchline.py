"""One-line digest of a bench.py JSON line read from stdin (tuning runs)."""
import json
import sys

for ln in sys.stdin:
    ln = ln.strip()
    if not ln.startswith("{"):
        continue
    d = json.loads(ln)
    k = d.get("kernel_ms_per_step", {})
    e = d.get("e2e") or {}
    print("%.4g seg/s  step %.0f ms | closest %.1f vismap %.1f anyhit %.1f shade %.1f sort %.1f splat %.1f | iters %d | e2e %s (%.0f ms, create %.0f ms) | roofline %.3f" % (
        d["value"], d["ms_per_step"], k.get("closest", 0), k.get("vismap", 0), k.get("anyhit", 0), k.get("shade", 0), k.get("sort", 0), k.get("splat", 0),
        d["launches_per_step"]["closest"], ("%.4g" % e["value"]) if e.get("value") else "-", e.get("ms_per_step", 0),
        e.get("scene_create_ms", 0), d["roofline"]["frac"]))
