"""Print one replayed step of a trace written by scripts/timeline_step.py."""
import json, sys
tr = json.load(open(sys.argv[1]))
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
minus = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
ev = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
n = len(ev) // nsteps
step = ev[n:2 * n]
t0 = step[0]["ts"]
for e in step:
    s, d = e["ts"] - t0, e["dur"]
    if d < minus:
        continue
    print(f"{s:8.1f} {d:7.1f} end {s + d:8.1f} st{e['args'].get('stream')} {e['name'][:64]}")
