"""Print the interesting parts of a bench.py JSON line.  usage: show_bench.py FILE"""
import json
import sys

lines = [ln for ln in open(sys.argv[1]).read().strip().splitlines() if ln.startswith("{")]
d = json.loads(lines[-1])
print({k: round(d[k], 2) for k in ["value", "encode_gbs", "decode_gbs", "decode_with_block_index_gbs", "ms_per_step"] if k in d})
print("e2e", d.get("e2e"))
print("strong", d.get("strong"))
print("config5", d.get("config5_python_streaming"))
for r in d.get("sweep", []) or []:
    print("  ", r)
r = d.get("roofline") or {}
print("roofline", {k: v for k, v in r.items() if k != "kernel_ms"})
print("kernel_ms", r.get("kernel_ms"))
print("cpu", d.get("cpu_baseline"))
print("clocks", d.get("clocks"))
