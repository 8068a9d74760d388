"""GPU stress over data shapes and block sizes: huf_encode vs oracle, huf_decode round trip,
device-resident decode with lane statistics.  usage: stress_gpu.py [mib]   (also handy under
compute-sanitizer --tool synccheck / memcheck)"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ.setdefault("HUF_B200_ACCEPT_1025", "1")
import libhuffman_b200
from libhuffman_b200 import datagen
from oracle import harness

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = mib << 20
lib = libhuffman_b200.load()
shapes = {
    "english": lambda: datagen.english_text(n, seed=1),
    "zipf256": lambda: datagen.zipf(n, 256, seed=2),
    "zipf255": lambda: datagen.zipf(n, 255, seed=2),
    "uniform": lambda: datagen.uniform(n, 256, seed=3),
    "fibonacci": lambda: datagen.fibonacci(n, 65536, seed=4),
    "geometric": lambda: datagen.geometric(n, seed=4),
    # 1 MiB Fibonacci blocks: 28-bit code words, the 64-bit table entries of the general packing lane
    "fib1m": lambda: datagen.fibonacci(max(n, 1 << 20) + 70001, 1 << 20, seed=5),
}
if os.environ.get("HUF_STRESS_SHAPES"):  # a comma-separated subset (sanitizer runs on a budget)
    shapes = {k: v for k, v in shapes.items() if k in os.environ["HUF_STRESS_SHAPES"].split(",")}
bad = 0
for name, gen in shapes.items():
    data = gen()
    for bs in (4096, 65536, 1 << 20):
        want = harness.oracle_encode(data, bs)
        rc, got = lib.encode(data, bs)
        ok_e = rc == 0 and got == want
        rc2, back = lib.decode(want)
        ok_d = rc2 == 0 and back == data
        print(f"{name:10s} bs {bs:8d} encode {'ok' if ok_e else 'FAIL rc=%d' % rc}  decode {'ok' if ok_d else 'FAIL rc=%d' % rc2}",
              flush=True)
        bad += (not ok_e) + (not ok_d)
print("failures:", bad)
sys.exit(1 if bad else 0)
