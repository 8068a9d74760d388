"""Steady-state e2e (page-locked memory streams, reused) against span size and slot count.
usage: python scripts/e2e_span_sweep.py [mib]"""
import ctypes as C, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import libhuffman_b200
from libhuffman_b200 import datagen
from libhuffman_b200.capi import Config
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = mib << 20
lib = libhuffman_b200.load()
host = datagen.zipf(n, 255, seed=2)
cap = lib.dll.huf_b200_encode_bound(n, 65536)
src = lib.memstream(n); mid = lib.memstream(cap); dst = lib.memstream(n)


def rnd():
    for s_ in (src, mid, dst):
        lib.dll.huf_memrewind(s_.rw)
    src.write(host)
    t0 = time.perf_counter()
    cfg = Config(length=n, blocksize=65536, reader=src.rw, writer=mid.rw)
    assert lib.dll.huf_encode(C.byref(cfg)) == 0
    t1 = time.perf_counter()
    cfg = Config(length=len(mid), reader=mid.rw, writer=dst.rw)
    assert lib.dll.huf_decode(C.byref(cfg)) == 0
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


for _ in range(5):
    rnd()
for span, slots in ((32, 5), (16, 5), (8, 6), (64, 5), (128, 4), (16, 6), (24, 6), (48, 5)):
    os.environ["HUF_B200_SPAN_MIB"] = str(span)
    os.environ["HUF_B200_SLOTS"] = str(slots)
    rnd()
    r = [rnd() for _ in range(4)]
    e = min(x[0] for x in r); d = min(x[1] for x in r)
    print(f"span {span:4d} MiB slots {slots}: encode {e*1e3:6.2f} ms decode {d*1e3:6.2f} ms  e2e {2*n/(e+d)/1e9:6.2f} GB/s", flush=True)
