"""Time huf_encode / huf_decode over memory streams with host buffers (the e2e leg of bench.py),
with HUF_B200_DEBUG=1 printing the stage timings.  usage: e2e_probe.py [mib] [reuse]
`reuse`: keep the three streams across iterations (rewound), i.e. output pages already touched."""
import ctypes as C, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import libhuffman_b200
from libhuffman_b200 import datagen
from libhuffman_b200.capi import Config
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reuse = len(sys.argv) > 2 and sys.argv[2] == "reuse"
n = mib << 20
lib = libhuffman_b200.load()
host = datagen.zipf(n, 255, seed=2)
cap = lib.dll.huf_b200_encode_bound(n, 65536)
print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), flush=True)
streams = None
for it in range(8):
    if streams is None or not reuse:
        src = lib.memstream(n); src.write(host)
        mid = lib.memstream(cap); dst = lib.memstream(n)
        streams = (src, mid, dst)
    else:
        src, mid, dst = streams
        lib.dll.huf_memrewind(mid.rw); lib.dll.huf_memrewind(dst.rw)
        lib.dll.huf_memrewind(src.rw); src.write(host)
    t0 = time.perf_counter()
    cfg = Config(length=n, blocksize=65536, reader=src.rw, writer=mid.rw)
    assert lib.dll.huf_encode(C.byref(cfg)) == 0
    t1 = time.perf_counter()
    clen = len(mid)
    cfg = Config(length=clen, reader=mid.rw, writer=dst.rw)
    assert lib.dll.huf_decode(C.byref(cfg)) == 0
    t2 = time.perf_counter()
    print(f"iter {it}: encode {t1-t0:.3f}s decode {t2-t1:.3f}s  e2e {2*n/(t2-t0)/1e9:.2f} GB/s", flush=True)
    if not reuse:
        for s_ in (src, mid, dst): s_.close()
