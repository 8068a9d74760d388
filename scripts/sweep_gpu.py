"""Device-resident encode/decode throughput over the named shapes and block sizes
(BASELINE.json configs 2-4, single GPU).  Prints one JSON line per case.
usage: sweep_gpu.py [mib]"""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ["HUF_B200_ACCEPT_1025"] = "1"   # 256-symbol shapes round-trip only with the opt-in (Q2)
import torch

import libhuffman_b200
from libhuffman_b200 import datagen
from libhuffman_b200.capi import DeviceCodec

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = mib << 20
dev = torch.device("cuda", 0)
lib = libhuffman_b200.load()
enc = DeviceCodec(lib, 0)
dec = DeviceCodec(lib, 0, accept_1025=True)
st = torch.cuda.current_stream().cuda_stream
peak = 6461.2
pk = Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"
if pk.exists():
    peak = json.loads(pk.read_text()).get("hbm_gbs", peak)


def host_shape(name, bs):
    small = 64 << 20
    gen = {"fibonacci": lambda: datagen.fibonacci(small, bs if bs <= (1 << 20) else 65536, seed=4),
           "geometric": lambda: datagen.geometric(small, seed=4),
           "english": lambda: datagen.english_text(small, seed=1)}[name]
    t = torch.frombuffer(bytearray(gen()), dtype=torch.uint8).to(dev)
    return t.repeat(n // small)[:n].contiguous()


cases = []
for bs in (4096, 16384, 65536, 262144, 1 << 20):
    cases.append(("zipf255", bs))
for name in ("zipf256", "uniform", "fibonacci", "geometric", "english"):
    cases.append((name, 65536))
cases.append(("fibonacci", 1 << 20))

for name, bs in cases:
    if name == "zipf255":
        x = datagen.zipf_torch(n, dev, 255, seed=2)
    elif name == "zipf256":
        x = datagen.zipf_torch(n, dev, 256, seed=2)
    elif name == "uniform":
        x = torch.randint(0, 256, (n,), dtype=torch.uint8, device=dev, generator=torch.Generator(dev).manual_seed(3))
    else:
        x = host_shape(name, bs)
    cap = enc.encode_bound(n, bs)
    comp = torch.empty(cap, dtype=torch.uint8, device=dev)
    back = torch.empty(n + 64, dtype=torch.uint8, device=dev)

    def e():
        enc.encode_async(x.data_ptr(), n, bs, comp.data_ptr(), cap, st)

    e()
    csize = enc.encode_finish()

    def d():
        dec.decode_async(comp.data_ptr(), csize, csize, back.data_ptr(), n + 64, st)

    d()
    r = dec.decode_finish()
    ok = r == (0, n, csize) and torch.equal(back[:n], x)
    slow = dec.slow_blocks()

    def timed(fn, fin, reps=5):
        for _ in range(2):
            fn(); fin()
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            fn()
        t1.record(); torch.cuda.synchronize(); fin()
        return t0.elapsed_time(t1) / reps / 1e3

    te = timed(e, enc.encode_finish)
    td = timed(d, dec.decode_finish)
    print(json.dumps({"shape": name, "blocksize": bs, "mib": mib, "ratio": round(csize / n, 4), "roundtrip_ok": bool(ok),
                      "slow_lane_blocks": slow, "encode_gbs": round(n / te / 1e9, 1), "decode_gbs": round(n / td / 1e9, 1),
                      "encode_frac_of_measured_hbm": round((n + csize) / te / 1e9 / peak, 4),
                      "decode_frac_of_measured_hbm": round((n + csize) / td / 1e9 / peak, 4)}), flush=True)
    del x, comp, back
