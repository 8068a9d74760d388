"""Per-kernel CUDA-event times of one encode + decode of a named shape (device resident).
usage: kernel_times.py shape [mib] [blocksize]"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ["HUF_B200_ACCEPT_1025"] = "1"
import torch
import libhuffman_b200
if os.environ.get("HUF_EXP_LIB"):  # an experiment build of the library (scripts/_bin/)
    libhuffman_b200.LIB_PATH = Path(os.environ["HUF_EXP_LIB"]).resolve()
from libhuffman_b200 import datagen
from libhuffman_b200.capi import DeviceCodec
shape = sys.argv[1]; mib = int(sys.argv[2]) if len(sys.argv) > 2 else 256; bs = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
n = mib << 20; small = min(n, 64 << 20)
gen = {"fibonacci": lambda: datagen.fibonacci(small, min(bs, 1 << 20), seed=4), "geometric": lambda: datagen.geometric(small, seed=4),
       "english": lambda: datagen.english_text(small, seed=1), "zipf255": lambda: datagen.zipf(small, 255, seed=2),
       "uniform": lambda: datagen.uniform(small, 256, seed=3)}[shape]
x = torch.frombuffer(bytearray(gen()), dtype=torch.uint8).cuda().repeat(n // small)[:n].contiguous()
lib = libhuffman_b200.load(); enc = DeviceCodec(lib, 0); dec = DeviceCodec(lib, 0, accept_1025=True)
cap = enc.encode_bound(n, bs); comp = torch.empty(cap, dtype=torch.uint8, device="cuda"); back = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
for rep in range(2):
    enc.set_kernel_timing(rep == 1); dec.set_kernel_timing(rep == 1)
    enc.encode_async(x.data_ptr(), n, bs, comp.data_ptr(), cap, 0); c = enc.encode_finish()
    te = enc.kernel_times() if rep else None
    dec.decode_async(comp.data_ptr(), c, c, back.data_ptr(), n + 64, 0); r = dec.decode_finish()
    td = dec.kernel_times() if rep else None
print(shape, "mib", mib, "bs", bs, "ratio", round(c / n, 4), "decode", r, "slow", dec.slow_blocks(), "ok", bool(torch.equal(back[:n], x)))
for name, ms in te + td:
    print(f"  {name:16s} {ms:8.3f} ms")
