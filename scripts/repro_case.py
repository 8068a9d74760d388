"""Run one named shape through huf_encode/huf_decode on the GPU and compare with the oracle.
usage: repro_case.py shape [mib] [blocksize]   (handy under compute-sanitizer)"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import libhuffman_b200
from libhuffman_b200 import datagen
from oracle import harness

shape = sys.argv[1]
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 8
bs = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
n = mib << 20
data = {
    "english": lambda: datagen.english_text(n, seed=1),
    "zipf256": lambda: datagen.zipf(n, 256, seed=2),
    "zipf255": lambda: datagen.zipf(n, 255, seed=2),
    "uniform": lambda: datagen.uniform(n, 256, seed=3),
    "fibonacci": lambda: datagen.fibonacci(n, bs, seed=4),
    "geometric": lambda: datagen.geometric(n, seed=4),
}[shape]()
lib = libhuffman_b200.load()
want = harness.oracle_encode(data, bs)
rc, got = lib.encode(data, bs)
print("encode rc", rc, "equal", got == want)
rc, back = lib.decode(want)
print("decode rc", rc, "equal", back == data)

# device-resident entry points: more detail
import ctypes as C
import torch
from libhuffman_b200.capi import DeviceCodec
codec = DeviceCodec(lib, 0)
x = torch.frombuffer(bytearray(want), dtype=torch.uint8).cuda()
out = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
codec.decode_async(x.data_ptr(), len(want), len(want), out.data_ptr(), n + 64, 0)
try:
    print("device decode:", codec.decode_finish(), "slow blocks", codec.slow_blocks())
except Exception as e:
    print("device decode failed:", e)
print("cuda error state:", torch.cuda.synchronize())
