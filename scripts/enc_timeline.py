"""Timeline of one pipelined encode call (HUF_B200_OPT_KERNEL_TIMING = 2): start and duration of
every launch relative to the first.  usage: python scripts/enc_timeline.py [mib] [blocksize]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

import libhuffman_b200
from libhuffman_b200 import datagen
from libhuffman_b200.capi import DeviceCodec

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
lib = libhuffman_b200.load()
dev = torch.device("cuda", 0)
n = mib << 20
st = torch.cuda.current_stream().cuda_stream
x = datagen.zipf_torch(n, dev, 255, seed=2)
c = DeviceCodec(lib, 0)
cap = c.encode_bound(n, bs)
out = torch.empty(cap, dtype=torch.uint8, device=dev)
for _ in range(3):
    c.encode_async(x.data_ptr(), n, bs, out.data_ptr(), cap, st)
    c.encode_finish()
lib.check(lib.dll.huf_b200_ctx_set_option(c.ctx, 2, 2), "opt")
c.encode_async(x.data_ptr(), n, bs, out.data_ptr(), cap, st)
c.encode_finish()
rows = []
for name, ms in c.kernel_times():
    k, at = name.split("@")
    rows.append((float(at), k, ms))
for i, (at, k, ms) in enumerate(rows):
    print(f"{i:3d} pass {i // 7}  {k:16s} start {at:7.3f}  dur {ms:6.3f}  end {at + ms:7.3f}")
print("total", max(a + m for a, _, m in rows))
