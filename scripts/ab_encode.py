"""A/B of the pipelined encode on one GPU: one stream against the staggered pass pipeline at a few
slot counts / pass sizes / block sizes.  usage: python scripts/ab_encode.py [mib]"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

import libhuffman_b200
from libhuffman_b200 import datagen
from libhuffman_b200.capi import DeviceCodec

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lib = libhuffman_b200.load()
dev = torch.device("cuda", 0)
n = mib << 20
st = torch.cuda.current_stream().cuda_stream


def run(x, bs, env, no_overlap, reps=10):
    for k in ("HUF_B200_ENC_SLOTS", "HUF_B200_ENC_PIPE_PASS", "HUF_B200_ENC_PIPE_MIN"):
        os.environ.pop(k, None)
    os.environ.update(env)
    c = DeviceCodec(lib, 0)
    lib.check(lib.dll.huf_b200_ctx_set_option(c.ctx, 4, int(no_overlap)), "opt")
    cap = c.encode_bound(n, bs)
    out = torch.empty(cap, dtype=torch.uint8, device=dev)
    for _ in range(3):
        c.encode_async(x.data_ptr(), n, bs, out.data_ptr(), cap, st)
        size = c.encode_finish()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        c.encode_async(x.data_ptr(), n, bs, out.data_ptr(), cap, st)
    e1.record()
    c.encode_finish()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    h = int(out[:size].to(torch.int64).sum().item())
    c.close()
    return ms, size, h


x = datagen.zipf_torch(n, dev, 255, seed=2)
combos = ((8, 4), (8, 4), (8, 4))
if os.environ.get("AB_SLOTS"):  # e.g. AB_SLOTS=2,4,6,8: one pipeline run per slot count
    combos = tuple((int(v), 4) for v in os.environ["AB_SLOTS"].split(","))
sizes = (65536, 262144, 65536, 262144, 1 << 20, 16384)
if os.environ.get("AB_BS"):
    sizes = tuple(int(v) for v in os.environ["AB_BS"].split(","))
for bs in sizes:
    base = run(x, bs, {}, True)
    print(f"bs {bs:8d}  one stream      {base[0]:.3f} ms  {n / base[0] / 1e6:7.1f} GB/s", flush=True)
    for slots, pass_mib in combos:
        env = {"HUF_B200_ENC_SLOTS": str(slots), "HUF_B200_ENC_PIPE_PASS": str(pass_mib << 20)}
        r = run(x, bs, env, False)
        ok = r[1:] == base[1:]
        print(f"bs {bs:8d}  slots {slots} pass>={pass_mib:3d}M {r[0]:.3f} ms  {n / r[0] / 1e6:7.1f} GB/s  same={ok}", flush=True)
