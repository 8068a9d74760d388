"""GPU diagnostic: encode N MiB of Zipf data and check every block's header against the data."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import libhuffman_b200
from libhuffman_b200 import datagen
from libhuffman_b200.capi import DeviceCodec
from oracle import harness

lib = libhuffman_b200.load()
codec = DeviceCodec(lib, 0)
bs = 65536
for mib in (64, 256, 1024, 1024):
    n = mib << 20
    x = datagen.zipf_torch(n, "cuda", 256, seed=2)
    torch.cuda.synchronize()
    cap = codec.encode_bound(n, bs)
    out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    codec.encode_async(x.data_ptr(), n, bs, out.data_ptr(), cap, torch.cuda.current_stream().cuda_stream)
    size = codec.encode_finish()
    ptr, nb = codec.block_offsets()
    offs = np.empty(nb + 1, dtype=np.uint64)
    lib.dll.huf_b200_copy_d2h(offs.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), 8 * (nb + 1))
    o = torch.from_numpy(offs[:-1].astype(np.int64)).cuda()
    tl = out[o + 8].to(torch.int64) | (out[o + 9].to(torch.int64) << 8)
    # expected distinct symbols per block
    xb = x.view(nb, bs)
    nsym = torch.zeros(nb, dtype=torch.int64, device="cuda")
    for lo in range(0, nb, 1024):
        chunk = xb[lo:lo + 1024].to(torch.int64)
        oh = torch.zeros(chunk.shape[0], 256, dtype=torch.int32, device="cuda")
        oh.scatter_add_(1, chunk, torch.ones_like(chunk, dtype=torch.int32))
        nsym[lo:lo + 1024] = (oh > 0).sum(1)
    bad = torch.nonzero(tl != 4 * nsym + 1).flatten().cpu().tolist()
    print(f"{mib} MiB: size={size} nblocks={nb} bad_tree_len_blocks={len(bad)} first={bad[:10]} last={bad[-5:]}")
    for b in bad[:3]:
        print("   block", b, "tree_len", int(tl[b]), "expected", int(4 * nsym[b] + 1))
    # full compare of a few blocks
    rng = np.random.default_rng(0)
    mism = []
    for b in [0, 1, nb - 2, nb - 1, *rng.integers(0, nb, 12).tolist()]:
        blk = xb[b].cpu().numpy().tobytes()
        got = out[int(offs[b]):int(offs[b + 1])].cpu().numpy().tobytes()
        if got != harness.oracle_encode(blk, 0):
            mism.append(b)
    print("   sampled full-compare mismatches:", mism)
codec.close()
