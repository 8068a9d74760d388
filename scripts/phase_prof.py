"""Phase timing of k_decode (thread 0 of every CTA, clock64 between barriers).

  python scripts/phase_prof.py build          # here: nvcc -DHUF_PHASE_PROF -> libhuffman_b200/_build/libhuffman_b200_prof.so
  python scripts/phase_prof.py run [shape] [mib]   # on the GPU box

The profiling library is a debug build of the same sources; it is never loaded by the package."""
import ctypes, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from libhuffman_b200 import build as B
PROF = B.OBJ / "libhuffman_b200_prof.so"

if sys.argv[1] == "build":
    B.build()
    obj = B.OBJ / "huf_b200_prof.o"
    subprocess.run([B.NVCC, *B.NVCC_FLAGS, "-DHUF_PHASE_PROF", "-c", str(B.CUDA_SOURCES[0]), "-o", str(obj)], check=True,
                   capture_output=True)
    objs = [str(B.OBJ / (s.stem + ".o")) for s in B.HOST_SOURCES] + [str(obj)]
    subprocess.run([B.NVCC, "-shared", "-o", str(PROF), *objs, "-Xlinker", "-Bsymbolic", "-lpthread"], check=True)
    print(PROF)
    sys.exit(0)

os.environ["HUF_B200_ACCEPT_1025"] = "1"
import torch
import libhuffman_b200
from libhuffman_b200 import datagen
from libhuffman_b200.capi import B200Lib, DeviceCodec
shape = sys.argv[2] if len(sys.argv) > 2 else "zipf255"
mib = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
bs = 65536
n = mib << 20; small = min(n, 64 << 20)
gen = {"fibonacci": lambda: datagen.fibonacci(small, bs, seed=4), "geometric": lambda: datagen.geometric(small, seed=4),
       "english": lambda: datagen.english_text(small, seed=1), "zipf255": lambda: datagen.zipf(small, 255, seed=2),
       "uniform": lambda: datagen.uniform(small, 256, seed=3)}[shape]
x = torch.frombuffer(bytearray(gen()), dtype=torch.uint8).cuda().repeat(n // small)[:n].contiguous()
lib = B200Lib(PROF)
enc = DeviceCodec(lib, 0); dec = DeviceCodec(lib, 0, accept_1025=True)
cap = enc.encode_bound(n, bs); comp = torch.empty(cap, dtype=torch.uint8, device="cuda"); back = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
enc.encode_async(x.data_ptr(), n, bs, comp.data_ptr(), cap, 0); c = enc.encode_finish()
raw = ctypes.CDLL(str(PROF))
out = (ctypes.c_ulonglong * 16)()
for rep in range(2):
    raw.huf_b200_debug_phase(out, 1)
    dec.set_kernel_timing(True)
    dec.decode_async(comp.data_ptr(), c, c, back.data_ptr(), n + 64, 0); r = dec.decode_finish()
raw.huf_b200_debug_phase(out, 0)
v = list(out)
names = ["0 handout+table", "1 chunk head+stage", "2 thread0 walk", "3 wait+verify rounds", "4 scan+fin", "5 barrier behind region copy", "6 copy-out", None, None, None, "10 next-chunk request", "11 region copy (thread 0)"]
tot = sum(v[:7]) + v[10] + v[11]
print(shape, mib, "MiB ok", bool(torch.equal(back[:n], x)), [t for t in dec.kernel_times() if t[0] == "k_decode"])
for k, nm in enumerate(names):
    if nm is None:
        continue
    print(f"  {nm:24s} {100.0 * v[k] / tot:6.2f} %   {v[k] / max(v[7], 1):9.0f} cycles per chunk")
print(f"  chunks {v[7]}  blocks {v[9]}  repair rounds {v[8]}  chunks/block {v[7] / max(v[9], 1):.2f}  rounds/chunk {v[8] / max(v[7], 1):.3f}")
print(f"  cycles per chunk (thread 0 timeline) {tot / max(v[7], 1):.0f}")
print(f"  thread 0 walks {v[14]}: bulk phase kept {v[12]}, thrown away {v[13]}")
