#!/usr/bin/env python
"""bench.py — BASELINE.json's headline metric on its headline config.

A step = one pass of the hot path over one batch: huf_encode of 1 GiB of Zipf(1.1) bytes at
64 KiB blocks followed by huf_decode of the resulting stream (configs[1]).  `value` is
uncompressed GB/s of that round trip with inputs resident in HBM (2N / t_step, GB = 1e9 B),
`e2e` the same through the reference-facing C API (huf_encode / huf_decode over memory
streams, host buffers, host<->device copies inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Under torchrun (N > 1) every rank codes its own 1 GiB shard (contiguous block ranges of an
N GiB input; blocks are independent, so there is no data-path collective): weak scaling, that
is `value`.  The same line carries, as extra keys: `strong` (BASELINE configs[3]: ONE 64 GiB job
over the N GPUs, stitched stream, sharded decode, `stitched_ok`), and at N = 1 `sweep` (configs
2-4: shapes and block sizes) and `config5_python_streaming` (configs[4]).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GB = 1e9
BLOCK = 65536
METRIC = "encode+decode GB/s (uncompressed)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mib", type=int, default=1024, help="uncompressed MiB per GPU")
    ap.add_argument("--nsym", type=int, default=255,
                    help="alphabet of the Zipf source (256 trips the reference decoder's tree_len limit)")
    ap.add_argument("--blocksize", type=int, default=BLOCK)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample-mib", type=int, default=48)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the shape/block-size sweep and the config-5 leg (N=1)")
    ap.add_argument("--no-strong", action="store_true", help="skip the one-job-over-N-GPUs leg")
    ap.add_argument("--sweep-mib", type=int, default=1024)
    ap.add_argument("--config5-mib", type=int, default=4096)
    ap.add_argument("--strong-gib", type=int, default=64)
    ap.add_argument("--strong-steps", type=int, default=3)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks during the timed region (profiling recipe's nvidia-smi line)
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.lines: list[str] = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the unmodified reference (oracle/_ref) on the host cores
# ------------------------------------------------------------------------------------------------

def _cpu_worker(args):
    """Encode + decode one contiguous block range with the compiled reference; returns seconds."""
    seed, nbytes, nsym, blocksize, rounds = args
    from libhuffman_b200 import datagen
    from oracle import harness
    ref = harness.reference() if harness.reference_available() else None
    data = datagen.zipf(nbytes, nsym, seed=seed)
    times = []
    for _ in range(rounds):
        t0 = time.perf_counter()
        if ref is not None:
            rc, stream = ref.encode(data, blocksize, 65536, 65536)
            assert rc == 0
            t1 = time.perf_counter()
            rc, back = ref.decode(stream, None, 65536, 65536)
            assert rc == 0 and back == data
        else:
            stream = harness.oracle_encode(data, blocksize)
            t1 = time.perf_counter()
            rc, back, _ = harness.oracle_decode(stream)
            assert rc == 0 and back == data
        t2 = time.perf_counter()
        times.append((t1 - t0, t2 - t1))
    return times


def cpu_reference_run(procs: int, mib_per_proc: int, nsym: int, blocksize: int, rounds: int):
    """One process per core over contiguous block ranges; per round the slowest process counts."""
    from oracle import harness
    harness.build()
    kind = "reference" if harness.reference_available() else "port"
    nbytes = mib_per_proc << 20
    jobs = [(100 + p, nbytes, nsym, blocksize, rounds) for p in range(procs)]
    if procs == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_cpu_worker, jobs)
    per_round = []
    for r in range(rounds):
        enc = max(res[p][r][0] for p in range(procs))
        dec = max(res[p][r][1] for p in range(procs))
        per_round.append((enc, dec))
    return kind, nbytes * procs, per_round


def workload_name(args) -> str:
    """BASELINE.json configs[1]; both arms report the same string."""
    return (f"{args.mib} MiB Zipf(1.1) over {args.nsym} symbols per GPU, {args.blocksize} B blocks, "
            "huf_encode then huf_decode")


def measured_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed
    ncu --set full capture (profiles/traffic.json), or None."""
    f = ROOT / "profiles" / "traffic.json"
    if not f.exists():
        return None
    try:
        return json.loads(f.read_text()).get(kernel)
    except (ValueError, OSError):
        return None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    mib = max(1, min(8, args.cpu_sample_mib))
    rounds = args.warmup + args.steps
    # keep the whole run within a few minutes: ~13 MB/s per core, enc + dec
    est = rounds * 2 * mib / 12.0
    if est > 240:
        mib = max(1, int(mib * 240 / est))
    kind, total, per_round = cpu_reference_run(cores, mib, args.nsym, args.blocksize, rounds)
    timed = per_round[args.warmup:]
    t = sum(e + d for e, d in timed)
    value = 2 * total * len(timed) / t / GB
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / len(timed), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "impl": "reference",
        # (the same workload as the b200 arm: its throughput on the CPU does not depend on the size,
        # so every step codes a bounded sample of it -- said in `sample_of`, not as another config)
        "config": {"workload": workload_name(args), "blocksize": args.blocksize,
                   "blocks_per_gpu": (args.mib << 20) // args.blocksize},
        "sample_of": f"{workload_name(args)}: each step codes {mib} MiB per process on {cores} host processes over "
                     "block ranges (reference CPU path)",
        "encode_gbs": total * len(timed) / sum(e for e, _ in timed) / GB,
        "decode_gbs": total * len(timed) / sum(d for _, d in timed) / GB,
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": kind,
                         "sample": f"{mib} MiB per process x {cores} processes per step, bufio 64 KiB"},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

NOMINAL_HBM_GBS = 8000.0   # north_star's nominal peak (SURVEY.md §8d asks for both fractions)


def _device_timer(torch, dist, world, dev):
    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, fin):
        """Device time of `steps` back-to-back invocations, max over ranks, in seconds."""
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        fin()
        t = torch.tensor([e0.elapsed_time(e1) / 1e3], device=dev, dtype=torch.float64)
        barrier()
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    return barrier, timed


def _shape_tensor(torch, datagen, name, n, dev, bs):
    """The named shapes of BASELINE.json configs 2-3 in HBM (host-generated ones are tiled)."""
    if name in ("zipf255", "zipf256"):
        return datagen.zipf_torch(n, dev, 255 if name == "zipf255" else 256, seed=2)
    if name == "uniform":
        return datagen.uniform_torch(n, dev, 256, seed=3)
    small = min(n, 64 << 20)
    gen = {"fibonacci": lambda: datagen.fibonacci(small, bs if bs <= (1 << 20) else 65536, seed=4),
           "geometric": lambda: datagen.geometric(small, seed=4),
           "english": lambda: datagen.english_text(small, seed=1)}[name]
    t = torch.frombuffer(bytearray(gen()), dtype=torch.uint8).to(dev)
    return t.repeat((n + small - 1) // small)[:n].contiguous()


def sweep_leg(torch, lib, dev, peak, mib):
    """Configs 2-4 on one GPU, device resident: every named shape at 64 KiB blocks, the Zipf
    block-size sweep 4 KiB ... 1 MiB, the 256-symbol variants through the opt-in decoder mode."""
    from libhuffman_b200 import datagen
    from libhuffman_b200.capi import DeviceCodec
    n = mib << 20
    enc = DeviceCodec(lib, dev.index)
    dec = DeviceCodec(lib, dev.index, accept_1025=True)
    st = torch.cuda.current_stream().cuda_stream
    cases = [("zipf255", bs) for bs in (4096, 16384, 262144, 1 << 20)]
    cases += [(name, 65536) for name in ("zipf256", "uniform", "fibonacci", "geometric", "english")]
    cases.append(("fibonacci", 1 << 20))
    rows = []
    for name, bs in cases:
        x = _shape_tensor(torch, datagen, name, n, dev, bs)
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
        ok = dec.decode_finish() == (0, n, csize) and bool(torch.equal(back[:n], x))

        def timed(fn, fin, reps=5):
            for _ in range(2):
                fn()
                fin()
            torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True)
            t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(reps):
                fn()
            t1.record()
            torch.cuda.synchronize()
            fin()
            return t0.elapsed_time(t1) / reps / 1e3

        te, td = timed(e, enc.encode_finish), timed(d, dec.decode_finish)
        rows.append({"shape": name, "blocksize": bs, "mib": mib, "ratio": round(csize / n, 4), "roundtrip_ok": ok,
                     "encode_gbs": round(n / te / GB, 1), "decode_gbs": round(n / td / GB, 1),
                     "encode_frac": round((n + csize) / te / GB / peak, 4),
                     "decode_frac": round((n + csize) / td / GB / peak, 4)})
        del x, comp, back
    enc.close()
    dec.close()
    return rows


def config5_leg(mib: int):
    """BASELINE configs[4]: the reference's HuffmanCompressor (unchanged cffi package, built against
    this library by scripts/build_reference_suites.py) streaming `mib` MiB in 64 MiB chunks."""
    pkg = ROOT / "oracle" / "_ref_suites" / "huffmanfile_gpu"
    if not (pkg / "huffmanfile").is_dir():
        return {"unavailable": "oracle/_ref_suites not built (needs the reference checkout at build time)"}
    code = f"""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, {str(ROOT)!r})
import huffmanfile
from libhuffman_b200 import datagen
chunk = 64 << 20
data = datagen.zipf(chunk, 255, seed=5)
c = huffmanfile.HuffmanCompressor()
c.compress(data)                      # warm-up: context, arenas, pinned buffers
t0 = time.perf_counter(); n = 0
for _ in range({mib} // 64):
    n += len(c.compress(data))
n += len(c.flush())
dt = time.perf_counter() - t0
one = huffmanfile.compress(data)
huffmanfile.HuffmanDecompressor().decompress(one)   # warm-up of the decode side
dt2 = 0.0
for _ in range(4):                                   # (one object decompresses once, huffmanfile.py:391-392)
    d = huffmanfile.HuffmanDecompressor()
    t1 = time.perf_counter(); back = d.decompress(one); dt2 += (time.perf_counter() - t1) / 4
assert back == data
print("RESULT", {mib} << 20, n, dt, chunk / dt2)
"""
    try:
        proc = subprocess.run([sys.executable, "-c", code], cwd=pkg, capture_output=True, text=True, timeout=900)
    except subprocess.TimeoutExpired:
        return {"unavailable": "timed out"}
    rows = [ln for ln in proc.stdout.splitlines() if ln.startswith("RESULT")]
    if proc.returncode != 0 or not rows:
        return {"unavailable": (proc.stderr or proc.stdout)[-300:]}
    _, total, out, dt, dec_bps = rows[0].split()
    return {"workload": f"HuffmanCompressor.compress()+flush(), {mib} MiB Zipf(1.1)/255 in 64 MiB chunks, default blocksize 131072",
            "compress_gbs": int(total) / float(dt) / GB, "decompress_gbs": float(dec_bps) / GB,
            "ratio": int(out) / int(total), "api": "reference huffmanfile package (cffi) on libhuffman_b200.so"}


def strong_leg(args, torch, dist, lib, dev, rank, world, peak):
    """BASELINE configs[3]: ONE job sharded over the GPUs (libhuffman_b200/sharded.py): the input is
    partitioned by contiguous block ranges, every GPU encodes its range, the slabs form one
    stream (exclusive scan of their sizes); that stream is split by BYTE range for decode, every
    GPU finds and decodes the blocks that start in its range, the host validates the chain.
    Strong scaling: the total is fixed as N grows (N = 1 runs the first half as a sample: the
    whole job with input, stream and output resident does not fit one GPU)."""
    from libhuffman_b200 import datagen, shard
    from libhuffman_b200.sharded import ShardedCodec
    barrier, timed = _device_timer(torch, dist, world, dev)
    bs = args.blocksize
    total = args.strong_gib << 30
    job = total if world > 1 else total // 2
    nb = job // bs
    chunk = 64 << 20                                      # generator granule: seed = chunk index
    b_lo, b_hi = shard.block_range(nb, rank, world)
    lo, hi = b_lo * bs, b_hi * bs
    n = hi - lo
    x = torch.empty(n, dtype=torch.uint8, device=dev)

    def gen(into, at_lo, at_hi):
        """bytes [at_lo, at_hi) of the job: chunk c holds zipf_torch(chunk, seed=1000+c)."""
        c0 = at_lo // chunk
        while c0 * chunk < at_hi:
            piece = datagen.zipf_torch(chunk, dev, args.nsym, seed=1000 + c0)
            a, b = max(at_lo, c0 * chunk), min(at_hi, (c0 + 1) * chunk)
            into[a - at_lo:b - at_lo] = piece[a - c0 * chunk:b - c0 * chunk]
            c0 += 1

    gen(x, lo, hi)
    sc = ShardedCodec(lib, rank, world, dev, dev.index, accept_1025=args.nsym > 255)
    st = torch.cuda.current_stream().cuda_stream
    cap = sc.enc.encode_bound(n, bs)
    comp = torch.empty(cap, dtype=torch.uint8, device=dev)

    def enc_step():
        sc.encode_async(x, bs, comp, st)

    enc_step()
    size = sc.encode_finish()
    sizes = [r[0] for r in sc.all_gather_ints([size])]
    offs = shard.slab_offsets(sizes)
    csize = offs[-1]
    # position-dependent checksum of the one stream (equal across N <=> same stitched bytes)
    s1 = s2 = 0
    step = 256 << 20
    for at in range(0, size, step):
        v = comp[at:min(size, at + step)].to(torch.int64)
        pos = (torch.arange(v.numel(), device=dev, dtype=torch.int64) + (offs[rank] + at)) % 65521 + 1
        s1 += int(v.sum().item())
        s2 += int((v * pos).sum().item()) % (1 << 61)
        del v, pos
    sums = sc.all_gather_ints([s1, s2 % (1 << 61), b_hi * bs])
    half = job // 2 if world > 1 else job
    first_half = [sum(r[0] for r in sums if r[2] <= half), sum(r[1] for r in sums if r[2] <= half) % (1 << 61)]
    t_enc = timed(enc_step, args.strong_steps, sc.encode_finish)
    del x   # (the check below regenerates what it needs; the stream laid out by byte range needs the room)

    # ---- the one stream by byte range (setup, timed separately), then the sharded decode
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    overlap = max(8 << 20, 4 * bs)
    buf, base, cuts = sc.redistribute(comp, sizes, overlap)
    torch.cuda.synchronize()
    t_layout = time.perf_counter() - t0
    moved = buf.numel() - max(0, min(cuts[rank + 1] + overlap, offs[rank + 1]) - max(base, offs[rank]))
    if world > 1:
        del comp   # (at N = 1 the buffer is a view of it)
        comp = None
    est = sc.decode_plan(buf, base, cuts, st)
    back = torch.empty(est + 64, dtype=torch.uint8, device=dev)

    def dec_step():
        sc.decode_async(buf, base, cuts, back, st)

    dec_step()
    mine = sc.decode_finish(base)
    ok, end, outs = sc.validate(mine, cuts)
    stitched_ok = bool(ok and end == csize and outs and outs[-1] == job)
    if stitched_ok:
        # my decoded slab is the job's bytes [outs[rank], outs[rank+1]): regenerate and compare
        want = torch.empty(outs[rank + 1] - outs[rank], dtype=torch.uint8, device=dev)
        gen(want, outs[rank], outs[rank + 1])
        stitched_ok = bool(torch.equal(back[:mine[3]], want))
        del want
    flags = sc.all_gather_ints([int(stitched_ok)])
    stitched_ok = all(f[0] for f in flags)
    t_dec = timed(dec_step, args.strong_steps, lambda: sc.decode_finish(base))
    moved_all = sum(r[0] for r in sc.all_gather_ints([int(moved)]))
    sc.close()
    del comp, back, buf
    torch.cuda.empty_cache()
    return {
        "scaling": "strong",
        "workload": f"{args.strong_gib} GiB Zipf(1.1)/{args.nsym} in {bs} B blocks as ONE job over the GPUs"
                    + ("" if world > 1 else f" (N=1: the first {job >> 30} GiB as a sample, the rest does not fit beside it)"),
        "job_bytes": job, "stream_bytes": csize, "n_gpus": world,
        "encode_gbs": job * args.strong_steps / t_enc / GB,
        "decode_gbs": job * args.strong_steps / t_dec / GB,
        "value": 2 * job * args.strong_steps / (t_enc + t_dec) / GB,
        "encode_frac_of_n_peaks": (job + csize) * args.strong_steps / t_enc / GB / (peak * world),
        "decode_frac_of_n_peaks": (job + csize) * args.strong_steps / t_dec / GB / (peak * world),
        "stitched_ok": stitched_ok,
        "stream_checksum": [csize, sum(r[0] for r in sums), sum(r[1] for r in sums) % (1 << 61)],
        "first_half_checksum": first_half,
        "layout_by_byte_range_ms": 1e3 * t_layout, "bytes_moved_between_gpus": moved_all,
        "decode_split": "byte ranges of the one stream + 8 MiB overlap, header scan per range, host chain check",
        "collectives_in_timed_region": "none (sizes and (first, end, n) tuples only, outside it)",
    }


def e2e_leg(args, torch, dist, lib, dev, world, host: bytes, csize: int, barrier):
    """The same round trip through the reference-facing C API with HOST buffers: huf_encode then
    huf_decode over huf_memopen streams; host<->device copies are inside the timed region.  Two
    protocols: streams opened once and rewound per step (steady state, what a streaming caller
    like the Python compressor does) and streams opened fresh for every step (first call: the
    output pages have never been touched)."""
    from libhuffman_b200.capi import Config
    n = len(host)
    bs = args.blocksize
    cap = lib.dll.huf_b200_encode_bound(n, bs)

    def one(src, mid, dst, check):
        barrier()
        t0 = time.perf_counter()
        cfg = Config(length=n, blocksize=bs, reader=src.rw, writer=mid.rw)
        rc = lib.dll.huf_encode(C.byref(cfg))
        assert rc == 0, rc
        clen = len(mid)
        cfg = Config(length=clen, reader=mid.rw, writer=dst.rw)
        rc = lib.dll.huf_decode(C.byref(cfg))
        assert rc == 0, rc
        dt = time.perf_counter() - t0
        assert clen == csize and len(dst) == n
        if check:   # (untimed) the decoded bytes, not just their count
            assert C.string_at(dst.buf, n) == host, "e2e round trip differs from the input"
        return dt

    def reduce_max(t):
        tt = torch.tensor([t], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # steady state: one set of streams, rewound (input re-written untimed) every step
    src, mid, dst = lib.memstream(n), lib.memstream(cap), lib.memstream(n)
    t_reuse = 0.0
    # (five untimed rounds: streams that keep being used are page-locked by the library at their
    # fourth codec call -- 0.3 s per GiB once -- and used in place from then on)
    for it in range(-5, args.e2e_steps):
        for s_ in (src, mid, dst):
            lib.dll.huf_memrewind(s_.rw)
        src.write(host)
        dt = one(src, mid, dst, check=it == -1)
        if it >= 0:
            t_reuse += dt
    for s_ in (src, mid, dst):
        s_.close()
    # first call: fresh streams every step
    t_fresh = 0.0
    for it in range(args.e2e_steps):
        src, mid, dst = lib.memstream(n), lib.memstream(cap), lib.memstream(n)
        src.write(host)
        t_fresh += one(src, mid, dst, check=False)
        for s_ in (src, mid, dst):
            s_.close()
    t_reuse, t_fresh = reduce_max(t_reuse), reduce_max(t_fresh)
    return {"value": 2 * n * world * args.e2e_steps / t_reuse / GB, "unit": "GB/s",
            "h2d_bytes_per_step": n + csize, "d2h_bytes_per_step": csize + n,
            "protocol": "huf_memopen streams opened once, rewound and refilled (untimed) per step; 5 untimed warm-up rounds; "
                        "the library page-locks stream buffers that live through 4 codec calls (HUF_B200_PIN_AFTER) and "
                        "then copies straight from / into them",
            "direct_copies": int(lib.dll.huf_b200_direct_copy_count()),
            "fresh_streams_value": 2 * n * world * args.e2e_steps / t_fresh / GB,
            "fresh_streams_protocol": "three new huf_memopen streams per step: output pages are first-touched inside the timed region",
            "api": "huf_encode + huf_decode over huf_memopen streams (host buffers of the caller; spans of 32 MiB pipelined: "
                   "H2D, kernels, D2H overlap; fresh streams go through the library's pinned bounce buffers with a "
                   "threaded host copy on either side)",
            "copy_threads": os.environ.get("HUF_B200_COPY_THREADS", "cores - 2")}


def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and "HUF_B200_COPY_THREADS" not in os.environ:
        # the ranks of one box share its cores: the copy pool of every rank gets its share
        os.environ["HUF_B200_COPY_THREADS"] = str(max(2, (os.cpu_count() or 8) // world))

    import libhuffman_b200
    from libhuffman_b200 import datagen
    from libhuffman_b200.capi import DeviceCodec

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if args.nsym > 255:
        os.environ["HUF_B200_ACCEPT_1025"] = "1"   # Q1/Q2: opt-in so 256-symbol blocks round-trip

    lib = libhuffman_b200.load()
    enc_ctx = DeviceCodec(lib, local)
    dec_ctx = DeviceCodec(lib, local, accept_1025=args.nsym > 255)
    n = args.mib << 20
    bs = args.blocksize
    x = datagen.zipf_torch(n, dev, args.nsym, seed=2 + rank)
    cap = enc_ctx.encode_bound(n, bs)
    comp = torch.empty(cap, dtype=torch.uint8, device=dev)
    back = torch.empty(n + 64, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    barrier, timed = _device_timer(torch, dist, world, dev)

    # one checked round trip: sizes, parity of the data path used below
    enc_ctx.encode_async(x.data_ptr(), n, bs, comp.data_ptr(), cap, stream)
    csize = enc_ctx.encode_finish()
    dec_ctx.decode_async(comp.data_ptr(), csize, csize, back.data_ptr(), n + 64, stream)
    rc, m, used = dec_ctx.decode_finish()
    assert (rc, m, used) == (0, n, csize), (rc, m, used)
    assert torch.equal(back[:n], x), "round trip mismatch"
    launches_per_step = enc_ctx.launches() + dec_ctx.launches()

    def enc_step():
        enc_ctx.encode_async(x.data_ptr(), n, bs, comp.data_ptr(), cap, stream)

    def dec_step():
        dec_ctx.decode_async(comp.data_ptr(), csize, csize, back.data_ptr(), n + 64, stream)

    def both():
        enc_step()
        dec_step()

    def finish():
        assert enc_ctx.encode_finish() == csize
        r = dec_ctx.decode_finish()
        assert r == (0, n, csize), r

    for _ in range(max(3, args.warmup)):
        both()
        finish()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_rt = timed(both, args.steps, finish)
    t_enc = timed(enc_step, args.steps, lambda: enc_ctx.encode_finish())
    t_dec = timed(dec_step, args.steps, lambda: dec_ctx.decode_finish())
    # decode with the encoder's block-offset array as an index: reported next to
    # the headline, never instead of it -- the headline decodes a bare stream and has to find the blocks
    enc_step()
    enc_ctx.encode_finish()
    off_ptr, off_n = enc_ctx.block_offsets()

    def dec_step_hinted():
        dec_ctx.decode_hint_offsets(off_ptr, off_n)
        dec_step()

    t_dech = timed(dec_step_hinted, args.steps, lambda: dec_ctx.decode_finish())
    clocks = sampler.stop() if rank == 0 else None

    # per-kernel durations (CUDA events on the launching stream, separate pass) for the roofline
    kt: dict[str, list[float]] = {}
    for ctx, fn, fin in ((enc_ctx, enc_step, lambda: enc_ctx.encode_finish()),
                         (dec_ctx, dec_step, lambda: dec_ctx.decode_finish())):
        ctx.set_kernel_timing(True)
        for _ in range(3):
            fn()
            fin()
            per_call: dict[str, float] = {}
            for name, ms in ctx.kernel_times():   # (a kernel may be launched once per pass: sum them)
                per_call[name] = per_call.get(name, 0.0) + ms
            for name, ms in per_call.items():
                kt.setdefault(name, []).append(ms)
        ctx.set_kernel_timing(False)
    kavg = {k: sum(v) / len(v) for k, v in kt.items()}

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s"

    # end to end through the C API with host buffers (memory streams)
    e2e = None
    if args.e2e_steps > 0:
        host = x.cpu().numpy().tobytes()
        e2e = e2e_leg(args, torch, dist, lib, dev, world, host, csize, barrier)
        del host
    enc_ctx.close()
    dec_ctx.close()
    del x, comp, back
    torch.cuda.empty_cache()

    # extra legs: the other configs of BASELINE.json, each a driver-visible key of the same line
    extra = {}
    if world == 1 and not args.no_extra:
        extra["sweep"] = sweep_leg(torch, lib, dev, peak, args.sweep_mib)
        extra["config5_python_streaming"] = config5_leg(args.config5_mib)
    strong = None
    if not args.no_strong:
        strong = strong_leg(args, torch, dist, lib, dev, rank, world, peak)

    if rank == 0:
        # dominant kernel of the step and its algorithmic bytes: N + C either way
        dom = max(kavg, key=kavg.get) if kavg else None
        algo = float(n + csize)
        roof = None
        if dom:
            achieved = algo / (kavg[dom] / 1e3) / GB
            enc_gbs_algo = (n + csize) * args.steps / t_enc / GB
            dec_gbs_algo = (n + csize) * args.steps / t_dec / GB
            roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": measured_traffic(dom), "peak_source": peak_src,
                    "peak_nominal": NOMINAL_HBM_GBS, "frac_nominal": achieved / NOMINAL_HBM_GBS,
                    "algorithmic_bytes_per_launch": algo,
                    "kernel_ms": {k: round(v, 4) for k, v in sorted(kavg.items())},
                    # (per GPU: n and csize are one rank's bytes, the time is the slowest rank's)
                    "encode_path_frac": enc_gbs_algo / peak, "decode_path_frac": dec_gbs_algo / peak,
                    "encode_path_frac_nominal": enc_gbs_algo / NOMINAL_HBM_GBS,
                    "decode_path_frac_nominal": dec_gbs_algo / NOMINAL_HBM_GBS}
        line = {
            "metric": METRIC,
            "value": 2 * n * world * args.steps / t_rt / GB,
            "unit": "GB/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(3, args.warmup),
            "ms_per_step": 1e3 * t_rt / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "blocksize": bs, "blocks_per_gpu": n // bs, "compressed_bytes_per_gpu": csize,
                       "ratio": csize / n, "parallelism": f"block ranges x{world}, no collective",
                       "residency": "device resident (inputs in HBM before the timed region)",
                       "l2": "inputs larger than L2 (1 GiB in, ~0.9 GiB stream)",
                       "decoder_mode": "strict (reference parity)" if args.nsym <= 255 else "accept_1025 opt-in"},
            "encode_gbs": n * world * args.steps / t_enc / GB,
            "decode_gbs": n * world * args.steps / t_dec / GB,
            "decode_with_block_index_gbs": n * world * args.steps / t_dech / GB,
            "roofline": roof,
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps,
            "strong": strong,
        }
        line.update(extra)
        if not args.no_cpu_baseline and world == 1:
            mib = args.cpu_sample_mib
            kind, total, per_round = cpu_reference_run(1, mib, args.nsym, bs, 1)
            e, d = per_round[0]
            line["cpu_baseline"] = {"value": 2 * total / (e + d) / GB, "unit": "GB/s", "cores": 1, "kind": kind,
                                    "sample": f"{mib} MiB of the same Zipf source, one thread, bufio 64 KiB",
                                    "encode_gbs": total / e / GB, "decode_gbs": total / d / GB}
        print(json.dumps(line), flush=True)

    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
