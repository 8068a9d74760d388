#!/usr/bin/env python
"""bench.py — BASELINE.json's headline metric on its headline config.

A step = one pass of the hot path over one batch: huf_encode of 1 GiB of Zipf(1.1) bytes at
64 KiB blocks followed by huf_decode of the resulting stream (configs[1]).  `value` is
uncompressed GB/s of that round trip with inputs resident in HBM (2N / t_step, GB = 1e9 B),
`e2e` the same through the reference-facing C API (huf_encode / huf_decode over memory
streams, host buffers, host<->device copies inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Under torchrun (N > 1) every rank codes its own 1 GiB shard (contiguous block ranges of an
N GiB input; blocks are independent, so there is no data-path collective): weak scaling.
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
        "config": {"workload": workload_name(args), "blocksize": args.blocksize,
                   "blocks_per_gpu": (args.mib << 20) // args.blocksize,
                   "parallelism": f"{cores} host processes over block ranges (reference CPU path)",
                   "sample": f"each step codes a bounded sample of the workload: {mib} MiB per process"},
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

def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    import libhuffman_b200
    from libhuffman_b200 import datagen
    from libhuffman_b200.capi import DeviceCodec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
            for name, ms in ctx.kernel_times():
                kt.setdefault(name, []).append(ms)
        ctx.set_kernel_timing(False)
    kavg = {k: sum(v) / len(v) for k, v in kt.items()}

    # end to end through the C API with host buffers (memory streams)
    e2e = None
    h2d = d2h = 0
    if args.e2e_steps > 0:
        from libhuffman_b200.capi import Config
        host = x.cpu().numpy().tobytes()
        t_e2e = 0.0
        clen = 0
        for it in range(-1, args.e2e_steps):   # it == -1: one untimed warm-up (context, arenas, pinned buffers)
            # untimed: put the step's input into a host memory stream, as a caller would have it
            src = lib.memstream(n)
            src.write(host)
            mid = lib.memstream(cap)
            dst = lib.memstream(n)
            barrier()
            t0 = time.perf_counter()
            cfg = Config(length=n, blocksize=bs, reader=src.rw, writer=mid.rw)
            rc = lib.dll.huf_encode(C.byref(cfg))
            assert rc == 0, rc
            clen = len(mid)
            cfg = Config(length=clen, reader=mid.rw, writer=dst.rw)
            rc = lib.dll.huf_decode(C.byref(cfg))
            assert rc == 0, rc
            torch.cuda.synchronize()
            if it >= 0:
                t_e2e += time.perf_counter() - t0
            ok = len(dst) == n
            if ok and it < 0:
                # (untimed) the decoded bytes, not just their count: memcmp against the input
                ok = C.string_at(dst.buf, n) == host
            for s_ in (src, mid, dst):
                s_.close()
            assert ok, "e2e round trip through huf_encode/huf_decode differs from the input"
        del host
        assert clen == csize
        tt = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
        barrier()
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt.item())
        e2e = 2 * n * world * args.e2e_steps / t_e2e / GB
        h2d = n + csize
        d2h = csize + n

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s"
        # dominant kernel of the step and its algorithmic bytes: N + C either way
        dom = max(kavg, key=kavg.get) if kavg else None
        algo = float(n + csize)
        roof = None
        if dom:
            achieved = algo / (kavg[dom] / 1e3) / GB
            roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": measured_traffic(dom), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": algo,
                    "kernel_ms": {k: round(v, 4) for k, v in sorted(kavg.items())},
                    "encode_path_frac": (n + csize) * args.steps / t_enc / GB / peak,
                    "decode_path_frac": (n + csize) * args.steps / t_dec / GB / peak}
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
            "e2e": {"value": e2e, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "huf_encode + huf_decode over huf_memopen streams (pageable host buffers, pinned bounce buffers inside the library)"},
            "gpu_launches": launches_per_step * args.steps,
        }
        if not args.no_cpu_baseline and world == 1:
            mib = args.cpu_sample_mib
            kind, total, per_round = cpu_reference_run(1, mib, args.nsym, bs, 1)
            e, d = per_round[0]
            line["cpu_baseline"] = {"value": 2 * total / (e + d) / GB, "unit": "GB/s", "cores": 1, "kind": kind,
                                    "sample": f"{mib} MiB of the same Zipf source, one thread, bufio 64 KiB",
                                    "encode_gbs": total / e / GB, "decode_gbs": total / d / GB}
        print(json.dumps(line), flush=True)

    enc_ctx.close()
    dec_ctx.close()
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
