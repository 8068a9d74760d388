"""Turn the files of a `gpurun` capture (gpurun_out/) into the tracked summaries under profiles/.
usage: make_profiles.py TAG      (e.g. r01_e)"""
import collections, csv, json, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
G, P = ROOT / "gpurun_out", ROOT / "profiles"
tag = sys.argv[1]

def summary(rep):
    return subprocess.run([sys.executable, str(ROOT / "scripts" / "ncu_summary.py"), str(rep)], capture_output=True, text=True).stdout

# launch list
rows = [r for r in csv.reader(open(G / "launches.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("hufb200::", "")
    if name.startswith("k_"):
        d = agg.setdefault(name, [0, 0.0]); d[0] += 1; d[1] += float(r[-1])
tot = sum(v[1] for v in agg.values())
with open(P / f"{tag}_launches.csv", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400, python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline\n")
    f.write(f"# per-launch device time under ncu (cold cache, serialised): compare SHARES with kernel_ms of profiles/{tag}_bench.json\n")
    f.write("kernel,launches,avg_us,share_pct\n")
    for k, (n, t) in agg.items():
        f.write(f"{k},{n},{t / n / 1000:.1f},{100 * t / tot:.1f}\n")
# bench + sweep
(P / f"{tag}_bench.json").write_text((G / "bench.log").read_text())
_line = json.loads([l for l in open(G / "bench.log") if l.startswith("{")][0])
(P / f"{tag}_sweep.jsonl").write_text("".join(json.dumps(r) + "\n" for r in _line.get("sweep", [])))
# ncu full
src = subprocess.run(["ncu", "-i", str(G / "prof_dec.ncu-rep"), "--page", "source", "--print-source", "cuda,sass", "--csv",
                      "--kernel-name", "k_decode"], capture_output=True, text=True).stdout
Path("/tmp/src_dec.csv").write_text(src)
phase = subprocess.run([sys.executable, str(ROOT / "scripts" / "ncu_phase.py"), "/tmp/src_dec.csv",
                        str(ROOT / "libhuffman_b200/csrc/cuda/dec_fast.cuh"), "win:The three staged words from",
                        "look4:Four consecutive table entries", "step:One exact decode step", "ctasync:CTA barrier behind divergent",
                        "regcopy:Copy symbols [s0, s0 + n)", "kernel_head:k_decode(DecArgs a)", "lutfill:lookup table from the ordered",
                        "chunkhead:// ---- chunk loop", "stage:(0) stage the chunk", "warm:(1) warm-up in front",
                        "walk+verify:(1b) decode my sub-block", "scan:(3) symbol-count scan", "compact:(4) compaction into",
                        "tail:if (status == kOk && end_bit"], capture_output=True, text=True).stdout
(P / f"{tag}_ncu_full.md").write_text(
    "# ncu --set full --clock-control none, one launch of each kernel (bench.py --steps 1 --warmup 1)\n\n## encode\n\n"
    + summary(G / "prof.ncu-rep") + "\n## decode\n\n" + summary(G / "prof_dec.ncu-rep")
    + "\n## k_decode by source phase (share of executed warp instructions / of stall samples)\n\n```\n" + phase + "```\n")
d = json.loads([l for l in open(G / "bench.log") if l.startswith("{")][0])
km = d["roofline"]["kernel_ms"]; t = sum(km.values())
print({k: round(100 * v / t, 1) for k, v in km.items()})
print("enc", d["encode_gbs"], "dec", d["decode_gbs"], "value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"],
      d["roofline"]["encode_path_frac"], d["roofline"]["decode_path_frac"], "ms", d["ms_per_step"])
