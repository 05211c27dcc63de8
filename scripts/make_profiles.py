"""Turn the raw round-end captures in gpurun_out/ (scripts/profile_round.sh) into the tracked summaries
under profiles/.  Needs ncu (no GPU).   python scripts/make_profiles.py r1"""
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import defaultdict

R = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# ---- launch list of the bench command
rows = [r for r in csv.reader(open(os.path.join(G, f"{R}_bench_launches_raw.csv"))) if len(r) > 14 and r[0].isdigit()]
launches = [(int(r[0]), r[4], float(r[14].replace(",", ""))) for r in rows if r[12] == "gpu__time_duration.sum"]
unit = next((r[13] for r in rows if r[12] == "gpu__time_duration.sum"), "ns")
scale = {"ns": 1.0, "us": 1e3, "usecond": 1e3, "nsecond": 1.0, "ms": 1e6, "msecond": 1e6}.get(unit, 1.0)
with open(os.path.join(P, f"{R}_bench_launches.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 600 python bench.py --steps 5 --warmup 3 --no-attack --no-cpu-baseline\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
    f.write("id,kernel,gpu__time_duration_ns\n")
    for i, k, t in launches:
        f.write(f'{i},"{k[:110]}",{t * scale:.0f}\n')


def shares(sel, title, out):
    agg = defaultdict(lambda: [0, 0.0])
    for _, k, t in launches:
        if sel(k):
            agg[k][0] += 1
            agg[k][1] += t * scale
    tot = sum(v[1] for v in agg.values()) or 1.0
    out.write(title + "\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        out.write(f"{100 * t / tot:6.2f}%  n={n:3d}  avg={t / n / 1e3:8.1f} us  {k[:100]}\n")
    out.write("\n")


with open(os.path.join(P, f"{R}_bench_kernel_shares.txt"), "w") as f:
    shares(lambda k: "sampler_" in k, "shares of the sampler step (the timed region of bench.py) from the ncu launch list", f)
    shares(lambda k: "allpairs_tc" in k or "lookup_fwd" in k or "prep_kmajor" in k or "lookup_convc1" in k or "volgrad" in k,
           "shares of the RAFT build + lookup from the ncu launch list", f)
    shares(lambda k: True, "all kernels of the bench command (first 600 launches)", f)

# ---- full captures
summ = os.path.join(ROOT, "scripts", "ncu_summary.py")
for name, what in (("sampler_full", "sampler"), ("raft_full", "raft: all-pairs + lookup forward"),
                   ("raft_aux_full", "raft: lookup backward, volume backward (tcgen05 GEMMs), alt_cuda_corr forward / backward"),
                   ("lookup_convc1_full", "lookup fused with the motion encoder's 1x1 convolution (tcgen05)"),
                   ("merge_full", "fused FlowNetC merge block: forward kernel (1/C + LeakyReLU epilogue, concat slice) + backward pre-pass")):
    rep = os.path.join(G, f"{R}_{name}.ncu-rep")
    if not os.path.exists(rep):
        continue
    txt = subprocess.run([sys.executable, summ, rep], capture_output=True, text=True).stdout
    with open(os.path.join(P, f"{R}_{name}_summary.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on (B200, round {R[1:]}, final kernels) -- {what}\n")
        f.write(txt)

# ---- DRAM traffic of the sampler kernels per launch (bench.py's roofline.traffic)
rep = os.path.join(G, f"{R}_sampler_full.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units, data = rr[0], rr[1], rr[2:]

    def val(row, key):
        i = hdr.index(key)
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        return float(row[i].replace(",", "")) * mult

    ki = hdr.index("Kernel Name")
    fwd = [r for r in data if "sampler_fwd" in r[ki]]
    bwd = [r for r in data if "sampler_bwd" in r[ki]]
    tr = {}
    if fwd:
        tr["sampler_fwd_kernel_dram_bytes_per_launch"] = sum(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in fwd) / len(fwd)
    if bwd:
        tr["sampler_bwd_kernel_dram_bytes_per_launch"] = sum(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in bwd) / len(bwd)
    tr["source"] = f"profiles/{R}_sampler_full_summary.txt (ncu --set full, B=8, 256x48x160)"
    json.dump(tr, open(os.path.join(P, "sampler_traffic.json"), "w"), indent=1)

# ---- DRAM traffic of the RAFT kernels per launch (bench.py's raft.roofline_build / roofline_lookup .traffic)
rep = os.path.join(G, f"{R}_raft_full.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units, data = rr[0], rr[1], rr[2:]

    def val2(row, key):
        i = hdr.index(key)
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        return float(row[i].replace(",", "")) * mult

    ki = hdr.index("Kernel Name")
    tr = {"source": f"profiles/{R}_raft_full_summary.txt (ncu --set full, B=4, 256x48x160; cold L2, one launch each)"}
    for key, pat in (("allpairs_tc_kernel", "allpairs_tc"), ("lookup_fwd_kernel", "lookup_fwd")):
        rows_k = [r for r in data if pat in r[ki]]
        if rows_k:
            tr[key + "_dram_read_bytes"] = sum(val2(r, "dram__bytes_read.sum") for r in rows_k) / len(rows_k)
            tr[key + "_dram_write_bytes"] = sum(val2(r, "dram__bytes_write.sum") for r in rows_k) / len(rows_k)
    json.dump(tr, open(os.path.join(P, "raft_traffic.json"), "w"), indent=1)

for f in (f"{R}_bench_line.json", f"{R}_vs_reference_cuda.json", f"{R}_sweep_cfg5.json", f"{R}_attack_timeline_g1.json",
          f"{R}_attack_timeline_g8.json", f"{R}_bench_line_g2.json", f"{R}_bench_line_g8.json"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f))
print(sorted(os.listdir(P)))
