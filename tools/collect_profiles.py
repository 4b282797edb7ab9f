#!/usr/bin/env python3
"""Turn the raw evidence in gpurun_out/ (tools/make_profiles.sh) into the tracked summaries under profiles/.
    python tools/collect_profiles.py r01"""
import collections
import csv
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src, dst = ROOT / "gpurun_out", ROOT / "profiles"
dst.mkdir(exist_ok=True)

# launch list -> share table
ll = src / f"{tag}_launches_bench.csv"
if ll.exists():
    rows = [r for r in csv.reader(open(ll)) if len(r) > 5]
    hdr = rows[0]
    iN, iV, iM, iU = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit"))
    agg, n = collections.OrderedDict(), collections.Counter()
    for r in rows[1:]:
        if r[iM] != "gpu__time_duration.sum":
            continue
        v = float(r[iV].replace(",", ""))
        v = v / 1000 if r[iU] == "ns" else (v * 1000 if r[iU] == "ms" else v)
        name = r[iN].split("(")[0][:90]
        agg[name] = agg.get(name, 0) + v
        n[name] += 1
    tot = sum(agg.values())
    with open(dst / f"{tag}_launch_shares_bench.txt", "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 3 --warmup 3 ({tag})\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"# total {tot:.1f} us over {sum(n.values())} launches\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1]):
            f.write(f"{v:10.1f} us {100 * v / tot:5.1f}%  x{n[k]:3d}  avg {v / n[k]:8.1f} us  {k}\n")
    shutil.copy(ll, dst / ll.name)

for rep, vox in ((f"{tag}_mind_noise", 2 * 192 ** 3), (f"{tag}_mind_clean", 2 * 192 ** 3), (f"{tag}_gin3333", 2 * 192 ** 3),
                 (f"{tag}_sampler", 2 * 128 ** 3), (f"{tag}_philox", 2 * 12 * 192 ** 3), (f"{tag}_closs", 2 * 128 ** 3)):
    path = src / f"{rep}.ncu-rep"
    if not path.exists():
        continue
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(path), str(vox)],
                         capture_output=True, text=True).stdout
    (dst / f"{rep}_ncu_summary.txt").write_text(
        f"# ncu --set full --clock-control none --import-source on  ({rep}); summarised by tools/ncu_summary.py\n" + out)

for name in (f"{tag}_bench.json", f"{tag}_bench_reference.json", f"{tag}_bench_tta.json", f"{tag}_kernel_times.txt",
             f"{tag}_eager_vs_ours.json", f"{tag}_nvidia_smi.csv", f"{tag}_sanitizer_memcheck.log",
             f"{tag}_sanitizer_memcheck_small.log", f"{tag}_sanitizer_racecheck.log"):
    if (src / name).exists():
        shutil.copy(src / name, dst / name)
for name, to in (("chain_errors.json", f"{tag}_chain_errors.json"), ("tta_dice.json", f"{tag}_tta_dice.json")):
    if (src / name).exists():
        shutil.copy(src / name, dst / to)
# per-phase split of the MIND capture (S1 / S2 / C between the kernel's BAR.SYNCs)
rep = src / f"{tag}_mind_noise.ncu-rep"
if rep.exists():
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_phases.py"), str(rep), str(2 * 192 ** 3), "6"],
                         capture_output=True, text=True).stdout
    (dst / f"{tag}_mind_noise_phases.txt").write_text(
        f"# tools/ncu_phases.py on the same capture as {tag}_mind_noise_ncu_summary.txt: segments between BAR.SYNCs (1 = S1, 2 = S2, 3 = C)\n" + out)
print(sorted(p.name for p in dst.iterdir()))
