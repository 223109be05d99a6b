"""Turn the raw ncu outputs of tools/profile.sh (gpurun_out/) into the small tracked files under profiles/.

usage: python tools/profile_summaries.py <tag>      e.g. r01b
  launches_<w>_<tag>.csv       -> profiles/launches_<w>_<tag>.csv + _summary.txt (per-kernel count, mean, share)
  pipes_<k>_<tag>.csv          -> profiles/ncu_pipes_<k>_<tag>.json
  prof_<k>_<tag>.ncu-rep       -> profiles/ncu_<k>_<tag>.json   (tools/ncu_summary.py)
"""
import collections, csv, glob, io, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def ncu_rows(path):
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    return list(csv.DictReader(io.StringIO("".join(lines))))


def launches(path, out_txt, cmd):
    agg = collections.OrderedDict()
    for r in ncu_rows(path):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(r["Metric Unit"], 1e-3)
        agg.setdefault(r["Kernel Name"], []).append(v)
    tot = sum(sum(v) for v in agg.values()) or 1.0
    with open(out_txt, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {cmd}\n"
                "# (cold-cache, serialised launches: compare shares, not absolutes)\n")
        for k, v in agg.items():
            f.write(f"{k[:70]:70s} n={len(v):3d} mean_us={sum(v) / len(v):10.2f} share={100 * sum(v) / tot:5.1f}%\n")


def pipes(path, out_json):
    rows = ncu_rows(path)
    if not rows:
        return
    d = {"kernel": rows[0]["Kernel Name"], "metrics": {}}
    for r in rows:
        d["metrics"][r["Metric Name"]] = {"value": r["Metric Value"].replace(",", ""), "unit": r["Metric Unit"]}
    json.dump(d, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    tag = sys.argv[1]
    for f in sorted(glob.glob(os.path.join(G, f"launches_*_{tag}.csv"))):
        base = os.path.basename(f)[:-4]
        shutil.copy(f, os.path.join(P, base + ".csv"))
        w = re.match(r"launches_(.*)_" + re.escape(tag), base).group(1)
        launches(f, os.path.join(P, base + "_summary.txt"), f"python bench.py --workload {w} --steps 2 --warmup 3")
    for f in sorted(glob.glob(os.path.join(G, f"pipes_*_{tag}.csv"))):
        pipes(f, os.path.join(P, "ncu_" + os.path.basename(f)[:-4] + ".json"))
    for f in sorted(glob.glob(os.path.join(G, f"prof_*_{tag}.ncu-rep"))):
        k = re.match(r"prof_(.*)_" + re.escape(tag), os.path.basename(f)).group(1)
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), f,
                        os.path.join(P, f"ncu_{k}_{tag}.json")], stdout=subprocess.DEVNULL)
    print("\n".join(sorted(os.listdir(P))))
