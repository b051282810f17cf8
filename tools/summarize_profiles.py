"""Turns the ncu artefacts under gpurun_out/ into the small text summaries committed under profiles/.
usage: python tools/summarize_profiles.py <tag-in-gpurun_out> <round-name>"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(csv_path, out_path, title):
    rows = list(csv.reader(l for l in open(csv_path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "us" else (v / 1e6 if r[ui] == "ns" else v)
        agg[name].append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out_path, "w") as f:
        f.write(f"# {title}\n# source: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# launches={len(rows)-1} total_kernel_ms={tot:.3f}\n")
        f.write(f"{'kernel':60s} {'n':>6s} {'sum_ms':>10s} {'avg_us':>10s} {'max_us':>10s} {'share%':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k[:60]:60s} {len(v):6d} {sum(v):10.3f} {sum(v)/len(v)*1e3:10.1f} {max(v)*1e3:10.1f} {100*sum(v)/tot:7.1f}\n")
    print("wrote", out_path)


def full(rep_path, out_path, title):
    raw = subprocess.run(["ncu", "-i", rep_path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {}
        for i, h in enumerate(hdr):
            if h == "Kernel Name" or h in KEYS:
                d[h] = (r[i], units[i])
        res.append(d)
    with open(out_path, "w") as f:
        f.write(f"# {title}\n# source: ncu --set full --clock-control none --import-source on ({os.path.basename(rep_path)})\n")
        for d in res:
            for k, (v, u) in d.items():
                f.write(f"{k} = {v} {u}\n")
            f.write("---\n")
    print("wrote", out_path)
    return res


if __name__ == "__main__":
    tag, rnd = sys.argv[1], sys.argv[2]
    os.makedirs(OUT, exist_ok=True)
    g = os.path.join(ROOT, "gpurun_out")
    for f in sorted(os.listdir(g)):
        if not f.startswith(tag):
            continue
        p = os.path.join(g, f)
        if f.endswith("launches.csv"):
            launches(p, os.path.join(OUT, f"{rnd}_{f[len(tag)+1:-4]}_summary.txt"), f"launch list {f}")
        elif f.endswith(".ncu-rep"):
            res = full(p, os.path.join(OUT, f"{rnd}_{f[len(tag)+1:-8]}.txt"), f"ncu full capture {f}")
            if "prof_gemm" in f and res:
                def gb(x):
                    v, u = x
                    v = float(v.replace(",", ""))
                    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}.get(u, 1)
                tr = [gb(d["dram__bytes_read.sum"]) + gb(d["dram__bytes_write.sum"]) for d in res]
                json.dump({"gemm_tc_kernel_rns_dram_bytes_per_launch": sum(tr) / len(tr), "source": f, "launches": len(tr)},
                          open(os.path.join(OUT, "roofline_traffic.json"), "w"), indent=1)
        elif f.endswith(("bench.json", "bench_ref.json", "pytest.txt", "smoke.txt", "gpu.txt")):
            open(os.path.join(OUT, f"{rnd}_{f[len(tag)+1:]}"), "w").write(open(p).read())
