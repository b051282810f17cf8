"""Compiles every CUDA source with `-Xptxas -v` (objects discarded) and writes the per-kernel resource usage -- registers, spills, static
shared memory, barriers -- to profiles/<tag>_ptxas_summary.txt: the check DESIGN.md's occupancy statements rest on (the library itself
is git-ignored)."""
import os
import re
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpufinitefieldmatrices.jl_b200"))
import build as B  # noqa: E402


def one(src):
    with tempfile.TemporaryDirectory() as td:
        r = subprocess.run([B.NVCC] + B.FLAGS + ["-Xptxas", "-v", "-c", os.path.join(B.CSRC, src), "-o", os.path.join(td, "x.o")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr)
    rows, cur = [], None
    for line in r.stderr.splitlines():
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
        if m:
            cur = {"name": m.group(1), "src": src, "spill_st": 0, "spill_ld": 0, "stack": 0, "regs": 0, "smem": 0, "bar": 0}
            rows.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            cur["stack"], cur["spill_st"], cur["spill_ld"] = map(int, m.groups())
        m = re.search(r"Used (\d+) registers", line)
        if m:
            cur["regs"] = int(m.group(1))
            b = re.search(r"used (\d+) barriers", line)
            cur["bar"] = int(b.group(1)) if b else 0
            sm = re.search(r"(\d+) bytes smem", line)
            cur["smem"] = int(sm.group(1)) if sm else 0
    return rows


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    with ThreadPoolExecutor(max_workers=4) as ex:
        rows = [r for rs in ex.map(one, B.SOURCES) for r in rs]
    names = subprocess.run(["c++filt", "-p"], input="\n".join(r["name"] for r in rows), capture_output=True, text=True).stdout.splitlines()
    out = [f"# nvcc {' '.join(B.FLAGS)} -Xptxas -v  ({B._nvcc_version()})",
           f"# {'kernel':<66s} {'source':<13s} {'regs':>4s} {'bar':>3s} {'static smem':>11s} {'stack':>6s} {'spill st/ld':>11s}"]
    for r, n in zip(rows, names):
        n = n.replace("(anonymous namespace)::", "").replace("void ", "")
        out.append(f"{n[:68]:<68s} {r['src']:<13s} {r['regs']:>4d} {r['bar']:>3d} {r['smem']:>11d} {r['stack']:>6d} {r['spill_st']:>5d}/{r['spill_ld']:<5d}")
    spilled = [n for r, n in zip(rows, names) if r["spill_st"] or r["spill_ld"]]
    out.append("")
    out.append(f"# kernels: {len(rows)}; with register spills: {len(spilled)}" + (" (" + ", ".join(s.replace('(anonymous namespace)::', '').replace('void ', '')[:50] for s in spilled) + ")" if spilled else ""))
    path = os.path.join(ROOT, "profiles", f"{tag}_ptxas_summary.txt")
    open(path, "w").write("\n".join(out) + "\n")
    print(out[-1])
    print(path)


if __name__ == "__main__":
    main()
