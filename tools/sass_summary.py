"""Counts the Blackwell-specific SASS mnemonics per kernel of the built libgffm.so (cuobjdump -sass) and writes
profiles/<tag>_sass_summary.txt -- committed evidence for DESIGN.md's tcgen05 / TMA / cluster claims (the .so itself is git-ignored)."""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gpufinitefieldmatrices.jl_b200", "lib", "libgffm.so")
PATTERNS = ["UTCIMMA.2CTA", "UTCIMMA", "UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "STAS", "SYNCS", "CREDUX", "REDUX", "IDP.4A", "IDP.2A",
            "UCGABAR", "ACQBULK", "LDG.E.128", "STG.E.128", "LDS.128", "STS.128", "MEMBAR", "ERRBAR"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        per[cur]["_total"] += 1
        for p in PATTERNS:
            if op.startswith(p):
                per[cur][p] += 1
                break
    demangle = subprocess.run(["c++filt", "-p"], input="\n".join(per.keys()), capture_output=True, text=True).stdout.splitlines()
    out = [f"# cuobjdump -sass summary of libgffm.so (sha256 {hashlib.sha256(open(LIB, 'rb').read()).hexdigest()[:16]}, sm_100a)",
           "# per kernel: total SASS instructions, then counts of Blackwell-specific / wide-access mnemonics (prefix match)", ""]
    tot = collections.Counter()
    for (name, cnt), dn in zip(per.items(), demangle):
        dn = dn.replace("(anonymous namespace)::", "").replace("void ", "")
        marks = "  ".join(f"{p}={cnt[p]}" for p in PATTERNS if cnt[p])
        out.append(f"{dn:<70s} {cnt['_total']:>7d}  {marks}")
        tot.update(cnt)
    out.append("")
    out.append("TOTAL " + "  ".join(f"{p}={tot[p]}" for p in PATTERNS if tot[p]) + f"  kernels={len(per)} instructions={tot['_total']}")
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt")
    open(path, "w").write("\n".join(out) + "\n")
    print(out[-1])
    print(path)


if __name__ == "__main__":
    main()
