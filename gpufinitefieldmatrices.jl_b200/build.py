"""Builds libgffm.so (sm_100a only) in-tree: gpufinitefieldmatrices.jl_b200/lib/libgffm.so

An object is reused only when the SHA-256 of everything that went into it (its source, every header of csrc/ and include/, the nvcc
version and the flags) equals the one recorded in build/manifest.json -- not by file times, which a checkout or a copy to another
machine rewrites.  build() returns the library path; `last_build_report()` says which objects were compiled and which were reused, and
lib/build_info.json records the same next to the library (sources, hashes, flags, compiler) for whoever receives only the binary."""
import hashlib
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libgffm.so")
MANIFEST = os.path.join(BUILD, "manifest.json")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
SOURCES = ["api.cu", "gemm_tc.cu", "pluq.cu", "karatsuba.cu", "gemv.cu", "mg.cu", "wide.cu"]
_report = {}


def _sha(paths, extra=()):
    h = hashlib.sha256()
    for x in extra:
        h.update(x.encode())
        h.update(b"\0")
    for p in paths:
        h.update(os.path.basename(p).encode())
        h.update(b"\0")
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _nvcc_version():
    try:
        out = subprocess.run([NVCC, "--version"], capture_output=True, text=True).stdout
        return out.strip().splitlines()[-1]
    except OSError as ex:  # no compiler: only a prebuilt library can be used
        return f"unavailable ({ex})"


def last_build_report():
    """{"compiled": [...], "reused": [...], "linked": bool} of the most recent build() call in this process"""
    return dict(_report)


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    headers = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    headers.append(os.path.join(HERE, "..", "include", "gffm.h"))
    nvcc_version = _nvcc_version()
    try:
        manifest = json.load(open(MANIFEST))
    except (OSError, ValueError):
        manifest = {}
    objs, jobs, compiled, reused, hashes = [], [], [], [], {}
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(BUILD, s.replace(".cu", ".o"))
        objs.append(obj)
        hashes[s] = _sha([src] + headers, extra=[nvcc_version] + FLAGS)
        if force or not os.path.exists(obj) or manifest.get("objects", {}).get(s) != hashes[s]:
            jobs.append([NVCC] + FLAGS + ["-c", src, "-o", obj])
            compiled.append(s)
        else:
            reused.append(s)

    def run(cmd):
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        for out in ex.map(run, jobs):
            if verbose and out:
                print(out)
    link = bool(jobs) or not os.path.exists(LIB) or force or manifest.get("linked_from") != hashes
    if link:
        run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"])
    manifest = {"objects": hashes, "linked_from": hashes, "nvcc": nvcc_version, "flags": FLAGS}
    json.dump(manifest, open(MANIFEST, "w"), indent=1)
    info = {"library": os.path.basename(LIB), "library_sha256": _sha([LIB]), "nvcc": nvcc_version, "flags": FLAGS, "link": "nvcc -shared (cudart static)",
            "sources": {s: _sha([os.path.join(CSRC, s)]) for s in SOURCES}, "headers": {os.path.basename(h): _sha([h]) for h in headers},
            "object_inputs_sha256": hashes}
    json.dump(info, open(os.path.join(LIBDIR, "build_info.json"), "w"), indent=1)
    _report.clear()
    _report.update({"compiled": compiled, "reused": reused, "linked": link})
    if verbose:
        print(json.dumps(_report))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
