"""CPU-only checks of the drop-in boundary: libgffm.so loads, exports every symbol include/gffm.h declares, the ctypes
mirror covers them all, and the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gffm.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gffm_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    import gffm_b200 as g
    lib = ctypes.CDLL(g.capi.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 45
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gffm.h but not exported"


def test_ctypes_mirror_covers_header():
    import gffm_b200 as g
    names = set(declared_functions())
    bound = set(g.capi.SIGNATURES) | set(g.capi.STRING_FUNCS)
    assert names == bound, (names - bound, bound - names)


def test_no_cpu_fallback():
    import torch
    import gffm_b200 as g
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert g.capi.device_count() == 0
    with pytest.raises(g.GffmError) as ei:
        g.Context(0)
    assert ei.value.code == g.capi.ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gpufinitefieldmatrices.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.replace("oracle/", "").lower() or f.endswith(".md"), f"{f} mentions the oracle"


def test_julia_shim_binds_every_symbol():
    shim = os.path.join(ROOT, "gpufinitefieldmatrices.jl_b200", "julia", "GPUFiniteFieldMatricesB200.jl")
    txt = open(shim).read()
    used = set(re.findall(r":(gffm_[a-z0-9_]+)", txt))
    missing = set(declared_functions()) - used
    assert not missing, f"Julia shim does not ccall: {sorted(missing)}"


def _build_c_client(tmp_path):
    import subprocess
    import gffm_b200 as g
    exe = str(tmp_path / "c_client")
    libdir = os.path.dirname(g.capi.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tools", "c_client.c"), "-L" + libdir, "-lgffm", "-Wl,-rpath," + libdir, "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return subprocess.run([exe], capture_output=True, text=True)


def test_plain_c_client_compiles_links_and_fails_loudly_without_gpu(tmp_path):
    """include/gffm.h is valid ISO C (no C++ types anywhere in the boundary) and libgffm.so links from gcc; without a GPU the
    client reports GFFM_ERR_NO_DEVICE from gffm_create (tools/c_client.c)."""
    import torch
    run = _build_c_client(tmp_path)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "libgffm" in run.stdout
    if not torch.cuda.is_available():
        assert "no CPU fallback" in run.stdout
    else:
        assert "C = [" in run.stdout


@pytest.mark.gpu
def test_plain_c_client_runs_on_the_gpu(tmp_path):
    """The same plain-C program on a B200: creates a context, uploads, multiplies through gffm_gemm, downloads and prints the product
    (the reference's 2x3 * 3x2 mod 11 literal, test/CuModMatrix/matmul_operations_test.jl:14-67)."""
    run = _build_c_client(tmp_path)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "C = [" in run.stdout and "no CPU fallback" not in run.stdout


def test_build_reuses_objects_by_content_hash_not_by_file_time(tmp_path):
    """build.py decides by the SHA-256 of source + headers + flags + compiler version (build/manifest.json): an untouched tree reuses every
    object whatever the file times say, a changed input hash recompiles exactly that object (checked on the decision, without running nvcc)."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("gffm_build_t", os.path.join(ROOT, "gpufinitefieldmatrices.jl_b200", "build.py"))
    bm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bm)
    bm.build()  # brings the tree up to date if it is not (the driver's build() has normally done that already)
    os.utime(os.path.join(bm.CSRC, "gemv.cu"))  # newer file time, same content
    bm.build()
    rep = bm.last_build_report()
    assert rep["compiled"] == [] and sorted(rep["reused"]) == sorted(bm.SOURCES) and rep["linked"] is False
    man = json.load(open(bm.MANIFEST))
    assert set(man["objects"]) == set(bm.SOURCES) and all(len(h) == 64 for h in man["objects"].values())
    info = json.load(open(os.path.join(bm.LIBDIR, "build_info.json")))
    assert info["object_inputs_sha256"] == man["objects"] and "compute_100a" in " ".join(info["flags"])


def test_mg_owner_ranges_is_a_pure_function():
    """Column ranges of B owned by the ranks (gffm_mg_owner_ranges): equal widths, multiples of 256, covering [0, n) -- no GPU needed."""
    import gffm_b200 as g
    assert g.multigpu.owner_ranges(16384, 8) == [2048 * q for q in range(9)]
    assert g.multigpu.owner_ranges(32768, 4) == [8192 * q for q in range(5)]
    assert g.multigpu.owner_ranges(1000, 4) == [0, 256, 512, 768, 1000]
    assert g.multigpu.owner_ranges(300, 4) == [0, 256, 300, 300, 300]
    assert g.multigpu.owner_ranges(0, 3) == [0, 0, 0, 0]
    with pytest.raises(g.GffmError):
        g.multigpu.owner_ranges(10, 0)
    # root-free ranges (peer-memory transports from 6 ranks on): the root owns nothing, 256-column blocks dealt to the others
    assert g.multigpu.owner_ranges_root_free(16384, 8, 0) == [0, 0, 2560, 4864, 7168, 9472, 11776, 14080, 16384]
    assert g.multigpu.owner_ranges_root_free(16384, 8, 7) == [0, 2560, 4864, 7168, 9472, 11776, 14080, 16384, 16384]
    assert g.multigpu.owner_ranges_root_free(1000, 2, 1) == [0, 1000, 1000]
    for n, nr, root in [(300, 4, 2), (0, 3, 1), (32768, 8, 3), (5000, 6, 5), (257, 32, 31)]:
        off = g.multigpu.owner_ranges_root_free(n, nr, root)
        assert off[0] == 0 and off[-1] == n and off[root] == off[root + 1] and all(b >= a for a, b in zip(off, off[1:]))
        assert all(o % 256 == 0 or o == n for o in off)
        widths = [b - a for q, (a, b) in enumerate(zip(off, off[1:])) if q != root]
        assert max(widths) - min(widths) <= 256 or n < 256 * (nr - 1)
    with pytest.raises(g.GffmError):
        g.multigpu.owner_ranges_root_free(10, 1, 0)
    assert len(g.multigpu.MultiGpu.unique_id()) == 128  # NCCL is dlopen'ed on demand, no link-time dependency


def test_enum_values_agree_between_header_ctypes_mirror_and_julia_shim():
    """Every enumerator of include/gffm.h that the ctypes mirror (capi.py) or the Julia shim names carries the header's value -- the
    boundary passes plain integers, so a renumbered enum would silently select another operation / transport."""
    import gffm_b200 as g
    hdr = open(os.path.join(ROOT, "include", "gffm.h")).read()
    enums = {name: int(val) for name, val in re.findall(r"\b(GFFM_[A-Z0-9_]+)\s*=\s*(-?\d+)", hdr)}
    assert enums["GFFM_MG_P2P_RAW"] == 5 and enums["GFFM_MG_DISTRIBUTED"] == -1 and len(enums) > 40
    checked = 0
    for name, val in enums.items():
        short = name[len("GFFM_"):]
        if hasattr(g.capi, short):
            assert getattr(g.capi, short) == val, name
            checked += 1
    assert checked >= 25, checked
    jl = open(os.path.join(ROOT, "gpufinitefieldmatrices.jl_b200", "julia", "GPUFiniteFieldMatricesB200.jl")).read()
    for jname, jval in re.findall(r"const (MG_[A-Z0-9_]+) = (-?\d+)", jl):
        assert enums["GFFM_" + jname] == int(jval), jname
