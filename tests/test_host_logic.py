"""Host-side logic and the C-ABI surface, no GPU needed: mesh generators, dof numbering, partitioning, golden fixtures,
and that libfinegpu.so loads and exports every symbol include/fegpu.h declares (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported_and_bound(fe):
    from finetools_jl_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "fegpu.h")).read()
    declared = set(re.findall(r"\b(fegpu_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert os.path.exists(_lib.LIB_PATH), "libfinegpu.so missing: run __graft_entry__.build()"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    _lib.lib()  # binds argtypes for all of them


def test_no_device_fails_loudly(fe):
    """No CPU fallback: creating a context without a CUDA device is an error, not a silent slow path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(fe.FEGPUError, match="no CUDA device"):
        fe.GPUContext(0)


def test_product_never_imports_oracle():
    """The product path may not import, link or execute anything under oracle/ (it may mention it in prose)."""
    pkg = os.path.join(ROOT, "finetools.jl_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|libfe_oracle|orc_[a-z_]+\s*\(|oracle/fe_oracle|#include\s+\".*oracle", re.M)
    for top in (pkg, os.path.join(ROOT, "finetools_jl_b200")):
        for dirpath, _, files in os.walk(top):
            for f in files:
                if f.endswith((".py", ".cu", ".h", ".jl")) or f == "Makefile":
                    src = open(os.path.join(dirpath, f)).read()
                    assert not bad.search(src), (dirpath, f)


def test_h8block_numbering(fe):
    fens, fes = fe.H8block(2.0, 3.0, 4.0, 2, 3, 4)
    assert fens.count() == 3 * 4 * 5 and fes.count() == 24
    # nodes x-fastest (MeshHexahedronModule.jl:76-86)
    np.testing.assert_allclose(fens.xyz[1], [1.0, 0.0, 0.0])
    np.testing.assert_allclose(fens.xyz[3], [0.0, 1.0, 0.0])
    # elements z-fastest (:88-105): element 2 sits above element 1
    np.testing.assert_array_equal(fes.conn[0], [1, 2, 5, 4, 13, 14, 17, 16])
    np.testing.assert_array_equal(fes.conn[1], fes.conn[0] + 12)
    # positive Jacobians: volume via the H8 mass identity is checked in test_oracle_pins


def test_refined_meshes_counts(fe):
    """Node-count closed forms of SURVEY.md appendix B; test/test_meshing.jl:3384-3394 style counts."""
    for n in (1, 2, 3):
        f, s = fe.H20block(1, 1, 1, n, n, n)
        assert f.count() == 4 * n ** 3 + 9 * n ** 2 + 6 * n + 1 and s.conn.shape == (n ** 3, 20)
        f, s = fe.H27block(1, 1, 1, n, n, n)
        assert f.count() == (2 * n + 1) ** 3 and s.conn.shape == (n ** 3, 27)
        f, s = fe.T10block(1, 1, 1, n, n, n)
        assert f.count() == (2 * n + 1) ** 3 and s.conn.shape == (6 * n ** 3, 10)
        f, s = fe.T4block(1, 1, 1, n, n, n, "ca")
        assert s.conn.shape == (5 * n ** 3, 4)
    # mid-edge nodes sit at edge midpoints, numbered in first-encounter order
    f, s = fe.H20block(2.0, 2.0, 2.0, 1, 1, 1)
    np.testing.assert_array_equal(s.conn[0, :8], [1, 2, 4, 3, 5, 6, 8, 7])
    np.testing.assert_array_equal(s.conn[0, 8:], np.arange(9, 21))
    np.testing.assert_allclose(f.xyz[8], [1.0, 0.0, 0.0])
    f, s = fe.T4block(1, 1, 1, 1, 1, 1)
    f10, s10 = fe.T4toT10(f, s)
    np.testing.assert_array_equal(s10.conn[0, 4:7], [9, 10, 11])
    np.testing.assert_allclose(f10.xyz[8], 0.5 * (f.xyz[s.conn[0, 0] - 1] + f.xyz[s.conn[0, 1] - 1]))


def test_meshboundary(fe):
    fens, fes = fe.H8block(1, 1, 1, 3, 4, 5)
    b = fe.meshboundary(fes)
    assert isinstance(b, fe.FESetQ4) and b.count() == 2 * (12 + 20 + 15)
    # lexicographic order of the sorted node ids
    key = np.sort(b.conn, axis=1)
    assert all(tuple(key[i]) < tuple(key[i + 1]) for i in range(len(key) - 1))
    fens, fes = fe.T4block(1, 1, 1, 2, 2, 2)
    bt = fe.meshboundary(fes)
    assert isinstance(bt, fe.FESetT3) and bt.count() == 2 * 6 * 4


def test_numberdofs_free_first_then_fixed(fe):
    u = fe.NodalField(np.zeros((5, 2)))
    fe.setebc(u, [2, 4], True, 1, 0.0)
    fe.setebc(u, [4], True, 2, 7.0)
    fe.numberdofs(u)
    # free dofs in node order, component inner; then the fixed ones (FieldModule.jl:360-377)
    np.testing.assert_array_equal(u.dofnums, [[1, 2], [8, 3], [4, 5], [9, 10], [6, 7]])
    assert u.nfreedofs() == 7 and u.nalldofs() == 10
    assert u.values[3, 1] == 7.0
    v = fe.gathersysvec(u)
    assert v[9] == 7.0


def test_linearspace_matches_exact_progression(fe):
    x = fe.linearspace(0.0, 1.1, 21)
    assert x[0] == 0.0 and x[-1] == 1.1 and len(x) == 21
    from fractions import Fraction
    assert x[7] == float(Fraction(11, 10) * 7 / 20)
    np.testing.assert_array_equal(fe.linearspace(0.0, 1.0, 129), np.arange(129) / 128.0)


def test_pointpartitioning_labels(fe):
    """test/test_miscellaneous2.jl:6-25, 63-75: every label 1..2^k is present."""
    fens, _ = fe.H8block(3.0, 1.0, 1.0, 12, 4, 4)
    for npart in (2, 4, 8):
        p = fe.pointpartitioning(fens.xyz, npart)
        assert sorted(np.unique(p)) == list(range(1, npart + 1))
        assert p.shape == (fens.count(),)
    own = fe.slab_owner(fens.count(), 4)
    assert (np.diff(own) >= 0).all() and own[0] == 0 and own[-1] == 3
    # planar point sets (_nodepartitioning2, MeshModificationModule.jl:886-951): a 4 x 1 strip is cut across its long direction
    f2, _ = fe.Q4block(4.0, 1.0, 16, 4)
    p2 = fe.pointpartitioning(f2.xyz, 4)
    assert sorted(np.unique(p2)) == [1, 2, 3, 4]
    for lab in (1, 2, 3, 4):
        xs = f2.xyz[p2 == lab, 0]
        assert xs.max() - xs.min() <= 1.0 + 1e-12          # every part is one quarter of the strip
    with pytest.raises(ValueError):
        fe.pointpartitioning(np.zeros((5, 1)), 2)


def test_golden_fixtures_against_oracle(orc, fe):
    """tests/golden/*.npz were produced by tests/golden/make_golden.py (oracle run in the build container); the oracle
    must keep reproducing them bit for bit in the pattern and to 1e-15 in the values."""
    import glob
    from helpers import oracle_csc
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
    assert files, "golden fixtures missing"
    from golden.make_golden import CASES, build_case
    for f in files:
        g = np.load(f)
        name = os.path.basename(f)[:-4]
        fens, fes, u, rule, coef, form, et, kw = build_case(fe, CASES[name])
        (cp, rv, nz), _ = oracle_csc(orc, form, et, fes, fens, u, rule, coef, **kw)
        np.testing.assert_array_equal(cp, g["colptr"])
        np.testing.assert_array_equal(rv, g["rowval"])
        assert np.abs(nz - g["nzval"]).max() <= 1e-15 * np.abs(g["nzval"]).max()


def _julia_ccalls(src):
    """Every `ccall((:name, LIB), Ret, (T1, T2, ...), args...)` of the Julia shim: (name, return type, [argument types])."""
    import re
    out = []
    for m in re.finditer(r"ccall\(\(:(\w+),\s*LIB\),\s*(\w+),\s*\(", src):
        i, depth = m.end(), 1
        while depth:  # balanced scan of the argument-type tuple
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        tup = src[m.end():i - 1]
        parts, depth, cur = [], 0, ""
        for ch in tup:
            if ch == "," and depth == 0:
                parts.append(cur.strip())
                cur = ""
            else:
                depth += {"{": 1, "}": -1}.get(ch, 0)
                cur += ch
        if cur.strip():
            parts.append(cur.strip())
        out.append((m.group(1), m.group(2), parts))
    return out


def test_julia_shim_ccalls_match_the_abi(fe):
    """The Julia shim cannot be executed here (no julia on the image), so its FFI layer is checked statically: every ccall
    names a symbol of include/fegpu.h and passes the right number and kind of arguments (pointer / Int32 / Int64 / Float64),
    judged against the ctypes table that the header test pins and the GPU tests exercise."""
    import ctypes as C
    from finetools_jl_b200 import _lib
    src = open(os.path.join(ROOT, "finetools.jl_b200", "julia", "FinEtoolsGPU.jl")).read()
    calls = _julia_ccalls(src)
    assert len(calls) >= 35
    kind_c = {C.c_int32: "i32", C.c_int64: "i64", C.c_double: "f64"}
    kind_jl = {"Int32": "i32", "Cint": "i32", "Int64": "i64", "Float64": "f64", "Cdouble": "f64"}
    ret_jl = {"Int32": C.c_int32, "Int64": C.c_int64, "Cstring": C.c_char_p, "Ptr{UInt8}": C.c_char_p}
    seen = set()
    for name, ret, args in calls:
        assert name in _lib.SIGNATURES, "Julia shim calls unknown symbol %s" % name
        restype, argtypes = _lib.SIGNATURES[name]
        assert ret_jl.get(ret) is restype, "%s: return type %s" % (name, ret)
        assert len(args) == len(argtypes), "%s: %d arguments in the shim, %d in the ABI" % (name, len(args), len(argtypes))
        for k, (a, t) in enumerate(zip(args, argtypes)):
            want = kind_c.get(t, "ptr")
            got = "ptr" if a.startswith(("Ptr{", "Ref{")) else kind_jl.get(a)
            assert got == want, "%s: argument %d is %s in the shim, %s in the ABI" % (name, k + 1, a, want)
        seen.add(name)
    # the path's own entry points are all bound by the shim
    for must in ("fegpu_create", "fegpu_mesh_upload", "fegpu_dofmap_upload", "fegpu_rule_set", "fegpu_bilform_diffusion",
                 "fegpu_bilform_lin_elastic", "fegpu_bilform_dot", "fegpu_startassembly", "fegpu_assemble", "fegpu_makematrix",
                 "fegpu_makematrix_sizes", "fegpu_makematrix_copy", "fegpu_set_async", "fegpu_cache_release",
                 # multi-GPU split, the nomatrixresult flow, values-only re-assembly, device-resident results, windowed geometry
                 "fegpu_partition_set", "fegpu_coo_copy", "fegpu_makematrix_copy_values", "fegpu_makematrix_device",
                 "fegpu_geom_update_window", "fegpu_pattern_was_cached", "fegpu_host_alloc", "fegpu_host_free", "fegpu_makematrix_view"):
        assert must in seen, must


def test_julia_shim_eligibility_and_caching_rules():
    """What the shim must refuse / must share, checked on its source (it cannot run here):
    - a DataCache whose `_fillcache!` is not the constant constructor's closure is an error (DataCacheModule.jl:65-89), as is a
      user-supplied other-dimension function (IntegDomainModule.jl:73-104): never the initial buffer read silently;
    - ONE context per device and a per-device mesh cache, so an assembler per call (FEMMBaseModule.jl:1374) re-uses the device
      mesh and the cached pattern; assemblers do not create or destroy contexts;
    - the quadrature tables are compared on every call (two FEMMs on one FESet with different rules);
    - SysmatAssemblerFFBlock{<:SysmatAssemblerSparseGPU} has its own bilform methods (no CPU element loop)."""
    src = open(os.path.join(ROOT, "finetools.jl_b200", "julia", "FinEtoolsGPU.jl")).read()
    assert re.search(r'occursin\("_fillcache_constant!", string\(nameof\(typeof\(cf\._fillcache!\)\)\)\)\s*\|\|\s*\n?\s*error\(', src)
    assert "cf._cache" not in re.sub(r"function _constant_cache.*?\nend\n", "", src, flags=re.S), "forms must read the cache through _constant_cache"
    assert re.search(r"f === otherdimensionunity && return 1\.0", src) and 'occursin("otherdimensionfu"' in src
    assert re.search(r"function _otherdim.*?error\(\"only the unit or a constant other-dimension", src, flags=re.S)
    # _eligible runs both checks before anything is uploaded, and every form calls _eligible first
    elig = re.search(r"function _eligible.*?\nend\n", src, flags=re.S).group(0)
    assert "_otherdim(self)" in elig and "_constant_cache(cf)" in elig
    for form in ("bilform_diffusion", "bilform_lin_elastic", "bilform_dot", "bilform_convection", "bilform_div_grad", "bilform_masslike", "linform_dot"):
        body = re.search(r"function %s\(self::FEMMBase, assembler::(SysmatAssemblerSparseGPU|SysvecAssemblerGPU).*?\nend\n" % form, src, flags=re.S).group(0)
        assert body.index("_eligible(") < body.index("_device("), form
    # one context per device; assemblers neither create nor destroy it
    assert src.count(":fegpu_create") == 1 and "const _DEVICES = Dict{Int,DeviceState}()" in src
    assert ":fegpu_destroy" not in src
    ctor = re.search(r"function SysmatAssemblerSparseGPU\(z::Float64.*?\nend\n", src, flags=re.S).group(0)
    assert "_device_state(device)" in ctor and ":fegpu_create" not in ctor
    # rule tables compared on every call
    dev = re.search(r"function _device\(self::FEMMBase.*?\nend\n", src, flags=re.S).group(0)
    assert "t.rule != (npts, N, dN, ww)" in dev and ":fegpu_rule_set" in dev
    assert "t.owner[1] != own" in dev and ":fegpu_partition_set" in dev
    # FFBlock over a GPU assembler: dedicated methods for every bilinear form on the path
    assert "const FFBlockGPU = SysmatAssemblerFFBlock{<:SysmatAssemblerSparseGPU}" in src
    for form in ("bilform_diffusion", "bilform_lin_elastic", "bilform_dot"):
        assert re.search(r"%s\(self::FEMMBase, assembler::FFBlockGPU" % form, src), form


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) prints exactly one JSON line with the keys of
    the contract; a non-zero rank under torchrun prints nothing and exits 0.  Tiny sample so the test takes a second."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3", "--ref-edge", "6",
           "--ref-threads", "2", "--gpus", "2"]
    env = dict(os.environ, RANK="0", WORLD_SIZE="2")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "elements/s" and d["higher_is_better"] is True and d["n_gpus"] == 2
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 3 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert other.returncode == 0 and other.stdout.strip() == ""


def test_bench_gpu_arm_refuses_to_run_without_a_device():
    """No CPU fallback anywhere on the measured path: without a CUDA device the GPU arm of bench.py exits with an error instead
    of timing something else.  (Skipped on a GPU box, where the arm would really run.)"""
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)
    assert not [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]


def test_gather_loads_are_spread_over_scoreboards():
    """Static guard for the finding of profiles/r02_scoreboards.txt: in the default instantiations of k_gather_tile (H8, scalar and
    3-dof, value planes, MODE 1) ptxas must not put (nearly) all global loads on one scoreboard -- if it does, the first add of an
    element waits for the loads of the next one and the software pipeline overlaps nothing.  Read from the SASS control codes of the
    in-tree object (cuobjdump; skipped when the object or the tool is missing)."""
    import shutil
    import subprocess
    from collections import Counter
    obj = os.path.join(ROOT, "finetools.jl_b200", "csrc", "fegpu_tile.o")
    if not os.path.exists(obj) or not shutil.which("cuobjdump"):
        pytest.skip("needs the built object and cuobjdump")
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur:
            funcs[cur].append(line)
    checked = 0
    for name, lines in funcs.items():
        if not re.search(r"k_gather_tileILi8ELi8ELi[13]ELb1ELb1ELi1E", name):
            continue
        wb = Counter()
        i = 0
        while i + 1 < len(lines):
            m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", lines[i])
            m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", lines[i + 1])
            if m and m2:
                if re.match(r"(@!?U?P[0-9T]+ )?LDG", m.group(1).strip()):
                    wb[((int(m2.group(1), 16) >> 41) >> 5) & 7] += 1   # write-barrier index of the control code
                i += 2
            else:
                i += 1
        total = sum(wb.values())
        assert total >= 60, (name, wb)
        # the pathological build had 207 of 214 loads (97 %) on one scoreboard; the measured-best one has at most 62 %
        assert max(wb.values()) <= 0.75 * total and len(wb) >= 3, "loads crowd one scoreboard: %r" % (wb,)
        checked += 1
    assert checked == 2
