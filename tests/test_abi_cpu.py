"""CPU-side checks of the drop-in boundary: librt_b200.so loads, exports every symbol include/rt_b200.h declares,
and fails LOUDLY (no CPU fallback) when there is no CUDA device.  No compute entry point is called without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import raytracing_jl_b200 as rt
from raytracing_jl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(rt_[a-z_0-9]+)\s*\(", src))
    names -= {"rt_batch_cb"}  # a function-pointer typedef, not an export
    return names


@pytest.fixture(scope="module")
def so():
    return rt.build()


def test_header_and_binding_agree(so):
    hdr = declared_symbols()
    assert hdr == set(_lib.SYMBOLS), (hdr ^ set(_lib.SYMBOLS))
    assert len(hdr) >= 24


def test_library_exports_every_declared_symbol(so):
    L = C.CDLL(so)
    for name in declared_symbols():
        assert hasattr(L, name), name
    L.rt_version.restype = C.c_char_p
    assert b"sm_100a" in L.rt_version()


def test_library_does_not_depend_on_the_oracle(so):
    import subprocess

    deps = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    assert "oracle" not in deps
    pkg = os.path.join(ROOT, "raytracing.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "rt_oracle" not in txt, f


def test_status_codes_match_header():
    src = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    assert re.search(r"RT_ERR_TRACK\s*=\s*-8", src) and re.search(r"RT_ERR_NO_EXIT\s*=\s*-4", src)
    for name, val in (("RT_SEG_LITERAL", 1), ("RT_SEG_NO_VOLUMES", 2), ("RT_SEG_COUNT_ONLY", 4), ("RT_SEG_NO_CHUNKS", 8), ("RT_SEG_SEQUENTIAL", 16)):
        assert re.search(rf"{name}\s*=\s*{val}\b", src) and getattr(_lib, name) == val
    assert [int(rt.Vacuum), int(rt.Reflective), int(rt.Periodic)] == [0, 1, 2]  # src/boundary.jl:12-16
    assert [int(rt.Forward), int(rt.Backward)] == [0, 1]  # src/track.jl:11-14


def _has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_cuda(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(pincell_model):
    h = C.c_void_p()
    assert _lib.lib().rt_create(C.byref(h), 0) == -1  # RT_ERR_CUDA
    with pytest.raises(rt.RTError):
        rt.TrackGenerator(pincell_model, 8, 0.02)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(rt.RTError):
        _lib.lib()


# ---- host logic that needs no device ---------------------------------------------------------------
def test_track_layout_counts(pincell_mesh):  # test/runtests.jl:14-19
    from tests.golden import runtests_goldens as G

    lay = rt.TrackLayout(pincell_mesh, G.MAIN["n_azim"], G.MAIN["delta"])
    assert lay.n_total_tracks == G.MAIN["n_total_tracks"]
    assert lay.n_tracks_x.tolist() == G.MAIN["n_tracks_x"] and lay.n_tracks_y.tolist() == G.MAIN["n_tracks_y"]
    from raytracing_jl_b200.api import _angle_tables

    _angle_tables(lay)
    aq = lay.azimuthal_quadrature
    assert np.allclose(aq.phis, G.MAIN["phis"], rtol=1e-12) and np.allclose(aq.deltas, G.MAIN["delta_eff"], rtol=1e-12)
    ph, w = aq.phis, aq.weights  # src/azimuthal_quad.jl:39-51 with N4 = 2
    assert w[0] == w[3] == (ph[1] - ph[0]) / (4 * np.pi) and w[1] == w[2] == (np.pi - ph[1] - ph[0]) / (4 * np.pi)


def test_domain_errors_are_host_side(pincell_mesh):  # src/azimuthal_quad.jl:22-25
    for args in [(0, 0.1), (-4, 0.1), (6, 0.1), (8, 0.0), (8, -1.0)]:
        with pytest.raises(rt.DomainError):
            rt.TrackLayout(pincell_mesh, *args)


def test_mesh_tables_are_gridap_layout(pincell_model, pincell_mesh):
    ptrs, data = pincell_mesh.node_cells
    assert ptrs[0] == 1 and ptrs[-1] == data.size + 1 and data.min() == 1 and data.max() == pincell_model.num_cells
    for node in (1, 17, pincell_model.num_nodes):  # cells around a node in ascending cell id (Gridap get_faces(topo,0,2))
        cells = data[ptrs[node - 1] - 1:ptrs[node] - 1]
        assert np.all(np.diff(cells) > 0)
        for c in cells:
            assert node in pincell_model.cell_data[3 * (c - 1):3 * c]
    assert pincell_mesh.bb_min.tolist() == [0.0, 0.0] and pincell_mesh.bb_max.tolist() == [1.6, 1.6]


def test_reference_arm_contract():
    """bench.py --impl reference (the CPU arm the driver runs next to ours): ONE JSON line on rank 0 with the contract's keys,
    nothing from the other ranks."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "pincell"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env={**os.environ, "RANK": "0", "WORLD_SIZE": "1"})
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "segments/sec for segmentize!" and d["unit"] == "segments/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    out = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, timeout=300, env={**os.environ, "RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_every_option_of_rt_set_option_is_documented_in_the_header():
    """include/rt_b200.h lists the tuning knobs; the list must not drift from what rt_set_option accepts"""
    import re

    src = open(os.path.join(ROOT, "raytracing.jl_b200", "csrc", "rt_b200.cu")).read()
    body = src[src.index("int rt_set_option("):]
    body = body[:body.index("\n}\n")]
    names = set(re.findall(r'n == "([a-z_0-9]+)"', body))
    assert len(names) > 10
    hdr = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    missing = sorted(n for n in names if f'"{n}"' not in hdr)
    assert not missing, missing


def _c_prototypes():
    """{name: [kind of every parameter]} from include/rt_b200.h; kinds: ptr, i32, i64, f64"""
    import re

    hdr = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|void|const char \*)\s*(rt_[a-z_0-9]+)\s*\(([^;{}]*?)\)\s*;", hdr, flags=re.S):
        name, args = m.group(1), " ".join(m.group(2).split())
        kinds = []
        for a in ([] if args in ("", "void") else args.split(",")):
            a = a.strip()
            if "*" in a or "[" in a or a.startswith("rt_batch_cb"):
                kinds.append("ptr")
            elif re.match(r"(const )?(int32_t|uint32_t|int|unsigned)\b", a):
                kinds.append("i32")
            elif re.match(r"(const )?(int64_t|uint64_t|long long|size_t)\b", a):
                kinds.append("i64")
            elif re.match(r"(const )?double\b", a):
                kinds.append("f64")
            else:
                raise AssertionError(f"unparsed parameter {a!r} of {name}")
        protos[name] = kinds
    return protos


def test_julia_glue_calls_match_the_header():
    """julia/RayTracingB200.jl cannot run here (no Julia in the image): at least every `ccall` in it must name an exported
    function and pass the same number of arguments, of the same machine kind (pointer / 32-bit / 64-bit integer / double) and in
    the same order, as the prototype in include/rt_b200.h."""
    import re

    protos = _c_prototypes()
    assert len(protos) >= 30 and protos["rt_create"] == ["ptr", "i32"]
    src = open(os.path.join(ROOT, "julia", "RayTracingB200.jl")).read()
    kind = {"Int32": "i32", "Cint": "i32", "UInt32": "i32", "Int64": "i64", "UInt64": "i64", "Csize_t": "i64", "Float64": "f64",
            "Cdouble": "f64", "Cstring": "ptr"}
    calls = list(re.finditer(r"ccall\(\(:(rt_[a-z_0-9]+),\s*LIBRT_B200\),\s*(\w+),\s*\(", src))
    assert len(calls) >= 15
    for m in calls:
        name, ret = m.group(1), m.group(2)
        assert name in protos, name
        i, depth, start = m.end(), 1, m.end()
        while depth:  # the argument-type tuple ends at the matching parenthesis
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        types, cur, d = [], "", 0
        for ch in src[start:i - 1]:
            if ch == "," and d == 0:
                types.append(cur.strip())
                cur = ""
            else:
                d += {"{": 1, "}": -1}.get(ch, 0)
                cur += ch
        if cur.strip():
            types.append(cur.strip())
        got = ["ptr" if t.startswith("Ptr{") else kind[t] for t in types]
        assert got == protos[name], (name, got, protos[name])
        assert ret in ("Cint", "Cvoid", "Cstring")


def test_ctypes_table_matches_the_header():
    """raytracing.jl_b200/_lib.py SYMBOLS against include/rt_b200.h: same functions, same number, order and machine kind of
    arguments (a wrong width in a ctypes prototype corrupts a call silently)"""
    import ctypes as C

    from raytracing_jl_b200 import _lib

    protos = _c_prototypes()
    assert set(_lib.SYMBOLS) == set(protos), set(_lib.SYMBOLS) ^ set(protos)

    def kind(t):
        if t in (C.c_int, C.c_int32, C.c_uint32):
            return "i32"
        if t in (C.c_int64, C.c_uint64, C.c_size_t, C.c_longlong):
            return "i64"
        if t is C.c_double:
            return "f64"
        return "ptr"  # c_void_p, c_char_p, POINTER(...), ndpointer, CFUNCTYPE

    for name, (_, args) in _lib.SYMBOLS.items():
        assert [kind(t) for t in args] == protos[name], (name, [kind(t) for t in args], protos[name])
