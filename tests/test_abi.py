"""CPU tests of the drop-in boundary: libsmcb200.so builds, loads, exports every symbol that
include/smcb200.h declares, fails loudly without a GPU, and its host-side helpers (Philox normals,
simulate) are bit-identical to the oracle's."""
import os
import re

import numpy as np
import pytest

import sequential_monte_carlo_b200 as smc
from sequential_monte_carlo_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "smcb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smcb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/smcb200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"
    assert lib.smcb_version() == 100
    assert lib.smcb_state_dim(smc.KIND_UCSV) == 3 and lib.smcb_state_dim(smc.KIND_LG1D) == 1 and lib.smcb_state_dim(9) == -1


def build_c_client():
    """tests/abi_client.c compiled as plain C against include/smcb200.h (what an FFI binding sees); returns the binary"""
    import subprocess
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe, src = os.path.join(out_dir, "abi_client"), os.path.join(ROOT, "tests", "abi_client.c")
    deps = [src, os.path.join(ROOT, "include", "smcb200.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["gcc", "-std=gnu11", "-O1", "-Wall", "-Wextra", "-Werror", "-ffp-contract=off", src, "-o", exe, "-ldl", "-lm"])
    return exe


def run_c_client(*args):
    import subprocess
    _lib.load()
    out = subprocess.run([build_c_client(), _lib.library_path(), *args], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    return dict(line.split(None, 1) for line in out.stdout.strip().splitlines())


def test_header_is_plain_c_and_a_c_client_resolves_every_symbol(oracle):
    """the drop-in boundary from C: the header compiles as C11 with -Wall -Wextra -Werror, dlopen + dlsym find every
    declared entry point, and the host-side helpers give the oracle's numbers"""
    names = _declared()
    assert run_c_client("symbols", *names)["resolved"] == str(len(names))
    got = run_c_client("host")
    assert got["version"] == "100" and got["dims"] == "1 1 3 -1"
    _, yo = oracle.simulate(0, [0.5, 1.0, 0.9, 0.8, 0.0, 1.0], 40, 1998)
    for t in range(3):
        assert float(got[f"y{t}"]) == yo[t]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(smc.SMCBError) as e:
        smc.Context(0, 1)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sequential_monte_carlo_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt and '#include "../../oracle' not in txt, f


def test_host_rng_matches_oracle(oracle):
    for purpose, comp in ((4, 0), (6, 2), (8, 1)):
        a = _lib.rng_normals(1998, 5, 7, 11, purpose, comp, 1001)
        b = oracle.normals(1998, 5, 7, 11, purpose, comp, 1001)
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(_lib.rng_uniforms64(3, 1, 2, 9, 7, 77), oracle.uniforms64(3, 1, 2, 9, 7, 77))
    np.testing.assert_array_equal(_lib.rng_uniforms01(3, 1, 2, 9, 7, 77), oracle.uniforms01(3, 1, 2, 9, 7, 77))


@pytest.mark.parametrize("kind,params", [(0, [0.5, 1.0, 0.9, 0.8, 0.0, 1.0]), (1, [-1.0, 0.9, 0.3]), (2, [0.2, 0.2, 3.0, 1.0, 1.0])])
def test_simulate_matches_oracle(oracle, kind, params):
    x, y = _lib.simulate(kind, params, 500, 1998)
    xo, yo = oracle.simulate(kind, params, 500, 1998)
    np.testing.assert_array_equal(x, xo)
    np.testing.assert_array_equal(y, yo)
    assert x.shape == (3 if kind == 2 else 1, 500)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times beside the CUDA arm) prints ONE JSON line with the
    contract's keys, on the CUDA arm's metric / unit / config, without touching a GPU"""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "particle-updates/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("particle-updates/sec") and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["config"]["N"] == 1 << 24 and d["config"]["T"] == 1000 and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"] > 1e5
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_julia_shim_ccalls_match_the_header():
    """julia/SequentialMonteCarloB200.jl cannot be executed here (no Julia in the image), so its bindings are checked statically:
    every `ccall((:smcb_…, LIB), Ret, (ArgTypes…), …)` names a function the header declares, passes as many argument types as the
    prototype has parameters, and each type is of the class of the C parameter (Float64 ↔ double, Cint ↔ int, Int64 ↔ int64_t,
    UInt32 / UInt64 ↔ uint32_t / uint64_t, Ptr / Ref / arrays ↔ pointers)."""
    import re
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "smcb200.h")).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|const char\s*\*|int64_t|void|double)\s+(smcb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
    assert len(protos) >= 70

    def c_class(p):
        if "*" in p or "[" in p:
            return "ptr"
        for t, k in (("double", "f64"), ("uint64_t", "u64"), ("uint32_t", "u32"), ("int64_t", "i64"), ("int", "i32")):
            if re.search(r"\b" + t + r"\b", p):
                return k
        raise AssertionError("unclassified C parameter: " + p)

    def jl_class(t):
        t = t.strip()
        if t.startswith(("Ptr{", "Ref{")) or t in ("Cstring",):
            return "ptr"
        return {"Float64": "f64", "Cdouble": "f64", "UInt64": "u64", "UInt32": "u32", "Int64": "i64", "Clonglong": "i64", "Cint": "i32", "Int32": "i32"}[t]

    jl = open(os.path.join(ROOT, "julia", "SequentialMonteCarloB200.jl")).read()
    seen, pos = set(), 0
    while True:
        k = jl.find("ccall((:", pos)
        if k < 0:
            break
        pos = k + 6
        m = re.match(r"ccall\(\(:(smcb_[a-z0-9_]+),\s*LIB\),\s*([A-Za-z0-9{}\. ]+?),\s*\(", jl[k:])
        assert m, jl[k:k + 80]
        name, start, depth = m.group(1), k + m.end() - 1, 0
        for j in range(start, len(jl)):
            depth += (jl[j] == "(") - (jl[j] == ")")
            if depth == 0:
                break
        parts, d, cur = [], 0, ""
        for ch in jl[start + 1:j]:
            d += (ch in "({[") - (ch in ")}]")
            if ch == "," and d == 0:
                parts.append(cur.strip())
                cur = ""
            else:
                cur += ch
        if cur.strip():
            parts.append(cur.strip())
        assert name in protos, name + " is not declared in include/smcb200.h"
        assert len(parts) == len(protos[name]), (name, protos[name], parts)
        for cp, jt in zip(protos[name], parts):
            assert c_class(cp) == jl_class(jt), (name, cp, jt)
        seen.add(name)
    assert len(seen) >= 35            # the shim binds the filter, batch, Kalman, sampler, communicator and host-only entry points
