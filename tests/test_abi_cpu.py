"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol include/pimc_b200.h
declares, fails loudly (no CPU fallback) on compute calls, and the RNG spec header reproduces the Random123 KATs."""
import ctypes as C
import os
import re
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pimc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pimc_[A-Za-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from pimc_jl_b200 import _lib
    return _lib


def test_every_declared_symbol_is_exported_and_bound(lib):
    L = lib.load()
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/pimc_b200.h but not exported by libpimc_b200.so"
        assert n in lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(lib.SIGNATURES) == names
    assert L.pimc_version() >= 100 and L.pimc_launch_count() == 0


def test_constants_mirror_the_header(lib):
    """the ctypes layer's enum / option / compat constants are the ones include/pimc_b200.h defines"""
    src = open(os.path.join(ROOT, "include", "pimc_b200.h")).read()
    vals = {k: int(v, 0) for k, v in re.findall(r"#define\s+(PIMC_[A-Z0-9_]+)\s+(-?(?:0x[0-9a-fA-F]+|\d+))\b", src)}
    for body in re.findall(r"enum\s*\{([^}]*)\}", src):
        nxt = 0
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, v = [t.strip() for t in item.split("=")]
                nxt = int(v, 0)
            else:
                name = item
            vals[name] = nxt
            nxt += 1
    checked = 0
    for name in dir(lib):
        if name.startswith(("POT_", "DV_", "UPD_", "SCHED_", "OPT_", "COMPAT_")) and name != "COMPAT_ALL":
            assert ("PIMC_" + name) in vals, f"{name} has no counterpart in the header"
            assert vals["PIMC_" + name] == getattr(lib, name), name
            checked += 1
    assert checked >= 16


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pimc_jl_b200 as pj
    with pytest.raises(pj.PimcError) as ei:
        pj.Engine(chains=1)
    assert "no CUDA device" in str(ei.value)
    x = np.zeros(4)
    rc = lib.load().pimc_teleport(4, x.ctypes.data_as(lib.f64p), 1.0, x.ctypes.data_as(lib.f64p))
    assert rc == -2


def test_product_never_touches_oracle():
    """nothing under pimc_jl_b200/ may import, include, link or load anything from oracle/"""
    pat = re.compile(r"(import\s+oracle|from\s+oracle|oracle_binding|libpimc_oracle|pimc_oracle\.|oracle/|ora_[a-z_]+\s*\()")
    for dp, _, fs in os.walk(os.path.join(ROOT, "pimc_jl_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dp, f)).read()
                txt = txt.replace("nothing in this package routes through oracle/ or any other CPU path", "")
                assert not pat.search(txt), f


def test_sass_is_sm100a(lib):
    out = subprocess.run(["cuobjdump", "-lelf", lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_philox_kat_and_gaussian_moments(oracle):
    """Random123 known-answer vectors for Philox4x32-10 through the C header, and moments of the Box-Muller map."""
    src = r'''
#include <stdio.h>
#include "pimc_rng.h"
int main(){ pimc_u4 o;
 o = pimc_philox4x32_10(0,0,0,0,0,0); printf("%08x %08x %08x %08x\n",o.w[0],o.w[1],o.w[2],o.w[3]);
 o = pimc_philox4x32_10(~0u,~0u,~0u,~0u,~0u,~0u); printf("%08x %08x %08x %08x\n",o.w[0],o.w[1],o.w[2],o.w[3]);
 o = pimc_philox4x32_10(0x243f6a88,0x85a308d3,0x13198a2e,0x03707344,0xa4093822,0x299f31d0); printf("%08x %08x %08x %08x\n",o.w[0],o.w[1],o.w[2],o.w[3]);
 return 0; }'''
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    cfile, exe = os.path.join(d, "_kat.c"), os.path.join(d, "_kat")
    open(cfile, "w").write(src)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe, "-lm"])
    out = subprocess.run([exe], capture_output=True, text=True).stdout.split("\n")
    assert out[0] == "6627e8d5 e169c58d bc57ac4c 9b00dbd8"
    assert out[1] == "408f276d 41c83b0e a20bc7c6 6d5451fd"
    assert out[2] == "d16cfe09 94fdcceb 5001e420 24126ea1"
    ob = oracle
    a, b = C.c_double(), C.c_double()
    g = np.zeros((20000, 2))
    for i in range(20000):
        ob.lib().ora_gauss_pair(99, 3, 7, i & 0xFFFF, 2, i >> 16, i % 9000, C.byref(a), C.byref(b))
        g[i] = (a.value, b.value)
    x = g.ravel()
    assert abs(x.mean()) < 0.03 and abs(x.var() - 1) < 0.03 and abs((x ** 4).mean() - 3) < 0.15
    assert abs(np.corrcoef(g[:, 0], g[:, 1])[0, 1]) < 0.03


def test_header_is_plain_c_and_the_c_driver_links(tmp_path):
    """include/pimc_b200.h must be consumable by a C compiler (the drop-in boundary is a C ABI: `ccall` / cgo / ctypes bind plain symbols),
    and the plain-C driver of examples/density_SRL_lattice.jl (tests/c/) must compile and link against libpimc_b200.so with -Wall -Werror
    -- no GPU needed: nothing is executed here."""
    import subprocess
    inc, so_dir = os.path.join(ROOT, "include"), os.path.join(ROOT, "pimc_jl_b200")
    probe = tmp_path / "probe.c"
    probe.write_text('#include "pimc_b200.h"\n#include "pimc_rng.h"\nint main(void) { pimc_config c; (void)c; return sizeof(pimc_measurements) > 0 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c11", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I" + inc, "-c", "-o", str(tmp_path / "probe.o"), str(probe)])
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I" + inc, "-o", str(tmp_path / "example_srl"),
                           os.path.join(ROOT, "tests", "c", "example_density_srl_lattice.c"), "-L" + so_dir, "-lpimc_b200", "-lm", "-Wl,-rpath," + so_dir])
