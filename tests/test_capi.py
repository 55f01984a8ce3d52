"""CPU tests of the drop-in boundary: the CUDA libraries load, export every symbol include/ode_b200.h
declares, and fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import numpy as np
import pytest
from parity_util import B, ROOT, gpu_lib
from ode_b200 import scenes


def _declared():
    txt = open(os.path.join(ROOT, "include", "ode_b200.h")).read()
    return sorted(set(re.findall(r"\b(odeb_[a-z_0-9]+)\s*\(", txt)))


@pytest.mark.parametrize("prec", ("single", "double"))
def test_exports_every_declared_symbol(prec):
    lib = C.CDLL(os.path.join(ROOT, "ode_b200", "libode_b200_%s.so" % prec))
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "%s missing from libode_b200_%s.so" % (n, prec)


def test_struct_layout_matches_header():
    """ctypes mirrors must have the sizes the C compiler gives the structs of the header."""
    import subprocess, tempfile
    src = '#include "ode_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(OdebWorldParams), sizeof(OdebBodyDesc), sizeof(OdebGeomDesc), sizeof(OdebJointDesc));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    assert [int(x) for x in out] == [C.sizeof(B.OdebWorldParams), C.sizeof(B.OdebBodyDesc), C.sizeof(B.OdebGeomDesc), C.sizeof(B.OdebJointDesc)]


def _have_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_no_cpu_fallback():
    """Without a GPU the product must refuse to run instead of silently computing on the host."""
    if _have_cuda():
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError) as e:
        B.Batch(gpu_lib("single"), scenes.box_stack(nworlds=1, nboxes=2))
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_package_does_not_touch_oracle():
    """Nothing under ode_b200/ may import, link or open anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ode_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                for line in txt.splitlines():
                    s = line.strip()
                    if s.startswith(("#include", "import ", "from ")) or "CDLL" in s or "dlopen" in s:
                        assert "oracle" not in s and "orc_" not in s, (f, s)


def _atan2_inputs(n, seed):
    rng = np.random.default_rng(seed)
    parts = []
    q = n // 4
    parts.append((rng.integers(0, 2**32, q, dtype=np.uint64).astype(np.uint32).view(np.float32), rng.integers(0, 2**32, q, dtype=np.uint64).astype(np.uint32).view(np.float32)))
    parts.append(((rng.random(q) * 4 - 2).astype(np.float32), (rng.random(q) * 4 - 2).astype(np.float32)))
    parts.append((((rng.integers(0, 1024, q) - 512) * 0.001953125).astype(np.float32), ((rng.integers(0, 1024, q) - 512) * 0.001953125).astype(np.float32)))   # symmetric angles
    parts.append((((rng.random(q) - 0.5) * 1e3).astype(np.float32), ((rng.random(q) - 0.5) * 1e-3).astype(np.float32)))
    y = np.ascontiguousarray(np.concatenate([p[0] for p in parts]))
    x = np.ascontiguousarray(np.concatenate([p[1] for p in parts]))
    return y, x


def _host_atan2f(y, x):
    import ctypes.util
    libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.atan2f.restype = C.c_float
    libm.atan2f.argtypes = [C.c_float, C.c_float]
    return np.array([libm.atan2f(float(a), float(b)) for a, b in zip(y, x)], dtype=np.float32)


def _lib_atan2f(y, x, on_device):
    lib = C.CDLL(os.path.join(ROOT, "ode_b200", "libode_b200_single.so"))
    out = np.empty_like(y)
    f = lib.odeb_test_atan2f
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    assert f(y.ctypes.data, x.ctypes.data, out.ctypes.data, len(y), on_device) == 1
    return out


def test_atan2f_replica_matches_host_libm():
    """The library computes single-precision atan2 (cullPoints, hinge / universal / motor angles) with the fdlibm algorithm restated in
    odeb_math.cuh; the reference calls the host libm.  Same bits on the host side of the shared source (NaN payloads aside)."""
    y, x = _atan2_inputs(200000, 11)
    want, got = _host_atan2f(y, x), _lib_atan2f(y, x, 0)
    ok = (want.view(np.uint32) == got.view(np.uint32)) | (np.isnan(want) & np.isnan(got))
    assert ok.all(), (y[~ok][:5], x[~ok][:5], want[~ok][:5], got[~ok][:5])


@pytest.mark.gpu
def test_atan2f_on_the_device_matches_host_libm():
    """... and the same bits from the sm_100a code (no FMA contraction, IEEE division)."""
    y, x = _atan2_inputs(400000, 12)
    want, got = _host_atan2f(y, x), _lib_atan2f(y, x, 1)
    ok = (want.view(np.uint32) == got.view(np.uint32)) | (np.isnan(want) & np.isnan(got))
    assert ok.all(), (y[~ok][:5], x[~ok][:5], want[~ok][:5], got[~ok][:5])
