"""The C-ABI library loads, exports every symbol include/gs_b200.h declares, and has no CPU fallback."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared(header, prefix):
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    return sorted(set(re.findall(rf"\b({prefix}_\w+)\s*\(", text)))


def test_c_abi_exports_every_declared_symbol(gs):
    names = declared(ROOT / "include" / "gs_b200.h", "gsb")
    assert names, "no declarations parsed"
    assert sorted(gs.EXPORTED_SYMBOLS) == names
    for n in names:
        assert hasattr(gs.lib, n), n


def test_host_bridge_exports_every_declared_symbol(gs):
    names = declared(ROOT / "3dgs.cpp_b200" / "host" / "gs_b200_host.h", "gsh")
    assert sorted(gs.HOST_EXPORTED_SYMBOLS) == names
    for n in names:
        assert hasattr(gs.host, n), n


def test_abi_version_and_struct_sizes(gs):
    assert gs.lib.gsb_abi_version() == 3
    assert C.sizeof(gs.Uniforms) == 160  # Renderer::UniformBuffer, std140
    assert gs.ATTR_DTYPE.itemsize == 64  # VertexAttribute
    assert C.sizeof(gs.Stats) == 6 * 8 + 2 * 4 + 7 * 4 + 3 * 4 + 8 * 4 + 2 * 4 + 8 + 8 + 8 + 2 * 4


def test_no_cpu_fallback_without_device(gs):
    """On a box without a GPU the product must fail loudly, not fall back to a CPU path."""
    if gs.lib.gsb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = gs.lib.gsb_create(0, C.byref(h))
    assert rc == gs.ERR_NO_DEVICE and not h.value
    assert b"no CPU path" in gs.lib.gsb_last_error(None)
    with pytest.raises(gs.GsbError):
        gs.Context(0)


def test_product_never_references_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may include, link, load or import it
    (comments that cite it for documentation are fine)."""
    pkg = ROOT / "3dgs.cpp_b200"
    bad = re.compile(r'#\s*include\s*[<"][^>"]*oracle|liboracle|-loracle|\bimport\s+oracle\b|\bfrom\s+oracle\b|dlopen\([^)]*oracle|\.\./oracle')
    for p in pkg.rglob("*"):
        if p.suffix in {".cu", ".cuh", ".cpp", ".h", ".py", ".txt", ".cmake"} or p.name in {"Makefile", "CMakeLists.txt"}:
            m = bad.search(p.read_text(errors="ignore"))
            assert m is None, (p, m.group(0))
