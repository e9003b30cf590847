"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol that
include/ecloop_b200.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "ecloop_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ecl_[a-z0-9_]+)\s*\(", text)))


def test_header_is_plain_c():
    # the boundary must compile as C with no CUDA / torch types
    src = '#include "ecloop_b200.h"\nint main(void){ ecl_hit h; return (int)sizeof(h) - 32; }\n'
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-x", "c", "-", "-o", "/dev/null"],
                       input=src.encode(), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()


def test_library_exports_every_declared_symbol():
    import ecloop_b200 as E

    lib = E.load_library()
    syms = declared_symbols()
    assert set(syms) == set(E.ABI_SYMBOLS)
    for s in syms:
        assert getattr(lib, s) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", str(E.library_path())], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (ecl_[a-z0-9_]+)", out))
    assert set(syms) <= exported
    assert lib.ecl_abi_version() == 2


def test_is_sm100a_native():
    import ecloop_b200 as E

    out = subprocess.run(["cuobjdump", "-lelf", str(E.library_path())], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_probe_fetches_use_the_safe_address_form():
    """ptxas 12.9 can emit LDGSTS [R+UR+imm] with an undefined uniform register for the probe pipe's cp.async
    ("illegal instruction" at run time): every LDGSTS of the library must be in the [R(+imm)] form"""
    import ecloop_b200 as E

    out = subprocess.run(["cuobjdump", "-sass", str(E.library_path())], capture_output=True, text=True).stdout
    ldgsts = [l for l in out.splitlines() if "LDGSTS" in l]
    assert len(ldgsts) >= 12  # six HBM instances of the add kernel
    assert not [l for l in ldgsts if re.search(r"\[R\d+\+UR", l)]


def test_no_cpu_fallback():
    import ecloop_b200 as E

    lib = E.load_library()
    if lib.ecl_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(E.EclError) as ei:
        E.Device(0)
    assert ei.value.code == -1 and "no CPU path" in str(ei.value)


def test_product_never_touches_the_oracle():
    for p in list((ROOT / "ecloop_b200").rglob("*.py")) + list((ROOT / "ecloop_b200").rglob("*.cu*")) + list((ROOT / "ecloop_b200").rglob("*.[ch]")) + list((ROOT / "include").glob("*.h")):
        t = p.read_text()
        assert "import oracle" not in t and "ecl_oracle" not in t and "oracle/" not in t.replace("never links or imports oracle/", ""), p
