"""TEST INFRASTRUCTURE ONLY — ctypes face of oracle/ecl_oracle.c (our CPU restatement) and oracle/_ref
(the unmodified reference compiled from /root/reference by oracle/Makefile).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this
package; the product (ecloop_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB = HERE / "libecl_oracle.so"
REF_DIR = HERE / "_ref"

A33, A65, ENDO = 1, 2, 4
N_ORDER = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
P_FIELD = 2**256 - 2**32 - 977

FE = C.c_uint64 * 4


class Hit(C.Structure):
    _fields_ = [
        ("key_off", C.c_uint64),
        ("h160", C.c_uint32 * 5),
        ("endo", C.c_uint8),
        ("kind", C.c_uint8),
        ("pad", C.c_uint8 * 2),
        ("pk", C.c_uint64 * 4),
    ]


class Filter(C.Structure):
    _fields_ = [
        ("bits", C.POINTER(C.c_uint64)),
        ("size", C.c_uint64),
        ("list", C.POINTER(C.c_uint32)),
        ("count", C.c_uint64),
    ]


def build(force: bool = False) -> None:
    """Compile the restatement (and oracle/_ref when /root/reference is present)."""
    if force or not LIB.exists() or LIB.stat().st_mtime < (HERE / "ecl_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-s", "-C", str(HERE), "oracle"], check=True)
    if Path("/root/reference/main.c").exists() and (force or not (REF_DIR / "ecloop_ref").exists()):
        subprocess.run(["make", "-s", "-C", str(HERE), "ref"], check=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.orc_add_span.restype = C.c_uint64
        _lib.orc_add_range.restype = C.c_uint64
        _lib.orc_mul_batch.restype = C.c_uint64
    return _lib


def to_fe(v: int) -> FE:
    return FE(*[(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)])


def from_fe(a) -> int:
    return sum(int(a[i]) << (64 * i) for i in range(4))


def h160_hex(words) -> str:
    return "".join("%08x" % int(w) for w in words)


def hex_to_h160(s: str):
    return [int(s[i : i + 8], 16) for i in range(0, 40, 8)]


# ---------------------------------------------------------------- field / group / hash wrappers


def _fp2(name, a, b):
    r = FE()
    getattr(lib(), name)(r, to_fe(a), to_fe(b))
    return from_fe(r)


def fp_mul(a, b):
    return _fp2("orc_fp_mul", a, b)


def fp_add(a, b):
    return _fp2("orc_fp_add", a, b)


def fp_sub(a, b):
    return _fp2("orc_fp_sub", a, b)


def fn_mul(a, b):
    return _fp2("orc_fn_mul", a, b)


def fp_inv(a):
    r = FE()
    lib().orc_fp_inv(r, to_fe(a))
    return from_fe(r)


def ec_mul_g(k: int):
    x, y = FE(), FE()
    inf = lib().orc_ec_mul_g(x, y, to_fe(k))
    return None if inf else (from_fe(x), from_fe(y))


def hash160_33(x: int, y: int) -> str:
    h = (C.c_uint32 * 5)()
    lib().orc_hash160_33(h, to_fe(x), to_fe(y))
    return h160_hex(h)


def hash160_65(x: int, y: int) -> str:
    h = (C.c_uint32 * 5)()
    lib().orc_hash160_65(h, to_fe(x), to_fe(y))
    return h160_hex(h)


def calc_priv(start: int, stride: int, off: int, endo: int) -> int:
    r = FE()
    lib().orc_calc_priv(r, to_fe(start), to_fe(stride), C.c_uint64(off), C.c_uint8(endo))
    return from_fe(r)


def blf_positions(h160_words, size_words: int):
    pos = (C.c_uint64 * 20)()
    lib().orc_blf_positions(pos, (C.c_uint32 * 5)(*h160_words), C.c_uint64(size_words))
    return list(pos)


# ---------------------------------------------------------------- filters


@dataclass
class HostFilter:
    """bloom bits + optional sorted unique list, as load_filter builds them (main.c:71-131)."""

    bits: "C.Array"
    size: int
    words: list | None  # sorted unique list of 5-tuples, or None (bloom-only)
    _list_arr: object = None

    def c_struct(self) -> Filter:
        f = Filter()
        f.bits = C.cast(self.bits, C.POINTER(C.c_uint64))
        f.size = self.size
        if self.words is not None:
            flat = [w for t in self.words for w in t]
            self._list_arr = (C.c_uint32 * len(flat))(*flat)
            f.list = C.cast(self._list_arr, C.POINTER(C.c_uint32))
            f.count = len(self.words)
        else:
            f.list = None
            f.count = 0
        return f

    def bits_list(self):
        return list(self.bits)


def filter_from_hashes(hex_hashes) -> HostFilter:
    """list mode: sorted unique h160 + bloom of 2*count words (main.c:113-130)."""
    words = sorted({tuple(hex_to_h160(h)) for h in hex_hashes})
    size = 2 * len(words)
    bits = (C.c_uint64 * size)()
    for t in words:
        lib().orc_blf_add(bits, C.c_uint64(size), (C.c_uint32 * 5)(*t))
    return HostFilter(bits, size, words)


def filter_from_text_file(path) -> HostFilter:
    """The reference keeps only 40-character chunks (main.c:96-98); a longer line is consumed in 40-char pieces."""
    hashes = []
    with open(path, "rb") as fh:
        data = fh.read()
    # emulate fgets(buf, 41): chunks end at '\n' or after 40 chars
    i = 0
    while i < len(data):
        j = data.find(b"\n", i, i + 40)
        chunk = data[i : j + 1] if j != -1 else data[i : i + 40]
        i += len(chunk)
        if len(chunk) == 40 and b"\n" not in chunk:
            words = []
            for k in range(0, 40, 8):
                try:
                    words.append(int(chunk[k : k + 8], 16))
                except ValueError:
                    words.append(0)
            hashes.append("".join("%08x" % w for w in words))
    return filter_from_hashes(hashes)


def filter_all_ones(size_words: int = 1) -> HostFilter:
    bits = (C.c_uint64 * size_words)(*([0xFFFFFFFFFFFFFFFF] * size_words))
    return HostFilter(bits, size_words, None)


def filter_bloom_only(hex_hashes, size_words: int) -> HostFilter:
    bits = (C.c_uint64 * size_words)()
    for h in hex_hashes:
        lib().orc_blf_add(bits, C.c_uint64(size_words), (C.c_uint32 * 5)(*hex_to_h160(h)))
    return HostFilter(bits, size_words, None)


# ---------------------------------------------------------------- hot-path drivers


def _hits_to_tuples(arr, n):
    out = []
    for i in range(n):
        h = arr[i]
        out.append((int(h.key_off), int(h.endo), int(h.kind), h160_hex(h.h160), from_fe(h.pk)))
    return out


def add_span(start: int, stride: int, n_keys: int, flags: int, flt: HostFilter, cap: int = 1 << 20):
    """-> (n_found, [(key_off, endo, kind, h160hex, pk)...]) in the reference's -t 1 emission order."""
    hits = (Hit * cap)()
    fs = flt.c_struct()
    n = lib().orc_add_span(to_fe(start), to_fe(stride), C.c_uint64(n_keys), C.c_uint32(flags), C.byref(fs), hits,
                           C.c_uint64(cap))
    return int(n), _hits_to_tuples(hits, min(int(n), cap))


def add_range(range_s: int, range_e: int, ord_offs: int, flags: int, flt: HostFilter, cap: int = 1 << 20):
    hits = (Hit * cap)()
    fs = flt.c_struct()
    kc = C.c_uint64(0)
    n = lib().orc_add_range(to_fe(range_s), to_fe(range_e), C.c_uint32(ord_offs), C.c_uint32(flags), C.byref(fs),
                            hits, C.c_uint64(cap), C.byref(kc))
    return int(n), _hits_to_tuples(hits, min(int(n), cap)), int(kc.value)


def mul_batch(pks, flags: int, flt: HostFilter, cap: int = 1 << 20):
    arr = (FE * len(pks))(*[to_fe(k) for k in pks])
    hits = (Hit * cap)()
    fs = flt.c_struct()
    n = lib().orc_mul_batch(arr, C.c_uint64(len(pks)), C.c_uint32(flags), C.byref(fs), hits, C.c_uint64(cap))
    return int(n), _hits_to_tuples(hits, min(int(n), cap))


def pubkey_hashes(pks):
    """-> list of (x, y, h33hex, h65hex)"""
    n = len(pks)
    arr = (FE * n)(*[to_fe(k) for k in pks])
    o33 = (C.c_uint32 * (5 * n))()
    o65 = (C.c_uint32 * (5 * n))()
    oxy = (C.c_uint64 * (8 * n))()
    lib().orc_pubkey_hashes(arr, C.c_uint64(n), o33, o65, oxy)
    out = []
    for i in range(n):
        x = sum(int(oxy[8 * i + j]) << (64 * j) for j in range(4))
        y = sum(int(oxy[8 * i + 4 + j]) << (64 * j) for j in range(4))
        out.append((x, y, h160_hex(o33[5 * i : 5 * i + 5]), h160_hex(o65[5 * i : 5 * i + 5])))
    return out


def format_found(kind: int, h160hex: str, pk: int, tab: bool = True) -> str:
    """`-o` file line (main.c:193-196) or stdout line (main.c:187-189)."""
    label = "addr33" if kind == 0 else "addr65"
    return ("%s\t%s\t%064x" if tab else "%s: %s <- %064x") % (label, h160hex, pk)


# ---------------------------------------------------------------- the unmodified reference binary


def ref_binary() -> Path | None:
    """oracle/_ref/ecloop_ref (native flags of the reference Makefile) or the portable -march=x86-64-v3 build
    if the native one cannot execute on this CPU."""
    for name in ("ecloop_ref", "ecloop_ref_v3"):
        p = REF_DIR / name
        if p.exists():
            try:
                r = subprocess.run([str(p), "-v"], capture_output=True, timeout=20)
                if r.returncode == 0 and b"ecloop v" in r.stdout:
                    return p
            except Exception:
                continue
    return None


def run_ref(args, stdin_bytes: bytes | None = None, timeout: float = 600.0):
    """Run the reference CLI; returns (returncode, stdout, stderr) as bytes."""
    exe = ref_binary()
    if exe is None:
        raise RuntimeError("oracle/_ref not built (needs /root/reference in the build container)")
    env = dict(os.environ, LC_ALL="C")
    r = subprocess.run([str(exe), *args], input=stdin_bytes, capture_output=True, timeout=timeout, env=env)
    return r.returncode, r.stdout, r.stderr


# ---------------------------------------------------------------- mirrors of the device-side filter tooling (tests only)


def synthetic_filter(size_words: int, fill: float, seed: int):
    """numpy restatement of ecl_filter_generate (csrc/filter_kernels.cuh): word i, bit 8k+b set iff byte b of
    splitmix64(seed + (8i+k+1)*gamma) < round(fill*256). -> numpy uint64 array"""
    import numpy as np

    thr = int(fill * 256.0 + 0.5)
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    out = np.zeros(size_words, dtype=np.uint64)
    with np.errstate(over="ignore"):
        idx = np.arange(size_words, dtype=np.uint64) * np.uint64(8)
        for k in range(8):
            z = (np.uint64(seed) + (idx + np.uint64(k + 1)) * np.uint64(0x9E3779B97F4A7C15)) & M
            z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M
            z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M
            z = z ^ (z >> np.uint64(31))
            for b in range(8):
                byte = (z >> np.uint64(8 * b)) & np.uint64(255)
                out |= (byte < np.uint64(thr)).astype(np.uint64) << np.uint64(8 * k + b)
    return out


def blf_gen_sequential(bits, hashes):
    """blf_gen's insert loop (lib/utils.c:453-465) word for word over python ints: `if has: continue; add; count += 1`.
    bits: list of ints (modified in place); hashes: 5-tuples. -> count"""
    size = len(bits)
    count = 0
    for h in hashes:
        pos = blf_positions(h, size)
        if all((bits[p >> 6] >> (p & 63)) & 1 for p in pos):
            continue
        for p in pos:
            bits[p >> 6] |= 1 << (p & 63)
        count += 1
    return count
