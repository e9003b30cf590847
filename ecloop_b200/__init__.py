"""ecloop_b200 — B200 (sm_100a) replacement for the hot path of vladkens/ecloop.

`Device` is a thin ctypes face of the C-ABI in include/ecloop_b200.h (libecloop_b200.so, built in-tree by
ecloop_b200/build.py). `host` mirrors the reference's host-side bookkeeping around that path (filter loading,
job plan, calc_priv, found-line format) so tests and bench.py read like the reference's own commands.

There is no CPU fallback: importing works anywhere (the library loads without a GPU so that its exports can be
checked), but opening a device without CUDA raises EclError.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

from . import build as _build

A33, A65, ENDO = 1, 2, 4
GROUP = 2048
OP_MUL, OP_SQR, OP_ADD, OP_SUB, OP_NEG, OP_INV = range(6)
# only libecloop_b200_exp.so (-DECL_EXPERIMENTAL, tests/test_gpu_experimental.py) implements these:
OP_MUL_F64, OP_MUL_F64_CHAIN, OP_AFFINE_F64_X, OP_AFFINE_F64_Y = range(6, 10)
MUL_DEPTH = 2
E_DEGENERATE = -6

ABI_SYMBOLS = (
    "ecl_abi_version", "ecl_device_count", "ecl_open", "ecl_close", "ecl_last_error", "ecl_set_stream",
    "ecl_set_filter", "ecl_set_stride", "ecl_add_submit", "ecl_mul_submit", "ecl_collect", "ecl_last_elapsed_ms",
    "ecl_set_tuning", "ecl_prim_fp", "ecl_prim_scalar_mul", "ecl_prim_hash160", "ecl_prim_bloom", "ecl_peak_bench",
    "ecl_peak_bench_kind", "ecl_filter_alloc", "ecl_filter_write", "ecl_filter_flush", "ecl_filter_commit",
    "ecl_filter_read", "ecl_filter_copy_peer", "ecl_filter_add", "ecl_filter_generate", "ecl_filter_fill",
    "ecl_host_alloc", "ecl_host_free", "ecl_mul_reserve",
)
PEAK_KINDS = ("lop3", "iadd3", "shf", "imad", "imad_wide", "lop3+imad", "imad_const", "imad_hi", "lop3+imad_const",
              "shf+imad_wide", "lop3+imad_hi", "lop3x5+imad_constx3", "add2", "lop3+imad_wide", "shf+imad", "lop3+shf", "dfma", "dfma+lop3", "dfma+imad")
PEAK_KINDS_EXPERIMENTAL = ("fe_mul_gmuls", "fe6_mul_f64_gmuls", "fe_mul+384alu_gmuls", "fe6_mul_f64+384alu_gmuls")  # kinds 19..22


class EclError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ecloop_b200 error {code}: {msg}")
        self.code = code


class Hit(C.Structure):
    _fields_ = [("key_off", C.c_uint64), ("h160", C.c_uint32 * 5), ("endo", C.c_uint8), ("kind", C.c_uint8),
                ("pad", C.c_uint8 * 2)]


assert C.sizeof(Hit) == 32

_lib = None


def library_path() -> Path:
    return _build.OUT


def load_library(rebuild: bool = False, experimental: bool = False) -> C.CDLL:
    """Load libecloop_b200.so, building it first if the sources changed. Fails loudly if it cannot be had.
    experimental=True loads libecloop_b200_exp.so instead (the same sources with -DECL_EXPERIMENTAL: the FP64-pipe
    field arithmetic and its microbenchmarks; test-only, never cached as the product library)."""
    global _lib
    if _lib is not None and not rebuild and not experimental:
        return _lib
    path = _build.OUT_EXP if experimental else _build.OUT
    override = os.environ.get("ECLOOP_B200_LIB")  # tuning variants (tools/build_variants.py); never a fallback
    if override and not experimental:
        path = Path(override)
        if not path.exists():
            raise EclError(-1, f"ECLOOP_B200_LIB={override} does not exist")
    else:
        try:
            _build.build()
        except Exception as e:  # no nvcc on this box: use the prebuilt library that travelled with the repo
            if not path.exists():
                raise EclError(-1, f"{path.name} is missing and cannot be built here: {e}") from e
    lib = C.CDLL(str(path))
    lib.ecl_last_error.restype = C.c_char_p
    lib.ecl_last_error.argtypes = [C.c_void_p]
    lib.ecl_open.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    lib.ecl_close.argtypes = [C.c_void_p]
    lib.ecl_close.restype = None
    lib.ecl_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.ecl_set_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.ecl_set_stride.argtypes = [C.c_void_p, C.c_void_p]
    lib.ecl_add_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32]
    lib.ecl_mul_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    lib.ecl_collect.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    lib.ecl_last_elapsed_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
    lib.ecl_set_tuning.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    lib.ecl_prim_fp.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    lib.ecl_prim_scalar_mul.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    lib.ecl_prim_hash160.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    lib.ecl_prim_bloom.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    lib.ecl_peak_bench.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.ecl_peak_bench_kind.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.ecl_filter_alloc.argtypes = [C.c_void_p, C.c_uint64]
    lib.ecl_filter_write.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.ecl_filter_flush.argtypes = [C.c_void_p]
    lib.ecl_filter_commit.argtypes = [C.c_void_p]
    lib.ecl_filter_read.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.ecl_filter_copy_peer.argtypes = [C.c_void_p, C.c_void_p]
    lib.ecl_filter_add.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]
    lib.ecl_filter_generate.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_uint64]
    lib.ecl_filter_fill.argtypes = [C.c_void_p]
    lib.ecl_filter_fill.restype = C.c_double
    lib.ecl_host_alloc.argtypes = [C.c_uint64]
    lib.ecl_host_alloc.restype = C.c_void_p
    lib.ecl_host_free.argtypes = [C.c_void_p]
    lib.ecl_host_free.restype = None
    lib.ecl_mul_reserve.argtypes = [C.c_void_p, C.c_uint32]
    if not experimental:
        _lib = lib
    return lib


def _fe(v: int):
    return (C.c_uint64 * 4)(*[(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)])


def _fe_array(vals):
    arr = (C.c_uint64 * (4 * len(vals)))()
    for i, v in enumerate(vals):
        for j in range(4):
            arr[4 * i + j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return arr


def _ints_from(arr, n, limbs=4):
    return [sum(int(arr[limbs * i + j]) << (64 * j) for j in range(limbs)) for i in range(n)]


class Device:
    """One GPU. Methods map 1:1 onto the C-ABI; see include/ecloop_b200.h for the reference call sites."""

    def __init__(self, ordinal: int = 0, experimental: bool = False):
        self._lib = load_library(experimental=experimental)
        self._h = C.c_void_p()
        rc = self._lib.ecl_open(C.byref(self._h), ordinal)
        if rc != 0:
            raise EclError(rc, self._lib.ecl_last_error(None).decode())
        self.ordinal = ordinal

    def _ck(self, rc: int):
        if rc != 0:
            raise EclError(rc, self._lib.ecl_last_error(self._h).decode())

    def close(self):
        if self._h:
            self._lib.ecl_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration
    def set_stream(self, cuda_stream_ptr: int | None):
        self._ck(self._lib.ecl_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def set_filter(self, bits, size_words: int | None = None):
        """bits: a ctypes uint64 array, a numpy uint64 array, or a list of ints (blf_t.bits)."""
        if hasattr(bits, "ctypes"):  # numpy
            n = int(bits.size) if size_words is None else size_words
            self._keep = bits
            self._ck(self._lib.ecl_set_filter(self._h, C.c_void_p(bits.ctypes.data), n))
        else:
            if not isinstance(bits, C.Array):
                bits = (C.c_uint64 * len(bits))(*bits)
            n = len(bits) if size_words is None else size_words
            self._ck(self._lib.ecl_set_filter(self._h, C.cast(bits, C.c_void_p), n))

    # ---- filters on the device (.blf tooling, lib/utils.c:362-475)
    def filter_alloc(self, size_words: int):
        self._ck(self._lib.ecl_filter_alloc(self._h, size_words))

    def filter_write(self, offset_words: int, bits):
        """bits: numpy uint64 array (any host memory)"""
        self._ck(self._lib.ecl_filter_write(self._h, offset_words, C.c_void_p(bits.ctypes.data), int(bits.size)))
        self._ck(self._lib.ecl_filter_flush(self._h))

    def filter_commit(self):
        self._ck(self._lib.ecl_filter_commit(self._h))

    def filter_read(self, offset_words: int, n_words: int):
        import numpy as np

        out = np.empty(n_words, dtype=np.uint64)
        self._ck(self._lib.ecl_filter_read(self._h, offset_words, C.c_void_p(out.ctypes.data), n_words))
        return out

    def filter_copy_peer(self, src: "Device"):
        self._ck(self._lib.ecl_filter_copy_peer(self._h, src._h))

    def filter_add(self, h160_words_list) -> int:
        """blf_gen's insert loop: hashes as 5-tuples of words (or an (n, 5) uint32 numpy array) -> number of new items"""
        if hasattr(h160_words_list, "ctypes"):
            n, ptr = int(h160_words_list.shape[0]), C.c_void_p(h160_words_list.ctypes.data)
            keep = h160_words_list
        else:
            n = len(h160_words_list)
            keep = (C.c_uint32 * (5 * n))(*[w for h in h160_words_list for w in h])
            ptr = C.cast(keep, C.c_void_p)
        new = C.c_uint64(0)
        self._ck(self._lib.ecl_filter_add(self._h, ptr, n, C.byref(new)))
        del keep
        return int(new.value)

    def filter_generate(self, size_words: int, fill: float, seed: int):
        self._ck(self._lib.ecl_filter_generate(self._h, size_words, fill, seed))

    def filter_fill(self) -> float:
        return float(self._lib.ecl_filter_fill(self._h))

    def set_stride(self, stride_k: int):
        self._ck(self._lib.ecl_set_stride(self._h, C.cast(_fe(stride_k), C.c_void_p)))

    def set_tuning(self, groups_per_thread: int = 0, hit_capacity: int = 0):
        self._ck(self._lib.ecl_set_tuning(self._h, groups_per_thread, hit_capacity))

    # ---- hot path
    def add_submit(self, start_pk: int, n_keys: int, flags: int = A33):
        self._ck(self._lib.ecl_add_submit(self._h, C.cast(_fe(start_pk), C.c_void_p), n_keys, flags))

    def mul_submit(self, pks, flags: int = A33):
        arr = _fe_array(pks) if not isinstance(pks, C.Array) else pks
        n = len(arr) // 4
        self._ck(self._lib.ecl_mul_submit(self._h, C.cast(arr, C.c_void_p), n, flags))

    def collect(self, cap: int = 1 << 16):
        """-> (hits, keys_done); hits = [(key_off, endo, kind, (h0..h4))], in the reference's -t 1 order."""
        while True:
            buf = (Hit * cap)()
            n = C.c_uint32(0)
            done = C.c_uint64(0)
            rc = self._lib.ecl_collect(self._h, buf, cap, C.byref(n), C.byref(done))
            if rc == -4 and cap < (1 << 28):  # caller buffer too small: the work is kept, ask again
                cap *= 8
                continue
            self._ck(rc)
            hits = [(int(h.key_off), int(h.endo), int(h.kind), tuple(int(w) for w in h.h160)) for h in buf[: n.value]]
            return hits, int(done.value)

    def batch_add(self, start_pk: int, n_keys: int, flags: int = A33, cap: int = 1 << 16):
        """batch_add + check_found_add up to the bloom decision (main.c:287-403) for one span."""
        self.add_submit(start_pk, n_keys, flags)
        return self.collect(cap)[0]

    def mul_batch(self, pks, flags: int = A33, cap: int = 1 << 16):
        self.mul_submit(pks, flags)
        return self.collect(cap)[0]

    def last_elapsed_ms(self):
        t, h, n = C.c_float(0), C.c_float(0), C.c_uint32(0)
        self._ck(self._lib.ecl_last_elapsed_ms(self._h, C.byref(t), C.byref(h), C.byref(n)))
        return float(t.value), float(h.value), int(n.value)

    # ---- primitives (parity entry points)
    def fp(self, op: int, a, b=None):
        n = len(a)
        out = (C.c_uint64 * (4 * n))()
        pb = C.cast(_fe_array(b), C.c_void_p) if b is not None else None
        self._ck(self._lib.ecl_prim_fp(self._h, op, C.cast(_fe_array(a), C.c_void_p), pb, C.cast(out, C.c_void_p), n))
        return _ints_from(out, n)

    def scalar_mul(self, ks):
        n = len(ks)
        out = (C.c_uint64 * (8 * n))()
        self._ck(self._lib.ecl_prim_scalar_mul(self._h, C.cast(_fe_array(ks), C.c_void_p), C.cast(out, C.c_void_p), n))
        return [(sum(int(out[8 * i + j]) << (64 * j) for j in range(4)),
                 sum(int(out[8 * i + 4 + j]) << (64 * j) for j in range(4))) for i in range(n)]

    def hash160(self, points):
        """points: [(x, y)] -> ([h33 hex], [h65 hex])"""
        n = len(points)
        xy = (C.c_uint64 * (8 * n))()
        for i, (x, y) in enumerate(points):
            for j in range(4):
                xy[8 * i + j] = (x >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
                xy[8 * i + 4 + j] = (y >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
        o33 = (C.c_uint32 * (5 * n))()
        o65 = (C.c_uint32 * (5 * n))()
        self._ck(self._lib.ecl_prim_hash160(self._h, C.cast(xy, C.c_void_p), C.cast(o33, C.c_void_p),
                                            C.cast(o65, C.c_void_p), n))
        fmt = lambda o, i: "".join("%08x" % int(o[5 * i + k]) for k in range(5))  # noqa: E731
        return [fmt(o33, i) for i in range(n)], [fmt(o65, i) for i in range(n)]

    def bloom_has(self, h160_words_list):
        n = len(h160_words_list)
        arr = (C.c_uint32 * (5 * n))(*[w for h in h160_words_list for w in h])
        out = (C.c_uint8 * n)()
        self._ck(self._lib.ecl_prim_bloom(self._h, C.cast(arr, C.c_void_p), C.cast(out, C.c_void_p), n))
        return [bool(b) for b in out]

    def peak_bench(self):
        out = (C.c_double * 8)()
        self._ck(self._lib.ecl_peak_bench(self._h, out))
        keys = ("lop3", "iadd3", "shf", "imad", "imad_wide", "lop3_imad_mix")
        d = {k + "_gops": float(out[i]) for i, k in enumerate(keys)}
        d["sm_mhz"] = float(out[6])
        return d


    def peak_bench_kinds(self):
        """every instruction kind / mix of csrc/peak.cuh -> {name: Gops/s}"""
        out = {}
        for k, name in enumerate(PEAK_KINDS):
            g, mhz = C.c_double(0), C.c_double(0)
            self._ck(self._lib.ecl_peak_bench_kind(self._h, k, C.byref(g), C.byref(mhz)))
            out[name] = round(float(g.value), 1)
        return out


def device_count() -> int:
    return int(load_library().ecl_device_count())
