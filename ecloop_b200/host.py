"""Host-side mirror of the reference's bookkeeping around the hot path (python stand-in for main.c's cold code;
the C host in ecloop_b200/host/ does the same for the drop-in CLI).

Nothing here touches a curve point or a hash: filter files, the job plan, private-key recovery for a reported
hit, and the found-line format. Names follow the reference (load_filter, cmd_add, cmd_mul, calc_priv ...).
"""
from __future__ import annotations

import hashlib
import struct
from dataclasses import dataclass, field

import numpy as np

from . import A33, A65, ENDO, GROUP, Device

N_ORDER = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
P_FIELD = 2**256 - 2**32 - 977
LAMBDA1 = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72  # A1, lib/ecc.c:36
LAMBDA2 = 0xAC9C52B33FA3CF1F5AD9E3FD77ED9BA4A880B9FC8EC739C2E0CFC810B51283CE  # A2, lib/ecc.c:37
MAX_JOB_SIZE = 2 * 1024 * 1024  # main.c:16
BLF_MAGIC, BLF_VERSION = 0x45434246, 1  # lib/utils.c:274-275
M64 = (1 << 64) - 1


# ---------------------------------------------------------------- bloom filter files (lib/utils.c:282-396)


def blf_positions(h, size_words: int):
    """the 20 bit positions of h160 words h[0..4] (blf_add / blf_has, lib/utils.c:290-326)"""
    a = [(h[0] << 32 | h[1]), (h[2] << 32 | h[3]), (h[4] << 32 | h[0]), (h[1] << 32 | h[2]), (h[3] << 32 | h[4])]
    out = []
    for s in (24, 28, 36, 40):
        for i in range(5):
            v = ((a[i] << s) | (a[(i + 1) % 5] >> s)) & M64
            out.append(v % (size_words * 64))
    return out


def blf_add(bits: np.ndarray, h) -> None:
    for p in blf_positions(h, bits.size):
        bits[p >> 6] |= np.uint64(1 << (p & 63))


def blf_has(bits: np.ndarray, h) -> bool:
    return all((int(bits[p >> 6]) >> (p & 63)) & 1 for p in blf_positions(h, bits.size))


def blf_save(path, bits: np.ndarray) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<IIQ", BLF_MAGIC, BLF_VERSION, bits.size))
        f.write(bits.astype("<u8").tobytes())


def blf_load(path) -> np.ndarray:
    with open(path, "rb") as f:
        hdr = f.read(16)
        if len(hdr) != 16:
            raise ValueError("failed to read bloom filter header")
        magic, ver, size = struct.unpack("<IIQ", hdr)
        if magic != BLF_MAGIC or ver != BLF_VERSION:
            raise ValueError("invalid bloom filter version; create a new filter with blf-gen command")
        bits = np.fromfile(f, dtype="<u8", count=size)
        if bits.size != size:
            raise ValueError("failed to read bloom filter bits")
    return bits.astype(np.uint64)


@dataclass
class Filter:
    """ctx->blf + ctx->to_find_hashes (main.c:48-51)"""

    bits: np.ndarray                      # uint64 words
    hashes: list | None = None            # sorted unique 5-tuples (list mode) or None (bloom-only mode)
    _set: set = field(default_factory=set, repr=False)

    def __post_init__(self):
        if self.hashes is not None:
            self._set = set(self.hashes)

    def check_exact(self, h) -> bool:
        """second stage of ctx_check_hash (main.c:214-216); the bloom stage already ran on the GPU"""
        return True if self.hashes is None else tuple(h) in self._set

    @property
    def label(self) -> str:
        return "bloom" if self.hashes is None else f"list ({len(self.hashes):,})"


def filter_from_hashes(hashes) -> Filter:
    """list mode: sort, dedupe, bloom of 2*count words (main.c:113-130). hashes: hex strings or 5-tuples."""
    words = sorted({tuple(int(h[i : i + 8], 16) for i in range(0, 40, 8)) if isinstance(h, str) else tuple(h) for h in hashes})
    bits = np.zeros(2 * len(words), dtype=np.uint64)
    for t in words:
        blf_add(bits, t)
    return Filter(bits, words)


def load_filter(path) -> Filter:
    """load_filter (main.c:71-131): `.blf` -> bloom-only; else text, one 40-hex-digit hash per line. Like the
    reference's fgets(41) loop, only chunks of exactly 40 characters count, and longer lines are consumed in
    40-character pieces (the comment line of data/btc-bw-hash becomes one bogus entry)."""
    path = str(path)
    if path.endswith(".blf"):
        return Filter(blf_load(path), None)
    data = open(path, "rb").read()
    hashes, i = [], 0
    while i < len(data):
        j = data.find(b"\n", i, i + 40)
        chunk = data[i : j + 1] if j != -1 else data[i : i + 40]
        i += len(chunk)
        if len(chunk) == 40 and b"\n" not in chunk:
            words = []
            for k in range(0, 40, 8):
                try:
                    words.append(int(chunk[k : k + 8], 16))  # sscanf("%8x")
                except ValueError:
                    words.append(0)
            hashes.append(tuple(words))
    return filter_from_hashes(hashes)


# ---------------------------------------------------------------- private-key recovery (main.c:267-276)


def calc_priv(start_pk: int, stride_k: int, pk_off: int, endo: int) -> int:
    pk = (start_pk + pk_off * stride_k) % N_ORDER
    if endo == 0:
        return pk
    if endo in (2, 3):
        pk = pk * LAMBDA1 % N_ORDER
    if endo in (4, 5):
        pk = pk * LAMBDA2 % N_ORDER
    if endo in (1, 3, 5):
        pk = (N_ORDER - pk) % N_ORDER
    return pk


def format_found(kind: int, h160, pk: int, tab: bool = True) -> str:
    """ctx_write_found (main.c:182-203): `-o` file line (tab=True) or stdout line"""
    label = "addr33" if kind == 0 else "addr65"
    hx = "".join("%08x" % w for w in h160)
    return ("%s\t%s\t%064x" if tab else "%s: %s <- %064x") % (label, hx, pk)


def fe_modn_from_hex(s: str) -> int:
    """fe_modn_from_hex (lib/ecc.c:81-95,262-265): right to left, non-hex characters skipped, one conditional -n"""
    v, cnt = 0, 0
    for ch in reversed(s):
        if ch in "0123456789abcdefABCDEF":
            if cnt < 64:
                v |= int(ch, 16) << (4 * cnt)
            cnt += 1
    return v - N_ORDER if v >= N_ORDER else v


def raw_to_key(line: bytes) -> int:
    """-raw (main.c:506-527): key = SHA-256(line), big-endian, not reduced"""
    return int.from_bytes(hashlib.sha256(line).digest(), "big")


# ---------------------------------------------------------------- commands


@dataclass
class Found:
    kind: int
    h160: tuple
    pk: int
    endo: int = 0

    def line(self, tab=True):
        return format_found(self.kind, self.h160, self.pk, tab)


def job_plan(range_s: int, range_e: int, ord_offs: int = 0):
    """cmd_add + cmd_add_worker's dispenser (main.c:405-454) -> (job_size, [job start keys])"""
    stride = 1 << ord_offs
    rsz = (range_e - range_s) % N_ORDER
    job = rsz if rsz < MAX_JOB_SIZE else MAX_JOB_SIZE
    inc = job * stride % N_ORDER
    starts, cur = [], range_s
    while not (cur >= range_e or cur < range_s):
        starts.append(cur)
        cur += inc
        if cur >= 1 << 256:  # fe_modn_add: subtract n only on a 2^256 carry (lib/ecc.c:174-187)
            cur -= N_ORDER
    return job, starts


class Searcher:
    """ctx_t + the three compute commands, with the GPU in place of the worker threads."""

    def __init__(self, dev: Device, flt: Filter, flags: int = A33):
        self.dev, self.flt, self.flags = dev, flt, flags
        dev.set_filter(np.ascontiguousarray(flt.bits, dtype=np.uint64))
        self.k_checked = 0
        self.found: list[Found] = []
        self.bloom_positives = 0

    def _take(self, hits, start_pk, stride):
        for key_off, endo, kind, h in hits:
            self.bloom_positives += 1
            if self.flt.check_exact(h):
                self.found.append(Found(kind, h, calc_priv(start_pk, stride, key_off, endo), endo))

    def cmd_add(self, range_s: int, range_e: int, ord_offs: int = 0, jobs_per_submit: int = 512):
        """`ecloop add -r range_s:range_e [-d offs:..]`: same keys visited, same hits, same k_checked.
        Consecutive reference jobs are contiguous (job i starts at range_s + i*job*stride), so they are fused
        into large spans for the GPU; the last job overshoots exactly like the reference (A.1)."""
        stride = 1 << ord_offs
        self.dev.set_stride(stride)
        job, starts = job_plan(range_s, range_e, ord_offs)
        visit = (job + GROUP - 1) // GROUP * GROUP
        i = 0
        while i < len(starts):
            n = min(jobs_per_submit, len(starts) - i) if visit == job else 1
            hits = self.dev.batch_add(starts[i], visit * n, self.flags)
            self._take(hits, starts[i], stride)
            self.k_checked += job * n * (6 if self.flags & ENDO else 1)  # main.c:431
            i += n
        return self.found

    def cmd_mul(self, keys, batch: int = 1 << 20):
        """`ecloop mul` after parsing: keys are integers (fe_modn_from_hex / raw_to_key)"""
        for b in range(0, len(keys), batch):
            part = keys[b : b + batch]
            for key_off, endo, kind, h in self.dev.mul_batch(part, self.flags & (A33 | A65)):
                self.bloom_positives += 1
                if self.flt.check_exact(h):
                    self.found.append(Found(kind, h, part[key_off]))
            self.k_checked += len(part)
        return self.found
