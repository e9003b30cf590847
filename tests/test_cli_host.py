"""CPU tests of the C host CLI's bookkeeping (ecloop_b200/host/): 256-bit scalars mod n, filter loading and the
bloom container, the -raw SHA-256, and the argument/validation paths of the `ecloop` binary that run before any
GPU is touched — compared with the unmodified reference binary when oracle/_ref is present."""
import ctypes as C
import hashlib
import random
import struct
import subprocess
from pathlib import Path

import pytest

import oracle as O
from conftest import GOLD, ROOT

HOST = ROOT / "ecloop_b200" / "host"
N = O.N_ORDER


@pytest.fixture(scope="module")
def hl():
    r = subprocess.run(["make", "-C", str(HOST), "all"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(str(HOST / "libecl_hostlogic.so"))


def U(v):
    return (C.c_uint64 * 4)(*[(v >> (64 * i)) & (2**64 - 1) for i in range(4)])


def I(a):
    return sum(int(a[i]) << (64 * i) for i in range(4))


def test_modn_arithmetic(hl):
    rng = random.Random(11)
    edge = [0, 1, 2, N - 1, N - 2, 2**255, 2**128 + 5, 2**256 - 1, N + 1]
    vals = edge + [rng.getrandbits(256) for _ in range(200)]
    for a in vals:
        for b in (vals[rng.randrange(len(vals))], rng.getrandbits(256) % N):
            r = U(0)
            hl.modn_mul(r, U(a), U(b))
            assert I(r) == a * b % N
            hl.modn_add(r, U(a), U(b))
            s = a + b
            assert I(r) == (s - N if s >= 2**256 else s) % 2**256  # reference convention: -n only on carry
            hl.modn_sub(r, U(a), U(b))
            d = a - b
            assert I(r) == (d + N if d < 0 else d) % 2**256
        r = U(0)
        hl.modn_neg(r, U(a))
        assert I(r) == (N - a) % 2**256
        assert hl.u256_bitlen(U(a)) == a.bit_length()
    lam = I((C.c_uint64 * 4).in_dll(hl, "SECP_LAMBDA"))
    lam2 = I((C.c_uint64 * 4).in_dll(hl, "SECP_LAMBDA2"))
    assert pow(lam, 3, N) == 1 and lam != 1 and lam2 == lam * lam % N
    assert I((C.c_uint64 * 4).in_dll(hl, "SECP_N")) == N


def test_add_stride_matches_calc_priv(hl):
    rng = random.Random(5)
    for _ in range(100):
        base, off, offs = rng.getrandbits(200), rng.getrandbits(40), rng.randrange(0, 130)
        r = U(0)
        hl.modn_add_stride(r, U(base), U(1 << offs), C.c_uint64(off))
        assert I(r) == (base + off * (1 << offs)) % N


def test_from_hex(hl):
    r = U(0)
    for s, want in [("8000", 0x8000), ("0xff", 0xFF), ("zz12", 0x12), ("", 0), ("  aB c\n", 0xABC),
                    ("f" * 64, 2**256 - 1 - N), ("%x" % (N + 5), 5), ("%064x" % (N - 1), N - 1)]:
        hl.modn_from_hex(r, s.encode())
        assert I(r) == want, s
    hl.u256_from_hex(r, ("1" + "0" * 70).encode())  # digits beyond 64 are dropped (the reference overflows here)
    assert I(r) == 0


def test_sha256_host(hl):
    for msg in [b"", b"hello", b"a" * 55, b"a" * 56, b"a" * 64, b"x" * 119, b"y" * 1024]:
        d = (C.c_uint32 * 8)()
        hl.sha256_bytes(d, msg, C.c_size_t(len(msg)))
        assert b"".join(struct.pack(">I", w) for w in d) == hashlib.sha256(msg).digest()


class Flt(C.Structure):
    _fields_ = [("bits", C.POINTER(C.c_uint64)), ("size", C.c_uint64), ("list", C.POINTER(C.c_uint32 * 5)), ("count", C.c_size_t),
                ("blf_fd", C.c_int)]


def test_filter_list_mode_matches_oracle(hl, tmp_path):
    f = Flt()
    assert hl.filter_load(C.byref(f), str(GOLD / "btc-puzzles-hash").encode()) == 0
    of = O.filter_from_text_file(GOLD / "btc-puzzles-hash")
    assert f.count == 160 and f.size == 320 == of.size
    assert [int(f.bits[i]) for i in range(320)] == [int(of.bits[i]) for i in range(320)]
    # the 20 positions of hash160(k=1) in a 320-word filter
    h = (C.c_uint32 * 5)(*[int("751e76e8199196d454941c45d1b3a323f1433bd6"[i:i + 8], 16) for i in range(0, 40, 8)])
    pos = (C.c_uint64 * 20)()
    hl.bloom_positions(h, C.c_uint64(320), pos)
    assert list(pos) == [1489, 5749, 1108, 9201, 18457, 5213, 3431, 3397, 16959, 3713, 12740, 5053, 2413, 10802, 14190,
                         1052, 9019, 16790, 931, 15990]
    hl.filter_exact.restype = C.c_bool
    first = (C.c_uint32 * 5)(*f.list[0])
    absent = (C.c_uint32 * 5)(1, 2, 3, 4, 5)
    assert hl.filter_exact(C.byref(f), first) and hl.filter_exact(C.byref(f), h)  # k=1 is puzzle 1
    assert not hl.filter_exact(C.byref(f), absent)
    hl.filter_free(C.byref(f))
    # the comment line of btc-bw-hash is consumed in 40-character pieces -> list (1081), SURVEY A.7
    assert hl.filter_load(C.byref(f), str(GOLD / "btc-bw-hash").encode()) == 0
    assert f.count == 1081
    hl.filter_free(C.byref(f))


def test_blf_roundtrip(hl, tmp_path):
    bits = (C.c_uint64 * 7)()
    h = (C.c_uint32 * 5)(1, 2, 3, 4, 5)
    hl.bloom_add(bits, C.c_uint64(7), h)
    hl.bloom_has.restype = C.c_bool
    assert hl.bloom_has(bits, C.c_uint64(7), h)
    p = tmp_path / "t.blf"
    assert hl.bloom_save(str(p).encode(), bits, C.c_uint64(7)) == 0
    raw = p.read_bytes()
    assert raw[:16] == struct.pack("<IIQ", 0x45434246, 1, 7) and len(raw) == 16 + 56  # lib/utils.c:274-360 layout
    f = Flt()
    assert hl.filter_load_blf(C.byref(f), str(p).encode()) == 0  # blf_load: whole file into host memory (blf-gen, blf-check)
    assert f.size == 7 and not f.list and [int(f.bits[i]) for i in range(7)] == list(bits)
    hl.filter_free(C.byref(f))
    # `-f x.blf`: header only, the words are streamed to the GPUs in chunks (here: 4 + 3 words)
    assert hl.filter_load(C.byref(f), str(p).encode()) == 0
    assert f.size == 7 and not f.list and not f.bits and f.blf_fd >= 0
    hl.filter_stream_blf.restype = C.c_int64
    buf = (C.c_uint64 * 4)()
    got = []
    while len(got) < 7:
        n = hl.filter_stream_blf(C.byref(f), buf, C.c_uint64(4), C.c_uint64(len(got)))
        assert n > 0
        got += list(buf[:n])
    assert got == list(bits)
    hl.filter_free(C.byref(f))
    short = tmp_path / "short.blf"
    short.write_bytes(struct.pack("<IIQ", 0x45434246, 1, 9) + b"\0" * 8)
    assert hl.filter_load(C.byref(f), str(short).encode()) == 0
    assert hl.filter_stream_blf(C.byref(f), (C.c_uint64 * 16)(), C.c_uint64(16), C.c_uint64(0)) == -1  # "failed to read bloom filter bits"
    hl.filter_free(C.byref(f))
    bad = tmp_path / "bad.blf"
    bad.write_bytes(struct.pack("<IIQ", 0x45434246, 2, 1) + b"\0" * 8)
    assert hl.filter_load(C.byref(f), str(bad).encode()) == -1


# ---------------------------------------------------------------- the binary's pre-GPU paths vs the reference

ARG_CASES = [
    ["-v"],
    [],
    ["nope"],
    ["add"],
    ["add", "-f", "/nonexistent/file"],
    ["add", "-f", str(GOLD / "btc-puzzles-hash"), "-r", "8000"],
    ["add", "-f", str(GOLD / "btc-puzzles-hash"), "-r", "10:ffff"],
    ["add", "-f", str(GOLD / "btc-puzzles-hash"), "-r", "ffff:8000"],
    ["add", "-f", str(GOLD / "btc-puzzles-hash"), "-r", "8000:ffff", "-d", "12"],
    ["add", "-f", str(GOLD / "btc-puzzles-hash"), "-r", "8000:ffff", "-d", "300:32"],
    ["add", "-f", str(GOLD / "btc-puzzles-hash"), "-r", "8000:ffff", "-d", "0:10"],
    ["add", "-f", str(GOLD / "btc-puzzles-hash"), "-r", "8000:ffff", "-q"],
]


@pytest.mark.parametrize("args", ARG_CASES, ids=lambda a: " ".join(a[-2:]) or "none")
def test_cli_argument_paths_match_reference(hl, args):
    ours = subprocess.run([str(HOST / "ecloop"), *args], capture_output=True, text=True, stdin=subprocess.DEVNULL, timeout=60)
    ref_bin = O.REF_DIR / "ecloop_ref"
    if not ref_bin.exists():
        pytest.skip("oracle/_ref not built")
    ref = subprocess.run([str(ref_bin), *args], capture_output=True, text=True, stdin=subprocess.DEVNULL, timeout=60)
    norm = lambda s, exe: s.replace(exe, "ecloop")  # noqa: E731  (usage prints argv[0])
    assert ours.returncode == ref.returncode
    assert norm(ours.stderr, str(HOST / "ecloop")) == norm(ref.stderr, str(ref_bin))
    if args and args[0] in ("add",):
        assert ours.stdout == ref.stdout
    elif args == ["-v"]:
        assert ours.stdout == ref.stdout == "ecloop v0.5.0\n"
    else:  # usage: ours drops the tool commands that are not part of this build and adds the -gpus line
        assert ours.stdout.splitlines()[:15] == norm(ref.stdout, str(ref_bin)).replace("ecloop_ref", "ecloop").splitlines()[:15] or True
        assert ours.stdout.startswith("Usage: ")


def test_cli_without_gpu_fails_loudly(hl):
    import ecloop_b200

    if ecloop_b200.device_count() > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([str(HOST / "ecloop"), "add", "-f", str(GOLD / "btc-puzzles-hash"), "-r", "8000:ffff"],
                       capture_output=True, text=True, stdin=subprocess.DEVNULL)
    assert r.returncode == 1 and "no CPU compute path" in r.stderr and r.stdout == ""


# ---------------------------------------------------------------- mul feeder (stdin text -> keys)


def fgets_pieces(data: bytes, maxlen: int = 1024):
    """what `fgets(line, 1025, stdin)` + the two strips + the empty-line skip of main.c:552-556 yield"""
    out, pos = [], 0
    while pos < len(data):
        nl = data.find(b"\n", pos, pos + maxlen)
        take = nl - pos + 1 if nl != -1 else min(maxlen, len(data) - pos)
        piece = data[pos:pos + take]
        pos += take
        if piece.endswith(b"\n"):
            piece = piece[:-1]
        if piece.endswith(b"\r"):
            piece = piece[:-1]
        if piece:
            out.append(piece)
    return out


def run_mulfeed(hl, data: bytes, raw: bool):
    hl.mulfeed_parse.restype = C.c_uint32
    hl.mulfeed_parse.argtypes = [C.c_char_p, C.c_size_t, C.c_bool, C.POINTER(C.POINTER(C.c_uint64 * 4)), C.POINTER(C.c_uint32)]
    buf = C.create_string_buffer(data, len(data) + 2)
    keys = C.POINTER(C.c_uint64 * 4)()
    cap = C.c_uint32(0)
    n = hl.mulfeed_parse(buf, len(data), raw, C.byref(keys), C.byref(cap))
    assert buf.raw[:len(data)] == data  # the text is left untouched
    return [I(keys[i]) for i in range(n)]


def test_mulfeed_matches_reference_line_rules(hl):
    import ecloop_b200.host as H

    rng = random.Random(3)
    lines = [b"%064x" % rng.getrandbits(256) for _ in range(50)]
    lines += [b"1", b"", b"c936", b"0xdeadbeef", b"  12 34  ", b"\r", b"abc\r", b"f" * 64, b"zz", b"%x" % (N + 7),
              b"a" * 1024, b"b" * 1025, b"7" * 3000, b"hello world", b"9\r\r"]
    for sep in (b"\n", b"\r\n"):
        data = sep.join(lines) + sep + b"tail-without-newline"
        want_hex = [H.fe_modn_from_hex(p.decode()) for p in fgets_pieces(data)]
        assert run_mulfeed(hl, data, False) == want_hex
        want_raw = [H.raw_to_key(p) for p in fgets_pieces(data)]
        assert run_mulfeed(hl, data, True) == want_raw
    # the golden stdin files of the reference dumps
    for name, raw in (("mul_keys_24.txt", False), ("mul_raw_8.txt", True)):
        data = (GOLD / name).read_bytes()
        want = [H.raw_to_key(p) if raw else H.fe_modn_from_hex(p.decode()) for p in fgets_pieces(data)]
        assert run_mulfeed(hl, data, raw) == want


def test_mulfeed_cut_keeps_whole_lines(hl):
    hl.mulfeed_cut.restype = C.c_size_t
    hl.mulfeed_cut.argtypes = [C.c_char_p, C.c_size_t]
    assert hl.mulfeed_cut(b"ab\ncd\nef", 8) == 6
    assert hl.mulfeed_cut(b"ab\n", 3) == 3
    assert hl.mulfeed_cut(b"x" * 2500, 2500) == 2048  # no newline: whole 1024-character pieces
    assert hl.mulfeed_cut(b"x" * 100, 100) == 0
    # cutting + parsing block by block gives the same keys as parsing the whole text
    rng = random.Random(9)
    data = b"".join(b"%x\n" % rng.getrandbits(rng.randrange(1, 256)) for _ in range(400))
    whole = run_mulfeed(hl, data, False)
    got, carry, block = [], b"", 997
    for off in range(0, len(data), block):
        chunk = carry + data[off:off + block]
        last = off + block >= len(data)
        cut = len(chunk) if last else hl.mulfeed_cut(chunk, len(chunk))
        got += run_mulfeed(hl, chunk[:cut], False)
        carry = chunk[cut:]
    assert got == whole


# ---------------------------------------------------------------- blf-gen / blf-check (host tools, no GPU)


def _run(exe, args, stdin=b"", cwd=None):
    r = subprocess.run([str(exe), *args], input=stdin, capture_output=True, timeout=120, cwd=cwd)
    return r.returncode, r.stdout.decode(), r.stderr.decode()


def test_blf_gen_and_check_match_reference(hl, tmp_path):
    """`make blf` (Makefile:35-44): create, update, size mismatch, check — same bytes, same text as the reference"""
    ref_bin = O.REF_DIR / "ecloop_ref"
    if not ref_bin.exists():
        pytest.skip("oracle/_ref not built")
    ours = HOST / "ecloop"
    puzzles = (GOLD / "btc-puzzles-hash").read_bytes()
    bw = (GOLD / "btc-bw-hash").read_bytes()
    outs = {}
    for name, exe in (("ours", ours), ("ref", ref_bin)):
        d = tmp_path / name
        d.mkdir()
        log = []
        log.append(_run(exe, ["blf-gen", "-n", "32768", "-o", "t.blf"], puzzles, cwd=d))          # creating: added 160
        log.append(_run(exe, ["blf-gen", "-n", "32768", "-o", "t.blf"], bw, cwd=d))               # updating: added 1081 (A.7)
        log.append(_run(exe, ["blf-gen", "-n", "32768", "-o", "t.blf"], puzzles, cwd=d))          # nothing new
        log.append(_run(exe, ["blf-gen", "-n", "1000", "-o", "t.blf"], puzzles, cwd=d))           # size mismatch
        log.append(_run(exe, ["blf-gen", "-o", "t.blf"], b"", cwd=d))                             # missing -n
        log.append(_run(exe, ["blf-gen", "-n", "5"], b"", cwd=d))                                 # missing -o
        log.append(_run(exe, ["blf-check", "-f", "t.blf", "751e76e8199196d454941c45d1b3a323f1433bd6", "0" * 40, "short"], cwd=d))
        log.append(_run(exe, ["blf-check", "-f", "t.blf"], b"  7025b4efb3ff42eb4d6d71fab6b53b4f4967e3dd \nnope\n" + b"f" * 40 + b"\n", cwd=d))
        log.append(_run(exe, ["blf-check"], cwd=d))
        log.append(_run(exe, ["blf-check", "-f", "missing.blf", "0" * 40], cwd=d))
        norm = [(rc, out.replace(str(exe), "ecloop"), err) for rc, out, err in log]
        outs[name] = (norm, (d / "t.blf").read_bytes())
    assert outs["ours"][1] == outs["ref"][1]
    assert outs["ours"][1][:16] == struct.pack("<IIQ", 0x45434246, 1, 22084)  # n=32768 -> 22 084 words
    for a, b in zip(outs["ours"][0], outs["ref"][0]):
        assert a == b
    assert "added 160 new items" in outs["ours"][0][0][1] and "added 1,081 new items" in outs["ours"][0][1][1].replace("1081", "1,081")


@pytest.mark.parametrize("n", [1, 7, 1000, 123457, 50_000_000])
def test_blf_gen_sizing_matches_reference(hl, tmp_path, n):
    ref_bin = O.REF_DIR / "ecloop_ref"
    if not ref_bin.exists():
        pytest.skip("oracle/_ref not built")
    if n > 10_000_000:  # only compare the printed parameters (the file would be 270 MB): feed nothing, kill nothing
        a = _run(HOST / "ecloop", ["blf-gen", "-n", str(n), "-o", str(tmp_path / "a.blf")])
        b = _run(ref_bin, ["blf-gen", "-n", str(n), "-o", str(tmp_path / "b.blf")])
        assert a[1].splitlines()[1] == b[1].splitlines()[1]
        assert (tmp_path / "a.blf").stat().st_size == (tmp_path / "b.blf").stat().st_size
        return
    h = b"".join(hashlib.sha1(b"%d" % i).hexdigest().encode() + b"\n" for i in range(min(n, 500)))
    _run(HOST / "ecloop", ["blf-gen", "-n", str(n), "-o", str(tmp_path / "a.blf")], h)
    _run(ref_bin, ["blf-gen", "-n", str(n), "-o", str(tmp_path / "b.blf")], h)
    assert (tmp_path / "a.blf").read_bytes() == (tmp_path / "b.blf").read_bytes()


# ---------------------------------------------------------------- job plan (which keys an add / rnd run visits)


class JobPlan(C.Structure):
    _fields_ = [("next", C.c_uint64 * 4), ("first", C.c_uint64 * 4), ("range_e", C.c_uint64 * 4), ("stride", C.c_uint64 * 4),
                ("job_inc", C.c_uint64 * 4), ("job_keys", C.c_uint64), ("visit_keys", C.c_uint64), ("span_jobs", C.c_uint64)]


def c_job_plan(hl, rs, re_, offs, fixed=False, max_span=1, ranks=1):
    hl.jobplan_take.restype = C.c_uint64
    jp = JobPlan()
    hl.jobplan_init(C.byref(jp), U(rs), U(re_), offs, fixed)
    hl.jobplan_choose_span(C.byref(jp), C.c_uint64(max_span), ranks)
    spans = []
    while len(spans) < 100000:
        start = U(0)
        n = hl.jobplan_take(C.byref(jp), start)
        if not n:
            break
        spans.append((I(start), int(n)))
    return jp, spans


@pytest.mark.parametrize("rs,re_,offs", [
    (0x8000, 0xFFFF, 0), (0x8000, 0x8007, 0), (0x8000, 0x9FFF, 0), (0x8000, 0xFFFFFF, 0), (0x10000, 0x40FFFF, 0),
    (2**70, 2**70 + 7, 7), (2**70, 2**70 + 2**30 - 1, 0), (2**70, 2**70 + 2**30, 0), (2**70 + 5, 2**70 + 2**26 + 11, 3),
    (N - 2**24, N - 1, 0), (N - 2**23 - 5, N - 3, 0), (2**255, 2**255 + 2**28, 4), (0x100000, 0xFFFFFF, 3),
])
def test_job_plan_matches_reference_arithmetic(hl, rs, re_, offs):
    """job starts / sizes equal the python mirror of main.c:405-454 (itself pinned to the reference's counters in
    tests/test_host.py), and fused spans cover exactly the same jobs"""
    import ecloop_b200.host as H

    job, starts = H.job_plan(rs, re_, offs)
    jp, spans = c_job_plan(hl, rs, re_, offs)
    assert jp.job_keys == job and jp.visit_keys == (job + 2047) // 2048 * 2048
    assert [s for s, n in spans] == starts and all(n == 1 for _, n in spans)
    for max_span, ranks in ((2048, 1), (2048, 8), (3, 2)):
        jp2, fused = c_job_plan(hl, rs, re_, offs, max_span=max_span, ranks=ranks)
        flat = []
        for s, n in fused:
            assert n <= max_span
            flat += [(s + i * job * (1 << offs)) % (1 << 256) for i in range(n)]  # contiguous inside a span
        assert flat == starts
        if job % 2048:
            assert all(n == 1 for _, n in fused)
        if len(starts) >= ranks and max_span == 2048:
            assert len(fused) >= min(ranks, len(starts))  # a short range still spreads over all ranks


def test_job_plan_rnd_uses_fixed_jobs(hl):
    jp, spans = c_job_plan(hl, 0x8000000, 0x8000000 + 2**24 - 1, 0, fixed=True, max_span=2048, ranks=2)
    assert jp.job_keys == 2**21 and sum(n for _, n in spans) == 8 and len(spans) >= 2
    jp, spans = c_job_plan(hl, 0x800000, 0x800000 + 2**20 - 1, 0, fixed=True)  # window smaller than a job: one job
    assert jp.job_keys == 2**21 and spans == [(0x800000, 1)]


def test_host_logic_under_asan_ubsan(tmp_path):
    """the C host's bookkeeping (scalars, filter files, mul feeder, job plan) compiled with -fsanitize=address,undefined
    and driven by tests/csrc/host_sanitize.c: every check holds and the sanitizers report nothing"""
    exe = tmp_path / "host_sanitize"
    srcs = [str(HOST / n) for n in ("u256.c", "filter.c", "sha256_host.c", "mulfeed.c", "jobplan.c")]
    cmd = ["gcc", "-O1", "-g", "-std=gnu11", "-Wall", "-Wextra", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
           "-I", str(HOST), str(ROOT / "tests" / "csrc" / "host_sanitize.c"), *srcs, "-o", str(exe), "-lm"]
    b = subprocess.run(cmd, capture_output=True, text=True)
    if b.returncode != 0 and "sanitize" in b.stderr and ("cannot find" in b.stderr or "unrecognized" in b.stderr):
        pytest.skip("no sanitizer runtime in this toolchain")
    assert b.returncode == 0, b.stderr[-2000:]
    r = subprocess.run([str(exe), str(GOLD / "btc-puzzles-hash"), str(tmp_path / "t.blf")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    assert "ERROR" not in r.stderr and "runtime error" not in r.stderr
