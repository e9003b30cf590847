"""Static checks of the compiled sm_100a code (no GPU needed): the properties DESIGN.md claims for the hot kernels are
read back from the SASS of the in-tree build, so a change that silently loses one of them fails here, not on the box."""
import json
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib_sass():
    import ecloop_b200 as E

    E.load_library()
    return subprocess.run(["cuobjdump", "-sass", str(E.library_path())], capture_output=True, text=True).stdout


def functions(sass):
    out, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            out[name].append(line)
    return out


def test_tables_are_staged_by_tma_and_probes_by_cp_async(lib_sass):
    f = functions(lib_sass)
    add = {k: v for k, v in f.items() if "add_kernel" in k}
    assert len(add) == 12  # six flag combinations x {filter in shared memory, filter in HBM}
    for name, body in add.items():
        text = "\n".join(body)
        assert "UBLKCP" in text, name  # cp.async.bulk global -> shared (table, step point, small filters)
        assert "SYNCS" in text, name   # the mbarrier it completes on
    hbm = [k for k in add if re.search(r"Lb1E(Li[12]E)?Ev9AddParams$", k)]
    assert len(hbm) == 6
    for name in hbm:
        assert any("LDGSTS" in l for l in add[name]), name


def test_pipelined_loop_is_branch_free_and_counts_match_the_design():
    """the pass-2 loop of add_kernel_sp<A33>: no branch between its barrier and its probe code apart from the probe's own
    (hit / filter) branches, ~309 IMAD.WIDE and ~2 400 ALU issue slots per key (profiles/add_kernel_traffic.json)"""
    r = subprocess.run(["python", str(ROOT / "tools" / "loop_census.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    c = json.loads((ROOT / "profiles" / "add_kernel_traffic.json").read_text())["loop_census"]
    assert 3200 <= c["loop_instructions_per_key"] <= 3500
    assert 295 <= c["imad_wide_per_key"] <= 325      # 3.5 products x 73 + 1 squaring x 45 + the 9 fold products
    assert 2300 <= c["alu_issue_slots_per_key"] <= 2480
    obj = ROOT / "build" / "obj" / "add_inst_1.o"
    sass = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
    body = functions(sass)
    sp = next(v for k, v in body.items() if "add_kernel_sp" in k)
    ops = [re.sub(r"^@!?U?P\d\s+", "", re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l).group(1).strip()).split()[0] for l in sp]
    # the two hashes of a step: > 1500 consecutive instructions without a branch, twice
    runs, cur = [], 0
    for op in ops:
        if op.startswith(("BRA", "BSSY", "BSYNC", "CALL", "RET", "EXIT")):
            runs.append(cur)
            cur = 0
        else:
            cur += 1
    runs.append(cur)
    assert sorted(runs)[-2] > 2500, sorted(runs)[-4:]  # block X and block Y are one basic block each


def test_resource_usage_of_the_hot_kernels():
    import ecloop_b200 as E

    out = subprocess.run(["cuobjdump", "-res-usage", str(E.library_path())], capture_output=True, text=True).stdout
    regs = dict(re.findall(r"Function (\S+):\s*\n\s*REG:(\d+)", out))
    for name, r in regs.items():
        if "add_kernel" in name:
            assert int(r) <= 128, (name, r)  # one CTA of 512 threads per SM needs <= 128 registers
        if "mul_points_kernel" in name:
            assert int(r) <= 128, (name, r)  # two CTAs of 256 threads per SM
    assert any("add_kernel_sp" in n for n in regs)


def test_launch_planner_invariants(tmp_path):
    """ecloop_b200/csrc/launch_plan.h on the CPU (tests/csrc/plan_test.cpp): every span size is covered, fits the resident
    threads, overhangs by less than one thread's share, and the shapes quoted in DESIGN.md K1 come out"""
    exe = tmp_path / "plan_test"
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-o", str(exe), str(ROOT / "tests" / "csrc" / "plan_test.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout
