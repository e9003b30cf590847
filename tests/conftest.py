import gzip
import json
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    meta = json.loads((GOLD / "golden.json").read_text())
    return {m["name"]: m for m in meta["fixtures"]}


def golden_lines(name):
    with gzip.open(GOLD / (name + ".txt.gz"), "rt") as f:
        return f.read().splitlines()


@pytest.fixture(scope="session")
def puzzles_filter():
    import oracle as O

    return O.filter_from_text_file(GOLD / "btc-puzzles-hash")
