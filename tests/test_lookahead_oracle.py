"""The CPU lookahead checker against its own frozen fingerprints (tests/golden/lookahead_golden.json).
Not a reference pin -- libx264 is absent from the reference tree -- but it keeps the target of the
GPU parity tests from moving unnoticed."""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_lookahead_golden as mk  # noqa: E402

GOLDEN = json.load(open(os.path.join(HERE, "golden", "lookahead_golden.json")))


@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
def test_checker_reproduces_its_golden_fingerprints(case):
    fp = mk.fingerprint(case["preset"], case["w"], case["h"], case["frames"], case["over"])
    for k in ("types", "coded_order", "costs", "qp_offset_fnv", "qp_offset_aq_fnv"):
        assert fp[k] == case[k], (case["name"], k)
