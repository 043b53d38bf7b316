"""bench.py contract pieces that can be checked without a GPU."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_follows_the_survey_formula():
    # cmps*(d*sizeof(T)+8) + hops*(max_degree+1)*4 + d*sizeof(T) + k*8 per query
    cmps = np.array([1000, 2000], np.uint32)
    hops = np.array([100, 110], np.uint32)
    got = bench.algorithmic_bytes(cmps, hops, 128, 4, 10, 83)
    want = 3000 * 520 + 210 * 336 + 2 * (512 + 80)
    assert got == want


def test_max_degree_is_the_reference_slack():
    assert bench.max_degree(64) == 83 and bench.max_degree(32) == 41  # config/mod.rs:269-275


def test_stdout_carries_exactly_one_json_line():
    """Anything a library writes to file descriptor 1 after start-up must not reach stdout."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "os.write(1, b'library banner\\n'); print('python noise'); bench.emit({'ok': 1})" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0
    assert p.stdout.strip().splitlines() == ['{"ok": 1}']
    assert "library banner" in p.stderr and "python noise" in p.stderr


def test_reference_arm_prints_one_json_line_without_a_gpu():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--workload", "small_100Kx128_f32_l2"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert p.returncode == 0
    lines = p.stdout.strip().splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
