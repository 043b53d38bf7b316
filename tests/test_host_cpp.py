"""The C++ host mirror (include/diskann_b200.hpp) over the C ABI: compiles, fails loudly without a
GPU, and on a GPU reproduces the reference's checked-in grid-search baselines."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "grid_search")


def build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples"), "-s"])


def test_cpp_host_mirror_compiles_and_reports_anns_error_without_gpu():
    import torch
    build()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([EXE, "3", "5", "1", "-1"], capture_output=True, text=True)
    assert r.returncode == 16 and "no CPU fallback" in r.stderr  # ANNError code 6 (DAB_ERR_NO_DEVICE)


@pytest.mark.gpu
def test_cpp_host_mirror_reproduces_grid_baselines():
    build()
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "grid_search.json")))
    for case in g["cases"]:
        qv = case["query"][0]
        r = subprocess.run([EXE, str(case["grid_dims"]), str(case["grid_size"]), str(case["beam_width"]), str(qv)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        res, stats = r.stdout.split("|")
        pairs = [p.split(":") for p in res.split()]
        want = case["results"][:case["num_results"]]
        assert [int(a) for a, _ in pairs] == [w[0] for w in want], case
        assert [float(b) for _, b in pairs] == [w[1] for w in want], case
        assert f"cmps={case['comparisons']} hops={case['hops']} count={case['num_results']}" in stats
