"""DiskANN on-disk formats (diskann_b200/formats.py): byte layouts as the reference writes them
(diskann-utils/src/io.rs, storage/pq_storage.rs, storage/bin.rs) and round trips."""
import struct

import numpy as np
import pytest

from diskann_b200 import formats as F


def test_bin_layout_and_round_trip(tmp_path):
    m = np.arange(12, dtype=np.float32).reshape(3, 4)
    p = tmp_path / "a.bin"
    assert F.write_bin(p, m) == 8 + 48
    raw = p.read_bytes()
    assert struct.unpack("<II", raw[:8]) == (3, 4) and raw[8:] == m.tobytes()  # io.rs:24-80
    assert np.array_equal(F.read_bin(p, np.float32), m)
    p.write_bytes(raw[:-4])
    with pytest.raises(ValueError):
        F.read_bin(p, np.float32)


def test_pq_pivot_file_layout(tmp_path):
    rng = np.random.default_rng(0)
    piv = rng.normal(size=(256, 10)).astype(np.float32)
    offs = np.array([0, 4, 7, 10], np.uint64)
    p = tmp_path / "pq_pivots.bin"
    F.write_pq_pivots(p, piv, offs)
    raw = p.read_bytes()
    # offset table: a .bin column of four u64 at byte 0; data starts after the 4 KiB metadata block
    assert struct.unpack("<II", raw[:8]) == (4, 1)
    o = struct.unpack("<4Q", raw[8:40])
    assert o[0] == 4096 and o[1] == 4096 + 8 + piv.nbytes and o[2] == o[1] + 8 + 40 and o[3] == o[2] + 8 + 16
    assert len(raw) == o[3]
    got, centroid, got_offs = F.read_pq_pivots(p)
    assert np.array_equal(got, piv) and not centroid.any() and np.array_equal(got_offs, offs)


def test_graph_file_layout_and_round_trip(tmp_path):
    adj = np.zeros((4, 5), np.uint32)
    adj[0, :3] = [2, 1, 3]
    adj[1, :2] = [1, 0]
    adj[3, :5] = [4, 0, 1, 2, 9]
    p = tmp_path / "graph"
    size = F.write_graph(p, adj, start_point=3, max_degree=4)
    raw = p.read_bytes()
    assert size == len(raw) == 24 + 4 * (4 + 7)
    assert struct.unpack("<QIIQ", raw[:24]) == (size, 4, 3, 1)  # bin.rs:343-350
    assert struct.unpack("<3I", raw[24:36]) == (2, 1, 3)
    got, md, start, extra = F.read_graph(p)
    assert (md, start, extra) == (4, 3, 1) and np.array_equal(got, adj)


def test_groundtruth_round_trip(tmp_path):
    ids = np.arange(6, dtype=np.uint32).reshape(2, 3)
    d = np.linspace(0, 1, 6, dtype=np.float32).reshape(2, 3)
    F.write_groundtruth(tmp_path / "gt", ids, d)
    a, b = F.read_groundtruth(tmp_path / "gt")
    assert np.array_equal(a, ids) and np.array_equal(b, d)
