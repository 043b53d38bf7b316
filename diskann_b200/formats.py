"""DiskANN on-disk formats (host side), so that an index built by the reference can be put on the
GPU and a device-built index handed back:

* `.bin` matrices — `u32 npoints, u32 ndims`, then row-major data (diskann-utils/src/io.rs:24-80);
* PQ pivot files — 4 KiB metadata block holding a `.bin` column of four u64 byte offsets, then the
  pivot table `[n_centers][dim] f32`, the centroid `[dim][1] f32` and the chunk offsets
  `[n_chunks + 1][1] u32`, each as a `.bin` (diskann-providers/src/storage/pq_storage.rs:72-146, 225-281);
  compressed vectors are a plain `.bin` of u8 `[npoints][n_chunks]`;
* saved graphs — `u64 index_size, u32 max_degree, u32 start_point, u64 additional_points`, then per node
  `u32 degree` + `degree` u32 ids (diskann-providers/src/storage/bin.rs:234-380);
* ground truth — `.bin` header `(n, k)`, `u32 ids[n][k]`, optionally `f32 dists[n][k]`.
"""
import struct

import numpy as np

METADATA_SIZE = 4096  # pq_storage.rs


def read_bin(path, dtype, offset=0):
    dtype = np.dtype(dtype)
    with open(path, "rb") as f:
        f.seek(offset)
        hdr = f.read(8)
        if len(hdr) != 8:
            raise ValueError(f"{path}: truncated .bin header")
        npts, dim = struct.unpack("<II", hdr)
        data = np.fromfile(f, dtype=dtype, count=npts * dim)
    if data.size != npts * dim:
        raise ValueError(f"{path}: expected {npts} x {dim} {dtype} values, found {data.size}")
    return data.reshape(npts, dim)


def write_bin(path, matrix, offset=None, mode="wb"):
    """Returns the number of bytes written (header included)."""
    matrix = np.ascontiguousarray(matrix)
    if matrix.ndim != 2:
        raise ValueError("write_bin expects a 2-d array")
    with open(path, mode) as f:
        if offset is not None:
            f.seek(offset)
        f.write(struct.pack("<II", matrix.shape[0], matrix.shape[1]))
        f.write(matrix.tobytes())
    return 8 + matrix.nbytes


def write_pq_pivots(path, pivots, chunk_offsets, centroid=None):
    pivots = np.ascontiguousarray(pivots, np.float32)
    n_centers, dim = pivots.shape
    centroid = np.zeros(dim, np.float32) if centroid is None else np.ascontiguousarray(centroid, np.float32)
    offs = np.ascontiguousarray(chunk_offsets).astype(np.uint32)
    cumul = [METADATA_SIZE, 0, 0, 0]
    with open(path, "wb") as f:
        f.write(b"\0" * METADATA_SIZE)
    cumul[1] = cumul[0] + write_bin(path, pivots, offset=cumul[0], mode="r+b")
    cumul[2] = cumul[1] + write_bin(path, centroid.reshape(dim, 1), offset=cumul[1], mode="r+b")
    cumul[3] = cumul[2] + write_bin(path, offs.reshape(-1, 1), offset=cumul[2], mode="r+b")
    write_bin(path, np.array(cumul, np.uint64).reshape(4, 1), offset=0, mode="r+b")


def read_pq_pivots(path):
    """-> (pivots [n_centers, dim] f32, centroid [dim] f32, chunk_offsets u64 [n_chunks + 1])."""
    table = read_bin(path, np.uint64)
    if table.shape != (4, 1):
        raise ValueError(f"{path}: offsets don't contain correct metadata, expected 4 x 1, found {table.shape}")
    o = table[:, 0].astype(np.int64)
    pivots = read_bin(path, np.float32, int(o[0]))
    centroid = read_bin(path, np.float32, int(o[1]))
    offs = read_bin(path, np.uint32, int(o[2]))
    if centroid.shape != (pivots.shape[1], 1) or offs.shape[1] != 1:
        raise ValueError(f"{path}: centroid / chunk offsets have the wrong shape")
    offs = offs[:, 0].astype(np.uint64)
    if offs[0] != 0 or offs[-1] != pivots.shape[1] or (np.diff(offs.astype(np.int64)) <= 0).any():
        raise ValueError(f"{path}: chunk offsets must start at 0, end at dim and increase")
    return pivots, centroid[:, 0].copy(), offs


def write_graph(path, adj, start_point, max_degree=None, additional_points=1):
    """adj: [n][stride] u32 rows `[degree, ids...]` (the layout of dab_download_graph)."""
    adj = np.ascontiguousarray(adj, np.uint32)
    deg = adj[:, 0]
    size = 24 + 4 * int(deg.size + deg.astype(np.int64).sum())
    md = int(deg.max()) if max_degree is None else int(max_degree)
    with open(path, "wb") as f:
        f.write(struct.pack("<QIIQ", size, md, int(start_point), int(additional_points)))
        for row in adj:
            f.write(row[:1 + int(row[0])].tobytes())
    return size


def read_graph(path, stride=None):
    """-> (adj [n][stride] u32 rows `[degree, ids...]`, max_degree, start_point, additional_points)."""
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size < 24:
        raise ValueError(f"{path}: truncated graph header")
    size, md, start, extra = struct.unpack("<QIIQ", raw[:24].tobytes())
    if size != raw.size:
        raise ValueError(f"{path}: header says {size} bytes, file has {raw.size}")
    words = raw[24:].view(np.uint32)
    rows, pos = [], 0
    while pos < words.size:
        d = int(words[pos])
        if pos + 1 + d > words.size:
            raise ValueError(f"{path}: adjacency list runs past the end of the file")
        rows.append(words[pos + 1:pos + 1 + d])
        pos += 1 + d
    width = max([md] + [len(r) for r in rows]) + 1 if stride is None else stride
    adj = np.zeros((len(rows), width), np.uint32)
    for i, r in enumerate(rows):
        if len(r) + 1 > width:
            raise ValueError(f"{path}: node {i} has {len(r)} neighbours, stride {width}")
        adj[i, 0] = len(r)
        adj[i, 1:1 + len(r)] = r
    return adj, md, start, extra


def read_groundtruth(path):
    with open(path, "rb") as f:
        n, k = struct.unpack("<II", f.read(8))
        ids = np.fromfile(f, np.uint32, n * k).reshape(n, k)
        rest = np.fromfile(f, np.float32)
    dists = rest.reshape(n, k) if rest.size == n * k else None
    return ids, dists


def write_groundtruth(path, ids, dists=None):
    ids = np.ascontiguousarray(ids, np.uint32)
    with open(path, "wb") as f:
        f.write(struct.pack("<II", ids.shape[0], ids.shape[1]))
        f.write(ids.tobytes())
        if dists is not None:
            f.write(np.ascontiguousarray(dists, np.float32).tobytes())
