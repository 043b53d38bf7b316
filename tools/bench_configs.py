#!/usr/bin/env python3
"""Search throughput on the other BASELINE.json shapes (not the contract bench; bench.py stays on
configs[1]).  One GPU, device-resident timing with CUDA events, recall against the exhaustive scan.

    python tools/bench_configs.py c3        # 1M x 768 f16, inner product (unit-normalised rows)
    python tools/bench_configs.py c4fp      # 1M x 128 i8, L2, full-precision traversal
    python tools/bench_configs.py c3 --n 200000 --steps 10

WRITTEN WITHOUT A GPU AT HAND (end of round 1): expect to fix small things on first use.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (generator, constants, roofline helpers)

SHAPES = {
    "c3": dict(dim=768, dtype="f16", metric="InnerProduct", centers=1024, seeds=(0xD15C0005, 0xD15C0006)),
    "c4fp": dict(dim=128, dtype="i8", metric="L2", centers=1024, seeds=(0xD15C0007, 0xD15C0008)),
}


def make(shape, seed, count, centers):
    x = bench.make_data({"dim": shape["dim"]}, seed, count, centers)
    if shape["dtype"] == "f16":
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        return x.astype(np.float16)
    return np.clip(np.round(x * 40.0), -127, 127).astype(np.int8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("shape", choices=sorted(SHAPES))
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--l-search", type=int, default=100)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch

    import diskann_b200 as dab

    sh = SHAPES[args.shape]
    dim, n, nq, R = sh["dim"], args.n, args.nq, 64
    md = bench.max_degree(R)
    centers = np.random.default_rng(sh["seeds"][0] ^ 0xC0FFEE).standard_normal((sh["centers"], dim), dtype=np.float32)
    base = make(sh, sh["seeds"][0], n, centers)
    queries = make(sh, sh["seeds"][1], nq, centers)
    mean = base.astype(np.float32).mean(0)
    medoid = base[np.argmin(((base.astype(np.float32) - mean) ** 2).sum(1))]
    ddt = dab.DType.f16 if sh["dtype"] == "f16" else dab.DType.i8
    metric = getattr(dab.Metric, sh["metric"])
    elem = 2 if sh["dtype"] == "f16" else 1

    g = dab.GpuIndex(ddt, metric, dim, n, 1, md)
    g.upload_vectors(base)
    g.upload_vectors(medoid[None, :], first=n)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    g.set_stream(stream.cuda_stream)
    t0 = time.time()
    g.build(R, 100, bench.ALPHA)
    t_build = time.time() - t0
    t0 = time.time()
    gt_ids, _ = g.flat_knn(queries, bench.K)
    t_gt = time.time() - t0

    def recall_of(ids, counts):
        return sum(len(set(gt_ids[i].tolist()) & set(ids[i, :counts[i]].tolist())) for i in range(nq)) / (nq * bench.K)

    sweep, min_l = [], None
    for L in bench.L_SWEEP:
        ids, _, counts, cmps, hops = g.search_batch(queries, bench.K, L, 1)
        r = recall_of(ids, counts)
        sweep.append({"l": L, "recall": round(r, 5), "mean_cmps": float(cmps.mean())})
        if r >= bench.TARGET_RECALL:
            min_l = L
            break
    L = max(args.l_search, min_l or bench.L_SWEEP[-1])

    tq = torch.from_numpy(queries.view(np.int16) if sh["dtype"] == "f16" else queries).cuda()
    d_ids = torch.empty((nq, bench.K), dtype=torch.int32, device="cuda")
    d_dists = torch.empty((nq, bench.K), dtype=torch.float32, device="cuda")
    d_counts = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_cmps = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_hops = torch.empty(nq, dtype=torch.int32, device="cuda")

    def step():
        g.search_batch_device(tq.data_ptr(), nq, bench.K, L, 1, d_ids.data_ptr(), d_dists.data_ptr(), d_counts.data_ptr(),
                              d_cmps.data_ptr(), d_hops.data_ptr())

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    cmps = d_cmps.cpu().numpy().view(np.uint32)
    hops = d_hops.cpu().numpy().view(np.uint32)
    rec = recall_of(d_ids.cpu().numpy().view(np.uint32), d_counts.cpu().numpy().view(np.uint32))
    alg = bench.algorithmic_bytes(cmps, hops, dim, elem, bench.K, md)
    peak, peak_src = bench.measured_peak_gbs()
    print(json.dumps({
        "shape": args.shape, "n_points": n, "dim": dim, "dtype": sh["dtype"], "metric": sh["metric"], "queries": nq, "l_search": L,
        "min_l_for_target_recall": min_l, "recall_at_10": round(rec, 5), "ms_per_step": ms, "queries_per_s": nq / (ms / 1e3),
        "mean_cmps": float(cmps.mean()), "mean_hops": float(hops.mean()), "algorithmic_bytes_per_step": alg,
        "roofline_frac": alg / (ms / 1e3) / 1e9 / peak, "peak_source": peak_src, "setup_s": {"build": round(t_build, 1), "ground_truth": round(t_gt, 2)},
        "l_sweep": sweep}))


if __name__ == "__main__":
    main()
