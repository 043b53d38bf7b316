#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200-native DiskANN distance hot path.

Metric (BASELINE.json): QPS at recall@10 >= 0.95 on synthetic 1M x 128 f32 L2 (R=64, max degree
83, L_build=100, alpha=1.2, batch of 10K queries, beam 1), plus the achieved fraction of the HBM
roofline of the search kernel.  A "step" is one pass of the hot path over the whole query batch.

    python bench.py --gpus N --steps K --warmup W            # this repo (GPU)
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference path

value : device-timed QPS with the queries already resident in HBM (dab_search_batch_device)
e2e   : the same through the reference-facing C-ABI call with pinned HOST buffers
        (dab_search_batch: H2D of the queries and D2H of ids/distances inside the timed region)
Multi-GPU: one process per GPU (torchrun), index replicated (graph broadcast once over NCCL),
each rank searches its own 10K-query shard with no collective on the search path -> weak scaling.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_points, dim, n_queries, n_centers, R, L_build)
    "c2_1Mx128_f32_l2": dict(n=1_000_000, dim=128, nq=10_000, centers=1024, R=64, l_build=100),
    "small_100Kx128_f32_l2": dict(n=100_000, dim=128, nq=10_000, centers=256, R=64, l_build=100),
}
ALPHA = 1.2
K = 10
TARGET_RECALL = 0.95
L_SWEEP = [10, 15, 20, 25, 30, 40, 50, 60, 70, 80, 90, 100, 120, 140, 160, 200, 250, 300, 400]
SEED_BASE, SEED_QUERY = 0xD15C0003, 0xD15C0004


def max_degree(R):
    return int(R * 1.3)  # graph slack factor, diskann/src/graph/config/defaults.rs:26


def make_data(cfg, seed, count, centers):
    """Clustered Gaussians: centre ~ N(0, I), point = centre + 0.3 * N(0, I)."""
    rng = np.random.default_rng(seed)
    out = np.empty((count, cfg["dim"]), np.float32)
    step = 1 << 18
    for i in range(0, count, step):
        m = min(step, count - i)
        which = rng.integers(0, centers.shape[0], m)
        out[i:i + m] = centers[which] + np.float32(0.3) * rng.standard_normal((m, cfg["dim"]), dtype=np.float32)
    return out


def make_centers(cfg):
    return np.random.default_rng(SEED_BASE ^ 0xC0FFEE).standard_normal((cfg["centers"], cfg["dim"]), dtype=np.float32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def algorithmic_bytes(cmps, hops, dim, elem, k, max_deg):
    """per query cmps*(d*sizeof(T)+8) + hops*(max_degree+1)*4 + d*sizeof(T) + k*8."""
    unit = dim * elem + 8
    return float(cmps.astype(np.float64).sum() * unit + hops.astype(np.float64).sum() * (max_deg + 1) * 4
                 + len(cmps) * (dim * elem + k * 8))


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("search_kernel_dram_bytes_per_launch")
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------ GPU arm

def run_gpu(args):
    import torch
    import torch.distributed as dist

    import diskann_b200 as dab

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # whatever NCCL logs, stdout stays the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = WORKLOADS[args.workload]
    n, dim, nq, R = cfg["n"], cfg["dim"], cfg["nq"], cfg["R"]
    md = max_degree(R)

    t0 = time.time()
    centers = make_centers(cfg)
    base = make_data(cfg, SEED_BASE, n, centers)
    queries = make_data(cfg, SEED_QUERY + rank, nq, centers)
    medoid = base[np.argmin(((base - base.mean(0, dtype=np.float64).astype(np.float32)) ** 2).sum(1))]
    t_data = time.time() - t0

    g = dab.GpuIndex(dab.DType.f32, dab.Metric.L2, dim, n, 1, md, device=local)
    g.upload_vectors(base)
    g.upload_vectors(medoid[None, :], first=n)
    # a real (non-default) stream shared by torch and the library, so the CUDA events below are
    # recorded on the stream the kernels are launched on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    g.set_stream(stream.cuda_stream)

    # index: built once on rank 0 with the device build, replicated with one NCCL broadcast
    t0 = time.time()
    adj_dev = torch.empty((n + 1, md + 1), dtype=torch.int32, device="cuda")
    if rank == 0:
        g.build(R, cfg["l_build"], ALPHA)
        adj_host = g.download_graph()
        if world > 1:
            adj_dev.copy_(torch.from_numpy(adj_host.view(np.int32)))
    if world > 1:
        dist.broadcast(adj_dev, src=0)
        if rank != 0:
            torch.cuda.synchronize()
            g.upload_graph_device(adj_dev.data_ptr(), md + 1, n + 1)
    t_build = time.time() - t0
    del adj_dev

    # ground truth (exhaustive scan, bit-identical distances) and the L sweep on this rank's shard
    t0 = time.time()
    gt_ids, _ = g.flat_knn(queries, K)
    t_gt = time.time() - t0

    def recall_of(ids, counts):
        hits = 0
        for i in range(ids.shape[0]):
            hits += len(set(gt_ids[i].tolist()) & set(ids[i, :counts[i]].tolist()))
        return hits / (ids.shape[0] * K)

    # BASELINE.json configs[1] names L_search=100; the sweep records the smallest L that already
    # reaches the recall target (reported, and used instead only if L=100 itself misses it).
    sweep = []
    l_search = args.l_search or cfg.get("l_search", 100)
    min_l = None
    if rank == 0:
        for L in L_SWEEP:
            ids, _, counts, cmps, hops = g.search_batch(queries, K, L, 1)
            r = recall_of(ids, counts)
            sweep.append({"l": L, "recall": round(r, 5), "mean_cmps": float(cmps.mean()), "mean_hops": float(hops.mean())})
            if r >= TARGET_RECALL:
                min_l = L
                break
        if min_l is None:
            min_l = L_SWEEP[-1]
        if min_l > l_search and not args.l_search:
            l_search = min_l
    if world > 1:
        t = torch.tensor([l_search or 0], device="cuda")
        dist.broadcast(t, src=0)
        l_search = int(t.item())

    # resident inputs / outputs for `value`
    d_q = torch.from_numpy(queries).cuda()
    d_ids = torch.empty((nq, K), dtype=torch.int32, device="cuda")
    d_dists = torch.empty((nq, K), dtype=torch.float32, device="cuda")
    d_counts = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_cmps = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_hops = torch.empty(nq, dtype=torch.int32, device="cuda")

    def step_device():
        g.search_batch_device(d_q.data_ptr(), nq, K, l_search, 1, d_ids.data_ptr(), d_dists.data_ptr(),
                              d_counts.data_ptr(), d_cmps.data_ptr(), d_hops.data_ptr())

    # pinned host buffers for `e2e` (the C-ABI call a Rust caller makes)
    h_q = torch.from_numpy(queries).pin_memory()
    h_ids = torch.empty((nq, K), dtype=torch.int32).pin_memory()
    h_dists = torch.empty((nq, K), dtype=torch.float32).pin_memory()
    lib = dab.lib()

    def step_e2e():
        dab._lib.check(lib.dab_search_batch(g._h, C.c_void_p(h_q.data_ptr()), nq, K, l_search, 1, C.c_void_p(h_ids.data_ptr()),
                                            C.c_void_p(h_dists.data_ptr()), None, None, None))

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = dab.launch_count()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        timed.launches = dab.launch_count() - l0
        ms = e0.elapsed_time(e1)
        if world > 1:
            dist.barrier()
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if args.profile_range:  # ncu --profile-from-start off: only the timed region is captured
        torch.cuda.profiler.start()
    ms_dev = timed(step_device, args.steps, args.warmup)
    launches = timed.launches
    ms_e2e = timed(step_e2e, args.steps, args.warmup)
    if args.profile_range:
        torch.cuda.profiler.stop()
    clocks = sampler.stop() if rank == 0 else None

    # correctness of what was timed: the resident and host paths agree, recall at the chosen L
    ids_dev = d_ids.cpu().numpy().view(np.uint32)
    counts = d_counts.cpu().numpy().view(np.uint32)
    cmps = d_cmps.cpu().numpy().view(np.uint32)
    hops = d_hops.cpu().numpy().view(np.uint32)
    assert np.array_equal(ids_dev, h_ids.numpy().view(np.uint32)), "device-resident and host C-ABI results differ"
    recall = recall_of(ids_dev, counts)

    # informative only: the same device-resident step at the smallest L of the sweep that already
    # meets the recall target (rank 0, after every collective of the timed runs)
    at_min_l = None
    if rank == 0 and min_l and min_l != l_search:
        def step_min():
            g.search_batch_device(d_q.data_ptr(), nq, K, min_l, 1, d_ids.data_ptr(), d_dists.data_ptr(),
                                  d_counts.data_ptr(), d_cmps.data_ptr(), d_hops.data_ptr())
        for _ in range(max(3, args.warmup)):
            step_min()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_min()
        e1.record(stream)
        torch.cuda.synchronize()
        ms_min = e0.elapsed_time(e1) / args.steps
        r_min = recall_of(d_ids.cpu().numpy().view(np.uint32), d_counts.cpu().numpy().view(np.uint32))
        at_min_l = {"l_search": min_l, "recall_at_10": round(r_min, 5), "ms_per_step": ms_min,
                    "queries_per_s_this_gpu": nq / (ms_min / 1e3)}

    ms_step = ms_dev / args.steps
    total_q = nq * world
    value = total_q / (ms_step / 1e3)
    e2e_value = total_q / ((ms_e2e / args.steps) / 1e3)
    alg_bytes = algorithmic_bytes(cmps, hops, dim, 4, K, md)
    peak, peak_src = measured_peak_gbs()
    achieved = alg_bytes / (ms_step / 1e3) / 1e9

    result = None
    if rank == 0:
        result = {
            "metric": "QPS @ recall@10>=0.95, 1Mx128 f32 L2 (Vamana R=64 greedy search, batch 10K)",
            "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "n_points": n, "dim": dim, "queries_per_gpu": nq, "metric": "L2",
                       "pruned_degree": R, "max_degree": md, "l_build": cfg["l_build"], "alpha": ALPHA, "k": K,
                       "l_search": l_search, "beam_width": 1, "recall_at_10": round(recall, 5),
                       "mean_cmps": float(cmps.mean()), "mean_hops": float(hops.mean()),
                       "generator": f"{cfg['centers']} Gaussian centres N(0,I), point = centre + 0.3 N(0,I); "
                                    f"seeds base {SEED_BASE:#x} query {SEED_QUERY:#x}+rank; start = copy of the medoid",
                       "index": "replicated per GPU, built on rank 0 by dab_build (device), one NCCL broadcast",
                       "parallelism": f"replica x{world}, queries sharded, no collective on the search path",
                       "l2_policy": f"no flush: index {(n * dim * 4 + (n + 1) * 4 * (md + 1)) / 1e6:.0f} MB >> 126 MB L2 and "
                                    "each step gathers ~GBs of random rows",
                       "setup_s": {"data": round(t_data, 1), "build": round(t_build, 1), "ground_truth": round(t_gt, 2)},
                       "min_l_for_target_recall": min_l, "l_sweep": sweep, "at_min_l": at_min_l},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": nq * dim * 4,
                    "d2h_bytes_per_step": nq * K * 8, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(), "kernel": "search_kernel_v2<float,L2,QT=4>",
                         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                         "note": "achieved = algorithmic bytes (cmps*520 + hops*336 + 512 + k*8 per query, run's own counters) "
                                 "/ CUDA-event step time on this rank"},
        }
        if not args.no_cpu_baseline and world == 1:
            result["cpu_baseline"] = cpu_baseline(base, medoid, g.download_graph(), queries, n, l_search, gt_ids)
    g.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(result)


# ------------------------------------------------------------------------------------------ CPU arms

def cpu_search_setup(base, medoid, adj, n):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O  # the CPU checker; only used for the CPU baseline legs
    vecs = np.concatenate([base, medoid[None, :]])
    return O, O.Index(vecs, adj, n, 1, O.L2)


def cpu_baseline(base, medoid, adj, queries, n, l_search, gt_ids, reps=3):
    """The CPU restatement of the reference path (AVX2, reference threading model:
    contiguous query partitions, one thread each) on this box's host cores."""
    O, oidx = cpu_search_setup(base, medoid, adj, n)
    threads = O.lib().orc_hardware_threads()
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        ids, _, counts, _, _ = oidx.search_batch(queries, K, l_search, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    t0 = time.perf_counter()
    sample = queries[:500]
    oidx.search_batch(sample, K, l_search, threads=1)
    dt1 = time.perf_counter() - t0
    return {"value": queries.shape[0] / best, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"all {queries.shape[0]} queries x {reps} reps (best), same graph/L as the GPU arm, AVX2 V3-order kernels",
            "recall_at_10": round(O.recall(gt_ids, ids, counts, K, K), 5),
            "single_thread_qps": 500 / dt1}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle restatement; the Rust workspace cannot
    be compiled here) on the host cores.  The graph is input data: it is produced once by the
    device build (untimed) because the sequential CPU build of 1M points would take hours."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    cfg = WORKLOADS[args.workload]
    n, dim, nq, R = cfg["n"], cfg["dim"], cfg["nq"], cfg["R"]
    md = max_degree(R)
    centers = make_centers(cfg)
    base = make_data(cfg, SEED_BASE, n, centers)
    queries = make_data(cfg, SEED_QUERY, nq, centers)
    medoid = base[np.argmin(((base - base.mean(0, dtype=np.float64).astype(np.float32)) ** 2).sum(1))]
    if not has_gpu:
        emit({"impl": "reference", "unavailable": "no GPU to prepare the 1M-point graph input for the CPU arm"})
        return
    import diskann_b200 as dab
    g = dab.GpuIndex(dab.DType.f32, dab.Metric.L2, dim, n, 1, md)
    g.upload_vectors(base)
    g.upload_vectors(medoid[None, :], first=n)
    g.build(R, cfg["l_build"], ALPHA)
    adj = g.download_graph()
    gt_ids, _ = g.flat_knn(queries, K)
    # the same L as the GPU arm: BASELINE.json configs[1] names L_search=100; a larger L only if
    # that misses the recall target (the search is deterministic, so the device sweep decides)
    l_search = args.l_search or cfg.get("l_search", 100)
    min_l = None
    for L in L_SWEEP:
        ids, _, counts, _, _ = g.search_batch(queries, K, L, 1)
        hits = sum(len(set(gt_ids[i].tolist()) & set(ids[i, :counts[i]].tolist())) for i in range(nq))
        min_l = L
        if hits / (nq * K) >= TARGET_RECALL:
            break
    if min_l > l_search and not args.l_search:
        l_search = min_l
    g.close()
    O, oidx = cpu_search_setup(base, medoid, adj, n)
    threads = O.lib().orc_hardware_threads()
    for _ in range(args.warmup):
        oidx.search_batch(queries, K, l_search, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ids, _, counts, cmps, hops = oidx.search_batch(queries, K, l_search, threads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    qps = nq / dt
    emit({
        "impl": "reference", "metric": "QPS @ recall@10>=0.95, 1Mx128 f32 L2 (Vamana R=64 greedy search, batch 10K)",
        "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "n_points": n, "dim": dim, "queries": nq, "l_search": l_search, "k": K,
                   "recall_at_10": round(O.recall(gt_ids, ids, counts, K, K), 5), "mean_cmps": float(cmps.mean()),
                   "mean_hops": float(hops.mean()), "min_l_for_target_recall": min_l},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": f"each step = the full {nq}-query batch on {threads} threads (contiguous partitions)"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries loaded later (NCCL prints its version
    banner there) write to file descriptor 1 directly, so fd 1 is pointed at stderr for the rest
    of the process and the JSON line goes to a private duplicate of the original stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(obj):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2_1Mx128_f32_l2", choices=sorted(WORKLOADS))
    ap.add_argument("--l-search", type=int, default=0, help="skip the sweep and use this L")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-range", action="store_true", help="cudaProfilerStart/Stop around the timed region (for ncu)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
