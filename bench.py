#!/usr/bin/env python
"""bench.py — PCM pair-checks/s (+ max-clique ms) at 50k loop closures, the metric BASELINE.json names.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
  python bench.py --impl reference ...                     (the CPU restatement on the host cores)

A "step" is one pass of the PCM hot path over one batch of synthetic loop closures:
  pairwise consistency matrix (K3) -> [all-gather of row chunks, N>1] -> mirror + degrees -> max clique (K4).
`value`   : pair-checks/s with all inputs already resident in HBM (device-timed with CUDA events).
`e2e`     : the same metric through the public host API (PcmGpu: odometry fold + closure append + inlier
            selection) from HOST buffers, host<->device copies inside the (wall-clock) timed region.
`roofline`: the pairwise kernel alone, algorithmic fp64 flop (6.2e3 per pair) over its
            CUDA-event duration, against the FP64 DFMA peak measured live by rpgo_fp64_peak().
`cpu_baseline`: the CPU oracle (a restatement of the reference, 1 thread) on a bounded sample.
Prints ONE JSON line on rank 0.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PAIR_3D_PCM = 6.2e3  # SURVEY.md §8(d) algorithmic flop per pair (dense count of the reference's arithmetic)
METRIC = "pcm_pair_checks_per_sec"
UNIT = "pair-checks/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--closures", type=int, default=50000)
    ap.add_argument("--poses", type=int, default=50000)
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--lc-threshold", type=float, default=5.0)
    ap.add_argument("--cpu-sample", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return ("config5-50k: 3D single-robot helix, P=%d poses, n=%d closures in one group, 50%% outliers, seed 4, "
            "Pcm3D(odom=-1, lc=%.1f)" % (a.poses, a.closures, a.lc_threshold))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                               "--format=csv,noheader,nounits"], text=True, timeout=5)
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_baseline(a, gph, sample):
    """The oracle (kind 'port', single thread like the reference) on the first `sample` closures."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    lcs = gph["lcs"][:sample]
    res = {}
    for shaped in (False, True):
        o = orc.OraclePcm(3, 0, odom_threshold=-1, lc_threshold=a.lc_threshold)
        o.set_reference_shaped(shaped)
        o.update(gph["odom"], gph["values"])
        t0 = time.perf_counter()
        o.update(lcs, [])
        dt = time.perf_counter() - t0
        res[shaped] = (o.pair_checks() / dt, dt, o.num_inliers())
    return {"value": res[False][0], "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "first %d closures of the same workload (%d pairs, %.1f s): oracle PCM incl. clique, packed "
                      "adjacency; reference-shaped (dense double matrices re-copied per closure) = %.3e %s"
                      % (sample, sample * (sample - 1) // 2, res[False][1], res[True][0], UNIT)}


def _ref_worker(args):
    seed, poses, n, thr = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    synth = importlib.import_module("kimera-rpgo_b200.synth")
    gph = synth.config2(seed=seed, P=poses, n=n)
    o = orc.OraclePcm(3, 0, odom_threshold=-1, lc_threshold=thr)
    o.set_reference_shaped(False)
    o.update(gph["odom"], gph["values"])
    t0 = time.perf_counter()
    o.update(gph["lcs"], [])
    return o.pair_checks(), time.perf_counter() - t0


def run_reference(a):
    """--impl reference: the reference's CPU path (oracle port; the reference itself needs GTSAM and cannot be
    built) on all host cores: one independent sample of the workload per core per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n = 600
    poses = 3000
    pool = mp.Pool(cores)
    jobs = [(100 + i, poses, n, a.lc_threshold) for i in range(cores)]
    for _ in range(max(a.warmup, 1)):
        pool.map(_ref_worker, jobs[:cores])
    t0 = time.perf_counter()
    pairs = 0
    for _ in range(a.steps):
        for p, _dt in pool.map(_ref_worker, jobs):
            pairs += p
    dt = time.perf_counter() - t0
    pool.close()
    val = pairs / dt
    sample = ("%d processes x (3D helix, P=%d, n=%d closures, 50%% outliers) per step; oracle port of Pcm.h "
              "(single-threaded per graph, like the reference)" % (cores, poses, n))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "reference_sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("kimera-rpgo_b200")
    synth = importlib.import_module("kimera-rpgo_b200.synth")
    par = importlib.import_module("kimera-rpgo_b200.parallel")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=device)
    assert world == a.gpus or world == 1, "launch with torch.distributed.run for --gpus > 1"

    gph = synth.config2(seed=4, P=a.poses, n=a.closures)
    arr = synth.as_arrays(gph)
    n = a.closures
    pairs = n * (n - 1) // 2
    params = dict(odom_threshold=-1.0, lc_threshold=a.lc_threshold)

    # ---- resident state for the kernel-timed leg -------------------------------------------------
    g = pkg.PcmGpu(3, 0, device=local, kernel=a.kernel, rank=rank, world=world, **params)
    g.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
    g.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])  # all-gathers when world > 1
    st = torch.cuda.ExternalStream(g.stream_ptr(), device=device)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2

    def one_step(timed):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        flush.fill_(1)  # L2 flush between iterations (outside the timed events)
        torch.cuda.synchronize()
        with torch.cuda.stream(st):
            ev[0].record(st)
            g.pairwise_only(0, 0)
            ev[1].record(st)
            if world > 1:
                par.allgather_adjacency(g, 0, device)
            else:
                g.finalize(0)
            ev[2].record(st)
            size, ids, _ = g.find_inliers_raw(0, pkg.CLIQUE_HEU)
            ev[3].record(st)
        ev[3].synchronize()
        return ev[0].elapsed_time(ev[3]), ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3]), size

    for _ in range(a.warmup):
        one_step(False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = g.launch_count()
    tot = k3 = cl = 0.0
    size = 0
    for _ in range(a.steps):
        t, tk, tc, size = one_step(True)
        tot += t
        k3 += tk
        cl += tc
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = g.launch_count() - l0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    tt = torch.tensor([tot, k3, cl], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    tot, k3, cl = [float(x) for x in tt.tolist()]
    ms_per_step = tot / a.steps
    value = pairs / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    import ctypes as C
    tf = C.c_double()
    g.lib.rpgo_fp64_peak(local, C.byref(tf))
    k3_ms = k3 / a.steps
    achieved = FLOP_PER_PAIR_3D_PCM * (pairs / world) / (k3_ms * 1e-3) / 1e12
    roofline = {"bound": "fp64", "achieved": achieved, "peak": tf.value, "unit": "TFLOP/s", "frac": achieved / tf.value,
                "traffic": None, "kernel": ("pairwise_direct_kernel<3>" if a.kernel in (1, 13, 14) else
                           "pairwise_tiled_kernel<3>" if a.kernel in (2, 20, 21, 22, 23, 24) else "pairwise_grouped_kernel<3,12,3,504>"),
                "peak_source": "measured live by rpgo_fp64_peak (DFMA micro-benchmark; MEASURED_PEAKS.json has no FP64 entry)",
                "algorithmic_flop_per_pair": FLOP_PER_PAIR_3D_PCM, "kernel_ms": k3_ms}
    # DRAM traffic of that kernel from the committed ncu --set full capture of this very configuration (one GPU)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_k3_traffic.json")))
        if tr["kernel"] == roofline["kernel"] and a.closures == 50000 and world == 1:
            roofline["traffic"] = tr["traffic_bytes_per_launch"]
            roofline["traffic_source"] = tr["source"]
    except (OSError, KeyError, ValueError):
        pass

    # ---- e2e through the public host API, host buffers ---------------------------------------------
    e2e = None
    if not a.no_e2e:
        times = []
        h2d = d2h = 0
        for it in range(5 + 1):  # one untimed warm-up, five timed; the median is reported, all five are listed
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            p = pkg.PcmGpu(3, 0, device=local, kernel=a.kernel, rank=rank, world=world, **params)
            h2d = p.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
            p.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])  # incl. all-gather when world > 1
            h2d += p.last_h2d_bytes
            sz, ids, _ = p.find_inliers_raw(0, pkg.CLIQUE_HEU)
            d2h = p.last_d2h_bytes + 4 * int(sz) + 8
            p.sync()
            dt = time.perf_counter() - t0
            p.close()
            if it > 0:
                times.append(dt)
        te = torch.tensor([sorted(times)[len(times) // 2]], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": pairs / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": float(te.item()) * 1e3, "ms_all_rank0": [round(t * 1e3, 1) for t in times],
               "what": "new handle -> odom_append(P-1 factors) -> lc_append(n closures) -> find_inliers, numpy host buffers"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "pairs_per_step": pairs, "l2": "flushed between iterations (256 MB fill)",
                   "parallelism": "rows sharded over %d GPU(s), NCCL all-gather of the adjacency bitset" % world},
        "k3_ms": k3_ms, "max_clique_ms": cl / a.steps, "max_clique_size": int(size),
        "roofline": roofline, "gpu_launches": int(launches), "clocks": sampler.summary(),
    }
    if e2e:
        out["e2e"] = e2e
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(a, gph, a.cpu_sample)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
