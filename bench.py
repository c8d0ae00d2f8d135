#!/usr/bin/env python
"""bench.py — PCM pair-checks/s (+ max-clique ms) at 50k loop closures, the metric BASELINE.json names.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
  python bench.py --impl reference ...                     (the CPU restatement on the host cores)

A "step" is one pass of the PCM hot path over one batch of synthetic loop closures:
  pairwise consistency matrix (K3, this rank's row chunks) -> [NCCL all-gather of the row chunks, N>1]
  -> mirror + degrees -> max clique (K4, candidates partitioned over ranks when N>1).
`value`   : pair-checks/s with all inputs already resident in HBM (device-timed with CUDA events on the handle's stream).
`e2e`     : the same metric through the public host API (PcmGpu: reset -> odometry fold -> closure append -> inlier
            selection) from HOST buffers, host<->device copies inside the (wall-clock) timed region; the handle, its
            stream, arena and communicator live as long as a RobustSolver would.
`roofline`: the pairwise kernel alone, algorithmic fp64 flop (6.2e3 per pair) over its
            CUDA-event duration, against the FP64 DFMA peak measured live by rpgo_fp64_peak() (`frac`) and against
            the nominal 148 SM x 64 FMA x 2 x 1.965 GHz = 37.2 TFLOP/s (`frac_nominal`).
`cpu_baseline`: the CPU oracle (a restatement of the reference, 1 thread) on a bounded sample.
`extra`   : the other named configurations at the run's N (config 4: 8 robots / 36 groups; config 5: 200k closures),
            per-mode kernel rooflines (PcmSimple3D, Pcm2D) and the clique stage on a planted-clique instance —
            measured outside the headline timing.
Prints ONE JSON line on rank 0.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PAIR = {"pcm3d": 6.2e3, "pcm2d": 9.6e2, "simple3d": 7.4e2}  # SURVEY.md §8(d), algorithmic flop per pair
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12
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
    ap.add_argument("--cpu-sizes", default="1000,2000", help="closure counts of the single-thread CPU baseline samples")
    ap.add_argument("--ref-closures", type=int, default=1000, help="closures per graph in the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--extras", default="config4,config5,modes,planted")
    ap.add_argument("--config5-closures", type=int, default=200000)
    return ap.parse_args()


def workload_name(a):
    return ("config5-50k: 3D single-robot helix, P=%d poses, n=%d closures in one group, 50%% outliers, seed 4, "
            "Pcm3D(odom=-1, lc=%.1f)" % (a.poses, a.closures, a.lc_threshold))


class ClockLog:
    """clocks + throttle reasons DURING the timed region: one `nvidia-smi -lms 200` child started before and stopped
    after it (B200_PROFILING.md clocks line).  No threads: the child writes a file that is parsed afterwards."""

    Q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpus):
        self.gpus = set(gpus)
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL, stdin=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=10)
            except subprocess.TimeoutExpired:
                self.p.kill()
                self.p.wait()
        self.f.flush()
        self.f.seek(0)
        per_gpu, reasons, mx = {}, set(), None
        for line in self.f:
            c = [x.strip() for x in line.split(",")]
            try:
                idx, mhz, mxm = int(c[0]), float(c[1]), float(c[2])
            except (ValueError, IndexError):
                continue
            if idx not in self.gpus:
                continue
            per_gpu.setdefault(idx, []).append(mhz)
            mx = mxm
            for nm, v in zip(self.NAMES, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        meds = [sorted(v)[len(v) // 2] for v in per_gpu.values() if v]
        return {"sm_mhz": min(meds) if meds else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": sum(len(v) for v in per_gpu.values()), "gpus_sampled": sorted(per_gpu)}


def cpu_baseline(a, gph):
    """The oracle (kind 'port', single thread like the reference) on prefixes of the same workload, sizes per
    SURVEY §8(d); `value` is the packed-adjacency port at the largest size, the reference-shaped figure (dense double
    matrices re-allocated and copied per closure, as Pcm.h:739-746 does) is reported beside it."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    sizes = [int(x) for x in a.cpu_sizes.split(",") if x]
    rows = []
    for n in sizes:
        lcs = gph["lcs"][:n]
        rec = {"closures": n, "pairs": n * (n - 1) // 2}
        for shaped in (False, True):
            if shaped and n > 1000:
                continue  # the O(n^3) matrix copies dominate beyond that
            o = orc.OraclePcm(3, 0, odom_threshold=-1, lc_threshold=a.lc_threshold)
            o.set_reference_shaped(shaped)
            o.update(gph["odom"], gph["values"])
            t0 = time.perf_counter()
            o.update(lcs, [])
            dt = time.perf_counter() - t0
            rec["reference_shaped" if shaped else "packed"] = {"pair_checks_per_s": o.pair_checks() / dt, "seconds": dt,
                                                                "inliers": o.num_inliers()}
        rows.append(rec)
    best = rows[-1]["packed"]
    shaped = [r["reference_shaped"]["pair_checks_per_s"] for r in rows if "reference_shaped" in r]
    return {"value": best["pair_checks_per_s"], "unit": UNIT, "cores": 1, "kind": "port",
            "reference_shaped_value": shaped[-1] if shaped else None, "samples": rows,
            "sample": "first n closures of the same workload for n in %s (oracle PCM incl. clique, one thread); value = packed "
                      "adjacency at n=%d; reference_shaped_value = dense double adj+dist matrices re-copied per closure "
                      "(Pcm.h:739-746) at n<=1000" % (sizes, sizes[-1])}


def _ref_worker(args):
    seed, poses, n, thr = args[:4]
    shaped = bool(args[4]) if len(args) > 4 else False
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    synth = importlib.import_module("kimera-rpgo_b200.synth")
    gph = synth.config2(seed=seed, P=poses, n=n)
    o = orc.OraclePcm(3, 0, odom_threshold=-1, lc_threshold=thr)
    o.set_reference_shaped(shaped)  # True: dense double adj + dist matrices re-allocated and copied per closure (Pcm.h:739-746)
    o.update(gph["odom"], gph["values"])
    t0 = time.perf_counter()
    o.update(gph["lcs"], [])
    return (o.pair_checks(), time.perf_counter() - t0)


def run_reference(a):
    """--impl reference: the reference's CPU path (oracle port; the reference itself needs GTSAM and cannot be built
    here) on all host cores: one independent config-2-shaped graph (same generator, the largest size the CPU path runs
    in seconds: P=2500 poses, n=1000 closures) per core per step.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n = a.ref_closures
    poses = max(2500, int(2.5 * n))
    ctx = mp.get_context("spawn")
    pool = ctx.Pool(cores)
    jobs = [(100 + i, poses, n, a.lc_threshold) for i in range(cores)]
    for _ in range(max(min(a.warmup, 2), 1)):
        pool.map(_ref_worker, jobs)
    t0 = time.perf_counter()
    pairs = 0
    per_core = []
    for _ in range(a.steps):
        for p, dt in pool.map(_ref_worker, jobs):
            pairs += p
            per_core.append(p / dt)
    dt = time.perf_counter() - t0
    # one extra, untimed pass in the reference's own memory layout (same graphs), reported beside the timed figure
    t1 = time.perf_counter()
    shaped_pairs = sum(p for p, _ in pool.map(_ref_worker, [j + (True,) for j in jobs]))
    shaped_val = shaped_pairs / (time.perf_counter() - t1)
    pool.close()
    pool.join()
    val = pairs / dt
    per_core.sort()
    sample = ("%d processes x (3D helix, P=%d, n=%d closures, 50%% outliers: BASELINE config 2's generator) per step; oracle "
              "port of Pcm.h with packed adjacency, single-threaded per graph like the reference" % (cores, poses, n))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "reference_sample": sample,
                   "same_config": "same generator, largest size the CPU path can run per step"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "per_core_median": per_core[len(per_core) // 2], "packed_value": val,
                         "reference_shaped_value": shaped_val},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def event_ms(torch, st, fn, reps=3):
    """best-of-reps CUDA-event time of fn() on stream st"""
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st)
            fn()
            e1.record(st)
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return
    import faulthandler
    faulthandler.enable()
    import ctypes as C

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
    handles = []

    def new_handle(d=3, mode=0, **params):
        h = pkg.PcmGpu(d, mode, device=local, kernel=a.kernel, rank=rank, world=world, **params)
        if world > 1:
            par.attach_comm(h)  # rank 0's NCCL id over the process group, then rpgo_comm_init on every rank
        handles.append(h)
        return h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    gph = synth.config2(seed=4, P=a.poses, n=a.closures)
    arr = synth.as_arrays(gph)
    n = a.closures
    pairs = n * (n - 1) // 2
    params = dict(odom_threshold=-1.0, lc_threshold=a.lc_threshold)

    # ---- resident state for the kernel-timed leg -------------------------------------------------
    g = new_handle(**params)
    g.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
    g.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])  # all-gathers inside when world > 1
    st = torch.cuda.ExternalStream(g.stream_ptr(), device=device)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2

    def one_step():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        flush.fill_(1)  # L2 flush between iterations (outside the timed events)
        torch.cuda.synchronize()
        with torch.cuda.stream(st):
            ev[0].record(st)
            g.pairwise_only(0, 0)
            ev[1].record(st)
            g.allgather(0)  # world == 1: mirror + degrees only
            ev[2].record(st)
            size, ids, _ = g.find_inliers_raw(0, pkg.CLIQUE_HEU)
            ev[3].record(st)
        ev[3].synchronize()
        return ev[0].elapsed_time(ev[3]), ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3]), ev[1].elapsed_time(ev[2]), size

    for _ in range(a.warmup):
        one_step()
    barrier()
    clocks = ClockLog(range(world)) if rank == 0 else None
    l0 = g.launch_count()
    tot = k3 = cl = xg = 0.0
    size = 0
    for _ in range(a.steps):
        t, tk, tc, tx, size = one_step()
        tot += t
        k3 += tk
        cl += tc
        xg += tx
    barrier()
    launches = g.launch_count() - l0
    clock_summary = clocks.stop() if clocks is not None else None
    tot, k3, cl, xg = max_over_ranks([tot, k3, cl, xg])
    ms_per_step = tot / a.steps
    value = pairs / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    tf = C.c_double()
    g.lib.rpgo_fp64_peak(local, C.byref(tf))
    k3_ms = k3 / a.steps
    achieved = FLOP_PER_PAIR["pcm3d"] * (pairs / world) / (k3_ms * 1e-3) / 1e12
    kname = {0: "pairwise_grouped_kernel<3,PCM,12,3,504,1>", 1: "pairwise_direct_kernel<3,PCM,2>", 2: "pairwise_grouped_kernel<3,PCM,12,3,504,1>",
             24: "pairwise_tiled_kernel<3,12,1,2>", 22: "pairwise_tiled_kernel<3,12,1,1>"}.get(a.kernel, "?")
    roofline = {"bound": "fp64", "achieved": achieved, "peak": tf.value, "unit": "TFLOP/s", "frac": achieved / tf.value,
                "peak_nominal": FP64_NOMINAL_TFLOPS, "frac_nominal": achieved / FP64_NOMINAL_TFLOPS,
                "traffic": None, "kernel": kname,
                "peak_source": "measured live by rpgo_fp64_peak (DFMA micro-benchmark, ncu capture in profiles/; "
                               "MEASURED_PEAKS.json has no FP64 entry); peak_nominal = 148 SM x 64 FMA/clk x 2 x 1.965 GHz",
                "algorithmic_flop_per_pair": FLOP_PER_PAIR["pcm3d"], "kernel_ms": k3_ms}
    # DRAM traffic of that kernel from the committed ncu --set full capture of this very configuration (one GPU)
    for name in ("r2_k3_traffic.json", "r1_k3_traffic.json"):
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", name)))
            if tr["kernel"] == roofline["kernel"] and a.closures == 50000 and world == 1:
                roofline["traffic"] = tr["traffic_bytes_per_launch"]
                roofline["traffic_source"] = tr["source"]
                break
        except (OSError, KeyError, ValueError):
            pass

    # ---- e2e through the public host API, host buffers ---------------------------------------------
    e2e = None
    if not a.no_e2e:
        p = new_handle(**params)  # a long-lived solver: created once, reset per graph
        times = []
        h2d = d2h = 0
        for it in range(5 + 1):  # one untimed warm-up, five timed; the median is reported, all five are listed
            barrier()
            t0 = time.perf_counter()
            p.reset()
            h2d = p.odom_append_arrays(arr["o_prev"], arr["o_new"], arr["o_pose"], arr["o_cov"], arr["o_init"])
            p.lc_append_arrays(arr["l_from"], arr["l_to"], arr["l_pose"], arr["l_cov"])  # incl. all-gather when world > 1
            h2d += p.last_h2d_bytes
            sz, ids, _ = p.find_inliers_raw(0, pkg.CLIQUE_HEU)
            d2h = p.last_d2h_bytes + 4 * int(sz) + 8
            p.sync()
            dt = time.perf_counter() - t0
            if it > 0:
                times.append(dt)
        assert int(sz) == int(size)
        (te,) = max_over_ranks([sorted(times)[len(times) // 2]])
        e2e = {"value": pairs / te, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": te * 1e3, "ms_all_rank0": [round(t * 1e3, 1) for t in times],
               "what": "rpgo_reset -> odom_append(P-1 factors) -> lc_append(n closures, incl. all-gather) -> find_inliers, "
                       "numpy host buffers; handle/stream/arena/communicator are long-lived like a RobustSolver"}
        p.close()

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "pairs_per_step": pairs, "l2": "flushed between iterations (256 MB fill)",
                   "parallelism": "rows sharded over %d GPU(s), NCCL all-gather of the adjacency bitset and NCCL all-reduce of "
                                  "the clique incumbent inside the C-ABI library" % world},
        "k3_ms": k3_ms, "allgather_mirror_degree_ms": xg / a.steps, "max_clique_ms": cl / a.steps, "max_clique_size": int(size),
        "roofline": roofline, "gpu_launches": int(launches), "clocks": clock_summary,
    }
    if e2e:
        out["e2e"] = e2e
    g.close()

    # ---- the other named configurations and per-mode kernels, outside the headline timing ---------------
    if not a.no_extras:
        import bench_extras
        ctx = dict(pkg=pkg, synth=synth, torch=torch, np=np, new_handle=new_handle, barrier=barrier, max_over_ranks=max_over_ranks,
                   world=world, rank=rank, device=device, event_ms=event_ms, peak_tflops=tf.value, flop=FLOP_PER_PAIR,
                   nominal=FP64_NOMINAL_TFLOPS, args=a)
        out["extra"] = bench_extras.run(ctx, [x for x in a.extras.split(",") if x])
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(a, gph)
    if rank == 0:
        print(json.dumps(out), flush=True)

    # ---- teardown: library handles first, then the process group ---------------------------------------
    for h in handles:
        h.close()
    del flush
    barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
