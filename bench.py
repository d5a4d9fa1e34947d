#!/usr/bin/env python
"""bench.py -- segment-pair tests/s and views/s of the Line3D++ matching -> scoring -> affinity
path on B200 (BASELINE.json metric), with the K1 roofline, the CPU baseline and the end-to-end
number.  One "step" = one full pass of stages 1-4 (l3d_match_images + l3d_affinity) over one
synthetic scene.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c4s]

The default single-GPU line also carries `config3_stream`: BASELINE config[2], the key-frame stream through
the incremental mode (what `--workload c3` measures on its own; `--no-stream` skips it).

N=1: BASELINE config[1] (50 views x 1000 segments x 10 neighbours, 640x480).  N>1 (torchrun, one
process per GPU): the scene grows with N (50*N views, weak scaling); every rank owns a contiguous
slice of reference views (matching, scoring rows, hypotheses, affinity edges), four NCCL
all-gathers per step make the results whole on every rank.
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
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

FLOP_PER_TEST = 114.0  # algorithmic FP32 flop of one segment-pair test


def make_workload(scene_mod, name, n_gpus):
    if name == "c2":
        return scene_mod.make_scene("c2", n_views=50 * n_gpus), "c2" if n_gpus == 1 else "c2x%d" % n_gpus
    if name == "c4":
        return scene_mod.make_scene("c4"), "c4"
    if name == "c4s":  # BASELINE config[3] shape, 100 of the 500 views (oracle-checkable in ~1 min)
        return scene_mod.make_scene("c4", n_views=100), "c4s"
    if name == "tiny":
        return scene_mod.make_scene("tiny"), "tiny"
    raise SystemExit("unknown workload %r" % name)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md).  The query
    loop is started before the warm-up (nvidia-smi needs a few hundred ms to produce its first line)
    and every line carries a timestamp, so the samples of the timed region can be picked afterwards."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        import datetime
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for seen, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except ValueError:
                ts = seen
            try:
                rows.append((ts, float(f[1]), float(f[2]), [n for n, v in zip(names, f[4:8]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        note = None
        inside = rows if t0 is None else [r for r in rows if t0 <= r[0] <= t1]
        if not inside and rows and t0 is not None:
            # the timed region is shorter than the sampling period: take the samples closest to it
            mid = 0.5 * (t0 + t1)
            inside = sorted(rows, key=lambda r: abs(r[0] - mid))[:3]
            note = "timed region (%.0f ms) shorter than the sampling period; nearest samples" % (1e3 * (t1 - t0))
        out = {"sm_mhz": float(np.median([r[1] for r in inside])) if inside else None,
               "sm_max_mhz": max(r[2] for r in inside) if inside else None, "samples": len(inside),
               "reasons": sorted({n for r in inside for n in r[3]})}
        if note:
            out["note"] = note
        return out


def run_cpu_oracle(scene, threads=0, snapshot=False):
    import oracle_py
    t0 = time.perf_counter()
    o = oracle_py.run_scene(scene, threads=threads, snapshot=snapshot)
    dt = time.perf_counter() - t0
    tests = o.pair_tests()
    tm = o.timers()
    cores = oracle_py.lib().orc_max_threads()
    o.close()
    return dict(seconds=dt, tests=tests, timers=tm, cores=cores)


def bench_reference(args, scene_mod):
    """--impl reference: the reference's own CPU implementation of the path = the oracle port
    (the reference cannot be compiled here: no Eigen/Boost/OpenCV), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "c3":
        # the key-frame stream through the oracle's incremental mode, same cycles as the product's arm
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_py
        import stream_utils
        fill = int(os.environ.get("L3D_C3_FILL", "24"))
        W, K = max(args.warmup, 3), min(args.steps, 40)
        st = scene_mod.make_stream(n_keyframes=5 + fill + W + K, n_seg=1000, window=20, nbrs=10, jitter=0.3)
        o, oc = stream_utils.oracle_driver(oracle_py, st)
        ts, tests = [], []
        for ci, cy in enumerate(st.cycles):
            before = o.pair_tests()
            t0 = time.perf_counter()
            oc["begin_cycle"]()
            for cam in cy.deletes:
                oc["delete"](cam)
            for v in cy.adds:
                oc["add"](v, v.worldpoints)
            for cam, R, t, md, lst in cy.updates:
                oc["update"](cam, R, t, md, lst)
            oc["match"](st.params)
            oc["reconstruct"]()
            dt = time.perf_counter() - t0
            if ci >= fill + W and len(ts) < K:
                ts.append(dt)
                tests.append(o.pair_tests() - before)
        cores = oracle_py.lib().orc_max_threads()
        T = float(np.sum(ts))
        value = float(np.sum(tests)) / T
        sample = "the same stream, %d steady-state cycles" % len(ts)
        print(json.dumps({
            "impl": "reference", "metric": "segment_pair_tests_per_s", "value": value, "unit": "tests/s",
            "n_gpus": args.gpus, "steps": len(ts), "warmup": W, "ms_per_step": 1e3 * T / len(ts),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "c3", "window": 20, "segments_per_view": 1000, "neighbours": 10, "sample": sample},
            "views_per_s": 20.0 * len(ts) / T,
            "cpu_baseline": {"value": value, "unit": "tests/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return
    # each step is a bounded sample of the arm's workload: the single-GPU scene of the same generator
    # (at N > 1 the product's scene has N times the views; tests/s of the CPU path does not depend on it)
    scene, sample_name = make_workload(scene_mod, args.workload, 1)
    n = max(args.gpus, 1)
    wname = sample_name if (n == 1 or args.workload != "c2") else "c2x%d" % n
    full_views = scene.num_views * (n if args.workload == "c2" else 1)
    times, tests, cores = [], 0, 1
    for i in range(args.warmup + args.steps):
        r = run_cpu_oracle(scene)
        tests, cores = r["tests"], r["cores"]
        if i >= args.warmup:
            times.append(r["timers"]["match_images"] + r["timers"]["reconstruct"])
    T = float(np.sum(times))
    value = tests * len(times) / T
    sample = "%s scene (%d views, same generator and per-view shape), stages 1-4, %d passes" % (sample_name, scene.num_views, len(times))
    line = {
        "impl": "reference", "metric": "segment_pair_tests_per_s", "value": value, "unit": "tests/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wname, "views": full_views, "segments_per_view": scene.views[0].segs.shape[0],
                   "neighbours": scene.params["num_neighbors"], "sample": sample},
        "views_per_s": scene.num_views * len(times) / T,
        "cpu_baseline": {"value": value, "unit": "tests/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def bench_stream(args, scene_mod, emit=True):
    """--workload c3: BASELINE config[2], the key-frame stream through the incremental mode
    (l3d_stream_*; 640x480, 1000 segments per key frame, window of 20, 10 neighbours).  One step = one
    L3DPPing cycle in steady state: delete the culled key frames, add the new one (host buffers),
    re-pose every current key frame, matchImages, reconstruct3Dlines -- the whole cycle is the public
    API with host inputs, so the step time IS the end-to-end time; `value` counts the cycle's new
    segment-pair tests.  Single GPU (the mode does not shard: replicas only)."""
    import torch
    api = importlib.import_module("3dline-slam_b200.api")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import stream_utils
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    # the walk over the views is a dependency chain and a cycle brings only ~10 new pairs: the mode does not
    # shard.  N > 1 = N independent replicas (one stream per GPU), whole-job value = N streams / slowest rank.
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    fill = int(os.environ.get("L3D_C3_FILL", "24"))  # cycles until the window of 20 is full and sliding (profiling runs shorten it)
    W, K = max(args.warmup, 3), min(args.steps, 100)
    st = scene_mod.make_stream(n_keyframes=5 + fill + W + K, n_seg=1000, window=20, nbrs=10, jitter=0.3)
    for cy in st.cycles:  # the caller holds its world-point lists as arrays (no per-call list conversion)
        cy.updates = [(cam, R, t, md, np.asarray(lst, dtype=np.uint32)) for cam, R, t, md, lst in cy.updates]
        for v in cy.adds:
            v.worldpoints = np.asarray(v.worldpoints, dtype=np.uint32)
    l3, calls = stream_utils.cuda_driver(api, st)
    ts, tests, launches = [], [], []
    state = {"t0": 0.0}

    parts = []

    def cycle(cy, calls):
        a = time.perf_counter()
        calls["begin_cycle"]()
        for cam in cy.deletes:
            calls["delete"](cam)
        for v in cy.adds:
            calls["add"](v, v.worldpoints)
        for cam, R, t, md, lst in cy.updates:
            calls["update"](cam, R, t, md, lst)
        b = time.perf_counter()
        calls["match"](st.params)
        c = time.perf_counter()
        calls["reconstruct"]()
        parts.append((b - a, c - b, time.perf_counter() - c))

    sampler = ClockSampler(dev.index)
    if not os.environ.get("L3D_BENCH_NO_SAMPLER"):
        sampler.start()
    if dist is not None:
        dist.barrier()
    wall0 = wall1 = time.time()
    stage = {}
    h2d = d2h = 0
    for ci, cy in enumerate(st.cycles):
        timed = ci >= fill + W and len(ts) < K
        if timed and not ts:
            wall0 = time.time()
        l3.reset_counters()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        cycle(cy, calls)
        ij, w = l3.edges()                      # A_ back on the host, like the reference's consumer
        ids = l3.cluster_ids()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if timed:
            c = l3.counts()
            ts.append(dt)
            tests.append(c["pair_tests"])
            launches.append(c["gpu_launches"])
            for k, v in l3.timings().items():
                stage[k] = stage.get(k, 0.0) + v
            h2d += sum(v.segs.nbytes for v in cy.adds) + 96 * len(cy.updates)
            d2h += ij.nbytes + w.nbytes + ids.nbytes
            wall1 = time.time()
    clocks = sampler.stop(wall0, wall1) if not os.environ.get("L3D_BENCH_NO_SAMPLER") else None
    T = float(np.sum(ts))
    n = len(ts)
    if dist is not None:  # replicas: the job is as slow as its slowest rank
        t = torch.tensor([T], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        T = float(t.item())
        dist.barrier()
        dist.destroy_process_group()
        if rank != 0:
            return
    if os.environ.get("L3D_BENCH_TRACE"):
        for i in np.argsort(-np.asarray(ts))[:4]:
            ci = fill + W + int(i)
            print("slow cycle %d: %.2f ms (host calls %.2f, match %.2f, reconstruct %.2f)" %
                  (ci, 1e3 * ts[i], 1e3 * parts[ci][0], 1e3 * parts[ci][1], 1e3 * parts[ci][2]), file=sys.stderr)
        pm = np.median(np.asarray(parts[fill + W:fill + W + n]), axis=0)
        print("median parts: host calls %.2f ms, match %.2f ms, reconstruct %.2f ms" % tuple(1e3 * pm), file=sys.stderr)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        import oracle_py
        o, oc = stream_utils.oracle_driver(oracle_py, st)
        ct, ctests = 0.0, 0
        for ci, cy in enumerate(st.cycles[:fill + W + min(K, 20)]):
            before = o.pair_tests()
            t0 = time.perf_counter()
            cycle(cy, oc)
            d = time.perf_counter() - t0
            if ci >= fill + W:
                ct += d
                ctests += o.pair_tests() - before
        cores = oracle_py.lib().orc_max_threads()
        o.close()
        cpu = {"value": ctests / ct, "unit": "tests/s", "cores": cores, "kind": "port",
               "sample": "the same stream, the first %d timed cycles" % min(K, 20),
               "ms_per_cycle": 1e3 * ct / min(K, 20)}
    value = world * float(np.sum(tests)) / T
    line = {
        "metric": "segment_pair_tests_per_s", "value": value, "unit": "tests/s", "n_gpus": world, "steps": n, "warmup": W,
        "ms_per_step": 1e3 * T / n, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
        "data": "synthetic",
        "config": {"workload": "c3", "keyframes": len(st.cycles) + 4, "window": 20, "segments_per_view": 1000,
                   "neighbours": 10, "new_keyframes_per_cycle": 1, "step": "one L3DPPing cycle in steady state, wall clock",
                   "parallelism": "1 rank" if world == 1 else "%d independent replicas (the mode does not shard)" % world,
                   "l2": "inputs larger than one kernel's footprint are not the bound here: launch-bound"},
        "views_per_s": world * 20.0 * n / T,
        "ms_per_cycle_p50": 1e3 * float(np.median(ts)), "ms_per_cycle_max": 1e3 * float(np.max(ts)),
        "stage_ms": {k: v / n for k, v in stage.items()},
        "e2e": {"value": value, "unit": "tests/s", "h2d_bytes_per_step": world * (h2d // n),
                "d2h_bytes_per_step": world * (d2h // n)},
        "gpu_launches": int(np.sum(launches)), "clocks": clocks,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    if emit:
        print(json.dumps(line))
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stream", action="store_true", help="skip the config-3 key-frame stream measurement")
    ap.add_argument("--check", action="store_true", help="also verify the result against the CPU oracle")
    ap.add_argument("--trace-phases", action="store_true", help="N>1: wall time per phase/exchange (stderr)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    scene_mod = importlib.import_module("3dline-slam_b200.scene")
    if args.impl == "reference":
        bench_reference(args, scene_mod)
        return
    if args.workload == "c3":
        bench_stream(args, scene_mod)
        return

    import torch
    api = importlib.import_module("3dline-slam_b200.api")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    torch.cuda.set_device(dev)
    n_gpus = world if world > 1 else 1
    if args.gpus != n_gpus and rank == 0:
        print("note: --gpus %d but WORLD_SIZE=%d; using %d" % (args.gpus, world, n_gpus), file=sys.stderr)

    scene, wname = make_workload(scene_mod, args.workload, n_gpus)
    prm = scene.params
    stream = torch.cuda.current_stream(dev)
    l3 = api.Line3D("", False, scene.max_image_width, 3000, False, True, dev.index, stream.cuda_stream)
    l3.shard = (rank, n_gpus)
    l3.load_scene(scene)
    l3.upload()  # tables resident in HBM before the timed region

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    sharding = importlib.import_module("3dline-slam_b200.sharding")
    xch = sharding.Exchanger(dist, torch, dev) if n_gpus > 1 else None

    def step():
        if n_gpus == 1:
            l3.matchImages(prm["sigma_p"], prm["sigma_a"], prm["num_neighbors"], prm["epipolar_overlap"], prm["knn"],
                           prm["const_reg_depth"])
            l3.affinity()
        else:
            # view slices per rank; forward matches, fold programs, hypotheses and edges are
            # all-gathered with NCCL (3dline-slam_b200/sharding.py)
            sharding.run_sharded(l3, xch, prm)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(dev.index)
    sampler.start()
    for _ in range(args.warmup):
        flush.zero_()
        step()
    barrier()

    l3.reset_counters()
    total_ms = 0.0
    stage_ms = {}
    k1_ms, k1_launches = 0.0, 0
    barrier()
    wall0 = time.time()
    for _ in range(args.steps):
        flush.zero_()  # flush L2 between timed iterations (inputs are smaller than L2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        step()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        total_ms += e0.elapsed_time(e1)
        # stage timers of the library (CUDA events on the same stream); affinity() resets them,
        # so read both halves
        for k, v in l3.timings().items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v
    barrier()
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1)
    cnt = l3.counts()
    launches = cnt["gpu_launches"]

    # max over ranks
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        tt = torch.tensor([float(cnt["pair_tests"])], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        tests_per_step = float(tt.item())
    else:
        tests_per_step = float(cnt["pair_tests"])

    # ---- separate instrumented passes for the per-stage / K1 numbers (same workload) ----
    if n_gpus == 1:
        l3.match_stage12(prm["sigma_p"], prm["sigma_a"], prm["num_neighbors"], prm["epipolar_overlap"], prm["knn"],
                         prm["const_reg_depth"])
        t12 = l3.timings()
        c12 = l3.counts()
        l3.match_stage3()
        t3 = l3.timings()
        l3.affinity()
        t4 = l3.timings()
    else:
        step()
        t12 = t3 = t4 = l3.timings()   # the stage timers accumulate over the phases of one step
        c12 = l3.counts()
    cfin = l3.counts()
    if args.trace_phases and n_gpus > 1:
        tr = {}
        for _ in range(10):
            barrier()
            sharding.run_sharded(l3, xch, prm, trace=tr)
        if rank == 0:
            print("phase ms (rank 0, synchronised after every phase): " +
                  ", ".join("%s %.3f" % (k, 1e2 * v) for k, v in tr.items()), file=sys.stderr)
    k1_s = max(t12["k1_kernel"], 1e-9) * 1e-3
    k1_tests = float(c12["pair_tests"])

    value = tests_per_step * args.steps / (total_ms * 1e-3)
    views_per_s = scene.num_views * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public API with host buffers (H2D of the scene, D2H of A_) ----
    # every rank uploads the (replicated) tables and reads back A_; wall clock, max over ranks
    h2d = scene.total_segments() * 16 + scene.num_views * (8 * 21 + 20) + sum(4 * len(v.neighbors) for v in scene.views)
    d2h = 0
    t_e2e = 0.0
    e2e_steps = min(args.steps, 20)
    for i in range(2 + e2e_steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        l3.upload()                      # host -> device: segments, cameras, neighbour lists
        step()
        ij, w = l3.edges()               # device -> host: A_ (what the CPU clustering consumes)
        l2g = l3.local2global()
        barrier()
        dt = time.perf_counter() - t0
        if i >= 2:
            t_e2e += dt
        d2h = ij.nbytes + w.nbytes + l2g.nbytes
    if dist is not None:
        t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e = {"value": tests_per_step * e2e_steps / t_e2e, "unit": "tests/s", "h2d_bytes_per_step": int(h2d) * n_gpus,
           "d2h_bytes_per_step": int(d2h) * n_gpus, "ms_per_step": 1e3 * t_e2e / e2e_steps, "steps": e2e_steps,
           "views_per_s": scene.num_views * e2e_steps / t_e2e,
           "what": "Line3D.upload (pageable host arrays -> HBM) + matchImages + affinity + edges()/local2global() D2H, per rank"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K1) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max = clocks.get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    fp32_nominal = n_sm * 128 * 2 * sm_max * 1e6 / 1e12
    fp32_measured = api.Context(dev.index, stream.cuda_stream).fp32_peak_tflops()
    achieved = FLOP_PER_TEST * k1_tests / k1_s / 1e12
    roofline = {
        "kernel": "k1_pairtest_kernel", "bound": "fp32", "achieved": achieved, "peak": fp32_measured,
        "unit": "TFLOP/s", "frac": achieved / fp32_measured if fp32_measured else None,
        "peak_source": "measured here: dense FFMA micro-benchmark (l3d_bench_fp32_peak); nominal %d SMs x 128 x 2 x %.0f MHz = %.1f TFLOP/s"
                       % (n_sm, sm_max, fp32_nominal),
        "peak_nominal": fp32_nominal, "frac_of_nominal": achieved / fp32_nominal,
        # the kernel's own formulation issues 35.5 SASS instructions per test (cuobjdump count over the
        # unrolled loop): fraction of the lane-issue capacity (SMs x 4 schedulers x 32 lanes x clock)
        "issued_instructions_per_test": 35.5,
        "lane_issue_frac": 35.5 * k1_tests / k1_s / (n_sm * 128 * sm_max * 1e6),
        "algorithmic_flop_per_test": FLOP_PER_TEST, "tests_per_launch": k1_tests / max(t12["k1_launches"], 1),
        "launch_ms": 1e3 * k1_s / max(t12["k1_launches"], 1), "k1_tests_per_s": k1_tests / k1_s,
        # dram__bytes_read.sum + dram__bytes_write.sum of one K1 launch on C2 (ncu --set full,
        # profiles/r1f_ncu_full_c2.md); the 31 MB bit mask stays in the 126 MB L2
        "traffic": 2435840 if wname == "c2" else None,
    }
    n_pairs_local = max(c12["num_pairs_local"], 1)
    seg_n = scene.views[0].segs.shape[0]
    k1_bytes = n_pairs_local * (32.0 + 16.0) * seg_n + k1_tests / 8.0 + 4.0 * n_pairs_local * seg_n
    roofline["hbm"] = {"algorithmic_bytes_per_launch": k1_bytes / max(t12["k1_launches"], 1),
                       "achieved_gbs": k1_bytes / k1_s / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                       "frac": (k1_bytes / k1_s / 1e9) / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                       "note": "descriptors + segments + 1 bit per test + counts; the kernel is FP32-pipe bound, not HBM bound"}

    # the other stages against the roofline that bounds them (SURVEY.md 8d figures); they are
    # latency / dependency bound at this scene size, which is what the small fractions say
    fp64_nominal = n_sm * 64 * 2 * sm_max * 1e6 / 1e12
    k2_flop = 330.0 * float(c12["candidates"])
    stage_rooflines = {
        "k2_exact": {"bound": "fp64", "algorithmic_flop_per_candidate": 330.0, "candidates": c12["candidates"],
                     "achieved_tflops": k2_flop / max(t12["exact"] * 1e-3, 1e-9) / 1e12, "peak_tflops": fp64_nominal,
                     "frac": k2_flop / max(t12["exact"] * 1e-3, 1e-9) / 1e12 / fp64_nominal},
        "k3_score": {"bound": "fp32+sfu, dependency chain over the views",
                     "sim_evals_per_s": float(cfin["sim_evals"]) / max(t3["score"] * 1e-3, 1e-9),
                     "algorithmic_flop_per_sim_eval": 40.0,
                     "achieved_tflops": 40.0 * float(cfin["sim_evals"]) / max(t3["score"] * 1e-3, 1e-9) / 1e12},
        "k4_affinity": {"bound": "hbm gather", "algorithmic_bytes_per_edge_test": 160.0,
                        "edge_tests": cfin["filtered_entries"],
                        "achieved_gbs": 160.0 * float(cfin["filtered_entries"]) / max(t4["affinity"] * 1e-3, 1e-9) / 1e9,
                        "peak_gbs": peaks.get("hbm_gbs")},
    }

    cpu_baseline = None
    if not args.no_cpu_baseline and n_gpus == 1:   # reported on rank 0 at N=1 only
        base_scene, bname = make_workload(scene_mod, args.workload, 1)
        r = run_cpu_oracle(base_scene)
        tcpu = r["timers"]["match_images"] + r["timers"]["reconstruct"]
        cpu_baseline = {"value": r["tests"] / tcpu, "unit": "tests/s", "cores": r["cores"], "kind": "port",
                        "sample": "full %s scene (%d views), stages 1-4, one pass, %.1f s" % (bname, base_scene.num_views, tcpu),
                        "views_per_s": base_scene.num_views / tcpu,
                        "stage1_tests_per_s": r["tests"] / max(r["timers"]["match"], 1e-9)}

    if args.check:
        import oracle_py
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from parity_utils import compare_full
        if n_gpus == 1:
            l3.reconstruct3Dlines()
            orc = oracle_py.run_scene(scene)
            print("check vs oracle:", compare_full(l3, orc, scene, check_scored=False), file=sys.stderr)
        else:  # the other ranks have left by now; sharded parity is tests/test_parity_gpu.py / test_full_size_gpu.py
            print("--check is a single-GPU option (sharded == unsharded == oracle is covered by the GPU tests)",
                  file=sys.stderr)

    line = {
        "metric": "segment_pair_tests_per_s", "value": value, "unit": "tests/s", "n_gpus": n_gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 pre-filter + f64 exact", "data": "synthetic",
        "config": {"workload": wname, "views": scene.num_views, "segments_per_view": scene.views[0].segs.shape[0],
                   "neighbours": prm["num_neighbors"], "image": "%dx%d" % (scene.views[0].width, scene.views[0].height),
                   "l2": "flushed between timed iterations (256 MB write)",
                   "parallelism": ("1 rank" if n_gpus == 1 else
                                   "%d contiguous view slices: matching, scoring rows, hypotheses and edges per slice; "
                                   "4 NCCL all-gathers per step; the score fold is replicated" % n_gpus)},
        "views_per_s": views_per_s,
        "stage1_tests_per_s": k1_tests / max((t12["pairtest"] + t12["exact"]) * 1e-3, 1e-9),
        "stage_ms": {"prep": t12["prep"], "k1_pairtest": t12["pairtest"], "k2_exact": t12["exact"],
                     "k3_score": t3["score"], "k4_affinity": t4["affinity"]},
        "counts": {k: cfin[k] for k in ("pair_tests", "candidates", "forward_matches", "scored_entries", "sim_evals",
                                         "filtered_entries", "num_pairs", "num_entries", "num_edges", "num_local_ids")},
        "roofline": roofline, "stage_rooflines": stage_rooflines, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    # BASELINE config 3 (key-frame stream, incremental mode) rides along on the default single-GPU run
    if n_gpus == 1 and args.workload == "c2" and not args.no_stream:
        try:
            import copy
            a3 = copy.copy(args)
            a3.steps, a3.warmup = 40, 3
            s3 = bench_stream(a3, scene_mod, emit=False)
            line["config3_stream"] = {
                "what": "bench.py --workload c3: one L3DPPing cycle in steady state (window 20, 1000 segments, 10 neighbours, "
                        "one new key frame per cycle), public API with host inputs, wall clock",
                "ms_per_cycle": s3["ms_per_step"], "ms_per_cycle_p50": s3["ms_per_cycle_p50"],
                "tests_per_s": s3["value"], "views_per_s": s3["views_per_s"], "stage_ms": s3["stage_ms"],
                "gpu_launches_per_cycle": s3["gpu_launches"] / max(s3["steps"], 1),
                "cpu_baseline": s3.get("cpu_baseline")}
        except Exception as e:  # the headline line must not depend on the extra measurement
            line["config3_stream"] = {"error": repr(e)}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
