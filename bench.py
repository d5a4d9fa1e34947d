#!/usr/bin/env python
"""bench.py -- segment-pair tests/s and views/s of the Line3D++ matching -> scoring -> affinity
path on B200 (BASELINE.json metric), with the rooflines of the dominant kernels, the CPU baseline,
the end-to-end number and a parity digest.  One "step" = one full pass of stages 1-4
(l3d_match_images + l3d_affinity) over one synthetic scene.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c4|c2|c3|c4s|c5s|c5|tiny]

Default workload at every N: BASELINE config[3], the 500-view x 3000-segment x 20-neighbour scene the
metric is quoted on "at 1/2/4/8 B200" -- the SAME scene at every N, i.e. strong scaling.  N>1
(torchrun, one process per GPU): every rank owns a contiguous slice of reference views (matching,
scoring rows, hypotheses, affinity edges); NCCL all-gathers make the results whole on every rank
(DESIGN.md section 6).  After the timed region every rank hashes its results (filtered lists,
hypotheses, A_, local2global_); the digests must agree on all ranks and the line carries
`parity_digest`, so runs at N = 1, 2, 4, 8 can be compared.  At N = 1 the line also carries
`config2` (BASELINE config[1], 50 x 1000 x 10, fully measured) and `config3_stream` (config[2], the
key-frame stream through the incremental mode), and a C4-shaped cut is compared with the CPU oracle.
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


# the CPU arm's bounded sample of a workload: a cut of the same generator and per-view shape (the CPU
# path's tests/s does not depend on the number of views), sized for a few seconds per pass
CPU_SAMPLE = {"c4": ("c4", 16), "c4s": ("c4", 16), "c2": ("c2", 50), "tiny": ("tiny", 8), "c5": ("c5", 8), "c5s": ("c5", 8)}


def make_workload(scene_mod, name, n_gpus=1):
    """(scene, workload name).  The scene does not depend on the number of GPUs: strong scaling."""
    if name == "c2":
        return scene_mod.make_scene("c2"), "c2"
    if name == "c2w":  # round-1 weak-scaling variant: the circle grows with N
        return scene_mod.make_scene("c2", n_views=50 * n_gpus), "c2w"
    if name == "c4":
        return scene_mod.make_scene("c4"), "c4"
    if name == "c4s":  # BASELINE config[3] shape, 100 of the 500 views (oracle-checkable in ~1 min)
        return scene_mod.make_scene("c4", n_views=100), "c4s"
    if name == "tiny":
        return scene_mod.make_scene("tiny"), "tiny"
    raise SystemExit("unknown workload %r" % name)


def workload_config(scene_mod, name):
    """The `config` object both arms print (identical dicts, so the driver can see it is the same job)."""
    base = {"c4s": "c4", "c2w": "c2", "c5s": "c5"}.get(name, name)
    V, N, NB, img = scene_mod.PRESET_SHAPES[base]
    if name == "c4s":
        V = 100
    if name == "c5s":
        V = 500
    return {"workload": name, "views": V, "segments_per_view": N, "neighbours": NB, "image": img}


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
    (the reference cannot be compiled here: no Eigen/Boost/OpenCV), all host threads.  Each step is a
    bounded sample of the product arm's workload: a cut of the same generator and per-view shape."""
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
            "config": C3_CONFIG,
            "views_per_s": 20.0 * len(ts) / T,
            "cpu_baseline": {"value": value, "unit": "tests/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return
    kind, nv = CPU_SAMPLE.get(args.workload, ("c4", 16))
    scene = scene_mod.make_scene(kind, n_views=nv)
    cfg = workload_config(scene_mod, args.workload)
    times, tests, cores = [], 0, 1
    for i in range(args.warmup + args.steps):
        r = run_cpu_oracle(scene)
        tests, cores = r["tests"], r["cores"]
        if i >= args.warmup:
            times.append(r["timers"]["match_images"] + r["timers"]["reconstruct"])
    T = float(np.sum(times))
    value = tests * len(times) / T
    sample = ("%d-view cut of the %s generator (same per-view shape: %d segments, %d neighbours), stages 1-4, %.2e tests per "
              "pass, %d passes" % (scene.num_views, kind, cfg["segments_per_view"], cfg["neighbours"], tests, len(times)))
    # views/s of the full workload at the CPU's tests/s (the cut has fewer views, the same tests per view pair)
    line = {
        "impl": "reference", "metric": "segment_pair_tests_per_s", "value": value, "unit": "tests/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "views_per_s_of_the_sample": scene.num_views * len(times) / T,
        "cpu_baseline": {"value": value, "unit": "tests/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


C3_CONFIG = {"workload": "c3", "keyframes": 300, "window": 20, "segments_per_view": 1000, "neighbours": 10,
             "image": "640x480", "new_keyframes_per_cycle": 1}


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
        "config": C3_CONFIG,
        "step": "one L3DPPing cycle in steady state (of a %d-key-frame prefix of the stream), wall clock" % (len(st.cycles) + 4),
        "parallelism": "1 rank" if world == 1 else "%d independent replicas (the mode does not shard)" % world,
        "l2_policy": "not flushed: a cycle is launch-bound, its working set is far below L2",
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


# SASS instruction counts of K1 (profiles/r2_k1_sass.md): per test the kernel RUNS (two tests per loop iteration,
# 90 instructions), and per 32-target word of every source row (the wedge test of the warp, ballot, mask store)
K1_ISSUED_PER_RUN_TEST = 45.0
K1_ISSUED_PER_WORD = 74.0
K2_FLOP_PER_CANDIDATE = 330.0  # SURVEY.md section 8(d) / DESIGN.md section 4.2: FP64 flop per K1 candidate


class Env:
    pass


def setup_env(args):
    import torch
    env = Env()
    env.torch = torch
    env.api = importlib.import_module("3dline-slam_b200.api")
    env.sharding = importlib.import_module("3dline-slam_b200.sharding")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    env.world = int(os.environ.get("WORLD_SIZE", "1"))
    env.rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    env.dist = None
    env.dev = torch.device("cuda", local_rank if env.world > 1 else 0)
    torch.cuda.set_device(env.dev)
    if env.world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=env.dev)
        env.dist = dist
    env.n_gpus = env.world if env.world > 1 else 1
    if args.gpus != env.n_gpus and env.rank == 0:
        print("note: --gpus %d but WORLD_SIZE=%d; using %d" % (args.gpus, env.world, env.n_gpus), file=sys.stderr)
    env.stream = torch.cuda.current_stream(env.dev)
    env.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=env.dev)  # > 126 MB L2
    return env


def barrier(env):
    if env.dist is not None:
        env.dist.barrier()
    env.torch.cuda.synchronize(env.dev)


def max_over_ranks(env, x):
    if env.dist is None:
        return float(x)
    t = env.torch.tensor([float(x)], dtype=env.torch.float64, device=env.dev)
    env.dist.all_reduce(t, op=env.dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(env, x):
    if env.dist is None:
        return float(x)
    t = env.torch.tensor([float(x)], dtype=env.torch.float64, device=env.dev)
    env.dist.all_reduce(t, op=env.dist.ReduceOp.SUM)
    return float(t.item())


def gather_digests(env, hexdigest):
    """Every rank's sha256 -> list of hex strings on every rank."""
    if env.dist is None:
        return [hexdigest]
    torch = env.torch
    mine = torch.tensor(list(bytes.fromhex(hexdigest)), dtype=torch.uint8, device=env.dev)
    allb = torch.empty(32 * env.world, dtype=torch.uint8, device=env.dev)
    env.dist.all_gather_into_tensor(allb, mine)
    raw = bytes(allb.cpu().tolist())
    return [raw[32 * q:32 * q + 32].hex() for q in range(env.world)]


def bench_batch(env, args, scene, wname, steps, warmup, e2e_steps, trace_phases=False):
    """Stages 1-4 over `scene` on env.n_gpus GPUs: device-timed steps (L2 flushed), an instrumented pass
    for the per-stage / per-kernel numbers, the end-to-end loop with host buffers, and the result digest."""
    torch, api, sharding, dist = env.torch, env.api, env.sharding, env.dist
    dev, stream, n_gpus, rank = env.dev, env.stream, env.n_gpus, env.rank
    prm = scene.params
    l3 = api.Line3D("", False, scene.max_image_width, 3000, False, True, dev.index, stream.cuda_stream)
    l3.shard = (rank, n_gpus)
    l3.load_scene(scene)
    l3.upload()  # tables resident in HBM before the timed region
    xch = sharding.Exchanger(dist, torch, dev) if n_gpus > 1 else None
    mp = (prm["sigma_p"], prm["sigma_a"], prm["num_neighbors"], prm["epipolar_overlap"], prm["knn"], prm["const_reg_depth"])

    def step():
        if n_gpus == 1:
            l3.matchImages(*mp)
            l3.affinity()
        else:
            sharding.run_sharded(l3, xch, prm)

    sampler = ClockSampler(dev.index)
    sampler.start()
    for _ in range(warmup):
        env.flush.zero_()
        step()
    barrier(env)
    l3.reset_counters()
    total_ms = 0.0
    barrier(env)
    wall0 = time.time()
    for _ in range(steps):
        env.flush.zero_()  # flush L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(env)
        e0.record(stream)
        step()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        total_ms += e0.elapsed_time(e1)
    barrier(env)
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1)
    cnt = l3.counts()
    launches = cnt["gpu_launches"]
    total_ms = max_over_ranks(env, total_ms)
    tests_per_step = sum_over_ranks(env, cnt["pair_tests"])   # the counter holds the last step's tests
    free_b, total_b = torch.cuda.mem_get_info(dev)
    hbm_used = max_over_ranks(env, total_b - free_b)

    # ---- parity digest of the last timed step's results, on every rank ----
    t0 = time.perf_counter()
    digest = api.result_digest(l3, [v.cam_id for v in scene.views])
    digests = gather_digests(env, digest)
    digest_s = time.perf_counter() - t0

    # ---- separate instrumented pass for the per-stage / per-kernel numbers (same workload) ----
    if n_gpus == 1:
        l3.match_stage12(*mp)
        t12, c12 = l3.timings(), l3.counts()
        l3.match_stage3()
        t3 = l3.timings()
        l3.affinity()
        t4 = l3.timings()
    else:
        step()
        t12 = t3 = t4 = l3.timings()   # the stage timers accumulate over the phases of one step
        c12 = l3.counts()
    cfin = l3.counts()
    phases = None
    if trace_phases and n_gpus > 1:
        tr = {}
        for _ in range(10):
            barrier(env)
            sharding.run_sharded(l3, xch, prm, trace=tr)
        phases = {k: 1e2 * v for k, v in tr.items()}  # ms per step (10 steps)
        allp = [None] * env.world
        dist.all_gather_object(allp, phases)
        phases = allp

    # ---- end to end through the public API with host buffers (H2D of the scene, D2H of A_) ----
    h2d = scene.total_segments() * 16 + scene.num_views * (8 * 21 + 20) + sum(4 * len(v.neighbors) for v in scene.views)
    d2h = 0
    t_e2e = 0.0
    for i in range(2 + e2e_steps):
        env.flush.zero_()
        barrier(env)
        t0 = time.perf_counter()
        l3.upload()                      # host -> device: segments, cameras, neighbour lists
        step()
        ij, w = l3.edges()               # device -> host: A_ (what the CPU clustering consumes)
        l2g = l3.local2global()
        barrier(env)
        dt = time.perf_counter() - t0
        if i >= 2:
            t_e2e += dt
        d2h = ij.nbytes + w.nbytes + l2g.nbytes
    t_e2e = max_over_ranks(env, t_e2e)
    e2e = {"value": tests_per_step * e2e_steps / t_e2e, "unit": "tests/s", "h2d_bytes_per_step": int(h2d) * n_gpus,
           "d2h_bytes_per_step": int(d2h) * n_gpus, "ms_per_step": 1e3 * t_e2e / e2e_steps, "steps": e2e_steps,
           "views_per_s": scene.num_views * e2e_steps / t_e2e,
           "what": "Line3D.upload (host arrays -> HBM) + matchImages + affinity + edges()/local2global() D2H, per rank"}
    return dict(l3=l3, xch=xch, total_ms=total_ms, steps=steps, tests_per_step=tests_per_step, t12=t12, t3=t3, t4=t4,
                c12=c12, cfin=cfin, launches=launches, clocks=clocks, e2e=e2e, digest=digest, digests=digests,
                digest_s=digest_s, hbm_used=hbm_used, phases=phases, wname=wname)


def load_traffic(wname):
    """DRAM bytes per launch from the ncu --set full capture of the same workload (tools/ncu_summary.py
    writes profiles/r2_traffic.json); None when no capture exists for this workload."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json"))).get(wname, {})
    except Exception:
        return {}


def build_rooflines(env, res, scene):
    """Roofline objects of K1 (FP32 pipe) and K2 (FP64 pipe) and the figures of K3/K4; `dominant` names
    the kernel with the larger share of the step."""
    torch, api = env.torch, env.api
    t12, t3, t4, c12, cfin = res["t12"], res["t3"], res["t4"], res["c12"], res["cfin"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max = res["clocks"].get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    n_sm = torch.cuda.get_device_properties(env.dev).multi_processor_count
    pk = api.Context(env.dev.index, env.stream.cuda_stream)
    fp32_measured, fp64_measured = pk.fp32_peak_tflops(), pk.fp64_peak_tflops()
    pk.close()
    fp32_nominal = n_sm * 128 * 2 * sm_max * 1e6 / 1e12
    fp64_nominal = n_sm * 64 * 2 * sm_max * 1e6 / 1e12
    traffic = load_traffic(res["wname"])
    step_ms = res["total_ms"] / res["steps"]

    k1_s = max(t12["k1_kernel"], 1e-9) * 1e-3
    k1_tests = float(c12["pair_tests"])
    k1_launches = max(t12["k1_launches"], 1)
    a1 = FLOP_PER_TEST * k1_tests / k1_s / 1e12
    k1_run = float(c12.get("pair_tests_run", 0)) or k1_tests
    k1_issued = K1_ISSUED_PER_RUN_TEST * k1_run + K1_ISSUED_PER_WORD * k1_tests / 32.0
    seg_n = scene.views[0].segs.shape[0]
    n_pairs_local = max(c12["num_pairs_local"], 1)
    k1_bytes = n_pairs_local * (32.0 + 16.0) * seg_n + k1_tests / 8.0 + 4.0 * n_pairs_local * seg_n
    k1 = {
        "kernel": "k1_pairtest_kernel (+ k1_rowsort_kernel)", "bound": "fp32", "achieved": a1, "peak": fp32_measured, "unit": "TFLOP/s",
        "frac": a1 / fp32_measured if fp32_measured else None,
        "frac_is": "ALGORITHMIC flop (114 per segment-pair test, SURVEY 8d, over ALL tests of the launch) / measured FFMA peak.  "
                   "It exceeds 1 because most tests are decided without being evaluated: the rows of a pair are sorted by "
                   "epipolar direction and a target outside the wedge of a warp's 32 rows is skipped for the warp "
                   "(tests_run_frac of the tests are evaluated), and an evaluated test needs 45 instructions, not 114 flop.  "
                   "issue_slot_frac is the hardware-side figure: instructions the kernel issues / lane-issue capacity",
        "tests_run_frac": k1_run / k1_tests if k1_tests else None,
        "issue_slot_frac": k1_issued / k1_s / (n_sm * 128 * sm_max * 1e6),
        "issued_instructions_per_run_test": K1_ISSUED_PER_RUN_TEST, "issued_instructions_per_word": K1_ISSUED_PER_WORD,
        "peak_source": "measured here: dense FFMA micro-benchmark (l3d_bench_fp32_peak); MEASURED_PEAKS.json has no FP32 entry; "
                       "nominal %d SMs x 128 x 2 x %.0f MHz = %.1f TFLOP/s" % (n_sm, sm_max, fp32_nominal),
        "peak_nominal": fp32_nominal, "algorithmic_flop_per_test": FLOP_PER_TEST,
        "tests_per_launch": k1_tests / k1_launches, "launches_per_step": k1_launches,
        "launch_ms": 1e3 * k1_s / k1_launches, "k1_tests_per_s": k1_tests / k1_s,
        "share_of_step": 1e3 * k1_s / step_ms if res["steps"] else None,
        "traffic": traffic.get("k1_pairtest_kernel"),
        "hbm": {"algorithmic_bytes_per_launch": k1_bytes / k1_launches, "achieved_gbs": k1_bytes / k1_s / 1e9,
                "peak_gbs": peaks.get("hbm_gbs"),
                "frac": (k1_bytes / k1_s / 1e9) / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                "note": "descriptors + segments + 1 bit per test + counts; the kernel is FP32-issue bound, not HBM bound"},
    }
    k2_s = max(t12["exact"], 1e-9) * 1e-3
    a2 = K2_FLOP_PER_CANDIDATE * float(c12["candidates"]) / k2_s / 1e12
    k2 = {
        "kernel": "k2 (exact re-test + triangulation + kNN + orientation)", "bound": "fp64", "achieved": a2,
        "peak": fp64_measured, "unit": "TFLOP/s", "frac": a2 / fp64_measured if fp64_measured else None,
        "peak_source": "measured here: dense DFMA micro-benchmark (l3d_bench_fp64_peak); nominal %d SMs x 64 x 2 x %.0f MHz = %.1f TFLOP/s"
                       % (n_sm, sm_max, fp64_nominal),
        "peak_nominal": fp64_nominal, "algorithmic_flop_per_candidate": K2_FLOP_PER_CANDIDATE,
        "candidates": c12["candidates"], "stage_ms": t12["exact"], "share_of_step": t12["exact"] / step_ms,
        "traffic": traffic.get("k2"),
    }
    other = {
        "k3_score": {"bound": "fp32+sfu, dependency chain over the views",
                     "sim_evals_per_s": float(cfin["sim_evals"]) / max(t3["score"] * 1e-3, 1e-9),
                     "algorithmic_flop_per_sim_eval": 40.0,
                     "achieved_tflops": 40.0 * float(cfin["sim_evals"]) / max(t3["score"] * 1e-3, 1e-9) / 1e12,
                     "share_of_step": t3["score"] / step_ms},
        "k4_affinity": {"bound": "hbm gather", "algorithmic_bytes_per_edge_test": 160.0,
                        "edge_tests": cfin["filtered_entries"],
                        "achieved_gbs": 160.0 * float(cfin["filtered_entries"]) / max(t4["affinity"] * 1e-3, 1e-9) / 1e9,
                        "peak_gbs": peaks.get("hbm_gbs"), "share_of_step": t4["affinity"] / step_ms},
    }
    # the dominant KERNEL: K2 is five kernels per batch, the largest of them (k2_front) 53 % of the stage in the
    # ncu launch list (profiles/r2b_launches_c4.md); K1's time is its two kernels
    dominant = k1 if (1e3 * k1_s) >= 0.53 * t12["exact"] else k2
    return dominant, {"k1_pairtest": k1, "k2_exact": k2, **other}


def cpu_baseline_and_parity(env, args, scene_mod, wname):
    """Rank 0, N=1: the CPU oracle timed on a bounded cut of the workload, and the product compared with it
    bit for bit on the same cut (the oracle here is the checker and the reported baseline, never the product)."""
    import oracle_py
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from parity_utils import compare_full
    kind, nv = CPU_SAMPLE.get(wname, ("c4", 16))
    if wname in ("c4", "c4s"):
        nv = 32   # ~10-30 s of CPU work on the box's cores
    cut = scene_mod.make_scene(kind, n_views=nv)
    t0 = time.perf_counter()
    orc = oracle_py.run_scene(cut)
    tests, tm = orc.pair_tests(), orc.timers()
    tcpu = tm["match_images"] + tm["reconstruct"]
    cores = oracle_py.lib().orc_max_threads()
    base = {"value": tests / tcpu, "unit": "tests/s", "cores": cores, "kind": "port",
            "sample": "%d-view cut of the %s generator (same per-view shape), stages 1-4, one pass, %.2e tests, %.1f s"
                      % (cut.num_views, kind, tests, tcpu),
            "views_per_s_of_the_sample": cut.num_views / tcpu, "stage1_tests_per_s": tests / max(tm["match"], 1e-9)}
    parity = None
    try:
        l3 = env.api.run_scene(cut, device=env.dev.index, stream=env.stream.cuda_stream)
        sizes = compare_full(l3, orc, cut, check_scored=False)
        parity = {"vs": "CPU oracle, the cut above", "result": "bit-exact", "compared": sizes}
    except AssertionError as e:
        parity = {"vs": "CPU oracle, the cut above", "result": "MISMATCH", "detail": str(e)[:300]}
    orc.close()
    return base, parity


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-riders", action="store_true", help="N=1: skip the config-2 and config-3 measurements")
    ap.add_argument("--no-stream", action="store_true", help="skip the config-3 key-frame stream measurement")
    ap.add_argument("--check", action="store_true", help="also verify the full result against the CPU oracle (slow on c4)")
    ap.add_argument("--trace-phases", action="store_true", help="N>1: wall time per phase/exchange on every rank")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    scene_mod = importlib.import_module("3dline-slam_b200.scene")
    if args.impl == "reference":
        bench_reference(args, scene_mod)
        return
    if args.workload == "c3":
        bench_stream(args, scene_mod)
        return
    if args.workload in ("c5", "c5s"):
        c5 = importlib.import_module("3dline-slam_b200.city")
        c5.bench_city(args, setup_env(args), sys.modules[__name__])
        return

    env = setup_env(args)
    n_gpus, rank, dist = env.n_gpus, env.rank, env.dist
    scene, wname = make_workload(scene_mod, args.workload, n_gpus)
    steps = args.steps
    res = bench_batch(env, args, scene, wname, steps, args.warmup, min(steps, 20), trace_phases=args.trace_phases)
    if len(set(res["digests"])) != 1:
        raise SystemExit("parity digest differs between ranks: %r" % (res["digests"],))
    value = res["tests_per_step"] * steps / (res["total_ms"] * 1e-3)
    views_per_s = scene.num_views * steps / (res["total_ms"] * 1e-3)

    rider2 = None
    if n_gpus == 1 and wname == "c4" and not args.no_riders:
        sc2, _ = make_workload(scene_mod, "c2")
        r2 = bench_batch(env, args, sc2, "c2", max(steps, 20), max(args.warmup, 3), 20)
        dom2, roofs2 = build_rooflines(env, r2, sc2)
        ms2 = r2["total_ms"] / r2["steps"]
        rider2 = {"what": "BASELINE config[1]: 50 views x 1000 segments x 10 neighbours, 640x480, one B200, L2 flushed",
                  "value": r2["tests_per_step"] / (ms2 * 1e-3), "unit": "tests/s", "ms_per_step": ms2,
                  "views_per_s": sc2.num_views / (ms2 * 1e-3), "steps": r2["steps"],
                  "stage_ms": {"prep": r2["t12"]["prep"], "k1_pairtest": r2["t12"]["pairtest"], "k2_exact": r2["t12"]["exact"],
                               "k3_score": r2["t3"]["score"], "k4_affinity": r2["t4"]["affinity"]},
                  "e2e": r2["e2e"], "roofline": dom2, "parity_digest": r2["digest"], "gpu_launches": r2["launches"]}
        r2["l3"].ctx.close()

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    dominant, roofs = build_rooflines(env, res, scene)
    cpu_baseline = parity = None
    if not args.no_cpu_baseline and n_gpus == 1:   # reported on rank 0 at N=1 only
        cpu_baseline, parity = cpu_baseline_and_parity(env, args, scene_mod, wname)

    if args.check:
        import oracle_py
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from parity_utils import compare_full
        if n_gpus == 1:
            res["l3"].reconstruct3Dlines()
            orc = oracle_py.run_scene(scene)
            print("check vs oracle:", compare_full(res["l3"], orc, scene, check_scored=False), file=sys.stderr)
        else:
            print("--check is a single-GPU option; at N > 1 compare parity_digest with the N = 1 run", file=sys.stderr)

    t12, t3, t4, cfin = res["t12"], res["t3"], res["t4"], res["cfin"]
    line = {
        "metric": "segment_pair_tests_per_s", "value": value, "unit": "tests/s", "n_gpus": n_gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": res["total_ms"] / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32 pre-filter + f64 exact", "data": "synthetic",
        "config": workload_config(scene_mod, wname),
        "l2_policy": "flushed between timed iterations (256 MB write)",
        "parallelism": ("1 rank" if n_gpus == 1 else
                        "%d contiguous view slices of the same scene: matching, scoring rows, hypotheses and edges per slice; "
                        "NCCL all-gathers make the results whole on every rank" % n_gpus),
        "views_per_s": views_per_s,
        "stage1_tests_per_s": float(res["c12"]["pair_tests"]) / max((t12["pairtest"] + t12["exact"]) * 1e-3, 1e-9),
        "stage_ms": {"prep": t12["prep"], "k1_pairtest": t12["pairtest"], "k2_exact": t12["exact"],
                     "k3_score": t3["score"], "k4_affinity": t4["affinity"]},
        "counts": {k: cfin[k] for k in ("pair_tests", "candidates", "forward_matches", "scored_entries", "sim_evals",
                                         "filtered_entries", "num_pairs", "num_entries", "num_edges", "num_local_ids")},
        "parity_digest": res["digest"],
        "parity_digest_what": "sha256 of filtered lists + hypotheses + A_ + local2global_ after the last timed step; equal on all %d rank(s)" % n_gpus,
        "hbm_used_bytes_max_rank": res["hbm_used"],
        "roofline": dominant, "rooflines": roofs, "cpu_baseline": cpu_baseline, "parity_vs_oracle": parity,
        "e2e": res["e2e"], "gpu_launches": res["launches"], "clocks": res["clocks"],
    }
    if res["phases"]:
        line["phase_ms_per_rank"] = res["phases"]
    if res["xch"] is not None:
        line["exchange"] = {"bytes_gathered_per_step_rank0": res["xch"].bytes_gathered // max(1, steps + args.warmup + 3 + min(steps, 20)),
                            "fallbacks": res["xch"].fallbacks}
    if rider2 is not None:
        line["config2"] = rider2
    # BASELINE config 3 (key-frame stream, incremental mode) rides along on the default single-GPU run
    if n_gpus == 1 and wname == "c4" and not args.no_riders and not args.no_stream:
        try:
            import copy
            a3 = copy.copy(args)
            a3.steps, a3.warmup = 40, 3
            a3.no_cpu_baseline = True if args.no_cpu_baseline else False
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
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
