"""BASELINE config 5: the city-scale scene (5000 views x 5000 segments x 20 neighbours, 1920x1080), sharded by
reference view over the GPUs of one box.  `bench.py --workload c5 --gpus 8` (full size) / `--workload c5s`
(L3D_C5_VIEWS views of the same generator, default 1000).

What is replicated and what is not: the camera and segment tables are replicated -- every rank
generates a slice of the views and the 16 B/segment table is all-gathered (400 MB at full size); matching,
scoring rows, hypotheses and affinity edges are per slice; the per-row match counts and the fold programs are
all-gathered, the match records of boundary pairs go all-to-all to the one rank that needs them.  The line carries
a capacity report: the sizes of every 32-bit index space and the HBM in use per rank.

Correctness inside the run: the result digest must agree on all ranks; rank 0 also runs a 12-view cut of the same
generator unsharded and compares it bit for bit with the CPU oracle."""
from __future__ import annotations

import importlib
import json
import os
import sys
import time

import numpy as np


def _gather_segments(env, scene_mod, V, N, nbrs, n_world):
    torch, dist = env.torch, env.dist
    world, rank = env.n_gpus, env.rank
    per = (V + world - 1) // world
    lo, hi = min(rank * per, V), min((rank + 1) * per, V)
    t0 = time.perf_counter()
    scene, segs, meds = scene_mod.make_city(V, N, nbrs, n_world, view_range=(lo, hi))
    gen_s = time.perf_counter() - t0
    if dist is None:
        return scene, gen_s
    mine = torch.zeros((per, N, 4), dtype=torch.float32, device=env.dev)
    mine[:hi - lo] = torch.from_numpy(segs).to(env.dev)
    allseg = torch.empty((world * per, N, 4), dtype=torch.float32, device=env.dev)
    dist.all_gather_into_tensor(allseg, mine)
    md = torch.zeros(per, dtype=torch.float32, device=env.dev)
    md[:hi - lo] = torch.from_numpy(meds).to(env.dev)
    allmd = torch.empty(world * per, dtype=torch.float32, device=env.dev)
    dist.all_gather_into_tensor(allmd, md)
    allseg_h = allseg.cpu().numpy()
    allmd_h = allmd.cpu().numpy()
    del allseg, mine
    for i, v in enumerate(scene.views):
        v.segs = allseg_h[i]
        v.median_depth = float(allmd_h[i])
    return scene, gen_s


def bench_city(args, env, bench_mod):
    torch, api, sharding, dist = env.torch, env.api, env.sharding, env.dist
    # memory, not time, is the limit at this scale: the 512 MB mask batches the published run used
    # (the library default of 2 GB would put four times the K2 scratch, ~ 48 GB, on every rank); read by the
    # library when it plans its first batch
    os.environ.setdefault("L3D_MASK_WORDS_LOG2", "27")
    scene_mod = importlib.import_module("3dline-slam_b200.scene")
    full = args.workload == "c5"
    V = 5000 if full else int(os.environ.get("L3D_C5_VIEWS", "1000"))
    N, nbrs, n_world = 5000, 20, 200000
    steps, warmup = min(args.steps, 5), max(1, min(args.warmup, 2))
    scene, gen_s = _gather_segments(env, scene_mod, V, N, nbrs, n_world)
    prm = scene.params
    dev, stream, n_gpus, rank = env.dev, env.stream, env.n_gpus, env.rank
    l3 = api.Line3D("", False, scene.max_image_width, 5000, False, True, dev.index, stream.cuda_stream)
    l3.shard = (rank, n_gpus)
    t0 = time.perf_counter()
    l3.load_scene(scene)
    l3.upload()
    load_s = time.perf_counter() - t0
    xch = sharding.Exchanger(dist, torch, dev) if n_gpus > 1 else None
    mp = (prm["sigma_p"], prm["sigma_a"], prm["num_neighbors"], prm["epipolar_overlap"], prm["knn"], prm["const_reg_depth"])

    def step():
        if n_gpus == 1:
            l3.matchImages(*mp)
            l3.affinity()
        else:
            sharding.run_sharded(l3, xch, prm)

    sampler = bench_mod.ClockSampler(dev.index)
    sampler.start()
    for _ in range(warmup):
        step()
    bench_mod.barrier(env)
    l3.reset_counters()
    total_ms = 0.0
    wall0 = time.time()
    for _ in range(steps):
        env.flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        bench_mod.barrier(env)
        e0.record(stream)
        step()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        total_ms += e0.elapsed_time(e1)
    bench_mod.barrier(env)
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1)
    cnt = l3.counts()
    tm = l3.timings()
    total_ms = bench_mod.max_over_ranks(env, total_ms)
    tests_per_step = bench_mod.sum_over_ranks(env, cnt["pair_tests"])
    cands = bench_mod.sum_over_ranks(env, cnt["candidates"])
    free_b, total_b = torch.cuda.mem_get_info(dev)
    hbm_used = bench_mod.max_over_ranks(env, total_b - free_b)
    phases = None
    if n_gpus > 1:
        tr = {}
        bench_mod.barrier(env)
        sharding.run_sharded(l3, xch, prm, trace=tr)
        phases = {k: 1e3 * v for k, v in tr.items()}
        allp = [None] * env.world
        dist.all_gather_object(allp, phases)
        phases = allp
    t0 = time.perf_counter()
    digest = api.result_digest(l3, [v.cam_id for v in scene.views])
    digests = bench_mod.gather_digests(env, digest)
    digest_s = time.perf_counter() - t0
    if len(set(digests)) != 1:
        raise SystemExit("parity digest differs between ranks: %r" % (digests,))
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    # ---- a cut of the same generator, unsharded, against the CPU oracle ----
    parity = None
    if not args.no_cpu_baseline:
        import oracle_py
        sys.path.insert(0, os.path.join(bench_mod.ROOT, "tests"))
        from parity_utils import compare_full
        nv = 12
        cut, csegs, cmeds = scene_mod.make_city(V, N, nbrs, n_world, view_range=(0, nv))
        cut.views = cut.views[:nv]
        for v in cut.views:
            v.neighbors = [j for j in v.neighbors if j < nv]
        t0 = time.perf_counter()
        orc = oracle_py.run_scene(cut)
        tcpu = orc.timers()["match_images"] + orc.timers()["reconstruct"]
        one = api.run_scene(cut, device=dev.index, stream=stream.cuda_stream)
        try:
            sizes = compare_full(one, orc, cut, check_scored=False)
            parity = {"vs": "CPU oracle on the first %d views of the same generator, unsharded" % nv, "result": "bit-exact",
                      "compared": sizes, "cpu_tests_per_s": orc.pair_tests() / tcpu,
                      "cpu_cores": oracle_py.lib().orc_max_threads()}
        except AssertionError as e:
            parity = {"result": "MISMATCH", "detail": str(e)[:300]}
        orc.close()
    S = scene.total_segments()
    value = tests_per_step * steps / (total_ms * 1e-3)
    u32 = float(2 ** 32)
    line = {
        "metric": "segment_pair_tests_per_s", "value": value, "unit": "tests/s", "n_gpus": n_gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32 pre-filter + f64 exact", "data": "synthetic",
        "config": {"workload": args.workload, "views": V, "segments_per_view": N, "neighbours": nbrs, "image": "1920x1080"},
        "l2_policy": "flushed between timed iterations (256 MB write)",
        "parallelism": "%d contiguous view slices; tables replicated (generated per slice, all-gathered); match counts "
                       "and fold programs all-gathered; boundary match records all-to-all" % n_gpus,
        "views_per_s": V * steps / (total_ms * 1e-3),
        "stage_ms_rank0": {"prep": tm["prep"], "k1_pairtest": tm["pairtest"], "k2_exact": tm["exact"], "k3_score": tm["score"],
                           "k4_affinity": tm["affinity"]},
        "counts": {"pair_tests": tests_per_step, "candidates": cands, "forward_matches": cnt["forward_matches"],
                   "filtered_entries": cnt["filtered_entries"], "num_pairs": cnt["num_pairs"], "num_entries": cnt["num_entries"],
                   "num_edges": cnt["num_edges"], "num_local_ids": cnt["num_local_ids"]},
        "capacity": {
            "what": "every index space of the path is 32-bit per GPU; fraction of 2^32 in use",
            "segments": S / u32, "pair_rows": (cnt["num_pairs"] * N) / u32, "forward_matches": cnt["forward_matches"] / u32,
            "edges": cnt["num_edges"] / u32, "hbm_used_bytes_max_rank": hbm_used, "hbm_total_bytes": float(total_b),
            "segment_table_bytes_replicated": S * 16,
        },
        "phase_ms_per_rank": phases,
        "exchange": {"bytes_per_step_rank0": (xch.bytes_gathered // (steps + warmup + 1)) if xch else 0,
                     "fallbacks": xch.fallbacks if xch else 0},
        "setup_s": {"generate_slice": gen_s, "load_and_upload": load_s, "digest": digest_s},
        "parity_digest": digest, "parity_vs_oracle": parity, "gpu_launches": cnt["gpu_launches"], "clocks": clocks,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
