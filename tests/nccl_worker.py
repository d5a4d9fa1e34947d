"""Worker of tests/test_nccl_gpu.py: one process per GPU (torchrun), real NCCL all-gathers.  Every rank runs
the sharded path on its slice, compares the whole result with the CPU oracle and prints its digest; rank 0
also runs the unsharded path on its GPU and prints that digest."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    import oracle_py
    from parity_utils import compare_full
    api = importlib.import_module("3dline-slam_b200.api")
    scene_mod = importlib.import_module("3dline-slam_b200.scene")
    sharding = importlib.import_module("3dline-slam_b200.sharding")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    kind, nv, nseg = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    scene = scene_mod.make_scene(kind, n_views=nv, n_seg=nseg)
    views = [v.cam_id for v in scene.views]
    stream = torch.cuda.current_stream(dev)
    l3 = api.Line3D("", False, scene.max_image_width, 3000, False, True, dev.index, stream.cuda_stream)
    l3.shard = (rank, world)
    l3.load_scene(scene)
    l3.upload()
    xch = sharding.Exchanger(dist, torch, dev)
    out = {"rank": rank, "world": world}
    for it in range(3):   # first step: size exchange; afterwards the self-describing blobs
        sharding.run_sharded(l3, xch, scene.params)
        torch.cuda.synchronize(dev)
        out["digest_step%d" % it] = api.result_digest(l3, views)
    out["fallbacks"] = xch.fallbacks
    l3._ck(l3.L.l3d_cluster(l3.h))
    orc = oracle_py.run_scene(scene)
    out["sizes"] = compare_full(l3, orc, scene, check_scored=False)   # raises on any difference
    orc.close()
    if rank == 0:
        one = api.run_scene(scene, device=dev.index, stream=stream.cuda_stream)
        out["digest_unsharded"] = api.result_digest(one, views)
    print("NCCLWORKER " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
