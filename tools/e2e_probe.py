"""Where the end-to-end step spends its host time: upload / step / edges / local2global, wall clock with a device
synchronisation after each (tools only; bench.py's e2e figure is the un-instrumented loop)."""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


class Args:
    gpus = 1


def main():
    env = bench.setup_env(Args())
    scene_mod = importlib.import_module("3dline-slam_b200.scene")
    wname = sys.argv[1] if len(sys.argv) > 1 else "c4"
    scene, _ = bench.make_workload(scene_mod, wname)
    api, torch = env.api, env.torch
    prm = scene.params
    l3 = api.Line3D("", False, scene.max_image_width, 3000, False, True, env.dev.index, env.stream.cuda_stream)
    l3.shard = (0, 1)
    l3.load_scene(scene)
    mp = (prm["sigma_p"], prm["sigma_a"], prm["num_neighbors"], prm["epipolar_overlap"], prm["knn"], prm["const_reg_depth"])
    acc = {}
    for i in range(6):
        for name, fn in (("upload", l3.upload), ("matchImages", lambda: l3.matchImages(*mp)), ("affinity", l3.affinity),
                         ("edges", l3.edges), ("local2global", l3.local2global)):
            torch.cuda.synchronize(env.dev)
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize(env.dev)
            if i >= 2:
                acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3 / 4
        if i >= 2:
            t = l3.timings()
            acc["device_total"] = acc.get("device_total", 0.0) + t.get("total", 0.0) / 4
    print({k: round(v, 3) for k, v in acc.items()})
    print(l3.timings())


if __name__ == "__main__":
    main()
