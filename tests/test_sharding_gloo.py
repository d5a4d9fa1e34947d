"""CPU, world_size 2 over gloo: the host-side exchange of the per-shard forward-match lists
(all-gather-v by padding + merge into the canonical row order)."""
import importlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

shd = importlib.import_module("3dline-slam_b200.sharding")


class FakeShard:
    """Stands in for api.Line3D on a box without a GPU: owns rows r with (r // 3) % world == rank."""

    def __init__(self, rank, world, n_rows=37):
        self.rank, self.world, self.n_rows = rank, world, n_rows
        rng = np.random.default_rng(123)            # same stream on every rank: the "global truth"
        self.cnt_all = rng.integers(0, 4, size=n_rows).astype(np.uint32)
        self.recs_all = np.zeros(int(self.cnt_all.sum()), dtype=shd.FWD_DTYPE)
        self.recs_all["c"] = np.arange(len(self.recs_all))
        self.recs_all["overlap"] = rng.random(len(self.recs_all)).astype(np.float32)
        own = (np.arange(n_rows) // 3) % world == rank
        self.cnt = np.where(own, self.cnt_all, 0).astype(np.uint32)
        off = np.concatenate([[0], np.cumsum(self.cnt_all.astype(np.int64))])
        self.recs = np.concatenate([self.recs_all[off[r]:off[r + 1]] for r in range(n_rows) if own[r]] or
                                   [np.zeros(0, dtype=shd.FWD_DTYPE)])
        self.merged = None

    def forward_blob_size(self):
        return (self.n_rows + 7) // 8 * 8 * 4 + self.recs.nbytes

    def export_forward(self, ptr, cap, device_ptr):
        assert not device_ptr and cap >= self.forward_blob_size()
        buf = (np.ctypeslib.as_array((__import__("ctypes").c_uint8 * cap).from_address(ptr)))
        rows_pad = (self.n_rows + 7) // 8 * 8
        head = np.zeros(rows_pad, dtype=np.uint32)
        head[:self.n_rows] = self.cnt
        blob = head.tobytes() + self.recs.tobytes()
        buf[:len(blob)] = np.frombuffer(blob, dtype=np.uint8)

    def import_forward(self, ptr, stride, world, device_ptr):
        assert not device_ptr
        raw = bytes((__import__("ctypes").c_uint8 * (stride * world)).from_address(ptr))
        blobs = [raw[i * stride:(i + 1) * stride] for i in range(world)]
        self.merged = shd.merge_blobs_numpy(blobs, self.n_rows)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = FakeShard(rank, world)
    shd.exchange_forward(s, dist, torch, torch.device("cpu"))
    cnt, recs = s.merged
    ok = bool((cnt == s.cnt_all).all() and recs.tobytes() == s.recs_all.tobytes())
    q.put((rank, ok, int(cnt.sum())))
    dist.destroy_process_group()


def test_exchange_forward_world2_gloo():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), res
    assert res[0][2] == res[1][2] > 0
