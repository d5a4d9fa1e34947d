"""CPU, world_size 2 over gloo: the host-side exchange logic of a sharded run (sizes all-gather,
all-gather-v by padding, import with every rank's blob and size), with fake shards standing in for
the GPU contexts."""
import ctypes
import importlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

shd = importlib.import_module("3dline-slam_b200.sharding")


class FakeShard:
    """Stands in for api.Line3D on a box without a GPU: every exchange kind carries a blob whose
    length and content depend on (rank, kind); import checks that all blobs arrive intact."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world
        self.seen = {}

    @staticmethod
    def blob(rank, kind):
        rng = np.random.default_rng(1000 * kind + rank)
        return rng.integers(0, 256, size=37 + 101 * rank + 13 * kind, dtype=np.uint8)

    def shard_blob_size(self, kind):
        return int(self.blob(self.rank, kind).size)

    def shard_export(self, kind, ptr, cap, device_ptr):
        b = self.blob(self.rank, kind)
        assert not device_ptr and cap >= b.size
        np.ctypeslib.as_array((ctypes.c_uint8 * cap).from_address(ptr))[:b.size] = b

    def shard_import(self, kind, ptr, stride, world, sizes, device_ptr):
        assert not device_ptr and world == self.world and stride % 32 == 0
        raw = np.ctypeslib.as_array((ctypes.c_uint8 * (stride * world)).from_address(ptr))
        ok = True
        for q in range(world):
            b = self.blob(q, kind)
            ok &= int(sizes[q]) == b.size and bool((raw[q * stride:q * stride + b.size] == b).all())
        self.seen[kind] = ok

    # forward records, all-to-all: rank r sends (3 + 2 r + d) records of 32 bytes to rank d, filled with (r, d)
    def _n(self, src, dst):
        return 0 if src == dst else 3 + 2 * src + dst

    def shard_forward_plan(self):
        send = np.array([self._n(self.rank, d) for d in range(self.world)], dtype=np.uint64)
        recv = np.array([self._n(q, self.rank) for q in range(self.world)], dtype=np.uint64)
        return send, recv

    def shard_forward_pack(self, ptr, cap_bytes, device_ptr):
        send, _ = self.shard_forward_plan()
        assert not device_ptr and cap_bytes == int(send.sum()) * 32
        raw = np.ctypeslib.as_array((ctypes.c_uint8 * max(cap_bytes, 1)).from_address(ptr))
        o = 0
        for d in range(self.world):
            n = int(send[d]) * 32
            raw[o:o + n] = (16 * self.rank + d) & 255
            o += n

    def shard_forward_unpack(self, ptr, nbytes, device_ptr):
        _, recv = self.shard_forward_plan()
        assert not device_ptr and nbytes == int(recv.sum()) * 32
        raw = np.ctypeslib.as_array((ctypes.c_uint8 * max(nbytes, 1)).from_address(ptr))
        o, ok = 0, True
        for q in range(self.world):
            n = int(recv[q]) * 32
            ok &= bool((raw[o:o + n] == ((16 * q + self.rank) & 255)).all())
            o += n
        self.seen["fwd"] = ok


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = FakeShard(rank, world)
    xch = shd.Exchanger(dist, torch, torch.device("cpu"))
    for rep in range(2):                    # buffers persist and are reused
        for kind in (shd.X_FORWARD, shd.X_PROGRAMS, shd.X_HYPOTHESES, shd.X_EDGES):
            xch.exchange(s, kind)
            if kind == shd.X_FORWARD:
                xch.forward_records(s)      # all-to-all-v (point-to-point under gloo)
    q.put((rank, all(s.seen.get(k, False) for k in (0, 1, 2, 3, "fwd")), xch.bytes_gathered))
    dist.destroy_process_group()


def test_exchanger_world2_gloo():
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
    assert res[0][2] > 0 and res[1][2] > 0     # (the all-to-all part differs per rank by construction)


def test_local_group_uses_the_same_protocol():
    shards = [FakeShard(r, 3) for r in range(3)]
    g = shd.LocalGroup(shards)
    for kind in range(4):
        g.exchange(kind)
        if kind == shd.X_FORWARD:
            g.forward_records()
    assert all(all(s.seen[k] for k in (0, 1, 2, 3, "fwd")) for s in shards)
