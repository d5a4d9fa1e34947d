"""Multi-GPU plumbing: one process per GPU, reference views split into contiguous slices.

A rank matches the pairs whose source view it owns (stages 1-2), builds and finishes the scoring
rows of its views and computes their affinity edges; four exchanges (all-gather over NCCL / NVLink
on the GPU box, gloo in the CPU tests) make every result whole on every rank:

    match_stage12 -> FORWARD (+ forward records, all-to-all) -> score_build -> PROGRAMS -> score_fold
                  -> HYPOTHESES -> affinity_edges -> EDGES -> affinity_ids

FORWARD carries the per-row match counts only (every rank needs the global record numbering).  The match
records of a boundary pair are needed by ONE other rank, the owner of the pair's target view, so they travel
all-to-all (`Exchanger.forward_records`): all-gathering them sent every record to every rank and was the
largest exchange of a step (C4 on 8 GPUs: 3.9 of 26 ms).

Each exchange is an all-gather-v by padding into a buffer that persists between steps (the fold
programs are read in place by the next phase).  The first time, the blob sizes are gathered first
(8 bytes per rank) to fix the stride; afterwards the blobs describe themselves (32-byte header
written by the device), the stride of the previous step is reused and the size exchange -- and the
sender's wait for its own cursors -- disappears; a blob that does not fit makes every rank fall
back to the size exchange for that step.  The blob layouts are documented in csrc/abi.cu.
"""
from __future__ import annotations

import numpy as np

X_FORWARD, X_PROGRAMS, X_HYPOTHESES, X_EDGES = 0, 1, 2, 3
KINDS = ("forward", "programs", "hypotheses", "edges")


class Exchanger:
    """Persistent send / receive buffers of the four exchanges of one rank."""

    def __init__(self, dist, torch, device):
        self.dist, self.torch, self.device = dist, torch, device
        self.world = dist.get_world_size()
        self.on_gpu = device.type == "cuda"
        self.mine = {}
        self.all = {}
        self.stride = {}       # kind -> stride of the self-describing exchange
        self.fallbacks = 0
        self.bytes_gathered = 0

    def _buf(self, store, kind, nbytes):
        t = store.get(kind)
        if t is None or t.numel() < nbytes:
            t = self.torch.empty(int(nbytes * 1.25) + 64, dtype=self.torch.uint8, device=self.device)
            store[kind] = t
        return t

    def exchange(self, shard, kind):
        """shard offers shard_blob_size(kind) / shard_export(kind, ptr, cap, device_ptr) /
        shard_import(kind, ptr, stride, world, sizes, device_ptr) (api.Line3D does)."""
        torch, dist, world = self.torch, self.dist, self.world
        if self.on_gpu and kind in self.stride and hasattr(shard, "shard_export_hdr"):
            stride = self.stride[kind]
            mine = self._buf(self.mine, kind, stride)
            allb = self._buf(self.all, kind, stride * world)
            shard.shard_export_hdr(kind, mine.data_ptr(), stride)
            dist.all_gather_into_tensor(allb[:stride * world], mine[:stride])
            redo, sizes = shard.shard_import_hdr(kind, allb.data_ptr(), stride, world)
            self.bytes_gathered += stride * world
            if not redo:
                if int(sizes.max()) * 2 < stride:      # the blobs shrank a lot: tighten the stride
                    self.stride[kind] = self._stride_for(int(sizes.max()))
                return sizes
            self.fallbacks += 1
        nbytes = shard.shard_blob_size(kind)
        sz = torch.tensor([nbytes], dtype=torch.int64, device=self.device)
        if self.on_gpu:
            szs = torch.empty(world, dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(szs, sz)
            sizes = szs.cpu().numpy().astype(np.uint64)
        else:
            parts = [torch.zeros_like(sz) for _ in range(world)]
            dist.all_gather(parts, sz)
            sizes = np.array([int(p.item()) for p in parts], dtype=np.uint64)
        stride = (int(sizes.max()) + 31) // 32 * 32
        stride = max(stride, 32)
        mine = self._buf(self.mine, kind, stride)
        allb = self._buf(self.all, kind, stride * world)
        shard.shard_export(kind, mine.data_ptr(), stride, self.on_gpu)
        if self.on_gpu:
            dist.all_gather_into_tensor(allb[:stride * world], mine[:stride])
        else:
            parts = [torch.empty(stride, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(parts, mine[:stride])
            allb[:stride * world] = torch.cat(parts)
        shard.shard_import(kind, allb.data_ptr(), stride, world, sizes, self.on_gpu)
        self.bytes_gathered += stride * world
        self.stride[kind] = self._stride_for(int(sizes.max()))
        return sizes

    REC = 32   # bytes per forward-match record on the wire (FwdRec)

    def forward_records(self, shard):
        """All-to-all-v of the boundary pairs' match records, after the FORWARD exchange."""
        torch, dist, world = self.torch, self.dist, self.world
        send, recv = shard.shard_forward_plan()
        ns, nr = int(send.sum()) * self.REC, int(recv.sum()) * self.REC
        inp = self._buf(self.mine, "fwd_send", max(ns, 32))
        out = self._buf(self.all, "fwd_recv", max(nr, 32))
        shard.shard_forward_pack(inp.data_ptr(), ns, self.on_gpu)
        in_split = [int(x) * self.REC for x in send]
        out_split = [int(x) * self.REC for x in recv]
        if self.on_gpu:
            dist.all_to_all_single(out[:nr], inp[:ns], output_split_sizes=out_split, input_split_sizes=in_split)
        else:   # gloo has no all-to-all: point-to-point per peer
            ops, io, oo = [], 0, 0
            me = dist.get_rank()
            for q in range(world):
                if in_split[q]:
                    ops.append(dist.P2POp(dist.isend, inp[io:io + in_split[q]], q))
                if out_split[q]:
                    ops.append(dist.P2POp(dist.irecv, out[oo:oo + out_split[q]], q))
                io += in_split[q]
                oo += out_split[q]
            assert in_split[me] == 0 and out_split[me] == 0
            for w in (dist.batch_isend_irecv(ops) if ops else []):
                w.wait()
        shard.shard_forward_unpack(out.data_ptr(), nr, self.on_gpu)
        self.bytes_gathered += nr
        return ns, nr

    @staticmethod
    def _stride_for(max_payload):
        """Stride of the self-describing exchange: 25 % head room over the largest blob seen."""
        return (int(max_payload * 1.25) + 64 + 32 + 31) // 32 * 32


def run_sharded(l3, xch, params, trace=None):
    """One full pass of stages 1-4 on this rank's slice with the four exchanges.
    trace: optional dict, receives the wall time of every phase (adds a device sync per phase)."""
    p = params
    steps = [("match_stage12", lambda: l3.match_stage12(p["sigma_p"], p["sigma_a"], p["num_neighbors"],
                                                        p["epipolar_overlap"], p["knn"], p["const_reg_depth"])),
             ("x_forward", lambda: xch.exchange(l3, X_FORWARD)),
             ("x_forward_records", lambda: xch.forward_records(l3)),
             ("score_build", l3.score_build),
             ("x_programs", lambda: xch.exchange(l3, X_PROGRAMS)),
             ("score_fold", l3.score_fold),
             ("x_hypotheses", lambda: xch.exchange(l3, X_HYPOTHESES)),
             ("affinity_edges", l3.affinity_edges),
             ("x_edges", lambda: xch.exchange(l3, X_EDGES)),
             ("affinity_ids", l3.affinity_ids)]
    if trace is None:
        for _, fn in steps:
            fn()
        return
    import time
    for name, fn in steps:
        t0 = time.perf_counter()
        fn()
        if xch.on_gpu:
            xch.torch.cuda.synchronize(xch.device)
        trace[name] = trace.get(name, 0.0) + (time.perf_counter() - t0)


class LocalGroup:
    """Several shards living in one process (tests on one GPU / on the CPU): the same exchange with
    host buffers instead of a collective."""

    def __init__(self, shards, torch=None, device=None):
        self.shards = shards
        self.keep = {}
        self.torch, self.device = torch, device   # given: also exercise the self-describing exchange
        self.stride = {}
        self.fallbacks = 0

    def exchange(self, kind):
        world = len(self.shards)
        if self.torch is not None and kind in self.stride:
            stride = self.stride[kind]
            buf = self.torch.zeros(stride * world, dtype=self.torch.uint8, device=self.device)
            for r, s in enumerate(self.shards):
                s.shard_export_hdr(kind, buf.data_ptr() + r * stride, stride)
            self.keep[kind] = buf
            res = [s.shard_import_hdr(kind, buf.data_ptr(), stride, world) for s in self.shards]
            assert len({r[0] for r in res}) == 1, "ranks disagree on redo"
            if not res[0][0]:
                return res[0][1]
            self.fallbacks += 1
        sizes = np.array([s.shard_blob_size(kind) for s in self.shards], dtype=np.uint64)
        stride = max((int(sizes.max()) + 31) // 32 * 32, 32)
        buf = np.zeros(stride * world, dtype=np.uint8)
        for r, s in enumerate(self.shards):
            s.shard_export(kind, buf[r * stride:].ctypes.data, stride, False)
        self.keep[kind] = buf
        for s in self.shards:
            s.shard_import(kind, buf.ctypes.data, stride, world, sizes, False)
        self.stride[kind] = Exchanger._stride_for(int(sizes.max()))
        return sizes

    def forward_records(self):
        """The all-to-all of the forward-match records between the shards of this process (host buffers)."""
        world = len(self.shards)
        plans = [s.shard_forward_plan() for s in self.shards]
        packed = []
        for s, (send, _) in zip(self.shards, plans):
            buf = np.zeros(max(int(send.sum()) * Exchanger.REC, 1), dtype=np.uint8)
            s.shard_forward_pack(buf.ctypes.data, int(send.sum()) * Exchanger.REC, False)
            packed.append(buf)
        for d, s in enumerate(self.shards):
            parts = []
            for q in range(world):
                send = plans[q][0]
                assert int(send[d]) == int(plans[d][1][q]), "send / receive plans disagree"
                o = int(send[:d].sum()) * Exchanger.REC
                parts.append(packed[q][o:o + int(send[d]) * Exchanger.REC])
            got = np.ascontiguousarray(np.concatenate(parts)) if parts else np.zeros(0, np.uint8)
            s.shard_forward_unpack(got.ctypes.data if got.size else 0, int(got.size), False)

    def run(self, params):
        p = params
        for s in self.shards:
            s.match_stage12(p["sigma_p"], p["sigma_a"], p["num_neighbors"], p["epipolar_overlap"], p["knn"],
                            p["const_reg_depth"])
        self.exchange(X_FORWARD)
        self.forward_records()
        for s in self.shards:
            s.score_build()
        self.exchange(X_PROGRAMS)
        for s in self.shards:
            s.score_fold()
        self.exchange(X_HYPOTHESES)
        for s in self.shards:
            s.affinity_edges()
        self.exchange(X_EDGES)
        for s in self.shards:
            s.affinity_ids()
