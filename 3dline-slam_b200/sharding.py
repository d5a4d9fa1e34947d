"""Multi-GPU plumbing: one process per GPU, pairs sharded in contiguous blocks (slices of reference views, balanced by test count) for stages 1-2, the
forward-match lists of all shards all-gathered (NCCL over NVLink on the GPU box, gloo in the CPU
tests) and merged into the canonical layout before the scoring wavefront.

Blob layout of one shard (l3d_export_forward): `uint32 cnt[rows_pad] | FwdRec recs[total]` with
rows_pad = total rows rounded up to 8 (records 32-byte aligned); rows the shard does not own have
cnt = 0; records are in row order.
"""
from __future__ import annotations

import numpy as np

FWD_DTYPE = np.dtype([("c", "<u4"), ("overlap", "<f4"), ("d_p1", "<f4"), ("d_p2", "<f4"), ("d_q1", "<f4"),
                      ("d_q2", "<f4"), ("score", "<f4"), ("flags", "<u4")])


def exchange_forward(shard, dist, torch, device):
    """All-gather the forward-match blobs of every rank and import the merged lists.

    `shard` offers forward_blob_size() / export_forward(ptr, cap, device_ptr) /
    import_forward(ptr, stride, world, device_ptr) (api.Line3D does).  Returns the bytes gathered."""
    world = dist.get_world_size()
    on_gpu = device.type == "cuda"
    nbytes = shard.forward_blob_size()
    sz = torch.tensor([nbytes], dtype=torch.int64, device=device)
    szs = [torch.zeros_like(sz) for _ in range(world)]
    dist.all_gather(szs, sz)                      # blob sizes differ per rank: all-gather-v by padding
    stride = max(int(s.item()) for s in szs)
    stride = (stride + 31) // 32 * 32
    mine = torch.zeros(stride, dtype=torch.uint8, device=device)
    shard.export_forward(mine.data_ptr(), stride, on_gpu)
    allb = torch.empty(stride * world, dtype=torch.uint8, device=device)
    if on_gpu:
        dist.all_gather_into_tensor(allb, mine)
        torch.cuda.current_stream(device).synchronize()
    else:
        parts = [torch.empty(stride, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(parts, mine)
        allb = torch.cat(parts)
    shard.import_forward(allb.data_ptr(), stride, world, on_gpu)
    return stride * world


def merge_blobs_numpy(blobs, n_rows):
    """Reference merge of per-shard blobs (host, numpy): returns (cnt[n_rows], recs in row order).
    This is what l3d_import_forward does on the device."""
    rows_pad = (n_rows + 7) // 8 * 8
    cnts = [np.frombuffer(b, dtype=np.uint32, count=n_rows) for b in blobs]
    total = np.sum(cnts, axis=0).astype(np.uint32)
    recs = [np.frombuffer(b, dtype=FWD_DTYPE, offset=rows_pad * 4, count=int(c.sum())) for b, c in zip(blobs, cnts)]
    offs = [np.concatenate([[0], np.cumsum(c.astype(np.int64))]) for c in cnts]
    out = np.zeros(int(total.sum()), dtype=FWD_DTYPE)
    pos = 0
    for r in range(n_rows):
        for c, o, rr in zip(cnts, offs, recs):
            n = int(c[r])
            if n:
                out[pos:pos + n] = rr[o[r]:o[r] + n]
                pos += n
    return total, out
