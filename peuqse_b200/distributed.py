"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink), contiguous index ranges or
chain blocks per rank, and the three tiny collectives the sweeps / chains need.

Replaces PEUQSE/parallel_processing.py + the pickle-file rendezvous of consolidate_parallel_sampling_data
(PEUQSE/InverseProblem.py:1130-1155, 1209-1319): "keep only the best" for multistart becomes an all-gather
of (maxLogP, first index) with a lowest-index tie-break (preserving the strict '>' of IP.py:933), and the
statistics become one all-reduce(sum) of the moment accumulators plus min/max.  No data-path collective:
points and chains are independent, so nothing is exchanged while kernels run.
"""
import numpy as np
import torch

try:
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None


def world():
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(total, rank=None, size=None):
    """Contiguous global-index range [first, first+count) owned by `rank`."""
    if rank is None:
        rank, size = world()
    first = (total * rank) // size
    last = (total * (rank + 1)) // size
    return first, last - first


def _comm_device():
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def allreduce_np(a, op="sum"):
    """All-reduce a small float64 numpy array; identity when not distributed."""
    rank, size = world()
    a = np.asarray(a, dtype=np.float64)
    if size == 1:
        return a
    t = torch.from_numpy(a.copy()).to(_comm_device())
    dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}[op])
    return t.cpu().numpy()


def broadcast_np(a, src=0):
    """Rank `src`'s float64 array on every rank (same shape everywhere); identity when not distributed."""
    rank, size = world()
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if size == 1:
        return a
    t = torch.from_numpy(a.copy()).to(_comm_device())
    dist.broadcast(t, src)
    return t.cpu().numpy()


def broadcast_int(v, src=0):
    """Rank `src`'s integer (up to 63 bits) on every rank."""
    rank, size = world()
    if size == 1:
        return int(v)
    t = torch.tensor([int(v)], dtype=torch.int64, device=_comm_device())
    dist.broadcast(t, src)
    return int(t.item())


def allreduce_int_np(a):
    rank, size = world()
    a = np.asarray(a, dtype=np.int64)
    if size == 1:
        return a
    t = torch.from_numpy(a.copy()).to(_comm_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def merge_best(pairs):
    """pairs: iterable of (logp, index) with index -1 for 'none'.  Highest logp, lowest index among equals,
    NaN / -inf never selected: the result the reference's sequential loop (IP.py:824, 933-938) would give."""
    best_v, best_i = float("-inf"), -1
    for v, i in pairs:
        if i < 0 or not (v > float("-inf")):
            continue
        if v > best_v or (v == best_v and i < best_i):
            best_v, best_i = float(v), int(i)
    return best_v, best_i


def allgather_best(logp, index):
    rank, size = world()
    if size == 1:
        return merge_best([(logp, index)])
    dev = _comm_device()
    mine_v = torch.tensor([float(logp)], dtype=torch.float64, device=dev)
    mine_i = torch.tensor([int(index)], dtype=torch.int64, device=dev)
    vs = [torch.empty_like(mine_v) for _ in range(size)]
    is_ = [torch.empty_like(mine_i) for _ in range(size)]
    dist.all_gather(vs, mine_v)
    dist.all_gather(is_, mine_i)
    return merge_best([(float(v.item()), int(i.item())) for v, i in zip(vs, is_)])


def allgather_rows(t):
    """Concatenate a per-rank [n_r, ...] tensor along dim 0 over ranks (sample gather)."""
    rank, size = world()
    if size == 1:
        return t
    dev = _comm_device()
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=dev)
    ns = [torch.empty_like(n) for _ in range(size)]
    dist.all_gather(ns, n)
    ns = [int(x.item()) for x in ns]
    m = max(ns)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
    pad[: t.shape[0]] = t.to(dev)
    out = [torch.empty_like(pad) for _ in range(size)]
    dist.all_gather(out, pad)
    return torch.cat([o[:c] for o, c in zip(out, ns)], dim=0).to(t.device)


def sharded_rows(fn, x, min_rows_per_rank=1):
    """fn(rows) -> values, row by row independent: every rank evaluates its contiguous shard of x[B, ...] and the shards are
    all-gathered back into the full [B] result, identical on all ranks (ensemble slice sampling v2: the walkers
    of the active half are shared out, one all-gather per stepping-out / shrink round).  The shard sizes follow from B
    alone, so no sizes are exchanged: one all_gather_into_tensor of the padded shards."""
    rank, size = world()
    B = x.shape[0]
    if size == 1 or B < size * min_rows_per_rank:
        return fn(x)
    first, count = shard_range(B, rank, size)
    mine = fn(x[first:first + count])
    m = max(shard_range(B, r, size)[1] for r in range(size))
    dev = _comm_device()
    pad = torch.zeros((m,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=dev)
    pad[:count] = mine.to(dev)
    out = torch.empty((size * m,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=dev)
    dist.all_gather_into_tensor(out, pad)
    parts = [out[r * m:r * m + shard_range(B, r, size)[1]] for r in range(size)]
    return torch.cat(parts, dim=0).to(mine.device)


_native_comm = None


def native_comm():
    """The library's own NCCL communicator (collectives enqueued on the compute stream by libpeuqse_b200, without the
    per-call stream hand-over of torch.distributed).  Created on first use from the torch.distributed NCCL group: rank 0
    makes the 128-byte id, a broadcast distributes it.  None when not running multi-GPU on NCCL."""
    global _native_comm
    rank, size = world()
    if size == 1 or dist.get_backend() != "nccl":
        return None
    if _native_comm is None:
        import ctypes as C
        from . import _lib
        L = _lib.load()
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            _lib.check(L.peuqse_comm_unique_id(buf))
        t = torch.tensor(list(buf), dtype=torch.uint8, device=_comm_device())
        dist.broadcast(t, 0)
        idb = (C.c_ubyte * 128)(*t.cpu().tolist())
        h = C.c_void_p()
        _lib.check(L.peuqse_comm_create(idb, rank, size, C.byref(h)))
        _native_comm = h
    return _native_comm


_native_peer = None
_native_peer_failed = False


def native_peer():
    """The library's peer-memory exchange (csrc/peer.cu): every rank owns a small mailbox in its device memory, maps the
    peers' mailboxes through CUDA IPC and the sweep's collectives become one push-wait-merge kernel each.  Needs all
    ranks on one node (NVLink/NVSwitch) and world <= 16; returns None otherwise, or when IPC is not available (the NCCL
    communicator is used then).  PEUQSE_B200_NO_PEER=1 disables it."""
    global _native_peer, _native_peer_failed
    import os
    rank, size = world()
    if size == 1 or size > 16 or dist.get_backend() != "nccl" or _native_peer_failed or os.environ.get("PEUQSE_B200_NO_PEER") == "1":
        return None
    if _native_peer is None:
        import ctypes as C
        from . import _lib
        L = _lib.load()
        local = C.c_void_p()
        hbuf = (C.c_ubyte * 64)()
        ok = L.peuqse_peer_alloc(C.byref(local), hbuf) == 0
        # every rank must take the same branch: agree on success before mapping
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=_comm_device())
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        mine = torch.tensor(list(hbuf), dtype=torch.uint8, device=_comm_device())
        allh = torch.empty((size, 64), dtype=torch.uint8, device=_comm_device())
        dist.all_gather_into_tensor(allh, mine.reshape(1, 64))
        h = C.c_void_p()
        if int(flag.item()) == 1:
            raw = bytes(allh.cpu().numpy().tobytes())
            ok = L.peuqse_peer_create(local, raw, rank, size, C.byref(h)) == 0
        else:
            ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=_comm_device())
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) != 1:
            if ok:
                L.peuqse_peer_destroy(h)
            _native_peer_failed = True
            return None
        torch.cuda.synchronize()
        dist.barrier()                       # every mailbox is mapped and zeroed before the first exchange
        _native_peer = h
    return _native_peer
