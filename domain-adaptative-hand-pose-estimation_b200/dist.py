"""Batch sharding across the GPUs of one box + the single collective of the hot path.

Every stage is independent per sample, so ranks own contiguous slices of the batch
and exchange nothing until the end: ONE all-reduce(sum) of the int64 partial vector
``[mse_fx, kl_fx, n_maps, n_elems, hits[K], valid[K], 6 non-finite counters]`` (4+2K+6 int64 = 416 B
at K=21; the loss sums are fixed point, value * 2**40, so the sum over ranks is exact and order
independent), after which every rank finalises exactly what ``accuracy()`` / ``loss.mean()`` would
give on the concatenated batch.  NCCL over NVLink on the
GPUs; the same code runs on ``gloo`` for the CPU tests of the host logic.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of ``total`` samples owned by ``rank`` (sizes differ by at most 1)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def is_distributed(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def allreduce_partial(partial: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum of the partial vector over the ranks of ``group`` (no-op when single-process)."""
    if is_distributed(group):
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


class PeerExchange:
    """The path's single collective without NCCL on the per-step path: every rank owns a mailbox in device
    memory (``hp_peer_alloc``), exported with CUDA IPC and mapped by all ranks of the node; per step ONE small
    kernel (``hp_pipeline_finalize_peer``) stores this rank's int64 partial vector into every mailbox over
    NVLink, waits (bounded) for the other ranks' vectors of the same step, sums them in rank order and
    finalises - all ranks get bit-identical results, with no collective launch latency on the host.

    Setup uses ``torch.distributed`` once (all-gather of the 64-byte IPC handles).  Single node only."""

    def __init__(self, device, group=None):
        import ctypes as C
        from . import _lib
        self._lib, self._C = _lib, C
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.device(device)
        self._mapped = []
        with _lib.on_device(self.device):
            own = C.c_void_p()
            _lib.call("hp_peer_alloc", self.world, C.byref(own))
            self._own = own
            handle = (C.c_ubyte * 64)()
            _lib.call("hp_peer_export", own, handle)
            mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=self.device)
            every = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(every, mine, group=group)
            ptrs = []
            for r, h in enumerate(every):
                if r == self.rank:
                    ptrs.append(own.value)
                    continue
                raw = (C.c_ubyte * 64)(*h.cpu().tolist())
                mapped = C.c_void_p()
                _lib.call("hp_peer_import", raw, C.byref(mapped))
                self._mapped.append(mapped)
                ptrs.append(mapped.value)
            self._table = (C.c_void_p * self.world)(*ptrs)
            torch.cuda.synchronize(self.device)
        dist.barrier(group=group)          # every mailbox is mapped (and zeroed) before the first step

    def finalize(self, partial, K, result, partial_out=None, stream=None):
        """Enqueue the exchange + finalise of this step on ``stream`` (default: current)."""
        C, _lib = self._C, self._lib
        st = C.c_void_p(stream.cuda_stream) if stream is not None else _lib.stream_ptr(self.device)
        _lib.call("hp_pipeline_finalize_peer", _lib.ptr(partial), self._table, self.rank, self.world, int(K),
                  C.c_int64(0), _lib.ptr(partial_out), _lib.ptr(result), st)      # 0: step counted on the device

    def flush(self, workspace, stream=None):
        """Complete the outstanding step of a train of deferred exchanges that used ``workspace``."""
        C, _lib = self._C, self._lib
        st = C.c_void_p(stream.cuda_stream) if stream is not None else _lib.stream_ptr(self.device)
        _lib.call("hp_pipeline_flush_peer", _lib.ptr(workspace), self._table, self.rank, self.world, st)

    def pck_finalize(self, counts, K, counts_out, acc_out, stream=None):
        """configs[3]: integer PCK counts int32 [2K] of this rank -> totals + accuracies, one small kernel."""
        C, _lib = self._C, self._lib
        st = C.c_void_p(stream.cuda_stream) if stream is not None else _lib.stream_ptr(self.device)
        _lib.call("hp_pck_finalize_peer", _lib.ptr(counts), self._table, self.rank, self.world, int(K),
                  _lib.ptr(counts_out), _lib.ptr(acc_out), st)

    def fuse_decode_pck(self, lo, a_lo, mid, a_mid, hi, a_hi, tgt, B, K, H, W, thr, pred_xy, maxvals, counts, acc, ws,
                        stream=None, partial=None, result=None):
        """configs[3], sharded: fuse + decode + PCK of this rank's samples with the counts exchange folded into the kernel's
        last block (one launch per step; other geometries: the one-warp exchange kernel follows inside the C call).
        ``partial`` / ``result`` given: a DEFERRED step (HP_PIPE_DEFER_EXCHANGE) - see :meth:`MultiscaleEval.step`."""
        C, _lib = self._C, self._lib
        st = C.c_void_p(stream.cuda_stream) if stream is not None else _lib.stream_ptr(self.device)
        _lib.call("hp_fuse_decode_pck_peer", _lib.ptr(lo), lo.shape[2], lo.shape[3], C.c_float(a_lo), _lib.ptr(mid),
                  mid.shape[2], mid.shape[3], C.c_float(a_mid), _lib.ptr(hi), C.c_float(a_hi), _lib.ptr(tgt), B, K, H, W,
                  C.c_double(thr), _lib.ptr(pred_xy), _lib.ptr(maxvals), _lib.ptr(counts), _lib.ptr(acc), _lib.ptr(ws),
                  self._table, self.rank, self.world, 2 if partial is not None else 0, _lib.ptr(partial), _lib.ptr(result), st)

    def close(self):
        _SHARED.pop((self.device.index, id(self.group)), None)
        if self._own is None:
            return
        try:
            torch.cuda.synchronize(self.device)
            if dist.is_initialized():
                dist.barrier(group=self.group)     # nobody unmaps while a peer may still write
            with self._lib.on_device(self.device):
                for m in self._mapped:
                    self._lib.call("hp_peer_close", m)
                self._lib.call("hp_peer_free", self._own)
        finally:
            self._mapped, self._own = [], None


_SHARED = {}


def shared_peer_exchange(device, group=None) -> PeerExchange:
    """One mailbox per (device, process group), shared by everything that exchanges on it (the step number is counted
    in the mailbox, so all users must issue their exchanges in the same order on every rank - SPMD)."""
    device = torch.device(device)
    key = (device.index if device.index is not None else torch.cuda.current_device(), id(group))
    px = _SHARED.get(key)
    if px is None:
        px = _SHARED[key] = PeerExchange(device, group)
    return px


FX_SHIFT = 40                    # HP_LOSS_FX_SHIFT of include/hp_b200.h
FX_LIMIT = 2097152.0             # |per-map loss| >= 2**21 counts as infinite


def partial_len(K: int) -> int:
    """HP_PARTIAL_LEN(K)."""
    return 4 + 2 * K + 6


def loss_to_fx(values):
    """Per-map float losses -> (fixed-point int64 sum, [n_nan, n_pinf, n_ninf]) exactly like the kernels."""
    v = np.asarray(values, dtype=np.float64).ravel()
    nan = np.isnan(v)
    pinf = ~nan & (v >= FX_LIMIT)
    ninf = ~nan & (v <= -FX_LIMIT)
    fin = ~(nan | pinf | ninf)
    fx = np.rint(np.ldexp(v[fin], FX_SHIFT)).astype(np.int64).sum()
    return int(fx), [int(nan.sum()), int(pinf.sum()), int(ninf.sum())]


def _loss_from_fx(fx, n_nan, n_pinf, n_ninf, n_maps):
    if n_nan or (n_pinf and n_ninf):
        return float("nan")
    if n_pinf:
        return float("inf")
    if n_ninf:
        return float("-inf")
    return float(np.ldexp(np.float64(fx), -FX_SHIFT) / np.float64(n_maps))


def finalize_partial_host(partial, K: int):
    """Host mirror of ``hp_pipeline_finalize`` for an int64 partial vector already on the host:
    -> dict(mse, kl, avg_acc, cnt, acc[K], hits, valid).  Same order of float64 operations as
    utils/keypoint_detection.py:53-60, 80-90."""
    p = np.asarray(partial, dtype=np.int64)
    hits = p[4:4 + K].astype(np.int64)
    valid = p[4 + K:4 + 2 * K].astype(np.int64)
    cls = p[4 + 2 * K:4 + 2 * K + 6]
    acc = np.full(K, -1.0)
    total, cnt = 0.0, 0
    for k in range(K):
        if valid[k] > 0:
            acc[k] = hits[k] * 1.0 / valid[k]
        if acc[k] >= 0:
            total = total + acc[k]
            cnt += 1
    return dict(mse=_loss_from_fx(p[0], cls[0], cls[1], cls[2], p[2]), kl=_loss_from_fx(p[1], cls[3], cls[4], cls[5], p[2]),
                avg_acc=(total / cnt if cnt != 0 else 0.0), cnt=cnt, acc=acc, hits=hits, valid=valid)


# ------------------------------------------------------------------------------------------
# host placement: pinned staging buffers next to the GPU
# ------------------------------------------------------------------------------------------

def gpu_numa_node(device_index: int):
    """NUMA node of the PCIe root the GPU hangs off (``/sys/bus/pci/devices/<bus id>/numa_node``), or None."""
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(device_index: int):
    """Pin THIS process to the CPUs of the GPU's NUMA node (one process per GPU).  Host buffers allocated afterwards -
    the pinned staging memory of ``HeatmapPipeline.run_host`` in particular - are first-touched on that node, so the
    H2D DMA does not cross the socket interconnect (eight ranks pinning on one node halved the end-to-end rate at
    N = 8 in round 1).  Returns the node, or None when the topology is not exposed (nothing is changed then)."""
    import os
    node = gpu_numa_node(device_index)
    if node is None:
        return None
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _cpulist(f.read())
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None
