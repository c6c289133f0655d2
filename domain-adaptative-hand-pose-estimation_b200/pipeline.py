"""The fused hot path as one object: Gaussian target generation + JointsMSELoss + JointsKLLoss +
decode + PCK in a single pass over the prediction tensor (BASELINE.json metric), and the multiscale
fuse + decode + PCK evaluation of configs[3].

What it replaces per batch in the reference (``validate()``, ``train1.py:495-536``, plus the
dataset-side ``generate_target``)::

    target, weight = zip(*[generate_target(j, v, (W,H), sigma, image) for each sample])   # util.py:9-68
    mse  = JointsMSELoss()(y, target, weight)                                            # loss.py:55-65
    kl   = JointsKLLoss(epsilon=eps)(y, target, weight)                                  # loss.py:145-158
    acc, avg_acc, cnt, pred = accuracy(y.cpu().numpy(), target.cpu().numpy())            # keypoint_detection.py:63-92
"""
from __future__ import annotations

import collections
import ctypes as C
import weakref
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from . import dist as hpdist

_LOSS_BITS = {"mse": _lib.LOSS_MSE, "kl": _lib.LOSS_KL}


@dataclass
class PipelineResult:
    """Device-resident outputs; ``host()`` does the one small device->host read."""
    result: torch.Tensor      # float64 [4+K] = mse, kl, avg_acc, cnt, acc[K]
    partial: torch.Tensor     # int64 [4+2K+6] = mse_fx, kl_fx, n_maps, n_elems, hits[K], valid[K], 6 non-finite counters
    #                           (loss sums are fixed point, value * 2**40: exact under any schedule / sharding)
    pred_xy: torch.Tensor     # float32 [B,K,2]
    maxvals: torch.Tensor     # float32 [B,K,1]
    weight: torch.Tensor      # float32 [B,K,1]  (target_weight of generate_target)
    K: int
    ready: object = None      # CUDA event recorded after the last op that writes `result` (sharded path)
    owner: object = None      # weakref to the pipeline while this step's totals are still outstanding (deferred exchange)

    def wait(self):
        """Make the current stream wait for `result`: completes a deferred exchange that still has this step
        outstanding, and orders the stream after a collective that ran on a side stream."""
        pipe = self.owner() if self.owner is not None else None
        if pipe is not None and pipe._pending is self:
            pipe.join()
        if self.ready is not None:
            torch.cuda.current_stream(self.result.device).wait_event(self.ready)
        return self

    def host(self):
        self.wait()
        r = self.result.cpu().numpy() if isinstance(self.result, torch.Tensor) else np.asarray(self.result)
        K = self.K
        if np.isnan(r[3]):          # cnt is never NaN in a completed step: the kernels poison it on an exchange timeout
            raise RuntimeError("HeatmapPipeline: the cross-GPU exchange of this step timed out (a rank never delivered "
                               "its partial vector); the totals are incomplete")
        return dict(mse=float(r[0]), kl=float(r[1]), avg_acc=float(r[2]) if int(r[3]) else 0, cnt=int(r[3]),
                    acc=r[4:4 + K].copy())


class _Plan:
    """Owner of one C-side plan handle (``hp_plan_t``) and of the library-side tensors whose pointers it stores
    (outputs, workspace, table, mailboxes).  ``keep`` holds the caller's input tensors only for plans made through
    the explicit :meth:`HeatmapPipeline.plan`; the implicit per-call cache never pins inputs."""

    def __init__(self, handle, keep):
        self.handle, self.keep = handle, keep

    def __del__(self):
        try:
            if self.handle:
                _lib.load().hp_pipeline_plan_destroy(self.handle)
        except Exception:  # interpreter shutdown
            pass
        self.handle = None


class HeatmapPipeline:
    """gen + loss + decode + PCK, one kernel launch per batch (one more tiny one after the
    all-reduce when the batch is sharded over several GPUs)."""

    def __init__(self, num_keypoints=21, heatmap_size=(64, 64), image_size=(256, 256), sigma=2, kl_epsilon=0.0,
                 thr=0.5, losses=("mse", "kl"), device=None, group=None, collective="nccl", defer_exchange=True):
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: the B200 heatmap path has no CPU fallback")
        self.K = int(num_keypoints)
        if self.K > _lib.MAX_K:
            raise ValueError(f"num_keypoints={self.K} exceeds HP_MAX_K={_lib.MAX_K}")
        self.W, self.H = int(heatmap_size[0]), int(heatmap_size[1])
        self.sigma = sigma
        self.tmp = _lib.integer_tmp(sigma * 3)
        stride = np.array(image_size) / np.array(heatmap_size)          # util.py:36
        self.stride = (float(stride[0]), float(stride[1]))
        self.kl_epsilon = float(kl_epsilon)
        self.thr = float(thr)
        self.loss_mask = 0
        for name in losses:
            self.loss_mask |= _LOSS_BITS[name]
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.group = group
        with _lib.on_device(self.device):
            self.tab = _lib.gaussian_table(sigma, self.tmp, self.device)
        self._host_state = None
        self._ws = None
        self._plans = collections.OrderedDict()     # small LRU of pre-bound launches (see _cached_plan)
        self._comm_stream = None
        if collective not in ("nccl", "peer"):
            raise ValueError("collective must be 'nccl' or 'peer'")
        self.collective = collective      # how the sharded path sums the partial vectors (dist.PeerExchange / NCCL)
        if collective == "peer" and hpdist.is_distributed(group) and self.K > _lib.PEER_MAX_K:
            raise ValueError(f"num_keypoints={self.K} exceeds what the peer mailboxes carry (K <= {_lib.PEER_MAX_K}); "
                             "use collective='nccl'")
        self._peer = None
        # collective="peer": overlapped steps (a train) only SEND their partial vector; the totals of step s are
        # completed by step s+1's kernel - or by join() / PipelineResult.wait() / host() for the last one - so no step
        # waits for a peer that is less than a whole step late (HP_PIPE_DEFER_EXCHANGE).  Serialised steps exchange at once.
        self.defer_exchange = bool(defer_exchange)
        self._pending = None              # the PipelineResult whose totals are still outstanding
        _lib.load()

    # ------------------------------------------------------------------------------ device path
    def _check(self, pred, joints, vis):
        """-> contiguous CUDA (pred fp32, joints fp64, vis fp32) of the pipeline's shape; converts fp16/bf16 and
        non-contiguous views (fresh copies: such calls are never cached)."""
        pred = _lib.require_cuda(pred, "HeatmapPipeline(pred)")
        joints = _lib.require_cuda(joints, "HeatmapPipeline(joints)", torch.float64)
        vis = _lib.require_cuda(vis, "HeatmapPipeline(vis)")
        B, K, H, W = pred.shape
        if (K, H, W) != (self.K, self.H, self.W):
            raise ValueError(f"pred is {tuple(pred.shape)}, pipeline was built for K={self.K} H={self.H} W={self.W}")
        if joints.numel() != 2 * B * K or vis.numel() != B * K:
            raise ValueError("joints must be [B,K,2] and vis [B,K,1]")
        if pred.device != self.device:
            raise ValueError(f"inputs are on {pred.device}, the pipeline was built for {self.device}")
        return pred, joints, vis

    def plan(self, pred, joints, vis, out=None, finalize=True, overlap=False):
        """Validate once and return ``(launch, out)``: ``launch()`` enqueues the fused kernel on the
        current stream with pre-bound arguments (a ~15 us kernel leaves no room for per-call Python
        argument checking; this is also what gets captured into CUDA graphs).  The plan keeps the three
        input tensors alive; write new batches INTO them (``pred.copy_(...)``) between launches.

        ``overlap`` (HP_PIPE_OVERLAP_PREV): ``True`` or a depth 1..8 - the launch is part of a train of
        launches over batches that are already resident and may run concurrently with the previous
        ``depth - 1`` launches on the stream (it writes nothing before they have completed; results are
        bit-identical).  ``pred``, ``joints`` and ``vis`` must not be written by the kernel launched right
        before this one, and consecutive launches need different ``out`` buffers."""
        pred, joints, vis = self._check(pred, joints, vis)
        if out is None:
            out = self.alloc_outputs(pred.shape[0], pred.device)
        ws = self._workspace(pred.shape[0] * self.K)
        return self._bind(pred, joints, vis, out, ws, finalize, overlap, None, True), out

    def _bind(self, pred, joints, vis, out, ws, finalize, overlap, peer, keep_inputs, defer=False):
        """Arguments validated and stored once on the C side (``hp_pipeline_plan_create``); the returned ``launch()``
        is a two-argument FFI call (a step is a ~12 us kernel: per-call marshalling of 25 arguments costs as much)."""
        lib = _lib.load()
        B, K, H, W = pred.shape
        dev = pred.device
        handle = C.c_void_p()
        _lib.call("hp_pipeline_plan_create", _lib.ptr(pred), _lib.ptr(joints), _lib.ptr(vis), B, K, H, W,
                  C.c_double(self.stride[0]), C.c_double(self.stride[1]), self.tmp, _lib.ptr(self.tab),
                  C.c_float(self.kl_epsilon), C.c_double(self.thr), self.loss_mask, _lib.ptr(out.pred_xy),
                  _lib.ptr(out.maxvals), _lib.ptr(out.weight), _lib.ptr(out.partial), 0,
                  _lib.ptr(out.result) if finalize else None, _lib.ptr(ws),
                  peer._table if peer is not None else None, peer.rank if peer is not None else 0,
                  peer.world if peer is not None else 1, C.c_uint(_lib.pipe_flags(overlap, defer)), C.byref(handle))
        # the plan owns references to what the LIBRARY side allocated; the caller's inputs only on request
        plan = _Plan(handle, (out, ws, peer, self.tab) + ((pred, joints, vis) if keep_inputs else ()))
        fn = lib.hp_pipeline_plan_launch
        raw_stream = torch._C._cuda_getCurrentRawStream
        dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        last_error = lib.hp_last_error

        if peer is None:
            def launch(_plan=plan, _h=handle):
                rc = fn(_h, raw_stream(dev_index))
                if rc != 0:
                    raise RuntimeError(f"hp_pipeline_plan_launch failed (rc={rc}): {last_error().decode(errors='replace')}")
            return launch
        owner = weakref.ref(self)
        pending = out if defer else None     # a sharded step completes whatever was outstanding; a deferred one leaves itself

        def launch_sharded(_plan=plan, _h=handle):
            rc = fn(_h, raw_stream(dev_index))
            if rc != 0:
                raise RuntimeError(f"hp_pipeline_plan_launch failed (rc={rc}): {last_error().decode(errors='replace')}")
            self._pending = pending
            if pending is not None:
                pending.owner = owner
        return launch_sharded

    def plan_peer(self, pred, joints, vis, out=None, overlap=False):
        """Sharded step with the peer-memory exchange as ONE call (``hp_pipeline_fused_peer``): the fused kernel on
        this rank's slice, then the exchange + finalise kernel, both on the current stream."""
        pred, joints, vis = self._check(pred, joints, vis)
        if out is None:
            out = self.alloc_outputs(pred.shape[0], pred.device)
        ws = self._workspace(pred.shape[0] * self.K)
        return self._bind(pred, joints, vis, out, ws, True, overlap, self._peer_link(), True,
                          self.defer_exchange and bool(overlap)), out

    def _peer_link(self):
        if self._peer is None:
            if self.K > _lib.PEER_MAX_K:
                raise ValueError(f"num_keypoints={self.K} exceeds what the peer mailboxes carry "
                                 f"(K <= {_lib.PEER_MAX_K}); use collective='nccl'")
            self._peer = hpdist.shared_peer_exchange(self.device, self.group)
        return self._peer

    def _workspace(self, n_maps):
        """One zero-initialised workspace per pipeline object (a pipeline is used on one stream at a time)."""
        need = int(_lib.load().hp_workspace_bytes(int(n_maps), self.K))
        if self._ws is None or self._ws.numel() < need:
            with _lib.on_device(self.device):
                self._ws = torch.zeros(max(need, 1 << 16), dtype=torch.uint8, device=self.device)
            self._plans.clear()          # plans hold the old workspace pointer
        return self._ws

    _PLAN_CACHE = 32

    def _launcher(self, pred, joints, vis, out, finalize, overlap, peer):
        """``(launch, out)`` for one call.  Pre-bound launches are cached ONLY when the caller's tensors are used in
        place (CUDA, right dtype, contiguous: no private copy that could go stale) and the caller supplied ``out``;
        the key is the set of device addresses + the batch size, the cache holds no reference to the inputs (an
        entry whose memory was freed can only be hit again by tensors that own the same addresses) and is a
        32-entry LRU.  Anything else - fp16/bf16 or strided inputs, ``out=None`` - is validated, converted and
        bound per call."""
        cacheable = (out is not None and isinstance(pred, torch.Tensor) and pred.is_cuda and pred.dtype == torch.float32
                     and pred.is_contiguous() and isinstance(joints, torch.Tensor) and joints.is_cuda
                     and joints.dtype == torch.float64 and joints.is_contiguous() and isinstance(vis, torch.Tensor)
                     and vis.is_cuda and vis.dtype == torch.float32 and vis.is_contiguous())
        if cacheable:
            key = (pred.data_ptr(), joints.data_ptr(), vis.data_ptr(), out.result.data_ptr(), out.pred_xy.data_ptr(),
                   out.partial.data_ptr(), pred.shape[0], finalize, int(overlap), peer)
            plans = self._plans
            hit = plans.get(key)
            if hit is not None:
                plans.move_to_end(key)
                return hit, out
        pred, joints, vis = self._check(pred, joints, vis)
        if out is None:
            out = self.alloc_outputs(pred.shape[0], pred.device)
        ws = self._workspace(pred.shape[0] * self.K)
        launch = self._bind(pred, joints, vis, out, ws, finalize, overlap, self._peer_link() if peer else None,
                            not cacheable, bool(peer) and self.defer_exchange and bool(overlap))
        if cacheable:
            self._plans[key] = launch
            while len(self._plans) > self._PLAN_CACHE:
                self._plans.popitem(last=False)
        return launch, out

    def launch_local(self, pred, joints, vis, out=None, overlap=False) -> PipelineResult:
        """This rank's kernel only (finalised locally, no collective)."""
        launch, out = self._launcher(pred, joints, vis, out, True, overlap, False)
        launch()
        return out

    def __call__(self, pred, joints, vis, out=None, overlap=False) -> PipelineResult:
        """pred float32 [B,K,H,W], joints float64 [B,K,2] (image px), vis float32 [B,K,1]|[B,K]: CUDA
        tensors of THIS rank's slice of the batch.  Asynchronous on the current stream.  Single process:
        one kernel.  Sharded (torch.distributed initialised): kernel -> all-reduce of the 4+2K partial
        doubles (the path's only collective) -> finalise kernel; with ``collective="peer"`` the exchange over
        NVLink peer memory happens inside the kernel's last block instead (one kernel per step, no NCCL)."""
        sharded = hpdist.is_distributed(self.group)
        if not sharded or self.collective == "peer":
            launch, out = self._launcher(pred, joints, vis, out, True, overlap, sharded)
            launch()
            return out
        # NCCL collective: kernel on this stream, all-reduce + finalise on a side stream
        launch, out = self._launcher(pred, joints, vis, out, False, overlap, False)
        if out.ready is not None:
            # `out` is being reused: its previous collective (side stream) must have consumed it first
            torch.cuda.current_stream(self.device).wait_event(out.ready)
        launch()
        # NCCL: the collective and the finalise run on a side stream so that they overlap the NEXT step's
        # kernel (steps are independent; `out.ready` / `out.wait()` order consumers after them).
        main = torch.cuda.current_stream(self.device)
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.device)
        comm = self._comm_stream
        comm.wait_stream(main)
        with torch.cuda.stream(comm):
            hpdist.allreduce_partial(out.partial, self.group)
            _lib.call("hp_pipeline_finalize", _lib.ptr(out.partial), self.K, _lib.ptr(out.result),
                      C.c_void_p(comm.cuda_stream))
            if out.ready is None:
                out.ready = torch.cuda.Event()
            out.ready.record(comm)
        return out

    def close(self):
        """Release the peer mailboxes (collective; call on every rank before destroying the process group)."""
        self.join()
        self._plans.clear()
        if self._peer is not None:
            self._peer.close()
            self._peer = None

    def join(self):
        """Order the current stream after every outstanding collective of this pipeline: completes the last step of
        a train of deferred exchanges (one-warp flush kernel) and waits for the NCCL side stream."""
        if self._pending is not None:
            self._pending = None
            with _lib.on_device(self.device):
                self._peer.flush(self._ws)
        if self._comm_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)

    def alloc_outputs(self, B, device=None) -> PipelineResult:
        dev = device or self.device
        K = self.K
        return PipelineResult(
            result=torch.empty((4 + K,), dtype=torch.float64, device=dev),
            partial=torch.empty((hpdist.partial_len(K),), dtype=torch.int64, device=dev),
            pred_xy=torch.empty((B, K, 2), dtype=torch.float32, device=dev),
            maxvals=torch.empty((B, K, 1), dtype=torch.float32, device=dev),
            weight=torch.empty((B, K, 1), dtype=torch.float32, device=dev), K=K)

    # -------------------------------------------------------------------------------- host path
    def run_host(self, pred, joints, vis, slab=64, want_pred_xy=True):
        """End-to-end form for HOST inputs (what a caller holding numpy / CPU tensors uses):
        pinned host buffers -> slabbed H2D copies overlapped with the kernel -> one small D2H.
        Sharded (torch.distributed initialised, ``collective="peer"``): the inputs are THIS rank's slice; the ranks'
        partial vectors are exchanged over the peer mailboxes before the result is copied back, so every rank returns the
        losses / PCK of the whole batch (``pred_xy`` stays the rank's slice).
        ``pred`` float32 [B,K,H,W], ``joints`` float64 [B,K,2], ``vis`` float32 [B,K,1] (numpy arrays or
        CPU tensors; pinned tensors are used in place, anything else is staged through pinned memory).
        Returns ``dict(mse, kl, avg_acc, cnt, acc[K], pred_xy[B,K,2] numpy)`` after synchronising."""
        hp, hj, hv = self._pinned(pred, torch.float32), self._pinned(joints, torch.float64), self._pinned(vis, torch.float32)
        B, K, H, W = hp.shape
        if (K, H, W) != (self.K, self.H, self.W):
            raise ValueError(f"pred is {tuple(hp.shape)}, pipeline was built for K={self.K} H={self.H} W={self.W}")
        slab = max(1, min(int(slab), B))
        st = self._host_buffers(B, slab)
        dev = self.device
        sharded = hpdist.is_distributed(self.group)
        peer = self._peer_link() if (sharded and self.collective == "peer") else None
        if sharded and peer is None:
            raise RuntimeError("run_host on a sharded batch needs collective='peer' (the exchange runs on the step's "
                               "stream between the last slab's kernel and the device->host copy of the result)")
        self.join()
        with _lib.on_device(dev):
            ws = self._workspace(B * K)
            _lib.call("hp_pipeline_fused_host", _lib.ptr(hp), _lib.ptr(hj), _lib.ptr(hv), B, K, H, W,
                      C.c_double(self.stride[0]), C.c_double(self.stride[1]), self.tmp, _lib.ptr(self.tab),
                      C.c_float(self.kl_epsilon), C.c_double(self.thr), self.loss_mask, slab, _lib.ptr(st["d_pred"]),
                      _lib.ptr(st["d_joints"]), _lib.ptr(st["d_vis"]), _lib.ptr(st["d_xy"]), _lib.ptr(st["d_max"]),
                      _lib.ptr(st["d_w"]), _lib.ptr(st["d_partial"]), _lib.ptr(st["d_result"]), _lib.ptr(ws),
                      _lib.ptr(st["h_xy"]) if want_pred_xy else None, _lib.ptr(st["h_result"]),
                      peer._table if peer is not None else None, peer.rank if peer is not None else 0,
                      peer.world if peer is not None else 1,
                      _lib.stream_ptr(dev), C.c_void_p(st["copy_stream"].cuda_stream))
        r = st["h_result"].numpy()
        out = dict(mse=float(r[0]), kl=float(r[1]), avg_acc=float(r[2]) if int(r[3]) else 0, cnt=int(r[3]),
                   acc=r[4:4 + K].copy())
        if want_pred_xy:
            out["pred_xy"] = st["h_xy"][:B].numpy().copy()
        return out

    def host_bytes_per_call(self, B):
        """(h2d, d2h) bytes moved by :meth:`run_host` for a batch of B."""
        K = self.K
        h2d = B * K * self.H * self.W * 4 + B * K * 2 * 8 + B * K * 4
        d2h = (4 + K) * 8 + B * K * 2 * 4
        return h2d, d2h

    @staticmethod
    def _pinned(x, dtype):
        t = torch.as_tensor(x)
        if t.is_cuda:
            raise RuntimeError("run_host takes host arrays; call the pipeline object directly for CUDA tensors")
        t = t.to(dtype).contiguous()
        return t if t.is_pinned() else t.pin_memory()

    def _host_buffers(self, B, slab):
        st = self._host_state
        if st is None or st["B"] < B or st["slab"] != slab:
            dev, K = self.device, self.K
            st = dict(
                B=B, slab=slab,
                d_pred=torch.empty((2 * slab, K, self.H, self.W), dtype=torch.float32, device=dev),
                d_joints=torch.empty((B, K, 2), dtype=torch.float64, device=dev),
                d_vis=torch.empty((B, K), dtype=torch.float32, device=dev),
                d_xy=torch.empty((B, K, 2), dtype=torch.float32, device=dev),
                d_max=torch.empty((B, K), dtype=torch.float32, device=dev),
                d_w=torch.empty((B, K), dtype=torch.float32, device=dev),
                d_partial=torch.empty((hpdist.partial_len(K),), dtype=torch.int64, device=dev),
                d_result=torch.empty((4 + K,), dtype=torch.float64, device=dev),
                h_xy=torch.empty((B, K, 2), dtype=torch.float32).pin_memory(),
                h_result=torch.empty((4 + K,), dtype=torch.float64).pin_memory(),
                copy_stream=torch.cuda.Stream(device=dev),
            )
            self._host_state = st
        return st


class MultiscaleStep:
    """Result of :meth:`MultiscaleEval.step`: ``pred_xy`` at once; ``acc`` (float64 [K+2] = acc[K], avg_acc, cnt) and
    ``counts`` (int32 [2K] hits, valid) over ALL ranks once the step has been collected (next step / ``flush``)."""
    __slots__ = ("K", "pred_xy", "partial", "result", "_acc", "_counts")

    def __init__(self, K, pred_xy, partial, result, acc, counts):
        self.K, self.pred_xy, self.partial, self.result, self._acc, self._counts = K, pred_xy, partial, result, acc, counts

    @property
    def acc(self):
        if self._acc is None:
            K = self.K
            self._acc = torch.cat([self.result[4:4 + K], self.result[2:4]])
        return self._acc

    @property
    def counts(self):
        if self._counts is None:
            self._counts = self.partial[4:4 + 2 * self.K].to(torch.int32)
        return self._counts


class MultiscaleEval:
    """BASELINE.json configs[3]: fuse three resolutions (``0.5*up(lo) + up(mid) + hi``, the rule of
    train1.py:410-424 scaled up), decode the fused map and score PCK against label coordinates - the
    fused map lives in registers only.  Batch-sharded like :class:`HeatmapPipeline`."""

    def __init__(self, num_keypoints=21, thr=0.5, a_lo=0.5, a_mid=1.0, a_hi=1.0, group=None, collective="peer"):
        self.K = int(num_keypoints)
        self.thr = float(thr)
        self.coef = (float(a_lo), float(a_mid), float(a_hi))
        self.group = group
        if collective not in ("nccl", "peer"):
            raise ValueError("collective must be 'nccl' or 'peer'")
        # how the sharded path sums the 2K integer counts: one-warp kernel over the NVLink peer mailboxes (stream-ordered
        # right behind the fuse kernel, no NCCL launch) or torch.distributed.all_reduce + hp_pck_finalize
        self.collective = collective if self.K <= _lib.PEER_MAX_K else "nccl"
        _lib.load()

    def step(self, lo, mid, hi, target_xy):
        """One step of a TRAIN of sharded evaluations (peer collective, 32 / 64 / 128 geometry): the kernel only sends its
        counts; the totals over all ranks land in the returned :class:`MultiscaleStep` when the next ``step`` (same
        stream) or :meth:`flush` collects them - no step waits for a rank that is less than a step late.  Unsharded or
        with ``collective='nccl'`` it is the synchronous ``__call__``."""
        if not (hpdist.is_distributed(self.group) and self.collective == "peer"):
            acc, pred_xy, counts = self(lo, mid, hi, target_xy)
            return MultiscaleStep(self.K, pred_xy, None, None, acc, counts)
        lo = _lib.require_cuda(lo, "MultiscaleEval(lo)")
        mid = _lib.require_cuda(mid, "MultiscaleEval(mid)")
        hi = _lib.require_cuda(hi, "MultiscaleEval(hi)")
        tgt = _lib.require_cuda(target_xy, "MultiscaleEval(target_xy)")
        B, K, H, W = hi.shape
        dev = hi.device
        pred_xy = torch.empty((B, K, 2), dtype=torch.float32, device=dev)
        maxvals = torch.empty((B, K, 1), dtype=torch.float32, device=dev)
        counts = torch.empty((2 * K,), dtype=torch.int32, device=dev)
        acc = torch.empty((K + 2,), dtype=torch.float64, device=dev)
        partial = torch.empty((4 + 2 * K + 6,), dtype=torch.int64, device=dev)
        result = torch.empty((4 + K,), dtype=torch.float64, device=dev)
        with _lib.on_device(dev):
            ws = _lib.workspace(dev, B * K, K)
            self._pending = (dev, ws)
            hpdist.shared_peer_exchange(dev, self.group).fuse_decode_pck(
                lo, self.coef[0], mid, self.coef[1], hi, self.coef[2], tgt, B, K, H, W, self.thr, pred_xy, maxvals,
                counts, acc, ws, partial=partial, result=result)
        return MultiscaleStep(K, pred_xy, partial, result, None, None)

    def flush(self):
        """Complete the outstanding step of a train of :meth:`step` calls (one-warp kernel; a no-op when nothing is pending)."""
        pend = getattr(self, "_pending", None)
        if pend is not None:
            dev, ws = pend
            with _lib.on_device(dev):
                hpdist.shared_peer_exchange(dev, self.group).flush(ws)
            self._pending = None

    def __call__(self, lo, mid, hi, target_xy, local=False):
        """lo/mid/hi float32 [B,K,h,w] CUDA tensors (hi may be None -> output size = 2x mid); target_xy
        float32 [B,K,2].  -> (acc_vec float64 [K+2] = acc[K], avg_acc, cnt ; pred_xy [B,K,2] ; counts int32 [2K]).
        ``local=True``: this rank's batch only, no exchange (the single-GPU reference of a sharded run)."""
        lo = _lib.require_cuda(lo, "MultiscaleEval(lo)")
        mid = _lib.require_cuda(mid, "MultiscaleEval(mid)")
        hi = _lib.require_cuda(hi, "MultiscaleEval(hi)")
        tgt = _lib.require_cuda(target_xy, "MultiscaleEval(target_xy)")
        B, K, H, W = hi.shape
        dev = hi.device
        pred_xy = torch.empty((B, K, 2), dtype=torch.float32, device=dev)
        maxvals = torch.empty((B, K, 1), dtype=torch.float32, device=dev)
        counts = torch.empty((2 * K,), dtype=torch.int32, device=dev)
        acc = torch.empty((K + 2,), dtype=torch.float64, device=dev)
        sharded = not local and hpdist.is_distributed(self.group)
        with _lib.on_device(dev):
            ws = _lib.workspace(dev, B * K, K)
            if sharded and self.collective == "peer":
                # one launch: the kernel's last block sums the 2K integer counts over the NVLink peer mailboxes
                hpdist.shared_peer_exchange(dev, self.group).fuse_decode_pck(
                    lo, self.coef[0], mid, self.coef[1], hi, self.coef[2], tgt, B, K, H, W, self.thr, pred_xy, maxvals,
                    counts, acc, ws)
                return acc, pred_xy, counts
            _lib.call("hp_fuse_decode_pck", _lib.ptr(lo), lo.shape[2], lo.shape[3], C.c_float(self.coef[0]),
                      _lib.ptr(mid), mid.shape[2], mid.shape[3], C.c_float(self.coef[1]), _lib.ptr(hi),
                      C.c_float(self.coef[2]), _lib.ptr(tgt), B, K, H, W, C.c_double(self.thr), _lib.ptr(pred_xy),
                      _lib.ptr(maxvals), _lib.ptr(counts), _lib.ptr(acc), _lib.ptr(ws), _lib.stream_ptr(dev))
            if sharded:
                # the path's one collective, NCCL form: integer sum of the 2K hit / valid counts (exact, order-free)
                torch.distributed.all_reduce(counts, op=torch.distributed.ReduceOp.SUM, group=self.group)
                _lib.call("hp_pck_finalize", _lib.ptr(counts), K, _lib.ptr(acc), _lib.stream_ptr(dev))
        return acc, pred_xy, counts
