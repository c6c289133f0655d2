"""Run under torchrun on >= 2 GPUs: the batch-sharded pipeline (kernel -> NCCL all-reduce of the int64
partial vector -> finalise) must equal the single-GPU pipeline on the concatenated batch BIT FOR BIT, and
match the CPU oracle.  Launched by tests/test_gpu_parity.py::test_sharded_pipeline_two_gpus."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hp = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    B = 37
    d = hp.synth.make_host_batch(9001, B)
    lo, hi = hp.dist.shard_bounds(B, rank, world)
    pipe = hp.HeatmapPipeline(kl_epsilon=1e-7, device=dev)
    assert hp.dist.is_distributed()
    outs = []
    for rep in range(3):    # several steps in flight: exercises the side-stream overlap
        outs.append(pipe(torch.from_numpy(d["pred"][lo:hi]).to(dev), torch.from_numpy(d["joints"][lo:hi]).to(dev),
                         torch.from_numpy(d["vis"][lo:hi]).to(dev)))
    got = [o.host() for o in outs]
    part = outs[-1].wait().partial.cpu().numpy()
    # the same steps with the NVLink peer-memory exchange instead of NCCL (one kernel per step): 40 steps in flight,
    # separate outputs, overlapped launches - must be bit-identical on every rank
    peer = hp.HeatmapPipeline(kl_epsilon=1e-7, device=dev, collective="peer")
    x = [torch.from_numpy(d[k][lo:hi]).to(dev) for k in ("pred", "joints", "vis")]
    pouts = [peer.alloc_outputs(hi - lo, dev) for _ in range(8)]
    for overlap in (False, True, 2):
        for o in pouts:
            o.partial.zero_(); o.result.zero_()
        for rep in range(40):
            peer(x[0], x[1], x[2], out=pouts[rep % 8], overlap=overlap)
        peer.join()
        torch.cuda.synchronize()
        for o in pouts:
            assert np.array_equal(o.partial.cpu().numpy(), part), (overlap, o.partial.cpu().numpy(), part)
            assert torch.equal(o.result, outs[-1].result), (overlap, o.result, outs[-1].result)
    # deferred exchange made visible: right after a train WITHOUT join() the last step's totals are still outstanding;
    # reading them through PipelineResult.host() must complete the exchange first
    for rep in range(5):
        last = peer(x[0], x[1], x[2], out=pouts[rep % 8], overlap=True)
    assert peer._pending is last
    h_last = last.host()
    assert peer._pending is None and h_last["mse"] == got[-1]["mse"] and h_last["kl"] == got[-1]["kl"] and h_last["cnt"] == got[-1]["cnt"]
    # synchronous steps (no overlap) right after deferred ones complete what is outstanding themselves
    for o in pouts:
        o.partial.zero_(); o.result.zero_()
    peer(x[0], x[1], x[2], out=pouts[0], overlap=True)
    peer(x[0], x[1], x[2], out=pouts[1], overlap=False)
    assert peer._pending is None
    torch.cuda.synchronize()
    for o in pouts[:2]:
        assert np.array_equal(o.partial.cpu().numpy(), part) and torch.equal(o.result, outs[-1].result)
    # the host-buffer entry point, sharded: every rank gets the totals of the whole batch
    hr = peer.run_host(d["pred"][lo:hi], d["joints"][lo:hi], d["vis"][lo:hi], slab=8)
    assert hr["mse"] == got[-1]["mse"] and hr["kl"] == got[-1]["kl"] and hr["cnt"] == got[-1]["cnt"], (hr, got[-1])
    assert np.array_equal(hr["acc"], got[-1]["acc"]) and np.array_equal(hr["pred_xy"], pouts[0].pred_xy.cpu().numpy())
    # NCCL path with reused outputs and overlapped launches
    nouts = [pipe.alloc_outputs(hi - lo, dev) for _ in range(3)]
    for rep in range(30):
        pipe(x[0], x[1], x[2], out=nouts[rep % 3], overlap=True)
    pipe.join()
    torch.cuda.synchronize()
    for o in nouts:
        assert np.array_equal(o.partial.cpu().numpy(), part), ("nccl", o.partial.cpu().numpy(), part)
    # configs[3]: sharded fuse + decode + PCK; the all-reduced counts and accuracies must equal the unsharded ones
    Bm = 10
    dm = hp.synth.make_host_batch(9100, Bm, 21, 128, 128, image_size=512)
    mid_h, lo_h = hp.synth.make_lowres_heads(9101, dm["pred"], (64, 32))
    tgt_h = np.random.RandomState(9102).randint(0, 128, size=(Bm, 21, 2)).astype(np.float32)
    lo_m, hi_m = hp.dist.shard_bounds(Bm, rank, world)
    tm = lambda a, sl: torch.from_numpy(a[sl]).to(dev)
    sl = slice(lo_m, hi_m)
    acc_s, _, counts_s = hp.MultiscaleEval(21)(tm(lo_h, sl), tm(mid_h, sl), tm(dm["pred"], sl), tm(tgt_h, sl))   # peer exchange
    acc_n, _, counts_n = hp.MultiscaleEval(21, collective="nccl")(tm(lo_h, sl), tm(mid_h, sl), tm(dm["pred"], sl), tm(tgt_h, sl))
    assert torch.equal(acc_s, acc_n) and torch.equal(counts_s, counts_n), "peer and NCCL sums of the PCK counts differ"
    # a TRAIN of deferred steps (each only sends; the next step / flush collects): every step's totals == the synchronous ones
    ev_d = hp.MultiscaleEval(21)
    steps_d = [ev_d.step(tm(lo_h, sl), tm(mid_h, sl), tm(dm["pred"], sl), tm(tgt_h, sl)) for _ in range(4)]
    ev_d.flush()
    torch.cuda.synchronize()
    for st in steps_d:
        assert torch.equal(st.acc, acc_s) and torch.equal(st.counts, counts_s), "deferred fuse step differs from the synchronous one"
    acc_s, counts_s = acc_s.cpu().numpy(), counts_s.cpu().numpy()
    # a geometry the staged kernel (exchange folded into its last block) does not take - 16 / 32 / 64: the exchange is the
    # one-warp kernel behind the fuse kernel, inside the same C call
    d6 = hp.synth.make_host_batch(9110, Bm, 21, 64, 64)
    mid6, lo6 = hp.synth.make_lowres_heads(9111, d6["pred"], (32, 16))
    tgt6 = np.random.RandomState(9112).randint(0, 64, size=(Bm, 21, 2)).astype(np.float32)
    acc6_s, _, counts6_s = hp.MultiscaleEval(21)(tm(lo6, sl), tm(mid6, sl), tm(d6["pred"], sl), tm(tgt6, sl))
    acc6_n, _, counts6_n = hp.MultiscaleEval(21, collective="nccl")(tm(lo6, sl), tm(mid6, sl), tm(d6["pred"], sl), tm(tgt6, sl))
    assert torch.equal(acc6_s, acc6_n) and torch.equal(counts6_s, counts6_n), "peer and NCCL sums differ (16/32/64)"
    acc6_s, counts6_s = acc6_s.cpu().numpy(), counts6_s.cpu().numpy()
    peer.close()                    # the shared mailboxes: collective, before the process group goes away
    # single-GPU reference on the whole batch (group of one rank)
    solo_group = dist.new_group(ranks=[rank]) if False else None
    dist.barrier()
    dist.destroy_process_group()
    solo = hp.HeatmapPipeline(kl_epsilon=1e-7, device=dev)
    ref = solo(torch.from_numpy(d["pred"]).to(dev), torch.from_numpy(d["joints"]).to(dev),
               torch.from_numpy(d["vis"]).to(dev))
    want = ref.host()
    wpart = ref.partial.cpu().numpy()
    assert np.array_equal(part, wpart), (part, wpart)                      # every entry is an exact integer
    for g in got:
        assert g["mse"] == want["mse"] and g["kl"] == want["kl"], (g, want)
        assert np.array_equal(g["acc"], want["acc"]) and g["cnt"] == want["cnt"] and g["avg_acc"] == want["avg_acc"]
    full = slice(0, Bm)
    acc_f, _, counts_f = hp.MultiscaleEval(21)(tm(lo_h, full), tm(mid_h, full), tm(dm["pred"], full), tm(tgt_h, full))
    assert np.array_equal(counts_s, counts_f.cpu().numpy()) and np.array_equal(acc_s, acc_f.cpu().numpy()), "sharded C4 differs"
    acc6_f, _, counts6_f = hp.MultiscaleEval(21)(tm(lo6, full), tm(mid6, full), tm(d6["pred"], full), tm(tgt6, full))
    assert np.array_equal(counts6_s, counts6_f.cpu().numpy()) and np.array_equal(acc6_s, acc6_f.cpu().numpy()), "sharded 16/32/64 differs"
    if rank == 0:
        from oracle import hp_oracle as O
        o = O.pipeline(d["pred"], d["joints"], d["vis"], kl_epsilon=1e-7)
        assert np.array_equal(want["acc"], o["acc"]) and want["cnt"] == o["cnt"]
        np.testing.assert_allclose(want["mse"], o["mse"], rtol=1e-5)
        np.testing.assert_allclose(want["kl"], o["kl"], rtol=1e-5)
    print(f"rank {rank}: sharded (NCCL and peer-memory exchange) == single-GPU bit for bit")


if __name__ == "__main__":
    main()
