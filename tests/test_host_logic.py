"""CPU: host logic of the package and the C-ABI surface (no compute calls - there is no GPU here)."""
import ctypes
import importlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import hp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hp = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")
L = importlib.import_module("domain-adaptative-hand-pose-estimation_b200._lib")


@pytest.fixture(scope="module")
def built_lib():
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(L.LIB_PATH)


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in include/hp_b200.h but not exported"
    assert sorted(L.PROTOTYPES) == names, "ctypes prototypes out of sync with the header"


def test_prototype_arity_matches_header():
    text = open(os.path.join(ROOT, "include", "hp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, (_, args) in L.PROTOTYPES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", text, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), f"{name}: header has {n} parameters, ctypes binding {len(args)}"


def test_library_metadata_calls(built_lib):
    handle = L.load()
    assert handle.hp_version() == 100
    assert handle.hp_workspace_bytes(5376, 21) >= 1024
    assert isinstance(handle.hp_last_error(), bytes)


def test_product_path_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hp.get_max_preds(np.zeros((1, 2, 4, 4), np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hp.JointsMSELoss()(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hp.generate_target(np.zeros((21, 2)), np.ones((21, 1)), (64, 64), 2, (256, 256))
    with pytest.raises(RuntimeError):
        hp.HeatmapPipeline()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "domain-adaptative-hand-pose-estimation_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
            assert "/root/reference" not in src, fn


def test_gaussian_table_equals_reference_patch():
    for sigma, tmp in ((2, 6), (2, 4), (2, 3), (1, 3), (3, 9)):
        tab = L.gaussian_table_host(sigma, tmp)
        g = O.gaussian_patch(sigma, tmp)
        c = tmp
        for dy in range(-tmp, tmp + 1):
            for dx in range(-tmp, tmp + 1):
                assert tab[dx * dx + dy * dy] == g[dy + c, dx + c]
    assert L.gaussian_table_host(2, 6)[0] == 1.0
    with pytest.raises(NotImplementedError):
        L.integer_tmp(1.5 * 3)


def test_shard_bounds_partition():
    D = hp.dist
    for total in (0, 1, 7, 256, 2048, 8191):
        for world in (1, 2, 3, 8):
            spans = [D.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _oracle_partial(d, eps):
    """The int64 partial vector a rank would produce, built from the oracle's per-map losses."""
    import torch
    r = O.pipeline(d["pred"], d["joints"], d["vis"], kl_epsilon=eps)
    n = d["pred"].shape[0] * d["pred"].shape[1]
    hw = d["pred"].shape[2] * d["pred"].shape[3]
    tp, tt, tw = torch.from_numpy(d["pred"]), torch.from_numpy(r["target"]), torch.from_numpy(r["weight"])
    mse_map = O.joints_mse_loss(tp, tt, tw, reduction="none").numpy()
    kl_map = (torch.nn.functional.kl_div(torch.log_softmax(tp.reshape(tp.shape[0], tp.shape[1], -1), -1),
                                         (lambda q: q / q.sum(-1, keepdim=True))(tt.reshape(tp.shape[0], tp.shape[1], -1) + eps),
                                         reduction="none").sum(-1) * tw.view(tp.shape[0], tp.shape[1])).numpy()
    mfx, mcls = hp.dist.loss_to_fx(mse_map)
    kfx, kcls = hp.dist.loss_to_fx(kl_map)
    part = np.concatenate([[mfx, kfx, n, n * hw], r["hits"], r["valid"], mcls, kcls]).astype(np.int64)
    return part, r


def test_finalize_partial_host_matches_oracle():
    d = hp.synth.make_host_batch(31, 6)
    part, r = _oracle_partial(d, 1e-7)
    f = hp.dist.finalize_partial_host(part, 21)
    assert np.array_equal(f["acc"], r["acc"]) and f["cnt"] == r["cnt"] and f["avg_acc"] == r["avg_acc"]
    np.testing.assert_allclose(f["mse"], r["mse"], rtol=1e-6)
    np.testing.assert_allclose(f["kl"], r["kl"], rtol=1e-6)
    assert len(part) == hp.dist.partial_len(21)


def test_fixed_point_partials_are_exactly_associative():
    """Any split of the per-map losses sums to the same int64 - the property that makes the result
    independent of block schedule, slab split and GPU sharding."""
    rs = np.random.RandomState(3)
    v = rs.lognormal(0, 2, size=1000) * rs.choice([-1, 1], size=1000)
    whole, _ = hp.dist.loss_to_fx(v)
    for parts in (2, 3, 8):
        assert sum(hp.dist.loss_to_fx(c)[0] for c in np.array_split(v, parts)) == whole
    fx, cls = hp.dist.loss_to_fx([1.0, float("nan"), float("inf"), -3e7])
    assert fx == 1 << 40 and cls == [1, 1, 1]
    assert np.isnan(hp.dist._loss_from_fx(fx, *cls, 4))


def test_synth_is_deterministic_and_has_edge_cases():
    a = hp.synth.make_host_batch(5, 64)
    b = hp.synth.make_host_batch(5, 64)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    flat = a["pred"].reshape(64 * 21, -1)
    assert (flat.max(axis=1) <= 0).any(), "no all-nonpositive map"
    mx = flat.max(axis=1, keepdims=True)
    assert ((flat == mx).sum(axis=1) > 1).any(), "no duplicated maximum"


_WORKER = r'''
import os, sys, importlib, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from oracle import hp_oracle as O
hp = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
B = 7
d = hp.synth.make_host_batch(77, B)
lo, hi = hp.dist.shard_bounds(B, rank, world)
sl = {k: v[lo:hi] for k, v in d.items()}
sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from test_host_logic import _oracle_partial
part_np, r = _oracle_partial(sl, 1e-7)
part = torch.from_numpy(part_np)
assert hp.dist.is_distributed()
hp.dist.allreduce_partial(part)                      # the path's single collective
f = hp.dist.finalize_partial_host(part.numpy(), 21)
full = O.pipeline(d["pred"], d["joints"], d["vis"], kl_epsilon=1e-7)
assert np.array_equal(f["acc"], full["acc"]) and f["cnt"] == full["cnt"] and f["avg_acc"] == full["avg_acc"]
assert np.array_equal(f["hits"], full["hits"]) and np.array_equal(f["valid"], full["valid"])
np.testing.assert_allclose(f["mse"], full["mse"], rtol=1e-6)
np.testing.assert_allclose(f["kl"], full["kl"], rtol=1e-6)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_gloo_sharding_reproduces_full_batch(tmp_path):
    """N>1 host logic on CPU: shard the batch over 2 gloo ranks, all-reduce the partial vector,
    finalise -> identical PCK (bit-exact) and losses as the unsharded batch."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29617",
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_peer_exchange_rejects_more_joints_than_the_mailboxes_carry(built_lib):
    """K <= HP_MAX_K = 64 is fine on one GPU, but the peer mailboxes carry 2*(4+2K+6) <= 128
    tagged entries per source -> K <= 27.  Every entry point that binds a PeerLink with world > 1 must refuse
    K = 28..64 with a clean error instead of overrunning a (remote) mailbox.  Argument validation only: the fake
    pointers are never dereferenced and nothing is launched."""
    lib = L.load()
    C = ctypes
    fake = 0x10000                                   # non-null, 16-byte aligned
    boxes = (C.c_void_p * 2)(fake, fake)
    plan = C.c_void_p()

    def create(K, world):
        return lib.hp_pipeline_plan_create(fake, fake, fake, 4, K, 64, 64, C.c_double(4.0), C.c_double(4.0), 6, fake,
                                           C.c_float(1e-7), C.c_double(0.5), 3, fake, fake, fake, fake, 0, fake, fake,
                                           boxes, 0, world, C.c_uint(0), C.byref(plan))
    assert create(32, 2) != 0 and b"peer mailbox" in lib.hp_last_error()
    assert create(28, 2) != 0
    assert create(27, 2) == 0 and plan.value
    lib.hp_pipeline_plan_destroy(plan)
    assert create(64, 1) == 0 and plan.value          # unsharded: up to HP_MAX_K
    lib.hp_pipeline_plan_destroy(plan)
    rc = lib.hp_pipeline_fused_peer(fake, fake, fake, 4, 32, 64, 64, C.c_double(4.0), C.c_double(4.0), 6, fake,
                                    C.c_float(1e-7), C.c_double(0.5), 3, fake, fake, fake, fake, fake, fake, boxes, 0, 2,
                                    C.c_int64(0), C.c_uint(0), None)
    assert rc != 0 and b"peer mailbox" in lib.hp_last_error()
    assert L.PEER_MAX_K == 27


def test_generate_target_refuses_to_run_inside_a_loader_worker(monkeypatch):
    """the per-sample CUDA generate_target must not be reached from a forked DataLoader worker."""
    import torch.utils.data
    T = importlib.import_module("domain-adaptative-hand-pose-estimation_b200.target")
    monkeypatch.setattr(torch.utils.data, "get_worker_info", lambda: object())
    with pytest.raises(RuntimeError, match="DataLoader worker"):
        T.generate_target(np.zeros((21, 2)), np.ones((21, 1), np.float32), (64, 64), 2, (256, 256))
