"""CPU, only where /root/reference exists (this container, not the GPU box): run the real
reference next to the oracle restatement on the same seeded inputs and demand BIT equality for
every output - this is what pins the oracle (the reference has no tests)."""
import numpy as np
import pytest

from oracle import api, cases, ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


@pytest.mark.parametrize("name", list(cases.CASES))
def test_restatement_is_bit_equal_to_reference(name):
    ref = cases.CASES[name](ref_loader.load(), "cpu")
    orc = cases.CASES[name](api.namespace(), "cpu")
    assert sorted(ref) == sorted(orc)
    for key in ref:
        a, b = np.asarray(ref[key]), np.asarray(orc[key])
        assert a.shape == b.shape and a.dtype == b.dtype, key
        same = (a == b) | (np.isnan(a.astype(np.float64)) & np.isnan(b.astype(np.float64)))
        assert same.all(), f"{key}: restatement differs from the reference"


def test_pipeline_composite_matches_reference_calls():
    """oracle.pipeline() == generate_target + JointsMSELoss + JointsKLLoss + accuracy of the reference."""
    import torch
    from oracle import hp_oracle as O
    ns = ref_loader.load()
    d = cases.synth.make_host_batch(907, 3)
    got = O.pipeline(d["pred"], d["joints"], d["vis"], kl_epsilon=1e-7)
    ts, ws = zip(*[ns.generate_target(d["joints"][b], d["vis"][b], (64, 64), 2, (256, 256)) for b in range(3)])
    t, w = np.stack(ts), np.stack(ws)
    tp, tt, tw = torch.from_numpy(d["pred"]), torch.from_numpy(t), torch.from_numpy(w)
    assert float(ns.JointsMSELoss()(tp, tt, tw)) == got["mse"]
    assert float(ns.JointsKLLoss(epsilon=1e-7)(tp, tt, tw)) == got["kl"]
    acc, avg, cnt, pred = ns.accuracy(d["pred"], t)
    assert (acc == got["acc"]).all() and avg == got["avg_acc"] and cnt == got["cnt"]
    assert (pred == got["pred_xy"]).all()
