"""CPU: the oracle restatement (oracle/hp_oracle.py) against the golden vectors that
oracle/gen_golden.py froze from the REAL reference.  Same host => the bar is the one in the key
prefix (x: bit-exact, c: rtol 1e-5)."""
import pytest

from oracle import api, cases
from oracle.gen_golden import digest


@pytest.mark.parametrize("name", list(cases.CASES))
def test_oracle_matches_reference_goldens(name, golden):
    got = digest(cases.CASES[name](api.namespace(), "cpu"))
    cases.compare(got, golden(name))


def test_manifest_lists_every_case():
    import json, os
    root = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(root, "golden", "MANIFEST.json")) as f:
        man = json.load(f)
    assert sorted(man["files"]) == sorted(f"{n}.npz" for n in cases.CASES)
    assert "real reference" in man["source"]
