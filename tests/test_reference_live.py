"""The committed goldens against the UNMODIFIED reference, live: /root/reference in the authoring container, the verbatim
copy under oracle/_ref on the GPU box (oracle/build_ref.py).  Skipped when neither is present."""
import pytest
import torch

from oracle import ref_import
from tests.conftest import load_golden, rel_err
from tests.helpers import case_inputs, cfg_of

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="no reference tree (run oracle/build_ref.py)")


@pytest.mark.parametrize("name", ["nat_b1_64", "nat_b2_64_task2d"])
def test_golden_is_what_the_reference_computes(name, cases):
    meta = cases[name]
    net = ref_import.build_reference(cfg_of(meta["model"]), seed=0)
    x, tid = case_inputs(meta)
    with torch.no_grad():
        y = net(x, tid)
    # bit-identical on the machine that generated the fixture; other CPUs / thread counts re-associate sums (2.4e-6 abs)
    assert rel_err(y, load_golden(name)["out"]) < 2e-6


def test_state_dict_layout_is_the_reference_layout():
    import json
    import os
    from tests.conftest import GOLDEN
    with open(os.path.join(GOLDEN, "state_dict_manifest.json")) as f:
        manifest = json.load(f)["natural"]
    net = ref_import.build_reference(cfg_of("natural"), seed=0)
    sd = net.state_dict()
    assert [e["key"] for e in manifest] == list(sd.keys())
    assert all(list(sd[e["key"]].shape) == e["shape"] for e in manifest)
