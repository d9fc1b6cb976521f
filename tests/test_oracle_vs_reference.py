"""CPU, build container only: fixture-drift guard.  When /root/reference is present the golden generators are
re-run on the VERBATIM reference (under oracle/ref_standins.py) and their output must be bit-identical to the
committed tests/golden/*.npz; the generators themselves assert oracle-vs-reference agreement while they run.
On the GPU box (no /root/reference) these tests skip; the committed fixtures are what travels."""
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "gotennet")), reason="needs /root/reference")

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


@pytest.fixture(scope="module")
def ref():
    from oracle.ref_standins import import_reference
    n = torch.get_num_threads()
    torch.set_num_threads(4)  # the thread count the committed fixtures were generated with
    yield import_reference()
    torch.set_num_threads(n)


def _same(out, path):
    gold = np.load(path)
    assert sorted(gold.files) == sorted(out), (sorted(set(gold.files) ^ set(out)))
    for k in gold.files:
        assert np.array_equal(np.asarray(out[k]), gold[k]), k


@pytest.mark.parametrize("name", ["cfg1", "yaml_l2", "l3_trunc", "eu_mlp", "eu_linw_ln_ev", "eu_linwa_postln_gated"])
def test_representation_fixture_regenerates_bit_identically(ref, name, golden_dir):
    import make_golden
    from oracle.golden_cases import CASES, NORM_CASES
    out = make_golden.run_case(ref, name, {**CASES, **NORM_CASES}[name], save=False)
    _same(out, os.path.join(golden_dir, name + ".npz"))


def test_head_fixture_regenerates_bit_identically(ref, golden_dir):
    import make_golden_head
    from oracle.golden_cases import HEAD_CASES
    out = make_golden_head.run_case(ref, "head_forces_l2", HEAD_CASES["head_forces_l2"], save=False)
    _same(out, os.path.join(golden_dir, "head_forces_l2.npz"))


def test_head2_fixture_regenerates_bit_identically(ref, golden_dir):
    import make_golden_heads2
    from oracle.golden_cases import HEAD2_CASES
    for name in ("dipole_l2", "ese_l2"):
        out = make_golden_heads2.run_case(ref, name, HEAD2_CASES[name], save=False)
        _same(out, os.path.join(golden_dir, name + ".npz"))
