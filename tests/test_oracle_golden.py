"""CPU: the oracle restatement (oracle/refmath.py) reproduces the frozen outputs of the UNMODIFIED reference
(tests/golden/reference_cases.pt, written by oracle/gen_golden.py from /root/reference)."""
import pytest
import torch

from oracle import cases, refmath

RTOL = 2e-6


def _cmp(ref, got, name):
    for k, x in ref.items():
        y = got.get(k)
        if x is None or y is None:
            assert (x is None or float(x.abs().max()) == 0) and (y is None or float(y.abs().max()) == 0), (name, k)
            continue
        assert x.shape == y.shape, (name, k)
        err = float((x - y).abs().max()) / max(float(x.abs().max()), 1e-12)
        assert err <= RTOL, "%s %s rel err %.3e" % (name, k, err)


def test_all_cases_present(golden):
    assert [c["case"]["name"] for c in golden["cases"]] == [c["name"] for c in cases.case_list()]


@pytest.mark.parametrize("idx", range(len(cases.case_list())))
def test_restatement_matches_reference(golden, idx):
    entry = golden["cases"][idx]
    _cmp(entry["reference"], cases.run_oracle(entry["case"]), entry["case"]["name"])


def test_case_inputs_regenerate_from_seed(golden):
    """The committed script + seeds regenerate exactly the frozen inputs."""
    for entry, fresh in zip(golden["cases"], cases.case_list()):
        for a, b in zip(entry["case"]["mods"], fresh["mods"]):
            for k in ("mu", "s", "W", "b", "target"):
                assert torch.equal(a[k], b[k]), (fresh["name"], k)
        for a, b in zip(entry["case"]["noise"], fresh["noise"]):
            assert torch.equal(a, b)


def test_chunk_maps_bit_exact(golden):
    """mixture_component_selection chunk bounds (mmvae_models.py:396-410): index work, bit exact."""
    for (S, B), ends in golden["chunk_ends"].items():
        st, en = refmath.mopoe_chunk_bounds(S, B)
        assert en == ends, (S, B)
        rm = refmath.mopoe_row_to_subset(S, B)
        assert [int((rm <= k).sum()) for k in range(S)] == ends


def test_unimodal_restatement_matches_reference(golden):
    from oracle.validate_against_reference import run_oracle_unimodal
    assert len(golden["unimodal"]) == 3
    for entry in golden["unimodal"]:
        _cmp(entry["reference"], run_oracle_unimodal(entry["case"]), entry["case"]["name"])


def test_kl_table_restatement_matches_reference_make_kl_df(golden):
    """refmath.kl_table against the frozen DataFrame values of the reference's utils.make_kl_df (utils.py:130-162)."""
    assert len(golden["kl_df"]) == 3
    for entry in golden["kl_df"]:
        c, ref = entry["case"], entry["reference"]
        t = refmath.kl_table(c["family"], c["locs"], c["scales"], c["loc0"], c["scale0"])
        mine = t.permute(0, 2, 1).reshape(-1).double()
        assert mine.shape == ref["values"].shape
        assert float((mine - ref["values"]).abs().max() / ref["values"].abs().max()) <= RTOL, c["name"]
