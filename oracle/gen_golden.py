"""TEST INFRASTRUCTURE (build container only): freeze reference outputs as golden fixtures.

    PYTHONHASHSEED=0 python -m oracle.gen_golden

Runs every case of oracle/cases.py through the UNMODIFIED reference classes (imported in place from
/root/reference, see oracle/ref_inplace.py) and stores inputs + reference outputs (loss, kld, reconstruction
terms, every gradient, the PoE subset order the reference used, MoPoE chunk maps) in tests/golden/.
The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)

from oracle import cases, ref_inplace, refmath  # noqa: E402
from oracle.validate_against_reference import (kl_df_cases, run_reference, run_reference_kl_df,  # noqa: E402
                                               run_reference_unimodal, unimodal_cases)

OUT = os.path.join(_ROOT, "tests", "golden")


def chunk_maps():
    """Row->component maps produced by the reference's MoPOE.moe_fusion on a (S,B,D) stack (function-level
    contract of mixture_component_selection, mmvae_models.py:396-410)."""
    models, _, _ = ref_inplace.load()
    mop = models.mopoe.__new__(models.mopoe)
    out = {}
    for S in (1, 2, 3, 7, 15, 31):
        for B in (1, 2, 3, 5, 6, 7, 8, 16, 24, 31, 32, 33, 64, 100, 255, 256, 1000, 4096, 65536):
            mus = torch.arange(B, dtype=torch.float32).reshape(1, B, 1).repeat(S, 1, 1) + \
                1e6 * torch.arange(S, dtype=torch.float32).reshape(S, 1, 1)
            sel, _ = mop.moe_fusion(mus, mus.clone(), (1 / float(S)) * torch.ones(S))
            ref_map = (sel.reshape(-1) // 1e6).to(torch.int32)
            # store run-length encoded: chunk end offsets
            ends = [int((ref_map <= k).sum()) for k in range(S)]
            out[(S, B)] = ends
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    blob = {"cases": [], "torch": torch.__version__, "hashseed": os.environ.get("PYTHONHASHSEED")}
    for case in cases.case_list():
        ref, subsets = run_reference(case)
        if subsets is not None:
            case["poe_subsets"] = subsets
        blob["cases"].append({"case": case, "reference": ref})
        print("froze", case["name"], float(ref["loss"]))
    blob["unimodal"] = [{"case": c, "reference": run_reference_unimodal(c)} for c in unimodal_cases()]
    blob["kl_df"] = [{"case": c, "reference": run_reference_kl_df(c)} for c in kl_df_cases()]
    blob["chunk_ends"] = chunk_maps()
    path = os.path.join(OUT, "reference_cases.pt")
    torch.save(blob, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
