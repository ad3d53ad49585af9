#!/usr/bin/env python
"""r2 helper (GPU): relative deviation of every drop-in plugin output / gradient from the frozen reference outputs."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_models_gpu as T  # noqa: E402

golden = torch.load(os.path.join(ROOT, "tests", "golden", "reference_cases.pt"), weights_only=False)
for idx, entry in enumerate(golden["cases"]):
    case, ref = entry["case"], entry["reference"]
    got, out = T.run_dropin(case)
    errs = []
    for k, x in ref.items():
        y = got.get(k)
        if k == "reconstruction_loss" or x is None or y is None or x.shape != y.shape:
            continue
        errs.append((T._rel(y, x), k))
    errs.sort(reverse=True)
    print("%2d %-22s B=%d K=%d D=%d  worst: %s" % (idx, case["name"], case["B"], case["K"], case["D"],
                                                 "  ".join("%s %.1e" % (k, e) for e, k in errs[:4])))
