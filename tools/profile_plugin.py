"""Launch list of the PLUGIN step (the `e2e` path of bench.py: model.objective(batch) + backward through the drop-in
classes with linear stand-in encoders / decoders), eager, fused tails -- run under
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_plugin.csv python tools/profile_plugin.py
to see how the step splits between the stand-in dense layers (torch / cuBLAS) and the mmvae:: kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import mmvae_b200.synthetic as syn  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2_moe_iwae_cdsprites_l5"
cfg = dict(syn.WORKLOADS[name])
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
fused = any(m["ltype"] == "bce" for m in cfg["mods"])
ps = bench.PluginStep(cfg, cfg["B"], dev, None, 1, 0, fused_tail=fused, graphed=False, fused_enc=True)
for _ in range(3):
    ps.step()
torch.cuda.synchronize()
