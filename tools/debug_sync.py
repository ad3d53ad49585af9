"""Debug harness (torchrun, 2+ ranks): in-step gradient all-reduce, eager and captured in the step's CUDA graph.
Every stage prints when it finishes; a watchdog dumps all Python stacks and exits if a stage hangs."""
import faulthandler
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(int(os.environ.get("DEBUG_SYNC_TIMEOUT", "60")), exit=True)
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)


def say(*a):
    print("[rank %d %.2fs]" % (rank, time.perf_counter() - T0), *a, flush=True)


T0 = time.perf_counter()
try:
    import mmvae_b200.workloads as W
    torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
    name = sys.argv[1] if len(sys.argv) > 1 else "c2_moe_iwae_cdsprites_l5"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    cfg, t = W.make_leaves(name, B=B, seed=11 + rank)
    t["pz_logits"] = torch.randn(1, cfg["D"], generator=torch.Generator().manual_seed(2)) * 0.3
    base = W.LeafStep(cfg, t, device=dev, group=dist.group.WORLD, global_batch=B * world)
    base.run()
    ref = base.pz_logits.grad.clone()
    dist.all_reduce(ref)
    torch.cuda.synchronize()
    say("reference step + eager all-reduce ok", ref.flatten()[:3].tolist())
    s2 = W.LeafStep(cfg, t, device=dev, group=dist.group.WORLD, global_batch=B * world, sync_grads=True)
    say("armed:", s2.sync is not None and bool(s2.sync._handles))
    for i in range(3):
        s2.run()
        torch.cuda.synchronize()
        say("eager in-step sync run", i, "max diff", float((s2.pz_logits.grad - ref).abs().max()))
    gs = W.GraphedStep(s2)
    torch.cuda.synchronize()
    say("captured")
    for i in range(3):
        gs.run()
        torch.cuda.synchronize()
        say("replay", i, "max diff", float((s2.pz_logits.grad - ref).abs().max()))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    a.record()
    for _ in range(50):
        gs.run()
    b.record()
    torch.cuda.synchronize()
    say("50 replays: %.1f us/step" % (a.elapsed_time(b) * 1e3 / 50))
    gs.close()
    torch.cuda.synchronize()
    say("graph destroyed")
except Exception:
    say("EXCEPTION\n" + traceback.format_exc())
finally:
    try:
        dist.destroy_process_group()
        say("process group destroyed")
    except Exception:
        pass
    faulthandler.cancel_dump_traceback_later()
