"""GPU, world_size 2, NCCL: the batch-sharded CUDA path (LeafStep with a process group) reproduces the single-GPU
full-batch objective: summed loss, all-reduced prior-logit gradient, local gradients of the shard rows -- including the
three forward exchanges (DReG batch sums, MoPoE global batch mean, no exchange for IWAE/ELBO).  Skipped with < 2 GPUs."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard(t, cfg, lo, hi):
    B = cfg["B"]
    out = {"mu": t["mu"][:, lo:hi].contiguous(), "s": t["s"][:, lo:hi].contiguous(), "pz_logits": t["pz_logits"],
           "targets": [x[lo:hi].contiguous() for x in t["targets"]]}
    out["recon"] = [r.view(r.shape[0] // B, B, *r.shape[1:])[:, lo:hi].reshape(-1, *r.shape[1:]).contiguous()
                    for r in t["recon"]]
    out["noise"] = [n[:, lo:hi].contiguous() for n in t["noise"]]
    c = dict(cfg)
    c["B"] = hi - lo
    return c, out


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import mmvae_b200.parallel as par
        import mmvae_b200.workloads as W
        res = {}
        for name, B in (("c2_moe_iwae_cdsprites_l5", 6), ("c4_moe_dreg_mnistsvhn", 8), ("c3_mopoe_elbo_sprites", 6),
                        ("c1_poe_elbo_cdsprites_l1", 6)):
            cfg, t = W.make_leaves(name, B=B, seed=11)
            t["pz_logits"] = torch.randn(1, cfg["D"], generator=torch.Generator().manual_seed(2)) * 0.3
            full = W.LeafStep(cfg, t, device="cuda")
            full_loss = full.run().detach().clone()
            lo, hi = par.shard_range(B, rank, world)
            c2, t2 = _shard(t, cfg, lo, hi)
            step = W.LeafStep(c2, t2, device="cuda", group=dist.group.WORLD, global_batch=B)
            loss = step.run().detach().clone()
            if cfg["obj"] != "dreg":  # the DReG loss is already a global quantity (computed from all-reduced sums)
                dist.all_reduce(loss)
            if step.pz_logits.grad is not None:
                par.GradSync([step.pz_logits], dist.group.WORLD)()
            rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
            e = [rel(loss, full_loss), rel(step.mu.grad, full.mu.grad[:, lo:hi])]
            if full.pz_logits.grad is not None and float(full.pz_logits.grad.abs().max()) > 0:
                e.append(rel(step.pz_logits.grad, full.pz_logits.grad))
                # in-step gradient sync (hook -> side stream), eager and captured in the step's CUDA graph
                s2 = W.LeafStep(c2, t2, device="cuda", group=dist.group.WORLD, global_batch=B, sync_grads=True)
                s2.run()
                torch.cuda.synchronize()
                e.append(rel(s2.pz_logits.grad, full.pz_logits.grad))
                e.append(rel(s2.mu.grad, step.mu.grad))
                if cfg["obj"] != "dreg":  # (a forward collective keeps the DReG step eager)
                    gs = W.GraphedStep(s2)
                    for _ in range(3):
                        gs.run()
                    torch.cuda.synchronize()
                    e.append(rel(s2.pz_logits.grad, full.pz_logits.grad))
                    e.append(rel(s2.mu.grad, step.mu.grad))
                    gs.close()  # a graph that captured the communicator must be destroyed before the process group
                s2.sync.disarm()
            res[name] = max(e)
        if rank == 0:
            q.put(res)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(600)
def test_sharded_cuda_path_matches_single_gpu():
    import torch.multiprocessing as mp
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(500)
        assert p.exitcode == 0
    res = q.get()
    for k, v in res.items():
        assert v < 2e-5, (k, v)
