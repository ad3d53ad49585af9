"""GPU, world_size 2, NCCL: the batch-sharded CUDA path (LeafStep with a process group, captured in a CUDA graph with its
collectives) reproduces the single-GPU full-batch objective: summed loss, all-reduced prior-logit gradient, local
gradients of the shard rows -- including the forward exchanges (DReG batch sums, MoPoE global batch mean, optimal_sigma
sum + element count with UNEVEN shards; no exchange for IWAE/ELBO).  Skipped with < 2 GPUs; `bench.py --gpus N` runs the
same check (parallel.sharded_parity) on every multi-GPU bench and reports it as `parity_n`."""
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


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import mmvae_b200.parallel as par
        res = {}
        peer = par.PeerGroup(dist.group.WORLD, dev)  # collectives fused into our kernels over NVLink peer memory
        for mode, coll in (("nccl", dist.group.WORLD), ("peer", peer)):
            for name, B in (("c2_moe_iwae_cdsprites_l5", 6), ("c4_moe_dreg_mnistsvhn", 9), ("c3_mopoe_elbo_sprites", 6),
                            ("c1_poe_elbo_cdsprites_l1", 7), ("c3_mopoe_elbo_vilanro", 7), ("c4_moe_dreg_latent_only", 11)):
                err = torch.tensor([par.sharded_parity(name, B, coll, dev)], device=dev)
                dist.all_reduce(err, op=dist.ReduceOp.MAX)
                res[mode + ":" + name] = float(err)
        assert not peer.error(), "a peer-memory wait timed out"
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
        # same kernels on a different split of the batch: only the summation order of the batch reductions differs
        # (DReG: its softmax over K amplifies that, see tests/test_workloads_gpu.py)
        assert v < (1e-4 if "dreg" in k else 2e-5), (k, v)
