"""Several GPUs of one node: perturbed nodes sharded over the ranks, every rank's reduce kernel storing its
influence entries into all ranks' matrices through CUDA IPC peer memory (lt_attack_efficient_comm).  Needs ≥ 2 GPUs:
skipped on a single-GPU box, run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs ≥ 2 GPUs")]


def _worker(rank, world, port, workload, n_test, f, ok):
    import torch.distributed as dist
    from linkteller_b200 import attack, ops, workloads
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    group = dist.group.WORLD
    wl = workloads.make(workload, device=dev, n_test=n_test, f=f)
    T = wl.T
    full = ops.attack_efficient(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta)       # all rows, this GPU alone
    comm = attack.make_node_comm(group, T, dev, n=wl.n, hidden=wl.h)
    good = True
    for rep in range(3):                                                                             # epochs reuse the buffers
        with ops.profile() as prof:
            m = attack.influence_matrix(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta, group=group, comm=comm)
            torch.cuda.synchronize()
        good &= bool(torch.equal(m, full))
        # f % 4 == 0 and f < h: the ranks shard the restricted precompute by frontier rows and exchange sign records
        good &= ("comm_push_signs" in prof.report()) == (f is not None)
    with ops.option("shard_rows", 0):                                                                # per-slice precompute: same bits
        m = attack.influence_matrix(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta, group=group, comm=comm)
        torch.cuda.synchronize()
        good &= bool(torch.equal(m, full))
    # the NCCL path (dense all-gather) still agrees
    out = attack.influence_matrix(wl.adj_hat, wl.x, wl.w1, wl.b1, wl.w2, wl.test_nodes, wl.delta, group=group)
    good &= bool(torch.equal(out, full))
    # host-buffer entry: sliced upload + all-gather of the inputs, result on rank 0 only
    hat = wl.adj_hat
    host, h2d, d2h = attack.influence_matrix_host(hat.indptr.cpu(), hat.indices.cpu(), hat.values.cpu(), wl.x.cpu(), wl.w1.cpu(),
                                                  wl.b1.cpu(), wl.w2.cpu(), wl.test_nodes.cpu(), wl.delta, device=dev,
                                                  group=group, comm=comm)
    if rank == 0:
        good &= bool(torch.equal(host, full.cpu()))
    else:
        good &= host is None
    good &= not comm.timed_out()
    ok[rank] = int(good)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


# PPI shape with f = 52 (sharded precompute + sign exchange), PPI as is (f = 50: per-slice precompute), ragged slices
@pytest.mark.parametrize("workload,n_test,f", [("ppi", 500, 52), ("ppi", 500, None), ("small", 7, None)])
def test_sharded_attack_through_peer_memory_equals_one_gpu(workload, n_test, f):
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    ok = mp.get_context("spawn").Array("i", [0] * world)
    port = 29600 + (os.getpid() + n_test) % 300
    mp.spawn(_worker, args=(world, port, workload, n_test, f, ok), nprocs=world, join=True)
    assert list(ok) == [1] * world
