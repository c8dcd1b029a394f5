"""FlatBucketAllReduce (the gradient exchange of BASELINE config #4) on a world-size-2 gloo group (CPU): bucketed, averaged
gradients equal the mean of the per-rank gradients; the NCCL / bf16 twin of this path runs on the GPU box."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, dtype_name, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from upsnet_b200.training import FlatBucketAllReduce
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.ReLU(), torch.nn.Linear(64, 11), torch.nn.Linear(11, 3))
        model[2].weight.requires_grad_(False)                       # frozen parameters are skipped (backbone_freeze_at)
        x = torch.randn(5, 37, generator=torch.Generator().manual_seed(100 + rank))
        model(x).pow(2).sum().backward()
        local = [p.grad.clone() for p in model.parameters() if p.requires_grad]
        ar = FlatBucketAllReduce(model.parameters(), bucket_bytes=128, reduce_dtype=getattr(torch, dtype_name))
        assert len(ar.buckets) >= 2
        ar()
        got = [p.grad.clone() for p in model.parameters() if p.requires_grad]
        # reference: explicit per-tensor fp32 all-reduce
        for g in local:
            dist.all_reduce(g)
            g /= world
        tol = 1e-6 if dtype_name == "float32" else 2e-2
        ok = all((a - b).abs().max().item() <= tol * max(1.0, b.abs().max().item()) for a, b in zip(got, local))
        out.put((rank, ok))
    finally:
        dist.destroy_process_group()


def _run(dtype_name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dtype_name, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok in res), res


def test_flat_bucket_allreduce_fp32_gloo():
    _run("float32")


def test_flat_bucket_allreduce_bf16_gloo():
    _run("bfloat16")
