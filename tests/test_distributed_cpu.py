"""CPU, gloo, world size 2: the N>1 plumbing of the harness.  The attention path shards on batch with no
data-path collective; the only collective is DDP's gradient all-reduce.  The B200 module
has no CPU path, so the oracle attention module stands in for it inside the same MsViT harness."""
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARCH = "l1,h2,d16,n1,s1,g1,p4,f4,a0_l2,h2,d32,n1,s1,g1,p2,f4_l3,h2,d32,n1,s0,g1,p2,f7"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build():
    from oracle.vil_oracle import OracleLong2DSCSelfAttention
    from vision_longformer_b200 import MsViT
    torch.manual_seed(0)
    return MsViT(arch=ARCH, img_size=32, num_classes=7, sharew=True, norm_embed=True, drop_path_rate=0.0,
                 attn_cls=OracleLong2DSCSelfAttention).double()


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    net = _build()
    ddp = torch.nn.parallel.DistributedDataParallel(net)
    gen = torch.Generator().manual_seed(42)
    x = torch.randn(4, 3, 32, 32, generator=gen, dtype=torch.float64)
    y = torch.randint(0, 7, (4,), generator=gen)
    shard = slice(rank * 2, rank * 2 + 2)           # batch sharding: each rank sees its own images only
    loss = torch.nn.functional.cross_entropy(ddp(x[shard]), y[shard])
    loss.backward()
    grads = {n: p.grad.clone() for n, p in net.named_parameters()}
    # max-over-ranks reduction used by bench.py for its timing
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    if rank == 0:
        torch.save(grads, os.path.join(out_dir, "grads.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_world2_matches_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(tmp_path, "grads.pt"))
    net = _build()
    gen = torch.Generator().manual_seed(42)
    x = torch.randn(4, 3, 32, 32, generator=gen, dtype=torch.float64)
    y = torch.randint(0, 7, (4,), generator=gen)
    torch.nn.functional.cross_entropy(net(x), y).backward()       # mean over the global batch == mean of shard means
    for n, p in net.named_parameters():
        assert torch.allclose(got[n], p.grad, rtol=1e-9, atol=1e-12), n


def test_reference_arm_only_runs_on_rank0():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
