"""BASELINE configs[3] in miniature: ray-sharded data-parallel training.  Two ranks (torch.distributed, gloo backend so
that it also runs on a one-GPU box: both ranks then share cuda:0; NCCL at N = 2..8 is exercised by bench.py's train
arm) wrap the training system in DistributedDataParallel; each renders its own ray batch through render_rays() (tensor-core
forward + backward); the gradients DDP leaves on every rank must be the mean of the two single-process gradients."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import cases, helpers

pytestmark = pytest.mark.gpu


def _system(inp, dev):
    from torch import nn
    from object_nerf_b200 import Embedding, render_rays, synthetic as S

    class System(nn.Module):
        def __init__(self):
            super().__init__()
            self.coarse = helpers.make_model(inp["weights"]["coarse"], True, dev).train()
            self.fine = helpers.make_model(inp["weights"]["fine"], True, dev).train()
            self.emb = helpers.GridModule(inp["grid"]).to(dev)
            self.lib = S.make_code_library(inp["code_table"]).to(dev)

        def forward(self, b, rand):
            codes = self.lib.lookup(b["instance_ids"])
            c = cases.GRAD_CASE
            out = render_rays({"coarse": self.coarse, "fine": self.fine}, {"xyz": self.emb, "dir": Embedding(3, 4)}, b["rays"],
                              N_samples=c["n_samples"], perturb=c["perturb"], noise_std=c["noise_std"],
                              N_importance=c["n_importance"], embedding_instance=codes, frustum_bound_th=c["frustum_bound_th"],
                              pass_through_mask=b["pass_through_mask"], is_eval=False, precision="bf16", _rand=rand)
            return cases.total_loss(out, b["batch"])
    return System()


def _batch(inp, sl, dev):
    b = {"rays": inp["rays"][sl].to(dev), "instance_ids": inp["instance_ids"][sl].to(dev),
         "pass_through_mask": inp["pass_through_mask"][sl].to(dev),
         "batch": {k: v[sl].to(dev) for k, v in inp["batch"].items()}}
    rand = {k: v[sl].to(dev) for k, v in inp["rand"].items()}
    return b, rand


def _grads(system):
    return [p.grad.detach().clone() for p in system.parameters()]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    ndev = torch.cuda.device_count()
    dev = torch.device("cuda", rank % ndev)
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        inp = cases.build_grad_case(n_rays=96)
        half = slice(rank * 48, (rank + 1) * 48)
        # single-process gradients of BOTH halves (every rank computes them: same code, same inputs)
        singles = []
        for r in range(world):
            s = _system(inp, dev)
            b, rand = _batch(inp, slice(r * 48, (r + 1) * 48), dev)
            s(b, rand).backward()
            singles.append(_grads(s))
        want = [(a + b) / 2 for a, b in zip(*singles)]
        s = _system(inp, dev)
        ddp = torch.nn.parallel.DistributedDataParallel(s, device_ids=[dev.index], broadcast_buffers=False)
        b, rand = _batch(inp, half, dev)
        ddp(b, rand).backward()
        got = _grads(s)
        errs = []
        for (name, _), g, w in zip(s.named_parameters(), got, want):
            scale = w.abs().max().item() + 1e-12
            errs.append(((g - w).abs().max().item() / scale, name, g.abs().max().item(), scale))
        errs.sort(reverse=True)
        ret[rank] = errs[:4]
    finally:
        dist.destroy_process_group()


def test_ddp_gradients_are_the_mean_of_the_per_rank_gradients():
    world = 2
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    # fp32 atomics reorder sums between runs: equality up to accumulation order
    assert len(ret) == world and max(v[0][0] for v in ret.values()) < 2e-3, dict(ret)
