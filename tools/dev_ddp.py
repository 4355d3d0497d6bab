import os, sys, socket
sys.path.insert(0, "/root/repo")
import torch, torch.distributed as dist, torch.multiprocessing as mp
from tests import cases
from tests.test_gpu_ddp import _system, _batch, _grads

def worker(rank, world, port, bucket_view):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dev = torch.device("cuda", rank % torch.cuda.device_count()); torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    inp = cases.build_grad_case(n_rays=96)
    s = _system(inp, dev)
    b, rand = _batch(inp, slice(rank * 48, rank * 48 + 48), dev)
    s(b, rand).backward()
    single = {n: p.grad.clone() for n, p in s.named_parameters()}
    s = _system(inp, dev)
    ddp = torch.nn.parallel.DistributedDataParallel(s, device_ids=[dev.index], broadcast_buffers=False, gradient_as_bucket_view=bucket_view)
    ddp(b, rand).backward()
    torch.cuda.synchronize()
    if rank == 0:
        z = [n for n, p in s.named_parameters() if p.grad.abs().max().item() == 0 and single[n].abs().max().item() > 0]
        nz = [n for n, p in s.named_parameters() if p.grad.abs().max().item() > 0]
        print("bucket_view", bucket_view, "zero:", len(z), z[:6], "... nonzero:", len(nz), nz[:3], nz[-3:])
    dist.destroy_process_group()

if __name__ == "__main__":
    for bv in (False, True):
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]
        mp.spawn(worker, args=(2, port, bv), nprocs=2, join=True)
