"""Per-object latent codes (reference models/code_library.py:5-28): an embedding table looked up by instance id.
On a CUDA device the lookup and its gradient (scatter-add of the per-ray code gradients into the table) are kernels of
libonerf_sm100.so (`onerf_code_gather` / `onerf_code_scatter_add`); the parameter keeps the reference's name
(`embedding_instance.weight`), so checkpoints and optimizers are interchangeable."""
import torch
from torch import nn

from . import _lib


class _CodeLookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, ids):
        dev = table.device
        ids = ids.reshape(-1).to(torch.int64).contiguous()
        t = table.detach().contiguous().float()
        out = torch.empty(ids.numel(), t.shape[1], dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().onerf_code_gather(_lib.ctx(dev), t.data_ptr(), ids.data_ptr(), ids.numel(), t.shape[0],
                                                     out.data_ptr(), _lib.stream()))
        ctx.save_for_backward(ids)
        ctx.shape = tuple(t.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        (ids,) = ctx.saved_tensors
        g = g.contiguous().float()
        grad = torch.zeros(ctx.shape, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(_lib.load().onerf_code_scatter_add(_lib.ctx(g.device), g.data_ptr(), ids.data_ptr(), ids.numel(),
                                                          ctx.shape[0], grad.data_ptr(), _lib.stream()))
        return grad, None


class CodeLibrary(nn.Module):
    def __init__(self, model_config):
        super().__init__()
        get = model_config.get if hasattr(model_config, "get") else (lambda k, d: getattr(model_config, k, d))
        self.embedding_instance = nn.Embedding(get("N_max_objs", 64), get("N_obj_code_length", 64))

    def lookup(self, instance_ids: torch.Tensor) -> torch.Tensor:
        """(N,) or (N,1) int64 ids -> (N,64) codes."""
        w = self.embedding_instance.weight
        if w.shape[1] != 64:
            raise RuntimeError("object_nerf_b200 kernels are built for 64-long object codes")
        return _CodeLookup.apply(w, instance_ids)

    def forward(self, inputs):
        out = {}
        if "instance_ids" in inputs:
            out["embedding_instance"] = self.lookup(inputs["instance_ids"])
        return out
