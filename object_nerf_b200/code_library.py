"""Per-object latent codes (reference models/code_library.py:5-28): an embedding table looked up by
instance id.  The lookup is a plain index (host-side orchestration, as in the reference); the codes it
returns are consumed by the fused kernels."""
from torch import nn


class CodeLibrary(nn.Module):
    def __init__(self, model_config):
        super().__init__()
        get = model_config.get if hasattr(model_config, "get") else (lambda k, d: getattr(model_config, k, d))
        self.embedding_instance = nn.Embedding(get("N_max_objs", 64), get("N_obj_code_length", 64))

    def forward(self, inputs):
        out = {}
        if "instance_ids" in inputs:
            out["embedding_instance"] = self.embedding_instance(inputs["instance_ids"].squeeze())
        return out
