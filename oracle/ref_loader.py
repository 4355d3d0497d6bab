"""Import the REAL reference (oracle/_ref: byte-compiled by oracle/build_ref.py, or a source checkout) with the
third-party modules it imports at module scope but never calls on the per-ray path stubbed out.
Test / baseline infrastructure: only tests/, tools/ and bench.py's baseline legs may import this."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BUILD = os.path.join(HERE, "_ref")


class AttrDict(dict):
    """dict with attribute access (the reference indexes its config both ways; OmegaConf stand-in)."""

    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


def available(root=REF_BUILD):
    return os.path.exists(os.path.join(root, "models", "rendering.pyc")) or \
        os.path.exists(os.path.join(root, "models", "rendering.py"))


def _stub(name):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


class _FakePcd:
    def __init__(self, pts):
        self.points = pts


_PCD_REGISTRY = {}
_ORIG_CUDA = (torch.Tensor.cuda, torch.nn.Module.cuda)


def cuda_noop(flag: bool):
    """The reference calls `.cuda()` unconditionally (models/embedding_helper.py:103,125,163,166,193,200,367).  To run it
    on CPU tensors (CPU baseline, CPU-only container) those calls are made no-ops; flag=False restores torch's methods."""
    if flag:
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    else:
        torch.Tensor.cuda, torch.nn.Module.cuda = _ORIG_CUDA


def register_pointcloud(path, pts):
    _PCD_REGISTRY[path] = np.asarray(pts, dtype=np.float64)


def install(root=REF_BUILD, cuda_noop=None):
    """Put the reference on sys.path and stub what is not installed.  cuda_noop (default: no GPU present) turns the
    reference's unconditional `.cuda()` calls into no-ops so its voxel helper runs on CPU tensors."""
    if root not in sys.path:
        sys.path.insert(0, root)
    for name in ("torch_optimizer", "matplotlib", "matplotlib.pyplot", "ipdb", "imageio", "mcubes", "test_tube"):
        _stub(name)
    k = _stub("kornia")

    def create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=None):
        """kornia is not installed: its published create_meshgrid restated (1 x H x W x 2, last dim = (x, y))."""
        xs = torch.linspace(0, width - 1, width)
        ys = torch.linspace(0, height - 1, height)
        if normalized_coordinates:
            xs = (xs / (width - 1) - 0.5) * 2
            ys = (ys / (height - 1) - 0.5) * 2
        base = torch.stack(torch.meshgrid([xs, ys], indexing="ij"), dim=-1)
        return base.permute(1, 0, 2).unsqueeze(0)
    k.create_meshgrid = create_meshgrid  # datasets/ray_utils.py:2,17
    k.__path__ = []
    kl = _stub("kornia.losses")
    kl.ssim = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("kornia.losses.ssim stub"))  # utils/metrics.py:2
    k.losses = kl
    o3d = _stub("open3d")
    io = _stub("open3d.io")
    io.read_point_cloud = lambda p: _FakePcd(_PCD_REGISTRY[p])
    o3d.io = io
    if "pytorch_lightning" not in sys.modules:
        install_lightning_shim()
    if "omegaconf" not in sys.modules:
        oc = _stub("omegaconf")

        class OmegaConf:
            @staticmethod
            def create(d=None):
                return to_attr(d or {})

            @staticmethod
            def merge(*cfgs):
                out = AttrDict()
                for c in cfgs:
                    _merge(out, c)
                return out

            @staticmethod
            def load(path):
                import yaml
                with open(path) as f:
                    return to_attr(yaml.safe_load(f))

            @staticmethod
            def from_cli(args=None):
                return AttrDict()

            @staticmethod
            def to_container(c, resolve=False):
                return dict(c)
        oc.OmegaConf = OmegaConf
    if cuda_noop is None:
        cuda_noop = not torch.cuda.is_available()
    globals()["cuda_noop"](bool(cuda_noop))


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    if isinstance(d, (list, tuple)):
        return [to_attr(v) for v in d]
    return d


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = to_attr(v)


def install_lightning_shim():
    """pytorch_lightning 1.5 is not installed: the few names train.py touches at import / class-definition time, and a
    LightningModule base that is a plain nn.Module (the tests call training_step directly)."""
    pl = _stub("pytorch_lightning")

    class LightningModule(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self.logged = {}
            self.current_epoch = 0
            self.global_step = 0

        def log(self, name, value, *a, **k):
            self.logged[name] = value

        def save_hyperparameters(self, *a, **k):
            pass

    def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kwargs):
        """pytorch_lightning 1.5 semantics for the one call the reference makes (editable_renderer.py:76)."""
        obj = cls(**kwargs)
        ck = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
        obj.load_state_dict(ck["state_dict"], strict=True)
        return obj
    LightningModule.load_from_checkpoint = classmethod(load_from_checkpoint)
    pl.LightningModule = LightningModule
    pl.Trainer = type("Trainer", (), {"__init__": lambda self, *a, **k: None})
    cb = _stub("pytorch_lightning.callbacks")
    cb.ModelCheckpoint = type("ModelCheckpoint", (), {"__init__": lambda self, *a, **k: None})
    cb.LearningRateMonitor = type("LearningRateMonitor", (), {"__init__": lambda self, *a, **k: None})
    pl.callbacks = cb
    lg = _stub("pytorch_lightning.loggers")
    lg.TestTubeLogger = type("TestTubeLogger", (), {"__init__": lambda self, *a, **k: None})
    lg.TensorBoardLogger = type("TensorBoardLogger", (), {"__init__": lambda self, *a, **k: None})
    pl.loggers = lg
    pl.seed_everything = lambda seed=None, **k: torch.manual_seed(seed or 0)


def default_model_config(use_voxel=True):
    # values of config/default_conf.yml:7-36
    return AttrDict(
        use_voxel_embedding=use_voxel, N_freq_xyz=10, N_freq_dir=4, N_freq_voxel=6, D=8, W=256,
        skips=[4], N_scn_voxel_size=16, inst_D=4, inst_W=128, inst_skips=[2], N_obj_voxel_size=8,
        N_samples=64, N_importance=64, frustum_bound=0.05, use_disp=False, perturb=1, noise_std=1,
        N_max_objs=64, N_obj_code_length=64, N_max_voxels=800000,
    )


# ------------------------------------------------------------------------------------------------
# the reference's own modules filled with a synthetic scene (object_nerf_b200/synthetic.py)
# ------------------------------------------------------------------------------------------------
def ref_model(w, use_voxel, device="cpu"):
    """Reference ObjectNeRF (models/nerf_model.py) holding the weight dict `w`."""
    from models.nerf_model import ObjectNeRF
    from object_nerf_b200.synthetic import REF_NAMES
    m = ObjectNeRF(default_model_config(use_voxel))
    sd = {}
    for k, (W, b) in w.items():
        sd[REF_NAMES[k] + ".weight"] = W
        sd[REF_NAMES[k] + ".bias"] = b
    m.load_state_dict(sd, strict=True)
    return m.eval().to(device)


def ref_voxel_embedding(grid, device="cpu"):
    """Build the reference EmbeddingVoxel on a throw-away cloud, then overwrite the buffers the hot path
    reads with the synthetic grid (the cold-path constructor is pinned separately in fixture 'gridbuild')."""
    from models.embedding_helper import EmbeddingVoxel
    register_pointcloud("tiny.ply", np.array([[0.0, 0, 0], [0.2, 0.2, 0.2]]))
    extra = AttrDict(pcd_path="tiny.ply", scene_center=[0, 0, 0], scale_factor=1.0, voxel_size=0.1,
                     neighbor_marks=3)
    emb = EmbeddingVoxel(24, 6, grid["table"].shape[0], extra)
    emb.voxel_size = grid["voxel_size"].clone().to(device)
    emb.voxel_offset = grid["offset"].clone().to(device)
    emb.voxel_shape = grid["shape"].clone().to(device)
    emb.voxel_idx_map = grid["idx_map"].clone().to(device)
    emb = emb.to(device)
    with torch.no_grad():
        emb.embedding_space_ftr.weight.copy_(grid["table"])
    return emb


def ref_render_setup(weights, grid, device="cpu"):
    """-> (models, embeddings) as the reference's render_rays() takes them."""
    from models.embedding_helper import Embedding
    models = {k: ref_model(w, grid is not None, device) for k, w in weights.items()}
    emb_xyz = ref_voxel_embedding(grid, device) if grid is not None else Embedding(3, 10)
    return models, {"xyz": emb_xyz, "dir": Embedding(3, 4)}
