"""Synthetic scenes live in the package (object_nerf_b200/synthetic.py) so that bench.py and smoke() do not import the
test package; the tests keep using them under this name."""
from object_nerf_b200.synthetic import *  # noqa: F401,F403
from object_nerf_b200.synthetic import N_CODE, N_VOX_CH, N_OBJ_CH  # noqa: F401
