"""Import helper for the *real* reference source checkout (only usable where /root/reference exists, i.e. the build
container).  Used by tools/make_golden.py to pin oracle/ against the reference.  Never imported by the product.
The stubs live in oracle/ref_loader.py (shared with the byte-compiled oracle/_ref that travels to the GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import AttrDict, default_model_config, register_pointcloud  # noqa: E402,F401
from oracle import ref_loader as _rl  # noqa: E402

REF_ROOT = "/root/reference"


def install():
    _rl.install(REF_ROOT)
