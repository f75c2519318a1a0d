"""Loader of the torch C++ extension over the C-ABI (csrc_ext/tgn_torch.cpp -> torch.ops.tgn.*), built in-tree by
`python setup.py build_ext --inplace` (reference README.md:1-2).  `available()` is False when it has not been built;
the Python modules then use their ctypes route to the same entry points (tgn_b200/_cabi.py) -- both are native,
neither is a fallback to anything but libtgn_b200.so."""
import glob
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_loaded = None


def path():
    hits = sorted(glob.glob(os.path.join(_HERE, "_tgn_torch*.so")))
    return hits[0] if hits else None


def available() -> bool:
    global _loaded
    if _loaded is None:
        p = path()
        _loaded = False
        if p is not None:
            from . import _cabi
            _cabi.lib()                      # libtgn_b200.so first: fails loudly if the C-ABI library is missing
            torch.ops.load_library(p)
            _loaded = int(torch.ops.tgn.abi_version()) == 1
    return _loaded
