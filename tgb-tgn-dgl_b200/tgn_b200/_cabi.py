"""ctypes binding of the C-ABI in include/tgn_b200.h.

The prototypes are parsed from the header itself, so Python can never drift
from the declared ABI; `declared_symbols()` is what the CPU test-suite checks
against the exported symbols of the shared library.

There is no fallback: if `lib/libtgn_b200.so` is missing, importing the ops
raises.  Build it with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C tgb-tgn-dgl_b200/csrc`.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REPO_DIR = os.path.dirname(_PKG_DIR)
HEADER = os.path.join(_REPO_DIR, "include", "tgn_b200.h")
LIB_PATH = os.path.join(_PKG_DIR, "lib", "libtgn_b200.so")

TGN_OK, TGN_EINVAL, TGN_ECUDA = 0, -1, -2
AGG_LAST, AGG_MEAN = 0, 1
SAMPLE_RECENT, SAMPLE_UNIFORM = 0, 1
SORT_MAX = 8192


class MsgStoreStruct(ctypes.Structure):
    """Mirror of `struct tgn_msgstore` (field order must match the header)."""

    _fields_ = [
        ("num_nodes", ctypes.c_int64),
        ("capacity", ctypes.c_int64),
        ("raw_dim", ctypes.c_int32),
        ("t_is_float", ctypes.c_int32),
        ("ev_src", ctypes.c_void_p),
        ("ev_dst", ctypes.c_void_p),
        ("ev_t", ctypes.c_void_p),
        ("ev_msg", ctypes.c_void_p),
        ("s_perm", ctypes.c_void_p),
        ("d_perm", ctypes.c_void_p),
        ("s_start", ctypes.c_void_p),
        ("s_cnt", ctypes.c_void_p),
        ("s_last", ctypes.c_void_p),
        ("d_start", ctypes.c_void_p),
        ("d_cnt", ctypes.c_void_p),
        ("d_last", ctypes.c_void_p),
    ]


class GemmDesc(ctypes.Structure):
    """Mirror of `struct tgn_gemm_desc` (field order must match the header)."""

    _fields_ = [
        ("a", ctypes.c_void_p), ("b", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("c", ctypes.c_void_p),
        ("m_dev", ctypes.c_void_p), ("k_dev", ctypes.c_void_p),
        ("m", ctypes.c_int32), ("n", ctypes.c_int32), ("k", ctypes.c_int32),
        ("lda", ctypes.c_int32), ("ldb", ctypes.c_int32), ("ldc", ctypes.c_int32),
        ("trans_a", ctypes.c_int32), ("trans_b", ctypes.c_int32),
        ("mode", ctypes.c_int32), ("split_k", ctypes.c_int32),
    ]


_SCALARS = {
    "int32_t": ctypes.c_int32,
    "int64_t": ctypes.c_int64,
    "uint64_t": ctypes.c_uint64,
    "uint32_t": ctypes.c_uint32,
    "float": ctypes.c_float,
}


def _strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object]]]:
    """name -> (restype, argtypes) for every function the header declares."""
    src = _strip_comments(open(path).read())
    protos: Dict[str, Tuple[object, List[object]]] = {}
    for m in re.finditer(r"(?:^|\n)\s*(const\s+char\s*\*|int32_t|int64_t)\s+(tgn_\w+)\s*\(([^;{]*?)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = ctypes.c_char_p if "char" in ret else _SCALARS[ret]
        argtypes: List[object] = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    if "tgn_msgstore" in a:
                        argtypes.append(ctypes.POINTER(MsgStoreStruct))
                    else:
                        argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.replace("const ", "").split()[0]
                    argtypes.append(_SCALARS[ty])
        protos[name] = (restype, argtypes)
    return protos


def declared_symbols() -> List[str]:
    return sorted(parse_header().keys())


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the sm_100a C-ABI library has not been built "
                "(run __graft_entry__.build() or `make -C tgb-tgn-dgl_b200/csrc`). "
                "There is no CPU fallback."
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in parse_header().items():
            fn = getattr(handle, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.tgn_abi_version() != 1:
            raise ImportError("libtgn_b200.so ABI version mismatch")
        _lib = handle
    return _lib


class TgnError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != TGN_OK:
        msg = lib().tgn_last_error()
        kind = {TGN_EINVAL: "TGN_EINVAL", TGN_ECUDA: "TGN_ECUDA"}.get(rc, str(rc))
        raise TgnError(f"{kind}: {msg.decode() if msg else ''}")
