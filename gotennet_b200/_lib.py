"""ctypes binding of libgotennet_b200.so, generated from include/gotennet_b200.h.

The prototypes are parsed from the header itself, so the Python side cannot drift
from the C ABI.  Loading fails loudly (no CPU fallback, no alternative backend).
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

from . import _build

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "gotennet_b200.h")

_CTYPES = {
    "int": ctypes.c_int,
    "int64_t": ctypes.c_int64,
    "int32_t": ctypes.c_int32,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "void": None,
}


class GotenError(RuntimeError):
    pass


def parse_header(path: str = HEADER) -> Dict[str, Tuple[str, List[Tuple[str, str]]]]:
    """{name: (return_type, [(type, argname), ...])} for every `goten_*` prototype."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    text = re.sub(r"^\s*#[^\n]*$", " ", text, flags=re.M)
    text = text.replace('extern "C" {', " ")
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w \*]*?)\b(goten_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        params: List[Tuple[str, str]] = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                params.append((mm.group(1).strip(), mm.group(2)))
        protos[name] = (ret, params)
    return protos


def _to_ctype(t: str):
    t = t.replace("const ", "").strip()
    if t.endswith("*"):
        base = t[:-1].strip()
        if base == "char":
            return ctypes.c_char_p
        return ctypes.c_void_p
    return _CTYPES[t]


class _Lib:
    def __init__(self):
        path = _build.lib_path()
        if not os.path.exists(path):
            raise GotenError(
                f"{path} is missing: build it with `python -m gotennet_b200._build` "
                "(needs nvcc); gotennet_b200 has no CPU or PyTorch fallback."
            )
        self.path = path
        self.profile = None
        self.cdll = ctypes.CDLL(path)
        self.protos = parse_header()
        for name, (ret, params) in self.protos.items():
            fn = getattr(self.cdll, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = _to_ctype(ret) if ret != "const char*" else ctypes.c_char_p
            fn.argtypes = [_to_ctype(t) for t, _ in params]
        ver = self.cdll.goten_abi_version()
        if ver != 1:
            raise GotenError(f"ABI version mismatch: library {ver}, binding 1")

    def call(self, name: str, *args):
        if self.profile is not None:  # bench.py: CUDA events around every entry point (never on by default)
            import torch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = getattr(self.cdll, name)(*args)
            e1.record()
            self.profile.append((name, args, e0, e1))
        else:
            rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            raise GotenError(f"{name} failed: {self.cdll.goten_last_error().decode()}")

    def raw(self, name: str):
        return getattr(self.cdll, name)


_LIB = None


def lib() -> _Lib:
    global _LIB
    if _LIB is None:
        _LIB = _Lib()
    return _LIB
