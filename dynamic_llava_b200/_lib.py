"""ctypes binding of libdynllava_b200.so (bfloat16) and libdynllava_b200_f16.so (float16).  Signatures are parsed from include/dynllava_b200.h, so the header is the
single source of truth for the C ABI.  There is NO fallback: if the library is missing or a call fails, this raises."""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "dynllava_b200.h")
LIB_PATH = os.path.join(HERE, "csrc", "libdynllava_b200.so")
LIB_PATH_F16 = os.path.join(HERE, "csrc", "libdynllava_b200_f16.so")

_CTYPES = {
    "int": ctypes.c_int,
    "float": ctypes.c_float,
    "int64_t": ctypes.c_int64,
    "uint64_t": ctypes.c_uint64,
    "uint32_t": ctypes.c_uint32,
    "int32_t": ctypes.c_int32,
    "cudaStream_t": ctypes.c_void_p,
}


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object]]]:
    """-> {name: (restype, [argtypes])} for every `dl_*` prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(?m)^\s*((?:const\s+)?[\w]+\s*\*?)\s*(dl_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if "*" in ret:
            restype = ctypes.c_char_p if "char" in ret else ctypes.c_void_p
        else:
            restype = _CTYPES[ret.replace("const", "").strip()]
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_char_p if re.match(r"const\s+char", a) else ctypes.c_void_p)
                else:
                    ty = a.replace("const", "").split()[0]
                    argtypes.append(_CTYPES[ty])
        protos[name] = (restype, argtypes)
    return protos


class DynLlavaLib:
    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} not found: the CUDA extension is required (no CPU/eager fallback exists). "
                "Build it with `python -m dynamic_llava_b200.build`."
            )
        self._dll = ctypes.CDLL(path)
        self.protos = parse_header()
        for name, (restype, argtypes) in self.protos.items():
            fn = getattr(self._dll, name)  # AttributeError here == header/library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
            setattr(self, "_" + name, fn)

    def __getattr__(self, name):
        # dl.foo(...) -> checked call of dl_foo
        raw = self.__dict__.get("_dl_" + name)
        if raw is None:
            raise AttributeError(name)

        def call(*args):
            rc = raw(*args)
            if raw.restype is ctypes.c_int and rc != 0 and name not in ("version", "element_type", "device_is_sm100", "skinny_gemm_splits"):
                raise RuntimeError(f"dl_{name} failed ({rc}): {self._dl_last_error_string().decode()}")
            return rc

        self.__dict__[name] = call
        return call


_LIBS: Dict[str, DynLlavaLib] = {}


def lib(dtype=None) -> DynLlavaLib:
    """The library whose element type is `dtype` (torch.bfloat16 — the default — or torch.float16; strings accepted)."""
    key = "f16" if str(dtype) in ("torch.float16", "float16", "fp16", "f16") else "bf16"
    if dtype is not None and key == "bf16" and str(dtype) not in ("torch.bfloat16", "bfloat16", "bf16"):
        raise RuntimeError(f"no kernel library computes in {dtype}: load the model in torch.bfloat16 or torch.float16")
    if key not in _LIBS:
        L = DynLlavaLib(LIB_PATH_F16 if key == "f16" else LIB_PATH)
        if L.element_type() != (1 if key == "f16" else 0):
            raise RuntimeError(f"{L._dll._name}: element type {L.element_type()} does not match the requested {key} build")
        _LIBS[key] = L
    return _LIBS[key]
