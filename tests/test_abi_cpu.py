"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a without a GPU, loads, and exports
exactly the symbols include/dynllava_b200.h declares; the product path fails loudly without a CUDA device."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from dynamic_llava_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", params=["bf16", "f16"])
def so_path(request):
    """both builds of the same sources: libdynllava_b200.so (bfloat16) and libdynllava_b200_f16.so (float16)"""
    build.build_extension()
    return build.LIB if request.param == "bf16" else build.LIB_F16


def test_library_builds_and_loads(so_path):
    assert os.path.exists(so_path)
    dll = ctypes.CDLL(so_path)
    dll.dl_version.restype = ctypes.c_int
    assert dll.dl_version() == 7
    dll.dl_element_type.restype = ctypes.c_int
    assert dll.dl_element_type() == (1 if so_path.endswith("_f16.so") else 0)
    dll.dl_last_error_string.restype = ctypes.c_char_p
    assert isinstance(dll.dl_last_error_string(), bytes)


def test_every_declared_symbol_is_exported(so_path):
    protos = _lib.parse_header()
    assert len(protos) >= 30
    dll = ctypes.CDLL(so_path)
    missing = [n for n in protos if not hasattr(dll, n)]
    assert not missing, missing
    # and nothing is exported under the dl_ prefix that the header does not declare
    out = subprocess.run(["nm", "-D", "--defined-only", so_path], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\b(dl_\w+)\b", out))
    assert exported == set(protos), exported ^ set(protos)


def test_header_cites_reference_for_each_entry_point():
    src = open(_lib.HEADER).read()
    assert src.count("DML:") + src.count("CU:") + src.count("ARCH:") + src.count("CTL:") >= 20
    assert "extern \"C\"" in src and "torch" not in src.replace("PyTorch", "")


def test_sass_contains_blackwell_native_instructions(so_path):
    """tcgen05.mma -> UTCHMMA, TMA -> UTMALDG/UBLKCP, tcgen05.ld -> LDTM (B200_PROFILING.md)"""
    out = subprocess.run(["cuobjdump", "-sass", so_path], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM"):
        assert mnemonic in out, mnemonic
    assert "HGMMA" not in out


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from dynamic_llava_b200 import DynamicLlavaConfig, DynamicLlavaLlamaForCausalLM
    m = DynamicLlavaLlamaForCausalLM(DynamicLlavaConfig.from_preset("toy"), device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 4, dtype=torch.long))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dynamic_llava_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith(".py"):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn
