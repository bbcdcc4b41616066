"""The C-ABI library loads without a GPU, exports every symbol include/forkergl_b200.h declares, and refuses to run
(loudly) when there is no CUDA device — there is no CPU fallback in the product path."""
import ctypes as C
import os
import re

import pytest

import parity as P
from forkerrenderer_b200 import binding as B

HEADER = os.path.join(P.REPO, "include", "forkergl_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fgl_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_whole_surface():
    syms = declared_symbols()
    assert len(syms) >= 35
    for must in ("fgl_create", "fgl_draw_mesh", "fgl_draw_screen_space_pixels", "fgl_ssao", "fgl_blur", "fgl_ssaa_resolve", "fgl_read_plane"):
        assert must in syms


@pytest.mark.parametrize("lib", [B.CUDA_LIB, P.ORACLE_LIB])
def test_library_exports_every_declared_symbol(lib):
    if not os.path.exists(lib):
        import __graft_entry__ as g
        g.build()
    L = C.CDLL(lib, mode=C.RTLD_LOCAL)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing


def test_backend_names():
    cuda = C.CDLL(B.CUDA_LIB, mode=C.RTLD_LOCAL)
    cuda.fgl_backend_name.restype = C.c_char_p
    assert cuda.fgl_backend_name() == b"cuda-sm_100a"
    orc = C.CDLL(P.ORACLE_LIB, mode=C.RTLD_LOCAL)
    orc.fgl_backend_name.restype = C.c_char_p
    assert orc.fgl_backend_name() == b"oracle-cpu"


def test_product_library_does_not_link_the_oracle():
    import subprocess
    out = subprocess.run(["ldd", B.CUDA_LIB], stdout=subprocess.PIPE, text=True).stdout
    out += subprocess.run(["ldd", B.HOST_LIB], stdout=subprocess.PIPE, text=True).stdout
    assert "oracle" not in out
    needed = subprocess.run(["readelf", "-d", B.HOST_LIB], stdout=subprocess.PIPE, text=True).stdout
    assert "libforkergl_b200.so" in needed


def test_no_gpu_means_a_loud_failure_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(B.FglError) as e:
        B.product_fgl(0)
    assert "CUDA" in str(e.value) or "cuda" in str(e.value)


def test_default_params_match_the_reference_constants():
    for lib in (B.CUDA_LIB, P.ORACLE_LIB):
        L = C.CDLL(lib, mode=C.RTLD_LOCAL)
        p = B.FglParams()
        L.fgl_default_params(C.byref(p))
        assert p.shadow_mode == B.SHADOW_PCSS                       # shadow.h:16 ships PCSS
        assert p.pcf_filter_size == 0.007 and p.pcss_blocker_filter_size == 0.005   # shadow.h:19-22
        assert abs(p.area_light_size - 2.5) < 1e-7 and abs(p.ssao_radius - 0.075) < 1e-7
        assert p.ssao_range_check == 1 and p.materialize_frame_f32 == 1
