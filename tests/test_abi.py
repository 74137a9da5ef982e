"""CPU: the C-ABI library loads and exports every symbol include/vrcoc.h declares (no compute calls without a GPU);
argument validation errors are reported through the return code + vrcoc_last_error, never by crashing."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "vrcoc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vrcoc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from vrcoc import _lib
    syms = _header_symbols()
    assert len(syms) >= 15
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/vrcoc.h but not exported by libvrcoc.so"
    assert sorted(_lib.SIGNATURES) == syms, "python binding table and header disagree"
    assert "sm_100a" in _lib.version()


def test_conv_desc_layout_matches_header():
    """field order of the ctypes mirror == field order of struct vrcoc_conv_desc"""
    from vrcoc._lib import ConvDesc
    src = open(os.path.join(ROOT, "include", "vrcoc.h")).read()
    body = src[src.index("typedef struct vrcoc_conv_desc {"):src.index("} vrcoc_conv_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S).split("{", 1)[1]
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        parts = decl.replace("*", " ").split(",")
        names.append(parts[0].split()[-1])
        names += [p.strip() for p in parts[1:]]
    assert names == [f[0] for f in ConvDesc._fields_]


def test_validation_errors_do_not_need_a_gpu():
    from vrcoc import _lib
    lib = _lib.lib
    # null pointers -> VRCOC_EINVAL with a message, no crash, no CUDA call
    rc = lib.vrcoc_cluster_core_fwd(None, 0, None, 0, None, 0, None, None, None, None, 1, 1, 8, 16, 16, 2, 2, 2, 2, 0, 0, 0, None)
    assert rc == -1 and b"null pointer" in lib.vrcoc_last_error()
    rc = lib.vrcoc_conv_fwd(None, None)
    assert rc == -1 and b"null descriptor" in lib.vrcoc_last_error()
    with pytest.raises(_lib.VrcocError, match="null descriptor"):
        _lib.check(rc, "conv_fwd")
    d = _lib.ConvDesc()
    d.B = d.H_in = d.W_in = d.H_out = d.W_out = d.C0 = d.O = d.kh = d.kw = d.stride = 1
    d.H_out = 3
    rc = lib.vrcoc_conv_fwd(ctypes.byref(d), None)
    assert rc == -1 and b"inconsistent" in lib.vrcoc_last_error()


def test_new_entry_points_validate_without_a_gpu():
    """vrcoc_patch_embed(_supported), vrcoc_im2col_rows and the row-tap mode of vrcoc_conv_desc (k_order 2): the argument checks
    run before any CUDA call"""
    from vrcoc import _lib
    lib = _lib.lib
    BF16, F32 = _lib.BF16, _lib.F32
    assert lib.vrcoc_patch_embed_supported(BF16, 3, 2, 512, 512, 64, 4) == 1
    assert lib.vrcoc_patch_embed_supported(BF16, 4, 2, 1024, 1024, 64, 4) == 1
    assert lib.vrcoc_patch_embed_supported(F32, 3, 2, 512, 512, 64, 4) == 0          # bf16 only
    assert lib.vrcoc_patch_embed_supported(BF16, 3, 2, 512, 512, 32, 4) == 0         # 64 output channels only
    assert lib.vrcoc_patch_embed_supported(BF16, 7, 2, 512, 512, 64, 4) == 0         # at most 8 input channels
    assert lib.vrcoc_patch_embed_supported(BF16, 3, 2, 512, 480, 64, 4) == 0         # W % 64
    assert lib.vrcoc_patch_embed_supported(BF16, 3, 2, 512, 512, 64, 16) == 0        # patch 4 only
    rc = lib.vrcoc_patch_embed(None, None, 0, None, None, None, None, BF16, 1, 3, 2, 512, 512, 64, 4, None)
    assert rc == -1 and b"patch_embed" in lib.vrcoc_last_error()
    rc = lib.vrcoc_im2col_rows(None, None, BF16, 1, 64, 16, 16, 3, 1, None)
    assert rc == -1 and b"im2col_rows" in lib.vrcoc_last_error()
    # row-tap descriptor with a horizontal kernel extent: rejected with the contract in the message
    buf = ctypes.create_string_buffer(64)
    d = _lib.ConvDesc()
    d.B, d.H_in, d.W_in, d.H_out, d.W_out, d.C0, d.O = 1, 64, 64, 64, 64, 192, 64
    d.kh, d.kw, d.stride, d.pad, d.dil, d.k_order = 3, 3, 1, 1, 1, 2
    d.src0 = d.weight = d.out = ctypes.addressof(buf)
    d.src0_dtype = d.weight_dtype = d.out_dtype = BF16
    d.O_split = 64
    rc = lib.vrcoc_conv_fwd(ctypes.byref(d), None)
    assert rc == -1 and b"row-tap" in lib.vrcoc_last_error()
    d.kw, d.C0 = 1, 100                                                               # C0 % 64
    rc = lib.vrcoc_conv_fwd(ctypes.byref(d), None)
    assert rc == -1 and b"row-tap" in lib.vrcoc_last_error()
    # a row-tap descriptor is not a 1x1 projection for the weight-gradient entry points
    d.C0 = 192
    assert lib.vrcoc_conv1x1_wgrad_workspace(ctypes.byref(d)) == -1


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    import subprocess
    import sys
    env = dict(os.environ, VRCOC_LIB=str(tmp_path / "nope.so"), PYTHONPATH=os.path.join(ROOT, "asy-vrnet_b200"))
    r = subprocess.run([sys.executable, "-c", "import vrcoc"], env=env, capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU / PyTorch fallback" in r.stderr
