"""The C-ABI library loads without a GPU, exports every symbol include/vidil_b200.h declares, and fails loudly —
never falls back — when asked to compute on a machine without a B200."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from vidil_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vidil_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vidil_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for must in ["vidil_vit_forward", "vidil_clip_forward", "vidil_sim_topk", "vidil_vit_forward_host",
                 "vidil_encoder_create", "vidil_encoder_load", "vidil_last_error"]:
        assert must in names


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/vidil_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(declared_functions()), "ctypes signature table out of sync with the header"


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "vidil_b200.h"\nint main(void){ vidil_encoder_cfg c; (void)c; return VIDIL_B200_ABI_VERSION - 2; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o",
                        str(tmp_path / "t.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_sass_contains_tcgen05_and_tma():
    """UTCHMMA = tcgen05.mma, UTMALDG = TMA load, LDTM = tcgen05.ld (guides/B200_PROFILING.md)."""
    sass = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, f"{mnemonic} missing from the SASS of libvidil_b200.so"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_entry_points_fail_loudly_without_a_gpu():
    lib = _lib.load()
    assert lib.vidil_abi_version() == _lib.ABI_VERSION
    cfg = _lib.EncoderCfg(img_size=224, patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_dim=4096,
                          ln_eps=1e-6, act=0, patch_bias=1, pre_ln=0, proj_dim=0, dtype=0, cta_group=0)
    h = ctypes.c_void_p()
    assert lib.vidil_encoder_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert not h
    assert len(_lib.last_error()) > 0
    buf = (ctypes.c_float * 64)()
    assert lib.vidil_op_layernorm(buf, buf, buf, buf, 1, 128, 1e-6, None) != 0


def test_bad_configurations_are_rejected_with_a_message():
    lib = _lib.load()
    h = ctypes.c_void_p()
    for kw in [dict(patch_size=15), dict(embed_dim=1000), dict(num_heads=8), dict(dtype=7), dict(act=9), dict(cta_group=3)]:
        base = dict(img_size=224, patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_dim=4096, ln_eps=1e-6,
                    act=0, patch_bias=1, pre_ln=0, proj_dim=0, dtype=0, cta_group=0)
        base.update(kw)
        cfg = _lib.EncoderCfg(**base)
        assert lib.vidil_encoder_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
        assert _lib.last_error()


def _med_cfg(**kw):
    base = dict(vocab_size=30524, max_positions=512, hidden=768, depth=12, num_heads=12, mlp_dim=3072, encoder_width=1024,
                ln_eps=1e-12, lm_head=1, cls_out=0, dtype=0, cta_group=0)
    base.update(kw)
    return _lib.MedCfg(**base)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_text_stack_entry_points_fail_loudly_without_a_gpu():
    lib = _lib.load()
    h = ctypes.c_void_p()
    cfg = _med_cfg()
    assert lib.vidil_med_create(ctypes.byref(cfg), ctypes.byref(h)) != 0 and not h
    assert len(_lib.last_error()) > 0
    assert lib.vidil_med_generate(None, None, 1, 1, None, 4, 3, 20, 5, 102, 0, 1.0, None, None, None, None, 0, None) != 0
    assert lib.vidil_med_forward(None, None, 1, 1, None, None, None, 0, 1, 8, 1, None, None, None, None, 0, None) != 0
    assert lib.vidil_med_sample(None, None, 1, 1, None, 4, 20, 5, 102, 0, 50, 0.9, 1.1, None, None, None, None, None, 0, None) != 0
    buf = (ctypes.c_float * 64)()
    prompt = (ctypes.c_int32 * 4)(1, 2, 3, 4)
    # argument checks come first and name the offending values
    assert lib.vidil_op_beam_search(buf, 3, 1, 9, 8, ctypes.cast(prompt, ctypes.c_void_p), 4, 7, 0, 1, 0, 1.0, buf, buf, buf, buf, 64,
                                    None) != 0
    assert "num_beams" in _lib.last_error()
    # workspace queries are pure arithmetic and work without a device
    assert lib.vidil_op_beam_search_workspace_bytes(1024, 3, 20) > 1024 * 3 * 20 * 4
    assert lib.vidil_med_generate_workspace_bytes(None, 8, 197, 3, 20, 4) == 0


def test_text_stack_bad_configurations_are_rejected_with_a_message():
    lib = _lib.load()
    h = ctypes.c_void_p()
    for kw in [dict(vocab_size=30523), dict(hidden=700), dict(num_heads=8), dict(encoder_width=1000), dict(mlp_dim=3000), dict(dtype=5),
               dict(cls_out=-1), dict(depth=0), dict(cta_group=4)]:
        cfg = _med_cfg(**kw)
        assert lib.vidil_med_create(ctypes.byref(cfg), ctypes.byref(h)) != 0, kw
        assert _lib.last_error(), kw
