"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, the kernels' window geometry equals the reference's gather map, the drop-in module mirrors
the reference interface (names, kwargs, errors).  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from oracle import panoswin_oracle as O

import panoswintransformerobjectdetection_b200 as P
from panoswintransformerobjectdetection_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "panoswin_b200.h")).read()
    return sorted(set(re.findall(r"PSW_API\s+(?:const\s+char\*|int64_t|int)\s+(psw_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    names = _header_symbols()
    assert len(names) >= 11 and "psw_window_attn_fwd" in names and "psw_linear_fwd" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/panoswin_b200.h but not exported"
    assert lib.psw_abi_version() == _lib.ABI_VERSION
    # the ctypes table binds exactly the header's functions (besides the error-string getter)
    assert sorted(list(_lib.SIGNATURES) + ["psw_last_error_string"]) == names


def test_library_does_not_leak_cudart_symbols():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.library_path()], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert exported and all(s.startswith("psw_") for s in exported), exported


def test_argument_errors_do_not_need_a_gpu(lib):
    assert lib.psw_layernorm_fwd(None, None, None, None, None, 4, 96, 0, 1e-5, 0, 0, None) == -1
    assert b"null" in lib.psw_last_error_string()
    hp, wp = C.c_int(), C.c_int()
    assert lib.psw_window_source_map(4, 8, 7, 7, 1, None, 0, C.byref(hp), C.byref(wp)) == -1   # shift >= window
    # every entry point validates before it touches the device: null pointers, unsupported shapes
    assert lib.psw_layernorm2_fwd(None, None, None, None, None, None, None, None, 4, 96, 0, 1e-5, 1e-5, 1, None) == -1
    assert b"psw_layernorm2_fwd" in lib.psw_last_error_string()
    assert lib.psw_mlp_fused_fwd(None, None, None, None, None, None, 128, 96, 384, None) == -1
    assert b"psw_mlp_fused_fwd" in lib.psw_last_error_string()
    fake = C.c_void_p(1 << 20)                                   # never dereferenced: the shape check comes first
    assert lib.psw_mlp_fused_fwd(fake, fake, None, fake, None, fake, 128, 96, 384, None) == -1   # b1 / b2 are required
    assert lib.psw_mlp_fused_fwd(fake, fake, fake, fake, fake, fake, 128, 192, 768, None) == -2
    assert b"C = 96" in lib.psw_last_error_string()


def _source_map(lib, H, W, ws, s, pano):
    hp, wp = C.c_int(), C.c_int()
    assert lib.psw_window_source_map(H, W, ws, s, pano, None, 0, C.byref(hp), C.byref(wp)) == 0
    m = np.zeros((hp.value, wp.value), dtype=np.int32)
    assert lib.psw_window_source_map(H, W, ws, s, pano, m.ctypes.data, m.size, C.byref(hp), C.byref(wp)) == 0
    return m


def test_kernel_geometry_equals_reference_gather_map(lib, known_answers):
    """psw::source_token (used by both attention kernels) against the maps dumped from the reference's
    WindowTransition with the arange trick (oracle/make_golden.py)."""
    keys = [k for k in known_answers.files if k.startswith("pano_src_")]
    for k in keys:
        H, W, s = (int(v) for v in k.split("_")[2:])
        ref = known_answers[k]
        m = _source_map(lib, H, W, 7, s, 1)
        assert (m[:ref.shape[0], :ref.shape[1]] == ref).all(), k
        assert (m[ref.shape[0]:] == -1).all() and (m[:, ref.shape[1]:] == -1).all()   # window padding


@pytest.mark.parametrize("hw", [(128, 256), (64, 128), (32, 64), (16, 32), (25, 50), (13, 25), (99, 199)])
def test_kernel_geometry_is_a_bijection(lib, hw):
    H, W = hw
    for s in (0, 3):
        m = _source_map(lib, H, W, 7, s, 1).reshape(-1)
        real = np.sort(m[m >= 0])
        assert real.size == H * W and (real == np.arange(H * W)).all()
        assert np.array_equal(_source_map(lib, H, W, 7, s, 1)[:2 * H, :(W + 1) // 2], O.pano_source_index(H, W, s).numpy())


def test_planar_geometry(lib):
    for (H, W, s) in [(12, 31, 3), (20, 16, 0), (14, 14, 3)]:
        m = _source_map(lib, H, W, 7, s, 0)
        Hp, Wp = m.shape
        base = -np.ones((Hp, Wp), dtype=np.int64)
        base[:H, :W] = np.arange(H * W).reshape(H, W)
        want = base[(np.arange(Hp) + s) % Hp][:, (np.arange(Wp) + s) % Wp]
        assert np.array_equal(m, want)


def test_host_constants_match_reference(known_answers):
    assert np.array_equal(P.make_relative_position_index(7).numpy(), known_answers["rel_index_7"])
    assert np.array_equal(P.make_relative_position_index(3).numpy(), known_answers["rel_index_3"])
    for H, W in [(2, 4), (128, 256), (13, 25)]:
        assert np.array_equal(P.make_uv_hw2(H, W).numpy(), known_answers[f"uv_{H}_{W}"])
    for k in [k for k in known_answers.files if k.startswith("planar_mask_")]:
        H, W, s = (int(v) for v in k.split("_")[2:])
        assert np.array_equal(P.planar_attention_mask(H, W, 7, s).numpy(), known_answers[k])
    with pytest.raises(ValueError):
        P.make_uv_hw2(8, 4)


def test_state_dict_names_are_the_reference_names():
    meta, _ = load_golden("tiny_pano")          # keys recorded from the real reference (strict load)
    cfg = meta["cfg"]
    m = P.SimplePanoSwinTransformer(embed_dim=cfg["embed_dim"], depths=list(cfg["depths"]), num_heads=list(cfg["num_heads"]), ape=True)
    assert sorted(m.state_dict().keys()) == meta["keys"]
    res = m.load_state_dict(O.make_state_dict(cfg, 1), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    t = PANOSWIN_T = P.SimplePanoSwinTransformer(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], ape=True)
    assert sum(p.numel() for p in t.parameters()) == 27_657_876       # SURVEY.md §8c [probe]
    assert t.num_features == [96, 192, 384, 768]


def test_registry_and_constructor_contract():
    assert P.BACKBONES.get("SimplePanoSwinTransformer") is P.SimplePanoSwinTransformer
    m = P.build_backbone(dict(type="SimplePanoSwinTransformer", embed_dim=32, depths=[2, 2], num_heads=[1, 2],
                              out_indices=(0, 1), ape=True, drop_path_rate=0.1, frozen_stages=-1, use_checkpoint=False))
    assert isinstance(m, torch.nn.Module) and m.out_indices == (0, 1) and m.pano_mode is True
    with pytest.raises(TypeError):                            # unknown keys fail like plain Python in the reference
        P.build_backbone(dict(type="SimplePanoSwinTransformer", emb_conv_type="x"))
    with pytest.raises(TypeError):
        m.init_weights(pretrained=123)
    m.init_weights(None)
    assert float(m.layers[0].blocks[0].attn.qkv.bias.abs().sum()) == 0.0
    assert m.train(False) is None and m.training is False
    with pytest.raises(AssertionError):                       # reference :450
        P.backbone.PanoSwinTransformerBlock(dim=32, num_heads=1, window_size=7, shift_size=7)
    with pytest.raises(AssertionError):                       # reference :284
        P.backbone.WindowAttention(dim=30, window_size=7, num_heads=4)
    # the reference's DEFAULT constructor (depths=[2,2,7,2], :785) builds: the odd stage ends in a PitchAttentionModule
    d = P.SimplePanoSwinTransformer()
    assert isinstance(d.layers[2].blocks[6], P.backbone.PitchAttentionModule) and len(d.layers[2].blocks) == 7
    assert d.pano_mode is True and d.ape is False


def test_odd_depth_state_dict_and_mode_switch():
    """Odd stage depths append a PitchAttentionModule with the reference's parameter names (strict load of the
    state_dict the real reference accepted, tests/golden/planar_odd.npz) and follow set_pano_mode like any block."""
    meta, _ = load_golden("planar_odd")
    cfg = meta["cfg"]
    m = P.SimplePanoSwinTransformer(embed_dim=cfg["embed_dim"], depths=list(cfg["depths"]), num_heads=list(cfg["num_heads"]),
                                    out_indices=tuple(cfg["out_indices"]), ape=cfg["ape"], pano_mode=cfg["pano_mode"])
    res = m.load_state_dict(O.make_state_dict(cfg, meta["param_seed"]), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    pam = m.layers[0].blocks[2]
    assert isinstance(pam, P.backbone.PitchAttentionModule) and pam.pano_mode is False
    assert sorted(k for k, _ in pam.named_parameters()) == sorted(
        ["proj.weight", "proj.bias", "sphere_position_alpha_table_Te", "sphere_position_beta_table_Te", "mlp.fc1.weight",
         "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias", "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias",
         "q_linear.weight", "q_linear.bias", "k_linear.weight", "k_linear.bias", "v_linear.weight", "v_linear.bias"])
    m.set_pano_mode(True)
    assert pam.pano_mode is True


def test_set_pano_mode_propagates():
    m = P.SimplePanoSwinTransformer(embed_dim=32, depths=[2, 2], num_heads=[1, 2], out_indices=(0, 1), ape=True)
    m.set_pano_mode(False)
    assert not any(b.pano_mode or b.attn.pano_mode or b.window_transition.pano_mode for l in m.layers for b in l.blocks)
    m.switch_pano_mode()
    assert all(b.pano_mode and b.attn.pano_mode for l in m.layers for b in l.blocks)
    assert [b.shift_size for b in m.layers[0].blocks] == [0, 3]


def test_no_cpu_fallback():
    m = P.SimplePanoSwinTransformer(embed_dim=32, depths=[2, 2], num_heads=[1, 2], out_indices=(0, 1), ape=True)
    m.eval()
    with pytest.raises(P.ops.PanoSwinB200Error):
        m(torch.rand(1, 3, 64, 128))
    with pytest.raises(P.ops.PanoSwinB200Error):
        P.ops.linear(torch.rand(4, 8), torch.rand(8, 8))


def test_checkpoint_loading_strips_prefixes(tmp_path):
    cfg = O.make_config(embed_dim=32, depths=(2, 2), num_heads=(1, 2), out_indices=(0, 1))
    sd = O.make_state_dict(cfg, 5)
    path = tmp_path / "ckpt.pth"
    torch.save({"state_dict": {"module." + k: v for k, v in sd.items()}}, path)
    m = P.SimplePanoSwinTransformer(embed_dim=32, depths=[2, 2], num_heads=[1, 2], out_indices=(0, 1), ape=True)
    m.init_weights(str(path))
    assert torch.equal(m.layers[1].blocks[1].attn.sphere_position_beta_table_Te.data,
                       sd["layers.1.blocks.1.attn.sphere_position_beta_table_Te"])
