"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

TEST INFRASTRUCTURE ONLY.  Usage (container with /root/reference mounted):

    python -m oracle.make_golden            # writes tests/golden/<case>.npz and tests/golden/fpn_<case>.npz
    python -m oracle.make_golden --fpn      # only the FPN-neck fixtures

Each fixture stores the case definition (config, parameter seed, image seed/shape/kind) and the
reference's outputs: full stage features for the small cases, a strided sample plus L2 norms for
PanoSwin-T at 512x1024 (full maps would be 24 MB each).  Parameters and images are re-derived from
the seeds with `oracle.panoswin_oracle.make_state_dict / make_image` (numpy legacy RandomState, so
they are identical on every machine); loading them into the reference with strict=True pins the
state_dict key names.  Also stores the reference's in-file known answers.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

from . import fpn_oracle as FO
from . import panoswin_oracle as O
from . import ref_loader as R

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (config, image shape, image kind, full outputs?)
    "tiny_pano": (O.make_config(embed_dim=32, depths=(2, 2, 2, 2), num_heads=(1, 2, 4, 8)), (2, 3, 64, 128), "rand", True),
    "odd_pano": (O.make_config(embed_dim=32, depths=(2, 2, 2, 2), num_heads=(1, 2, 4, 8)), (1, 3, 98, 197), "randn", True),
    "hd_var_pano": (O.make_config(embed_dim=24, depths=(2, 2, 2, 2), num_heads=(1, 1, 2, 2)), (1, 3, 64, 128), "rand", True),
    "planar": (O.make_config(embed_dim=32, depths=(2, 2, 2, 2), num_heads=(1, 2, 4, 8), pano_mode=False, ape=False),
               (2, 3, 45, 123), "rand", True),
    "planar_tall": (O.make_config(embed_dim=32, depths=(2, 2), num_heads=(1, 2), out_indices=(0, 1), pano_mode=False, ape=True),
                    (1, 3, 78, 64), "randn", True),
    # planar mode with ODD depths: every stage ends in a PitchAttentionModule (the reference default depths=[2,2,7,2]
    # path; it only executes in planar mode)
    "planar_odd": (O.make_config(embed_dim=32, depths=(3, 1, 2), num_heads=(1, 2, 4), out_indices=(0, 1, 2), pano_mode=False,
                                 ape=False), (2, 3, 60, 100), "rand", True),
    # BASELINE.json configs[3] family (embed_dim 128, heads 4-8-16-32) at a size the CPU reference finishes quickly
    "panoswin_b_shaped": (O.make_config(embed_dim=128, depths=(2, 2, 2, 2), num_heads=(4, 8, 16, 32)), (1, 3, 224, 448), "rand", False),
    "panoswin_t_512": (O.PANOSWIN_T, (1, 3, 512, 1024), "rand", False),
    "panoswin_t_512_randn": (O.PANOSWIN_T, (1, 3, 512, 1024), "randn", False),
}
SAMPLE_STRIDE = 997          # prime: samples walk through every channel / row / column phase
PARAM_SEED, IMAGE_SEED = 1, 2
FPN_CASES = ("tiny_pano", "panoswin_t_512")                 # backbone cases whose outputs also go through the FPN neck
FPN_SEED, FPN_OUT, FPN_LEVELS = 3, 256, 5


def sample(t: torch.Tensor) -> np.ndarray:
    return t.reshape(-1)[::SAMPLE_STRIDE].numpy().copy()


def fpn_golden():
    """tests/golden/fpn_<case>.npz: the unmodified reference FPN on the unmodified reference backbone's outputs."""
    for name in FPN_CASES:
        cfg, shape, kind, full = CASES[name]
        model = R.build_reference_model(cfg, O.make_state_dict(cfg, PARAM_SEED))
        outs = R.reference_forward(model, O.make_image(shape, IMAGE_SEED, kind))
        chans = [int(o.shape[1]) for o in outs]
        fsd = FO.make_fpn_state(chans, FPN_OUT, FPN_SEED)
        fpn = R.build_reference_fpn(chans, fsd, FPN_OUT, FPN_LEVELS)
        with torch.no_grad():
            levels = fpn(tuple(outs))
        rec = {"meta": np.array(json.dumps(dict(case=name, in_channels=chans, out_channels=FPN_OUT, num_outs=FPN_LEVELS,
                                               fpn_seed=FPN_SEED, full=full, stride=SAMPLE_STRIDE, n_out=len(levels),
                                               keys=sorted(fsd.keys()))))}
        for i, o in enumerate(levels):
            rec[f"out{i}_shape"] = np.array(o.shape)
            rec[f"out{i}_norm"] = np.array(float(o.double().norm()))
            rec[f"out{i}"] = o.numpy() if full else sample(o)
        path = os.path.join(GOLDEN_DIR, f"fpn_{name}.npz")
        np.savez_compressed(path, **rec)
        print(f"fpn_{name}: {len(levels)} levels {[tuple(o.shape) for o in levels]}, {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    if "--fpn" in sys.argv[1:]:                            # only the FPN fixtures (the others are unchanged)
        fpn_golden()
        return
    only = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--only=")]
    ref = R.load_reference()
    if only:                                                # regenerate just the named end-to-end cases
        end_to_end(only)
        return
    # --- the reference's own known answers ---------------------------------------------------
    kat = {
        "rel_index_3": ref.make_relative_position_index(3).numpy(),
        "rel_index_7": ref.make_relative_position_index(7).numpy(),
        "uv_2_4": ref.make_uv_hw2(2, 4).numpy(),
        "uv_128_256": ref.make_uv_hw2(128, 256).numpy(),
        "uv_13_25": ref.make_uv_hw2(13, 25).numpy(),
    }
    # pano shift gather maps (arange trick) incl. the shapes of _test_WindowTransition (:1276-1283)
    for (H, W, s) in [(4, 8, 0), (4, 8, 3), (13, 25, 3), (7, 13, 3), (15, 29, 0), (6, 11, 3), (16, 32, 3)]:
        t = ref.WindowTransition(shift_size=s, pano_mode=True)
        x = torch.arange(1, H * W + 1, dtype=torch.float32).view(1, H, W, 1)
        y = t(x)
        assert torch.equal(t(y, reverse=True), x), "reference round trip must be the identity"
        kat[f"pano_src_{H}_{W}_{s}"] = (y[0, ..., 0].long() - 1).numpy()
    import importlib
    gc = sys.modules["lzx.models.great_circle"]
    uv1 = torch.tensor([[-77, 39], [121.489, 31.225]]) / 180 * np.pi
    uv2 = torch.tensor([[116.4, 39.9]] * 2) / 180 * np.pi
    kat["haversine_cities_km"] = (gc.haversine22(uv1, uv2) * 6400).numpy()
    g = torch.Generator().manual_seed(7)
    uvr = torch.stack([torch.rand(5, 49, generator=g) * 2 * np.pi - np.pi, torch.rand(5, 49, generator=g) * np.pi - np.pi / 2], -1)
    kat["haversine_in"] = uvr.numpy()
    kat["haversine_out"] = gc.haversine22(uvr, uvr).numpy()
    for (H, W, s) in [(12, 31, 3), (20, 16, 3)]:
        layer = ref.BasicLayer(dim=8, depth=2, num_heads=1, window_size=7, pano_mode=False)
        kat[f"planar_mask_{H}_{W}_{s}"] = layer._get_attention_mask(torch.zeros(1), H, W).numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "known_answers.npz"), **kat)
    print("known_answers.npz:", sorted(kat))

    end_to_end(list(CASES))
    fpn_golden()


def end_to_end(names):
    # --- end-to-end cases --------------------------------------------------------------------
    for name in names:
        cfg, shape, kind, full = CASES[name]
        sd = O.make_state_dict(cfg, PARAM_SEED)
        model = R.build_reference_model(cfg, sd)            # strict=True: pins key names
        img = O.make_image(shape, IMAGE_SEED, kind)
        outs, blocks = R.reference_forward(model, img, return_blocks=True)
        rec = {"meta": np.array(json.dumps(dict(cfg=cfg, shape=shape, kind=kind, param_seed=PARAM_SEED,
                                               image_seed=IMAGE_SEED, full=full, stride=SAMPLE_STRIDE,
                                               n_out=len(outs), n_blocks=len(blocks),
                                               keys=sorted(sd.keys()) if name == "tiny_pano" else None)))}
        for i, o in enumerate(outs):
            rec[f"out{i}_shape"] = np.array(o.shape)
            rec[f"out{i}_norm"] = np.array(float(o.double().norm()))
            rec[f"out{i}"] = o.numpy() if full else sample(o)
        for i, b in enumerate(blocks):
            rec[f"block{i}_norm"] = np.array(float(b.double().norm()))
            rec[f"block{i}"] = sample(b)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: {len(outs)} outputs, {len(blocks)} blocks, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
