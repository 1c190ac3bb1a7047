"""CPU-side checks of the C-ABI boundary: the library builds, loads, and exports exactly what
include/tedm_b200.h declares (no compute calls: there is no GPU here)."""
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tedm_b200.h")).read()
    return sorted(set(re.findall(r"TEDM_API\s+[\w\s\*]+?\b(tedm_\w+)\s*\(", src)))


def test_header_declares_symbols():
    names = _declared()
    assert "tedm_conv_igemm_fwd" in names and "tedm_q_sample" in names and len(names) >= 25


def test_library_builds_and_exports_every_declared_symbol():
    from tedm_b200 import _build, native
    _build.build()
    lib = native.load()
    assert lib.tedm_version() == 1
    declared = _declared()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(native.SIGNATURES) == declared, "native.SIGNATURES and include/tedm_b200.h disagree"


def test_conv_tile_selection_predicates_are_host_logic():
    """Which 3x3 convs accept a fused input GroupNorm (tedm_conv_src_affine_supported mirrors the tile selection of
    tedm_conv_igemm_fwd: halo tiles, or the four-row weight-stationary kernel) and how many GroupNorm partial slots a conv
    writes: pure host functions, no device needed."""
    from tedm_b200 import native
    lib = native.load()
    ok = lib.tedm_conv_src_affine_supported
    assert ok(64, 64, 128, 128) == 1 and ok(16, 16, 512, 512) == 1 and ok(32, 32, 256, 256) == 1     # halo tiles
    assert ok(128, 128, 64, 64) == 1                       # four-row weight-stationary kernel, one channel block
    assert ok(128, 128, 128, 64) == 0                      # ... two channel blocks: no transform warps
    assert ok(64, 64, 64, 64) == 0                         # 64-pixel rows of two images
    assert ok(8, 8, 256, 256) == 0 and ok(16, 12, 64, 128) == 0 and ok(48, 48, 64, 128) == 0   # below a tile / not a power of two
    try:
        lib.tedm_conv_set_halo(0)
        assert ok(64, 64, 128, 128) == 0 and ok(128, 128, 64, 64) == 1
        lib.tedm_conv_set_ws(0)
        assert ok(128, 128, 64, 64) == 0
    finally:
        lib.tedm_conv_set_halo(1)
        lib.tedm_conv_set_ws(1)
    assert lib.tedm_conv_gn_parts(128, 128) == 128 and lib.tedm_conv_gn_parts(16, 16) == 2 and lib.tedm_conv_gn_parts(8, 8) == 1


def test_no_cpu_fallback():
    import torch
    from tedm_b200.models import Unet
    m = Unet(dim=64, dim_mults=[1, 2])
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 1, 32, 32), torch.zeros(1, dtype=torch.long))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tedm_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("the oracle", ""), f"{f} references the oracle"


def test_state_dict_layout_matches_reference():
    import json
    from argparse import Namespace
    from tedm_b200.models import DatasetDM, Unet, tedm_classifier
    inv = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    shapes = lambda m: {k: list(v.shape) for k, v in m.state_dict().items()}
    assert shapes(Unet()) == inv["unet_default"]
    assert shapes(Unet(64, dim_mults=[1, 2], channels=2, out_dim=3)) == inv["unet_mults12_outdim3"]
    d = DatasetDM(Namespace(normalize=True, saved_diffusion_model="/nonexistent",
                            t_steps_to_save=[1, 10, 25, 50, 200, 400, 600, 800]))
    assert shapes(d) == inv["ledme"]
    d.classifier = tedm_classifier(8)
    assert shapes(d) == inv["tedm"] and len(inv["tedm"]) == 301


def test_schedule_buffers_bit_exact(golden):
    import numpy as np
    from argparse import Namespace
    from tedm_b200.models import DiffusionModel
    g = golden["schedule"]
    for kind in ("cosine", "linear"):
        m = DiffusionModel(Namespace(normalize=True, beta_schedule=kind, p2_loss_weight_gamma=1.0, dim_mults=[1]))
        for k, v in m.state_dict().items():
            if not k.startswith("model."):
                assert np.array_equal(v.numpy(), g[f"{kind}.{k}"]), (kind, k)


def test_cli_options_match_reference_defaults():
    """tedm_b200/config.py keeps the reference's option names and defaults (config.py:13-83; fixture written from the live
    reference's parser: tests/golden/config_defaults.json), minus the SegDiff / contrastive-only options."""
    import json
    from tedm_b200.config import parser
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "config_defaults.json")))
    ours = {a.dest: a for a in parser._actions}
    out_of_scope = {"seg_out_dim", "img_out_dim", "img_inter_dim", "tau", "global_model_path", "glob_loc_model_path",
                    "unfreeze_weights_at_step", "augment_at_finetuning"}
    for name, spec in ref.items():
        if name in out_of_scope:
            continue
        assert name in ours, f"option --{name} of the reference is missing"
        d = ours[name].default
        d = list(d) if isinstance(d, (list, tuple)) else d
        assert d == spec["default"], (name, d, spec["default"])
    cfg = parser.parse_args(["--experiment", "TEDM", "--n_labelled_images", "6"])
    assert cfg.experiment == "TEDM" and cfg.batch_size == 16 and cfg.timesteps == 1000 and cfg.cuda_graph


def test_trainers_import_and_refuse_cpu():
    import torch
    from tedm_b200.dataloaders.device_loader import DeviceLoader, SyntheticXray
    from tedm_b200.trainers import train_baseline, train_CXR14, train_datasetDM  # noqa: F401
    ds = SyntheticXray(4, 32, labelled=True)
    img, masks = ds[1]
    assert img.dtype == torch.uint8 and img.shape == (1, 32, 32) and masks.shape == (2, 32, 32)
    assert torch.equal(ds[1][0], img)                       # deterministic per index
    with pytest.raises(RuntimeError, match="CUDA"):
        DeviceLoader(ds, 2, False, 0, device="cpu", labelled=True)


def test_snapshot_grid_matches_torchvision_make_grid():
    """sample_plot_image assembles its snapshot grids like torchvision.utils.make_grid(nrow=4) (trainers/utils.py:91-93)."""
    tv = pytest.importorskip("torchvision.utils")
    import torch
    from tedm_b200.trainers.utils import snapshot_grid
    g = torch.Generator().manual_seed(0)
    for n, c, h in ((8, 1, 16), (5, 1, 12), (8, 3, 8)):
        imgs = torch.rand(n, c, h, h, generator=g)
        assert torch.equal(snapshot_grid(imgs), tv.make_grid(imgs, nrow=4)), (n, c, h)


def test_vendored_reference_is_unmodified_and_is_what_the_cpu_leg_times():
    """oracle/_ref (git-ignored, made by oracle/make_ref.py) holds byte-identical copies of the reference's own files, and
    bench.py's CPU leg runs THEM (cpu_baseline.kind == "reference"); without it the leg falls back to the oracle port."""
    import hashlib
    import json
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isfile(os.path.join(ref_dir, "MANIFEST.json")):
        pytest.skip("oracle/_ref not built (python oracle/make_ref.py needs /root/reference)")
    manifest = json.load(open(os.path.join(ref_dir, "MANIFEST.json")))["sha256"]
    for rel, digest in manifest.items():
        assert hashlib.sha256(open(os.path.join(ref_dir, rel), "rb").read()).hexdigest() == digest, rel
        src = os.path.join("/root/reference", rel)
        if os.path.isfile(src):
            assert open(src, "rb").read() == open(os.path.join(ref_dir, rel), "rb").read(), rel
    import subprocess
    tracked = subprocess.run(["git", "ls-files", "oracle/_ref"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    assert tracked == "", "reference sources must never be committed"
    sys.path.insert(0, ROOT)
    import bench
    v, dt, kind = bench.cpu_tedm_images_per_s(1)
    assert kind == "reference" and v > 0
