"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the REAL reference (imported from /root/reference via
oracle/ref_loader.py) on seeded inputs and weights.  Run in the authoring container:  python -m oracle.make_golden

Each fixture stores the configuration and seeds (inputs and weights are re-drawn from numpy's PCG64 by the tests, see
``golden_case``), the three outputs on a stride-3 pixel grid, the loss terms, the L2 norm of every parameter gradient, a few
full gradients and the updated BatchNorm running statistics digest.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import towerunet_port as port  # noqa: E402

CASES = {
    "small_masked": dict(B=2, C=3, T=7, H=24, W=24, hidden=8, dilations=[1, 2], seed=11, y_low=-1),
    "odd_dil3": dict(B=1, C=2, T=6, H=25, W=25, hidden=8, dilations=[1, 2, 3], seed=23, y_low=0),
    # constructor-argument variants (the reference's tests/test_cultionet.py:67-78 builds spatial_channel + pool_by_max)
    "sca_maxpool": dict(B=2, C=3, T=6, H=24, W=24, hidden=8, dilations=[1, 2], seed=31, y_low=-1, attention=1, pool_by_max=1),
    "res_bnfirst": dict(B=2, C=2, T=6, H=20, W=20, hidden=8, dilations=[1, 2], seed=37, y_low=0, attention=2, res=1, batchnorm_first=1),
    "bnfirst_maxpool_odd": dict(B=1, C=2, T=6, H=25, W=25, hidden=8, dilations=[1, 2], seed=41, y_low=-1, batchnorm_first=1,
                                pool_by_max=1),
    # GeoEmbeddings broadcast channel block in the towers (use_latlon=True, nn/modules/unet_parts.py:739-750, geo_encoding.py:5-26)
    "latlon": dict(B=2, C=2, T=6, H=20, W=20, hidden=8, dilations=[1, 2], seed=43, y_low=0, use_latlon=1),
}
FULL_GRADS = [
    "pre_unet.conv3.seq.0.weight",
    "final_combine.edge_gamma2",
    "final_combine.final_edge.1.gamma",
    "decoder.up_cu.res_conv.attention_conv.2.qkv.bias",
    "final_a.fuse_conv.seq.0.weight",
    "encoder.down_b.pool_conv.seq.1.weight",
    # variants
    "decoder.up_au.res_conv.attention_conv.gamma",
    "decoder.up_bu.res_conv.attention_conv.channel_attention.fc2.0.weight",
    "decoder.up_cu.res_conv.attention_conv.spatial_attention.conv.weight",
    "encoder.down_b.res_conv.skip.weight",
    "encoder.down_c.pool_conv.bias",
    "tower_fusion.tower_a.res_conv.seq.block.0.seq.0.weight",
    "tower_fusion.tower_b.res_conv.res_modules.1.block.0.seq.0.bias",
    "tower_fusion.tower_a.geo_embeddings.coord_embedding.weight",
    "tower_fusion.tower_c.geo_embeddings.coord_embedding.bias",
]
ATTENTION = {0: "natten", 1: "spatial_channel", 2: None}


def variant_kwargs(cfg: dict) -> dict:
    """TowerUNet constructor arguments of a case beyond the defaults (shared by the generator and the tests)."""
    return dict(attention_weights=ATTENTION[cfg.get("attention", 0)], pool_by_max=bool(cfg.get("pool_by_max", 0)),
                batchnorm_first=bool(cfg.get("batchnorm_first", 0)), res_block_type="res" if cfg.get("res", 0) else "resa",
                use_latlon=bool(cfg.get("use_latlon", 0)))


def is_variant(cfg: dict) -> bool:
    return any(cfg.get(k, 0) for k in ("attention", "pool_by_max", "batchnorm_first", "res", "use_latlon"))


def golden_latlon(cfg: dict):
    """(lon, lat) in decimal degrees per sample for the use_latlon cases, else None."""
    if not cfg.get("use_latlon", 0):
        return None
    rng = np.random.default_rng(cfg["seed"] + 2000)
    lon = rng.uniform(-180.0, 180.0, size=(cfg["B"], 1))
    lat = rng.uniform(-90.0, 90.0, size=(cfg["B"], 1))
    return torch.from_numpy(np.concatenate([lon, lat], axis=1).astype(np.float32))


def golden_case(cfg: dict, spec=None):
    """Seeded weights and inputs of a case (shared by the generator and the tests).  Default-architecture cases take the parameter
    inventory from the port; variant cases pass the inventory stored in their fixture."""
    if spec is None:
        spec = port.param_spec(cfg["C"], cfg["T"], cfg["hidden"], cfg["dilations"])
    sd = port.synth_state_dict(spec, seed=cfg["seed"])
    rng = np.random.default_rng(cfg["seed"] + 1000)
    x = torch.from_numpy(rng.random((cfg["B"], cfg["C"], cfg["T"], cfg["H"], cfg["W"]), dtype=np.float32))
    y = torch.from_numpy(rng.integers(cfg["y_low"], 3, size=(cfg["B"], cfg["H"], cfg["W"]))).long()
    bdist = torch.from_numpy(rng.random((cfg["B"], cfg["H"], cfg["W"]), dtype=np.float32))
    return spec, sd, x, y, bdist


def run_reference(cfg: dict):
    from oracle.ref_loader import load_reference

    ref = load_reference()
    model = ref.TowerUNet(in_channels=cfg["C"], in_time=cfg["T"], hidden_channels=cfg["hidden"], dilations=cfg["dilations"],
                          **variant_kwargs(cfg))
    spec = None
    if is_variant(cfg):  # the parameter inventory of a variant comes from the reference model itself
        spec = [(k.replace("._orig_mod.", "."), tuple(v.shape)) for k, v in model.state_dict().items()]
    spec, sd, x, y, bdist = golden_case(cfg, spec)
    compiled = {k.replace("._orig_mod.", "."): k for k in model.state_dict()}  # torch.compile wrappers prefix their keys
    model.load_state_dict({compiled[k]: v for k, v in sd.items()}, strict=True)
    model.train()
    out = model(x, latlon_coords=golden_latlon(cfg))
    cls = ref.TanimotoComplementLoss()
    reg = ref.TanimotoComplementLoss(transform_logits=False, one_hot_targets=False)
    mask = (y != -1).long().unsqueeze(1) if int(y.min()) == -1 else None
    d = reg(out["distance"], bdist, mask=mask)
    e = cls(out["edge"], (y == 2).long(), mask=mask)
    c = cls(out["crop"], ((y > 0) & (y < 2)).long(), mask=mask)
    loss = (d + e + c) / 3.0
    loss.backward()
    grads = {n.replace("._orig_mod.", "."): p.grad.detach().clone() for n, p in model.named_parameters()}
    buffers = {n.replace("._orig_mod.", "."): b.detach().clone() for n, b in model.named_buffers()}
    return out, (loss, d, e, c), grads, buffers, spec


def main() -> None:
    out_dir = ROOT / "tests" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    only = set(sys.argv[1:])
    for name, cfg in CASES.items():
        if only and name not in only:
            continue
        out, losses, grads, buffers, spec = run_reference(cfg)
        names = sorted(grads)
        arrays = {
            "grad_names": np.array(names),
            "grad_norms": np.array([float(grads[n].double().norm()) for n in names], dtype=np.float64),
            "losses": np.array([float(v) for v in losses], dtype=np.float64),
        }
        for k in ("distance", "edge", "crop"):
            arrays["out_" + k] = out[k].detach().numpy()[:, :, ::3, ::3].astype(np.float32)
        for n in FULL_GRADS:
            if n in grads:
                arrays["grad::" + n] = grads[n].numpy().astype(np.float32)
        if is_variant(cfg):
            arrays["spec_names"] = np.array([k for k, _ in spec])
            arrays["spec_shapes"] = np.array([",".join(map(str, shp)) for _, shp in spec])
        rm = sorted(n for n in buffers if n.endswith("running_mean"))
        arrays["bn_names"] = np.array(rm)
        arrays["bn_mean_norms"] = np.array([float(buffers[n].double().norm()) for n in rm])
        arrays["bn_var_norms"] = np.array([float(buffers[n.replace("running_mean", "running_var")].double().norm()) for n in rm])
        arrays["cfg_keys"] = np.array(sorted(k for k in cfg if k != "dilations"))
        arrays["cfg_vals"] = np.array([cfg[k] for k in sorted(k for k in cfg if k != "dilations")], dtype=np.int64)
        arrays["dilations"] = np.array(cfg["dilations"], dtype=np.int64)
        np.savez_compressed(out_dir / f"towerunet_{name}.npz", **arrays)
        print(name, "loss", arrays["losses"], "size", (out_dir / f"towerunet_{name}.npz").stat().st_size)


if __name__ == "__main__":
    main()
