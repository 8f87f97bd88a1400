"""Debug: eval-mode bf16 at hidden 64 vs the fp32 port (localise the layer where they part)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
import torch
from cultionet_b200 import _lib, functional as F
from oracle import towerunet_port as port
from tests.util import mine_from_state_dict, rel_err
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
for hidden, B in ((64, 2), (64, 8)):
    cfg = dict(B=B, C=5, T=12, H=140, W=140, hidden=hidden, dilations=[1, 2])
    spec = port.param_spec(cfg["C"], cfg["T"], cfg["hidden"], cfg["dilations"])
    sd = port.synth_state_dict(spec, seed=3)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(cfg["B"], cfg["C"], cfg["T"], cfg["H"], cfg["W"], generator=g).to(dev)
    with torch.no_grad():
        calib = mine_from_state_dict(cfg, sd, dev, torch.float32).train()
        for m in calib.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.momentum = 1.0
        calib(x)
        sd2 = {k: v.detach().clone() for k, v in calib.state_dict().items()}
        taps = {}
        want = port.towerunet_forward({k: v.to(dev) for k, v in sd2.items()}, x, cfg["dilations"], training=False, taps=taps)
        for dtype in (torch.float32, torch.bfloat16):
            for fuse in (True, False):
                F.FUSE_EVAL_EPILOGUE = fuse
                F._EVAL_FUSE_REJECTED.clear()
                model = mine_from_state_dict(cfg, sd2, dev, dtype).eval()
                got = {}
                hooks = []
                def mk(name):
                    def hook(mod, inp, out):
                        got[name] = out
                    return hook
                for name, mod in (("e0", model.pre_unet), ("enc", model.encoder), ("dec", model.decoder), ("tow", model.tower_fusion),
                                  ("h_a", model.final_a), ("h_b", model.final_b), ("h_c", model.final_c)):
                    hooks.append(mod.register_forward_hook(mk(name)))
                out = model(x)
                errs = {k: round(rel_err(out[k], want[k]), 5) for k in ("distance", "edge", "crop")}
                lv = {}
                def cmp(a, key):
                    lv[key] = round(rel_err(a.float().permute(0, 3, 1, 2), taps[key]), 5)
                cmp(got["e0"], "e0")
                for k in ("x_a", "x_b", "x_c", "x_d"): cmp(got["enc"][k], k)
                for k in ("x_du", "x_cu", "x_bu", "x_au"): cmp(got["dec"][k], k)
                for k, kk in (("x_tower_c", "t_c"), ("x_tower_b", "t_b"), ("x_tower_a", "t_a")): cmp(got["tow"][k], kk)
                for k in ("h_a", "h_b", "h_c"): cmp(got[k], k)
                print(f"hidden {hidden} B {B} {dtype} fuse={fuse}: out {errs}\n    levels {lv}", flush=True)
                for h in hooks: h.remove()
