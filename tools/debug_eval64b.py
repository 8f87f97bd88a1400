"""Debug: tower_a at B=8, 140x140, hidden 64 (eval): which operand / operator parts from torch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
import torch
import torch.nn.functional as TF
from oracle import towerunet_port as port
from tests.util import mine_from_state_dict, rel_err
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = dict(B=B, C=5, T=12, H=140, W=140, hidden=64, dilations=[1, 2])
spec = port.param_spec(cfg["C"], cfg["T"], cfg["hidden"], cfg["dilations"])
sd = port.synth_state_dict(spec, seed=3)
g = torch.Generator().manual_seed(3)
x = torch.rand(cfg["B"], cfg["C"], cfg["T"], cfg["H"], cfg["W"], generator=g).to(dev)
with torch.no_grad():
    model = mine_from_state_dict(cfg, sd, dev, torch.float32).eval()
    sdd = {k: v.to(dev) for k, v in model.state_dict().items()}
    taps = {}
    port.towerunet_forward(sdd, x, cfg["dilations"], training=False, taps=taps)
    got = {}
    ta = model.tower_fusion.tower_a
    for name, mod in (("bd", ta.backbone_down_conv), ("dd", ta.decode_down_conv), ("tc", ta.tower_conv), ("res", ta.res_conv),
                      ("skip_in", ta.res_conv)):
        mod.register_forward_hook((lambda n: (lambda m, i, o: got.__setitem__(n, (i, o))))(name))
    model(x)
    c = port._Ctx(sdd, False, None)
    p = "tower_fusion.tower_a"
    size = (140, 140)
    w_bd = port._convT_fwd(c, taps["x_b"], p + ".backbone_down_conv", size)
    w_dd = port._convT_fwd(c, taps["x_bu"], p + ".decode_down_conv", size)
    w_tc = port._convT_fwd(c, taps["t_b"], p + ".tower_conv", size)
    nchw = lambda t: t.float().permute(0, 3, 1, 2)
    print("B", B)
    print("convT backbone_down", rel_err(nchw(got["bd"][1]), w_bd))
    print("convT decode_down", rel_err(nchw(got["dd"][1]), w_dd))
    print("convT tower", rel_err(nchw(got["tc"][1]), w_tc))
    cat = torch.cat([taps["x_a"], w_bd, taps["x_au"], w_dd, w_tc], dim=1)
    print("cat", tuple(cat.shape), cat.numel())
    # the pieces of ResidualAConv in torch, each against the same op on a per-sample loop (cuDNN algorithm check)
    wsk, bsk = sdd[p + ".res_conv.skip.weight"], sdd[p + ".res_conv.skip.bias"]
    full = TF.conv2d(cat, wsk, bsk)
    per = torch.cat([TF.conv2d(cat[b:b + 1], wsk, bsk) for b in range(B)], 0)
    print("torch 1x1 skip: batched vs per-sample", rel_err(full, per))
    w0 = sdd[p + ".res_conv.res_modules.0.block.0.seq.0.weight"]
    full3 = TF.conv2d(cat, w0, None, padding=1)
    per3 = torch.cat([TF.conv2d(cat[b:b + 1], w0, None, padding=1) for b in range(B)], 0)
    print("torch 3x3 960->256: batched vs per-sample", rel_err(full3, per3))
    full3d = TF.conv2d(cat.double(), w0.double(), None, padding=1).float()
    print("torch 3x3 fp32 batched vs fp64", rel_err(full3, full3d), " per-sample vs fp64", rel_err(per3, full3d))
    from cultionet_b200 import functional as F
    srcs = [t.permute(0, 2, 3, 1).contiguous() for t in (taps["x_a"], w_bd, taps["x_au"], w_dd, w_tc)]
    mine3 = F.conv2d(srcs, w0, None, ksize=3, stride=1, pad=1)
    print("mine 3x3 fp32 vs fp64", rel_err(nchw(mine3), full3d))
    mine3b = F.conv2d([s.bfloat16() for s in srcs], w0, None, ksize=3, stride=1, pad=1)
    print("mine 3x3 bf16 vs fp64", rel_err(nchw(mine3b), full3d))
    want_ta = port._resa_fwd(c, cat, p + ".res_conv", 3, 2, [1, 2])
    print("tower_a: mine vs port", rel_err(nchw(got["res"][1]), want_ta), " port vs taps", rel_err(want_ta, taps["t_a"]))
