import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests.test_models_gpu import build_net
from tests.util import rel_l2
from multimodal_segmentation_b200 import engine as E, ops
from oracle import ref_models as RM
H = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net, conf = build_net(H=H, filters=64, rounding=False, use_tc=True)
rs = np.random.RandomState(0)
x = rs.uniform(-1, 1, size=(2, H, H, 1)).astype(np.float32)
enc = net.Encoders_Anatomy[0]
W = {k: torch.from_numpy(v) for k, v in enc.named_weights().items()}
st = RM.BNState(W, True)
# oracle stage by stage
ref = {}
l = torch.from_numpy(x); skips = []
for i in range(4):
    d = RM.conv_block(W, "enc1_d%d" % i, l, st); ref['d%d' % i] = d; skips.append(d); l = RM.R.maxpool2(d)
l = RM.conv_block(W, "shared_bt", l, st); ref['bt'] = l
for i in reversed(range(4)):
    up = RM.upsample_block(W, "shared_u%d_up" % i, l, st); ref['up%d' % i] = up
    l = RM.conv_block(W, "shared_u%d" % i, torch.cat([up, skips[i]], -1), st); ref['u%d' % i] = l
# product stage by stage using the same layer objects (forward only, training stats)
# find the python objects through the closure of the forward
fwd = enc._forward
cells = {n: c.cell_contents for n, c in zip(fwd.__code__.co_freevars, fwd.__closure__)}
down, up_ = cells['down'], cells['up']
for mode in (False, True):
    E.USE_TC = mode
    ctx = E.Ctx(None, True)
    l = E.Var(torch.from_numpy(x).cuda()); sk = []
    out = {}
    for i, b in enumerate(down.blocks):
        d = b(ctx, l); out['d%d' % i] = d; sk.append(d); l = E.maxpool2(ctx, d)
    l = up_.bottleneck(ctx, l); out['bt'] = l
    n = len(up_.ups)
    for j, (u, b) in enumerate(zip(up_.ups, up_.blocks)):
        upv = u(ctx, l); out['up%d' % (n - 1 - j)] = upv
        l = b(ctx, [upv, sk[n - 1 - j]]); out['u%d' % (n - 1 - j)] = l
    print('TC' if mode else 'fp32', {k: '%.1e' % rel_l2(v.data.float().cpu().numpy(), ref[k].numpy()) for k, v in out.items()})
