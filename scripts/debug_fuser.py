import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests.test_models_gpu import build_net
from tests.util import rel_l2
from oracle import ref_models as RM, ref_ops as R
from multimodal_segmentation_b200 import engine as E, ops

net, conf = build_net(H=64, filters=16, rounding=False, use_tc=False)
rs = np.random.RandomState(0)
for soft in (True, False):
    a1 = rs.uniform(size=(2, 64, 64, 8)).astype(np.float32)
    a2 = rs.uniform(size=(2, 64, 64, 8)).astype(np.float32)
    if not soft:
        a1, a2 = (a1 > 0.6).astype(np.float32), (a2 > 0.6).astype(np.float32)
    g = rs.normal(size=a1.shape).astype(np.float32)
    F = net.Anatomy_Fuser
    W = {k: torch.from_numpy(v).double().requires_grad_(True) for k, v in F.named_weights().items()}
    t1 = torch.from_numpy(a1).double().requires_grad_(True)
    t2 = torch.from_numpy(a2).double().requires_grad_(True)
    d, f, th = RM.anatomy_fuser(W, t1, t2)
    th.retain_grad()
    (d * torch.from_numpy(g).double()).sum().backward()
    for p in F.params():
        p.grad.zero_()
    tape = E.Tape(); ctx = E.Ctx(tape, True)
    v1, v2 = E.Var(torch.from_numpy(a1).cuda(), True), E.Var(torch.from_numpy(a2).cuda(), True)
    theta = F.locnet(ctx, v1, v2)
    out = E.tps_warp(ctx, v1, theta)
    out.grad = torch.from_numpy(g).cuda()
    # capture dtheta
    dvol, dth = ops.tps_warp_bwd(v1.data, theta.data, out.grad)
    print('soft' if soft else 'binary', 'theta', rel_l2(theta.data.cpu().numpy(), th.detach().numpy()),
          'out', rel_l2(out.data.cpu().numpy(), d.detach().numpy()),
          'dtheta', rel_l2(dth.cpu().numpy(), th.grad.numpy()))
    tape.backward(); torch.cuda.synchronize()
    print('  da1', rel_l2(v1.grad.cpu().numpy(), t1.grad.numpy()), 'da2', rel_l2(v2.grad.cpu().numpy(), t2.grad.numpy()))
    for p in F.params():
        print('  ', p.name, rel_l2(p.grad.cpu().numpy(), W[p.name].grad.numpy()))
    # direct kernel check with the oracle's own theta
    thn = th.detach().numpy().astype(np.float32)
    dvol2, dth2 = ops.tps_warp_bwd(v1.data, torch.from_numpy(thn).cuda(), out.grad)
    print('  dtheta(oracle theta)', rel_l2(dth2.cpu().numpy(), th.grad.numpy()), np.abs(thn).max())
    # per-sample, per-component error
    e = dth.cpu().numpy() - th.grad.numpy()
    print('  per-comp rel', np.linalg.norm(e[..., 0]) / np.linalg.norm(th.grad.numpy()[..., 0]), np.linalg.norm(e[..., 1]) / np.linalg.norm(th.grad.numpy()[..., 1]))
