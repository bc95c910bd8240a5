import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests.test_models_gpu import build_net, make_batch, oracle_step, product_step, _run_component
from tests.util import rel_l2
from oracle import ref_models as RM
from multimodal_segmentation_b200 import engine as E
net, conf = build_net(H=64, filters=64, rounding=False, use_tc=True)
rs = np.random.RandomState(0)
RM.BF16_EMULATION = True
s = rs.uniform(size=(2, 64, 64, 8)).astype(np.float32)
x = rs.uniform(-1, 1, size=(2, 64, 64, 1)).astype(np.float32)
for name, model, fn, inp in (
    ('Segmentor', net.Segmentor, lambda W, a: RM.segmentor(W, a, RM.BNState(W, True)), [s]),
    ('D_Image1', net.D_Image1, lambda W, a: RM.discriminator(W, "D_Image1", a), [x]),
    ('EncAnatomy', net.Encoders_Anatomy[0], lambda W, a: RM.anatomy_encoder(W, a, RM.BNState(W, True), "enc1_", "shared_", rounding=False), [x])):
    W = {k: torch.from_numpy(v).double().requires_grad_(True) for k, v in model.named_weights().items()}
    tin = [torch.from_numpy(a).double().requires_grad_(True) for a in inp]
    yr = fn(W, *tin)
    g = rs.normal(size=tuple(yr.shape)).astype(np.float32)
    (yr * torch.from_numpy(g).double()).sum().backward()
    for p in model.params(): p.grad.zero_()
    tape = E.Tape(); ctx = E.Ctx(tape, True)
    vin = [E.Var(torch.from_numpy(a).cuda(), True) for a in inp]
    y = model(ctx, *vin)
    print(name, 'fwd', rel_l2(y.data.cpu().numpy(), yr.detach().numpy()))
    y.grad = torch.from_numpy(g).cuda(); tape.backward(); torch.cuda.synchronize()
    print('   dinput', rel_l2(vin[0].grad.cpu().numpy(), tin[0].grad.numpy()))
    for p in model.params():
        r = W[p.name].grad
        if r is not None and np.linalg.norm(r.numpy()) > 1e-3:
            print('   %-28s %.2e' % (p.name, rel_l2(p.grad.cpu().numpy(), r.numpy())))
