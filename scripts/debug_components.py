import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests.test_models_gpu import build_net
from tests.util import rel_l2
from oracle import ref_models as RM, ref_ops as R
from multimodal_segmentation_b200 import engine as E, ops

net, conf = build_net(H=64, filters=16, rounding=False, use_tc=False)
rs = np.random.RandomState(0)
B = 2

def run(model, oracle_fn, inputs, name):
    W = {k: torch.from_numpy(v).double().requires_grad_(True) for k, v in model.named_weights().items()}
    tin = [torch.from_numpy(a).double().requires_grad_(True) for a in inputs]
    yr = oracle_fn(W, *tin)
    g = rs.normal(size=tuple(yr.shape)).astype(np.float32)
    (yr * torch.from_numpy(g).double()).sum().backward()
    for p in model.params():
        p.grad.zero_()
    tape = E.Tape(); ctx = E.Ctx(tape, True)
    vin = [E.Var(torch.from_numpy(a).cuda(), True) for a in inputs]
    y = model(ctx, *vin)
    print(name, 'fwd', rel_l2(y.data.cpu().numpy(), yr.detach().numpy()))
    y.grad = torch.from_numpy(g).cuda()
    tape.backward(); torch.cuda.synchronize()
    for v, t in zip(vin, tin):
        if t.grad is not None and v.grad is not None:
            print('   dinput', rel_l2(v.grad.cpu().numpy(), t.grad.numpy()))
    errs = [(rel_l2(p.grad.cpu().numpy(), W[p.name].grad.numpy()), p.name) for p in model.params()
            if np.linalg.norm(W[p.name].grad.numpy()) > 1e-9]
    errs.sort(reverse=True)
    print('   worst', errs[:4])

s = rs.uniform(size=(B, 64, 64, 8)).astype(np.float32)
x = rs.uniform(-1, 1, size=(B, 64, 64, 1)).astype(np.float32)
z = rs.normal(size=(B, 8)).astype(np.float32)
run(net.Segmentor, lambda W, a: RM.segmentor(W, a, RM.BNState(W, True)), [s], 'Segmentor')
run(net.Encoders_Anatomy[0], lambda W, a: RM.anatomy_encoder(W, a, RM.BNState(W, True), 'enc1_', 'shared_', rounding=False), [x], 'EncAnatomy')
run(net.Decoder, lambda W, a, b: RM.decoder_film(W, a, b), [s, z], 'Decoder')
m = rs.uniform(size=(B, 64, 64, 4)).astype(np.float32)
run(net.D_Mask, lambda W, a: RM.discriminator(W, 'D_Mask', a), [m], 'D_Mask')
class Mu:
    def __init__(s, e): s.e = e
    def named_weights(s): return s.e.named_weights()
    def params(s): return s.e.params()
    def __call__(s, ctx, a, b): return s.e.forward_mu(ctx, a, b)
run(Mu(net.Enc_Modality), lambda W, a, b: RM.modality_encoder(W, a, b)[0], [s, x], 'EncM_mu')

# seg loss on softmax outputs at graph size
pred = R.softmax(torch.from_numpy(rs.normal(size=(B, 64, 64, 5)))).float().numpy()
lab = rs.randint(0, 5, size=(B, 64, 64)); tgt = np.eye(5, dtype=np.float32)[lab]
pt = torch.from_numpy(pred).double().requires_grad_(True)
(10.0 * R.combined_dice_bce(torch.from_numpy(tgt).double(), pt, 4)).backward()
loss = ops.zeros(1)
dp = ops.segloss(torch.from_numpy(pred).cuda(), torch.from_numpy(tgt).cuda(), 4, 1, 10.0, loss)
print('segloss grad', rel_l2(dp.cpu().numpy(), pt.grad.numpy()))
# realistic: near-uniform predictions (as at init)
pred = R.softmax(torch.from_numpy(rs.normal(size=(B, 64, 64, 5)) * 0.05)).float().numpy()
pt = torch.from_numpy(pred).double().requires_grad_(True)
(10.0 * R.combined_dice_bce(torch.from_numpy(tgt).double(), pt, 4)).backward()
dp = ops.segloss(torch.from_numpy(pred).cuda(), torch.from_numpy(tgt).cuda(), 4, 1, 10.0, loss)
print('segloss grad (uniform)', rel_l2(dp.cpu().numpy(), pt.grad.numpy()))
