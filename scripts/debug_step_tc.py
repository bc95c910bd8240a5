import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests.test_models_gpu import build_net, make_batch, oracle_step, product_step
from tests.util import rel_l2
net, conf = build_net(H=64, filters=64, rounding=False, use_tc=True)
batch = make_batch(conf, 2)
W, total, L, inter, st = oracle_step(net, conf, batch, True, dtype=torch.float32)
tr = product_step(net, batch, True)
print('loss', tr.book.buf.cpu().numpy().sum(), total.item())
for p in net.generator_params():
    g = p.grad.cpu().numpy().astype(np.float64); r = W[p.name].grad.numpy()
    if np.linalg.norm(r) > 1e-4:
        print('%-32s err %.2e  |g| %.3e |ref| %.3e' % (p.name, rel_l2(g, r), np.linalg.norm(g), np.linalg.norm(r)))
