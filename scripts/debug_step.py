import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests.test_models_gpu import build_net, make_batch, oracle_step, product_step
from tests.util import rel_l2
net, conf = build_net(H=64, filters=16, rounding=False, use_tc=False)
batch = make_batch(conf, 2)
W, total, L, inter, st = oracle_step(net, conf, batch, True)
tr = product_step(net, batch, True)
rows = []
for p in net.generator_params():
    g = p.grad.cpu().numpy().astype(np.float64); r = W[p.name].grad.numpy()
    rows.append((rel_l2(g, r), np.linalg.norm(r), p.name))
for e, n, name in rows:
    print('%-32s err %.2e  |g| %.3e' % (name, e, n))
