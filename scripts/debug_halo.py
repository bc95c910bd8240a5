import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multimodal_segmentation_b200 import ops
from oracle import ref_ops as R
from tests.util import cpu, gpu, rel_l2, t
def bf(a): return torch.as_tensor(a).to(torch.bfloat16).float().numpy()
for case in [(2,16,16,64,0,64),(1,24,40,64,0,64),(2,16,16,128,0,128),(1,56,56,64,0,64)]:
    N,H,W,C0,C1,Cout=case
    r=np.random.RandomState(1)
    x=bf(r.normal(size=(N,H,W,C0+C1)).astype(np.float32))
    w=bf((r.normal(size=(3,3,C0+C1,Cout))/np.sqrt(9*(C0+C1))).astype(np.float32))
    yr=R.conv2d(t(x,torch.float64),t(w,torch.float64),None,1,"same").numpy()
    wp=ops.pack_conv3x3(gpu(w))
    y=cpu(ops.conv3x3_tc_fwd(gpu(x,torch.bfloat16),None,wp,None,Cout))
    print(case,"err",rel_l2(y,yr), "center err", rel_l2(y[:,2:-2,2:-2],yr[:,2:-2,2:-2]), flush=True)
