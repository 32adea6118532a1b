import torch, sys
sys.path.insert(0, '.')
import oracle
from enhanced_unet_b200.models import EnhancedUNet
from enhanced_unet_b200 import lib
def nerr(a,b): return float((a.double()-b.double()).abs().max()/b.double().abs().max())
sd = oracle.make_state_dict(2)
for dtype in ("fp32","bf16"):
    for halo in (1,0):
        lib.set_option("conv_halo", halo)
        for res in (128, 512):
            m = EnhancedUNet(3, dtype=dtype); m.load_state_dict(sd); m = m.cuda().train()
            x1 = oracle.make_input(1, res, res, 3).cuda()
            with torch.no_grad():
                y1 = m(x1)
                y4 = m(x1.expand(4,3,res,res).contiguous())
                y1b = m(x1)
            print(dtype, 'halo',halo, res, 'rerun B1', nerr(y1b[0], y1[0]), 'B4[0] vs B1', nerr(y4[0], y1[0]), 'B4[3] vs B4[0]', nerr(y4[3], y4[0]))
