import sys, time; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from tacex_b200 import synth, calib, engine
from oracle import canon
H,W=240,320
T = calib.TaximTables.load(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))+'/tests/golden/gsmini_tables_320x240.npz')
taps = T.params.blur_taps((H,W))
cn = canon.CanonTaxim(H,W,T.poly_grad.numpy(), T.background.numpy(), None, taps)
eng = engine.TactileEngine(T, max_envs=64, marker_rows=9, marker_cols=11)
hm = torch.cat([synth.height_map_mm(synth.config0()['depth_m']), synth.golden_config1()])
hmd = hm.cuda()
press = eng.indentation_depth(hmd); torch.cuda.synchronize()
pc = cn.indentation_depth(hm.numpy())
print('press eq', np.array_equal(press.cpu().numpy(), pc), pc)
N=hm.shape[0]
dg = torch.empty((N,H,W),device='cuda'); mk=torch.empty((N,H,W),device='cuda',dtype=torch.uint8); dep=torch.empty(N,device='cuda')
rgb = eng.render(hmd, None, depth_out=dep, deformed_out=dg, mask_out=mk); torch.cuda.synchronize()
o = cn.render(hm.numpy(), pc)
print('depth fused eq', np.array_equal(dep.cpu().numpy(), pc))
print('mask diff', (mk.cpu().numpy()!=o['mask']).sum())
d = np.abs(dg.cpu().numpy()-o['deformed']); print('deformed maxdiff', d.max(), 'bitwise', np.array_equal(dg.cpu().numpy(), o['deformed']))
dr = np.abs(rgb.cpu().numpy()-o['rgb']); print('rgb maxdiff', dr.max(), 'n>1e-6', (dr>1e-6).sum())
rgb2 = eng.render(hmd, press); torch.cuda.synchronize(); print('explicit press same', torch.equal(rgb, rgb2))
