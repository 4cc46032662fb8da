import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import scenes
from lustrine_b200 import lgpu, slabs
for side in (12, 16, 20):
  domain, pos = scenes.dam_break(side)
  vel0 = np.zeros_like(pos); vel0[:,0] = 12.0*np.sin(pos[:,2])
  for world in (2,3,4):
    for K in (1,2):
      kw = dict(dt=0.01, iterations=K, literal_lambda_index=0, exact_math=1)
      G = lgpu.Context(domain, capacity_sand=len(pos)); G.upload_sand(pos, vel0)
      V = slabs.VirtualSlabs(domain, pos, world, vel=vel0)
      G.step_fluid(**kw); V.step(1, **kw)
      rp, rv, _ = G.download(); sp, sv, _ = V.gather()
      err = np.abs(sp-rp).max(1)
      bad = np.nonzero(err>1e-4)[0]
      cs = slabs.cell_size(); cx0 = slabs.cell_x(pos, cs)
      infos=[c.G.slab_info() for c in V.ctx]
      print("side",side,"world",world,"K",K,"plan",V.slabs,"nbad",len(bad),"max %.3g"%err.max(),"badcols",np.unique(cx0[bad]).tolist(),
            "n",[i['owned']+i['ghosts'] for i in infos], "gh",[i['ghosts'] for i in infos])
      V.close(); G.close()
