import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import psfr_oracle as orc
from muse_psfr_b200 import psfrec, _lib
psfrec.set_device(0)
ctx=psfrec.get_context()
psd=orc.simul_psd_wfm([0.7,0.3],(100,10000),1.0,25.)
lam=np.array([500.,700.,900.])
ref=orc.psf_muse(psd,lam)
def run(g,f,c=64.):
    ctx.set_option(_lib.OPT_EXP_GRADE,g); ctx.set_option(_lib.OPT_F32_ROWS,f); ctx.set_option(_lib.OPT_EXP_CUT,c)
    out=psfrec.psf_muse(psd,lam)
    return [float(np.abs(out[k]-ref[k]).max()/ref[k].max()) for k in range(3)]
for g,f in [(1e30,1e30),(25,1e30),(1e30,30),(25,30),(40,1e30),(60,1e30)]:
    print('grade',g,'f32',f,['%.2e'%e for e in run(g,f)])
print('cut off, all off', ['%.2e'%e for e in run(1e30,1e30,1000.)])
