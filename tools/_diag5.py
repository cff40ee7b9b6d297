import sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import numpy as np
from tests.helpers import mandala_scenario, model_kwargs, state_err, cov_err
from dvi_ekf_b200 import BatchFilter
g=np.load('tests/golden/reference_golden.npz')
sc=mandala_scenario(g,n_frames=140,ifv=10)
for (a,b) in ((0.1,0.1),(0.01,0.01),(0.4,0.1)):
  Qd=sc.Qd.copy(); Qd[6:9]*=b**2; Qd[9:12]*=a**2
  print('scales',a,b)
  kf=sc.new_oracle(); kf.Q=np.diag(Qd)
  with BatchFilter(1, variant=3, **model_kwargs(sc.cfg)) as bf:
    bf.set_noise(Qd[None], sc.Rd[None], sc.sig_om[None])
    bf.set_state(sc.x0[None], sc.P0[None], sc.u0[None], None)
    k=0
    for e in range(139):
        n=int(sc.n_prop[e])
        bf.run(sc.dt[k:k+n], sc.om_acc[k:k+n], sc.n_prop[e:e+1], sc.cam_meas[e:e+1], sc.notch_meas[e:e+1], want_stats=False)
        for _ in range(n):
            kf.propagate(sc.dt[k],sc.om_acc[k,:3],sc.om_acc[k,3:]); k+=1
        kf.update(sc.cam_meas[e,:3],sc.cam_meas[e,3:],sc.notch_meas[e])
        xg,Pg,_,_,st=bf.get_state(); xr,Pr,_,_=kf.get_vectors()
        if not np.isfinite(Pg).all(): print(e,'nonfinite'); break
        asg=np.abs(Pg[0]-Pg[0].T).max()/np.abs(Pg[0]).max(); asr=np.abs(Pr-Pr.T).max()/np.abs(Pr).max()
        if e in (0,5,10,15,20,30,50,80,110,138):
            print(e,'st',st[0],'serr %.1e cerr %.1e'%(state_err(xg[0],xr),cov_err(Pg[0],Pr,sc.Rd)),'asym gpu %.1e ora %.1e'%(asg,asr),'|p| %.4g %.4g'%(np.abs(xg[0,:3]).max(),np.abs(xr[:3]).max()))
