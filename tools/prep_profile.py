import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import gsb200
from gsb200 import forward as gf, backward as gb, scene, _lib
n,w,h=300000,800,800
params, cam, target = scene.synthetic_scene(n, w, h, 0.004, 0.02)
dev=torch.device("cuda",0)
P={k: torch.from_numpy(v).to(dev) for k,v in params.items()}
bg=np.zeros(3,dtype=np.float32)
L=_lib.lib()
orig_f, orig_b = L.gsb_forward, L.gsb_backward
marks={}
def wrap(name, fn):
    def inner(*a):
        marks[name+"_enter"]=time.perf_counter()
        r=fn(*a)
        marks[name+"_exit"]=time.perf_counter()
        return r
    return inner
class LL:  # proxy
    def __getattr__(self, k):
        if k=="gsb_forward": return wrap("f", orig_f)
        if k=="gsb_backward": return wrap("b", orig_b)
        return getattr(L,k)
_lib._lib = LL()
res=[]
for it in range(12):
    torch.cuda.synchronize()
    kw=scene.render_kwargs(P, cam, background=bg)
    t0=time.perf_counter()
    img,_d,buf=gf.render_gaussians(**kw)
    t1=time.perf_counter()
    dpix=torch.empty((h,w,3),device=dev)
    bkw=scene.backward_kwargs(P, cam, buf, dpix, background=bg)
    t2=time.perf_counter()
    g=gb.backward(**bkw)
    t3=time.perf_counter()
    res.append(((marks["f_enter"]-t0)*1e6,(marks["f_exit"]-marks["f_enter"])*1e6,(t1-marks["f_exit"])*1e6,(marks["b_enter"]-t2)*1e6,(marks["b_exit"]-marks["b_enter"])*1e6,(t3-marks["b_exit"])*1e6))
print("fwd prep / C call / post | bwd prep / C call / post (us, median of last 8)")
print(np.median(np.array(res[4:]),axis=0).round(1))
