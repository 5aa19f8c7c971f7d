import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import gsb200
from gsb200 import forward as gf, scene, _lib
n,w,h=300000,800,800
params, cam, target = scene.synthetic_scene(n, w, h, 0.004, 0.02)
dev=torch.device("cuda",0)
P={k: torch.from_numpy(v).to(dev) for k,v in params.items()}
bg=np.zeros(3,dtype=np.float32)
kw=scene.render_kwargs(P, cam, background=bg)
ctx=_lib.context()
def t(name, fn, reps=300):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(reps): fn()
    dt=(time.perf_counter()-t0)/reps*1e6
    torch.cuda.synchronize()
    print(f"{name:40s} {dt:8.2f} us")
t("context()", lambda: _lib.context())
t("torch.device", lambda: torch.device("cuda", 0))
t("to_device means", lambda: _lib.to_device(kw["means3D"], device=dev, shape=(-1,3)))
t("to_device sh + reshape", lambda: _lib.to_device(kw["sh"], device=dev).reshape(-1,3))
t("make_frame", lambda: _lib.make_frame(kw["viewmatrix"], kw["projmatrix"], kw["campos"], kw["tan_fovx"], kw["tan_fovy"], w, h, kw["background"], 3, True, 1.0))
f32,i32=torch.float32, torch.int32
t("carve floats", lambda: _lib.carve(dev, f32, [("points_xy_image", (n, 2)), ("depths", (n,)), ("colors", (n, 3)), ("cov3Ds", (n, 6)), ("conic_opacity", (n, 4)), ("clamped_state", (n, 3)), ("final_Ts", (h, w)), ("image", (h, w, 3)), ("depth", (h, w))]))
t("carve ints", lambda: _lib.carve(dev, i32, [("radii", (n,)), ("point_offsets", (n,)), ("ranges", (2500, 2)), ("n_contrib", (h, w)), ("point_list", (1900000,)), ("block_masks", (1900000,))]))
t("torch.empty", lambda: torch.empty((n,3), dtype=f32, device=dev))
t("stream_ptr", lambda: _lib.stream_ptr(0))
x=P["positions"]
t("ptr", lambda: _lib.ptr(x))
t("render_kwargs", lambda: scene.render_kwargs(P, cam, background=bg))
cs=torch.cuda.Stream(device=dev)
pinned=torch.from_numpy(target).pin_memory()
def h2d():
    main=torch.cuda.current_stream()
    with torch.cuda.stream(cs):
        tg=pinned.to(dev, non_blocking=True)
        ev=cs.record_event()
    main.wait_event(ev); tg.record_stream(main)
t("h2d on copy stream (host time)", h2d, 100)
