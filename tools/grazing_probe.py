import sys, os, numpy as np, torch
sys.path.insert(0,'/root/repo')
import rtb200
rtb200.hostlib.set_num_threads(os.cpu_count() or 1)
m = rtb200.Mesh().terrain(707, 100.0).finish(diffuse=(0.7,0.7,0.7)); A=m.arrays(); b=rtb200.FlatBVH.build(m)
w,h=1920,1080
ctx=rtb200.Context(0); stream=torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
ctx.upload_scene(A,b.nodes,b.tri_indices)
flush=torch.empty(256<<20,dtype=torch.uint8,device="cuda")
def timeit(fn, steps=10, warm=3):
    for _ in range(warm):
        flush.fill_(1); torch.cuda.synchronize()
        with torch.cuda.stream(stream): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(steps):
        flush.fill_(1); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(); fn(); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return round(float(np.median(ts)),4)
n=w*h
for light in ((-23.0,200.0,3.0),(-150.0,25.0,3.0)):
    params,_=rtb200.camera_params(w,h,A["aabb_min"],A["aabb_max"],light_pos=light); ctx.set_params(params)
    d_hits=torch.zeros((n,4),device="cuda"); d_rays=torch.zeros((n,8),device="cuda"); d_sh=torch.zeros((n,4),device="cuda"); vis=torch.zeros((h,w),dtype=torch.int32,device="cuda"); img=torch.zeros((h,w),dtype=torch.int32,device="cuda")
    with torch.cuda.stream(stream): ctx.primary_device(w,h,d_hits,d_rays)
    torch.cuda.synchronize()
    for ie in (0,1):
        ctx.set_option("inner_exit_batch", ie)
        for sched in (0,1):
            ctx.set_option("scheduler", sched)
            print("light",light[0],"inner_exit",ie,"scheduler",sched,"shadow",timeit(lambda: ctx.shadow_device(n,d_rays,d_hits,d_sh)))
        ctx.set_option("scheduler",-1)
        print("light",light[0],"inner_exit",ie,"fused",timeit(lambda: ctx.primary_shadow_device(w,h,None,None,vis)),"frame",timeit(lambda: ctx.render_frame_device(w,h,img)))
