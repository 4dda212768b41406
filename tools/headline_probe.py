import sys, os, subprocess, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
import rtb200
rtb200.hostlib.set_num_threads(os.cpu_count() or 1)
m = rtb200.Mesh().terrain(707, 100.0).finish(diffuse=(0.7,0.7,0.7)); A=m.arrays(); b=rtb200.FlatBVH.build(m)
w,h=1920,1080
params,_=rtb200.camera_params(w,h,A["aabb_min"],A["aabb_max"])
ctx=rtb200.Context(0); stream=torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
ctx.upload_scene(A,b.nodes,b.tri_indices); ctx.set_params(params)
d_hits=torch.zeros((w*h,4),device="cuda"); flush=torch.empty(256<<20,dtype=torch.uint8,device="cuda")
def timeit(fn, steps=20, warm=5, use_flush=True):
    with torch.cuda.stream(stream):
        for _ in range(warm): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(steps):
        if use_flush: flush.fill_(1); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(); fn(); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return round(float(np.mean(ts)),4), round(float(np.median(ts)),4), [round(t,3) for t in ts[:6]]
for br in (4,16,4,16):
    print("band_rows",br, timeit(lambda: ctx.primary_device(w,h,d_hits,None,part=0,n_parts=1,band_rows=br)))
p=subprocess.Popen(["nvidia-smi","--query-gpu=clocks.sm,power.draw","--format=csv,noheader,nounits","-lms","50"],stdout=subprocess.DEVNULL)
time.sleep(1.0)
for br in (4,16):
    print("with nvidia-smi sampler: band_rows",br, timeit(lambda: ctx.primary_device(w,h,d_hits,None,part=0,n_parts=1,band_rows=br)))
p.terminate()
ctx.set_option("gate_cull",0)
print("gate_cull=0 band_rows 16", timeit(lambda: ctx.primary_device(w,h,d_hits,None,part=0,n_parts=1,band_rows=16)))
