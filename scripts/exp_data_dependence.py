import importlib, sys, numpy as np, time
sys.path.insert(0,'.')
pkg=importlib.import_module("3dfluidsimulation_b200")
import bench
n=512
s=pkg.NativeSolver(n,n,n,iters_diffuse=20,iters_pressure=80,enable_obstacle=False,use_cuda_graph=False)
for fill,name in ((2,"zeros"),(1,"random")):
    for kind,kn in ((1,"jacobi"),(0,"smooth")):
        for reps in (1,5,20,50):
            ms,by=s.bench_sweep(kind+16*fill,0,reps)
            print(f"{name:7s} {kn:7s} reps={reps:3d} {ms*1e3:8.1f} us  {by/ms/1e6:8.1f} GB/s",flush=True)
px,py,pz,fall=bench.plume(n,n)
dt=0.1*128/n; vsrc=2.5/(dt*(n-2))
for step in range(1,9):
    s.add_source_cells(px,py,pz,density=np.float32(100)*fall,ay=np.float32(vsrc)*fall)
    s.step(dt,1e-4,1e-4)
    if step in (1,2,4,8):
        ms,by=s.bench_sweep(1,0,20)
        f=s.get_field("vx0"); den=np.count_nonzero((np.abs(f)<7.9e-31)&(f!=0)); 
        print(f"after step {step}: jacobi as-is {ms*1e3:8.1f} us {by/ms/1e6:8.1f} GB/s ; vx0 tiny-nonzero cells {den} ({100*den/f.size:.2f}%) zeros {100*np.count_nonzero(f==0)/f.size:.1f}%",flush=True)
s.timer_start(); 
for _ in range(3): s.step(dt,1e-4,1e-4)
print("step ms", s.timer_stop()/3)
