"""A/B timing of the relaxation sweep variants on one GPU (512^3, random normals, 30 back-to-back launches)."""
import importlib, os, sys
sys.path.insert(0, '.')
pkg = importlib.import_module("3dfluidsimulation_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
s = pkg.NativeSolver(n, n, n, iters_diffuse=20, iters_pressure=80, enable_obstacle=False, use_cuda_graph=False)
tag = "PF=" + os.environ.get("FS_RELAX_PREFETCH", "default")
for kind, kn in ((1, "jacobi"), (0, "smooth")):
    for rep in range(2):
        ms, by = s.bench_sweep(kind + 16, 0, 30)
        print(f"{tag:12s} {n}^3 {kn:7s} {ms*1e3:8.1f} us  {by/ms/1e6:8.1f} GB/s", flush=True)
s.close()
