import sys, os, time, threading
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("voxelengine-cpp_b200","oracle","tests"): sys.path.insert(0, os.path.join(ROOT,p))
sys.path.insert(0, ROOT)
import numpy as np, ctypes as C
import bench
from vxgpu import abi, content as ct, shard, gpu
K = int(sys.argv[1]); STEPS = int(sys.argv[2]) if len(sys.argv)>2 else 100
table = ct.ContentTable()
full = (-1,-1,34,34)
chunks = bench.generate_world("terrain", 2019, *full)
subs = shard.plan(*full, K)
ctxs=[]
for sh in subs:
    gw = gpu.GpuWorld(table, *sh.window, device=0)
    gw.put_chunks((k[0],k[1],chunks[k],None) for k in sh.resident)
    ctxs.append((gw, gpu.keys_array(sh.resident), gpu.keys_array(sh.lit), gpu.keys_array(sh.owned), len(sh.owned)))
def step(c):
    gw, res, lit, own, n = c
    gw.light_clear(); gw.prebuild_batch(res); gw.light_build(lit, True); gw.mesh_build(own, 800000)
def run(c, n):
    for _ in range(n): step(c)
for c in ctxs: run(c, 3)
bar = threading.Barrier(K+1)
def worker(c):
    bar.wait()
    run(c, STEPS)
    c[0].timer_begin(); c[0].timer_end()   # drains the stream
ths=[threading.Thread(target=worker, args=(c,)) for c in ctxs]
for t in ths: t.start()
bar.wait(); t0=time.perf_counter()
for t in ths: t.join()
dt=time.perf_counter()-t0
print("contexts", K, "ms per 1024-chunk step", round(dt/STEPS*1e3,4), "chunks/s", round(1024*STEPS/dt))
