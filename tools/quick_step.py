"""Step time of configs[1] with whatever libvxgpu.so is in place: three timings of 60 steps, nothing else
(no end-to-end leg, no CPU baseline) — 6 s on the GPU box, for sweeps over build-time knobs
(make EXTRA=-DVX_EMIT_UPT=8 / -DVX_EMIT_MINB / -DVX_SEED_MINB / -DVX_CLASSIFY_MINB / -DVX_SOLVE_PER_SM, one library
per value, copied over voxelengine-cpp_b200/vxgpu/libvxgpu.so between runs).
Usage: python tools/quick_step.py [label]"""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "voxelengine-cpp_b200")]
import bench
from vxgpu import gpu, shard, content as ct
wl = bench.WORKLOADS["terrain"]
side = wl["side"]
full = (-1, -1, side + 2, side + 2)
sh = shard.plan(*full, 1)[0]
chunks = bench.generate_world(wl["gen"], wl["seed"], *full, sub=sh.window)
gw = gpu.GpuWorld(ct.ContentTable(), *sh.window, device=0)
for at in range(0, len(sh.resident), 512):
    gw.put_chunks((k[0], k[1], chunks[k], None) for k in sh.resident[at:at + 512])
all_keys, lit, owned = (gpu.keys_array(k) for k in (sh.resident, sh.lit, sh.owned))
n_owned = len(sh.owned)
def step():
    gw.light_clear(); gw.prebuild_batch(all_keys); gw.light_build(lit, True)
    bench.lib_check(gw, gw.lib.vx_mesh_build(gw.h, owned, n_owned, bench.CAPACITY))
for _ in range(5): step()
best = []
for rep in range(3):
    gw.timer_begin()
    for _ in range(60): step()
    best.append(gw.timer_end() / 60)

print(sys.argv[1] if len(sys.argv) > 1 else "", "ms/step:", " ".join("%.4f" % b for b in best))
