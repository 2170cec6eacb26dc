#!/usr/bin/env python
"""bench.py — chunks lit+meshed per second on B200 (BASELINE.json metric).

A step is one full rebuild of the workload from voxels resident in HBM:
Lighting::clear + prebuildSkyLight + buildSkyLight/onChunkLoaded for every
inner chunk + BlocksRenderer::build for every inner chunk, all through the C
ABI of include/vxgpu.h.  Workload (configs[1]): 32x32 = 1,024 terrain chunks
plus a one-chunk ring of border chunks (1,156 resident), seed 2019; at N GPUs
every rank owns its own 32-column-wide range of a (32N)x32 world and holds a
duplicated one-chunk halo (weak scaling, no data-path collective: light
travels < 15 voxels, so a lit halo chunk reproduces the neighbour shard's
influence exactly).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints one JSON line (rank 0).  --impl reference times the reference's own CPU
implementation (oracle/_ref/libvxref.so when it was built, else the oracle
port) on the host cores.
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "voxelengine-cpp_b200"), os.path.join(ROOT, "oracle"),
           os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

from vxgpu import abi, content as ct, shard, worlds  # noqa: E402

SIDE = 32            # inner chunks per side and per rank
SEED = 2019
CAPACITY = 800000    # graphics.chunk-max-vertices-dense (settings.hpp:73-75)
METRIC = "chunks lit+meshed/sec (16x256x16)"


def _gen(args):
    cx, cz = args
    return cx, cz, worlds.terrain_chunk(cx, cz, SEED)


def generate_world(ox, oz, w, d):
    """{(cx,cz): u32[65536]} for the window; deterministic in (cx, cz, SEED).
    Cached under /tmp ("generator output cached to disk", BASELINE.json
    configs[1]); a cached world also means no fork() in a profiled process."""
    keys = [(cx, cz) for cz in range(oz, oz + d) for cx in range(ox, ox + w)]
    cache = "/tmp/vxbench_world_s%d_%d_%d_%d_%d.npy" % (SEED, ox, oz, w, d)
    if os.path.exists(cache):
        try:
            arr = np.load(cache)
            if arr.shape == (len(keys), abi.CHUNK_VOL):
                return {k: arr[i] for i, k in enumerate(keys)}
        except (OSError, ValueError):
            pass
    res = _generate(keys)
    try:
        tmp = cache + ".%d.tmp.npy" % os.getpid()
        np.save(tmp, np.stack([res[k] for k in keys]))
        os.replace(tmp, cache)
    except OSError:
        pass
    return res


def _generate(keys):
    procs = max(1, min(len(os.sched_getaffinity(0)), 48))
    if procs > 1:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_gen, keys, chunksize=max(1, len(keys) // (procs * 4)))
    else:
        res = [_gen(k) for k in keys]
    return {(cx, cz): v for cx, cz, v in res}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20", "-i", str(self.device)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout at the first collective; stdout
        # must carry exactly one JSON line, so the banner is sent to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier(device_ids=[local])
            torch.cuda.synchronize(local)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    return rank, world, local


def barrier_and_max(world, local, value):
    if world == 1:
        return value
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device="cuda:%d" % local)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world, local):
    if world > 1:
        import torch
        import torch.distributed as dist
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize(local)


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def run_ours(args):
    from vxgpu import gpu
    rank, world, local = dist_setup(args.gpus)
    table = ct.ContentTable()
    # the whole job is a (32N+2) x 34 world; rank r owns 32 of its lit chunk
    # columns and holds a one-chunk halo (vxgpu/shard.py)
    sh = shard.plan(-1, -1, SIDE * world + 2, SIDE + 2, world)[rank]
    ox, oz, w, d = sh.window
    t_gen = time.time()
    chunks = generate_world(ox, oz, w, d)
    t_gen = time.time() - t_gen
    all_keys, owned, lit = sh.resident, sh.owned, sh.lit
    gw = gpu.GpuWorld(table, ox, oz, w, d, device=local)
    lib = gw.lib
    all_keys_l, owned_l, lit_l = all_keys, owned, lit
    # ctypes key arrays are built once, outside every timed region
    all_keys, owned, lit = (gpu.keys_array(k) for k in (all_keys_l, owned_l, lit_l))

    # pinned host copies of the inputs (the e2e leg uploads from these)
    n_all = len(all_keys)
    vox_bytes = abi.CHUNK_VOL * 4
    h_vox = lib.vx_host_alloc(n_all * vox_bytes)
    assert h_vox, "pinned allocation failed"
    cin = (abi.ChunkIn * n_all)()
    for i, k in enumerate(all_keys_l):
        C.memmove(h_vox + i * vox_bytes, chunks[k].ctypes.data, vox_bytes)
        cin[i].cx, cin[i].cz = k
        cin[i].voxels = h_vox + i * vox_bytes
        cin[i].light = None
    del chunks
    gw.put_chunks_raw(cin, n_all)

    def step():
        gw.light_clear()
        gw.prebuild_batch(all_keys)
        gw.light_build(lit, True)
        gw.mesh_build(owned, CAPACITY)

    # clocks / throttle reasons are sampled from the warm-up to the end of the e2e
    # leg (the device-timed region alone lasts a few tens of milliseconds)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier(world, local)
    l0 = gw.timing()[2]
    gw.timer_begin()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dev_ms = gw.timer_end()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = gw.timing()[2] - l0
    barrier(world, local)
    ms = barrier_and_max(world, local, max(dev_ms, 0.0))
    wall_ms = barrier_and_max(world, local, wall_ms)
    total_chunks = len(owned) * world
    value = total_chunks * args.steps / (ms * 1e-3)

    # ---- mesh volume of one step (for algorithmic bytes)
    sizes = gw.mesh_arena_sizes()
    n_vf = sum(gw.mesh_info(i).n_vertex_floats for i in range(len(owned)))
    n_ix = sum(gw.mesh_info(i).n_indices for i in range(len(owned)))
    n_en = sum(gw.mesh_info(i).n_entries for i in range(len(owned)))
    n_ef = sum(gw.mesh_info(i).n_entry_floats for i in range(len(owned)))
    mesh_bytes = 4 * (n_vf + n_ix + n_ef) + 20 * n_en
    # read voxels + write lightmap + write mesh, per owned chunk
    alg_step = len(owned) * (262144 + 131072) + mesh_bytes

    # ---- per-kernel profile (events around every launch) on extra steps
    gw.profile_enable(True)
    PROF_STEPS = 2
    for _ in range(PROF_STEPS):
        step()
    prof = gw.profile()
    gw.profile_enable(False)
    per_step = {k: (v[0] / PROF_STEPS, v[1] / PROF_STEPS) for k, v in prof.items()}
    dom = max(per_step, key=lambda k: per_step[k][0])
    n_solve = len(set((cx + dx, cz + dz) for (cx, cz) in lit_l for dx in (-1, 0, 1)
                      for dz in (-1, 0, 1)) & set(all_keys_l))
    alg_kernel = {
        # algorithmic bytes per step of each kernel class
        # emit reads the unit list (16 B per unit) and the light cells around each
        # face, and writes the mesh; classify reads the voxels and writes the units
        "k_mesh_emit": len(owned) * 131072 + mesh_bytes,
        "k_mesh_classify": len(owned) * 262144,
        "k_solve": n_solve * 2 * 131072,
        "k_prebuild": n_all * 2 * 131072,
        "k_clear": n_all * 131072,
        "k_ystar": len(lit) * 131072,
        "k_seed": n_solve * (32768 + 8192),
    }
    # dram__bytes_read.sum + dram__bytes_write.sum per launch of each kernel class on
    # this workload, from the ncu --set full capture summarised in
    # profiles/r1s_ncu_full_summary.md (bytes)
    ncu_traffic = {"k_solve": 120899584 + 19842304, "k_mesh_emit": 168654080 + 538202368,
                   "k_mesh_classify": 119277568 + 40020992, "k_seed": 98705920 + 9662464,
                   "k_mesh_emit_generic": 19557888 + 16596736}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dom_ms, dom_launches = per_step[dom]
    achieved = alg_kernel.get(dom, 0) / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak,
        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
        "unit": "GB/s", "frac": round(achieved / peak, 4),
        "traffic": ncu_traffic.get(dom),
        "algorithmic_bytes_per_launch": int(alg_kernel.get(dom, 0) / max(dom_launches, 1)),
        "launch_ms": round(dom_ms / max(dom_launches, 1), 4),
        "launches_per_step": dom_launches,
        "whole_step": {"algorithmic_bytes": alg_step,
                       "achieved": round(alg_step / (ms / args.steps * 1e-3) / 1e9, 1),
                       "frac": round(alg_step / (ms / args.steps * 1e-3) / 1e9 / peak, 4)},
        "kernel_ms_per_step": {k: round(v[0], 4) for k, v in sorted(per_step.items())},
    }

    # ---- end to end: host buffers in, host buffers out, copies inside the timed region.
    # The rank's columns are cut into E2E_SPLIT sub-shards (same planner, same
    # duplicated halo), each with its own vx_ctx driven by its own host thread,
    # so one sub-shard's uploads / downloads overlap another's kernels.
    addr = {k: h_vox + i * vox_bytes for i, k in enumerate(all_keys_l)}
    subs = shard.plan(-1, -1, SIDE * world + 2, SIDE + 2, world * args.e2e_split)[
        rank * args.e2e_split:(rank + 1) * args.e2e_split]

    class Sub:
        pass
    units = []
    for sh2 in subs:
        u = Sub()
        u.gw = gw if args.e2e_split == 1 else gpu.GpuWorld(table, *sh2.window, device=local)
        u.n = len(sh2.resident)
        u.cin = (abi.ChunkIn * u.n)()
        for i, k in enumerate(sh2.resident):
            u.cin[i].cx, u.cin[i].cz = k
            u.cin[i].voxels = addr[k]
            u.cin[i].light = None
        u.res, u.own, u.lit = (gpu.keys_array(x) for x in (sh2.resident, sh2.owned, sh2.lit))
        u.h_light = lib.vx_host_alloc(len(sh2.owned) * abi.CHUNK_VOL * 2)
        u.light_bytes = len(sh2.owned) * abi.CHUNK_VOL * 2
        u.bufs = None
        units.append(u)

    def step_e2e(u):
        u.gw.put_chunks_raw(u.cin, u.n)          # H2D voxels; lightmaps start at zero
        u.gw.prebuild_batch(u.res)
        u.gw.light_build(u.lit, True)
        u.gw.mesh_build(u.own, CAPACITY)
        sz = u.gw.mesh_arena_sizes()
        if u.bufs is None or any(a > b for a, b in zip(sz, u.cap)):  # first step only
            if u.bufs:
                for q in u.bufs:
                    lib.vx_host_free(q)
            u.cap = [int(x * 1.05) + 1024 for x in sz]
            u.bufs = [lib.vx_host_alloc(u.cap[0] * 4), lib.vx_host_alloc(u.cap[1] * 4),
                      lib.vx_host_alloc(u.cap[2] * 20), lib.vx_host_alloc(u.cap[3] * 4)]
            assert all(u.bufs), "pinned allocation failed"
        u.gw.get_light_batch(u.own, u.h_light)  # D2H lightmaps
        u.gw.mesh_read_arena(*u.bufs)           # D2H meshes
        u.sz = sz

    def run_all(steps):
        if len(units) == 1:
            for _ in range(steps):
                step_e2e(units[0])
            return
        ths = [threading.Thread(target=lambda u=u: [step_e2e(u) for _ in range(steps)]) for u in units]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    run_all(1)
    barrier(world, local)
    t0 = time.perf_counter()
    E2E_STEPS = max(2, min(args.steps, 5))
    run_all(E2E_STEPS)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier(world, local)
    e2e_ms = barrier_and_max(world, local, e2e_ms)
    d2h = sum(u.light_bytes + 4 * (u.sz[0] + u.sz[1] + u.sz[3]) + 20 * u.sz[2] for u in units)
    e2e = {
        "value": round(total_chunks * E2E_STEPS / (e2e_ms * 1e-3), 1), "unit": "chunks/s",
        "steps": E2E_STEPS, "ms_per_step": round(e2e_ms / E2E_STEPS, 3),
        "h2d_bytes_per_step": sum(u.n for u in units) * vox_bytes * world,
        "d2h_bytes_per_step": d2h * world,
        "contexts_per_gpu": len(units),
        "api": "vx_chunks_put + vx_light_prebuild + vx_light_build + vx_mesh_build + "
               "vx_light_get_batch + vx_mesh_read_arena (pinned host buffers), "
               "%d sub-shard contexts per GPU on host threads" % len(units),
    }
    clocks = sampler.stop() if rank == 0 else None
    for u in units:
        lib.vx_host_free(u.h_light)
        for q in (u.bufs or []):
            lib.vx_host_free(q)
        if u.gw is not gw:
            u.gw.close()

    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": round(value, 1), "unit": "chunks/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
            "wall_ms_per_step": round(wall_ms / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16",
            "data": "synthetic",
            "config": {"workload": "configs[1]: 1,024 terrain chunks (32x32 inner + 1-chunk ring "
                                   "= 1,156 resident) per GPU, native demo-like terrain seed 2019, "
                                   "full light + mesh rebuild",
                       "chunks_per_gpu": len(owned), "resident_chunks_per_gpu": n_all,
                       "capacity": CAPACITY, "l2": "inputs larger than L2 (444 MB resident per GPU)",
                       "sharding": "chunk-column ranges, 1-chunk duplicated halo, no collective",
                       "world_gen_s": round(t_gen, 1)},
            "gpu_launches": int(launches),
            "faces_per_chunk": round(n_ix / 6 / max(1, len(owned)), 1),
            "clocks": clocks, "e2e": e2e, "roofline": roofline,
        }
        out["cpu_baseline"] = cpu_baseline(table, sample_side=SIDE)
    lib.vx_host_free(h_vox)
    gw.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


# ---------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/)
# ---------------------------------------------------------------------------
def _cpu_world(table, side):
    import pyoracle
    kind = pyoracle.best_kind()
    ox, oz, w, d = -1, -1, side + 2, side + 2
    chunks = generate_world(ox, oz, w, d)
    ow = pyoracle.OracleWorld(table, ox, oz, w, d, kind)
    keys = sorted(chunks.keys(), key=lambda k: (k[1], k[0]))
    for k in keys:
        ow.put_chunk(k[0], k[1], chunks[k])
    inner = [(cx, cz) for (cx, cz) in keys if 0 <= cx < side and 0 <= cz < side]
    return ow, kind, keys, inner


def _cpu_step(ow, keys, inner, threads):
    """One CPU step = the same work as a GPU step: clear, prebuild, light on one
    thread (the reference lights on the main thread), mesh on the thread pool."""
    t0 = time.perf_counter()
    ow.light_clear()
    for k in keys:
        ow.prebuild_sky(*k)
    t_light = ow.light_batch(inner, True)
    t1 = time.perf_counter()
    t_mesh, faces = ow.mesh_batch(inner, CAPACITY, threads)
    return time.perf_counter() - t0, t_light + (t1 - t0 - t_light), t_mesh, faces


def cpu_baseline(table, sample_side):
    cores = len(os.sched_getaffinity(0))
    ow, kind, keys, inner = _cpu_world(table, sample_side)
    total, t_light, t_mesh, faces = _cpu_step(ow, keys, inner, cores)
    ow.close()
    return {"value": round(len(inner) / total, 1), "unit": "chunks/s", "cores": cores,
            "kind": kind,
            "sample": "%d chunks (%dx%d inner of the same terrain world), one step: clear + "
                      "prebuild + light on 1 thread (%.2f s) then mesh on %d threads (%.2f s)"
                      % (len(inner), sample_side, sample_side, t_light, cores, t_mesh),
            "light_chunks_per_s_1thread": round(len(inner) / t_light, 1),
            "mesh_chunks_per_s": round(len(inner) / t_mesh, 1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    table = ct.ContentTable()
    cores = len(os.sched_getaffinity(0))
    side = 16  # bounded sample: 256 of the 1,024 chunks per step
    ow, kind, keys, inner = _cpu_world(table, side)
    for _ in range(args.warmup):
        _cpu_step(ow, keys, inner, cores)
    t0 = time.perf_counter()
    tl = tm = 0.0
    for _ in range(args.steps):
        _, a, b, faces = _cpu_step(ow, keys, inner, cores)
        tl += a
        tm += b
    total = time.perf_counter() - t0
    ow.close()
    value = len(inner) * args.steps / total
    sample = ("%d chunks per step (%dx%d inner + ring of the configs[1] terrain world): clear + "
              "prebuild + light on 1 thread then mesh on %d threads" % (len(inner), side, side, cores))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": "chunks/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(total / args.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": {"workload": "configs[1] terrain chunks, bounded sample", "sample": sample},
        "cpu_baseline": {"value": round(value, 1), "unit": "chunks/s", "cores": cores, "kind": kind,
                         "sample": sample,
                         "light_s_per_step": round(tl / args.steps, 3),
                         "mesh_s_per_step": round(tm / args.steps, 3)},
        "e2e": {"value": round(value, 1), "unit": "chunks/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-split", type=int, default=4, help="sub-shard contexts per GPU in the e2e leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
