"""Sanity check of the GPU parity tests themselves: shows that GPU-produced
bytes are what gets compared, and that the comparison can fail."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("voxelengine-cpp_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import cases, pyoracle
from vxgpu import gpu

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
print("modules:", cases.__file__, pyoracle.__file__, gpu.__file__)
print("oracle kind:", pyoracle.best_kind(), "| gpu lib:", gpu.LIB_PATH)

stats = {}
class Spy(gpu.GpuWorld):
    def close(self):
        if getattr(self, "h", None):  # __del__ calls close() again with h = None
            stats["timing"] = self.timing()
        super().close()

for name in ("c1_s1", "edits_s41", "random_s7_full_c800k"):
    case = cases.named_cases()[name]()
    got = cases.run_case(lambda t, ox, oz, w, d, s: Spy(t, ox, oz, w, d, s), case, batch_light=True)
    want = cases.run_case(lambda t, ox, oz, w, d, s: pyoracle.OracleWorld(t, ox, oz, w, d, pyoracle.best_kind(), s), case)
    lt, mt, launches = stats["timing"]
    print("\n==", name, "light_ms=%.3f mesh_ms=%.3f kernel_launches=%d" % (lt, mt, launches))
    shared = [k for k in got if got[k] is want[k] or np.shares_memory(np.asarray(got[k]), np.asarray(want[k]))]
    print("  arrays shared between gpu and oracle results:", shared)
    mk = case.mesh_keys[0]; lk = case.light_keys[0]
    for lab in ("mesh_v_%d_%d" % mk, "mesh_i_%d_%d" % mk, "mesh_e_%d_%d" % mk, "mesh_f_%d_%d" % mk):
        print("  %-14s gpu n=%d oracle n=%d" % (lab, got[lab].size, want[lab].size))
    L = got["light_%d_%d" % lk]
    print("  light: nonzero=%d distinct=%d S<15&S>0=%d rgb_nonzero=%d sample=%s" % (
        np.count_nonzero(L), np.unique(L).size, int(np.count_nonzero(((L >> 12) > 0) & ((L >> 12) < 15))),
        int(np.count_nonzero(L & 0x0FFF)), [hex(int(v)) for v in L[L != 0][::4001][:6]]))
    mism = [k for k in want if np.asarray(got[k]).tobytes() != np.asarray(want[k]).tobytes()]
    dig = cases.digest(got)
    print("  mismatches vs oracle:", mism[:5], "| golden mismatches:", [k for k in GOLDEN[name] if dig[k] != GOLDEN[name][k]][:5])

# negative controls: the comparison must be able to fail
case = cases.named_cases()["multi_s24_noexpand"]()
case.expand = True   # golden was made with expand=False
got = cases.run_case(lambda t, ox, oz, w, d, s: gpu.GpuWorld(t, ox, oz, w, d, s), case, batch_light=True)
dig = cases.digest(got)
bad = [k for k in GOLDEN["multi_s24_noexpand"] if dig[k] != GOLDEN["multi_s24_noexpand"][k]]
print("\nnegative control A (expand flipped): %d golden mismatches (must be > 0)" % len(bad))
case = cases.named_cases()["c1_s2"]()
got = cases.run_case(lambda t, ox, oz, w, d, s: gpu.GpuWorld(t, ox, oz, w, d, s), case, batch_light=True)
got["mesh_v_0_0"] = got["mesh_v_0_0"].copy(); got["mesh_v_0_0"][7] ^= 1
dig = cases.digest(got)
bad = [k for k in GOLDEN["c1_s2"] if dig[k] != GOLDEN["c1_s2"][k]]
print("negative control B (one vertex bit flipped): mismatches =", bad, "(must name mesh_v_0_0)")
