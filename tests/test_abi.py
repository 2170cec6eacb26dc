"""CPU-only: the C-ABI shared library builds, loads and exports every symbol
include/vxgpu.h declares; without a CUDA device it fails loudly instead of
falling back to a CPU path."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vxgpu.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vx_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from vxgpu import gpu
    if not os.path.exists(gpu.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "voxelengine-cpp_b200", "csrc")],
                              stdout=subprocess.DEVNULL)
    return gpu.load()


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "libvxgpu.so does not export %s" % n


def test_binding_lists_every_symbol():
    from vxgpu import gpu
    assert sorted(gpu.SYMBOLS) == declared_symbols()


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from vxgpu import gpu
    import cases
    with pytest.raises(gpu.VxError, match="CUDA device"):
        gpu.GpuWorld(cases.table(), 0, 0, 3, 3)


def test_struct_sizes_match_header():
    """ctypes mirrors (vxgpu/abi.py) against sizes computed by the C compiler."""
    from vxgpu import abi
    src = '#include "%s"\n#include <stdio.h>\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu\\n",' \
          'sizeof(vx_block_def),sizeof(vx_model_vertex),sizeof(vx_model_mesh),sizeof(vx_content),' \
          'sizeof(vx_chunk_in),sizeof(vx_chunk_info),sizeof(vx_edit),sizeof(vx_sort_entry),' \
          'sizeof(vx_mesh_info));return 0;}' % HEADER
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    mirrors = [abi.BlockDef, abi.ModelVertex, abi.ModelMesh, abi.Content, abi.ChunkIn,
               abi.ChunkInfo, abi.Edit, abi.SortEntry, abi.MeshInfo]
    assert sizes == [ctypes.sizeof(m) for m in mirrors]
