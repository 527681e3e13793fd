"""C-ABI checks that need no GPU: the library loads, exports every symbol the header declares,
the ctypes structs have the C layout, and the host-side validation rejects bad calls."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

from conftest import ROOT
from diffnet_b200 import _lib as L

HEADER = os.path.join(ROOT, "include", "diffnet_fem.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(L.LIB_PATH):
        from diffnet_b200.build import build
        build(verbose=False)
    return L.lib()


def header_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dn_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = header_functions()
    assert len(names) >= 13
    assert sorted(L.PROTOTYPES) == names, "ctypes table and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (dn_[a-z0-9_]+)", out))
    assert set(names) <= exported


def test_abi_version(lib):
    assert lib.dn_abi_version() == 1


def test_struct_layout_matches_c(tmp_path):
    """sizeof/offsetof of every struct, compiled from the header with gcc."""
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "diffnet_fem.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(dn_field), sizeof(dn_mask),'
                   ' sizeof(dn_geom), sizeof(dn_consts), offsetof(dn_mask, value), offsetof(dn_geom, hx),'
                   ' offsetof(dn_geom, mean_count), offsetof(dn_consts, reduction), sizeof(dn_slab_link),'
                   ' offsetof(dn_slab_link, put_plane), offsetof(dn_slab_link, max_spins), offsetof(dn_slab_link, world));'
                   ' return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(L.dn_field), C.sizeof(L.dn_mask), C.sizeof(L.dn_geom), C.sizeof(L.dn_consts),
            L.dn_mask.value.offset, L.dn_geom.hx.offset, L.dn_geom.mean_count.offset,
            L.dn_consts.reduction.offset, C.sizeof(L.dn_slab_link), L.dn_slab_link.put_plane.offset,
            L.dn_slab_link.max_spins.offset, L.dn_slab_link.world.offset]
    assert got == want


def test_status_codes_and_flags_match_the_header(tmp_path):
    """dn_status values, DN_F_LOAD_VECTOR and offsetof(dn_consts, flags) as gcc sees them in the header against _lib.py."""
    src = tmp_path / "codes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "diffnet_fem.h"\n'
                   'int main(){printf("%d %d %d %d %d %d %d %zu\\n", DN_OK, DN_EINVAL, DN_EARCH, DN_ECUDA, DN_EWORKSPACE,'
                   ' DN_ENOSTREAM, DN_F_LOAD_VECTOR, offsetof(dn_consts, flags)); return 0;}\n')
    exe = tmp_path / "codes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    assert got == [L.DN_OK, L.DN_EINVAL, L.DN_EARCH, L.DN_ECUDA, L.DN_EWORKSPACE, L.DN_ENOSTREAM, L.DN_F_LOAD_VECTOR,
                   L.dn_consts.flags.offset]


def test_workspace_query_is_host_only(lib):
    g = L.dn_geom(2, 64, 256, 256, 1, 2, 1 / 255, 1 / 255, 0.0, 0, 0, 0.0)
    n = lib.dn_fem_workspace_bytes(C.byref(g))
    assert 64 < n < (1 << 24)
    g3 = L.dn_geom(3, 16, 64, 64, 64, 2, 1 / 63, 1 / 63, 1 / 63, 0, 0, 0.0)
    assert 64 < lib.dn_fem_workspace_bytes(C.byref(g3)) < (1 << 24)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_no_cpu_fallback(lib):
    """Without a CUDA device the entry points fail loudly; CPU tensors are rejected up front."""
    assert lib.dn_device_check() != 0
    assert lib.dn_last_error()
    from diffnet_b200 import DiffNet2DFEM
    fem = DiffNet2DFEM(None, domain_size=8)
    u = torch.zeros(1, 1, 8, 8)
    with pytest.raises(L.DiffNetFEMError, match="no CPU fallback"):
        fem.energy_loss(u)
    with pytest.raises(L.DiffNetFEMError, match="no CPU fallback"):
        fem.gauss_pt_evaluation(u)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under diffnet_b200/ may reference it."""
    pkg = os.path.join(ROOT, "diffnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".inl")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "/root/reference" not in text, fn
