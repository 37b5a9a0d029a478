"""CPU: the C-ABI shared library builds/loads and exports every symbol include/fluidsolver_b200.h declares.
No compute calls are made here (there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

from conftest import REPO


def _declared_functions():
    text = open(os.path.join(REPO, "include", "fluidsolver_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fs_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from solver import _native as N
    assert _declared_functions() == N.exported_symbols()


def test_library_exports_every_declared_symbol():
    from solver import _native as N
    if not os.path.exists(N.LIB_PATH):
        import build  # python-fluid-simulation_b200/build.py
        build.build_library()
    lib = ctypes.CDLL(N.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/fluidsolver_b200.h but not exported"
    N.load()
    assert N.load().fs_abi_version() == 1


def test_workspace_queries_without_gpu():
    from solver import _native as N
    lib = N.load()
    assert lib.fs_visc3d_workspace_bytes(0, 4, 4, N.FS_F32) == 0
    assert lib.fs_visc3d_workspace_bytes(4, 4, 4, 7) == 0
    b32 = lib.fs_visc3d_workspace_bytes(64, 64, 64, N.FS_F32)
    b64 = lib.fs_visc3d_workspace_bytes(64, 64, 64, N.FS_F64)
    NL = 65 * 65 * 68
    assert b32 >= 22 * NL * 4 + 7 * NL and b64 >= 22 * NL * 8 + 7 * NL and b64 > b32
    assert lib.fs_visc2d_workspace_bytes(32, 32, N.FS_F64) > 14 * 33 * 36 * 8
    assert lib.fs_press_workspace_bytes(16, 16, 16) > 0 and lib.fs_press_workspace_bytes(16, 16, 0) > 0


def test_product_has_no_cpu_fallback():
    """The solver classes must refuse to run without CUDA rather than fall back to the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ViscosityCGSolver3D((8, 8, 8), (1.0, 1.0, 1.0))
    # and nothing under the product package imports the oracle
    pkg = os.path.join(REPO, "python-fluid-simulation_b200")
    for root, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, fn)).read()
                assert "numpy_oracle" not in src and "import oracle" not in src and "from oracle" not in src, fn


def test_header_is_plain_c_and_links(tmp_path):
    """include/fluidsolver_b200.h must be consumable from C (no C++ types): compile a small C program against it with gcc,
    link it to the shared library and run the calls that need no GPU."""
    import shutil
    import subprocess
    from solver import _native as N
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    if not os.path.exists(N.LIB_PATH):
        import build
        build.build_library()
    src = tmp_path / "abi_check.c"
    src.write_text(r'''
#include <stdio.h>
#include "fluidsolver_b200.h"
int main(void) {
    fs_cg_stats st;                       /* plain C struct */
    st.iterations = 0;
    if (fs_abi_version() != FS_ABI_VERSION) return 2;
    if (fs_visc3d_workspace_bytes(0, 4, 4, FS_F64) != 0) return 3;          /* bad grid -> 0, no CUDA call */
    size_t v = fs_visc3d_workspace_bytes(32, 32, 32, FS_F64);
    size_t p = fs_press_workspace_bytes(32, 32, 32);
    size_t p2 = fs_press_workspace_bytes(32, 32, 0);
    if (!v || !p || !p2) return 4;
    /* every entry point a binding needs is a linkable C symbol */
    void* fns[] = {(void*)fs_visc3d_solve, (void*)fs_visc3d_pack, (void*)fs_visc3d_set_active_mode, (void*)fs_visc3d_set_cg_mode,
                   (void*)fs_visc2d_solve, (void*)fs_press_cg, (void*)fs_press_set_operator, (void*)fs_solidfrac3d, (void*)fs_solidfrac2d,
                   (void*)fs_dens3d_scatter, (void*)fs_dens3d_gather, (void*)fs_comm_create, (void*)fs_visc3d_set_peers};
    for (unsigned i = 0; i < sizeof(fns) / sizeof(fns[0]); ++i) if (!fns[i]) return 5;
    printf("%zu %zu %zu %d\n", v, p, p2, (int)st.iterations);
    return 0;
}
''')
    exe = tmp_path / "abi_check"
    libdir = os.path.dirname(N.LIB_PATH)
    cmd = [gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-Wno-pedantic", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe),
           "-L", libdir, "-lfluidsolver_b200", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    v, p, p2, _ = r.stdout.split()
    assert int(v) == N.load().fs_visc3d_workspace_bytes(32, 32, 32, N.FS_F64)


def test_set_option_accepts_known_switches_only():
    """fs_set_option: the tuning switches between result-equivalent kernel forms (no GPU needed to set them)."""
    from solver import _native as N
    lib = N.load()
    for name in ("resident_form", "k1_block", "k1_tile", "sparse_setup"):
        assert lib.fs_set_option(name.encode(), 1) == 0
        assert lib.fs_set_option(name.encode(), -1) == 0          # back to the default
    assert lib.fs_set_option(b"no_such_switch", 1) < 0
    assert b"no_such_switch" in lib.fs_last_error()
    assert lib.fs_set_option(None, 1) < 0
    with pytest.raises(N.NativeError):
        N.set_option("llred", 1)                                   # (an experiment that was measured and removed)
