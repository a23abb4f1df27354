"""CPU tests of the product's host side: the C-ABI library loads, exports every symbol the header declares,
loads URDFs exactly like the oracle-side mirror of from_urdf, and fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import CHAIN32, FR3, ROOT
from oracle.rb_oracle import parse_urdf
from oracle.rb_oracle_np import ChainNP

HEADER = os.path.join(ROOT, "include", "rigidbody.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(multibody_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol(rb):
    from rigidbody_rs_b200 import _lib
    names = header_functions()
    assert len(names) >= 30
    for nm in names:
        assert hasattr(_lib.lib, nm), f"{nm} declared in include/rigidbody.h but not exported"
    assert set(names) == set(_lib.PROTOTYPES), set(names) ^ set(_lib.PROTOTYPES)


def test_header_is_valid_c_and_cpp(tmp_path):
    (tmp_path / "t.c").write_text('#include "rigidbody.h"\nint main(void){ RbChainDesc d; (void)d; return RB_OK; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    "-c", str(tmp_path / "t.c"), "-o", str(tmp_path / "t.o")], check=True)
    (tmp_path / "t.cpp").write_text('#include "rigidbody.h"\nint main(){ RbChainDesc d{}; (void)d; return RB_OK; }\n')
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-c", str(tmp_path / "t.cpp"), "-o", str(tmp_path / "t2.o")], check=True)


def _host_model(lib, path):
    mb = lib.multibody_new_from_urdf(path.encode())
    assert mb, lib.multibody_last_error()
    n = lib.multibody_n_joints(mb)
    R, t, m, h, I = np.empty((n, 3, 3)), np.empty((n, 3)), np.empty(n), np.empty((n, 3)), np.empty((n, 6))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.multibody_get_model(mb, dp(R), dp(t), dp(m), dp(h), dp(I)) == 0
    lib.multibody_free(mb)
    return n, R, t, m, h, I


@pytest.mark.parametrize("urdf", [FR3, CHAIN32])
def test_cpp_urdf_loader_matches_oracle_side_loader(rb, urdf):
    """csrc/rb_host_model.cpp (product) vs oracle/rb_oracle.py + rb_oracle_np.py (independent mirror of
    multibody.rs:65-77 / joint.rs:53-68 / inertia.rs:21-35)."""
    from rigidbody_rs_b200 import _lib
    n, R, t, m, h, I = _host_model(_lib.lib, urdf)
    c = ChainNP(parse_urdf(urdf))
    assert n == c.n
    np.testing.assert_allclose(R, c.Rp, rtol=0, atol=1e-15)       # product snaps 6e-17 -> 0
    np.testing.assert_array_equal(t, c.tp)
    np.testing.assert_array_equal(m, c.m)
    np.testing.assert_allclose(h, c.h, rtol=1e-15)
    I6 = np.stack([c.Io[:, 0, 0], c.Io[:, 0, 1], c.Io[:, 0, 2], c.Io[:, 1, 1], c.Io[:, 1, 2], c.Io[:, 2, 2]], 1)
    np.testing.assert_allclose(I, I6, rtol=1e-14, atol=1e-18)
    # the FR3's fixed rotations are exact signed permutations after snapping
    assert set(np.unique(R)) <= {-1.0, 0.0, 1.0}


def test_urdf_errors_are_statuses_not_crashes(rb, tmp_path):
    from rigidbody_rs_b200 import _lib
    lib = _lib.lib
    assert not lib.multibody_new_from_urdf(b"/nonexistent/robot.urdf")
    assert b"cannot open" in lib.multibody_last_error()
    bad = tmp_path / "bad.urdf"
    bad.write_text("<robot><link name='a'><joint></robot>")
    assert not lib.multibody_new_from_urdf(str(bad).encode())
    assert b"parse error" in lib.multibody_last_error()
    fixed = tmp_path / "fixed.urdf"
    fixed.write_text("<robot name='r'><link name='a'/><joint name='j' type='fixed'><origin xyz='0 0 0'/></joint></robot>")
    assert not lib.multibody_new_from_urdf(str(fixed).encode())
    assert b"no movable joint" in lib.multibody_last_error()
    yaxis = tmp_path / "y.urdf"
    yaxis.write_text("<robot name='r'><link name='a'><inertial><origin xyz='0 0 0'/><mass value='1'/>"
                     "<inertia ixx='1' ixy='0' ixz='0' iyy='1' iyz='0' izz='1'/></inertial></link>"
                     "<joint name='j' type='revolute'><origin xyz='0 0 0' rpy='0 0 0'/><axis xyz='0 1 0'/></joint></robot>")
    assert not lib.multibody_new_from_urdf(str(yaxis).encode())
    assert b"only +z" in lib.multibody_last_error()
    lib.multibody_free(None)            # null-safe like the reference (lib.rs:73-78)
    lib.multibody_free_result(None)


def test_null_and_bad_arguments(rb):
    from rigidbody_rs_b200 import _lib
    lib = _lib.lib
    out = C.c_void_p()
    assert lib.multibody_gpu_new(None, 0, C.byref(out)) == _lib.RB_ERR_NULL
    assert lib.multibody_rnea_batch(None, None, None, None, None, 1, 0, 0, 0, None) == _lib.RB_ERR_NULL
    assert lib.multibody_gpu_n_joints(None) == _lib.RB_ERR_NULL
    d = _lib.RbChainDesc()
    d.n_joints = 0
    assert lib.multibody_gpu_new(C.byref(d), 0, C.byref(out)) == _lib.RB_ERR_ARG
    d.n_joints = 65
    assert lib.multibody_gpu_new(C.byref(d), 0, C.byref(out)) == _lib.RB_ERR_UNSUPPORTED
    # branching tree -> unsupported (reference is serial-only, multibody.rs:148)
    n = 2
    R = np.tile(np.eye(3).reshape(-1), n); t = np.zeros(3 * n); m = np.ones(n); c = np.zeros(3 * n)
    Ic = np.tile(np.eye(3).reshape(-1), n); par = np.array([-1, -1], dtype=np.int32)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    d.n_joints = n
    d.parent_rot, d.parent_trans, d.mass, d.com, d.inertia_com = dp(R), dp(t), dp(m), dp(c), dp(Ic)
    d.parent = par.ctypes.data_as(C.POINTER(C.c_int32))
    assert lib.multibody_gpu_new(C.byref(d), 0, C.byref(out)) == _lib.RB_ERR_UNSUPPORTED
    assert b"serial" in lib.multibody_last_error()


def test_no_cpu_fallback(rb):
    """Without a CUDA device the engine refuses to exist; on a GPU box this test checks the opposite."""
    import torch
    if torch.cuda.is_available():
        mb = rb.Multibody.from_urdf(FR3)
        assert mb.kernel_variant == "fr3-specialised"
        return
    with pytest.raises(rb.RigidBodyError) as e:
        rb.Multibody.from_urdf(FR3)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)
    from rigidbody_rs_b200 import _lib
    mbh = _lib.lib.multibody_new_from_urdf(FR3.encode())
    z = (C.c_double * 7)()
    assert not _lib.lib.multibody_rnea(mbh, z, z, z)        # single-state call: NULL + message, never a CPU answer
    assert b"no CUDA device" in _lib.lib.multibody_last_error()
    _lib.lib.multibody_free(mbh)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under rigidbody_rs_b200/ may import, link or open it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rigidbody_rs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "rb_oracle" not in src and "oracle/" not in src and "import oracle" not in src, f


def test_bench_fp64_instruction_counts_match_the_built_library(built):
    """bench.py's executed-instruction roofline uses static SASS counts; keep them equal to the library's."""
    import importlib.util
    out = subprocess.run(["bash", os.path.join(ROOT, "tools", "sass_count.sh"),
                          os.path.join(ROOT, "rigidbody_rs_b200", "librigidbody_b200.so")],
                         capture_output=True, text=True, check=True).stdout
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
    for kernel, want in bench.FP64_INSTR.items():
        line = [l for l in out.splitlines() if l.startswith(kernel + "I7CtModelI6TabFr3dELb0E")]    # fp64, SoA instantiation
        assert len(line) == 1, out
        got = int(line[0].split("FP64 total")[1].split()[0])
        assert got == want, (kernel, got, want)


def _random_chain(n, seed):
    """A physically valid random serial chain with general (non-axis-aligned) fixed rotations."""
    rng = np.random.default_rng(seed)
    from scipy.spatial.transform import Rotation
    R = Rotation.random(n, random_state=seed).as_matrix()
    t = rng.uniform(-0.3, 0.3, (n, 3))
    m = rng.uniform(0.5, 4.0, n)
    c = rng.uniform(-0.1, 0.1, (n, 3))
    A = rng.uniform(-1, 1, (n, 3, 3))
    Ic = np.einsum("nij,nkj->nik", A, A) * 0.01 + np.eye(3) * 0.01
    return R, t, m, c, Ic


def test_jit_precompile_without_gpu(rb, tmp_path, monkeypatch):
    """The run-time specialisation compiles with NVRTC alone (no GPU): FR3 and a random 5-joint chain, then cache hits."""
    from rigidbody_rs_b200 import _lib
    monkeypatch.setenv("RIGIDBODY_B200_CACHE", str(tmp_path / "cache"))
    log = C.create_string_buffer(8192)
    assert _lib.lib.multibody_jit_precompile(None, FR3.encode(), log, 8192) == 0, (_lib.lib.multibody_last_error(), log.value)
    assert b"cache hit" not in log.value
    assert _lib.lib.multibody_jit_precompile(None, FR3.encode(), log, 8192) == 0
    assert b"cache hit" in log.value
    R, t, m, c, Ic = _random_chain(5, 11)
    d = _lib.RbChainDesc()
    keep = [np.ascontiguousarray(x) for x in (R, t, m, c, Ic)]
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    d.n_joints = 5
    d.parent_rot, d.parent_trans, d.mass, d.com, d.inertia_com = (dp(x) for x in keep)
    d.gravity[:] = [0.0, 0.0, 9.81]
    assert _lib.lib.multibody_jit_precompile(C.byref(d), None, log, 8192) == 0, (_lib.lib.multibody_last_error(), log.value)
    assert len(os.listdir(tmp_path / "cache")) == 2
    # chains longer than the register-resident limit are refused, not miscompiled
    assert _lib.lib.multibody_jit_precompile(None, CHAIN32.encode(), log, 8192) == _lib.RB_ERR_UNSUPPORTED
