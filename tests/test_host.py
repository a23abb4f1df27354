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
    # declarations the Rust bridge crate provides (not this library): #ifdef RIGIDBODY_HAVE_RUST_BRIDGE ... #endif
    src = re.sub(r"#ifdef RIGIDBODY_HAVE_RUST_BRIDGE.*?#endif", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(multibody_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol(rb):
    from rigidbody_rs_b200 import _lib
    names = header_functions()
    assert len(names) >= 30
    for nm in names:
        assert hasattr(_lib.lib, nm), f"{nm} declared in include/rigidbody.h but not exported"
    assert set(names) == set(_lib.PROTOTYPES), set(names) ^ set(_lib.PROTOTYPES)


def _c_prototypes():
    """name -> number of parameters, for every function include/rigidbody.h declares (Rust-bridge block excluded)."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"#ifdef RIGIDBODY_HAVE_RUST_BRIDGE.*?#endif", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(multibody_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_rust_crate_binds_the_whole_header():
    """rust/rigidbody_gpu_bindings cannot be compiled here (no cargo): instead its `extern "C"` block is held against
    include/rigidbody.h -- every Part-2 symbol (everything but the reference's own single-state symbols, which the Rust
    cdylib rigidbody_bindings keeps exporting) must be bound, with the same number of parameters, and nothing else."""
    rs = open(os.path.join(ROOT, "rust", "rigidbody_gpu_bindings", "src", "lib.rs")).read()
    block = rs[rs.index('extern "C" {'):]
    block = block[: block.index("\n}\n")]
    block = re.sub(r"//.*", "", block)
    rust = {}
    for m in re.finditer(r"pub fn (multibody_[a-z0-9_]+)\s*\(([^;]*?)\)\s*(->[^;]*)?;", block, flags=re.S):
        args = m.group(2).strip()
        rust[m.group(1)] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    c = _c_prototypes()
    part1 = {"multibody_new", "multibody_new_from_urdf", "multibody_fwd_kin", "multibody_jac", "multibody_rnea", "multibody_crba",
             "multibody_free", "multibody_free_result", "multibody_n_joints", "multibody_get_model", "multibody_get_chain"}
    want = {k: v for k, v in c.items() if k not in part1}
    assert set(rust) == set(want), sorted(set(rust) ^ set(want))
    for name, nargs in want.items():
        assert rust[name] == nargs, (name, rust[name], nargs)
    # the crate's one export is the bridge the header declares under RIGIDBODY_HAVE_RUST_BRIDGE
    assert "pub unsafe extern \"C\" fn multibody_gpu_from_rust(mb: *const Multibody, device: c_int, out: *mut *mut RbGpu) -> c_int" in rs
    assert "int multibody_gpu_from_rust(const Multibody* mb, int device, RbGpu** out);" in open(HEADER).read()


def _build_rust_double(tmp_path):
    exe = str(tmp_path / "rust_crate_double")
    libdir = os.path.join(ROOT, "rigidbody_rs_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "rust_crate_double.c"),
                    "-L" + libdir, "-lrigidbody_b200", "-Wl,-rpath," + libdir, "-lm", "-o", exe], check=True)
    return exe


def test_rust_crate_double_builds_and_flattens_the_chain(rb, tmp_path):
    """examples/rust_crate_double.c (the crate's call sequence in C99) builds warning-free against the header; the part
    that needs no GPU -- ChainArrays::from_multibody -> RbChainDesc -- is checked here: the arrays multibody_get_chain
    returns, fed back through a descriptor, give the same flattened model as the URDF loader, bit for bit."""
    import torch
    from rigidbody_rs_b200 import _lib
    exe = _build_rust_double(tmp_path)
    r = subprocess.run([exe, FR3], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout
    else:
        assert r.returncode == 2 and "no CPU fallback" in r.stdout
    lib = _lib.lib
    mb = lib.multibody_new_from_urdf(FR3.encode())
    n = lib.multibody_n_joints(mb)
    par = (C.c_int32 * n)()
    arrs = [np.empty(k * n) for k in (3, 9, 3, 1, 3, 9)]
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.multibody_get_chain(mb, par, *[dp(a) for a in arrs]) == 0
    assert list(par) == list(range(-1, n - 1))
    axis, R, t, m, com, Ic = arrs
    assert np.array_equal(axis.reshape(n, 3), np.tile([0.0, 0.0, 1.0], (n, 1)))
    from oracle.rb_oracle import parse_urdf
    mo = parse_urdf(FR3)
    assert np.array_equal(m, mo.mass) and np.array_equal(com.reshape(n, 3), mo.com) and np.array_equal(t.reshape(n, 3), mo.xyz)
    lib.multibody_free(mb)


def test_header_is_valid_c_and_cpp(tmp_path):
    # with and without the Rust bridge declaration
    (tmp_path / "t.c").write_text('#define RIGIDBODY_HAVE_RUST_BRIDGE 1\n#include "rigidbody.h"\nint main(void){ RbChainDesc d; (void)d; (void)multibody_gpu_from_rust; return RB_OK; }\n')
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
    mb = lib.multibody_new_from_urdf(str(yaxis).encode())          # non-z axes load (re-based to z on the host)
    assert mb and lib.multibody_n_joints(mb) == 1
    lib.multibody_free(mb)
    zero = tmp_path / "zero.urdf"
    zero.write_text(yaxis.read_text().replace("0 1 0", "0 0 0"))
    assert not lib.multibody_new_from_urdf(str(zero).encode())
    assert b"non-zero" in lib.multibody_last_error()
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
    # parents must come first (topological order); a valid tree gets past the loader (and stops at "no CUDA device" here)
    n = 2
    R = np.tile(np.eye(3).reshape(-1), n); t = np.zeros(3 * n); m = np.ones(n); c = np.zeros(3 * n)
    Ic = np.tile(np.eye(3).reshape(-1), n); par = np.array([-1, -1], dtype=np.int32)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    d.n_joints = n
    d.parent_rot, d.parent_trans, d.mass, d.com, d.inertia_com = dp(R), dp(t), dp(m), dp(c), dp(Ic)
    d.parent = par.ctypes.data_as(C.POINTER(C.c_int32))
    par[:] = [-1, 1]
    assert lib.multibody_gpu_new(C.byref(d), 0, C.byref(out)) == _lib.RB_ERR_ARG
    assert b"topological" in lib.multibody_last_error()
    par[:] = [-1, -1]
    import torch
    if not torch.cuda.is_available():
        assert lib.multibody_gpu_new(C.byref(d), 0, C.byref(out)) == _lib.RB_ERR_CUDA
    log = C.create_string_buffer(4096)
    assert lib.multibody_jit_precompile(C.byref(d), None, log, 4096) == 0, (lib.multibody_last_error(), log.value)   # trees are unrolled too


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
    for kernel, want in bench.FP64_STATIC.items():
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


AXES = [[1, 0, 0], [0, 1, 0], [0, 0, -1], [0, 0, 2], [1, 1, 0], [0.3, -0.5, 0.8], [-0.2, 0.1, -0.97]]


def _axis_urdf(path, R_rpy, t, m, c, Ic, axes):
    """Writes a serial-chain URDF with the given joint axes (k-th joint pairs with k-th link, multibody.rs:70)."""
    f = lambda v: " ".join(repr(float(x)) for x in v)
    out = ["<robot name='axes'>"]
    for i in range(len(m)):
        out.append(f"<link name='l{i}'><inertial><origin xyz='{f(c[i])}' rpy='0 0 0'/><mass value='{float(m[i])!r}'/>"
                   f"<inertia ixx='{float(Ic[i][0][0])!r}' ixy='{float(Ic[i][0][1])!r}' ixz='{float(Ic[i][0][2])!r}' iyy='{float(Ic[i][1][1])!r}' "
                   f"iyz='{float(Ic[i][1][2])!r}' izz='{float(Ic[i][2][2])!r}'/></inertial></link>")
    for i in range(len(m)):
        out.append(f"<joint name='j{i}' type='revolute'><origin xyz='{f(t[i])}' rpy='{f(R_rpy[i])}'/>"
                   f"<axis xyz='{f(axes[i])}'/><limit lower='-2' upper='2' velocity='2' effort='50'/></joint>")
    out.append("</robot>")
    path.write_text("\n".join(out))
    return str(path)


def _random_tree(n, seed):
    """parent[i] drawn from the earlier joints (or the base): a random kinematic tree in topological order."""
    rng = np.random.default_rng(1000 + seed)
    par = np.array([-1] + [int(rng.integers(-1 if i > 2 else 0, i)) for i in range(1, n)], dtype=np.int32)
    par[n - 1] = max(par[n - 1], 0)                      # give the tip at least one supporting joint
    return par


@pytest.mark.parametrize("tree", [False, True])
def test_twin_with_general_axes_is_self_consistent(tree):
    """The numpy twin's S = (axis; 0) and parent-index generalisations, checked against physics they do not assume:
    gravity torque = dU/dq (central differences), H = d tau / d ddq, H SPD, power balance of the Coriolis terms."""
    R, t, m, c, Ic = _random_chain(7, 5)
    ch = ChainNP.from_arrays(R, t, m, c, Ic, axis=AXES, parent=_random_tree(7, 3) if tree else None)
    if tree:
        assert not np.array_equal(ch.parent, np.arange(7) - 1)
    rng = np.random.default_rng(0)
    q, dq, ddq = rng.uniform(-2, 2, (3, 6, 7))
    g = ch.rnea(q, 0 * q, 0 * q)
    eps = 1e-6
    for j in range(7):
        e = np.zeros(7); e[j] = eps
        dU = (ch.potential_energy(q + e) - ch.potential_energy(q - e)) / (2 * eps)
        np.testing.assert_allclose(g[:, j], dU, rtol=0, atol=2e-8)
    H = ch.crba(q, symmetric=True)
    assert np.all(np.linalg.eigvalsh(H) > 0)
    np.testing.assert_allclose(ch.rnea(q, dq, ddq) - ch.rnea(q, dq, 0 * q), np.einsum("bij,bj->bi", H, ddq), rtol=0, atol=1e-12)
    # Coriolis terms: power balance  d/dt (1/2 dq^T H dq) = dq . (tau - g)  along ddq = 0  =>  dq^T Hdot dq / 2 = dq . c
    cor = ch.rnea(q, dq, 0 * q) - g
    h = 1e-6
    Hd = (ch.crba(q + h * dq, symmetric=True) - ch.crba(q - h * dq, symmetric=True)) / (2 * h)
    np.testing.assert_allclose(np.einsum("bi,bi->b", dq, cor), 0.5 * np.einsum("bi,bij,bj->b", dq, Hd, dq), rtol=0, atol=5e-8)


def test_axis_rebasing_on_the_host(rb, tmp_path):
    """csrc/rb_host_model.cpp turns non-z axes into z by re-basing the joint frames.  The re-based rows, run through
    the twin's z-only formulation, must give the tau / H / tip position of the chain with its axes as written."""
    from rigidbody_rs_b200 import _lib
    from scipy.spatial.transform import Rotation
    n = 7
    R, t, m, c, Ic = _random_chain(n, 21)
    rpy = Rotation.from_matrix(R).as_euler("xyz")                  # Rz(y) Ry(p) Rx(r), joint.rs:59-63
    path = _axis_urdf(tmp_path / "axes.urdf", rpy, t, m, c, Ic, AXES)
    n2, Rk, tk, mk, hk, Ik = _host_model(_lib.lib, path)
    assert n2 == n
    canon = ChainNP.__new__(ChainNP)
    canon.n, canon.Rp, canon.tp, canon.m, canon.h = n, Rk, tk, mk, hk
    canon.Io = np.stack([[[a[0], a[1], a[2]], [a[1], a[3], a[4]], [a[2], a[4], a[5]]] for a in Ik])
    canon.axis = np.tile(np.array([0.0, 0.0, 1.0]), (n, 1))
    canon.parent = np.arange(n) - 1
    direct = ChainNP.from_arrays(Rotation.from_euler("xyz", rpy).as_matrix(), t, m, c, Ic, axis=AXES)
    rng = np.random.default_rng(2)
    q, dq, ddq = rng.uniform(-2.5, 2.5, (3, 16, n))
    np.testing.assert_allclose(canon.rnea(q, dq, ddq), direct.rnea(q, dq, ddq), rtol=0, atol=2e-12)
    np.testing.assert_allclose(canon.crba(q), direct.crba(q), rtol=0, atol=1e-13)
    np.testing.assert_allclose(canon.fwd_kin(q)[1], direct.fwd_kin(q)[1], rtol=0, atol=1e-13)
    # +z joints keep their frame exactly when the previous joint did too (joint 3: axis (0,0,2) after (0,0,-1) does not)
    zz = ChainNP.from_arrays(R, t, m, c, Ic)
    path = _axis_urdf(tmp_path / "z.urdf", rpy, t, m, c, Ic, [[0, 0, 1]] * n)
    _, Rz_, tz_, *_ = _host_model(_lib.lib, path)
    np.testing.assert_allclose(Rz_, zz.Rp, rtol=0, atol=1e-15)
    np.testing.assert_array_equal(tz_, t)


def test_jit_cache_is_verified_not_trusted(rb, tmp_path, monkeypatch):
    """The on-disk kernel cache: a file is used only if its stored key material equals the request byte for byte and the
    image checksum holds; a flipped bit, a truncated file or a file planted under another chain's name is a miss (and is
    recompiled); a cache directory that group/others can write to is not used at all; no $HOME means no cache."""
    from rigidbody_rs_b200 import _lib
    cache = tmp_path / "cache"
    monkeypatch.setenv("RIGIDBODY_B200_CACHE", str(cache))
    log = C.create_string_buffer(8192)
    def desc(seed):
        keep = [np.ascontiguousarray(x) for x in _random_chain(2, seed)]
        d = _lib.RbChainDesc()
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        d.n_joints = 2
        d.parent_rot, d.parent_trans, d.mass, d.com, d.inertia_com = (dp(x) for x in keep)
        d.gravity[:] = [0.0, 0.0, 9.81]
        return d, keep
    d, keep = desc(11)
    pre = lambda: _lib.lib.multibody_jit_precompile(C.byref(d), None, log, 8192)
    assert pre() == 0 and b"cache hit" not in log.value
    assert (os.stat(cache).st_mode & 0o777) == 0o700
    (name,) = os.listdir(cache)
    path = cache / name
    assert (os.stat(path).st_mode & 0o777) == 0o600
    assert pre() == 0 and b"cache hit" in log.value
    good = path.read_bytes()
    # 1. bit rot inside the image
    bad = bytearray(good); bad[len(bad) // 2] ^= 0x10
    path.write_bytes(bytes(bad))
    assert pre() == 0 and b"cache hit" not in log.value          # detected, recompiled, rewritten
    assert path.read_bytes() == good
    # 2. truncation
    path.write_bytes(good[: len(good) - 100])
    assert pre() == 0 and b"cache hit" not in log.value
    # 3. another chain's (valid) file under this chain's name: the stored key material differs
    d2, keep2 = desc(12)
    assert _lib.lib.multibody_jit_precompile(C.byref(d2), None, log, 8192) == 0
    other = [n for n in os.listdir(cache) if n != name][0]
    path.write_bytes((cache / other).read_bytes())
    assert pre() == 0 and b"cache hit" not in log.value
    assert pre() == 0 and b"cache hit" in log.value
    # 4. a directory others can write to is not trusted: nothing is read from it or written to it
    loose = tmp_path / "loose"; loose.mkdir(); os.chmod(loose, 0o777)
    monkeypatch.setenv("RIGIDBODY_B200_CACHE", str(loose))
    assert pre() == 0 and b"cache hit" not in log.value
    assert os.listdir(loose) == []
    # 5. no cache directory configured and no HOME: compile every time, write nowhere
    monkeypatch.delenv("RIGIDBODY_B200_CACHE")
    monkeypatch.delenv("HOME", raising=False)
    assert pre() == 0 and b"cache hit" not in log.value


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
    # chains beyond the unrolled kernels' limit (32 joints) are refused, not miscompiled
    R, t, m, c, Ic = _random_chain(33, 12)
    keep = [np.ascontiguousarray(x) for x in (R, t, m, c, Ic)]
    d.n_joints = 33
    d.parent_rot, d.parent_trans, d.mass, d.com, d.inertia_com = (dp(x) for x in keep)
    assert _lib.lib.multibody_jit_precompile(C.byref(d), None, log, 8192) == _lib.RB_ERR_UNSUPPORTED


def _build_example(tmp_path):
    import subprocess
    exe = str(tmp_path / "batch_demo")
    libdir = os.path.join(ROOT, "rigidbody_rs_b200")
    subprocess.run(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "batch_demo.cpp"),
                    "-L" + libdir, "-lrigidbody_b200", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    return exe


def test_cpp_example_builds_against_the_c_abi(rb, tmp_path):
    """examples/batch_demo.cpp: a C++ program using only include/rigidbody.h links against the library; without a GPU
    it reports the error instead of computing anything."""
    import subprocess
    import torch
    exe = _build_example(tmp_path)
    r = subprocess.run([exe, FR3], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout
    else:
        assert r.returncode == 2 and "no CPU fallback" in r.stdout
