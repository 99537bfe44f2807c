"""CPU: the C-ABI library loads without a GPU and exports every symbol include/mppi_b200.h declares; argument
validation that needs no device is exercised (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from mppi_b200 import capi
    capi.build()
    return capi


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "mppi_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mppi_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported_and_bound(capi):
    lib = capi.lib()
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in capi.SYMBOLS, f"{n} has no ctypes signature in capi.SYMBOLS"
    assert sorted(capi.SYMBOLS) == names


def test_abi_version_and_strerror(capi):
    lib = capi.lib()
    assert lib.mppi_abi_version() == 4
    assert lib.mppi_strerror(0) == b"ok"
    assert b"invalid" in lib.mppi_strerror(-1)
    assert b"terrain" in lib.mppi_strerror(-3)


def test_default_params_are_the_reference_values(capi):
    p = capi.default_params(1000, 100)           # config.yaml + kernel literals
    got = {f: getattr(p, f) for f, _ in p._fields_}
    exp = dict(K=1000, T=100, dt=0.045, u1_min=-1, u1_max=1, u2_min=-1, u2_max=1, v_min=0, v_max=2, w_min=-1, w_max=1,
               lam=0.3, r_wheels=1.2, filt_k=3.5, filt_a=0.96, opt_k=3.0, opt_a=0.92, wheel_offset=0.2, cw_path=100.5,
               cw_slope=50.5, cw_speed=0.5, cw_obs=25.0, lethal_thresh=0.99, lethal_penalty=1e5, near_goal_cut=2.0,
               speed_eps=1e-4, pf_eps=1e-6, pf_near_gain=10.0, slope_eps=1e-6, slope_gain=5.0, horizon=9.0,
               target_speed=2.0, cw_orient=0, cw_slope_path=0, cw_goal_angle=0, goal_angle_radius=0.5, cw_roll=0,
               cw_pitch=0, cw_effort=0, reserved=0)
    for k, v in exp.items():
        assert got[k] == pytest.approx(v, rel=1e-6), k


def test_struct_layouts_match_the_header(capi):
    assert C.sizeof(capi.MppiParams) == 4 * 4 + 30 * 4 + 4 + 7 * 4 + 4      # + input_model (v2) + optional critics, reserved (v3)
    assert C.sizeof(capi.MppiState) == 48
    assert C.sizeof(capi.MppiTerrain) == 40
    assert C.sizeof(capi.MppiOutputs) == 8 * 8
    assert C.sizeof(capi.MppiDebugDump) == 15 * 8


def test_argument_validation_without_a_device(capi):
    lib = capi.lib()
    assert lib.mppi_default_params(None, 10, 10) == -1
    p = capi.MppiParams()
    assert lib.mppi_default_params(C.byref(p), 0, 10) == -1
    assert lib.mppi_default_params(C.byref(p), 10, 1) == -1
    assert lib.mppi_partial_floats(100) == 204
    assert lib.mppi_destroy(None) == 0
    assert lib.mppi_get_outputs(None, None) == -1
    h = C.c_void_p()
    bad = capi.default_params(16, 8)
    bad.lam = 0.0
    assert lib.mppi_create(C.byref(bad), 0, 1, C.byref(h)) == -1


def test_product_fails_loudly_without_cuda(capi):
    """No CPU fallback: on a box without a GPU, creating a controller must raise, not silently compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mppi_b200.core import Core
    with pytest.raises(capi.MppiError):
        Core(64, 10)
    lib = capi.lib()
    h = C.c_void_p()
    p = capi.default_params(64, 10)
    rc = lib.mppi_create(C.byref(p), 0, 1, C.byref(h))
    assert rc == -2 and b"CUDA" in lib.mppi_strerror(rc)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "husky-rover-mppi-isaacsim_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle_c" not in src, f
                if f.endswith((".cu", ".cuh")):
                    assert "oracle/" not in src.replace("oracle/mppi_oracle.c", "").replace("oracle/det_math.h", ""), f


def test_integration_md_binding_stub_matches_the_abi(capi):
    """INTEGRATION.md shows the ctypes binding a maintainer of the reference would add; its structures must be the
    header's (same field order and sizes), or the stub would silently corrupt arguments."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "INTEGRATION.md")).read()
    ns = {"C": C}
    for name in ("MppiParams", "MppiTerrain", "MppiState"):
        m = re.search(r"class %s\(C\.Structure\):.*?\n\n" % name, src, re.S)
        assert m, name
        exec(m.group(0), ns)
        mine, theirs = getattr(capi, name), ns[name]
        assert C.sizeof(mine) == C.sizeof(theirs), name
        norm = lambda n: "lam" if n == "lambda_" else n                                  # noqa: E731
        assert [(norm(n), t) for n, t in theirs._fields_] == [(n, t) for n, t in mine._fields_], name


def test_default_literals_are_the_reference_source_literals_if_present(capi):
    """`mppi_default_params` hard-codes the literals the reference scatters through its kernels and launch argument
    lists.  With the reference tree mounted they are read back from the source text, line by line."""
    base = "/root/reference/thesis_master/warp_implementation/"
    if not os.path.exists(base + "critics_warp.py"):
        pytest.skip("reference tree not mounted")
    p = capi.default_params(1000, 100)
    f32 = lambda v: C.c_float(v).value                                                    # noqa: E731

    def line(path, n):
        return open(base + path).read().split("\n")[n - 1]

    def num(path, n, pattern):
        m = re.search(pattern, line(path, n))
        assert m, (path, n, line(path, n))
        return f32(float(m.group(1)))

    assert p.filt_k == num("MPPI_isaac.py", 548, r"^\s*([0-9.]+),") and p.filt_a == num("MPPI_isaac.py", 549, r"^\s*([0-9.]+)")
    assert p.opt_k == num("MPPI_isaac.py", 688, r"^\s*([0-9.]+),") and p.opt_a == num("MPPI_isaac.py", 689, r"^\s*([0-9.]+)")
    assert p.wheel_offset == num("projection_warp.py", 333, r"offset = ([0-9.]+)")
    c = "critics_warp.py"
    assert p.goal_angle_radius == num(c, 33, r"dist_to_goal < ([0-9.]+)")
    assert p.pf_eps == num(c, 111, r"epsilon = ([0-9.e-]+)") and p.slope_eps == num(c, 188, r"epsilon = ([0-9.e-]+)")
    assert p.pf_near_gain == num(c, 126, r"cost \+= ([0-9.]+) \*")
    assert p.slope_gain == num(c, 209, r"\(1\.0 \+ ([0-9.]+)\*ratio_l\)")
    assert p.lethal_thresh == num(c, 251, r"costmap_cost > ([0-9.]+)") and p.lethal_penalty == num(c, 252, r"\+= ([0-9.]+)")
    assert p.near_goal_cut == num(c, 285, r"dist_to_goal < ([0-9.]+)")
    assert p.speed_eps == num(c, 297, r"\+ ([0-9.]+)\)\s*$")
    assert p.cw_path == num(c, 325, r"\+= ([0-9.]+)\*_path_follow_critic")
    assert p.cw_slope == num(c, 327, r"\+= ([0-9.]+)\*_avoid_slope_wheels")
    assert p.cw_speed == num(c, 328, r"\+= ([0-9.]+)\*_maximise_speed")
    assert p.cw_obs == num(c, 329, r"\+= ([0-9.]+)\*_avoid_obstacle")


def test_ctypes_mirrors_follow_the_header_field_by_field(capi):
    """Field names, order and scalar types of every structure in include/mppi_b200.h against the ctypes mirrors of
    capi.py (sizes alone would not notice two swapped floats)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "mppi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    ctype = {"int32_t": C.c_int32, "float": C.c_float, "uint32_t": C.c_uint32}
    for name in ("MppiParams", "MppiTerrain", "MppiState", "MppiOutputs", "MppiDebugDump"):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, re.S).group(1)
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            m = re.match(r"(?:const )?(\w+) (.*)", decl)
            base, rest = m.group(1), m.group(2)
            for item in rest.split(","):
                item = item.strip()
                ptr = item.startswith("*")
                fields.append((item.lstrip("*").strip(), C.c_void_p if ptr else ctype[base]))
        mirror = [("lambda" if n == "lam" else n, t) for n, t in getattr(capi, name)._fields_]
        assert mirror == fields, name


def test_library_holds_sm100a_code_for_every_kernel_family(capi):
    """The shipped .so carries sm_100a SASS (no PTX-only fallback) for both fused kernels in every flavour, and BOTH
    instantiations of the monolithic kernel (throughput and low-occupancy) -- read back with cuobjdump, no GPU needed."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-elf", capi.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    assert "sm_100a" in out
    for ns in ("strict", "fast", "strict_xc", "fast_xc"):
        tag = f"N4mppi{len(ns)}{ns}"
        assert f"{tag}22mppi_fused_pipe_kernelILi3ELb0EE" in out, ns
        assert f"{tag}17mppi_fused_kernelILi3ELb0ELb0EE" in out, ns          # throughput instantiation
        assert f"{tag}17mppi_fused_kernelILi3ELb0ELb1EE" in out, ns          # low-occupancy instantiation
