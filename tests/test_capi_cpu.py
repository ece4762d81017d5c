"""The C-ABI shared library on a box WITHOUT a GPU: it loads, exports every symbol include/drc_b200.h declares,
the host-only model API works, and every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from tests.conftest import LINK, SRDF, URDF

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def capi():
    from dyros_robot_controller_b200 import _capi, build
    build.build()                      # no-op when the in-tree .so matches the sources
    return _capi


def header_symbols():
    text = (ROOT / "include" / "drc_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(drc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(capi):
    L = capi.lib()
    names = header_symbols()
    assert len(names) >= 55
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"libdrc_b200.so does not export {missing}"
    assert sorted(capi.SYMBOLS) == names       # the ctypes table and the header agree
    assert L.drc_version() >= 100


def test_header_cites_the_reference_interface():
    text = (ROOT / "include" / "drc_b200.h").read_text()
    for cite in ("robot_data.cpp:7-70", "robot_data.cpp:91-124", "robot_controller.cpp", "QP_base.h"):
        assert cite in text


def test_model_api_is_host_only(capi):
    import dyros_robot_controller_b200 as drc
    m = drc.Model(URDF, SRDF)
    assert m.dof == 7 and m.info["pairs"] > 0 and m.info["skipped_meshes"] == 0
    assert m.frame_id(LINK) >= 0 and m.frame_id("no_such_link") == -1
    assert np.allclose(m.q_upper, [2.7437, 1.7837, 2.9007, -0.1518, 2.8065, 4.5169, 3.0159])
    assert m.joint_names[0] == "fr3_joint1" and "fr3_link8" in m.frame_names
    assert "Total nq = 7" in m.verbose()
    # missing URDF: an error code (the reference calls std::exit, robot_data.cpp:15-19)
    with pytest.raises(capi.DrcError) as e:
        drc.Model("/nonexistent/robot.urdf")
    assert e.value.code == -2 and "does not exist" in str(e.value)
    # malformed URDF: parse error, no crash
    h = C.c_void_p()
    rc = capi.lib().drc_model_create_from_text(b"<robot><link/></robot", None, C.byref(h))
    assert rc == -3


def test_compute_entry_points_fail_loudly_without_a_gpu(capi):
    import dyros_robot_controller_b200 as drc
    if drc.device_count() > 0:
        pytest.skip("a CUDA device is present")
    m = drc.Model(URDF, SRDF)
    with pytest.raises(capi.DrcError) as e:
        drc.Context(m, 16)
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)
    with pytest.raises(capi.DrcError):
        drc.fp64_peak_tflops(0)


def test_mobile_base_fails_loudly_without_a_gpu_and_type_records(capi):
    """Mobile::KinematicParam / JointIndex / ActuatorIndex mirrors are plain host records; the base itself needs the GPU."""
    import dyros_robot_controller_b200 as drc
    from dyros_robot_controller_b200.drc import ActuatorIndex, DriveType, JointIndex, KinematicParam
    kp = KinematicParam(type=DriveType.Mecanum, wheel_radius=0.12, roller_angles=[-0.78, 0.78, 0.78, -0.78],
                        base2wheel_positions=[np.array([0.2, 0.2]), np.array([0.2, -0.2]), np.array([-0.2, 0.2]), np.array([-0.2, -0.2])],
                        base2wheel_angles=[0, 0, 0, 0])
    d = kp.as_dict()
    assert d["type"] == 1 and d["max_lin_speed"] == 2.0 and len(d["base2wheel_positions"]) == 4   # defaults of drc/type_define.py:15-18
    with pytest.raises(AssertionError):
        KinematicParam(type=DriveType.Differential, wheel_radius=0.1)          # base_width is required (type_define.py:35-36)
    with pytest.raises(AssertionError):
        KinematicParam(type=DriveType.Caster, wheel_radius=0.05, base2wheel_positions=[(0.2, 0.1), (-0.2, -0.1)])   # wheel_offset
    assert JointIndex(0, 7, 3).as_dict() == dict(virtual_start=0, mani_start=7, mobi_start=3)
    assert ActuatorIndex(4, 0).as_dict() == dict(mani_start=4, mobi_start=0)
    if drc.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.DrcError) as e:
        drc.MobileBase(d)
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)


def test_product_package_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may import or call it."""
    pkg = ROOT / "dyros_robot_controller_b200"
    pat = re.compile(r"^\s*(import\s+oracle|from\s+oracle|#\s*include\s+[\"<].*(oracle|kernel_emu))|liboracle|libdrc_emu|CDLL\(.*oracle", re.M)
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.h")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")):
        assert not pat.search(p.read_text()), p
