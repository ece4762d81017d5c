"""The Python face of the boundary: dyros_robot_controller_b200.drc must offer every public method of the reference's `drc`
package (fixture tests/golden/ref_api_surface.json, written by tools/dump_reference_api.py from /root/reference/drc with `ast`)
with the same leading positional arguments.  Extra trailing keyword arguments (max_batch, device, verbose defaults) are allowed."""
import importlib
import inspect
import json
from pathlib import Path

import pytest

API = json.loads((Path(__file__).resolve().parent / "golden" / "ref_api_surface.json").read_text())
CLASSES = [k for k in API if not k.startswith("type_define.DriveType")]


@pytest.mark.parametrize("qual", CLASSES)
def test_mirror_has_every_reference_method(qual):
    mod, cls = qual.rsplit(".", 1)
    M = importlib.import_module("dyros_robot_controller_b200.drc." + mod)
    C = getattr(M, cls)
    missing, wrong = [], []
    for name, args in API[qual].items():
        f = getattr(C, name, None)
        if f is None:
            missing.append(name)
            continue
        params = [p for p in inspect.signature(f).parameters.values() if p.name != "self"]
        lead = [p.name for p in params[:len(args)]]
        # same leading parameters, same NAMES: the reference's examples call with keywords (examples/python/fr3_controller.py:131-176)
        if lead != args:
            wrong.append((name, args, lead))
    assert not missing, f"{qual} lacks {missing}"
    assert not wrong, f"{qual}: parameters differ from the reference: {wrong}"


def test_extension_module_surface():
    """the pybind11 module `dyros_robot_controller_cpp_wrapper` exports every class and method of the reference's Boost.Python
    module (src/bindings.cpp:219-447; fixture written by tools/dump_reference_api.py).  Import only -- no GPU needed."""
    import sys
    pkg = Path(__file__).resolve().parents[1] / "dyros_robot_controller_b200"
    sys.path.insert(0, str(pkg))
    try:
        import dyros_robot_controller_cpp_wrapper as w
    finally:
        sys.path.remove(str(pkg))
    ref = json.loads((Path(__file__).resolve().parent / "golden" / "ref_bindings_surface.json").read_text())
    assert ref["module"] == w.__name__
    for cls, members in ref["classes"].items():
        C = getattr(w, cls, None)
        assert C is not None, f"class {cls} missing"
        missing = [mname for mname in members if not hasattr(C, mname)]
        assert not missing, f"{cls} lacks {missing}"
    # without a CUDA device construction fails loudly (no CPU fallback) -- and succeeds on the GPU box
    import dyros_robot_controller_b200 as drc
    if drc.device_count() == 0:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            w.ManipulatorRobotData(drc.FR3_URDF, drc.FR3_SRDF, "")
