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
        # same number of leading positional parameters (names may differ in spelling: kp vs Kp)
        if len(lead) < len(args):
            wrong.append((name, args, lead))
    assert not missing, f"{qual} lacks {missing}"
    assert not wrong, f"{qual}: fewer positional parameters than the reference: {wrong}"
