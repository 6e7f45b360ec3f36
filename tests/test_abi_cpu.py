"""CPU: the C-ABI library loads, exports every symbol include/sfw_b200.h declares, the ctypes mirror matches
the header's struct layout, and — with no GPU — the product fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sfw_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sfw_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header():
    from social_force_window_planner_b200 import _lib
    lib = _lib.load()
    decl = _declared_functions()
    assert decl, "no functions parsed from the header"
    for name in decl:
        assert hasattr(lib, name), f"libsfw_b200.so does not export {name}"
    assert sorted(_lib.EXPORTS) == decl, "python EXPORTS list and header disagree"
    assert lib.sfw_abi_version() == 4


def test_struct_layout_matches_header(tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirror."""
    from social_force_window_planner_b200 import _abi
    probes = {
        "SfwParams": ["max_vel_x", "robot_radius", "social_weight", "vel_weight"],
        "SfwSfmParams": ["force_factor_desired", "relaxation_time"],
        "SfwRobot": ["x", "wpx", "agent_radius"],
        "SfwPed": ["x", "goal_radius", "has_goal", "id"],
        "SfwScene": ["robot", "costmap", "size_x", "resolution", "peds", "n_peds", "obstacles_xy", "n_footprint"],
        "SfwBest": ["valid", "index", "cost", "v", "w"],
        "SfwLimits": ["max_scenes", "max_cells"],
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for st, fields in probes.items():
        lines.append(f'printf("{st} %zu\\n", sizeof({st}));')
        for f in fields:
            lines.append(f'printf("{st}.{f} %zu\\n", offsetof({st}, {f}));')
    lines.append("return 0;}")
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    rename = {"lambda": "lambda_"}
    for line in out.splitlines():
        key, val = line.split()
        if "." in key:
            st, f = key.split(".")
            assert getattr(getattr(_abi, st), rename.get(f, f)).offset == int(val), key
        else:
            assert C.sizeof(getattr(_abi, key)) == int(val), key


def test_defaults_match_reference_header_values():
    from social_force_window_planner_b200 import _abi, _lib
    lib = _lib.load()
    p = _abi.SfwParams()
    lib.sfw_default_params(C.byref(p))
    q = _abi.default_params()
    for name, _ in _abi.SfwParams._fields_:
        assert getattr(p, name) == getattr(q, name), name
    # reference include/social_force_window_planner/sfw_planner.hpp:56-66
    assert (p.max_vel_x, p.sim_time, p.sim_granularity) == (0.7, 1.0, 0.025)
    assert (p.social_weight, p.costmap_weight, p.angle_weight, p.distance_weight, p.vel_weight) == (1.2, 2.0, 0.7, 1.0, 1.0)
    s = _abi.SfwSfmParams()
    lib.sfw_default_sfm_params(C.byref(s))
    assert (s.force_factor_social, s.lambda_, s.gamma, s.n, s.n_prime) == (2.1, 2.0, 0.35, 2.0, 3.0)


def test_no_cpu_fallback():
    """Without a CUDA device sfw_create must fail with SFW_ERR_CUDA and a message; nothing is computed."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from social_force_window_planner_b200.scorer import Scorer, SfwError
    with pytest.raises(SfwError) as ei:
        Scorer(0)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "social_force_window_planner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "sfw_oracle" not in txt and "libsfw_ref" not in txt, f
