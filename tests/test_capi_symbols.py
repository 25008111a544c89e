"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every function that
include/kba_b200.h declares, agrees with the ctypes struct mirror, and fails loudly (no CPU fallback) on compute calls."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "kba_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kba_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from limo_b200 import capi
    L = capi.lib()
    names = _declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), "symbol %s declared in include/kba_b200.h but not exported" % n
    assert set(capi.SYMBOLS) == set(names)
    assert L.kba_version() >= 1


def test_struct_sizes_match_header(tmp_path):
    """sizeof() of every POD struct as the C compiler sees it == size of the ctypes mirror"""
    from limo_b200 import capi_types as T
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include "kba_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(kba_window),sizeof(kba_options),sizeof(kba_iteration),sizeof(kba_solve_summary),'
                    'sizeof(kba_result),sizeof(kba_eval_out),sizeof(kba_counters),sizeof(kba_lidar_options));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    mirror = [C.sizeof(t) for t in (T.KbaWindow, T.KbaOptions, T.KbaIteration, T.KbaSolveSummary, T.KbaResult,
                                    T.KbaEvalOut, T.KbaCounters, T.KbaLidarOptions)]
    assert sizes == mirror


def test_default_options_match_reference_defaults():
    """bundle_adjuster_keyframes.hpp:79-89 and robust_solving.hpp:93-108 defaults; product and oracle agree"""
    from limo_b200 import capi
    from oracle import oracle as orc
    a, b = capi.default_options(), orc.default_options()
    for name, _ in a._fields_:
        assert getattr(a, name) == getattr(b, name), name
    assert (a.depth_thres, a.reprojection_thres, a.depth_quantile, a.reprojection_quantile) == (0.16, 1.6, 0.95, 0.95)
    assert (a.trim_solver_iterations, a.final_solver_iterations, a.min_residual_groups) == (2, 100, 30)


def test_no_cpu_fallback():
    """without a GPU the product path refuses to compute instead of silently using a CPU implementation"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from limo_b200 import capi
    from limo_b200.adjuster import BundleAdjusterKeyframes
    with pytest.raises(capi.KbaError, match="no CUDA device"):
        capi.Handle(0)
    from tests import ref_scenes as rs
    b, *_ = rs.build_adjuster(None, (0, 0), (0, 0, 0, 0), [rs.mono_extrinsics()])
    with pytest.raises(capi.KbaError):
        b.solve()


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under limo_b200/ may reference it"""
    pkg = os.path.join(ROOT, "limo_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dp, f)).read()
                assert "kba_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
