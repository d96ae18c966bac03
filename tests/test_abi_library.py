"""The C-ABI library loads and exports every symbol include/hmcmt_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from tests.helpers import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hmcmt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b([a-z_0-9]+_?)\s*\(", text)
    return sorted({n for n in names if n.startswith(("hmcmt_", "factor_mumps", "solve_mumps", "destroy_mumps"))})


def test_build_and_symbols():
    from hmcmt2d_b200 import build, lib
    path = build.build()
    assert os.path.exists(path)
    declared = _declared_symbols()
    assert len(declared) >= 25
    handle = ctypes.CDLL(path)
    missing = [s for s in declared if not hasattr(handle, s)]
    assert not missing, missing
    assert set(lib.EXPORTED_SYMBOLS) <= set(declared) | {"hmcmt_version"}
    L = lib.load()
    assert b"sm_100a" in L.hmcmt_version()


def test_sass_is_blackwell_native():
    """DMMA (FP64 tensor core) and UBLKCP (TMA bulk copy) must be present in the sm_100a SASS."""
    import shutil
    import subprocess
    from hmcmt2d_b200 import build
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "DMMA.8x8x4" in sass
    assert "UBLKCP" in sass


def test_no_device_fails_loudly():
    """Without a CUDA device the product must refuse to run (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    import scipy.sparse as sp
    from hmcmt2d_b200 import lib
    A = sp.diags([np.full(5, 4.0 + 1j), np.full(4, -1.0), np.full(4, -1.0)], [0, 1, -1], format="csc")
    with pytest.raises(lib.HmcmtError) as e:
        lib.factorMUMPS(A, 1)
    assert e.value.code == -98


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "hmcmt2d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
