"""The C-ABI library loads on a machine without a GPU and exports exactly what include/dgtta.h declares."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "dgtta.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(dgtta_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def libpath():
    from dg_tta_b200 import build
    return build.build()


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for must in ("dgtta_mind_ssc_fwd", "dgtta_gin_fwd", "dgtta_gin_layer_fwd", "dgtta_affine_sample_fwd",
                 "dgtta_affine_sample_bwd_input", "dgtta_mind_workspace_bytes", "dgtta_gin_workspace_bytes"):
        assert must in names


def test_library_exports_every_declared_symbol(libpath):
    handle = ctypes.CDLL(str(libpath))
    for name in declared_symbols():
        assert hasattr(handle, name), f"{name} declared in dgtta.h but not exported"
    assert handle.dgtta_abi_version() == 1


def test_python_binding_covers_the_header(libpath):
    from dg_tta_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert _lib.lib().dgtta_abi_version() == 1


def test_no_torch_or_cxx_types_in_the_abi(libpath):
    out = subprocess.run(["nm", "-D", "--defined-only", str(libpath)], capture_output=True, text=True).stdout
    exported = [ln.split()[-1] for ln in out.splitlines() if " T " in ln]
    assert all(not s.startswith("_Z") or "dgtta" in s for s in exported)
    assert not any("torch" in s or "at::" in s or "c10" in s for s in exported)


def test_argument_errors_without_a_gpu(libpath):
    """Pure host-side validation paths: no kernel is launched."""
    from dg_tta_b200 import _lib
    L = _lib.lib()
    assert L.dgtta_mind_workspace_bytes(0, 1, 1, 1) == 0
    assert L.dgtta_mind_workspace_bytes(1, 128, 128, 128) > 0
    assert L.dgtta_gin_workspace_bytes(2, 8, 8, 8, 1, 4, 2) >= 2 * 2 * 2 * 512 * 4
    rc = L.dgtta_mind_ssc_fwd(None, None, None, 1, 4, 4, 4, 1, None, 5, 0.0, 0, None, 0, 0, None, 0, None)
    assert rc == -2 and b"null" in L.dgtta_last_error()
    rc = L.dgtta_affine_sample_fwd(None, None, None, 1, 1, 4, 4, 4, 4, 4, 4, 0, 0, None)
    assert rc == -2
    # torch's offset bookkeeping for randn_like on 1x12x128^3 with 148 SMs x 2048 threads:
    # grid = 148*8 blocks of 256, unroll 4 -> ceil(25165824 / (256*1184*4)) * 4 = 21 * 4
    assert L.dgtta_mind_philox_offset_increment(1, 128, 128, 128, 148, 2048) == 84
