"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol that
include/zrk_b200.h declares.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from zrk_modulation_b200 import build as zb
    zb.build()
    from zrk_modulation_b200 import _cabi
    return _cabi.load()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "zrk_b200.h")).read()
    return sorted(set(re.findall(r"\b(zrk_[a-z_]+)\s*\(", src)) - {"zrk_handle"})


def test_header_and_binding_agree():
    from zrk_modulation_b200 import _cabi
    assert declared_symbols() == sorted(_cabi.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.zrk_abi_version() == 1


def test_struct_sizes_match_header(lib, tmp_path):
    """sizeof() of every ABI struct as gcc sees the header == the ctypes mirror."""
    import subprocess
    from zrk_modulation_b200 import _cabi
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "zrk_b200.h"\nint main(){printf("%zu %zu %zu %zu",'
                   'sizeof(zrk_msg),sizeof(zrk_options),sizeof(zrk_batch_desc),sizeof(zrk_state_view));return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(_cabi.ZrkMsg), C.sizeof(_cabi.ZrkOptions), C.sizeof(_cabi.ZrkBatchDesc),
                     C.sizeof(_cabi.ZrkStateView)]
    assert sizes[0] == 48


def test_sass_is_sm100a(lib):
    """The shipped library carries sm_100a SASS only (no PTX JIT fallback for another arch)."""
    import shutil
    import subprocess
    from zrk_modulation_b200 import _cabi
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _cabi.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_gpu_means_loud_failure(lib):
    """Without a device zrk_create must fail (ZRK_ERR_NO_DEVICE / CUDA error), never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from zrk_modulation_b200 import synthetic_batch
    from zrk_modulation_b200.engine import Engine, ZrkError
    with pytest.raises(ZrkError) as e:
        Engine(synthetic_batch(1, 2, n_radars=1, noise_period=2))
    assert e.value.rc in (-2, -3)


def test_package_never_imports_oracle():
    """The product path must not route through oracle/ or the emulator."""
    pkg = os.path.join(ROOT, "zrk_modulation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "libzrk_emu" not in src and "libzrk_oracle" not in src, f
                if f.endswith(".py"):
                    assert "emu_lib" not in src, f
