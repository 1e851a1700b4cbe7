"""CPU-side checks of the drop-in boundary: libomnissm.so loads, exports every symbol include/omnissm.h
declares, the ctypes structs mirror the header, and the reference-facing shim packages import."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "omnissm.h")


@pytest.fixture(scope="module")
def lib():
    from omnimamba_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        from omnimamba_b200.build import build
        build()
    return _cabi.lib()


def _declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"OMNI_API\s+[\w\s\*]+?\b(omni_\w+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/omnissm.h but not exported"


def test_cabi_covers_header():
    from omnimamba_b200 import _cabi
    assert sorted(list(_cabi.ENTRY_POINTS) + _cabi.OTHER_SYMBOLS) == _declared_symbols()


def test_struct_sizes_match_c_compiler(tmp_path, lib):
    """sizeof() of every params struct as gcc sees the header == ctypes.sizeof of the Python mirror."""
    from omnimamba_b200 import _cabi
    cname = {n: re.sub(r"^omni_", "", n) for n in _cabi.ENTRY_POINTS}
    hdr = open(HEADER).read()
    structs = {}
    for entry in _cabi.ENTRY_POINTS:
        m = re.search(r"int\s+%s\(const\s+(\w+)\*" % entry, hdr)
        assert m, entry
        structs[entry] = m.group(1)
    prog = '#include <stdio.h>\n#include "omnissm.h"\nint main(){\n'
    for entry, t in structs.items():
        prog += f'printf("{entry} %zu\\n", sizeof({t}));\n'
    prog += 'printf("tensor %zu\\n", sizeof(omni_tensor_t)); return 0; }\n'
    c = tmp_path / "sz.c"
    c.write_text(prog)
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    assert int(out["tensor"]) == ctypes.sizeof(_cabi.Tensor)
    for entry, struct in _cabi.ENTRY_POINTS.items():
        assert int(out[entry]) == ctypes.sizeof(struct), entry


def test_version_and_error_string(lib):
    assert lib.omni_version() == 8
    assert isinstance(lib.omni_last_error(), bytes)


def test_argument_errors_without_gpu(lib):
    """Validation happens before any launch: a null / malformed params block is rejected on a CPU-only box."""
    from omnimamba_b200 import _cabi
    p = _cabi.Ssu()
    rc = lib.omni_selective_state_update(ctypes.byref(p), None)
    assert rc == 1 and b"state" in lib.omni_last_error()
    p2 = _cabi.SsdFwd()
    assert lib.omni_ssd_chunk_scan_fwd(ctypes.byref(p2), None) != 0


def test_no_cpu_fallback():
    from omnimamba_b200.interface import selective_state_update
    st = torch.zeros(1, 2, 4, 16)
    with pytest.raises(RuntimeError, match="CUDA"):
        selective_state_update(st, torch.zeros(1, 2, 4), torch.zeros(1, 2, 4), torch.zeros(2, 4, 16),
                               torch.zeros(1, 1, 16), torch.zeros(1, 1, 16))


def test_product_never_imports_oracle():
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "omnimamba_b200")):
        for fn in fns:
            if fn.endswith(".py") and re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(dp, fn)).read(), re.M):
                bad.append(fn)
    assert not bad


def test_dropin_resolves_reference_imports():
    code = (
        "import omnimamba_b200 as o; o.install_dropin()\n"
        "from mamba_ssm.models.config_mamba import MambaConfig\n"
        "from mamba_ssm.modules.mamba_simple import Mamba\n"
        "from mamba_ssm.modules.mamba2 import Mamba2\n"
        "from mamba_ssm.modules.mha import MHA\n"
        "from mamba_ssm.modules.mlp import GatedMLP\n"
        "from mamba_ssm.utils.hf import load_config_hf, load_state_dict_hf\n"
        "from mamba_ssm.ops.triton.layer_norm import RMSNorm, layer_norm_fn, rms_norm_fn\n"
        "from mamba_ssm.ops.triton.ssd_combined import mamba_chunk_scan_combined, mamba_split_conv1d_scan_combined\n"
        "from mamba_ssm.ops.triton.selective_state_update import selective_state_update\n"
        "from mamba_ssm.ops.triton.layernorm_gated import RMSNorm as G, rmsnorm_fn\n"
        "from mamba_ssm.ops.selective_scan_interface import selective_scan_fn, selective_scan_ref, mamba_inner_fn\n"
        "from causal_conv1d import causal_conv1d_fn, causal_conv1d_update\n"
        "m = Mamba2(2048, layer_idx=0, device='meta')\n"
        "sd = {k: tuple(v.shape) for k, v in m.state_dict().items()}\n"
        "assert sd == {'dt_bias': (64,), 'A_log': (64,), 'D': (64,), 'in_proj.weight': (8512, 2048),"
        " 'conv1d.weight': (4352, 1, 4), 'conv1d.bias': (4352,), 'norm.weight': (4096,),"
        " 'out_proj.weight': (2048, 4096)}, sd\n"
        "print('ok')\n")
    out = subprocess.check_output([sys.executable, "-c", code], cwd=ROOT, text=True)
    assert out.strip().endswith("ok")
