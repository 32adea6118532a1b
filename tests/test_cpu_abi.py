"""CPU-only checks: the C-ABI library builds/loads and exports every symbol include/eunet.h declares (no
compute calls without a GPU), the drop-in keeps the reference's Python surface, and the product package
never touches the oracle."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from enhanced_unet_b200 import build, lib
    build.build_library()
    return lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "eunet.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eunet_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    from enhanced_unet_b200 import lib
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in include/eunet.h but not exported"
    typed = set(lib.SIGNATURES) | {"eunet_last_error"}
    assert set(names) == typed, set(names) ^ typed
    assert built_lib.eunet_abi_version() == lib.ABI_VERSION
    assert isinstance(lib.last_error(), str)


def test_library_is_sm100a_tensor_core_code(built_lib):
    """The shipped binary must contain tcgen05 / TMA machine code (UTCHMMA, UTMALDG loads, UTMASTG stores, LDTM) and
    the packed-fp32 epilogue arithmetic (FADD2 / FFMA2) for sm_100a."""
    import shutil
    import subprocess
    from enhanced_unet_b200 import lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "FADD2", "FFMA2"):
        assert mnemonic in out, mnemonic


def test_dropin_surface_matches_reference_contract():
    import inspect
    import oracle
    from enhanced_unet_b200 import models
    torch.manual_seed(0)
    m = models.EnhancedUNet(3)
    sd = m.state_dict()
    ref = oracle.make_state_dict(0)
    assert list(sd.keys()) == list(ref.keys()) and len(sd) == 109
    for k in sd:
        assert sd[k].shape == ref[k].shape and sd[k].dtype == ref[k].dtype, k
    assert sum(p.numel() for p in m.parameters()) == 7_790_790
    assert m.num_classes == 3 and m.get_aux_outputs() is None
    sig = inspect.signature(models.get_model)
    assert list(sig.parameters)[:6] == ["model_name", "num_classes", "device", "train_mode", "data_dir", "max_size"]
    with pytest.raises(ValueError, match="Unknown model"):
        models.get_model("does_not_exist")
    with pytest.raises(RuntimeError):          # parameters on CPU: there is no CPU fallback
        m(torch.rand(1, 3, 32, 32))


@pytest.mark.ref
def test_same_seed_gives_reference_initialisation():
    """Same construction order as the reference => identical default init for a given torch seed."""
    from oracle import ref_import
    from enhanced_unet_b200 import models
    ref_models, _, _ = ref_import.load()
    torch.manual_seed(0)
    r = ref_models.EnhancedUNet(3)
    torch.manual_seed(0)
    m = models.EnhancedUNet(3)
    rs, ms = r.state_dict(), m.state_dict()
    assert list(rs.keys()) == list(ms.keys())
    for k in rs:
        assert torch.equal(rs[k], ms[k]), k
    m.load_state_dict(rs, strict=True)
    r.load_state_dict(ms, strict=True)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "enhanced_unet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "/root/reference" not in text, f
