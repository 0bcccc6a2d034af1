"""C-ABI checks that need no GPU: the library builds for sm_100a, loads, exports every symbol include/udt_api.h
declares, its descriptor struct matches the ctypes mirror, and compute entry points FAIL LOUDLY (no CPU fallback)
when no sm_100 device is present."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "udt_api.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(udt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(udt_lib):
    from udifftext_b200 import lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(udt_lib, name), f"{name} declared in include/udt_api.h but not exported by libudt_b200.so"
    assert set(declared) == set(lib.EXPORTED_SYMBOLS), set(declared) ^ set(lib.EXPORTED_SYMBOLS)


def test_descriptor_layout_and_constants(udt_lib):
    from udifftext_b200 import lib, pack
    assert udt_lib.udt_version() == 5
    assert udt_lib.udt_sizeof_igemm_desc() == ctypes.sizeof(lib.IGemmDesc)
    assert udt_lib.udt_geglu_tile() == pack.GEGLU_TILE


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_calls_fail_loudly_without_sm100(udt_lib):
    from udifftext_b200 import lib
    rc = udt_lib.udt_layernorm(None, None, 1, 8, None, None, 1e-5, None)
    assert rc == -3  # UDT_ERR_ARCH
    assert b"no CPU path" in udt_lib.udt_last_error() or b"sm_" in udt_lib.udt_last_error()
    with pytest.raises(lib.UdtError):
        lib.check(rc, "udt_layernorm")
    d = lib.IGemmDesc()
    assert udt_lib.udt_igemm(ctypes.byref(d), None) == -3


def test_host_objects_refuse_cpu_devices():
    from udifftext_b200 import api
    eng = api.build_engine("tiny")
    with pytest.raises(RuntimeError):
        eng.to("cpu")


def test_sass_contains_blackwell_tensor_and_tma_instructions(udt_lib):
    """the built library carries tcgen05 (UTC*MMA), TMEM loads (LDTM) and TMA (UTMALDG/UTMASTG) SASS"""
    import shutil
    import subprocess
    from udifftext_b200 import lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "STTM"):
        assert mnemonic in sass, mnemonic
    assert "HMMA." not in sass.replace("UTCHMMA", "")  # no legacy mma.sync path


def test_production_kernels_read_no_environment():
    """every experiment switch goes through udt_host::tune_int, which reads the environment only in tuning builds
    (-DUDT_TUNING); no other getenv exists in the kernel sources, and a default build does not define UDT_TUNING"""
    import re
    from udifftext_b200 import build
    csrc = os.path.join(os.path.dirname(build.__file__), "csrc")
    hits = []
    for name in sorted(os.listdir(csrc)):
        text = open(os.path.join(csrc, name)).read()
        for m in re.finditer(r"\bgetenv\s*\(", text):
            hits.append((name, text.count("\n", 0, m.start()) + 1))
    assert [h[0] for h in hits] == ["udt_host.h"], hits
    host_h = open(os.path.join(csrc, "udt_host.h")).read()
    block = host_h[host_h.index("#ifdef UDT_TUNING"):host_h.index("#else", host_h.index("#ifdef UDT_TUNING"))]
    assert "getenv" in block
    os.environ.pop("UDT_TRACE", None)
    assert "-DUDT_TUNING" not in build._flags()


def test_only_test_infrastructure_touches_the_oracle():
    """oracle/ is the checker, never the product: nothing in the package, the scripts or the C sources may import, link or
    execute it; outside tests/ only __graft_entry__.smoke() and bench.py's baseline legs do."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)|oracle[/.](restated|parseq_restated|ref_import|make_golden)", re.M)
    offenders = []
    for sub in ("udifftext_b200", "scripts", "include"):
        for dirpath, _, files in os.walk(os.path.join(root, sub)):
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".h", ".sh", ".cpp")):
                    continue
                path = os.path.join(dirpath, f)
                text = open(path, encoding="utf-8", errors="ignore").read()
                if re.search(r"^\s*(from\s+oracle\b|import\s+oracle\b)", text, re.M):
                    offenders.append(os.path.relpath(path, root))
    assert offenders == [], offenders
    # the two allowed importers outside tests/ use it only in the legs the contract names
    imp = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)", re.M)
    for name, allowed in (("__graft_entry__.py", {"smoke"}),
                          ("bench.py", {"cpu_reference_sample", "run_reference", "gpu_library_baseline"})):
        text = open(os.path.join(root, name)).read()
        hits = list(imp.finditer(text))
        assert hits, name
        for m in hits:
            fn = re.findall(r"^def\s+(\w+)", text[: m.start()], re.M)[-1]
            assert fn in allowed, (name, fn)
