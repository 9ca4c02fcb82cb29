"""CPU-side checks: the C-ABI library loads and exports what include/prego_b200.h declares, the host
mirror of the reference interface behaves like the reference, and the product path fails loudly
without a GPU (no fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT, seeded_weights_checked

import prego_b200
from prego_b200 import _lib, synthetic
from prego_b200.registry import META_ARCHITECTURES, Registry, build_model


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def test_header_symbols_all_exported(lib):
    header = open(os.path.join(ROOT, "include", "prego_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(prego_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 11
    raw = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/prego_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes prototypes out of sync with the header"
    assert lib.prego_abi_version() == 4


def test_every_entry_point_is_mapped_to_the_reference_in_integration_md():
    """The drop-in boundary is documented function by function: every symbol include/prego_b200.h declares must appear in
    INTEGRATION.md's table (entry point -> the reference file:line it replaces)."""
    header = open(os.path.join(ROOT, "include", "prego_b200.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    declared = sorted(set(re.findall(r"\b(prego_[a-z0-9_]+)\s*\(", header)))
    missing = [n for n in declared if f"`{n}`" not in doc]
    assert not missing, f"declared in the header but absent from INTEGRATION.md: {missing}"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu(lib):
    dims = _lib.Dims(2048, 2048, 2048, 1024, 86)
    h = C.c_void_p()
    rc = lib.prego_model_create(C.byref(dims), 0, C.byref(h))
    assert rc != 0 and h.value is None
    assert len(lib.prego_last_error()) > 0


def test_bad_dims_rejected(lib):
    h = C.c_void_p()
    assert lib.prego_model_create(C.byref(_lib.Dims(0, 0, 2048, 1024, 86)), 0, C.byref(h)) == 1
    assert b"rgb" in lib.prego_last_error()
    assert lib.prego_model_create(C.byref(_lib.Dims(2048, 2048, 1000, 1024, 86)), 0, C.byref(h)) == 1


def test_registry_contract():
    r = Registry()

    @r.register("A")
    class A:  # noqa
        pass

    r.register("B", int)
    assert r["A"] is A and r["B"] is int
    with pytest.raises(AssertionError):
        r.register("A", float)
    assert "MiniROAD" in META_ARCHITECTURES and "OAD" in prego_b200.EVAL
    m = build_model(dict(synthetic.EPIC_TENT_O), None)
    assert isinstance(m, prego_b200.MROAD)


def test_state_dict_contract(golden_meta):
    m = seeded_weights_checked(golden_meta, "asm_b2_t160")  # hashes equal the reference's seeded init
    sd = m.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    assert shapes == {
        "gru.weight_ih_l0": (3072, 2048), "gru.weight_hh_l0": (3072, 1024), "gru.bias_ih_l0": (3072,),
        "gru.bias_hh_l0": (3072,), "layer1.0.weight": (2048, 4096), "layer1.0.bias": (2048,),
        "layer1.1.weight": (2048,), "layer1.1.bias": (2048,), "f_classification.0.weight": (86, 1024),
        "f_classification.0.bias": (86,)}
    assert sum(v.numel() for v in sd.values()) == 17926230
    m2 = build_model(dict(synthetic.ASSEMBLY101_O), None)
    m2.load_state_dict(sd, strict=True)
    assert "h0" not in sd and tuple(m.h0.shape) == (1, 1, 1024)


def test_cpu_tensors_and_train_mode_rejected():
    m = build_model(dict(synthetic.EPIC_TENT_O), None).eval()
    x = torch.zeros(1, 4, 2048)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, x)
    m.train()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, x)


def test_aggregate_needs_cuda():
    if torch.cuda.is_available():
        pytest.skip("no-GPU failure mode")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        prego_b200.aggregate_labels([[1, 2]], [[1, 2]])


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under prego_b200/ may reference it."""
    pkg = os.path.join(ROOT, "prego_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports oracle"
                assert "/root/reference" not in src, f"{f} reads the reference at run time"


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """Every ctypes mirror in prego_b200/_lib.py has the size and field offsets of the C struct it stands for
    (checked by compiling a probe against include/prego_b200.h with the host C compiler)."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no host C compiler")
    pairs = {"prego_dims_t": _lib.Dims, "prego_weights_t": _lib.Weights, "prego_forward_args_t": _lib.ForwardArgs,
             "prego_anticipation_args_t": _lib.AnticipationArgs, "prego_grads_t": _lib.Grads, "prego_train_args_t": _lib.TrainArgs, "prego_adamw_args_t": _lib.AdamWArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "prego_b200.h"', 'int main(void) {']
    for cname, ct in pairs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.check_call([cc, "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split("\n")
    seen = 0
    for ln in out:
        if not ln:
            continue
        cname, what, val = ln.split()
        ct = pairs[cname]
        if what == "size":
            assert C.sizeof(ct) == int(val), f"{cname}: ctypes size {C.sizeof(ct)} vs C {val}"
        else:
            assert getattr(ct, what).offset == int(val), f"{cname}.{what}: ctypes offset {getattr(ct, what).offset} vs C {val}"
        seen += 1
    assert seen > 60
