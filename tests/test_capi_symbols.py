"""The C-ABI library builds, loads, and exports every symbol include/doubletake_b200.h declares (CPU only)."""
import ctypes
import os
import re

from doubletake_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "doubletake_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(dtb200_[a-z0-9_]+)\s*\(", text))


def test_library_builds_and_loads():
    path = build.build()
    assert os.path.exists(path)
    lib = _lib.lib()
    assert lib.dtb200_abi_version() == 1


def test_every_declared_symbol_is_exported_and_bound():
    declared = header_symbols()
    assert declared, "no symbols parsed from the header"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    handle = ctypes.CDLL(build.build())
    for name in declared:
        assert hasattr(handle, name), name


def _struct_fields(name):
    text = open(os.path.join(ROOT, "include", "doubletake_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    structs = dict((n, b) for b, n in re.findall(r"typedef struct \{((?:(?!typedef).)*?)\} (\w+);", text, flags=re.S))
    names = []
    for decl in structs[name].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(float|int32_t|uint8_t|uint64_t|void)\s*\*?", "", decl).strip()
        for part in decl.split(","):
            names.append(re.sub(r"\[.*\]", "", part.replace("*", "")).strip())
    return names


def test_struct_layout_matches_header_field_order():
    assert _struct_fields("dtb200_conv_params") == [f[0] for f in _lib.ConvParams._fields_]
    assert _struct_fields("dtb200_cost_volume_params") == [f[0] for f in _lib.CostVolumeParams._fields_]


def test_errors_are_reported_without_a_gpu():
    lib = _lib.lib()
    rc = lib.dtb200_conv2d(None, None)
    assert rc == -1 and b"null params" in lib.dtb200_last_error()
    rc = lib.dtb200_cost_volume(None, None)
    assert rc == -1
