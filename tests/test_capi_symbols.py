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
    structs = dict((n, b) for b, n in re.findall(r"typedef struct(?: \w+)? \{((?:(?!typedef).)*?)\} (\w+);", text, flags=re.S))
    names = []
    for decl in structs[name].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(float|int32_t|uint8_t|uint64_t|void|dtb200_tsdf_frame)\s*\*?", "", decl).strip()
        for part in decl.split(","):
            names.append(re.sub(r"\[.*\]", "", part.replace("*", "")).strip())
    return names


def test_struct_layout_matches_header_field_order():
    assert _struct_fields("dtb200_conv_params") == [f[0] for f in _lib.ConvParams._fields_]
    assert _struct_fields("dtb200_cost_volume_params") == [f[0] for f in _lib.CostVolumeParams._fields_]
    assert _struct_fields("dtb200_tsdf_frame") == [f[0] for f in _lib.TsdfFrame._fields_]
    assert _struct_fields("dtb200_tsdf_integrate_params") == [f[0] for f in _lib.TsdfIntegrateParams._fields_]
    assert _struct_fields("dtb200_tsdf_raycast_params") == [f[0] for f in _lib.TsdfRaycastParams._fields_]
    assert _struct_fields("dtb200_instance_norm_params") == [f[0] for f in _lib.InstanceNormParams._fields_]


def test_struct_sizes_and_offsets_match_the_c_compiler(tmp_path):
    """sizeof / offsetof of every ABI struct as gcc lays them out == the ctypes mirrors (catches a wrong field TYPE, which
    the name comparison above cannot)."""
    import subprocess

    pairs = [("dtb200_conv_params", _lib.ConvParams), ("dtb200_cost_volume_params", _lib.CostVolumeParams),
             ("dtb200_tsdf_frame", _lib.TsdfFrame), ("dtb200_tsdf_integrate_params", _lib.TsdfIntegrateParams),
             ("dtb200_tsdf_raycast_params", _lib.TsdfRaycastParams), ("dtb200_instance_norm_params", _lib.InstanceNormParams)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "doubletake_b200.h")}"',
             "int main(void) {"]
    for cname, cls in pairs:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in pairs:
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)


def test_errors_are_reported_without_a_gpu():
    lib = _lib.lib()
    rc = lib.dtb200_conv2d(None, None)
    assert rc == -1 and b"null params" in lib.dtb200_last_error()
    rc = lib.dtb200_cost_volume(None, None)
    assert rc == -1


def test_tsdf_argument_validation_without_a_gpu():
    """Every rejection happens before the first CUDA call, so it is testable here."""
    lib = _lib.lib()
    assert lib.dtb200_tsdf_integrate(None, None) == -1
    p = _lib.TsdfIntegrateParams()
    p.values, p.weights = 16, 32  # never dereferenced: validation fails first
    p.num_frames = 9
    assert lib.dtb200_tsdf_integrate(ctypes.byref(p), None) == -1 and b"num_frames" in lib.dtb200_last_error()
    p.num_frames = 1
    p.dims = (ctypes.c_int32 * 3)(16, 16, 12)
    assert lib.dtb200_tsdf_integrate(ctypes.byref(p), None) == -1 and b"multiples of 8" in lib.dtb200_last_error()
    p.dims = (ctypes.c_int32 * 3)(16, 16, 16)
    p.img_h, p.img_w = 4, 4
    p.semantics = 7
    assert lib.dtb200_tsdf_integrate(ctypes.byref(p), None) == -1 and b"semantics" in lib.dtb200_last_error()
    p.semantics = 0
    assert lib.dtb200_tsdf_integrate(ctypes.byref(p), None) == -1 and b"no depth map" in lib.dtb200_last_error()
    p.frames[0].depth = 64
    p.vox_begin, p.vox_end = (ctypes.c_int32 * 3)(0, 0, 4), (ctypes.c_int32 * 3)(16, 16, 16)
    assert lib.dtb200_tsdf_integrate(ctypes.byref(p), None) == -1 and b"vox_begin" in lib.dtb200_last_error()
    p.vox_begin, p.vox_end = (ctypes.c_int32 * 3)(8, 8, 8), (ctypes.c_int32 * 3)(8, 8, 8)
    assert lib.dtb200_tsdf_integrate(ctypes.byref(p), None) == 0  # empty box: nothing to launch
    dims, origin = (ctypes.c_int32 * 3)(16, 16, 16), (ctypes.c_float * 3)(0, 0, 0)
    assert lib.dtb200_tsdf_sample(16, dims, origin, 0.04, 16, 16, 10, 5, None) == -1 and b"mode" in lib.dtb200_last_error()
    assert lib.dtb200_tsdf_sample(None, dims, origin, 0.04, 16, 16, 10, 0, None) == -1
    assert lib.dtb200_tsdf_raycast(None, None) == -1
    r = _lib.TsdfRaycastParams()
    r.values = r.weights = r.invK = r.world_T_cam = r.depth_hint = r.hint_mask = r.sampled_weights = 64
    r.batch, r.height, r.width = 1, 4, 4
    r.dims = (ctypes.c_int32 * 3)(16, 16, 16)
    r.voxel_size, r.z_near, r.z_far, r.max_steps = 0.04, 1.0, 0.5, 10
    assert lib.dtb200_tsdf_raycast(ctypes.byref(r), None) == -1 and b"depth range" in lib.dtb200_last_error()
