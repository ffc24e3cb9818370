"""Build + ctypes binding of ``libcova_b200.so`` (C ABI declared in ``include/cova_b200.h``).

The library is built IN-TREE with nvcc for sm_100a only (it travels to the GPU box with the repo snapshot).
There is no fallback: if the library is missing or fails to load, every op raises.
"""
import ctypes
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libcova_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "cova_b200.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--cudart", "static",
]

F32, BF16, BF16X2, U8, F16, F16X2 = 0, 1, 2, 3, 4, 5
ENGINE_SIMT, ENGINE_TCGEN05, ENGINE_TCGEN05_F16X2 = 0, 1, 2


def sources():
    return sorted(glob.glob(os.path.join(_CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(_CSRC, "*.cuh")) + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every ``csrc/*.cu`` into ``libcova_b200.so`` (nvcc cross-compiles without a GPU).  One nvcc per source
    (in parallel, objects under ``csrc/build/``, recompiled only when the source or a header changed), then one link."""
    if not force and not _stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(_CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = glob.glob(os.path.join(_CSRC, "*.cuh")) + [HEADER]
    hdr_t = max(os.path.getmtime(h) for h in hdrs)
    cflags = [f for f in NVCC_FLAGS if f not in ("-shared", "--cudart", "static")]
    logs = []

    def one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj
        cmd = [nvcc] + cflags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        logs.append(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(one, sources()))
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print("\n".join(logs))
    return LIB_PATH


_c = ctypes
_P, _I, _L, _F, _D = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float, _c.c_double

# name -> (restype, argtypes); must list every symbol include/cova_b200.h declares
SIGNATURES = {
    "cova_abi_version": (_I, []),
    "cova_last_error": (_c.c_char_p, []),
    "cova_device_info": (_I, [_c.POINTER(_I), _c.POINTER(_I)]),
    "cova_set_knob": (_I, [_I, _I]),
    "cova_debug_buffer": (_I, [_P, _L]),
    "cova_stem_fwd": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P, _I, _P]),
    "cova_conv3x3_bn_act_fwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _P]),
    "cova_conv1x1_bn_act_fwd": (_I, [_P, _P, _L, _I, _I, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P]),
    "cova_pack_conv_weight": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "cova_pack_stem_weight": (_I, [_P, _P, _P]),
    "cova_stem_conv_raw_fwd": (_I, [_P, _I, _I, _I, _I, _P, _I, _P, _P]),
    "cova_pack_stem_weight_f16": (_I, [_P, _P, _P]),
    "cova_pack_stem_weight_f16x2": (_I, [_P, _P, _P]),
    "cova_pack_conv_weight_f16": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "cova_pack_conv_weight_f16x2": (_I, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "cova_roi_fwd": (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _F, _I, _I, _P, _L, _P, _P]),
    "cova_roi_pool_bwd": (_I, [_P, _L, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "cova_roi_align_bwd": (_I, [_P, _L, _P, _I, _I, _I, _I, _F, _I, _I, _I, _I, _P, _P]),
    "cova_bbox_enc_fwd": (_I, [_P, _I, _P, _P, _P, _P, _I, _P, _L, _P]),
    "cova_affine_cols_fwd": (_I, [_P, _I, _I, _L, _P, _P, _P, _L, _P]),
    "cova_linear_fwd": (_I, [_P, _L, _I, _I, _P, _I, _P, _P, _P, _P, _L, _I, _I, _P, _P, _L, _I, _P]),
    "cova_pack_linear_weight": (_I, [_P, _I, _I, _P, _P]),
    "cova_gat_fwd": (_I, [_P, _L, _P, _P, _L, _F, _F, _P, _I, _I, _I, _P, _L, _P, _P]),
    "cova_gat_multihead_fwd": (_I, [_P, _L, _I, _I, _c.POINTER(_F), _F, _P, _I, _I, _P, _L, _P, _P]),
    "cova_gat_bwd": (_I, [_P, _L, _P, _L, _P, _P, _L, _F, _F, _P, _P, _I, _I, _I, _P, _L, _P, _P, _L, _P, _P]),
    "cova_ce_sum_fwd_bwd": (_I, [_P, _L, _P, _I, _I, _L, _P, _P, _P, _L, _P, _P]),
    "cova_adam_step": (_I, [_P, _P, _P, _P, _L, _D, _D, _D, _D, _D, _I, _D, _P]),
    "cova_topk_hits": (_I, [_P, _L, _P, _P, _I, _I, _I, _P, _P]),
    "cova_build_batch": (_I, [_P, _I, _I, _I, _P, _P, _P, _P]),
    "cova_bn_train_stats": (_I, [_P, _L, _I, _P, _P]),
    "cova_bn_train_finalize": (_I, [_P, _L, _I, _F, _F, _P, _P, _P, _P, _P]),
    "cova_bn_act_fwd": (_I, [_P, _L, _I, _P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _P, _P]),
    "cova_split_planes": (_I, [_P, _L, _P, _P, _I, _P]),
    "cova_split_planes_scaled": (_I, [_P, _L, _P, _P, _I, _I, _P, _P, _P]),
    "cova_conv3x3_wgrad": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "cova_conv1x1_raw_fwd": (_I, [_P, _P, _I, _L, _I, _I, _P, _P, _P, _P, _P]),
    "cova_stem_wgrad": (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P]),
    "cova_conv1x1_wgrad": (_I, [_P, _P, _P, _P, _L, _I, _I, _I, _P, _P, _P, _P]),
    "cova_bn_act_bwd": (_I, [_P, _P, _P, _L, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P]),
    "cova_bn_act_bwd_planes": (_I, [_P, _P, _P, _L, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "cova_bn_train_stats_t": (_I, [_P, _I, _L, _I, _P, _P]),
    "cova_bn_act_fwd_t": (_I, [_P, _I, _L, _I, _P, _P, _P, _P, _P, _I, _P, _I, _P, _P]),
    "cova_bn_act_bwd_t": (_I, [_P, _I, _P, _P, _I, _L, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P]),
    "cova_maxpool3x3s2_fwd_t": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "cova_maxpool3x3s2_bwd_t": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "cova_stem_conv_raw_fwd_bf16": (_I, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "cova_conv3x3_bn_act_stats_fwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _P, _P]),
    "cova_conv1x1_raw_stats_fwd": (_I, [_P, _P, _I, _L, _I, _I, _P, _P, _P, _P, _P, _P]),
    "cova_stem_conv_raw_stats_fwd": (_I, [_P, _I, _I, _I, _I, _P, _I, _P, _P, _P]),
    "cova_bn_relu_pool_fwd_t": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _P]),
    "cova_bn_relu_pool_bwd_t": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "cova_conv1x1_raw_res_fwd": (_I, [_P, _L, _I, _I, _P, _P, _P, _P, _P, _P]),
    "cova_conv3x3_scale_res_f32_fwd": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "cova_conv1x1_raw_res_f32_fwd": (_I, [_P, _P, _I, _L, _I, _I, _P, _P, _P, _P, _P, _P]),
    "cova_maxpool3x3s2_fwd": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P]),
    "cova_maxpool3x3s2_bwd": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "cova_bn_relu_pool_fwd": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P]),
    "cova_bn_relu_pool_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
}

_lib = None


def lib():
    """The loaded library.  Raises (never falls back) if it is absent or its ABI version is wrong."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(cova_b200 has no CPU or PyTorch fallback for its hot path)")
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        if h.cova_abi_version() != 1:
            raise RuntimeError("libcova_b200.so ABI version mismatch")
        _lib = h
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {lib().cova_last_error().decode()}")
