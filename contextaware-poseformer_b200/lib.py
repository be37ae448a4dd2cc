"""ctypes binding of libcapf_b200.so (include/capf_b200.h).

The shared library is the product: there is deliberately *no* fallback.  If it is missing, or a
compute call is made without a B200, the caller gets an exception -- never a silent CPU path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcapf_b200.so")

ABI_VERSION = 13

# enums (mirror capf_b200.h)
F32, F16, BF16 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
IMPL_SIMT, IMPL_TCGEN05 = 0, 1
(OP_CONV2D, OP_FUSE_SUM, OP_MAXPOOL, OP_BILINEAR, OP_LAYERNORM, OP_ATTENTION, OP_REF_SAMPLE,
 OP_DEFORM_SAMPLE, OP_EMBED_COORD, OP_LEVELS_TO_JOINT, OP_CROP_NORMALIZE, OP_CAST, OP_PREPROCESS_U8, OP_BASICBLOCK, OP_WARP_AFFINE_U8, OP_POSE_ERRORS,
 OP_GEMM_F32, OP_COLSUM, OP_LAYERNORM_BWD, OP_GELU, OP_GELU_BWD, OP_ATTENTION_BWD, OP_DEFORM_BWD, OP_ROWS_AXPY, OP_JOINT_TO_LEVELS,
 OP_ADAMW, OP_EXPAND_REDUCE, OP_MLP) = range(1, 29)

DTYPE_CODE = {"f32": F32, "f16": F16, "bf16": BF16}


class CapfOp(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("dtype_in", C.c_int32),
        ("dtype_out", C.c_int32),
        ("reserved", C.c_int32),
        ("i", C.c_int32 * 24),
        ("f", C.c_float * 4),
        ("inp", C.c_void_p * 6),
        ("out", C.c_void_p * 4),
    ]


class CapfError(RuntimeError):
    pass


_lib = None


def load():
    """dlopen the library once; raises CapfError (never falls back) if it is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise CapfError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU/PyTorch fallback for the lifting path.")
    lib = C.CDLL(LIB_PATH)
    lib.capf_abi_version.restype = C.c_int
    lib.capf_last_error.restype = C.c_char_p
    lib.capf_device_info.argtypes = [C.c_int, C.POINTER(C.c_int64)]
    lib.capf_plan_create.argtypes = [C.POINTER(CapfOp), C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.capf_plan_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.capf_plan_num_launches.argtypes = [C.c_void_p]
    lib.capf_plan_op_kernel.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
    lib.capf_plan_destroy.argtypes = [C.c_void_p]
    lib.capf_op_run.argtypes = [C.POINTER(CapfOp), C.c_int, C.c_void_p]
    lib.capf_crop_normalize.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.capf_jpeg_available.restype = C.c_int
    lib.capf_jpeg_info.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.capf_jpeg_decode_batch.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int, C.c_void_p, C.c_int, C.c_int,
                                           C.POINTER(C.c_int), C.c_int, C.c_void_p]
    if lib.capf_abi_version() != ABI_VERSION:
        raise CapfError(f"libcapf_b200.so ABI {lib.capf_abi_version()} != binding {ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = load().capf_last_error().decode(errors="replace")
        raise CapfError(f"{what or 'libcapf_b200'} failed ({status}): {msg}")


def device_info(device: int = 0):
    out = (C.c_int64 * 4)()
    check(load().capf_device_info(device, out), "capf_device_info")
    return {"sms": out[0], "cc": (out[1], out[2]), "mem_bytes": out[3]}


def exported_symbols():
    """Names declared in include/capf_b200.h (used by the CPU-side ABI test)."""
    return ["capf_abi_version", "capf_last_error", "capf_device_info", "capf_plan_create", "capf_plan_run",
            "capf_plan_num_launches", "capf_plan_op_kernel", "capf_plan_destroy", "capf_op_run", "capf_crop_normalize",
            "capf_jpeg_available", "capf_jpeg_info", "capf_jpeg_decode_batch"]
