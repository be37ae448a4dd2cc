"""Helpers for the GPU tests: run a single op through capf_op_run."""
import ctypes

import torch

from capf_b200 import lib

DT = {torch.float32: lib.F32, torch.float16: lib.F16, torch.bfloat16: lib.BF16}


def run_op(kind, dtype_in, dtype_out, ints, floats, ins, outs):
    op = lib.CapfOp()
    op.kind = kind
    op.dtype_in = DT[dtype_in]
    op.dtype_out = DT[dtype_out]
    for n, v in enumerate(ints):
        op.i[n] = int(v)
    for n, v in enumerate(floats):
        op.f[n] = float(v)
    for n, t in enumerate(ins):
        op.inp[n] = None if t is None else t.data_ptr()
    for n, t in enumerate(outs):
        op.out[n] = None if t is None else t.data_ptr()
    st = torch.cuda.current_stream().cuda_stream
    lib.check(lib.load().capf_op_run(ctypes.byref(op), torch.cuda.current_device(), st), f"op {kind}")
    torch.cuda.synchronize()
