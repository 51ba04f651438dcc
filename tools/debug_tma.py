import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import b200sr
from b200sr import _lib
lib = _lib.load()
x = torch.arange(256 * 128, dtype=torch.float32, device="cuda").remainder(997).to(torch.bfloat16).view(256, 128)
dims = (ctypes.c_uint64 * 2)(128, 256)
strides = (ctypes.c_uint64 * 1)(256)
box = (ctypes.c_uint32 * 2)(64, 32)
for swz in (0, 1):
    out = (ctypes.c_uint8 * 128)()
    lib.b200sr_debug_encode.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    rc = lib.b200sr_debug_encode(x.data_ptr(), 2, dims, strides, box, swz, out)
    mine = bytes(out)
    print("swz", swz, "encode rc", rc, _lib.last_error() if rc else "", mine[:64].hex())
    try:
        from cuda.bindings import driver as drv
        err, tm = drv.cuTensorMapEncodeTiled(drv.CUtensorMapDataType.CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x.data_ptr(),
                                             [drv.cuuint64_t(128), drv.cuuint64_t(256)], [drv.cuuint64_t(256)],
                                             [drv.cuuint32_t(64), drv.cuuint32_t(32)], [drv.cuuint32_t(1), drv.cuuint32_t(1)],
                                             drv.CUtensorMapInterleave.CU_TENSOR_MAP_INTERLEAVE_NONE,
                                             drv.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_128B if swz else drv.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_NONE,
                                             drv.CUtensorMapL2promotion.CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                             drv.CUtensorMapFloatOOBfill.CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
        theirs = bytes(np.array(tm.opaque, dtype=np.uint64).tobytes()) if hasattr(tm, "opaque") else None
        print("  cuda-python err", err, "equal", theirs == mine if theirs else "n/a")
        if theirs and theirs != mine:
            print("  theirs", theirs[:64].hex())
    except Exception as e:
        print("  cuda-python compare failed:", repr(e)[:200])
    dst = torch.zeros(32, 64, dtype=torch.bfloat16, device="cuda")
    lib.b200sr_debug_tma.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    rc = lib.b200sr_debug_tma(out, dst.data_ptr(), 32, 64, 32, None)
    try:
        torch.cuda.synchronize()
        ref = x[32:64, 64:128]
        print("  tma rc", rc, "match(no-swizzle view)", bool(torch.equal(dst, ref)), dst[0, :8].tolist(), ref[0, :8].tolist(), dst[1, :8].tolist())
    except Exception as e:
        print("  tma FAILED:", repr(e)[:300])
        break
