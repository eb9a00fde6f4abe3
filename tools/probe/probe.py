"""ctypes loader of tools/probe/libdmip_probe.so — tcgen05 building-block self-tests and micro-benchmarks.
Test / tooling infrastructure; the product package never imports this."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libdmip_probe.so")
_lib = None


def build():
    r = subprocess.run(["bash", os.path.join(_HERE, "build.sh")], capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("building libdmip_probe.so failed:\n" + r.stdout + r.stderr)
    return _PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise RuntimeError(f"{_PATH} is missing — run tools/probe/build.sh")
        L = C.CDLL(_PATH)
        L.dmip_probe_last_error.restype = C.c_char_p
        L.dmip_debug_umma.restype = C.c_int
        L.dmip_debug_umma.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.dmip_debug_umma2.restype = C.c_int
        L.dmip_debug_umma2.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]
        L.dmip_debug_mma_bench.argtypes = [C.c_int32] * 5 + [C.c_void_p, C.c_void_p]
        L.dmip_debug_mma_bench2.argtypes = [C.c_int32] * 7 + [C.c_void_p, C.c_void_p, C.c_void_p]
        L.dmip_debug_prim_bench.argtypes = [C.c_int32, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError(f"probe error {rc}: {lib().dmip_probe_last_error().decode()}")


def umma2(a, b, a_mn, b_mn, split, img, fld, iters=1, want_cycles=False):
    """D = A B^T (A: (128,k), B: (n,k) fp32 CUDA tensors) through the probe kernel; img / fld: 4 byte strides each
    (a_lbo, a_sbo, b_lbo, b_sbo) for the MN-major images and for the descriptor fields."""
    import torch
    L = lib()
    n, k = b.shape[0], a.shape[1]
    d = torch.full((128, n), float("nan"), device=a.device)
    cyc = torch.zeros(2, dtype=torch.int64, device=a.device)
    check(L.dmip_debug_umma2((C.c_int32 * 2)(a_mn, b_mn), n, k, int(split), iters, (C.c_uint32 * 4)(*img),
                             (C.c_uint32 * 4)(*fld), a.data_ptr(), b.data_ptr(), d.data_ptr(), cyc.data_ptr(),
                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return (d, cyc.cpu().tolist()) if want_cycles else d
