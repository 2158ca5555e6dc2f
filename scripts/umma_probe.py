"""Run the stand-alone tcgen05 TF32 probe (rllab_b200/csrc/experimental/umma_tf32_probe.cu) on a B200 and compare with
NumPy: one TF32 pass against tf32(A) @ tf32(B), then the three-pass split (a_hi b_hi + a_lo b_hi + a_hi b_lo) against the
full-precision product.  Usage (under gpurun):  timeout 120 python scripts/umma_probe.py
Exit code 0 = the mechanisms the round-2 kernel design relies on (TMEM A operand, MN-major no-swizzle descriptor,
mma/commit/ld, float32-grade accuracy of the split) work as documented; 2 = the MMA never completed; 1 = wrong numbers."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "rllab_b200", "csrc", "experimental", "umma_tf32_probe.cu")
LIB = os.path.join(ROOT, "rllab_b200", "csrc", "experimental", "libumma_probe.so")


def tf32(x):
    return (np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def main():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-Xcompiler",
                               "-fPIC", "-shared", "-o", LIB, SRC])
    lib = ctypes.CDLL(LIB)
    lib.umma_probe.restype = ctypes.c_int
    lib.umma_probe.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_void_p]
    rng = np.random.RandomState(0)
    A = np.tanh(rng.randn(128, 64)).astype(np.float32)
    B = (rng.randn(64, 64) / 8).astype(np.float32)
    dA, dB = torch.tensor(A, device="cuda"), torch.tensor(B, device="cuda")
    full = A.astype(np.float64) @ B.astype(np.float64)
    ref1 = tf32(A).astype(np.float64) @ tf32(B).astype(np.float64)
    ok = True
    for split, ref, tol, what in ((0, ref1, 1e-5, "one TF32 pass vs the tf32-truncated product"),
                                 (1, full, 2e-6, "three-pass split vs the full-precision product")):
        dD = torch.zeros((128, 64), dtype=torch.float32, device="cuda")
        st = torch.zeros(1, dtype=torch.int32, device="cuda")
        rc = lib.umma_probe(dA.data_ptr(), dB.data_ptr(), dD.data_ptr(), st.data_ptr(), split,
                            torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        status = int(st.item())
        print("split=%d: launch rc %d, kernel status %d (1 = completed, -1 = mbarrier never signalled)" % (split, rc, status))
        if rc != 0 or status != 1:
            sys.exit(2)
        D = dD.cpu().numpy()
        err = np.abs(D - ref).max() / np.abs(ref).max()
        print("  %s: max|D - ref| / max|ref| = %.3e (tolerance %.0e); vs full precision %.3e"
              % (what, err, tol, np.abs(D - full).max() / np.abs(full).max()))
        if err > tol:
            bad = np.abs(D - ref) > 10 * tol * np.abs(ref).max()     # which rows / columns are wrong?
            print("  bad rows:", np.where(bad.any(1))[0][:16], "bad cols:", np.where(bad.any(0))[0][:16])
            ok = False
    if not ok:
        sys.exit(1)
    print("OK")


if __name__ == "__main__":
    main()
