"""Run rllab_b200/csrc/experimental/umma_modes_probe.cu (operand source / major variants of one tcgen05 TF32 tile) on a
B200 and compare with NumPy.  Usage (under gpurun): timeout 120 python scripts/umma_modes_probe.py"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "rllab_b200", "csrc", "experimental", "umma_modes_probe.cu")
LIB = os.path.join(ROOT, "rllab_b200", "csrc", "experimental", "libumma_modes_probe.so")


def tf32(x):
    return (np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def main():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-Xcompiler",
                               "-fPIC", "-shared", "-o", LIB, SRC])
    lib = ctypes.CDLL(LIB)
    lib.umma_modes_probe.restype = ctypes.c_int
    lib.umma_modes_probe.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_void_p]
    rng = np.random.RandomState(0)
    A = np.tanh(rng.randn(128, 64)).astype(np.float32)
    B = (rng.randn(64, 64) / 8).astype(np.float32)
    dA, dB = torch.tensor(A, device="cuda"), torch.tensor(B, device="cuda")
    ref = tf32(A).astype(np.float64) @ tf32(B).astype(np.float64)
    names = {0: "SS A K-major, B K-major", 1: "SS A K-major, B MN-major", 2: "TS A TMEM, B MN-major",
             3: "SS A MN-major, B K-major", 4: "SS A MN-major, B MN-major"}
    bad = 0
    for mode in range(5):
        dD = torch.zeros((128, 64), dtype=torch.float32, device="cuda")
        dAb = torch.zeros((128, 64), dtype=torch.float32, device="cuda")
        st = torch.zeros(2, dtype=torch.int32, device="cuda")
        rc = lib.umma_modes_probe(dA.data_ptr(), dB.data_ptr(), dD.data_ptr(), dAb.data_ptr(), st.data_ptr(), mode,
                                  torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        D, Ab = dD.cpu().numpy(), dAb.cpu().numpy()
        s = st.cpu().numpy()
        err = np.abs(D - ref).max() / np.abs(ref).max()
        print("mode %d (%s): rc %d status %d tmem_base 0x%x  max|D| %.4f  sentinel-left %.2f  A-readback-ok %s  rel err %.3e"
              % (mode, names[mode], rc, s[0], int(s[1]) & 0xFFFFFFFF, np.abs(D).max(), float((D == 7.0).mean()),
                 bool((Ab == A).all()), err))
        if err > 1e-5:
            bad += 1
            # diagnose: does D match the product with a permuted / transposed operand?
            for nm, cand in (("A@B.T", tf32(A) @ tf32(B).T), ("first 8 k only", tf32(A)[:, :8] @ tf32(B)[:8]),
                             ("last 8 k only", tf32(A)[:, -8:] @ tf32(B)[-8:])):
                e2 = np.abs(D - cand).max() / np.abs(cand).max()
                print("     vs %s: %.3e" % (nm, e2))
            print("     D[0,:8] =", D[0, :8], " ref[0,:8] =", ref[0, :8].astype(np.float32))
    print("modes failing:", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
