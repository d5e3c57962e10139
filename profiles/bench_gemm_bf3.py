"""Device time of the TMA-fed 3xBF16 GEMM alone (the three modes of one hidden layer of BASELINE config 5:
65536 columns x 512 x 512).  The library prints the CUDA-event time per launch when HVI_GEMM_TIME is set.
Usage (GPU box): HVI_GEMM_TIME=1 python profiles/bench_gemm_bf3.py"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from helpers import hvi_b200  # noqa: E402

os.environ.setdefault("HVI_GEMM_TIME", "1")
lib = hvi_b200._abi.load_library()
p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
rows, W = int(os.environ.get("ROWS", 65536)), int(os.environ.get("WIDTH", 512))
rng = np.random.default_rng(0)
A = rng.standard_normal((rows, W)).astype(np.float32)
Wt = (rng.standard_normal((W, W)) / np.sqrt(W)).astype(np.float32)
b = rng.standard_normal(W).astype(np.float32)
H = np.tanh(rng.standard_normal((rows, W))).astype(np.float32)
out = np.zeros((rows, W), dtype=np.float32)
assert lib.hvi_gemm_bf3(0, 0, rows, W, W, p(A), p(Wt), p(b), 1, None, p(out)) == 0
assert lib.hvi_gemm_bf3(0, 1, rows, W, W, p(A), p(Wt), None, 1, p(H), p(out)) == 0
G = np.zeros((W, W), dtype=np.float32)
assert lib.hvi_gemm_bf3(0, 2, W, W, rows, p(A), p(H), None, 0, None, p(G)) == 0
