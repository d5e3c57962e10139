"""The TMA-fed tcgen05 GEMM of the wide applicator (bf16 hi/lo operand planes, three MMAs per K step,
persistent CTAs with two TMEM accumulators) against float64: layer forward, data gradient
(⊙ act'(H)) and the split-K weight gradient that reads the same planes as MN-major operands.
Tolerance 3e-5 relative to max|C|: operands carry 16 mantissa bits (2^-17 relative each), the lo·lo
product is dropped, accumulation is the tensor core's fp32."""
import ctypes as C

import numpy as np
import pytest

from helpers import hvi_b200

pytestmark = pytest.mark.gpu

p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None


def _act(z, act):
    return np.tanh(z) if act == 1 else 1 / (1 + np.exp(-z)) if act == 2 else z


@pytest.mark.parametrize("M,N,K,act", [(128, 128, 64, 0), (128, 256, 512, 1), (300, 256, 64, 1), (1000, 512, 512, 1),
                                       (77, 128, 96, 2), (40000, 512, 512, 1), (4096, 384, 128, 0)])
def test_forward(M, N, K, act):
    lib = hvi_b200._abi.load_library()
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    out = np.zeros((M, N), dtype=np.float32)
    rc = lib.hvi_gemm_bf3(0, 0, M, N, K, p(A), p(W), p(b), act, None, p(out))
    assert rc == 0, lib.hvi_last_error(None)
    ref = _act(A.astype(np.float64) @ W.astype(np.float64).T + b, act)
    err = np.max(np.abs(out - ref)) / np.max(np.abs(ref))
    assert err < 3e-5, err


@pytest.mark.parametrize("M,N,K,act", [(256, 128, 128, 1), (1000, 512, 512, 1), (333, 256, 512, 2), (515, 512, 256, 0)])
def test_data_gradient(M, N, K, act):
    lib = hvi_b200._abi.load_library()
    rng = np.random.default_rng(7 * M + N + K)
    D = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    H = np.tanh(rng.standard_normal((M, N))).astype(np.float32) if act == 1 else rng.uniform(0.05, 0.95, (M, N)).astype(np.float32)
    out = np.zeros((M, N), dtype=np.float32)
    rc = lib.hvi_gemm_bf3(0, 1, M, N, K, p(D), p(W), None, act, p(H), p(out))
    assert rc == 0, lib.hvi_last_error(None)
    h = H.astype(np.float64)
    dv = 1 - h * h if act == 1 else h * (1 - h) if act == 2 else np.ones_like(h)
    ref = (D.astype(np.float64) @ W.astype(np.float64).T) * dv
    err = np.max(np.abs(out - ref)) / np.max(np.abs(ref))
    assert err < 3e-5, err


@pytest.mark.parametrize("Md,Nd,rows", [(128, 128, 64), (512, 512, 65536), (256, 384, 5001), (512, 256, 34464), (128, 512, 96),
                                        (96, 128, 777)])
def test_weight_gradient_split_k(Md, Nd, rows):
    """G[Md x Nd] = Xᵀ Y, X (rows x Md), Y (rows x Nd): contraction over the rows cut into splits (ragged last
    split, rows not a multiple of 64, Md not a multiple of the 128-row tile)."""
    lib = hvi_b200._abi.load_library()
    rng = np.random.default_rng(Md + Nd + rows)
    X = rng.standard_normal((rows, Md)).astype(np.float32)
    Y = rng.standard_normal((rows, Nd)).astype(np.float32)
    out = np.zeros((Md, Nd), dtype=np.float32)
    rc = lib.hvi_gemm_bf3(0, 2, Md, Nd, rows, p(X), p(Y), None, 0, None, p(out))
    assert rc == 0, lib.hvi_last_error(None)
    ref = X.astype(np.float64).T @ Y.astype(np.float64)
    err = np.max(np.abs(out - ref)) / np.max(np.abs(ref))
    assert err < 3e-5, err
