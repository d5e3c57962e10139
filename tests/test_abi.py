"""The C-ABI boundary without a GPU: the library loads, exports every symbol that
include/hvi_b200.h declares, the ctypes mirror of hvi_desc has the C layout, the pure integer
index maps answer like the reference's, and the compute entry points fail loudly (no CPU
fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from helpers import ROOT, hvi_b200

A = hvi_b200._abi
HEADER = os.path.join(ROOT, "include", "hvi_b200.h")


def declared_symbols():
    src = open(HEADER, encoding="utf-8").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hvi_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = A.load_library()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(s[0] for s in A.SYMBOLS) == names


def test_desc_layout_matches_c():
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write('#include <stdio.h>\n#include <stddef.h>\n#include "hvi_b200.h"\n'
                             'int main(){printf("%zu %zu %zu %zu %zu %zu", sizeof(hvi_desc), offsetof(hvi_desc, scale_a),'
                             'offsetof(hvi_desc, prior_M), offsetof(hvi_desc, slot_fix), offsetof(hvi_desc, count_globals), offsetof(hvi_desc, n_site_total));return 0;}')
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = list(map(int, subprocess.check_output([exe]).split()))
    D = A.hvi_desc
    assert out == [C.sizeof(D), D.scale_a.offset, D.prior_M.offset, D.slot_fix.offset, D.count_globals.offset,
                   D.n_site_total.offset]


def test_index_maps_known_answers():
    # test/test_cholesky_structure.jl:109-119
    assert hvi_b200.utri2vec_pos(1, 1) == 1 and hvi_b200.utri2vec_pos(1, 2) == 2
    assert hvi_b200.utri2vec_pos(2, 2) == 3 and hvi_b200.utri2vec_pos(1, 3) == 4
    assert hvi_b200.utri2vec_pos(1, 4) == 7 and hvi_b200.utri2vec_pos(5, 5) == 15
    with pytest.raises(AssertionError):
        hvi_b200.utri2vec_pos(2, 1)
    for s in range(1, 5000):
        r, c = hvi_b200.vec2utri_pos(s)
        assert 1 <= r <= c and hvi_b200.utri2vec_pos(r, c) == s
    # :66-77
    for ends, n in (([], 0), ([1], 0), ([2], 1), ([3], 3), ([4], 6), ([1, 4], 3), ([2, 4], 2), ([3, 4], 3), ([2, 5], 4)):
        assert hvi_b200.get_cor_count(ends) == n


def test_plan_create_validates_and_fails_loudly_without_gpu():
    lib = A.load_library()
    h = C.c_void_p()
    d = hvi_b200.DoubleMMCase().to_desc()
    d.struct_size = 12
    assert lib.hvi_plan_create(C.byref(d), C.byref(h)) == A.HVI_EINVAL
    assert b"struct_size" in lib.hvi_last_error(None)
    d = hvi_b200.DoubleMMCase().to_desc()
    d.n_M = 0
    assert lib.hvi_plan_create(C.byref(d), C.byref(h)) == A.HVI_EINVAL
    d = hvi_b200.DoubleMMCase().to_desc()
    d.layers[1].n_in = 7
    assert lib.hvi_plan_create(C.byref(d), C.byref(h)) == A.HVI_EINVAL and b"chain" in lib.hvi_last_error(None)
    import torch
    if not torch.cuda.is_available():
        d = hvi_b200.DoubleMMCase().to_desc()
        rc = lib.hvi_plan_create(C.byref(d), C.byref(h))
        assert rc == A.HVI_ECUDA and b"no CPU fallback" in lib.hvi_last_error(None)
        with pytest.raises(hvi_b200.HVIError):
            hvi_b200.NegElboPlan(hvi_b200.DoubleMMCase())


def test_missing_library_raises():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        A.load_library(os.path.join(ROOT, "does_not_exist.so"))


def test_entry_points_reject_a_null_plan():
    """every plan-taking entry point returns HVI_EINVAL for plan == NULL instead of touching the device"""
    lib = A.load_library()
    n = None
    assert lib.hvi_plan_set_data(n, 1, n, n, n, n, A.HVI_MEM_HOST) == A.HVI_EINVAL
    assert lib.hvi_plan_prefetch_data(n, 1, n, n, n, n) == A.HVI_EINVAL
    assert lib.hvi_plan_commit_data(n) == A.HVI_EINVAL
    assert lib.hvi_plan_set_sites(n, None, 0) == A.HVI_EINVAL
    assert lib.hvi_neg_elbo_grad(n, 1, n, n, n, A.HVI_MEM_HOST, None, None, n) == A.HVI_EINVAL
    assert lib.hvi_neg_elbo_grad_rng(n, 1, n, 0, 0, A.HVI_MEM_HOST, None, None, n) == A.HVI_EINVAL
    assert lib.hvi_loss_gf_grad(n, n, A.HVI_MEM_HOST, None, n) == A.HVI_EINVAL
    assert lib.hvi_adam_init(n, n, 0.02, 0.9, 0.999, 1e-8) == A.HVI_EINVAL
    assert lib.hvi_adam_run(n, 1, 1, 0, 0, None) == A.HVI_EINVAL
    assert lib.hvi_adam_apply_dev(n, n, n) == A.HVI_EINVAL
    assert lib.hvi_adam_get_phi(n, n) == A.HVI_EINVAL
    assert lib.hvi_plan_destroy(n) == A.HVI_OK
    assert lib.hvi_phi_len(n) == -1 and lib.hvi_launch_count(n) == -1
