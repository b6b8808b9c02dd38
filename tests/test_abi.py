"""CPU tests: the C-ABI library builds for sm_100a, loads, and exports exactly the symbols
include/spalign.h declares (no compute calls -- there is no GPU here), plus the host logic."""
import os
import re

import numpy as np

from superpixel_align_b200 import _lib


def test_library_exports_every_header_symbol():
    import __graft_entry__ as ge
    ge.build()
    lib = _lib.load()
    syms = _lib.header_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(lib, s), s
    assert set(syms) == set(_lib.SIGNATURES)
    assert lib.spalign_abi_version() == _lib.ABI_VERSION


def test_header_constants_match_binding():
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                            'include', 'spalign.h')).read()
    defs = dict(re.findall(r'#define\s+(SPALIGN_[A-Z0-9_]+)\s+\(?(-?\d+)\)?', hdr))
    assert int(defs['SPALIGN_ABI_VERSION']) == _lib.ABI_VERSION
    assert (int(defs['SPALIGN_I32']), int(defs['SPALIGN_I64']), int(defs['SPALIGN_U8'])) == \
        (_lib.I32, _lib.I64, _lib.U8)
    assert (int(defs['SPALIGN_F32']), int(defs['SPALIGN_F64'])) == (_lib.F32, _lib.F64)
    assert int(defs['SPALIGN_F_NNZ_OVERFLOW']) == _lib.F_NNZ_OVERFLOW
    assert int(defs['SPALIGN_KM_RUNNING']) == _lib.KM_RUNNING
    assert int(defs['SPALIGN_KM_ITER_CAP']) == _lib.KM_ITER_CAP


def test_argument_errors_are_reported_without_a_gpu():
    lib = _lib.load()
    rc = lib.spalign_pool(None, 1, 512, 8, 8, None, 1, 1, None, None, None, None, None, None, 1,
                          None, 516, None)
    assert rc == 1 and b'NULL' in lib.spalign_last_error()
    rc = lib.spalign_kmeans_groups(None, 0, 516, 0, 0, 0, None, 514, 9, 10, None, 1, None, None,
                                   None, None, None, 0, None)
    assert rc != 0
    assert lib.spalign_overlap_workspace_bytes(2, 1024, 2048, 128, 256, 2000, 200000) > 2 * 200000 * 36


def test_host_init_matches_oracle_stream():
    from oracle import spalign_oracle as so
    from superpixel_align_b200 import batch_spalign_kmeans as bsk, pipeline
    rs = np.random.RandomState(0)
    w = rs.uniform(0, 1, 501)
    for k in (2, 4, 7):
        np.random.seed(1111)
        a = bsk._host_init(k, w)
        np.random.seed(1111)
        b = so.kmeans_init(k, w)
        assert np.array_equal(a, b.astype(np.int32))
    # speculative shuffles (drawn before the weights are known) replay the same stream
    sizes = [1000, 37, 1, 0, 513]
    np.random.seed(5)
    flat, off, m = pipeline.draw_shuffles(4, sizes)
    np.random.seed(5)
    for g, n in enumerate(sizes):
        if n == 0:
            continue
        wg = rs.uniform(0, 1, n)
        init = so.kmeans_init(4, wg)
        low = wg <= np.sort(wg)[n // 2]
        assert low.sum() == m[g]
        assert np.array_equal(init[low].astype(np.int32), flat[off[g]:off[g + 1]])


def test_prior_axes_product_equals_reference_map():
    from oracle import spalign_oracle as so
    from superpixel_align_b200 import ops
    gy, gx = ops.prior_axes(64, 128, 0.75, 0.5, 0.1, 0.1)
    np.testing.assert_allclose(np.outer(gy, gx), so.create_prior_map(64, 128, 0.75, 0.5, 0.1, 0.1),
                               rtol=4e-15, atol=1e-300)
