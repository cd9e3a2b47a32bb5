"""Error convention of the C ABI (include/hawkeye_b200.h), exercised WITHOUT a GPU: argument / shape / alignment /
workspace errors are detected before anything is launched (return < 0, hk_last_error() explains), never a slow path."""
import ctypes

import pytest


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build()
    from hawkeye_b200 import _lib
    return _lib.lib()


FAKE = 0x10000      # a non-null, 16-byte aligned address that must never be dereferenced on these paths


def err(lib):
    return lib.hk_last_error().decode()


def test_workspace_queries_are_pure(lib):
    a = lib.hk_bilinear_pool_fwd_workspace_bytes(32, 512, 196)
    assert a >= 32 * 4 and a == lib.hk_bilinear_pool_fwd_workspace_bytes(32, 512, 196)
    assert lib.hk_bilinear_pool_bwd_workspace_bytes(32, 512, 196) >= 32 * 512 * 512 * 4
    assert lib.hk_cbp_bwd_workspace_bytes(2, 512, 8192) == (2 * 512 * 512 + 2 * 8192) * 4
    assert lib.hk_conv3x3_wgrad_workspace_bytes(64, 64) >= 9 * 64 * 64 * 4


def test_bilinear_pool_argument_errors(lib):
    f = lib.hk_bilinear_pool_fwd
    assert f(None, FAKE, None, 2, 512, 196, FAKE, 1 << 30, None) == -1 and 'null' in err(lib)
    assert f(FAKE, FAKE, None, 2, 100, 196, FAKE, 1 << 30, None) == -3 and 'multiple of 128' in err(lib)
    # H*W % 4 != 0 is supported (zero-padded copy in the workspace): the workspace query grows accordingly
    q = lib.hk_bilinear_pool_fwd_workspace_bytes
    q.restype = ctypes.c_size_t
    assert q(2, 512, 49) > q(2, 512, 52) and f(FAKE, FAKE, None, 2, 512, 49, FAKE, q(2, 512, 52), None) == -4
    assert f(FAKE + 4, FAKE, None, 2, 512, 196, FAKE, 1 << 30, None) == -2 and 'aligned' in err(lib)
    assert f(FAKE, FAKE, None, 2, 512, 196, FAKE, 16, None) == -4 and 'workspace' in err(lib)
    assert f(FAKE, FAKE, None, 0, 512, 196, FAKE, 1 << 30, None) == -1
    b = lib.hk_bilinear_pool_bwd
    assert b(FAKE, None, FAKE, 2, 512, 196, FAKE, 1 << 40, None) == -2
    assert b(FAKE, FAKE, FAKE, 2, 512, 196, FAKE, 16, None) == -4


def test_conv_argument_errors(lib):
    assert lib.hk_conv3x3_fwd(None, FAKE, None, FAKE, 2, 8, 8, 64, 64, 1, None) == -1
    assert lib.hk_conv3x3_fwd(FAKE, FAKE, None, FAKE, 2, 8, 8, 48, 64, 1, None) == -3 and 'multiples of 32' in err(lib)
    assert lib.hk_conv3x3_fwd(FAKE + 8, FAKE, None, FAKE, 2, 8, 8, 64, 64, 1, None) == -2
    assert lib.hk_conv3x3_wgrad(FAKE, FAKE, FAKE, None, 2, 8, 8, 64, 64, FAKE, 16, None) == -4
    assert lib.hk_maxpool2x2_fwd(FAKE, FAKE, 2, 7, 8, 64, 0, None) == -3
    assert lib.hk_conv3x3_first_fwd(FAKE, FAKE, None, FAKE, 2, 8, 8, 300, FAKE, 1 << 30, None) == -3


def test_head_and_mpncov_argument_errors(lib):
    assert lib.hk_linear_fwd(None, FAKE, None, FAKE, 2, 64, 8, FAKE, 1 << 30, None) == -1
    assert lib.hk_linear_fwd(FAKE, FAKE, None, FAKE, 2, 63, 8, FAKE, 1 << 30, None) == -3
    assert lib.hk_softmax_ce_ls(None, FAKE, FAKE, None, None, 2, 200, ctypes.c_float(0.1), ctypes.c_float(1.0), None) == -1
    assert lib.hk_sgd_momentum(FAKE, FAKE + 4, FAKE, 64, ctypes.c_float(0.1), ctypes.c_float(0.9), ctypes.c_float(0.0),
                               ctypes.c_float(1.0), 1, None) == -2
    assert lib.hk_covpool_fwd(None, FAKE, FAKE, 2, 256, 195, None) == -1
    assert lib.hk_sqrtm_fwd(FAKE, FAKE, FAKE, 2, 256, 1, FAKE, 1 << 40, None) == -3 and 'iterN' in err(lib)


def test_errors_are_thread_local_text(lib):
    lib.hk_bilinear_pool_fwd(FAKE, FAKE, None, 2, 100, 196, FAKE, 1 << 30, None)
    msg = err(lib)
    import threading
    seen = []
    t = threading.Thread(target=lambda: seen.append(lib.hk_last_error().decode()))
    t.start(); t.join()
    assert 'multiple of 128' in msg and seen == ['']


def test_round2_entry_points_argument_errors(lib):
    """hk_conv3x3_fwd_pool, hk_bn_bwd_ex, hk_l2norm_rows_*, hk_npair_loss: bad arguments are rejected before any launch."""
    p = lib.hk_conv3x3_fwd_pool
    assert p(FAKE, FAKE, None, None, None, 2, 8, 8, 64, 64, 0, None) == -1 and 'null output' in err(lib)
    assert p(FAKE, FAKE, None, FAKE, None, 2, 7, 8, 64, 64, 0, None) == -3 and 'even H/W' in err(lib)
    assert p(FAKE, FAKE, None, FAKE, FAKE + 8, 2, 8, 8, 64, 64, 0, None) == -3            # unaligned code buffer
    assert p(FAKE, FAKE, None, FAKE, None, 2, 8, 8, 48, 64, 0, None) == -3 and 'multiples of 32' in err(lib)
    lib.hk_set_precise(1)
    try:
        assert p(FAKE, FAKE, None, FAKE, None, 2, 8, 8, 64, 64, 0, None) == -3 and '3xTF32' in err(lib)
    finally:
        lib.hk_set_precise(0)
    b = lib.hk_bn_bwd_ex
    assert b(FAKE, None, FAKE, FAKE, None, FAKE, FAKE, FAKE, None, FAKE, FAKE, 64, 64, 1, FAKE, 1 << 30, None) == -1  # relu needs y or beta
    assert b(FAKE, None, FAKE, FAKE, FAKE, FAKE, FAKE, FAKE, None, FAKE, FAKE, 64, 62, 1, FAKE, 1 << 30, None) == -3
    assert b(FAKE, None, FAKE, FAKE, FAKE, FAKE, FAKE, FAKE, None, FAKE, FAKE, 64, 64, 1, FAKE, 16, None) == -4
    assert lib.hk_l2norm_rows_fwd(None, FAKE, FAKE, 4, 64, None) == -1
    assert lib.hk_l2norm_rows_bwd(FAKE, FAKE, None, FAKE, 4, 64, None) == -1
    assert lib.hk_npair_loss(FAKE, FAKE, FAKE, None, FAKE, 8, None) == -1
    assert lib.hk_npair_loss(FAKE, FAKE, FAKE, FAKE, FAKE, 0, None) == -1
