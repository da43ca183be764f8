"""The packed 2-bit genotype container (format of the reference's pre-CRAN
CreatePackedBinary, MyPackage/RcppFunctions.cpp.gpu:224-345) against the oracle's restatement: bit-exact."""
import ctypes as C

import numpy as np
import pytest

from eagleeverything_b200 import synth


def test_oracle_pack_roundtrip_and_known_bits():
    from oracle import np_oracle as npo
    G = np.array([[0, 1, 2, 2, 1], [2, 2, 2, 0, 0]], dtype=np.uint8)
    W = npo.pack_2bit(G)
    assert W.shape == (2, 1) and W.dtype == np.uint64
    # row 0: codes 0,1,2,2,1 -> bits (LSB first) 00 10 01 01 10  = 0b01_10_10_01_00
    assert int(W[0, 0]) == 0b0110100100 and int(W[1, 0]) == 0b0000101010
    for rows, cols in [(3, 1), (2, 31), (2, 32), (5, 33), (4, 1000)]:
        G = synth.genotypes(rows, cols, seed=cols)
        assert np.array_equal(npo.unpack_2bit(npo.pack_2bit(G), cols), G)
        assert npo.pack_2bit(G).shape == (rows, (cols + 31) // 32)


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols", [(1, 1), (3, 31), (7, 32), (5, 33), (64, 127), (200, 128), (129, 1000), (300, 5000)])
def test_pack_unpack_bit_exact(rows, cols):
    import torch
    from eagleeverything_b200 import _lib, device
    from oracle import np_oracle as npo
    lib = device.init(0)
    G = synth.genotypes(rows, cols, seed=rows * 7 + cols)
    ref_words = npo.pack_2bit(G)
    buf = torch.from_numpy(np.concatenate([synth.ascii_image(G).reshape(-1), np.zeros(64, np.uint8)])).cuda()
    st, _ = device.decode(buf, cols + 1, rows, cols)              # row-major store
    kb, _ = device.decode_kb(buf, cols + 1, rows, cols)           # K-blocked store
    wpr = int(lib.eg_packed_words_per_row(cols))
    assert wpr == ref_words.shape[1]
    vp = lambda t: C.c_void_p(t.data_ptr())
    for store, pitch in ((st, st.stride(0)), (kb, 0)):
        words = torch.full((rows, wpr), -1, dtype=torch.int64, device="cuda")
        _lib.check(lib.eg_dev_pack_2bit(vp(store), rows, cols, pitch, vp(words), None))
        assert np.array_equal(words.cpu().numpy().view(np.uint64), ref_words)
    dwords = torch.from_numpy(ref_words.view(np.int64)).cuda()
    err = torch.zeros(4, dtype=torch.int32, device="cuda")
    out = torch.full_like(st, 5)
    _lib.check(lib.eg_dev_unpack_2bit(vp(dwords), rows, cols, vp(out), out.stride(0), vp(err), None))
    assert err[0].item() == 0 and torch.equal(out, st)            # including the zero pad of every row
    outk = torch.full_like(kb, 5)
    _lib.check(lib.eg_dev_unpack_2bit(vp(dwords), rows, cols, vp(outk), 0, vp(err), None))
    assert err[0].item() == 0 and torch.equal(outk, kb)


@pytest.mark.gpu
def test_packed_store_host_api_and_bad_code(tmp_path):
    """Host-level path: packed words -> store -> M.Mt equals the ASCII path bit for bit; a code 3 is refused."""
    import torch
    from eagleeverything_b200 import _lib, device
    from oracle import np_oracle as npo
    lib = device.init(0)
    n, L = 150, 3001
    G = synth.genotypes(n, L, seed=3)
    img = np.concatenate([synth.ascii_image(G).reshape(-1), np.zeros(64, np.uint8)])
    words = np.ascontiguousarray(npo.pack_2bit(G))
    h1, h2 = C.c_void_p(), C.c_void_p()
    _lib.check(lib.eg_store_from_host_ascii(img.ctypes.data_as(C.c_void_p), n, L, 0, L, C.byref(h1)))
    _lib.check(lib.eg_store_from_host_packed(words.ctypes.data_as(C.c_void_p), n, L, 1, C.byref(h2)))
    K1, K2 = np.empty((n, n), order="F"), np.empty((n, n), order="F")
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    _lib.check(lib.eg_store_mmt(h1, None, 0, dp(K1)))
    _lib.check(lib.eg_store_mmt(h2, None, 0, dp(K2)))
    assert np.array_equal(K1, K2)
    back = np.empty_like(words)
    _lib.check(lib.eg_store_to_host_packed(h1, back.ctypes.data_as(C.c_void_p)))
    assert np.array_equal(back, words)
    lib.eg_store_free(h1); lib.eg_store_free(h2)
    bad = words.copy()
    bad[17, 5] |= np.uint64(3) << np.uint64(10)
    h3 = C.c_void_p()
    rc = lib.eg_store_from_host_packed(bad.ctypes.data_as(C.c_void_p), n, L, 0, C.byref(h3))
    assert rc == _lib.EG_ERR_FORMAT and b"row 17" in lib.eg_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize("n,L,drop", [(150, 3001, [149, 77, 3, 0]), (40, 130, [5]), (257, 500, list(range(256, 0, -2)))])
def test_reshapeM_as_a_device_mask(n, L, drop):
    """ReshapeM_rcpp.cpp:59-109 on resident stores: M without the rows, Mt without the columns of the dropped
    individuals -- the scan inputs built from them equal those of freshly written reduced files."""
    import torch
    from eagleeverything_b200 import _lib, device
    lib = device.init(0)
    G = synth.genotypes(n, L, seed=n + L)
    keep = np.array([i for i in range(n) if i not in set(drop)])
    img = np.concatenate([synth.ascii_image(G).reshape(-1), np.zeros(64, np.uint8)])
    imgT = np.concatenate([synth.ascii_image(np.ascontiguousarray(G.T)).reshape(-1), np.zeros(64, np.uint8)])
    hM, hT, hM2, hT2 = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    _lib.check(lib.eg_store_from_host_ascii(img.ctypes.data_as(C.c_void_p), n, L, 0, L, C.byref(hM)))
    _lib.check(lib.eg_store_from_host_ascii_rows(imgT.ctypes.data_as(C.c_void_p), L, n, 0, L, C.byref(hT)))
    idx = (C.c_int64 * len(drop))(*drop)
    _lib.check(lib.eg_store_drop_individuals(hM, idx, len(drop), 1, C.byref(hM2)))
    _lib.check(lib.eg_store_drop_individuals(hT, idx, len(drop), 0, C.byref(hT2)))

    def dump(h):
        r, c, p, ptr = C.c_int64(), C.c_int64(), C.c_int64(), C.c_void_p()
        _lib.check(lib.eg_store_info(h, C.byref(r), C.byref(c), C.byref(p), C.byref(ptr)))
        return r.value, c.value, p.value
    nk = len(keep)
    assert dump(hM2)[:2] == (nk, L) and dump(hT2)[:2] == (L, nk)
    # M.Mt of the masked store == M.Mt of the reduced genotypes (exact integers)
    K = np.empty((nk, nk), order="F")
    _lib.check(lib.eg_store_mmt(hM2, None, 0, K.ctypes.data_as(C.POINTER(C.c_double))))
    Mk = G[keep].astype(np.float64) - 1
    assert np.array_equal(K, Mk @ Mk.T)
    # the masked Mt store == transpose of the masked M store
    hT3 = C.c_void_p()
    _lib.check(lib.eg_store_transpose(hM2, C.byref(hT3)))
    words = lambda h, r, c: (lambda w: (_lib.check(lib.eg_store_to_host_packed(h, w.ctypes.data_as(C.c_void_p))), w)[1])(np.empty((r, (c + 31) // 32), np.uint64))
    assert np.array_equal(words(hT2, L, nk), words(hT3, L, nk))
    bad = (C.c_int64 * 1)(n)
    assert lib.eg_store_drop_individuals(hM, bad, 1, 1, C.byref(C.c_void_p())) == _lib.EG_ERR_ARG
    for h in (hM, hT, hM2, hT2, hT3):
        lib.eg_store_free(h)


@pytest.mark.gpu
def test_pinned_upload_accumulates_mmt_under_the_copy():
    """eg_store_from_host_ascii on a page-locked image uploads in column chunks and accumulates M.Mt on the way
    (eg_store::C32): same store bytes and same product as the pageable row-block path, selected loci included,
    and a bad byte is still located."""
    import torch
    from eagleeverything_b200 import _lib, device
    lib = device.init(0)
    n, L = 300, 70001                                   # several chunks would need rows*cw > 256 MB: force via many rows? no: 1 chunk
    G = synth.genotypes(n, L, seed=8)
    img_np = np.concatenate([synth.ascii_image(G).reshape(-1), np.zeros(64, np.uint8)])
    pinned = torch.from_numpy(img_np.copy()).pin_memory()
    hP, hU = C.c_void_p(), C.c_void_p()
    _lib.check(lib.eg_store_from_host_ascii(C.c_void_p(pinned.data_ptr()), n, L, 0, L, C.byref(hP)))
    _lib.check(lib.eg_store_from_host_ascii(img_np.ctypes.data_as(C.c_void_p), n, L, 0, L, C.byref(hU)))
    wpr = (L + 31) // 32
    wP, wU = np.empty((n, wpr), np.uint64), np.empty((n, wpr), np.uint64)
    _lib.check(lib.eg_store_to_host_packed(hP, wP.ctypes.data_as(C.c_void_p)))
    _lib.check(lib.eg_store_to_host_packed(hU, wU.ctypes.data_as(C.c_void_p)))
    assert np.array_equal(wP, wU)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    M = G.astype(np.float64) - 1
    for zero in ([], [5, 69999, 123]):
        z = (C.c_int64 * max(1, len(zero)))(*zero) if zero else None
        KP, KU = np.empty((n, n), order="F"), np.empty((n, n), order="F")
        _lib.check(lib.eg_store_mmt(hP, z, len(zero), dp(KP)))
        _lib.check(lib.eg_store_mmt(hU, z, len(zero), dp(KU)))
        Mz = M.copy(); Mz[:, zero] = 0
        assert np.array_equal(KP, KU) and np.array_equal(KP, Mz @ Mz.T)
    lib.eg_store_free(hP); lib.eg_store_free(hU)
    pinned[13 * (L + 1) + 40000] = ord("7")
    h = C.c_void_p()
    rc = lib.eg_store_from_host_ascii(C.c_void_p(pinned.data_ptr()), n, L, 0, L, C.byref(h))
    assert rc == _lib.EG_ERR_FORMAT and b"row 13" in lib.eg_last_error()
