"""GPU tests of the device-level entry points (eg_dev_*), and full-size (BASELINE config 2)
checks through size-independent properties plus an independent torch evaluation on the device."""
import numpy as np
import pytest

from eagleeverything_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import torch
    from eagleeverything_b200 import device
    device.init(0)
    return device, torch


def test_synth_kernel_matches_numpy_twin(dev):
    device, torch = dev
    n, L = 77, 1001
    img = device.synth_ascii(n, L, synth.GENO_SEED)
    torch.cuda.synchronize()
    got = img[: n * (L + 1)].cpu().numpy().reshape(n, L + 1)
    assert np.array_equal(got, synth.ascii_image(synth.genotypes(n, L)))
    shard = device.synth_ascii(n, 300, synth.GENO_SEED, col_offset=512, n_total=n)
    assert np.array_equal(shard[: n * 301].cpu().numpy().reshape(n, 301)[:, :300], got[:, 512:812])


@pytest.mark.parametrize("rows,cols", [(1, 1), (3, 15), (5, 16), (7, 17), (150, 4998), (4998, 150), (33, 8191),
                                       (9, 8192), (4, 8193), (2, 20000), (1000, 2000), (3, 127), (2, 128)])
def test_decode_bit_exact_all_alignments(dev, rows, cols):
    device, torch = dev
    G = synth.genotypes(rows, cols, seed=rows * 31 + cols)
    img_np = synth.ascii_image(G).reshape(-1)
    for lead in (0, 1, 5, 16):  # shift the image inside the buffer: every source misalignment class
        buf = torch.zeros(lead + img_np.size + 64, dtype=torch.uint8, device="cuda")
        buf[lead:lead + img_np.size] = torch.from_numpy(img_np).cuda()
        out, err = device.decode(buf, cols + 1, rows, cols, src_offset=lead)
        torch.cuda.synchronize()
        o = out.cpu().numpy()
        assert err.cpu().numpy()[0] == 0
        assert np.array_equal(o[:, :cols], 1 - G.astype(np.int8))      # store encoding: 1 - code (decode.cu)
        assert not o[:, cols:].any(), "row padding must be zero"


def test_decode_column_shard_and_bad_byte(dev):
    device, torch = dev
    rows, cols = 40, 5000
    G = synth.genotypes(rows, cols, seed=5)
    img_np = synth.ascii_image(G).reshape(-1).copy()
    buf = torch.from_numpy(np.concatenate([img_np, np.zeros(64, np.uint8)])).cuda()
    out, err = device.decode(buf, cols + 1, rows, 1111, src_offset=777)  # columns [777, 1888)
    assert np.array_equal(out.cpu().numpy()[:, :1111], 1 - G[:, 777:1888].astype(np.int8)) and err[0].item() == 0
    img_np[13 * (cols + 1) + 4000] = ord("3")
    buf = torch.from_numpy(np.concatenate([img_np, np.zeros(64, np.uint8)])).cuda()
    out, err = device.decode(buf, cols + 1, rows, cols)
    e = err.cpu().numpy()
    assert e[0] == 1 and e[1] == 13 and 0 <= e[2] <= 4000  # (row, first column) of the offending 16 KB unit


def test_transpose_and_extract(dev):
    device, torch = dev
    for rows, cols in [(1, 1), (63, 65), (64, 64), (150, 4998), (1000, 333)]:
        G = synth.genotypes(rows, cols, seed=rows + cols)
        buf = torch.from_numpy(np.concatenate([synth.ascii_image(G).reshape(-1), np.zeros(64, np.uint8)])).cuda()
        st, _ = device.decode(buf, cols + 1, rows, cols)
        tt = device.transpose(st, rows, cols)
        t = tt.cpu().numpy()
        assert np.array_equal(t[:, :rows], (1 - G.astype(np.int8)).T) and not t[:, rows:].any()
        c = device.extract_col(st, rows, cols // 2).cpu().numpy()
        assert np.array_equal(c, G[:, cols // 2].astype(np.int32) - 1)


@pytest.mark.parametrize("rows,cols", [(1, 1), (3, 127), (5, 128), (7, 129), (150, 4998), (300, 20001), (4998, 150)])
def test_kblocked_layout_matches_row_major(dev, rows, cols):
    """The K-blocked M store ([col/128][row][128]) holds the same genotypes; SYRK and transpose agree."""
    device, torch = dev
    G = synth.genotypes(rows, cols, seed=rows * 13 + cols)
    buf = torch.from_numpy(np.concatenate([synth.ascii_image(G).reshape(-1), np.zeros(64, np.uint8)])).cuda()
    st, _ = device.decode(buf, cols + 1, rows, cols)
    kb, err = device.decode_kb(buf, cols + 1, rows, cols)
    assert err[0].item() == 0
    nb = (cols + 127) // 128
    flat = kb.permute(1, 0, 2).reshape(rows, nb * 128)
    assert torch.equal(flat[:, :cols], st[:, :cols]) and not flat[:, cols:].any()
    assert torch.equal(torch.triu(device.syrk_kb(kb, rows, cols)), torch.triu(device.syrk(st, rows, cols)))
    assert torch.equal(device.transpose_kb(kb, rows, cols), device.transpose(st, rows, cols))
    for c in (0, cols // 2, cols - 1):
        assert torch.equal(device.extract_col(kb.view(-1), rows, c, kblocked=True), device.extract_col(st, rows, c))


def test_argmax_semantics(dev):
    device, torch = dev
    rng = np.random.default_rng(1)
    L = 100003
    a = rng.standard_normal(L)
    v = rng.random(L) + 0.5
    a[[10, 50000, 99999]] = 9.0
    v[[10, 50000, 99999]] = 1.0            # three-way tie at 81 -> first index
    a[5], v[5] = np.nan, 1.0               # NaN ignored
    a[7], v[7] = 0.0, 0.0                  # 0/0 -> NaN ignored
    best, idx = device.argmax_tsq(torch.from_numpy(a).cuda(), torch.from_numpy(v).cuda())
    assert idx.item() == 10 and best.item() == 81.0
    a[20], v[20] = 1.0, 0.0                # +Inf is kept by max(na.rm=TRUE)
    best, idx = device.argmax_tsq(torch.from_numpy(a).cuda(), torch.from_numpy(v).cuda())
    assert idx.item() == 20 and np.isinf(best.item())
    nan = torch.full((1000,), float("nan"), dtype=torch.float64, device="cuda")
    best, idx = device.argmax_tsq(nan, nan)
    assert idx.item() == -1


def test_gemv_matches_float64(dev):
    device, torch = dev
    rows, cols = 3000, 517
    G = synth.genotypes(cols, rows, seed=8)  # Mt = G.T is rows x cols
    buf = torch.from_numpy(np.concatenate([synth.ascii_image(G.T).reshape(-1), np.zeros(64, np.uint8)])).cuda()
    st, _ = device.decode(buf, cols + 1, rows, cols)
    x = np.random.default_rng(2).standard_normal(cols)
    y = device.gemv_i8(st, rows, cols, torch.from_numpy(x).cuda(), 2.5).cpu().numpy()
    np.testing.assert_allclose(y, 2.5 * ((G.T.astype(np.float64) - 1) @ x), rtol=1e-12, atol=1e-11)


@pytest.mark.parametrize("n,cuts", [(777, [0, 777]), (1500, [0, 416, 1500]), (2048, [0, 2048])])
def test_preproducts_int8_slices_match_float64(dev, n, cuts, monkeypatch):
    """W = S (V S) on the int8 tensor cores (prep_i8.cu) against the cuBLAS FP64 path and torch float64:
    badly scaled symmetric inputs (per-column exponents), ragged n, column shards, run-to-run identical."""
    import ctypes as C
    device, torch = dev
    from eagleeverything_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(n)

    def sym(shift):
        A = rng.standard_normal((n, n)); A = (A + A.T) / np.sqrt(n) + shift * np.eye(n)
        d = 2.0 ** rng.integers(-12, 13, n)
        return A * d[:, None] * d[None, :]
    S, V = sym(2.0), sym(1.5)
    Sd, Vd = torch.from_numpy(S).cuda(), torch.from_numpy(V).cuda()
    Kpad = (n + 31) // 32 * 32
    tmp = torch.empty(n * n, dtype=torch.float64, device="cuda")

    def run(mode, pair="0"):
        monkeypatch.setenv("EAGLE_PREP_MODE", mode)
        monkeypatch.setenv("EAGLE_PREP_PAIR", pair)
        Wp = torch.zeros(lib.eg_scan_wp_elems(n), dtype=torch.float64, device="cuda")
        for c0, c1 in zip(cuts[:-1], cuts[1:]):
            _lib.check(lib.eg_dev_scan_prepare_cols(C.c_void_p(Sd.data_ptr()), C.c_void_p(Vd.data_ptr()), n, c0, c1, 1,
                                                    C.c_void_p(tmp.data_ptr()), C.c_void_p(Wp.data_ptr()), None))
        torch.cuda.synchronize()
        return Wp[: Kpad * n].view(n, Kpad).T[:n, :]     # column-major content -> W[i, j]
    Wi, Wi2, Wf = run("i8"), run("i8"), run("f64")
    assert torch.equal(torch.triu(Wi), torch.triu(Wi2))
    # the CTA-pair kernel (256-row tiles over two SMs) forms the same integer sums and combines them in the same order
    assert torch.equal(torch.triu(Wi), torch.triu(run("i8", pair="1")))
    X = Vd @ Sd
    ref = Sd @ X
    eps = np.finfo(np.float64).eps
    # FP64 GEMM: componentwise bound.  Digit slices: the residual is relative to (row max) x (column max) of each
    # product, 2^-56 per operand entry and 6 * 2^-56 for the dropped levels, n terms
    comp = 4 * n * eps * (Sd.abs() @ (Vd.abs() @ Sd.abs()))
    rmax = lambda A: A.abs().max(dim=1).values
    cmax = lambda A: A.abs().max(dim=0).values
    norm1 = torch.outer(rmax(Vd), cmax(Sd))                  # error scale of X = V S
    normw = 8 * n * 2.0 ** -56 * (torch.outer(rmax(Sd), cmax(X)) + Sd.abs() @ norm1)
    assert (torch.triu(Wf - ref).abs() <= torch.triu(comp)).all()
    assert (torch.triu(Wi - ref).abs() <= torch.triu(comp + normw)).all()
    # typical (not worst-case) behaviour: the integer path is about as close to the FP64 GEMM as two FP64 GEMMs
    # with different summation orders are to each other
    rel = (torch.triu(Wi - Wf).abs() / comp).max().item()
    assert rel < 2.0, rel


@pytest.mark.parametrize("mode", [0, 1], ids=["dmma_f64", "tcgen05_i8"])
def test_scan_modes_agree_at_scale(dev, mode):
    """n = 3000 (not a multiple of any tile), 40,000 markers: both contractions against torch float64."""
    device, torch = dev
    from eagleeverything_b200 import api
    n, L = 3000, 40000
    img = device.synth_ascii(L, n, synth.GENO_SEED + 1)       # an Mt.ascii image: L rows of n characters
    tt, err = device.decode(img, n + 1, L, n)
    S, V, a = synth.scan_inputs(n, 5)
    Sd, Vd, ad = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (S, V, a))
    prev = api.get_scan_mode()
    api.set_scan_mode(mode)
    try:
        Wp = device.scan_prepare(Sd, Vd, ad, n)
        oa, ov = device.scan(tt, L, n, Wp, zero_rows=[7, 39999])
        torch.cuda.synchronize()
    finally:
        api.set_scan_mode(prev)
    W = Sd @ (Vd @ Sd)
    Mr = -tt[:, :n].double()   # the reference's genotype values: the store holds their negation
    ra, rv = Mr @ (Sd @ ad), ((Mr @ W) * Mr).sum(1)
    ra[[7, 39999]] = 0
    rv[[7, 39999]] = 0
    assert ((oa - ra).abs() <= 1e-9 * ra.abs() + 1e-12 * ra.abs().max()).all()
    assert ((ov - rv).abs() <= 1e-9 * rv.abs() + 1e-12 * rv.abs().max()).all()


def test_scan_at_config5_width(dev):
    """n = 20,000 individuals (BASELINE config 5's width) x 3,000 markers: beyond n = 18,724 a level of the sliced
    pre-products no longer fits one int32 accumulation and is split; checked against torch float64 on the device."""
    device, torch = dev
    n, L = 20000, 3000
    img = device.synth_ascii(L, n, synth.GENO_SEED + 3)
    tt, err = device.decode(img, n + 1, L, n)
    assert err[0].item() == 0
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    S = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); S = (S + S.T) * (0.5 / n ** 0.5); S.diagonal().add_(2.0)
    V = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); V = (V + V.T) * (0.5 / n ** 0.5); V.diagonal().add_(1.5)
    a = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    Wp = device.scan_prepare(S, V, a, n)
    oa, ov = device.scan(tt, L, n, Wp)
    torch.cuda.synchronize()
    del Wp
    X = V @ S
    del V
    W = S @ X
    del X
    Mr = -tt[:, :n].double()   # the reference's genotype values: the store holds their negation
    ra, rv = Mr @ (S @ a), ((Mr @ W) * Mr).sum(1)
    assert ((oa - ra).abs() <= 1e-9 * ra.abs() + 1e-12 * ra.abs().max()).all()
    assert ((ov - rv).abs() <= 1e-9 * rv.abs() + 1e-12 * rv.abs().max()).all()


def test_config2_full_size_properties(dev):
    """BASELINE config 2 (n=2,000 x L=500,000) on one GPU: decode -> M.Mt -> scan, checked by
    properties the domain offers and by an independent torch evaluation on the device."""
    device, torch = dev
    n, L = 2000, 500000
    img = device.synth_ascii(n, L, synth.GENO_SEED)
    st, err = device.decode(img, L + 1, n, L)
    assert err[0].item() == 0
    # decode == image - '1' (torch elementwise as the independent checker)
    view = img[: n * (L + 1)].view(n, L + 1)
    assert torch.equal(st[:, :L], (49 - view[:, :L].to(torch.int16)).to(torch.int8))
    assert not st[:, L:].any()
    stk, _ = device.decode_kb(img, L + 1, n, L)          # the layout the SYRK streams
    C32 = device.syrk_kb(stk, n, L)
    assert torch.equal(torch.triu(C32), torch.triu(device.syrk(st, n, L)))
    K = device.mmt_finalize(C32, n)
    torch.cuda.synchronize()
    assert torch.equal(K, K.T)
    # trace = number of non-heterozygous genotypes; exact
    assert K.diagonal().sum().item() == float((st[:, :L] != 0).sum().item())
    # independent evaluation: fp32 GEMM on +-1/0 values is exact while |entries| <= L < 2^24
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = torch.zeros((n, n), dtype=torch.float32, device="cuda")
    for c0 in range(0, L, 50000):
        blk = st[:, c0:c0 + 50000].float()
        ref += blk @ blk.T
    assert torch.equal(K, ref.double())
    del ref, blk
    # linearity over marker shards (what the multi-GPU all-reduce relies on)
    half = (L // 2) // 128 * 128
    Ca = device.syrk(st[:, :half], n, half)
    stb = st[:, half:].contiguous()
    Cb = device.syrk(stb, n, L - half)
    assert torch.equal(torch.triu(Ca + Cb), torch.triu(C32))
    del Ca, Cb, stb
    # scan on the transposed store, against torch float64 on sampled marker rows
    tt = device.transpose_kb(stk, n, L)
    assert torch.equal(tt, device.transpose(st, n, L))
    del stk
    S, V, a = synth.scan_inputs(n)
    Sd, Vd, ad = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (S, V, a))
    Wp = device.scan_prepare(Sd, Vd, ad, n)
    oa, ov = device.scan(tt, L, n, Wp)
    torch.cuda.synchronize()
    W = Sd @ (Vd @ Sd)      # S, V symmetric: row-major view == column-major content
    v = Sd @ ad
    rows = torch.from_numpy(np.random.default_rng(0).choice(L, 4096, replace=False)).cuda()
    Mr = -tt[rows, :n].double()
    ra = Mr @ v
    rv = ((Mr @ W) * Mr).sum(1)
    tol = 1e-9
    assert ((oa[rows] - ra).abs() <= tol * torch.maximum(ra.abs(), tol * ra.abs().max())).all()
    assert ((ov[rows] - rv).abs() <= tol * torch.maximum(rv.abs(), tol * rv.abs().max())).all()
    # argmax kernel == torch on the full vectors
    best, idx = device.argmax_tsq(oa, ov)
    tsq = oa * oa / ov
    assert best.item() == tsq.max().item() and idx.item() == int((tsq == tsq.max()).nonzero()[0].item())


def test_config3_full_size_properties(dev):
    """BASELINE config 3 (n=10,000 x L=1,000,000) on one GPU, the shape bench.py times: decode -> M.Mt -> scan,
    checked through size-independent properties and spot checks recomputed independently with torch
    (64 M.Mt rows, 1,024 markers' a / var(a))."""
    device, torch = dev
    n, L = 10000, 1000000
    img = device.synth_ascii(n, L, synth.GENO_SEED)
    kb, err = device.decode_kb(img, L + 1, n, L)
    assert err[0].item() == 0
    rng = np.random.default_rng(3)
    rows = np.sort(rng.choice(n, 64, replace=False))
    rows_t = torch.from_numpy(rows).cuda()
    view = img[: n * (L + 1)].view(n, L + 1)
    Mrows = kb[:, rows_t, :].permute(1, 0, 2).reshape(64, -1)           # 64 decoded rows out of the K-blocked store
    assert torch.equal(Mrows[:, :L], (49 - view[rows_t, :L].to(torch.int16)).to(torch.int8))
    assert not Mrows[:, L:].any()
    del view
    C32 = device.syrk_kb(kb, n, L)
    K = device.mmt_finalize(C32, n)
    torch.cuda.synchronize()
    assert torch.equal(K, K.T)
    # trace = number of non-heterozygous genotypes (exact integer)
    nz = sum(int((kb[b0:b0 + 512] != 0).sum().item()) for b0 in range(0, kb.shape[0], 512))
    assert K.diagonal().sum().item() == float(nz)
    # 64 rows of M.Mt recomputed with fp32 GEMMs that are exact per 32,768-marker chunk (|partial| < 2^24)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = torch.zeros((64, n), dtype=torch.float64, device="cuda")
    for b0 in range(0, kb.shape[0], 256):
        blk = kb[b0:b0 + 256]                                              # [256][n][128]
        full = blk.permute(1, 0, 2).reshape(n, -1).float()
        ref += (full[rows_t] @ full.T).double()
    assert torch.equal(K[rows_t], ref)
    del ref, full, blk
    # linearity over marker shards (what the multi-GPU all-reduce relies on): two halves, 128-aligned cut
    hb = kb.shape[0] // 2
    Ca = device.syrk_kb(kb[:hb].contiguous(), n, hb * 128)
    Cb = device.syrk_kb(kb[hb:].contiguous(), n, L - hb * 128)
    assert torch.equal(torch.triu(Ca + Cb), torch.triu(C32))
    del Ca, Cb, C32, K
    tt = device.transpose_kb(kb, n, L)
    mk = torch.from_numpy(np.sort(rng.choice(L, 1024, replace=False))).cuda()
    cols = kb.permute(1, 0, 2).reshape(n, -1)[:, mk]                     # the same markers, read as columns of M
    assert torch.equal(tt[mk, :n], cols.T.contiguous())
    assert not tt[:, n:].any()
    del kb, img, cols
    g = torch.Generator(device="cuda"); g.manual_seed(1234)
    S = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); S = (S + S.T) * (0.5 / n ** 0.5); S.diagonal().add_(2.0)
    V = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); V = (V + V.T) * (0.5 / n ** 0.5); V.diagonal().add_(1.5)
    a = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    Wp = device.scan_prepare(S, V, a, n)
    oa, ov = device.scan(tt, L, n, Wp)
    torch.cuda.synchronize()
    W = S @ (V @ S)
    Mr = -tt[mk, :n].double()
    ra, rv = Mr @ (S @ a), ((Mr @ W) * Mr).sum(1)
    tol = 1e-9
    assert ((oa[mk] - ra).abs() <= tol * torch.maximum(ra.abs(), tol * ra.abs().max())).all()
    assert ((ov[mk] - rv).abs() <= tol * torch.maximum(rv.abs(), tol * rv.abs().max())).all()
    assert bool((ov > 0).all())                                           # W is positive definite here
    best, idx = device.argmax_tsq(oa, ov)
    tsq = oa * oa / ov
    assert best.item() == tsq.max().item() and idx.item() == int((tsq == tsq.max()).nonzero()[0].item())


def test_cache_eviction_never_frees_a_store_in_use(tmp_path, monkeypatch):
    """createMt_ASCII_rcpp holds the M store of its input while the transpose is allocated; when that allocation fails and
    the path cache is searched for memory to give back, the held store must survive (it used to be freed and then read):
    the call fails with EG_ERR_ALLOC and the cached store still answers.  EAGLE_TEST_FAIL_ALLOC injects the failure."""
    from eagleeverything_b200 import _lib, api
    from oracle import np_oracle as npo
    n, L = 64, 1000
    G = synth.genotypes(n, L, seed=3)
    m, mt, m2 = str(tmp_path / "M.ascii"), str(tmp_path / "Mt.ascii"), str(tmp_path / "M2.ascii")
    npo.write_ascii(m, G)
    npo.write_ascii(m2, G[:, ::-1])
    api.cache_clear()
    K0 = api.calculateMMt_rcpp(m, 8, 1, [api.NA_REAL], (n, L))
    api.calculateMMt_rcpp(m2, 8, 1, [api.NA_REAL], (n, L))              # a second cached store, not held by anyone
    monkeypatch.setenv("EAGLE_TEST_FAIL_ALLOC", "1")
    with pytest.raises(_lib.EagleGpuError) as ei:
        api.createMt_ASCII_rcpp(m, mt, "text", 8, (n, L))               # evicts M2's store, never M's
    assert ei.value.code == _lib.EG_ERR_ALLOC
    monkeypatch.delenv("EAGLE_TEST_FAIL_ALLOC")
    assert np.array_equal(api.calculateMMt_rcpp(m, 8, 1, [api.NA_REAL], (n, L)), K0)
    api.createMt_ASCII_rcpp(m, mt, "text", 8, (n, L))
    want = (G.T + ord("0")).astype(np.uint8)
    want = np.concatenate([want, np.full((L, 1), ord("\n"), np.uint8)], axis=1).tobytes()
    assert open(mt, "rb").read() == want
    api.cache_clear()
