"""numpy restatement of the same hot path (second, independent ORACLE).  TEST INFRASTRUCTURE ONLY.

Independent of oracle/eagle_oracle.c: the file is parsed with numpy byte arithmetic and every
product is numpy's (OpenBLAS dgemm, all host threads).  It doubles as the fastest honest CPU
baseline available in this image for bench.py (`cpu_baseline` / `--impl reference`), because the
reference's own GEMMs are Eigen's blocked kernels and OpenBLAS is the same class of code.

Also holds the restated file-format writers (the bytes the decode kernel eats):
  CreateASCIInospace.cpp:84-93, 119-122  (AA->'0', AB->'1', BB->'2', missing->'1', '\n' per row)
  createMt_ASCII_rcpp.cpp:87-118         (Mt.ascii = transpose, L lines x n chars)
Paths are relative to /root/reference/MyPackage/Eagle/src/.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------- file format
def write_ascii(path, G012):
    """G012: (rows, cols) uint8 in {0,1,2}.  Writes rows lines of cols chars + '\\n'."""
    G = np.asarray(G012, dtype=np.uint8)
    rows, cols = G.shape
    buf = np.empty((rows, cols + 1), dtype=np.uint8)
    buf[:, :cols] = G + ord("0")
    buf[:, cols] = ord("\n")
    with open(path, "wb") as f:
        f.write(buf.tobytes())


def ascii_image(G012):
    G = np.asarray(G012, dtype=np.uint8)
    rows, cols = G.shape
    buf = np.empty((rows, cols + 1), dtype=np.uint8)
    buf[:, :cols] = G + ord("0")
    buf[:, cols] = ord("\n")
    return buf


def create_ascii_nospace(text_path, out_path, AA="0", AB="1", BB="2", missing="NA"):
    """CreateASCIInospace.cpp:67-125: tokenise a space-separated genotype file."""
    rows = []
    with open(text_path) as f:
        for line in f:
            toks = line.split()
            if not toks:
                continue
            r = np.empty(len(toks), dtype=np.uint8)
            for i, t in enumerate(toks):
                if t == BB:
                    r[i] = 2
                elif t == AB:
                    r[i] = 1
                elif t == AA:
                    r[i] = 0
                elif t == missing:
                    r[i] = 1  # :91-93 missing genotypes become hets
                else:
                    raise ValueError(f"unexpected genotype token {t!r}")
            rows.append(r)
    L = len(rows[0])
    if any(len(r) != L for r in rows):
        raise ValueError("unequal number of columns per row")  # :108-117
    G = np.vstack(rows)
    write_ascii(out_path, G)
    return G


# ----------------------------------------------------------------------------- hot path
def ReadBlock(asciifname, start_row, numcols, numrows_in_block):
    """ReadBlock.cpp:16-68.  Assumes fixed line pitch numcols_in_file+1 (what the writers produce)."""
    with open(asciifname, "rb") as f:
        first = f.readline()
        pitch = len(first)
        f.seek(start_row * pitch)
        raw = f.read(numrows_in_block * pitch)
    if len(raw) < numrows_in_block * pitch - 1:
        raise ValueError("truncated file")
    if len(raw) == numrows_in_block * pitch - 1:
        raw += b"\n"
    a = np.frombuffer(raw, dtype=np.uint8).reshape(numrows_in_block, pitch)[:, :numcols]
    return np.asfortranarray(a.astype(np.float64) - 49.0)  # (c - '0') - 1


def _zero(selected_loci):
    s = np.atleast_1d(np.asarray(selected_loci, dtype=np.float64))
    if s.size == 0 or np.isnan(s[0]):
        return None
    return s.astype(np.int64)


def calculateMMt_rcpp(f_name_ascii, max_memory_in_Gbytes, num_cores, selected_loci, dims, quiet=True, message=None):
    n, L = int(dims[0]), int(dims[1])
    G = ReadBlock(f_name_ascii, 0, L, n)
    z = _zero(selected_loci)
    if z is not None:
        G[:, z] = 0.0
    return G @ G.T


def calculate_a_and_vara_rcpp(f_name_ascii, selected_loci, inv_MMt_sqrt, dim_reduced_vara,
                              max_memory_in_Gbytes, dims, a, quiet=True, message=None):
    L, n = int(dims[0]), int(dims[1])
    Mt = ReadBlock(f_name_ascii, 0, n, L)
    z = _zero(selected_loci)
    if z is not None:
        Mt[z, :] = 0.0
    S = np.asarray(inv_MMt_sqrt, dtype=np.float64)
    V = np.asarray(dim_reduced_vara, dtype=np.float64)
    av = np.asarray(a, dtype=np.float64).reshape(-1)
    ans = Mt @ (S @ av)                       # :90-91
    W = S @ (V @ S)                           # :97-98
    T = Mt @ W                                # :103
    vara = np.einsum("ij,ij->i", T, Mt)       # :107-112
    return {"a": ans.reshape(L, 1), "vara": vara.reshape(L, 1)}


def calculate_reduced_a_rcpp(f_name_ascii, varG, P, y, max_memory_in_Gbytes, dims, selected_loci,
                             quiet=True, message=None):
    n, L = int(dims[0]), int(dims[1])
    Mt = ReadBlock(f_name_ascii, 0, n, L)
    z = _zero(selected_loci)
    if z is not None:
        Mt[z, :] = 0.0
    ar = np.asarray(P, dtype=np.float64) @ np.asarray(y, dtype=np.float64).reshape(-1)
    return (varG * (Mt @ ar)).reshape(L, 1)


def extract_geno_rcpp(f_name_ascii, max_memory_in_Gbytes, selected_locus, dims):
    n, L = int(dims[0]), int(dims[1])
    with open(f_name_ascii, "rb") as f:
        raw = np.frombuffer(f.read(), dtype=np.uint8)
    col = raw[selected_locus::L + 1][:n]
    return col.astype(np.int32) - 49


def a_and_vara_longdouble(Mt_int8, S, V, a, rows):
    """Higher-precision (numpy longdouble, 64-bit mantissa on x86) evaluation of a / vara for a few
    marker rows; used to bound the rounding error of BOTH oracle and GPU results."""
    ld = np.longdouble
    S_, V_, a_ = S.astype(ld), V.astype(ld), np.asarray(a).reshape(-1).astype(ld)
    v = S_ @ a_
    W = S_ @ (V_ @ S_)
    out_a, out_v = [], []
    for r in rows:
        m = Mt_int8[r].astype(ld)
        out_a.append(m @ v)
        out_v.append((m @ W) @ m)
    return np.array(out_a, dtype=ld), np.array(out_v, dtype=ld)


# ----------------------------------------------------------------------------- packed 2-bit container
def pack_2bit(G):
    """MyPackage/RcppFunctions.cpp.gpu:224-345 (CreatePackedBinary), restated: G holds the file's codes 0/1/2
    (rows x cols).  Each row becomes ceil(cols/32) unsigned 64-bit words, genotype k in bits 2(k%32), 2(k%32)+1 of
    word k//32 (bit 2k' set for AB = 1, bit 2k'+1 for BB = 2, :316-326); a row starts on a fresh word (:304-306,
    :329-333) and unused bits stay 0 (packed.reset())."""
    G = np.asarray(G)
    rows, cols = G.shape
    wpr = (cols + 31) // 32
    P = np.zeros((rows, wpr * 32), dtype=np.uint64)
    P[:, :cols] = G.astype(np.uint64)
    P = P.reshape(rows, wpr, 32)
    shifts = (2 * np.arange(32, dtype=np.uint64))[None, None, :]
    return np.bitwise_or.reduce(P << shifts, axis=2)


def unpack_2bit(words, cols):
    """Inverse of pack_2bit -> codes 0/1/2/3 (rows x cols)."""
    words = np.asarray(words, dtype=np.uint64)
    rows, wpr = words.shape
    shifts = (2 * np.arange(32, dtype=np.uint64))[None, None, :]
    codes = ((words[:, :, None] >> shifts) & np.uint64(3)).reshape(rows, wpr * 32)
    return codes[:, :cols].astype(np.uint8)


# ----------------------------------------------------------------------------- ingest
# Independent restatements in plain Python of what the two ingest routines WRITE and WHY they stop, used to cross-check
# oracle/eagle_oracle.c (which is itself checked against the compiled reference).  Small inputs only.
_WS = b" \t\n\v\f\r"


def tokenise_text(data: bytes, cols: int, AA: str, AB: str, BB: str, missing: str):
    """CreateASCIInospace.cpp:67-125 -> (ok, rows written as bytes, reason).  reason: None | ("token", row1, token) |
    ("columns", row1, count)."""
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()  # getline does not return an empty piece after the final newline
    out = bytearray()
    codes = [(BB.encode(), b"2"), (AB.encode(), b"1"), (AA.encode(), b"0"), (missing.encode(), b"1")]
    for r, line in enumerate(lines):
        row = bytearray()
        for tok in line.translate(bytes.maketrans(_WS, b" " * len(_WS))).split():
            for c, o in codes:
                if tok == c:
                    row += o
                    break
            else:
                return False, bytes(out), ("token", r + 1, tok.decode("latin-1"))
        if len(row) != cols:
            return False, bytes(out), ("columns", r + 1, len(row))
        out += row + b"\n"
    return True, bytes(out), None


def plink_genotypes(data: bytes, ncols: int):
    """CreateASCIInospace_PLINK.cpp:55-200 for single-character allele tokens -> (ok, rows written, reason, warned).
    reason: None | ("columns", row1, count) | ("alleles", snp1, row1)."""
    nsnp = (ncols - 6) // 2
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    a0, a1 = [None] * nsnp, [None] * nsnp
    out = bytearray()
    warned = False
    for r, line in enumerate(lines):
        toks = line.translate(bytes.maketrans(_WS, b" " * len(_WS))).split()
        if len(toks) != ncols:
            return False, bytes(out), ("columns", r + 1, len(toks)), warned
        al = [t[:1] for t in toks[6:]]
        row = bytearray()
        for i in range(nsnp):
            a, b = al[2 * i], al[2 * i + 1]
            miss = a in (b"0", b"-") or b in (b"0", b"-")
            if r == 0:
                a0[i], a1[i] = (b"I", b"I") if miss else (a, b)
            if miss:
                warned = True
                a = b = b"I"
            for x in (b, a):  # the second allele is looked at first (:137)
                if x != a0[i] and x != a1[i] and x != b"I":
                    if a0[i] == b"I":
                        a0[i] = x
                    elif a1[i] == b"I":
                        a1[i] = x
                    elif a0[i] == a1[i]:
                        a1[i] = x
                    else:
                        return False, bytes(out), ("alleles", i + 1, r + 1), warned
            row += b"1" if (a == b"I" or b == b"I" or a != b) else (b"0" if a == a0[i] else b"2")
        out += row + b"\n"
    return True, bytes(out), None, warned
