"""ctypes front-end of the CPU ORACLE (oracle/eagle_oracle.c).  TEST INFRASTRUCTURE ONLY.

Function names, argument order and meaning follow the reference's Rcpp exports
(/root/reference/MyPackage/Eagle/src/RcppExports.cpp:9, 37, 54, 73, 129) so the parity
tests can call oracle and GPU path with the same arguments.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libeagle_oracle.so")

#: R's NA_real_ (a NaN with low word 1954, see R_IsNA); selected_loci = [NA] means "none".
NA_REAL = struct.unpack("<d", struct.pack("<Q", 0x7FF00000000007A2))[0]

ERRORS = {1: "ERROR: Could not open ", 2: "short line / truncated file", 3: "allocation failed",
          4: "soft failure (reference returns zeros)", 5: "block size 0 (reference divides by zero)"}


class OracleError(RuntimeError):
    pass


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "eagle_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None
_REF_SO = os.path.join(_HERE, "_ref", "libeagle_ref.so")


def _bind(path):
    L = C.CDLL(path)
    dp, lp, ip = C.POINTER(C.c_double), C.POINTER(C.c_long), C.POINTER(C.c_int)
    L.eo_ReadBlock.argtypes = [C.c_char_p, C.c_long, C.c_long, C.c_long, dp]
    L.eo_calculateMMt.argtypes = [C.c_char_p, C.c_double, C.c_int, dp, C.c_long, lp, dp, ip]
    L.eo_calculate_a_and_vara.argtypes = [C.c_char_p, dp, C.c_long, dp, dp, C.c_double, lp, dp, dp, dp, ip]
    L.eo_calculate_reduced_a.argtypes = [C.c_char_p, C.c_double, dp, dp, C.c_double, lp, dp, C.c_long, dp]
    L.eo_extract_geno.argtypes = [C.c_char_p, C.c_double, C.c_long, lp, ip, ip]
    L.eo_createM_ASCII.argtypes = [C.c_char_p] * 6 + [C.c_double, lp, C.c_int, C.c_char_p, C.c_char_p, C.c_long, ip]
    L.eo_createMt_ASCII.argtypes = [C.c_char_p] * 3 + [C.c_double, lp, C.c_int, C.c_char_p, C.c_long]
    L.eo_ReshapeM.argtypes = [C.c_char_p, C.c_char_p, lp, C.c_long, lp, lp]
    L.eo_getRowColumn.argtypes = [C.c_char_p, lp]
    L.eo_num_threads.restype = C.c_int
    if hasattr(L, "eo_set_dgemm"):
        L.eo_set_dgemm.argtypes = [C.c_void_p]
        L.eo_last_stages.argtypes = [dp]
    if hasattr(L, "ref_last_error"):
        L.ref_last_error.restype = C.c_char_p
    return L


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = _bind(_SO)
    return _lib


def reference_available() -> bool:
    """True when oracle/_ref/libeagle_ref.so exists: the reference's OWN sources compiled from
    /root/reference against the stand-in headers of oracle/refshim (built by oracle/refshim/Makefile)."""
    return os.path.exists(_REF_SO)


class use_reference:
    """Context manager: route this module's functions to the compiled reference sources."""

    def __enter__(self):
        global _lib
        self._saved = _lib
        _lib = _bind(_REF_SO)
        return self

    def __exit__(self, *exc):
        global _lib
        _lib = self._saved
        return False


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _check(rc, what):
    if rc:
        L = lib()
        if hasattr(L, "ref_last_error") and rc == 1:  # the compiled reference threw (Rcpp::stop -> exception)
            raise OracleError(f"{what}: {L.ref_last_error().decode('utf-8', 'replace')}")
        raise OracleError(f"{what}: {ERRORS.get(rc, rc)}")


def _sel(selected_loci):
    s = np.atleast_1d(np.asarray(selected_loci, dtype=np.float64)).copy()
    return s


def _dims(dims):
    return (C.c_long * 2)(int(dims[0]), int(dims[1]))


def ReadBlock(asciifname, start_row, numcols, numrows_in_block):
    """ReadBlock.cpp:16-68 -> (numrows x numcols) float64, column-major (Fortran order)."""
    M = np.empty((numrows_in_block, numcols), dtype=np.float64, order="F")
    _check(lib().eo_ReadBlock(os.fsencode(asciifname), start_row, numcols, numrows_in_block, _d(M)),
           f"ReadBlock({asciifname})")
    return M


def calculateMMt_rcpp(f_name_ascii, max_memory_in_Gbytes, num_cores, selected_loci, dims,
                      quiet=True, message=None, return_branch=False):
    """calculateMMt_rcpp.cpp:19-185.  dims = (n, L)."""
    n = int(dims[0])
    out = np.empty((n, n), dtype=np.float64, order="F")
    s = _sel(selected_loci)
    br = C.c_int(-1)
    _check(lib().eo_calculateMMt(os.fsencode(f_name_ascii), float(max_memory_in_Gbytes), int(num_cores),
                                 _d(s), len(s), _dims(dims), _d(out), C.byref(br)), "calculateMMt_rcpp")
    return (out, br.value) if return_branch else out


def calculate_a_and_vara_rcpp(f_name_ascii, selected_loci, inv_MMt_sqrt, dim_reduced_vara,
                              max_memory_in_Gbytes, dims, a, quiet=True, message=None,
                              return_branch=False):
    """calculate_a_and_vara_rcpp.cpp:22-241.  dims = (L, n) = dims of Mt.  -> dict(a, vara) of L x 1."""
    Lm, n = int(dims[0]), int(dims[1])
    S = np.asfortranarray(inv_MMt_sqrt, dtype=np.float64)
    V = np.asfortranarray(dim_reduced_vara, dtype=np.float64)
    av = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    assert S.shape == (n, n) and V.shape == (n, n) and av.shape == (n,)
    oa = np.empty(Lm, dtype=np.float64)
    ov = np.empty(Lm, dtype=np.float64)
    s = _sel(selected_loci)
    br = C.c_int(-1)
    rc = lib().eo_calculate_a_and_vara(os.fsencode(f_name_ascii), _d(s), len(s), _d(S), _d(V),
                                       float(max_memory_in_Gbytes), _dims(dims), _d(av), _d(oa), _d(ov),
                                       C.byref(br))
    if rc == 4:  # :141-142  List(a=0, vara=0)
        res = {"a": 0, "vara": 0}
    else:
        _check(rc, "calculate_a_and_vara_rcpp")
        res = {"a": oa.reshape(Lm, 1), "vara": ov.reshape(Lm, 1)}
    return (res, br.value) if return_branch else res


def calculate_reduced_a_rcpp(f_name_ascii, varG, P, y, max_memory_in_Gbytes, dims, selected_loci,
                             quiet=True, message=None):
    """calculate_reduced_a_rcpp.cpp:20-171.  dims = (n, L) = dims of M (file is Mt.ascii)."""
    n, Lm = int(dims[0]), int(dims[1])
    Pm = np.asfortranarray(P, dtype=np.float64)
    yv = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
    out = np.empty(Lm, dtype=np.float64)
    s = _sel(selected_loci)
    rc = lib().eo_calculate_reduced_a(os.fsencode(f_name_ascii), float(varG), _d(Pm), _d(yv),
                                      float(max_memory_in_Gbytes), _dims(dims), _d(s), len(s), _d(out))
    if rc == 4:
        return np.zeros((1, 1))
    _check(rc, "calculate_reduced_a_rcpp")
    return out.reshape(Lm, 1)


def extract_geno_rcpp(f_name_ascii, max_memory_in_Gbytes, selected_locus, dims, return_branch=False):
    """extract_geno_rcpp.cpp:17-86.  dims = (n, L); selected_locus 0-based.  -> int32[n]."""
    n = int(dims[0])
    out = np.empty(n, dtype=np.int32)
    br = C.c_int(-1)
    _check(lib().eo_extract_geno(os.fsencode(f_name_ascii), float(max_memory_in_Gbytes), int(selected_locus),
                                 _dims(dims), out.ctypes.data_as(C.POINTER(C.c_int)), C.byref(br)),
           "extract_geno_rcpp")
    return (out, br.value) if return_branch else out


def _messages(buf):
    raw = buf.value.decode("utf-8", "replace")
    return raw.split("\x1e") if raw else []


def createM_ASCII_rcpp(f_name, f_name_ascii, type, AA, AB, BB, max_memory_in_Gbytes, dims, quiet, missing):
    """createM_ASCII_rcpp.cpp:19-106 + CreateASCIInospace.cpp:17-164 and, for type "PLINK",
    CreateASCIInospace_PLINK.cpp:16-248.  -> (ok, [message, ...]); the no-space ASCII file is written to f_name_ascii."""
    buf = C.create_string_buffer(1 << 16)
    ok = C.c_int(0)
    e = lambda x: x.encode() if isinstance(x, str) else os.fsencode(x)
    _check(lib().eo_createM_ASCII(os.fsencode(f_name), os.fsencode(f_name_ascii), e(type), e(AA), e(AB), e(BB),
                                  float(max_memory_in_Gbytes), _dims(dims), int(bool(quiet)), e(missing), buf, len(buf),
                                  C.byref(ok)), "createM_ASCII_rcpp")
    return bool(ok.value), _messages(buf)


def createMt_ASCII_rcpp(f_name, f_name_ascii, type, max_memory_in_Gbytes, dims, quiet):
    """createMt_ASCII_rcpp.cpp:15-245.  dims = (n, L) of M.ascii (f_name); writes Mt.ascii to f_name_ascii. -> messages."""
    buf = C.create_string_buffer(1 << 16)
    _check(lib().eo_createMt_ASCII(os.fsencode(f_name), os.fsencode(f_name_ascii), type.encode(), float(max_memory_in_Gbytes),
                                   _dims(dims), int(bool(quiet)), buf, len(buf)), "createMt_ASCII_rcpp")
    return _messages(buf)


def ReshapeM_rcpp(fnameM, fnameMt, indxNA, dims):
    """ReshapeM_rcpp.cpp:16-117 -> [rows kept, line length]; writes fnameM + "tmp" and fnameMt + "tmp"."""
    idx = (C.c_long * max(1, len(indxNA)))(*[int(i) for i in indxNA])
    out = (C.c_long * 2)()
    _check(lib().eo_ReshapeM(os.fsencode(fnameM), os.fsencode(fnameMt), idx, len(indxNA), _dims(dims), out), "ReshapeM_rcpp")
    return [int(out[0]), int(out[1])]


def getRowColumn(fname):
    """getRowColumn.cpp:19-72 -> [rows, columns]."""
    out = (C.c_long * 2)()
    _check(lib().eo_getRowColumn(os.fsencode(fname), out), "getRowColumn")
    return [int(out[0]), int(out[1])]


def num_threads() -> int:
    return int(lib().eo_num_threads())


# ----------------------------------------------------------------------------- timing baseline only (bench.py)
_blas = None


def use_openblas_dgemm(threads: int) -> str:
    """bench.py's CPU legs only: route the oracle's dense products through the OpenBLAS that numpy bundles (the
    reference runs them in Eigen's GEBP kernel, which the plain loops of eagle_oracle.c do not match in speed) and set
    the thread counts explicitly -- torchrun exports OMP_NUM_THREADS=1.  Returns a description of the BLAS."""
    global _blas
    import glob
    cands = sorted(glob.glob(os.path.join(os.path.dirname(np.__file__), "..", "numpy.libs", "libscipy_openblas64_*.so")))
    if not cands:
        raise OracleError("numpy's bundled OpenBLAS (ILP64) not found")
    if _blas is None:
        _blas = C.CDLL(cands[0])
    _blas.scipy_openblas_set_num_threads64_.argtypes = [C.c_int]
    _blas.scipy_openblas_set_num_threads64_(int(threads))
    lib().eo_set_dgemm(C.cast(_blas.scipy_cblas_dgemm64_, C.c_void_p))
    return "OpenBLAS bundled with numpy " + np.__version__ + " (scipy_cblas_dgemm64_)"


def use_plain_loops():
    lib().eo_set_dgemm(None)


def last_stages():
    """seconds: ReadBlock, M.Mt product, S*a, the two n^3 pre-products, Mt*W, row dots, Mt*v of the last export called"""
    out = (C.c_double * 8)()
    lib().eo_last_stages(out)
    return dict(zip(["readblock", "mmt_gemm", "S_a", "pre_products", "Mt_W", "rowdots", "Mt_v"], list(out)[:7]))
