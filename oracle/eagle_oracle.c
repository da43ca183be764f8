/*
 * eagle_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the genome-scan hot path of Eagle / WMAM v1.0.3
 * (jcbowden/EagleEverything).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The
 * product path (eagleeverything_b200/csrc) never links, loads or calls it.
 *
 * PINNING.  The reference ships no tests, golden vectors or known-answer fixtures for this path
 * (SURVEY.md section 4), and the package itself needs R + Rcpp + RcppEigen, none of which exist in
 * this image.  What pins this file:
 *   - the reference's OWN five hot-path sources (ReadBlock.cpp, calculateMMt_rcpp.cpp,
 *     calculate_a_and_vara_rcpp.cpp, calculate_reduced_a_rcpp.cpp, extract_geno_rcpp.cpp) are compiled
 *     from where they lie under /root/reference against a stand-in for the RcppEigen headers
 *     (oracle/refshim -> oracle/_ref/libeagle_ref.so) and run here; tests/test_reference_shim.py
 *     checks this file against them in every branch (in-memory and blocked, with and without
 *     selected loci, soft failures, error text): M.Mt, ReadBlock and extract_geno bit-for-bit,
 *     a / var(a) / reduced-a to 1e-12.  Their control flow is the real thing; the dense products
 *     inside are the stand-in's loops, not Eigen's kernels (Eigen is not vendored in the reference);
 *   - M.Mt is exact integer arithmetic (|entry| <= L < 2^53): any correct implementation, in any
 *     summation order, produces the same bits;
 *   - an independent numpy restatement (oracle/np_oracle.py) and a long-double evaluation of
 *     a / var(a) agree with it (tests/test_oracle.py);
 *   - results on the shipped demo data are frozen in tests/golden/demo.npz.
 * What remains unpinned: Eigen's own summation order inside the products (a / var(a) can differ from
 * a real R run in the last bits), hence the 1e-9 tolerance on those two outputs.
 *
 * Paths cited below are relative to /root/reference/MyPackage/Eagle/.
 * Matrices are column-major doubles, exactly as Eigen::MatrixXd stores them.
 *
 * Arithmetic note: every GEMM / GEMV / dot the reference performs is done by
 * Eigen (header-only, via RcppEigen; NOT vendored in the reference and not
 * version pinned -- DESCRIPTION:42-43 "LinkingTo: RcppEigen, Rcpp").  Eigen's
 * blocked GEBP summation order is an implementation detail of that library;
 * this file uses a straightforward cache-blocked product.  The published
 * algorithm (C = A*B in IEEE double, round-to-nearest) is what is restated.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EO_OK 0
#define EO_ERR_OPEN 1      /* ReadBlock.cpp:42-45  Rcpp::stop("ERROR: Could not open ...") */
#define EO_ERR_SHORT 2     /* line shorter than numcols / file ends early (UB in the reference) */
#define EO_ERR_ALLOC 3
#define EO_ERR_SOFT 4      /* in-band soft failure (a=0, vara=0 / 1x1 zero) */
#define EO_ERR_BLOCK0 5    /* block size evaluates to 0 -> division by zero in the reference */

/* R's NA_real_ is a NaN whose low word is 1954 (R_IsNA).  Any NaN is taken
 * as "no selected loci": a non-NA NaN would index out of bounds in the
 * reference (calculateMMt_rcpp.cpp:88-92). */
static int eo_is_na(double x) { return isnan(x); }

/* ------------------------------------------------------------------ */
/* dense helpers (column-major)                                        */
/* ------------------------------------------------------------------ */

/* Timing baseline only (bench.py --impl reference / cpu_baseline): the reference's products run in Eigen's GEBP
 * kernel, which this file's plain loops do not match in speed.  eo_set_dgemm() installs an optimised column-major
 * dgemm with the ILP64 CBLAS signature (bench.py passes scipy_cblas_dgemm64_ of the OpenBLAS bundled with numpy) so
 * that THIS control flow can be timed at a BLAS rate.  Never set by the tests: parity runs on the loops below.
 * eo_last_stages(): wall seconds of the stages of the last export called (0 ReadBlock, 1 M.Mt product, 2 S*a and Mt*v,
 * 3 the two n^3 pre-products, 4 Mt*W, 5 row dots, 6 Mt*v). */
typedef void (*eo_dgemm_fn)(int order, int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha,
                            const double *A, int64_t lda, const double *B, int64_t ldb, double beta, double *C,
                            int64_t ldc);
static eo_dgemm_fn eo_dgemm_hook = NULL;
void eo_set_dgemm(void *fn) { eo_dgemm_hook = (eo_dgemm_fn)fn; }
static double eo_stage_s[8];
static double eo_now(void)
{
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}
void eo_last_stages(double *out) { memcpy(out, eo_stage_s, sizeof(eo_stage_s)); }

/* C(m x n) = A(m x k) * op(B);  transb==0: B is k x n;  transb==1: B is n x k (C = A*B^T) */
static void eo_gemm(long m, long n, long k, const double *A, long lda, const double *B, long ldb,
                    int transb, double *C, long ldc)
{
    const long MB = 256, KB = 256;
    if (eo_dgemm_hook) { /* 102 = CblasColMajor, 111 = CblasNoTrans, 112 = CblasTrans */
        eo_dgemm_hook(102, 111, transb ? 112 : 111, m, n, k, 1.0, A, lda, B, ldb, 0.0, C, ldc);
        return;
    }
#pragma omp parallel for schedule(dynamic, 8)
    for (long j = 0; j < n; j++) {
        double *c = C + j * ldc;
        for (long i = 0; i < m; i++) c[i] = 0.0;
    }
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
    for (long i0 = 0; i0 < m; i0 += MB) {
        for (long j0 = 0; j0 < n; j0 += 64) {
            long i1 = i0 + MB < m ? i0 + MB : m;
            long j1 = j0 + 64 < n ? j0 + 64 : n;
            for (long p0 = 0; p0 < k; p0 += KB) {
                long p1 = p0 + KB < k ? p0 + KB : k;
                for (long j = j0; j < j1; j++) {
                    double *c = C + j * ldc;
                    for (long p = p0; p < p1; p++) {
                        double b = transb ? B[j + p * ldb] : B[p + j * ldb];
                        const double *a = A + p * lda;
                        for (long i = i0; i < i1; i++) c[i] += a[i] * b;
                    }
                }
            }
        }
    }
}

/* y(m) = A(m x k) * x(k) */
static void eo_gemv(long m, long k, const double *A, long lda, const double *x, double *y)
{
#pragma omp parallel for schedule(static)
    for (long i0 = 0; i0 < m; i0 += 1024) {
        long i1 = i0 + 1024 < m ? i0 + 1024 : m;
        for (long i = i0; i < i1; i++) y[i] = 0.0;
        for (long p = 0; p < k; p++) {
            const double *a = A + p * lda;
            double xv = x[p];
            for (long i = i0; i < i1; i++) y[i] += a[i] * xv;
        }
    }
}

/* ------------------------------------------------------------------ */
/* ReadBlock  (src/ReadBlock.cpp:16-68)                                */
/* ------------------------------------------------------------------ */
/* M is numrows_in_block x numcols, column-major; M(r,c) = (line[c]-'0') - 1
 * (:52-55).  Lines 0..start_row-1 are read and discarded (:47-50). */
int eo_ReadBlock(const char *asciifname, long start_row, long numcols, long numrows_in_block,
                 double *M)
{
    FILE *f = fopen(asciifname, "r");
    if (!f) return EO_ERR_OPEN;
    char *line = NULL;
    size_t cap = 0;
    long rowi = 0;
    int rc = EO_OK;
    for (long rr = 0; rr < start_row + numrows_in_block; rr++) {
        ssize_t len = getline(&line, &cap, f);
        if (len < 0) { rc = EO_ERR_SHORT; break; }
        if (rr >= start_row) {
            if (len > 0 && line[len - 1] == '\n') len--;
            if (len < numcols) { rc = EO_ERR_SHORT; break; }
            for (long ii = 0; ii < numcols; ii++) {
                int tmp = line[ii] - '0';
                M[rowi + ii * numrows_in_block] = (double)tmp - 1;
            }
            rowi++;
        }
    }
    free(line);
    fclose(f);
    return rc;
}

/* ------------------------------------------------------------------ */
/* calculateMMt_rcpp  (src/calculateMMt_rcpp.cpp:19-185)               */
/* ------------------------------------------------------------------ */
/* dims = (n, L).  selected_loci are 0-based column indices, element 0 == NA
 * means "none" (:88).  Returns MMt n x n column-major.  *branch (optional)
 * reports 0 = in-memory (:84-95), 1 = blocked (:99-179). */
int eo_calculateMMt(const char *f_name_ascii, double max_memory_in_Gbytes, int num_cores,
                    const double *selected_loci, long n_selected, const long *dims, double *MMt,
                    int *branch)
{
#ifdef _OPENMP
    if (num_cores > 0) omp_set_num_threads(num_cores); /* :25-35 (sticky, as in the reference) */
#endif
    const long n = dims[0], L = dims[1];
    memset(MMt, 0, sizeof(double) * (size_t)n * (size_t)n); /* :54-55 */
    const int have_sel = (n_selected > 0) && !eo_is_na(selected_loci[0]);

    /* :75-76 */
    double memory_needed_in_Gb =
        ((double)n * n * sizeof(double) + 2.0 * ((double)n * L * sizeof(double))) / 1000000000.0;

    if (max_memory_in_Gbytes > memory_needed_in_Gb) { /* :84 */
        if (branch) *branch = 0;
        double *genoMat = (double *)malloc(sizeof(double) * (size_t)n * (size_t)L);
        if (!genoMat) return EO_ERR_ALLOC;
        double t0 = eo_now();
        int rc = eo_ReadBlock(f_name_ascii, 0, L, n, genoMat); /* :86 */
        eo_stage_s[0] = eo_now() - t0;
        if (rc) { free(genoMat); return rc; }
        if (have_sel) /* :88-92 */
            for (long ii = 0; ii < n_selected; ii++)
                memset(genoMat + (long)selected_loci[ii] * n, 0, sizeof(double) * (size_t)n);
        t0 = eo_now();
        eo_gemm(n, n, L, genoMat, n, genoMat, n, 1, MMt, n); /* :95 */
        eo_stage_s[1] = eo_now() - t0;
        free(genoMat);
        return EO_OK;
    }

    if (branch) *branch = 1;
    /* :103-106 */
    double part1 = -2.0 * (double)L;
    double part2 = 4.0 * (double)L * (double)L + 4.0 * max_memory_in_Gbytes * 1000000000.0 / sizeof(double);
    part2 = sqrt(part2);
    long num_rows_in_block = (long)((part1 + part2) / 2.2);
    if (num_rows_in_block <= 0) return EO_ERR_BLOCK0; /* :113 divides by it */

    long num_blocks = n / num_rows_in_block; /* :113-118 */
    if (n % num_rows_in_block) num_blocks++;

    double *B1 = (double *)malloc(sizeof(double) * (size_t)num_rows_in_block * (size_t)L);
    double *B2 = (double *)malloc(sizeof(double) * (size_t)num_rows_in_block * (size_t)L);
    double *S = (double *)malloc(sizeof(double) * (size_t)num_rows_in_block * (size_t)num_rows_in_block);
    if (!B1 || !B2 || !S) { free(B1); free(B2); free(S); return EO_ERR_ALLOC; }
    int rc = EO_OK;
    for (long i = 0; i < num_blocks && !rc; i++) { /* :121 */
        long start_row1 = i * num_rows_in_block;
        long nr1 = num_rows_in_block;
        if (start_row1 + nr1 > n) nr1 = n - start_row1;
        rc = eo_ReadBlock(f_name_ascii, start_row1, L, nr1, B1); /* :129 */
        if (rc) break;
        if (have_sel) /* :133-137 */
            for (long ii = 0; ii < n_selected; ii++)
                memset(B1 + (long)selected_loci[ii] * nr1, 0, sizeof(double) * (size_t)nr1);
        eo_gemm(nr1, nr1, L, B1, nr1, B1, nr1, 1, S, nr1); /* :138 */
        for (long c = 0; c < nr1; c++)                     /* :140 */
            for (long r = 0; r < nr1; r++) MMt[(start_row1 + r) + (start_row1 + c) * n] = S[r + c * nr1];
        for (long j = i + 1; j < num_blocks; j++) { /* :142 */
            long start_row2 = j * num_rows_in_block;
            long nr2 = num_rows_in_block;
            if (start_row2 + nr2 > n) nr2 = n - start_row2;
            rc = eo_ReadBlock(f_name_ascii, start_row2, L, nr2, B2); /* :148 */
            if (rc) break;
            if (have_sel) /* :156-160 */
                for (long jj = 0; jj < n_selected; jj++)
                    memset(B2 + (long)selected_loci[jj] * nr2, 0, sizeof(double) * (size_t)nr2);
            eo_gemm(nr1, nr2, L, B1, nr1, B2, nr2, 1, S, nr1); /* :161 */
            for (long c = 0; c < nr2; c++)                     /* :163-165 */
                for (long r = 0; r < nr1; r++) {
                    double v = S[r + c * nr1];
                    MMt[(start_row1 + r) + (start_row2 + c) * n] = v;
                    MMt[(start_row2 + c) + (start_row1 + r) * n] = v;
                }
        }
    }
    free(B1); free(B2); free(S);
    return rc;
}

/* ------------------------------------------------------------------ */
/* calculate_a_and_vara_rcpp  (src/calculate_a_and_vara_rcpp.cpp:22-241) */
/* ------------------------------------------------------------------ */
/* dims = (L, n): the dimensions of Mt (:39).  inv_MMt_sqrt and
 * dim_reduced_vara are n x n column-major; a is the n-vector of reduced BLUPs.
 * out_a, out_vara: L doubles each. */
static void eo_rowdots(long rows, long n, const double *T, const double *Mt, double *out)
{
    /* :107-112 / :210-216   var_ans(i) = T.row(i) * Mt.row(i)^T */
#pragma omp parallel for schedule(static)
    for (long i = 0; i < rows; i++) {
        double s = 0.0;
        for (long k = 0; k < n; k++) s += T[i + k * rows] * Mt[i + k * rows];
        out[i] = s;
    }
}

int eo_calculate_a_and_vara(const char *f_name_ascii, const double *selected_loci, long n_selected,
                            const double *inv_MMt_sqrt, const double *dim_reduced_vara,
                            double max_memory_in_Gbytes, const long *dims, const double *a,
                            double *out_a, double *out_vara, int *branch)
{
    const long L = dims[0], n = dims[1];
    const int have_sel = (n_selected > 0) && !eo_is_na(selected_loci[0]);

    /* :65  -- integer division of an integer product, then converted to double */
    double mem_bytes_needed = (double)((4UL * (unsigned long)n * (unsigned long)L * sizeof(double)) / 1000000000UL);

    double *ans_part1 = (double *)malloc(sizeof(double) * (size_t)n);
    double *W1 = (double *)malloc(sizeof(double) * (size_t)n * (size_t)n);
    double *W = (double *)malloc(sizeof(double) * (size_t)n * (size_t)n);
    if (!ans_part1 || !W1 || !W) { free(ans_part1); free(W1); free(W); return EO_ERR_ALLOC; }
    int rc = EO_OK;

    if (mem_bytes_needed < max_memory_in_Gbytes) { /* :74 */
        if (branch) *branch = 0;
        double *Mt = (double *)malloc(sizeof(double) * (size_t)L * (size_t)n);
        double *T = (double *)malloc(sizeof(double) * (size_t)L * (size_t)n);
        if (!Mt || !T) { free(Mt); free(T); rc = EO_ERR_ALLOC; goto done; }
        double t0 = eo_now();
        rc = eo_ReadBlock(f_name_ascii, 0, n, L, Mt); /* :76 */
        eo_stage_s[0] = eo_now() - t0;
        if (!rc) {
            if (have_sel) /* :79-84  Mt.row(sel).setZero() */
                for (long ii = 0; ii < n_selected; ii++) {
                    long r = (long)selected_loci[ii];
                    for (long k = 0; k < n; k++) Mt[r + k * L] = 0.0;
                }
            t0 = eo_now();
            eo_gemv(n, n, inv_MMt_sqrt, n, a, ans_part1);           /* :90 */
            eo_stage_s[2] = eo_now() - t0; t0 = eo_now();
            eo_gemv(L, n, Mt, L, ans_part1, out_a);                 /* :91 */
            eo_stage_s[6] = eo_now() - t0; t0 = eo_now();
            eo_gemm(n, n, n, dim_reduced_vara, n, inv_MMt_sqrt, n, 0, W1, n); /* :97 */
            eo_gemm(n, n, n, inv_MMt_sqrt, n, W1, n, 0, W, n);      /* :98 */
            eo_stage_s[3] = eo_now() - t0; t0 = eo_now();
            eo_gemm(L, n, n, Mt, L, W, n, 0, T, L);                 /* :103 */
            eo_stage_s[4] = eo_now() - t0; t0 = eo_now();
            eo_rowdots(L, n, T, Mt, out_vara);                      /* :107-112 */
            eo_stage_s[5] = eo_now() - t0;
        }
        free(Mt); free(T);
        goto done;
    }

    if (branch) *branch = 1;
    {
        /* :129-130 */
        long num_rows_in_block = (long)(max_memory_in_Gbytes * (1000000000) / (4 * n * sizeof(double)));
        if (num_rows_in_block < 0) { /* :133-144 soft failure: List(a=0, vara=0) */
            rc = EO_ERR_SOFT;
            goto done;
        }
        if (num_rows_in_block == 0) { rc = EO_ERR_BLOCK0; goto done; } /* :150 divides by it */
        long num_blocks = L / num_rows_in_block; /* :150-152 */
        if (L % num_rows_in_block) num_blocks++;
        double *Mt = (double *)malloc(sizeof(double) * (size_t)num_rows_in_block * (size_t)n);
        double *vt = (double *)malloc(sizeof(double) * (size_t)num_rows_in_block * (size_t)n);
        if (!Mt || !vt) { free(Mt); free(vt); rc = EO_ERR_ALLOC; goto done; }
        for (long i = 0; i < num_blocks && !rc; i++) { /* :157 */
            long start_row1 = i * num_rows_in_block;
            long nr1 = num_rows_in_block;
            if (start_row1 + nr1 > L) nr1 = L - start_row1;
            rc = eo_ReadBlock(f_name_ascii, start_row1, n, nr1, Mt); /* :165 */
            if (rc) break;
            if (have_sel) /* :176-190 */
                for (long ii = 0; ii < n_selected; ii++) {
                    if (selected_loci[ii] >= start_row1 && selected_loci[ii] < start_row1 + nr1) {
                        long r = (long)selected_loci[ii] - start_row1;
                        for (long k = 0; k < n; k++) Mt[r + k * nr1] = 0.0;
                    }
                }
            eo_gemv(n, n, inv_MMt_sqrt, n, a, ans_part1);          /* :192 (recomputed per block) */
            eo_gemv(nr1, n, Mt, nr1, ans_part1, out_a + start_row1); /* :193, :221-222 */
            eo_gemm(n, n, n, dim_reduced_vara, n, inv_MMt_sqrt, n, 0, W1, n); /* :197 */
            eo_gemm(n, n, n, inv_MMt_sqrt, n, W1, n, 0, W, n);     /* :198 */
            eo_gemm(nr1, n, n, Mt, nr1, W, n, 0, vt, nr1);         /* :204 */
            eo_rowdots(nr1, n, vt, Mt, out_vara + start_row1);     /* :210-216, :223 */
        }
        free(Mt); free(vt);
    }
done:
    free(ans_part1); free(W1); free(W);
    return rc;
}

/* ------------------------------------------------------------------ */
/* calculate_reduced_a_rcpp (src/calculate_reduced_a_rcpp.cpp:20-171)  */
/* ------------------------------------------------------------------ */
/* dims = (n, L): the dimensions of M (:37), although the file read is
 * Mt.ascii (:71 ReadBlock(f, 0, dims[0], dims[1])).  `sizeof(double)/1000000000`
 * is integer 0 (:56) so mem_bytes_needed == 0 and the in-memory branch (:65)
 * runs whenever max_memory_in_Gbytes > 0; the blocked branch mixes dims[0] and
 * dims[1] (:109, :120) and is not restated.
 * ar = varG * (Mt * (P * y))  (:82-84).  out: L doubles. */
int eo_calculate_reduced_a(const char *f_name_ascii, double varG, const double *P, const double *y,
                           double max_memory_in_Gbytes, const long *dims,
                           const double *selected_loci, long n_selected, double *out)
{
    const long n = dims[0], L = dims[1];
    const int have_sel = (n_selected > 0) && !eo_is_na(selected_loci[0]);
    double mem_bytes_needed = (double)(n * L + n * n + n) * (double)(sizeof(double) / (1000000000)); /* :56 */
    if (!(mem_bytes_needed < max_memory_in_Gbytes)) return EO_ERR_SOFT;
    double *Mt = (double *)malloc(sizeof(double) * (size_t)L * (size_t)n);
    double *py = (double *)malloc(sizeof(double) * (size_t)n);
    if (!Mt || !py) { free(Mt); free(py); return EO_ERR_ALLOC; }
    int rc = eo_ReadBlock(f_name_ascii, 0, n, L, Mt); /* :71 */
    if (!rc) {
        if (have_sel) /* :74-78 */
            for (long ii = 0; ii < n_selected; ii++) {
                long r = (long)selected_loci[ii];
                for (long k = 0; k < n; k++) Mt[r + k * L] = 0.0;
            }
        eo_gemv(n, n, P, n, y, py);       /* :82 */
        eo_gemv(L, n, Mt, L, py, out);    /* :83 */
        for (long i = 0; i < L; i++) out[i] = varG * out[i]; /* :84 */
    }
    free(Mt); free(py);
    return rc;
}

/* ------------------------------------------------------------------ */
/* extract_geno_rcpp (src/extract_geno_rcpp.cpp:17-86)                 */
/* ------------------------------------------------------------------ */
/* dims = (n, L); selected_locus 0-based column of M.  out: n ints in {-1,0,1}. */
int eo_extract_geno(const char *f_name_ascii, double max_memory_in_Gbytes, long selected_locus,
                    const long *dims, int *out, int *branch)
{
    const long n = dims[0], L = dims[1];
    double memory_needed_in_Gb = ((double)n * L * sizeof(double)) / 1000000000.0; /* :35-36 */
    if (max_memory_in_Gbytes > memory_needed_in_Gb) { /* :44 */
        if (branch) *branch = 0;
        double *genoMat = (double *)malloc(sizeof(double) * (size_t)n * (size_t)L);
        if (!genoMat) return EO_ERR_ALLOC;
        int rc = eo_ReadBlock(f_name_ascii, 0, L, n, genoMat); /* :46 */
        if (!rc)
            for (long i = 0; i < n; i++) out[i] = (int)genoMat[i + selected_locus * n]; /* :48 */
        free(genoMat);
        return rc;
    }
    if (branch) *branch = 1;
    long num_rows_in_block = (long)((max_memory_in_Gbytes * 1000000000.0) / (sizeof(double) * L)); /* :53 */
    if (num_rows_in_block <= 0) return EO_ERR_BLOCK0;
    long num_blocks = n / num_rows_in_block; /* :55-57 */
    if (n % num_rows_in_block) num_blocks++;
    double *blk = (double *)malloc(sizeof(double) * (size_t)num_rows_in_block * (size_t)L);
    if (!blk) return EO_ERR_ALLOC;
    int rc = EO_OK;
    for (long i = 0; i < num_blocks && !rc; i++) { /* :60 */
        long start_row1 = i * num_rows_in_block;
        long nr1 = num_rows_in_block;
        if (start_row1 + nr1 > n) nr1 = n - start_row1;
        rc = eo_ReadBlock(f_name_ascii, start_row1, L, nr1, blk); /* :67 */
        if (rc) break;
        for (long j = start_row1; j < start_row1 + nr1; j++) /* :72-77 */
            out[j] = (int)blk[(j - start_row1) + selected_locus * nr1];
    }
    free(blk);
    return rc;
}

/* ------------------------------------------------------------------ */
/* ingest: createM_ASCII_rcpp (text files) and createMt_ASCII_rcpp     */
/* ------------------------------------------------------------------ */
/* message(...) calls are collected into msgbuf, separated by 0x1e; R's message() pastes its arguments
 * without separators. */
#include <stdarg.h>
static void eo_msg(char *buf, long cap, const char *fmt, ...)
{
    if (!buf || cap <= 0) return;
    size_t used = strlen(buf);
    if (used && used + 1 < (size_t)cap) { buf[used++] = '\x1e'; buf[used] = 0; }
    if (used + 1 >= (size_t)cap) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf + used, (size_t)cap - used, fmt, ap);
    va_end(ap);
}
/* what operator>> (std::string) skips: isspace in the "C" locale */
static int eo_isspace(int c) { return c == ' ' || (c >= 9 && c <= 13); }

/* CreateASCIInospace (src/CreateASCIInospace.cpp:17-164).  dims = (rows, columns) of the text file. */
static int eo_CreateASCIInospace(const char *fname, const char *asciifname, const long *dims, const char *AA,
                                 const char *AB, const char *BB, int quiet, const char *missing, char *msgbuf,
                                 long msgcap, int *ok)
{
    *ok = 0;
    FILE *in = fopen(fname, "r");
    if (!in) { /* :38-41 */
        eo_msg(msgbuf, msgcap, "ERROR: Text file could not be opened with filename  %s\n", fname);
        return EO_OK;
    }
    FILE *out = fopen(asciifname, "w"); /* :43 */
    if (!out) { fclose(in); return EO_ERR_OPEN; }
    if (!quiet) { /* :44-50 */
        eo_msg(msgbuf, msgcap, "%s", "");
        eo_msg(msgbuf, msgcap, " Reading text File  ");
        eo_msg(msgbuf, msgcap, "%s", "");
        eo_msg(msgbuf, msgcap, " Loading file ");
    }
    char *rowinfile = (char *)malloc((size_t)dims[1] + 1); /* :58-63 */
    char *line = NULL;
    size_t cap = 0;
    ssize_t len;
    long counter = 0;
    int rc = EO_OK, failed = 0;
    if (!rowinfile) { fclose(in); fclose(out); return EO_ERR_ALLOC; }
    memset(rowinfile, '0', (size_t)dims[1]);
    while (!failed && (len = getline(&line, &cap, in)) >= 0) { /* :67 */
        if (len > 0 && line[len - 1] == '\n') len--;
        long i = 0, number_of_columns = 0;
        ssize_t p = 0;
        for (;;) { /* :79  while (streamA >> token) */
            while (p < len && eo_isspace((unsigned char)line[p])) p++;
            if (p >= len) break;
            ssize_t t0 = p;
            while (p < len && !eo_isspace((unsigned char)line[p])) p++;
            size_t tl = (size_t)(p - t0);
            const char *tok = line + t0;
            number_of_columns++;
            char code;
#define EO_TOKEQ(s) (strlen(s) == tl && memcmp(tok, (s), tl) == 0)
            if (EO_TOKEQ(BB)) code = '2';            /* :84 */
            else if (EO_TOKEQ(AB)) code = '1';       /* :86 */
            else if (EO_TOKEQ(AA)) code = '0';       /* :88 */
            else if (EO_TOKEQ(missing)) code = '1';  /* :90-92 missing -> het */
            else {                                   /* :93-105 */
                if (strcmp(AB, "NA") == 0)
                    eo_msg(msgbuf, msgcap, "\n Marker file contains marker genotypes that are different to AA=%s BB=%s", AA, BB);
                else
                    eo_msg(msgbuf, msgcap, "\n Marker file contains marker genotypes that are different to AA=%s AB=%s BB=%s", AA, AB, BB);
                eo_msg(msgbuf, msgcap, " For example , %.*s in row %ld", (int)tl, tok, counter + 1);
                eo_msg(msgbuf, msgcap, "\n ReadMarker has terminated with errors\n");
                failed = 1;
                break;
            }
#undef EO_TOKEQ
            if (i < dims[1]) rowinfile[i] = code; /* the reference writes out of bounds for i >= dims[1] (:85) */
            i++;
        }
        if (failed) break;
        if (number_of_columns != dims[1]) { /* :108-116 */
            eo_msg(msgbuf, msgcap, "\n");
            eo_msg(msgbuf, msgcap, "Error:  Marker text file contains an unequal number of columns per row.  ");
            eo_msg(msgbuf, msgcap, "        The error has occurred at row %ld which contains %ld but ", counter + 1, number_of_columns);
            eo_msg(msgbuf, msgcap, "        it should contain %ld columns of data. ", dims[1]);
            eo_msg(msgbuf, msgcap, "\n");
            eo_msg(msgbuf, msgcap, " ReadMarkerData has terminated with errors");
            failed = 1;
            break;
        }
        fwrite(rowinfile, 1, (size_t)number_of_columns, out); /* :118-121 */
        fputc('\n', out);
        counter++;
    }
    if (!failed) { /* :129-157 echo of the first lines */
        int nrowsp = dims[0] < 5 ? (int)dims[0] : 5, ncolsp = dims[1] < 12 ? (int)dims[1] : 12;
        eo_msg(msgbuf, msgcap, " First %d lines and %d columns of the marker text  file. ", nrowsp, ncolsp);
        rewind(in);
        char tmp[256] = "";
        long c2 = 0;
        while (c2 < nrowsp && (len = getline(&line, &cap, in)) >= 0) {
            if (len > 0 && line[len - 1] == '\n') len--;
            char rowline[4096] = "";
            ssize_t p = 0;
            for (int i = 0; i < ncolsp; i++) {
                while (p < len && eo_isspace((unsigned char)line[p])) p++;
                ssize_t t0 = p;
                while (p < len && !eo_isspace((unsigned char)line[p])) p++;
                if (p > t0) snprintf(tmp, sizeof(tmp), "%.*s", (int)(p - t0), line + t0); /* failed >> keeps tmp */
                strncat(rowline, tmp, sizeof(rowline) - strlen(rowline) - 2);
                strcat(rowline, " ");
            }
            eo_msg(msgbuf, msgcap, "%s", rowline);
            c2++;
        }
        *ok = 1;
    }
    free(line); free(rowinfile);
    fclose(in); fclose(out);
    return rc;
}

/* CreateASCIInospace_PLINK (src/CreateASCIInospace_PLINK.cpp:16-248).  dims = (rows, 6 + 2 * nsnp). */
static int eo_CreateASCIInospace_PLINK(const char *fname, const char *asciifname, const long *dims, int quiet,
                                       char *msgbuf, long msgcap, int *ok)
{
    (void)quiet;
    *ok = 0;
    const long n_of_cols_in_geno = (long)((dims[1] - 6) / 2.0); /* :20 */
    FILE *in = fopen(fname, "r");
    if (!in) { /* :38-42 */
        eo_msg(msgbuf, msgcap, "ERROR: PLINK ped file could not be opened with filename  %s", fname);
        eo_msg(msgbuf, msgcap, "ERROR: ReadMarkerData has terminated with errors.  ");
        return EO_OK;
    }
    FILE *out = fopen(asciifname, "w");
    if (!out) { fclose(in); return EO_ERR_OPEN; }
    long nrv = dims[1] - 6 > 0 ? dims[1] - 6 : 0;
    char *alleles0 = (char *)malloc((size_t)n_of_cols_in_geno + 1), *alleles1 = (char *)malloc((size_t)n_of_cols_in_geno + 1);
    char *rowvec = (char *)calloc((size_t)nrv + 1, 1), *rowinfile = (char *)malloc((size_t)n_of_cols_in_geno + 2);
    char *line = NULL;
    size_t cap = 0;
    ssize_t len;
    long counter = 0;
    int printOnlyOnce = 0, failed = 0;
    while (!failed && (len = getline(&line, &cap, in)) >= 0) { /* :55 */
        if (len > 0 && line[len - 1] == '\n') len--;
        memset(rowinfile, '0', (size_t)n_of_cols_in_geno);
        long numcols = 0; /* :64-67 number of whitespace-separated tokens */
        ssize_t p = 0;
        for (;;) {
            while (p < len && eo_isspace((unsigned char)line[p])) p++;
            if (p >= len) break;
            while (p < len && !eo_isspace((unsigned char)line[p])) p++;
            numcols++;
        }
        if (numcols != dims[1]) { /* :69-76 */
            eo_msg(msgbuf, msgcap, "\n");
            eo_msg(msgbuf, msgcap, "Error:  PLINK file contains an unequal number of columns per row.  ");
            eo_msg(msgbuf, msgcap, "        The error has occurred at row %ld which contains %ld but ", counter + 1, numcols);
            eo_msg(msgbuf, msgcap, "        it should contain %ld columns of data. ", dims[1]);
            eo_msg(msgbuf, msgcap, "\n");
            eo_msg(msgbuf, msgcap, " ReadMarkerData has terminated with errors");
            failed = 1;
            break;
        }
        p = 0;
        for (int i = 0; i <= 5; i++) { /* :87-89 six string extractions */
            while (p < len && eo_isspace((unsigned char)line[p])) p++;
            while (p < len && !eo_isspace((unsigned char)line[p])) p++;
        }
        for (long i = 6; i < dims[1]; i++) { /* :90-92  operator>>(char&): ONE non-blank character each */
            while (p < len && eo_isspace((unsigned char)line[p])) p++;
            if (p < len) rowvec[i - 6] = line[p++]; /* a failed extraction leaves the element as it was */
        }
        if (counter == 0) { /* :99-111 */
            for (long i = 0; i < n_of_cols_in_geno; i++) {
                char a = rowvec[2 * i], b = rowvec[2 * i + 1];
                if (a == '0' || b == '0' || a == '-' || b == '-') { alleles0[i] = 'I'; alleles1[i] = 'I'; }
                else { alleles0[i] = a; alleles1[i] = b; }
            }
        }
        for (long i = 0; i < n_of_cols_in_geno && !failed; i++) { /* :115-196 */
            if (rowvec[2 * i] == '0' || rowvec[2 * i + 1] == '0' || rowvec[2 * i] == '-' || rowvec[2 * i + 1] == '-') {
                if (printOnlyOnce == 0) { /* :119-129 */
                    eo_msg(msgbuf, msgcap, "\n");
                    eo_msg(msgbuf, msgcap, " Warning:  PLINK file contains missing alleles (i.e. 0 or - ) ");
                    eo_msg(msgbuf, msgcap, "           These missing genotypes should be imputed before running Eagle.");
                    eo_msg(msgbuf, msgcap, "           As an approximation, AMpus has set these missing genotypes to heterozygotes. ");
                    eo_msg(msgbuf, msgcap, "           Since Eagle assumes an additive model, heterozygote genotypes do not contribute to the estimation of ");
                    eo_msg(msgbuf, msgcap, "           the additive effects.  ");
                    eo_msg(msgbuf, msgcap, "\n");
                    printOnlyOnce = 1;
                }
                rowvec[2 * i] = 'I';
                rowvec[2 * i + 1] = 'I';
            }
            for (int j = 1; j >= 0; --j) { /* :137 second allele first */
                char x = rowvec[2 * i + j];
                if (x != alleles0[i] && x != alleles1[i]) {
                    if (x == 'I') { /* nothing */ }
                    else if (alleles0[i] == 'I') alleles0[i] = x;
                    else if (alleles1[i] == 'I') alleles1[i] = x;
                    else if (alleles0[i] == alleles1[i]) alleles1[i] = x;
                    else { /* :160-167 */
                        eo_msg(msgbuf, msgcap, "\n");
                        eo_msg(msgbuf, msgcap, "Error:  PLINK file cannot contain more than two alleles at a locus.");
                        eo_msg(msgbuf, msgcap, "        The error has occurred at snp locus %ld for individual %ld", i + 1, counter + 1);
                        eo_msg(msgbuf, msgcap, "\n");
                        eo_msg(msgbuf, msgcap, " ReadMarkerData has terminated with errors");
                        failed = 1;
                        break;
                    }
                }
                /* :177-190, evaluated in both passes of j; the pass j == 0 decides */
                if (rowvec[2 * i] == 'I' || rowvec[2 * i + 1] == 'I') rowinfile[i] = '1';
                else if (rowvec[2 * i + 1] != rowvec[2 * i]) rowinfile[i] = '1';
                else if (rowvec[2 * i] == alleles0[i]) rowinfile[i] = '0';
                else rowinfile[i] = '2';
            }
        }
        if (failed) break;
        fwrite(rowinfile, 1, (size_t)n_of_cols_in_geno, out); /* :198-199 */
        fputc('\n', out);
        counter++;
    }
    if (!failed) { /* :203-236 */
        int nrowsp = dims[0] < 5 ? (int)dims[0] : 5, ncolsp = dims[1] < 25 ? (int)dims[1] : 24;
        eo_msg(msgbuf, msgcap, " First %d lines and %d columns of the PLINK ped file. ", nrowsp, ncolsp);
        rewind(in);
        char tmp[256] = "";
        long c2 = 0;
        while (c2 < nrowsp && (len = getline(&line, &cap, in)) >= 0) {
            if (len > 0 && line[len - 1] == '\n') len--;
            char rowline[8192] = "";
            ssize_t p = 0;
            for (int i = 0; i < ncolsp; i++) {
                while (p < len && eo_isspace((unsigned char)line[p])) p++;
                ssize_t t0 = p;
                while (p < len && !eo_isspace((unsigned char)line[p])) p++;
                if (p > t0) snprintf(tmp, sizeof(tmp), "%.*s", (int)(p - t0), line + t0);
                strncat(rowline, tmp, sizeof(rowline) - strlen(rowline) - 2);
                strcat(rowline, " ");
            }
            eo_msg(msgbuf, msgcap, "%s", rowline);
            c2++;
        }
        *ok = 1;
    }
    free(line); free(alleles0); free(alleles1); free(rowvec); free(rowinfile);
    fclose(in); fclose(out);
    return EO_OK;
}

/* createM_ASCII_rcpp (src/createM_ASCII_rcpp.cpp:19-106) */
int eo_createM_ASCII(const char *f_name, const char *f_name_ascii, const char *type, const char *AA, const char *AB,
                     const char *BB, double max_memory_in_Gbytes, const long *dims, int quiet, const char *missing,
                     char *msgbuf, long msgcap, int *ok)
{
    (void)max_memory_in_Gbytes; /* :88-96: both branches call the same routine */
    if (msgbuf && msgcap > 0) msgbuf[0] = 0;
    if (strcmp(type, "PLINK") == 0) /* :71-78 */
        return eo_CreateASCIInospace_PLINK(f_name, f_name_ascii, dims, quiet, msgbuf, msgcap, ok);
    if (!quiet) eo_msg(msgbuf, msgcap, " A text file is being assumed as the input data file type. "); /* :85-86 */
    return eo_CreateASCIInospace(f_name, f_name_ascii, dims, AA, AB, BB, quiet, missing, msgbuf, msgcap, ok);
}

/* createMt_ASCII_rcpp (src/createMt_ASCII_rcpp.cpp:15-245).  dims = (n, L) of M.ascii.  Both situations (:66-122 in
 * memory, :123-218 column blocks) write the same bytes: line c of the output is column c of the input. */
int eo_createMt_ASCII(const char *f_name, const char *f_name_ascii, const char *type, double max_memory_in_Gbytes,
                      const long *dims, int quiet, char *msgbuf, long msgcap)
{
    const long n = dims[0], L = dims[1];
    if (msgbuf && msgcap > 0) msgbuf[0] = 0;
    const double max_mem_in_bytes = max_memory_in_Gbytes * 1000000000; /* :43-44 */
    const double mem_bytes = 3.5 * n * L * (31 / 8);                   /* :50  bits_in_int = 31, integer division */
    FILE *in = fopen(f_name, "r");
    if (!in) return EO_ERR_OPEN; /* :72-75, :150-153 */
    FILE *out = fopen(f_name_ascii, "w");
    if (!out) { fclose(in); return EO_ERR_OPEN; }
    long ncols_block = L, n_blocks = 1;
    if (!(mem_bytes < max_mem_in_bytes)) { /* :123-141 */
        if (!quiet) {
            eo_msg(msgbuf, msgcap, " A block transpose is being performed due to lack of memory.  ");
            eo_msg(msgbuf, msgcap, " Memory parameter availmemGb is set to %ggigabytes", max_memory_in_Gbytes);
            eo_msg(msgbuf, msgcap, " If possible, increase availmemGb parameter. ");
        }
        ncols_block = (long)(max_mem_in_bytes * 1.0 / (3.5 * n * (31 / 8.0)));
        if (ncols_block <= 0) { fclose(in); fclose(out); return EO_ERR_BLOCK0; }
        n_blocks = L / ncols_block;
        if (L % ncols_block != 0) n_blocks++;
        if (!quiet) {
            eo_msg(msgbuf, msgcap, " Block Transpose of ASCII genotype file beginning ... ");
            eo_msg(msgbuf, msgcap, "  Due to marker data exceeding memory, data being processed in blocks. Number of blocks being processed is %ld", n_blocks);
        }
    }
    char *M = (char *)malloc((size_t)n * (size_t)ncols_block), *row = (char *)malloc((size_t)n + 1);
    char *line = NULL;
    size_t cap = 0;
    int rc = (M && row) ? EO_OK : EO_ERR_ALLOC;
    for (long b = 0; b < n_blocks && !rc; b++) {
        if (n_blocks > 1 || !(mem_bytes < max_mem_in_bytes)) {
            if (!quiet) eo_msg(msgbuf, msgcap, " Processing block ... %ld of a total number of blocks of %ld", b, n_blocks);
            if (!quiet) eo_msg(msgbuf, msgcap, "\n\n");
        }
        long start_val = b * ncols_block, end_val = (b + 1) * ncols_block;
        if (end_val > L) end_val = L;
        long nc = end_val - start_val;
        rewind(in);
        for (long r = 0; r < n && !rc; r++) {
            ssize_t len = getline(&line, &cap, in);
            if (len < 0 || len < end_val) { rc = EO_ERR_SHORT; break; }
            memcpy(M + (size_t)r * nc, line + start_val, (size_t)nc);
        }
        for (long c = 0; c < nc && !rc; c++) {
            for (long r = 0; r < n; r++) row[r] = M[(size_t)r * nc + c];
            row[n] = '\n';
            fwrite(row, 1, (size_t)n + 1, out);
        }
    }
    free(line); free(M); free(row);
    fclose(in); fclose(out);
    if (rc) return rc;
    /* :224-243 */
    eo_msg(msgbuf, msgcap, "\n\n                    Summary of Marker File  ");
    eo_msg(msgbuf, msgcap, "                   ~~~~~~~~~~~~~~~~~~~~~~~~   ");
    eo_msg(msgbuf, msgcap, " File type:                   %s", type);
    eo_msg(msgbuf, msgcap, " Reformatted ASCII file name:  %s", f_name);
    eo_msg(msgbuf, msgcap, " Number of individuals:        %ld", n);
    eo_msg(msgbuf, msgcap, " Number of loci:               %ld", L);
    eo_msg(msgbuf, msgcap, " File size (gigabytes):       %g", mem_bytes / 1000000000);
    eo_msg(msgbuf, msgcap, " Available memory (gigabytes): %g", max_memory_in_Gbytes);
    eo_msg(msgbuf, msgcap, "\n\n");
    eo_msg(msgbuf, msgcap, " The marker file has been Uploaded");
    return EO_OK;
}

/* ReshapeM_rcpp (src/ReshapeM_rcpp.cpp:16-117).  indxNA 0-based.  Writes <fnameM>tmp and <fnameMt>tmp. */
int eo_ReshapeM(const char *fnameM, const char *fnameMt, const long *indxNA, long n_indx, const long *dims, long *newdims)
{
    (void)dims;
    newdims[0] = newdims[1] = 0;
    FILE *in = fopen(fnameM, "r");
    if (!in) return EO_ERR_OPEN; /* :44-47 */
    size_t ln = strlen(fnameM);
    char *outname = (char *)malloc(ln + 4);
    memcpy(outname, fnameM, ln); memcpy(outname + ln, "tmp", 4); /* :51 */
    FILE *out = fopen(outname, "w");
    free(outname);
    if (!out) { fclose(in); return EO_ERR_OPEN; }
    char *line = NULL;
    size_t cap = 0;
    ssize_t len;
    long rownum = 0;
    while ((len = getline(&line, &cap, in)) >= 0) { /* :58-75 */
        if (len > 0 && line[len - 1] == '\n') len--;
        int writeline = 1;
        for (long ii = 0; ii < n_indx; ii++)
            if (indxNA[ii] == rownum) writeline = 0;
        if (writeline) {
            fwrite(line, 1, (size_t)len, out);
            fputc('\n', out);
            newdims[0]++;
        }
        rownum++;
        newdims[1] = (long)len;
    }
    fclose(in); fclose(out);
    in = fopen(fnameMt, "r");
    if (!in) { free(line); return EO_ERR_OPEN; } /* :86-89 */
    ln = strlen(fnameMt);
    outname = (char *)malloc(ln + 4);
    memcpy(outname, fnameMt, ln); memcpy(outname + ln, "tmp", 4); /* :93 */
    out = fopen(outname, "w");
    free(outname);
    if (!out) { fclose(in); free(line); return EO_ERR_OPEN; }
    int rc = EO_OK;
    while (!rc && (len = getline(&line, &cap, in)) >= 0) { /* :99-110 */
        if (len > 0 && line[len - 1] == '\n') len--;
        for (long ii = 0; ii < n_indx; ii++) { /* line.erase(indxNA[ii], 1), one after the other */
            if (indxNA[ii] < 0 || indxNA[ii] > len) { rc = EO_ERR_SHORT; break; } /* std::out_of_range in the reference */
            if (indxNA[ii] < len) {
                memmove(line + indxNA[ii], line + indxNA[ii] + 1, (size_t)(len - indxNA[ii] - 1));
                len--;
            }
        }
        if (rc) break;
        fwrite(line, 1, (size_t)len, out);
        fputc('\n', out);
    }
    free(line);
    fclose(in); fclose(out);
    return rc;
}

/* getRowColumn (src/getRowColumn.cpp:19-72) */
int eo_getRowColumn(const char *fname, long *dimen)
{
    dimen[0] = dimen[1] = 0;
    FILE *in = fopen(fname, "r");
    if (!in) return EO_ERR_OPEN; /* :36-39 */
    char *line = NULL;
    size_t cap = 0;
    ssize_t len;
    while ((len = getline(&line, &cap, in)) >= 0) dimen[0]++; /* :43-47 */
    rewind(in); /* :51-52 */
    len = getline(&line, &cap, in);
    if (len > 0 && line[len - 1] == '\n') len--;
    for (ssize_t p = 0; p < len;) { /* :58-65 */
        while (p < len && eo_isspace((unsigned char)line[p])) p++;
        if (p >= len) break;
        while (p < len && !eo_isspace((unsigned char)line[p])) p++;
        dimen[1]++;
    }
    free(line);
    fclose(in);
    return EO_OK;
}

int eo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
