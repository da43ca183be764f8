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

/* C(m x n) = A(m x k) * op(B);  transb==0: B is k x n;  transb==1: B is n x k (C = A*B^T) */
static void eo_gemm(long m, long n, long k, const double *A, long lda, const double *B, long ldb,
                    int transb, double *C, long ldc)
{
    const long MB = 256, KB = 256;
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
        int rc = eo_ReadBlock(f_name_ascii, 0, L, n, genoMat); /* :86 */
        if (rc) { free(genoMat); return rc; }
        if (have_sel) /* :88-92 */
            for (long ii = 0; ii < n_selected; ii++)
                memset(genoMat + (long)selected_loci[ii] * n, 0, sizeof(double) * (size_t)n);
        eo_gemm(n, n, L, genoMat, n, genoMat, n, 1, MMt, n); /* :95 */
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
        rc = eo_ReadBlock(f_name_ascii, 0, n, L, Mt); /* :76 */
        if (!rc) {
            if (have_sel) /* :79-84  Mt.row(sel).setZero() */
                for (long ii = 0; ii < n_selected; ii++) {
                    long r = (long)selected_loci[ii];
                    for (long k = 0; k < n; k++) Mt[r + k * L] = 0.0;
                }
            eo_gemv(n, n, inv_MMt_sqrt, n, a, ans_part1);           /* :90 */
            eo_gemv(L, n, Mt, L, ans_part1, out_a);                 /* :91 */
            eo_gemm(n, n, n, dim_reduced_vara, n, inv_MMt_sqrt, n, 0, W1, n); /* :97 */
            eo_gemm(n, n, n, inv_MMt_sqrt, n, W1, n, 0, W, n);      /* :98 */
            eo_gemm(L, n, n, Mt, L, W, n, 0, T, L);                 /* :103 */
            eo_rowdots(L, n, T, Mt, out_vara);                      /* :107-112 */
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

int eo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
