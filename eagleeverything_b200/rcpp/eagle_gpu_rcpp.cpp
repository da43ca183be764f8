// Rcpp glue for libeaglegpu.so -- replaces the BODIES of five files in MyPackage/Eagle/src/
// (ReadBlock.cpp, calculateMMt_rcpp.cpp, calculate_a_and_vara_rcpp.cpp,
// calculate_reduced_a_rcpp.cpp, extract_geno_rcpp.cpp) and, optionally, of the two ingest exports
// (createM_ASCII_rcpp.cpp, createMt_ASCII_rcpp.cpp).  The exported prototypes are unchanged,
// so RcppExports.cpp / RcppExports.R / NAMESPACE / every R file stay byte-identical and AM(),
// SummaryAM() etc. run untouched (reference: src/RcppExports.cpp:9, 37, 54, 73, 129 and the
// registration table :154-170).
//
// NOT compiled in this repository: R, Rcpp and RcppEigen are not installed in the build image.
// It is the thin, mechanical layer INTEGRATION.md describes; everything it calls is the C ABI in
// include/eagle_gpu.h, which IS built and tested here (tests/ call the same entry points through
// ctypes with the same arguments).
//
// Build: see Makevars in this directory (replaces MyPackage/Makevars.gpu + nvblas.conf).

// [[Rcpp::depends(RcppEigen)]]
#include <RcppEigen.h>

#include "eagle_gpu.h"
#include "readblock.h"

namespace {

// message(...) of the reference is an R closure; the shim calls back on the calling (R main) thread only
void r_message(void* ctx, const char* text) {
    Rcpp::Function* f = static_cast<Rcpp::Function*>(ctx);
    (*f)(text);
}

inline void check(int rc) {
    if (rc != EG_OK) Rcpp::stop(eg_last_error());  // -> R error through END_RCPP, as ReadBlock.cpp:42-45 does
}

inline std::vector<int64_t> dims64(const std::vector<long>& d) { return std::vector<int64_t>(d.begin(), d.end()); }

}  // namespace

// [[Rcpp::export]]
Eigen::MatrixXd ReadBlock(std::string asciifname, long start_row, long numcols, long numrows_in_block) {
    Eigen::MatrixXd M(numrows_in_block, numcols);  // column-major, the layout eg_ReadBlock writes
    check(eg_ReadBlock(asciifname.c_str(), start_row, numcols, numrows_in_block, M.data()));
    return M;
}

// [[Rcpp::export]]
Eigen::MatrixXd calculateMMt_rcpp(Rcpp::CharacterVector f_name_ascii, double max_memory_in_Gbytes, int num_cores,
                                  Rcpp::NumericVector selected_loci, std::vector<long> dims, bool quiet,
                                  Rcpp::Function message) {
    std::string fname = Rcpp::as<std::string>(f_name_ascii);
    std::vector<int64_t> d = dims64(dims);
    Eigen::MatrixXd MMt(dims[0], dims[0]);
    // NA_real_ in selected_loci(0) is passed through untouched: the shim tests it like R_IsNA
    check(eg_calculateMMt_rcpp(fname.c_str(), max_memory_in_Gbytes, num_cores, selected_loci.begin(),
                               selected_loci.size(), d.data(), quiet, r_message, &message, MMt.data()));
    return MMt;
}

// [[Rcpp::export]]
Rcpp::List calculate_a_and_vara_rcpp(Rcpp::CharacterVector f_name_ascii, Rcpp::NumericVector selected_loci,
                                     Eigen::Map<Eigen::MatrixXd> inv_MMt_sqrt,
                                     Eigen::Map<Eigen::MatrixXd> dim_reduced_vara, double max_memory_in_Gbytes,
                                     std::vector<long> dims, Eigen::VectorXd a, bool quiet, Rcpp::Function message) {
    std::string fname = Rcpp::as<std::string>(f_name_ascii);
    std::vector<int64_t> d = dims64(dims);
    Eigen::MatrixXd ans(dims[0], 1), var_ans(dims[0], 1);
    // the Eigen::Map arguments are zero-copy views of the R matrices; the shim only reads them
    check(eg_calculate_a_and_vara_rcpp(fname.c_str(), selected_loci.begin(), selected_loci.size(),
                                       inv_MMt_sqrt.data(), dim_reduced_vara.data(), max_memory_in_Gbytes, d.data(),
                                       a.data(), quiet, r_message, &message, ans.data(), var_ans.data()));
    return Rcpp::List::create(Rcpp::Named("a") = ans, Rcpp::Named("vara") = var_ans);
}

// [[Rcpp::export]]
Eigen::MatrixXd calculate_reduced_a_rcpp(Rcpp::CharacterVector f_name_ascii, double varG,
                                         Eigen::Map<Eigen::MatrixXd> P, Eigen::Map<Eigen::MatrixXd> y,
                                         double max_memory_in_Gbytes, std::vector<long> dims,
                                         Rcpp::NumericVector selected_loci, bool quiet, Rcpp::Function message) {
    std::string fname = Rcpp::as<std::string>(f_name_ascii);
    std::vector<int64_t> d = dims64(dims);
    Eigen::MatrixXd ar(dims[1], 1);
    check(eg_calculate_reduced_a_rcpp(fname.c_str(), varG, P.data(), y.data(), max_memory_in_Gbytes, d.data(),
                                      selected_loci.begin(), selected_loci.size(), quiet, r_message, &message,
                                      ar.data()));
    return ar;
}

// [[Rcpp::export]]
Eigen::VectorXi extract_geno_rcpp(Rcpp::CharacterVector f_name_ascii, double max_memory_in_Gbytes,
                                  long selected_locus, std::vector<long> dims) {
    std::string fname = Rcpp::as<std::string>(f_name_ascii);
    std::vector<int64_t> d = dims64(dims);
    Eigen::VectorXi column_of_genos(dims[0]);
    check(eg_extract_geno_rcpp(fname.c_str(), max_memory_in_Gbytes, selected_locus, d.data(), column_of_genos.data()));
    return column_of_genos;
}

// ---------------------------------------------------------------------------------------------------
// the two ingest exports ReadMarker() calls (R/ReadMarker.R), same prototypes
// as src/createM_ASCII_rcpp.cpp:19-29 and src/createMt_ASCII_rcpp.cpp:15-19 (registered with 11 and 7 arguments,
// RcppExports.cpp:154-170).  Replacing their bodies removes CreateASCIInospace.cpp and CreateASCIInospace_PLINK.cpp
// from the build as well.
// ---------------------------------------------------------------------------------------------------

// [[Rcpp::export]]
bool createM_ASCII_rcpp(Rcpp::CharacterVector f_name, Rcpp::CharacterVector f_name_ascii, Rcpp::CharacterVector type,
                        std::string AA, std::string AB, std::string BB, double max_memory_in_Gbytes, std::vector<long> dims,
                        bool quiet, Rcpp::Function message, std::string missing) {
    std::string fname = Rcpp::as<std::string>(f_name), fascii = Rcpp::as<std::string>(f_name_ascii),
                ftype = Rcpp::as<std::string>(type);
    std::vector<int64_t> d = dims64(dims);
    int ok = 0;
    check(eg_createM_ASCII_rcpp(fname.c_str(), fascii.c_str(), ftype.c_str(), AA.c_str(), AB.c_str(), BB.c_str(),
                                max_memory_in_Gbytes, d.data(), quiet, r_message, &message, missing.c_str(), &ok));
    return ok != 0;  // false after the reference's own messages (bad token, ragged row, third allele, unreadable file)
}

// [[Rcpp::export]]
void createMt_ASCII_rcpp(Rcpp::CharacterVector f_name, Rcpp::CharacterVector f_name_ascii, Rcpp::CharacterVector type,
                         double max_memory_in_Gbytes, std::vector<long> dims, bool quiet, Rcpp::Function message) {
    std::string fname = Rcpp::as<std::string>(f_name), fascii = Rcpp::as<std::string>(f_name_ascii),
                ftype = Rcpp::as<std::string>(type);
    std::vector<int64_t> d = dims64(dims);
    check(eg_createMt_ASCII_rcpp(fname.c_str(), fascii.c_str(), ftype.c_str(), max_memory_in_Gbytes, d.data(), quiet, r_message,
                                 &message));
}

// [[Rcpp::export]]
std::vector<long> ReshapeM_rcpp(Rcpp::CharacterVector fnameM, Rcpp::CharacterVector fnameMt, std::vector<long> indxNA,
                                std::vector<long> dims) {
    std::string fM = Rcpp::as<std::string>(fnameM), fMt = Rcpp::as<std::string>(fnameMt);
    std::vector<int64_t> d = dims64(dims), idx(indxNA.begin(), indxNA.end());
    int64_t nd[2] = {0, 0};
    check(eg_ReshapeM_rcpp(fM.c_str(), fMt.c_str(), idx.data(), (int64_t)idx.size(), d.data(), nd));
    return std::vector<long>{(long)nd[0], (long)nd[1]};
}

// [[Rcpp::export]]
std::vector<long> getRowColumn(std::string fname) {
    int64_t nd[2] = {0, 0};
    check(eg_getRowColumn(fname.c_str(), nd));
    return std::vector<long>{(long)nd[0], (long)nd[1]};
}

// ---------------------------------------------------------------------------------------------------
// the n x n algebra between two scans.  These are NEW exports (the
// reference does this work in R: R/calculateMMt_sqrt_and_sqrtinv.R:26-31, R/calculateH.R:36, R/calculateP.R:27-28,
// R/calculate_reduced_a.R:31, R/calculate_reduced_vara.R:21-35).  Each R function keeps its name, arguments,
// checks and messages and replaces only its arithmetic by one call, e.g. in calculateP.R:
//     P <- calculateP_gpu(H, X)            instead of lines 27-28
// so AM() and find_qtl() stay untouched.
// ---------------------------------------------------------------------------------------------------

// [[Rcpp::export]]
Rcpp::List calculateMMt_sqrt_and_sqrtinv_gpu(Eigen::Map<Eigen::MatrixXd> MMt, bool checkres, Rcpp::Function message) {
    const long n = MMt.rows();
    Eigen::MatrixXd sq(n, n), inv(n, n);
    int ok = 0;
    check(eg_calculateMMt_sqrt_and_sqrtinv(MMt.data(), n, checkres, r_message, &message, sq.data(), inv.data(), &ok));
    if (!ok) return R_NilValue;  // the messages of calculateMMt_sqrt_and_sqrtinv.R:15-21 were already emitted
    return Rcpp::List::create(Rcpp::Named("sqrt_MMt") = sq, Rcpp::Named("inverse_sqrt_MMt") = inv);
}

// [[Rcpp::export]]
Eigen::MatrixXd calculateP_gpu(Eigen::Map<Eigen::MatrixXd> H, Eigen::Map<Eigen::MatrixXd> X) {
    Eigen::MatrixXd P(H.rows(), H.rows());
    check(eg_calculateP(H.data(), X.data(), H.rows(), (int)X.cols(), P.data()));
    return P;
}

// [[Rcpp::export]]
Eigen::MatrixXd calculate_reduced_a_gpu(double varG, Eigen::Map<Eigen::MatrixXd> P, Eigen::Map<Eigen::MatrixXd> MMtsqrt,
                                        Eigen::Map<Eigen::MatrixXd> y) {
    Eigen::MatrixXd a(P.rows(), 1);
    check(eg_calculate_reduced_a(varG, P.data(), MMtsqrt.data(), y.data(), P.rows(), a.data()));
    return a;
}

// [[Rcpp::export]]
Eigen::MatrixXd calculate_reduced_vara_gpu(Eigen::Map<Eigen::MatrixXd> X, double varE, double varG,
                                           Eigen::Map<Eigen::MatrixXd> MMtsqrt) {
    Eigen::MatrixXd V(MMtsqrt.rows(), MMtsqrt.rows());
    check(eg_calculate_reduced_vara(X.data(), MMtsqrt.rows(), (int)X.cols(), varE, varG, MMtsqrt.data(), V.data()));
    return V;
}

// The reference's `ngpu` argument made live (R/AM.R:185-196; forced to 0 at R/AM.R:214).  One NEW export: AM() calls it
// where it sets ngpu; every other export above then works on the marker-sharded stores unchanged (include/eagle_gpu.h,
// eg_init_multi).  Returns the number of GPUs in use.
// [[Rcpp::export]]
int eagle_gpu_init(int ngpu) {
    check(ngpu > 1 ? eg_init_multi(ngpu, NULL) : eg_init(0));
    return eg_gpu_count();
}
