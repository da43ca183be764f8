// C entry points over the REFERENCE'S OWN functions (compiled from /root/reference against the
// stand-in RcppEigen.h).  Same C signatures as oracle/eagle_oracle.c so that the Python front end
// can load either library.  TEST INFRASTRUCTURE ONLY.
#include <RcppEigen.h>
#include <omp.h>

Eigen::MatrixXd ReadBlock(std::string asciifname, long start_row, long numcols, long numrows_in_block);
Eigen::MatrixXd calculateMMt_rcpp(Rcpp::CharacterVector f_name_ascii, double max_memory_in_Gbytes, int num_cores,
                                  Rcpp::NumericVector selected_loci, std::vector<long> dims, bool quiet,
                                  Rcpp::Function message);
Rcpp::List calculate_a_and_vara_rcpp(Rcpp::CharacterVector f_name_ascii, Rcpp::NumericVector selected_loci,
                                     Eigen::Map<Eigen::MatrixXd> inv_MMt_sqrt,
                                     Eigen::Map<Eigen::MatrixXd> dim_reduced_vara, double max_memory_in_Gbytes,
                                     std::vector<long> dims, Eigen::VectorXd a, bool quiet, Rcpp::Function message);
Eigen::MatrixXd calculate_reduced_a_rcpp(Rcpp::CharacterVector f_name_ascii, double varG,
                                         Eigen::Map<Eigen::MatrixXd> P, Eigen::Map<Eigen::MatrixXd> y,
                                         double max_memory_in_Gbytes, std::vector<long> dims,
                                         Rcpp::NumericVector selected_loci, bool quiet, Rcpp::Function message);
Eigen::VectorXi extract_geno_rcpp(Rcpp::CharacterVector f_name_ascii, double max_memory_in_Gbytes,
                                  long selected_locus, std::vector<long> dims);

bool createM_ASCII_rcpp(Rcpp::CharacterVector f_name, Rcpp::CharacterVector f_name_ascii, Rcpp::CharacterVector type,
                        std::string AA, std::string AB, std::string BB, double max_memory_in_Gbytes, std::vector<long> dims,
                        bool quiet, Rcpp::Function message, std::string missing);
void createMt_ASCII_rcpp(Rcpp::CharacterVector f_name, Rcpp::CharacterVector f_name_ascii, Rcpp::CharacterVector type,
                         double max_memory_in_Gbytes, std::vector<long> dims, bool quiet, Rcpp::Function message);

std::vector<long> ReshapeM_rcpp(Rcpp::CharacterVector fnameM, Rcpp::CharacterVector fnameMt, std::vector<long> indxNA,
                                std::vector<long> dims);
std::vector<long> getRowColumn(std::string fname);

static thread_local std::string g_what;
// messages joined with the ASCII record separator into a caller-provided buffer
static void join_msgs(const std::vector<std::string>& msgs, char* buf, long cap) {
    if (!buf || cap <= 0) return;
    std::string all;
    for (size_t i = 0; i < msgs.size(); i++) {
        if (i) all += '\x1e';
        all += msgs[i];
    }
    if ((long)all.size() >= cap) all.resize(cap - 1);
    std::memcpy(buf, all.c_str(), all.size() + 1);
}
#define GUARD(...)                                                           \
    try { __VA_ARGS__; return 0; }                                           \
    catch (const std::exception& e) { g_what = e.what(); return 1; }         \
    catch (...) { g_what = "unknown"; return 1; }

extern "C" {

const char* ref_last_error() { return g_what.c_str(); }

int eo_ReadBlock(const char* f, long start_row, long numcols, long numrows, double* out) {
    GUARD({
        Eigen::MatrixXd M = ReadBlock(f, start_row, numcols, numrows);
        std::memcpy(out, M.data(), sizeof(double) * (size_t)M.size());
    })
}

int eo_calculateMMt(const char* f, double mem, int cores, const double* sel, long nsel, const long* dims, double* out,
                    int* branch) {
    GUARD({
        std::vector<std::string> msgs;
        Eigen::MatrixXd M = calculateMMt_rcpp(f, mem, cores, Rcpp::NumericVector(sel, nsel),
                                              std::vector<long>{dims[0], dims[1]}, true, Rcpp::Function(&msgs));
        std::memcpy(out, M.data(), sizeof(double) * (size_t)M.size());
        if (branch) {  // the blocked branch announces itself (calculateMMt_rcpp.cpp:107)
            *branch = 0;
            for (auto& m : msgs) if (m.find("number of rows in block") != std::string::npos) *branch = 1;
        }
    })
}

int eo_calculate_a_and_vara(const char* f, const double* sel, long nsel, const double* S, const double* V, double mem,
                            const long* dims, const double* a, double* out_a, double* out_vara, int* branch) {
    GUARD({
        const long L = dims[0], n = dims[1];
        Eigen::VectorXd av(n);
        std::memcpy(av.data(), a, sizeof(double) * n);
        std::vector<std::string> msgs;
        Rcpp::List r = calculate_a_and_vara_rcpp(f, Rcpp::NumericVector(sel, nsel), Eigen::Map<Eigen::MatrixXd>(S, n, n),
                                                 Eigen::Map<Eigen::MatrixXd>(V, n, n), mem, std::vector<long>{L, n}, av,
                                                 true, Rcpp::Function(&msgs));
        if (branch) {
            *branch = 0;
            for (auto& m : msgs) if (m.find("Increasing maxmemGb") != std::string::npos) *branch = 1;
        }
        if (r.items["a"].size() != L) return 4;  // soft failure: List(a=0, vara=0) (:141-142)
        std::memcpy(out_a, r.items["a"].data(), sizeof(double) * L);
        std::memcpy(out_vara, r.items["vara"].data(), sizeof(double) * L);
    })
}

int eo_calculate_reduced_a(const char* f, double varG, const double* P, const double* y, double mem, const long* dims,
                           const double* sel, long nsel, double* out) {
    GUARD({
        const long n = dims[0], L = dims[1];
        Eigen::MatrixXd r = calculate_reduced_a_rcpp(f, varG, Eigen::Map<Eigen::MatrixXd>(P, n, n),
                                                     Eigen::Map<Eigen::MatrixXd>(y, n, 1), mem,
                                                     std::vector<long>{n, L}, Rcpp::NumericVector(sel, nsel), true,
                                                     Rcpp::Function());
        if (r.size() != L) return 4;
        std::memcpy(out, r.data(), sizeof(double) * L);
    })
}

int eo_extract_geno(const char* f, double mem, long locus, const long* dims, int* out, int* branch) {
    GUARD({
        Eigen::VectorXi v = extract_geno_rcpp(f, mem, locus, std::vector<long>{dims[0], dims[1]});
        std::memcpy(out, v.data(), sizeof(int) * (size_t)v.size());
        if (branch) *branch = (mem > ((double)dims[0] * dims[1] * sizeof(double)) / 1000000000.0) ? 0 : 1;
    })
}

int eo_createM_ASCII(const char* f, const char* fascii, const char* type, const char* AA, const char* AB, const char* BB,
                     double mem, const long* dims, int quiet, const char* missing, char* msgbuf, long msgcap, int* ok) {
    GUARD({
        std::vector<std::string> msgs;
        *ok = createM_ASCII_rcpp(f, fascii, type, AA, AB, BB, mem, std::vector<long>{dims[0], dims[1]}, quiet != 0,
                                 Rcpp::Function(&msgs), missing) ? 1 : 0;
        join_msgs(msgs, msgbuf, msgcap);
    })
}

int eo_createMt_ASCII(const char* f, const char* fascii, const char* type, double mem, const long* dims, int quiet,
                      char* msgbuf, long msgcap) {
    GUARD({
        std::vector<std::string> msgs;
        createMt_ASCII_rcpp(f, fascii, type, mem, std::vector<long>{dims[0], dims[1]}, quiet != 0, Rcpp::Function(&msgs));
        join_msgs(msgs, msgbuf, msgcap);
    })
}

int eo_ReshapeM(const char* fM, const char* fMt, const long* indxNA, long n_indx, const long* dims, long* newdims) {
    GUARD({
        std::vector<long> r = ReshapeM_rcpp(fM, fMt, std::vector<long>(indxNA, indxNA + n_indx), std::vector<long>{dims[0], dims[1]});
        newdims[0] = r[0];
        newdims[1] = r[1];
    })
}

int eo_getRowColumn(const char* f, long* dimen) {
    GUARD({
        std::vector<long> r = getRowColumn(f);
        dimen[0] = r[0];
        dimen[1] = r[1];
    })
}

int eo_num_threads(void) { return omp_get_max_threads(); }
}
