// Stand-in for <RcppEigen.h> -- TEST INFRASTRUCTURE ONLY (oracle/refshim).
//
// Lets the REFERENCE'S OWN hot-path sources (ReadBlock.cpp, calculateMMt_rcpp.cpp,
// calculate_a_and_vara_rcpp.cpp, calculate_reduced_a_rcpp.cpp, extract_geno_rcpp.cpp, and the ingest
// routines createM_ASCII_rcpp.cpp, CreateASCIInospace.cpp, CreateASCIInospace_PLINK.cpp,
// createMt_ASCII_rcpp.cpp under /root/reference/MyPackage/Eagle/src) compile unmodified, from where they lie, without R, Rcpp or
// Eigen (none of which exist in this image).  Only the small part of the two libraries' API that
// those five files use is provided, with eager evaluation and straightforward loops.  What runs is
// therefore the reference's own control flow -- file parsing, memory tests, row blocking, the
// zeroing of selected loci, the order of the products -- on top of THIS file's dense arithmetic
// (not Eigen's blocked kernels).  It pins oracle/eagle_oracle.c against the reference's code; it
// is not a performance baseline.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace Eigen {

inline void initParallel() {}
inline void setNbThreads(int) {}

class MatrixXd;

class VectorXi {
  public:
    VectorXi() {}
    explicit VectorXi(long n) : v_(n, 0) {}
    int& operator()(long i) { return v_[i]; }
    int operator()(long i) const { return v_[i]; }
    long size() const { return (long)v_.size(); }
    int* data() { return v_.data(); }
    std::vector<int> v_;
};

// integer matrix of createMt_ASCII_rcpp.cpp:79-99 (element access and transpose only)
class MatrixXi {
  public:
    MatrixXi() : r_(0), c_(0) {}
    MatrixXi(long r, long c) : r_(r), c_(c), d_((size_t)r * (size_t)c, 0) {}
    long rows() const { return r_; }
    long cols() const { return c_; }
    int& operator()(long r, long c) { return d_[(size_t)r + (size_t)c * r_]; }
    int operator()(long r, long c) const { return d_[(size_t)r + (size_t)c * r_]; }
    MatrixXi transpose() const {
        MatrixXi t(c_, r_);
        for (long j = 0; j < c_; j++)
            for (long i = 0; i < r_; i++) t(j, i) = (*this)(i, j);
        return t;
    }
  private:
    long r_, c_;
    std::vector<int> d_;
};

class MatrixXd {
  public:
    MatrixXd() : r_(0), c_(0) {}
    MatrixXd(long r, long c) : r_(r), c_(c), d_((size_t)r * (size_t)c) {}
    static MatrixXd Zero(long r, long c) { MatrixXd m(r, c); m.setZero(); return m; }
    long rows() const { return r_; }
    long cols() const { return c_; }
    long size() const { return r_ * c_; }
    double* data() { return d_.data(); }
    const double* data() const { return d_.data(); }
    double& operator()(long r, long c) { return d_[(size_t)r + (size_t)c * r_]; }
    double operator()(long r, long c) const { return d_[(size_t)r + (size_t)c * r_]; }
    double& operator()(long i) { return d_[i]; }
    double operator()(long i) const { return d_[i]; }
    MatrixXd& setZero() { std::fill(d_.begin(), d_.end(), 0.0); return *this; }
    MatrixXd& noalias() { return *this; }
    void resize(long r, long c) { r_ = r; c_ = c; d_.assign((size_t)r * (size_t)c, 0.0); d_.shrink_to_fit(); }
    MatrixXd transpose() const {
        MatrixXd t(c_, r_);
        for (long j = 0; j < c_; j++)
            for (long i = 0; i < r_; i++) t(j, i) = (*this)(i, j);
        return t;
    }

    struct ColRef {
        MatrixXd* m; long j;
        void setZero() { for (long i = 0; i < m->r_; i++) (*m)(i, j) = 0.0; }
        double operator()(long i) const { return (*m)(i, j); }
        template <class T> VectorXi cast() const {
            VectorXi v(m->r_);
            for (long i = 0; i < m->r_; i++) v(i) = (T)(*m)(i, j);
            return v;
        }
    };
    struct RowT { const MatrixXd* m; long i; };  // a transposed row (a column vector)
    struct RowRef {
        MatrixXd* m; long i;
        void setZero() { for (long j = 0; j < m->c_; j++) (*m)(i, j) = 0.0; }
        RowT transpose() const { return RowT{m, i}; }
    };
    struct BlockRef {
        MatrixXd* m; long i0, j0, nr, nc;
        BlockRef& operator=(const MatrixXd& s) {
            for (long j = 0; j < nc; j++)
                for (long i = 0; i < nr; i++) (*m)(i0 + i, j0 + j) = s(i, j);
            return *this;
        }
    };
    ColRef col(long j) { return ColRef{this, j}; }
    RowRef row(long i) { return RowRef{this, i}; }
    BlockRef block(long i, long j, long nr, long nc) { return BlockRef{this, i, j, nr, nc}; }

  private:
    long r_, c_;
    std::vector<double> d_;
};

// row(i) * row(k).transpose()  ->  dot product
inline double operator*(const MatrixXd::RowRef& a, const MatrixXd::RowT& b) {
    double s = 0.0;
    for (long k = 0; k < a.m->cols(); k++) s += (*a.m)(a.i, k) * (*b.m)(b.i, k);
    return s;
}

class VectorXd : public MatrixXd {
  public:
    VectorXd() {}
    explicit VectorXd(long n) : MatrixXd(n, 1) {}
};

template <class M> class Map;
template <> class Map<MatrixXd> {
  public:
    Map(const double* p, long r, long c) : p_(p), r_(r), c_(c) {}
    operator MatrixXd() const {  // eager copy: every use in the reference is as a product operand
        MatrixXd m(r_, c_);
        std::memcpy(m.data(), p_, sizeof(double) * (size_t)r_ * (size_t)c_);
        return m;
    }
    long rows() const { return r_; }
    long cols() const { return c_; }
  private:
    const double* p_; long r_, c_;
};

// C = A * B, column-major, plain i-k-j ordering with the k loop innermost-sequential per entry
inline MatrixXd operator*(const MatrixXd& A, const MatrixXd& B) {
    if (A.cols() != B.rows()) throw std::runtime_error("refshim: product dimension mismatch");
    const long m = A.rows(), k = A.cols(), n = B.cols();
    MatrixXd C(m, n);
    C.setZero();
#pragma omp parallel for schedule(static)
    for (long j = 0; j < n; j++) {
        double* c = C.data() + (size_t)j * m;
        for (long p = 0; p < k; p++) {
            const double b = B(p, j);
            const double* a = A.data() + (size_t)p * m;
            for (long i = 0; i < m; i++) c[i] += a[i] * b;
        }
    }
    return C;
}
inline MatrixXd operator*(const Map<MatrixXd>& A, const MatrixXd& B) { return MatrixXd(A) * B; }
inline MatrixXd operator*(const MatrixXd& A, const Map<MatrixXd>& B) { return A * MatrixXd(B); }
inline MatrixXd operator*(const Map<MatrixXd>& A, const Map<MatrixXd>& B) { return MatrixXd(A) * MatrixXd(B); }
inline MatrixXd operator*(double s, const MatrixXd& A) {
    MatrixXd C(A.rows(), A.cols());
    for (long i = 0; i < A.size(); i++) C(i) = s * A(i);
    return C;
}

}  // namespace Eigen

namespace Rcpp {

class CharacterVector {
  public:
    CharacterVector(const char* s) : s_(s) {}
    CharacterVector(const std::string& s) : s_(s) {}
    std::string s_;
};
template <class T> T as(const CharacterVector& c) { return T(c.s_); }
inline std::ostream& operator<<(std::ostream& os, const CharacterVector& c) { return os << c.s_; }
static std::ostream Rcout(nullptr);  // progress output of the ingest routines: discarded

class NumericVector {
  public:
    NumericVector(const double* p, long n) : v_(p, p + n) {}
    double operator()(long i) const { return v_[i]; }
    long size() const { return (long)v_.size(); }
    std::vector<double> v_;
};

// R closure `message`: collect the text so that tests can look at it
class Function {
  public:
    explicit Function(std::vector<std::string>* sink = nullptr) : sink_(sink) {}
    template <class... A> void operator()(const A&... a) const {
        std::ostringstream os;
        (void)std::initializer_list<int>{((os << a), 0)...};
        if (sink_) sink_->push_back(os.str());
    }
  private:
    std::vector<std::string>* sink_;
};

inline void stop(const std::string& msg) { throw std::runtime_error(msg); }

struct NamedValue { std::string name; Eigen::MatrixXd value; };
struct Named {
    explicit Named(const char* n) : name(n) {}
    NamedValue operator=(const Eigen::MatrixXd& m) const { return NamedValue{name, m}; }
    NamedValue operator=(int v) const { Eigen::MatrixXd m(1, 1); m(0, 0) = v; return NamedValue{name, m}; }
    std::string name;
};
class List {
  public:
    static List create(const NamedValue& a, const NamedValue& b) {
        List l;
        l.items[a.name] = a.value;
        l.items[b.name] = b.value;
        return l;
    }
    std::map<std::string, Eigen::MatrixXd> items;
};

}  // namespace Rcpp

// R's NA_real_: a NaN whose low word is 1954 (arithmetic.c, R_IsNA)
inline bool R_IsNA(double x) {
    if (!std::isnan(x)) return false;
    uint64_t b;
    std::memcpy(&b, &x, 8);
    return (uint32_t)(b & 0xFFFFFFFFu) == 1954u;
}
