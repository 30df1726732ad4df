// block_matrix: a symmetry-blocked matrix = sorted DualIndex + one dense column-major block per sector.
// Mirrors the reference container (dmrg/block_matrix/block_matrix.h:34-178, block_matrix.hpp) and its
// dense element type alps::numeric::matrix<double> (column-major, alps/numeric/matrix/matrix.hpp:66).
// Transposed / adjoint "views" (block_matrix_algorithms.h:495-517) are re-sorted bases that point at the
// original blocks with a transpose flag, so no data is copied.
#pragma once
#include "index.hpp"
#include <cmath>
#include <cstring>
#include <functional>

extern "C" void scipy_dgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k,
                             const double* alpha, const double* a, const int* lda, const double* b, const int* ldb,
                             const double* beta, double* c, const int* ldc);

namespace qcm {

struct Matrix
{
    size_t rows = 0, cols = 0;
    std::vector<double> v;
    Matrix() {}
    Matrix(size_t r, size_t c, double init = 0.) : rows(r), cols(c), v(r * c, init) {}
    // shape without storage (device-resident block whose values have not been fetched)
    static Matrix shell(size_t r, size_t c) { Matrix m; m.rows = r; m.cols = c; return m; }
    double& operator()(size_t i, size_t j) { return v[i + j * rows]; }
    double const& operator()(size_t i, size_t j) const { return v[i + j * rows]; }
    double* data() { return v.data(); }
    double const* data() const { return v.data(); }
    // alps::numeric::resize semantics: keep the overlapping top-left part, zero fill the rest
    void resize(size_t r, size_t c)
    {
        if (r == rows && c == cols) return;
        Matrix n(r, c, 0.);
        for (size_t j = 0; j < std::min(c, cols); ++j)
            for (size_t i = 0; i < std::min(r, rows); ++i) n(i, j) = (*this)(i, j);
        *this = std::move(n);
    }
    Matrix& operator+=(Matrix const& o) { for (size_t i = 0; i < v.size(); ++i) v[i] += o.v[i]; return *this; }
    Matrix& operator*=(double a) { for (auto& x : v) x *= a; return *this; }
    bool operator==(Matrix const& o) const { return rows == o.rows && cols == o.cols && v == o.v; }
};

// a (possibly transposed) reference to a dense block
struct MatRef
{
    double const* p; size_t rows, cols, ld; bool trans;   // rows/cols are the LOGICAL shape (after transposition)
};

// C(m x n) = alpha * op(A) * op(B) + beta * C ; C column-major with leading dimension ldc
inline void dgemm(MatRef const& A, MatRef const& B, double alpha, double beta, double* C, size_t ldc)
{
    int m = (int)A.rows, n = (int)B.cols, k = (int)A.cols, lda = (int)A.ld, ldb = (int)B.ld, ldc_ = (int)ldc;
    if (m == 0 || n == 0) return;
    char ta = A.trans ? 'T' : 'N', tb = B.trans ? 'T' : 'N';
    if (lda < 1) lda = 1; if (ldb < 1) ldb = 1;
    scipy_dgemm_(&ta, &tb, &m, &n, &k, &alpha, A.p, &lda, B.p, &ldb, &beta, C, &ldc_);
}

class block_matrix
{
public:
    block_matrix() {}
    // block_matrix(Index rows, Index cols): diagonal pairing of two equally long indices (block_matrix.hpp:26-35)
    block_matrix(Index const& rows, Index const& cols)
    {
        assert(rows.size() == cols.size());
        for (size_t k = 0; k < rows.size(); ++k) {
            basis_.push_back_unsorted(QnBlock(rows[k].first, cols[k].first, rows[k].second, cols[k].second));
            data_.emplace_back(rows[k].second, cols[k].second);
        }
    }
    DualIndex const& basis() const { return basis_; }
    size_t n_blocks() const { return data_.size(); }
    Matrix& operator[](size_t k) { return data_[k]; }
    Matrix const& operator[](size_t k) const { return data_[k]; }
    MatRef block(size_t k) const { return MatRef{data_[k].data(), data_[k].rows, data_[k].cols, data_[k].rows, false}; }
    size_t find_block(Charge const& r, Charge const& c) const { return basis_.position(r, c); }
    bool has_block(Charge const& r, Charge const& c) const { return basis_.has(r, c); }
    Matrix& operator()(Charge const& r, Charge const& c) { return data_[basis_.position(r, c)]; }
    Matrix const& operator()(Charge const& r, Charge const& c) const { return data_[basis_.position(r, c)]; }
    Index left_basis() const { return basis_.left_basis(); }
    Index right_basis() const { return basis_.right_basis(); }

    size_t insert_block(Matrix mtx, Charge const& c1, Charge const& c2)
    {
        size_t i1 = basis_.insert(QnBlock(c1, c2, mtx.rows, mtx.cols));
        data_.insert(data_.begin() + i1, std::move(mtx));
        return i1;
    }
    void remove_block(size_t which)
    {
        basis_.erase(which);
        data_.erase(data_.begin() + which);
    }
    void resize_block(size_t pos, size_t r, size_t c)
    {
        data_[pos].resize(r, c);
        basis_[pos].ls = r; basis_[pos].rs = c;
    }
    void clear() { data_.clear(); basis_.clear(); }

    // block_matrix.hpp:400-429 -- may GROW an existing block to the element-wise max shape
    void match_and_add_block(Matrix const& mtx, Charge const& c1, Charge const& c2)
    {
        size_t match = find_block(c1, c2);
        if (match < n_blocks()) {
            Matrix& t = data_[match];
            if (mtx.rows == t.rows && mtx.cols == t.cols) t += mtx;
            else {
                size_t mr = std::max(mtx.rows, t.rows), mc = std::max(mtx.cols, t.cols);
                Matrix cpy(mtx);
                resize_block(match, mr, mc);
                cpy.resize(mr, mc);
                data_[match] += cpy;
            }
        } else
            insert_block(mtx, c1, c2);
    }
    block_matrix& operator+=(block_matrix const& rhs)
    {
        for (size_t k = 0; k < rhs.n_blocks(); ++k) {
            Charge r = rhs.basis_[k].lc, c = rhs.basis_[k].rc;
            if (has_block(r, c)) (*this)(r, c) += rhs.data_[k];
            else insert_block(rhs.data_[k], r, c);
        }
        return *this;
    }
    // *this += a * rhs without a temporary (blocks matched by charge, missing ones inserted)
    block_matrix& axpy(double a, block_matrix const& rhs)
    {
        for (size_t k = 0; k < rhs.n_blocks(); ++k) {
            Charge r = rhs.basis_[k].lc, c = rhs.basis_[k].rc;
            size_t j = (k < n_blocks() && basis_[k].lc == r && basis_[k].rc == c) ? k : find_block(r, c);
            if (j < n_blocks() && data_[j].rows == rhs.data_[k].rows && data_[j].cols == rhs.data_[k].cols) {
                double* y = data_[j].v.data(); const double* x = rhs.data_[k].v.data();
                const size_t n = data_[j].v.size();
                for (size_t i = 0; i < n; ++i) y[i] += a * x[i];
            } else {
                Matrix t = rhs.data_[k]; t *= a;
                if (j < n_blocks()) match_and_add_block(t, r, c); else insert_block(t, r, c);
            }
        }
        return *this;
    }
    block_matrix& operator*=(double a) { for (auto& m : data_) m *= a; return *this; }
    void generate(std::function<double()> g) { for (auto& m : data_) for (auto& x : m.v) x = g(); }
    double norm_square() const { double r = 0; for (auto const& m : data_) for (double x : m.v) r += x * x; return r; }
    double trace() const
    {
        double r = 0;
        for (auto const& m : data_) for (size_t i = 0; i < std::min(m.rows, m.cols); ++i) r += m(i, i);
        return r;
    }
    // <this, rhs> over matching sectors (mpstensor.hpp:363-395 matches blocks by charge)
    double scalar_overlap(block_matrix const& rhs) const
    {
        double r = 0;
        for (size_t k = 0; k < n_blocks(); ++k) {
            size_t j = rhs.find_block(basis_[k].lc, basis_[k].rc);
            if (j == rhs.n_blocks()) continue;
            Matrix const& a = data_[k]; Matrix const& b = rhs.data_[j];
            if (a.rows != b.rows || a.cols != b.cols) throw std::runtime_error("scalar_overlap: block shape mismatch");
            for (size_t i = 0; i < a.v.size(); ++i) r += a.v[i] * b.v[i];
        }
        return r;
    }
    size_t num_elements() const { size_t r = 0; for (auto const& m : data_) r += m.v.size(); return r; }

private:
    DualIndex basis_;
    std::vector<Matrix> data_;
};

// transpose / adjoint view (real arithmetic: conjugate is the identity)
class block_view
{
public:
    block_view(block_matrix const& m, bool transposed)
    {
        if (!transposed) {
            basis_ = m.basis();
            for (size_t k = 0; k < m.n_blocks(); ++k) refs_.push_back(m.block(k));
        } else {
            for (size_t k = 0; k < m.n_blocks(); ++k) {
                QnBlock const& b = m.basis()[k];
                size_t i = basis_.insert(QnBlock(b.rc, b.lc, b.rs, b.ls));
                MatRef r = m.block(k);
                refs_.insert(refs_.begin() + i, MatRef{r.p, r.cols, r.rows, r.ld, true});
            }
        }
    }
    DualIndex const& basis() const { return basis_; }
    size_t n_blocks() const { return refs_.size(); }
    MatRef block(size_t k) const { return refs_[k]; }
    Index left_basis() const { return basis_.left_basis(); }
    Index right_basis() const { return basis_.right_basis(); }

private:
    DualIndex basis_;
    std::vector<MatRef> refs_;
};
inline block_view transpose(block_matrix const& m) { return block_view(m, true); }
inline block_view adjoint(block_matrix const& m) { return block_view(m, true); }
inline block_view conjugate(block_matrix const& m) { return block_view(m, false); }
inline block_view plain(block_matrix const& m) { return block_view(m, false); }
// a transposed COPY (block (lc, rc) of size m x n becomes block (rc, lc) of size n x m)
inline block_matrix transposed(block_matrix const& A)
{
    block_matrix r;
    for (size_t k = 0; k < A.n_blocks(); ++k) {
        Matrix const& m = A[k]; Matrix t(m.cols, m.rows);
        for (size_t j = 0; j < m.cols; ++j) for (size_t i = 0; i < m.rows; ++i) t(j, i) = m(i, j);
        r.insert_block(t, A.basis()[k].rc, A.basis()[k].lc);
    }
    return r;
}

} // namespace qcm
