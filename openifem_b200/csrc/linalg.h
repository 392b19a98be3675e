// Device sparse matrices and vectors for the Krylov loop.
//
// Matrix layout ("row-plane BCSR"): block rows of R x C blocks with a node-level
// CSR pattern (rowptr, col). Inside a block row with nb blocks the R*C entries of
// the blocks are stored as R*C contiguous planes of nb doubles:
//     val[rowptr[i]*R*C + (r*C + c)*nb + j]   = block j of block row i, entry (r,c)
// so a warp that walks the blocks of a row reads every plane fully coalesced, and
// the column index is stored once per block instead of once per scalar entry
// (uu block of the 3-D Q2/Q1 system: 8.44 B/nnz instead of CSR's 12 B/nnz).
// R = C = 1 is plain CSR.
//
// Replaces PETScWrappers::MPI::SparseMatrix / BlockSparseMatrix::vmult
// (PETSc MatMult, reference call sites mpi_insim.cpp:117,388,
// mpi_hyper_elasticity.cpp:144) and PETSc VecDot/VecNorm/VecAXPY.
#pragma once
#include "device.cuh"
#include "mesh.h"

namespace ifem
{
  struct Bcsr
  {
    int R = 1, C = 1;
    int n_brows = 0, n_bcols = 0;
    int n_brows_spmv = -1; // rows a mat-vec covers (owned rows); -1: all stored rows
    int64_t n_blocks = 0;
    DevBuf<int64_t> rowptr;
    DevBuf<int> col;
    DevBuf<double> val;
    // optional fp32 copy for inexact inner solves only: same row-plane layout but every block row padded to a
    // multiple of 4 blocks (rowptr32 / col32), so that a lane streams 4 consecutive blocks of a plane with one
    // 128-bit load (padding: value 0, column 0)
    DevBuf<float> val32;
    DevBuf<int64_t> rowptr32;
    DevBuf<int> col32;
    int tpr = 32; // threads per block row chosen from the average row length

    void init(const Pattern &P, int R_, int C_, cudaStream_t s);
    void zero(cudaStream_t s) { val.zero(s); }
    int64_t n_rows() const { return (int64_t)n_brows * R; }
    int64_t n_cols() const { return (int64_t)n_bcols * C; }
    int64_t nnz() const { return n_blocks * R * C; }
    // bytes one SpMV with this matrix has to move (values + block column index +
    // row pointer, x read once, y written once) - the roofline numerator
    double spmv_bytes() const { return 8.0 * nnz() + 4.0 * n_blocks + 8.0 * (n_brows + 1) + 8.0 * n_cols() + 8.0 * n_rows(); }
    // scalar CSR copy on the host (parity tests)
    void to_host_csr(cudaStream_t s, std::vector<int64_t> &rp, std::vector<int> &ci, std::vector<double> &v) const;
  };

  // y = A x  (accumulate = false)   or   y += A x  (accumulate = true)
  void spmv(Context &ctx, const Bcsr &A, const double *x, double *y, bool accumulate = false);
  // refresh A.val32 from A.val (allocates on first use)
  void make_fp32_copy(Context &ctx, Bcsr &A);
  // y = A32 x: matrix entries read as fp32 (half the HBM traffic), x / y / accumulation in fp64.
  // Only legal inside a flexible preconditioner (the A~^-1 stand-in), never for the operator itself.
  void spmv_fp32(Context &ctx, const Bcsr &A, const double *x, double *y);

  // Index space of a distributed vector on this rank: the entries a Krylov method owns are up to two
  // contiguous segments [0, len0) and [off1, off1 + len1) of an allocation of n_alloc doubles (the rest
  // are ghost copies refreshed by halo exchange). A plain length converts to a single segment.
  struct VecSpace
  {
    int64_t len0 = 0, off1 = 0, len1 = 0, n_alloc = 0;
    VecSpace() = default;
    VecSpace(int64_t n) : len0(n), off1(0), len1(0), n_alloc(n) {}
    VecSpace(int64_t l0, int64_t o1, int64_t l1, int64_t alloc) : len0(l0), off1(o1), len1(l1), n_alloc(alloc) {}
    int64_t n_owned() const { return len0 + len1; }
  };

  // ---- BLAS-1 on device vectors (deterministic two-stage reductions; dot products are summed over
  //      the ranks of ctx.comm) -----------
  double dot(Context &ctx, const VecSpace &n, const double *x, const double *y);
  double nrm2(Context &ctx, const VecSpace &n, const double *x);
  void axpy(Context &ctx, const VecSpace &n, double a, const double *x, double *y);             // y += a x
  void axpby(Context &ctx, const VecSpace &n, double a, const double *x, double b, double *y);  // y = a x + b y
  void scale(Context &ctx, const VecSpace &n, double a, double *x);                              // x *= a
  void equ(Context &ctx, const VecSpace &n, double a, const double *x, double *y);               // y = a x
  void copy(Context &ctx, const VecSpace &n, const double *x, double *y);
  void fill(Context &ctx, const VecSpace &n, double v, double *x);
  // aux += a V ; return aux . W   (deal.II Vector::add_and_dot, used by FGMRES' MGS)
  double add_and_dot(Context &ctx, const VecSpace &n, double *aux, double a, const double *V, const double *W);
  // Arnoldi orthogonalisation by classical Gram-Schmidt with one re-orthogonalisation pass, fused: aux is made orthogonal to the
  // k basis vectors (k <= 64), h[t] receives the coefficient of basis[t], the return value is |aux| afterwards. Three reductions
  // and one host synchronisation per call whatever k is (the add_and_dot loop of the modified Gram-Schmidt needs k + 1 of each).
  double orthogonalise_cgs2(Context &ctx, const VecSpace &n, int k, const double *const *basis, double *aux, double *h);
  // z = x + a y + b w
  void lin3(Context &ctx, const VecSpace &n, double *z, const double *x, double a, const double *y, double b, const double *w);
  // x[idx[k]] = vals ? vals[k] : 0   (AffineConstraints::distribute for Dirichlet lines)
  void set_indexed(Context &ctx, int n_idx, const int *idx, const double *vals, double *x);
  // x[g] = vals ? vals[g] : 0 on every dof with flag[g] != 0 (AffineConstraints::distribute for Dirichlet lines, flags and values
  // resident on the device: no compacted index list to rebuild when a coupling step changes the lines)
  void set_flagged(Context &ctx, int64_t n, const unsigned char *flag, const double *vals, double *x);
  // y[i] = d[i] * x[i]
  void hadamard(Context &ctx, const VecSpace &n, const double *d, const double *x, double *y);
  // y[i] /= d[i]
  void divide(Context &ctx, const VecSpace &n, const double *d, double *y);
  // y[i] = 1 / x[i]
  void reciprocal(Context &ctx, const VecSpace &n, const double *x, double *y);
  // block-diagonal apply: y_node = Binv_node * x_node (dim x dim blocks, row-major)
  void block_diag_apply(Context &ctx, int n_nodes, int bs, const double *binv, const double *x, double *y);
  // extract and invert the bs x bs diagonal blocks of a square Bcsr with R = C = bs
  void block_diag_inverse(Context &ctx, const Bcsr &A, double *binv);
} // namespace ifem
