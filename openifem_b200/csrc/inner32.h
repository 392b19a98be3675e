// Mixed-precision inner solvers of the block Schur preconditioner
// (BlockSchurPreconditioner::vmult, reference source/mpi_insim.cpp:56-128):
//   * InnerSolver32: the stand-in for the MUMPS factorisation of
//     system_matrix.block(0,0) (mpi_insim.cpp:111-127; Krylov-for-A~ precedent
//     source/mpi_insimex.cpp:114-124) - BiCGStab + node-block Jacobi;
//   * InnerCG32: "CG for Sm", the unpreconditioned CG on B diag(M_u)^-1 B^T
//     (mpi_insim.cpp:88-109).
// FGMRES is flexible, so the preconditioner may be any approximate inverse: both run
// entirely in fp32 on a copy of their matrix laid out for streaming (the operator,
// residuals and Krylov basis of the outer FGMRES stay fp64 on the row-plane BCSR
// matrices, and converged Newton states do not depend on the inner precision).
//
// Layout of the copy ("sliced ELL of bs x bs blocks", SELL-32):
//   * block rows are re-ordered: the domain is cut into columns of T x T nodes along
//     the sweep axis (last coordinate), inside a column rows are sorted by plane and
//     then by row length, and 32 consecutive rows form a slice. Rows of equal length
//     share slices (padding 0.5 % at config 3) and everything a wave of CTAs touches
//     in x is a few MB, so x is read from HBM once;
//   * slice s with L_s block slots stores col[(off_s + j) * 32 + lane] and
//     val[((off_s + j) * bs*bs + k) * 32 + lane]: one lane per row, every load of a
//     warp is one full 128-byte line of a purely sequential stream;
//   * precision 16: values scaled by 1 / max|row| per scalar row and stored as half2 =
//     two consecutive slots of a row (int2 = their column indices): 22 instead of 40
//     bytes per 3x3 block, products and sums in fp32;
//   * the solvers' vectors live in the same permuted ("SELL") numbering, so the
//     product is written coalesced and the x gathers of a slice hit runs of
//     consecutive nodes; gather sources of bs > 1 are padded to float4 per node.
#pragma once
#include "halo.h"
#include "krylov.h"
#include "linalg.h"
#include "peer.h"

namespace ifem
{
  struct Sell32
  {
    int bs = 0;
    int precision = 32;          // storage of the matrix values: 32 = float, 16 = row-scaled half
    int n_rows = 0;              // owned block rows (= rows of a product)
    int n_cols = 0;              // local block columns (owned + ghosts)
    int n_slices = 0, n_pad = 0; // n_pad = 32 * n_slices >= n_rows
    int64_t n_slots = 0;         // sum of slice lengths
    int64_t n_hslots = 0;        // sum of ceil(slice length / 2): double slots of the fp16 storage
    int64_t n_blocks = 0;        // blocks of the owned rows (unpadded)
    DevBuf<int> slice_off;       // [n_slices + 1]
    DevBuf<int> perm_row;        // [n_pad] original block row of a SELL row, -1 = padding row
    DevBuf<int> pos;             // [n_cols] SELL position of an original node (ghosts: n_pad + ghost no.)
    DevBuf<int> col;             // [n_slots * 32] SELL position of the column node
    DevBuf<float> val;           // [n_slots * bs * bs * 32]
    // fp16 storage: half2 = two consecutive slots of a row, int2 = their column indices, one scale per scalar row
    DevBuf<int> hoff;            // [n_slices + 1] double-slot offsets
    DevBuf<int> col2;            // [n_hslots * 32] int2
    DevBuf<unsigned int> valh;   // [n_hslots * bs * bs * 32] half2
    DevBuf<float> row_scale;     // [n_pad * bs]
    std::vector<int> h_pos;
    int variant = 24;            // kernel shape: 10 * slots per step + resident CTAs per SM
    // halo plan in SELL numbering
    const Halo *halo_plan = nullptr;
    DevBuf<int> send_pos;
    DevBuf<float> send_buf;
    // peer-memory halo (peer.h): gather sources live in IPC-shared allocations, the pack kernel writes straight into the
    // neighbours' ghost segments. sources[k] = {own pointer, the same buffer of every rank}
    struct Source
    {
      float *local = nullptr;
      std::vector<void *> peers;
    };
    std::vector<Source> sources;
    DevBuf<float> plain_sources[4];        // gather sources when there is no peer link
    bool peer_halo = false;
    std::vector<int64_t> ghost_off;        // [size][size][kPeerMaxMsgs]: float offsets in rank r's source where rank s's messages land (-1: none)
    std::vector<void *> flag_peers;        // arrival flags of every rank ([kPeerMaxRanks] unsigned per rank)
    DevBuf<unsigned int> halo_state;       // [0] pushes carried out, [1] CTA counter of the push kernel

    bool built() const { return n_slices > 0; }
    int xs() const { return bs == 1 ? 1 : 4; } // floats per node of a gather source
    size_t x_len() const { return ((size_t)n_pad + (size_t)(n_cols - n_rows)) * xs(); }
    // pattern-only work, once per sparsity pattern: row order, slices, column map, halo plan in SELL numbering
    void build(Context &ctx, const Bcsr &A, const NodeTable &nodes, const Halo *halo, int precision);
    // values of A -> the copy
    void refresh(Context &ctx, const Bcsr &A);
    // y (bs floats per SELL row) = A x, x a gather source of x_len() floats with up-to-date ghosts
    void apply(Context &ctx, const float *x, float *y, const int *skip = nullptr) const;
    // refresh the ghost entries of a gather source from their owners; with `skip` (device flag) set the exchange is a no-op
    // on every rank (peer mode only - NCCL exchanges always run)
    void halo(Context &ctx, float *x, const int *skip = nullptr);
    // a zeroed vector of x_len() floats whose ghost segment the neighbours can write (slot: 0..3, reused on rebuild)
    float *gather_source(Context &ctx, int slot);
    void setup_peer_halo(Context &ctx);
    // bytes one product has to move: values + column index + slice offsets, x read once, y written once
    double spmv_bytes() const
    {
      const double per_value = precision == 16 ? 2.0 : 4.0, scales = precision == 16 ? 4.0 * bs * n_rows : 0.0;
      return per_value * n_blocks * bs * bs + 4.0 * n_blocks + 4.0 * (n_slices + 1) + 4.0 * xs() * n_cols + 4.0 * bs * n_rows + scales;
    }
    double padding() const
    {
      if (!n_blocks) return 1.0;
      return (precision == 16 ? 2.0 * double(n_hslots) : double(n_slots)) * 32.0 / double(n_blocks);
    }
  };

  class InnerSolver32
  {
  public:
    ~InnerSolver32();
    // precision: 32 (float values) or 16 (half values scaled per scalar row)
    void setup(Context &ctx, const Bcsr &A, const NodeTable &nodes, const Halo *halo, int precision = 32);
    // values of A and of the inverted diagonal blocks (row-major bs x bs per node) -> fp32, every solve
    void refresh(Context &ctx, const Bcsr &A, const double *binv);
    // dst ~= A^-1 src to |r| <= rel_tol * |src| (recurrence residual), x0 = 0; src_norm = |src| over all ranks
    SolveResult solve(Context &ctx, const double *src, double src_norm, double *dst, double rel_tol, int max_it);
    // y = A32 x with x, y fp64 vectors in the original numbering (tests, kernel timing): load x into the gather
    // buffer (ghosts exchanged), apply the product kernel, store the result
    void probe_load(Context &ctx, const double *x);
    void probe_apply(Context &ctx);
    void probe_store(Context &ctx, double *y);
    Sell32 S;

    int check_every = 6;       // iterations enqueued between two looks at the device state
    long long n_fallbacks = 0; // applications that fell back to one block-Jacobi step (no progress in fp32)

  private:
    template <int BS>
    SolveResult solve_impl(Context &ctx, const double *src, double src_norm, double *dst, double rel_tol, int max_it);
    DevBuf<float> r, r0, p, v, s, t, x; // [n_pad * bs]
    float *ph = nullptr, *sh = nullptr; // [x_len], gather sources (S.gather_source)
    DevBuf<float> binv;                 // [bs * bs][n_pad]
    DevBuf<double> partials, red;       // CTA partial sums; the reduced (all-rank) values of the last reduction
    DevBuf<unsigned int> counter;       // CTA arrival counter of the reducing kernels
    DevBuf<int> state;                  // BicgState (inner32.cu)
    void *h_state = nullptr;            // pinned mirror
    int grid = 0;
  };

  // Unpreconditioned CG in fp32 on the SELL-32 copy of a scalar (1 x 1) matrix, x0 = 0, absolute tolerance.
  // Both solvers are driven from device-resident scalars: every dot product is finished inside the kernel that produced
  // its partial sums (last CTA; summed over the ranks through the peer link, peer.h), the coefficients of the recurrences
  // are computed there and read by the next kernel, so there is no host round trip and no collective launch inside an
  // iteration. The host enqueues `check_every` iterations, then reads the state; iterations past convergence are no-ops.
  class InnerCG32
  {
  public:
    ~InnerCG32();
    void setup(Context &ctx, const Bcsr &A, const NodeTable &nodes, const Halo *halo, int precision = 32);
    void refresh(Context &ctx, const Bcsr &A) { S.refresh(ctx, A); }
    // dst ~= A^-1 src to |r| <= tol_abs; src_norm = |src| over all ranks
    SolveResult solve(Context &ctx, const double *src, double src_norm, double *dst, double tol_abs, int max_it);
    Sell32 S;
    int check_every = 10;
    // single-reduction recurrence (Chronopoulos-Gear): one fused all-reduce of two values and four launches per iteration instead
    // of two all-reduces and six launches (inner32.cu cg_gear_*); IFEM_CG_SM_GEAR=0 selects the classical recurrence
    bool single_reduction = true;
    // Two-level preconditioner for "CG for Sm" (the reference runs it unpreconditioned, mpi_insim.cpp:88-109: 4 200 iterations per
    // time step at config 3): M^-1 = diag(A)^-1 + Z E^+ Z^T with Z = piecewise constants over G^dim boxes of the domain (<= 729
    // aggregates) and E = Z^T A Z formed on the host whenever the matrix changes, inverted there once (rank-one shifted: S_m is
    // singular in closed cavities). Per iteration: a segmented sum per aggregate (fixed order: deterministic), one all-reduce of
    // the coarse vector over the ranks, a dense mat-vec with E^+ and a fused build of u = M^-1 r. Inside the preconditioner of
    // FGMRES only the accuracy of the result matters (tolerance 1e-3 |src| as in the reference). IFEM_CG_SM_COARSE=0 switches it off.
    bool coarse_space = true;
    int n_coarse = 0;
    // global bounding box of the nodes (every rank the same): set before refresh() so that the aggregates agree across ranks
    double box[6] = {0, 1, 0, 1, 0, 1};
    int dim = 3;
    void build_coarse(Context &ctx, const Bcsr &A, const NodeTable &nodes);

  private:
    DevBuf<int> agg;              // [n_pad] aggregate of a SELL row (-1: padding row)
    DevBuf<int> agg_ptr, agg_rows; // rows of every aggregate (segmented restriction in a fixed order)
    DevBuf<double> Einv, cvec, yvec;
    DevBuf<float> dinv;           // [n_pad] 1 / diag(A)
    DevBuf<float> r, ap, x; // [n_pad]
    DevBuf<float> pg, sg;   // [n_pad] search direction and its product (single-reduction variant)
    float *p = nullptr;     // [x_len] gather source (S.gather_source)
    float *rg = nullptr;    // [x_len] gather source holding the residual (single-reduction variant)
    DevBuf<double> partials, red;
    DevBuf<unsigned int> counter;
    DevBuf<int> state;      // CgState (inner32.cu)
    void *h_state = nullptr;
    int grid = 0;
  };
} // namespace ifem
