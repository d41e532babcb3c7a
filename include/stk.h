/* stk.h -- C ABI of libstk.so, the sm_100a device side of the space-time
 * Kronecker solve path (B200-native replacement for the NumPy/SciPy/PETSc/MPI
 * arithmetic behind /root/reference/source/{mpi_vector,mpi_kron,wavelets,
 * multigrid,linalg}.py).
 *
 * The reference has no FFI of its own (it is pure Python), so
 * every entry point below names the reference Python call site whose
 * arithmetic it replaces.  INTEGRATION.md shows the ctypes stubs a reference
 * maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; `stream` is a cudaStream_t passed as void*
 *    (NULL = legacy default stream).  All work is enqueued on that stream;
 *    nothing synchronises except the *_host entry points and stk_sync.
 *  - every function returns 0 on success or a non-zero cudaError_t / negative
 *    argument error; stk_last_error() returns the message of the last failure
 *    in this thread.  No exception crosses the ABI.
 *  - the library owns no device memory except inside stk_mg handles' small
 *    host-side descriptors; the caller allocates and frees every buffer.
 *
 * Data layout in HBM ("slab block"): the local part of a space-time vector,
 * n_t local time slices x M space dofs, is stored TIME-FASTEST:
 *      x[i * ld + t],  i in [0,M), t in [0,n_t),  ld = n_t rounded up to 4,
 * with the pad entries t in [n_t, ld) kept equal to zero by every kernel.
 * A CSR nonzero a_ij of a space operator therefore multiplies a contiguous run
 * of time values (coalesced, matrix read once per row for all slices), and a
 * time operator is a unit-stride stencil; no transpose is ever needed.
 * The reference's X_loc (n_t, M) row-major array is the transpose of this
 * block (stk_block_from_rowmajor / stk_block_to_rowmajor convert).
 */
#ifndef STK_H
#define STK_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define STK_VERSION 100

int stk_version(void);
const char *stk_last_error(void);
/* cudaStreamSynchronize(stream). */
int stk_sync(void *stream);
/* Number of kernels this library has launched in this process (bench.py's
 * gpu_launches). */
int64_t stk_launch_count(void);

/* ---- layout conversion / host boundary ---------------------------------
 * mpi_vector.py:62-71 (X_loc storage), :124-138 (scatter/gather). */
/* dst block (M, ld) <- src (n_t, M) row-major, both on the device. */
int stk_block_from_rowmajor(const double *src, int n_t, int M, double *dst,
                            int ld, void *stream);
/* dst (n_t, M) row-major <- src block (M, ld), both on the device. */
int stk_block_to_rowmajor(const double *src, int ld, int n_t, int M,
                          double *dst, void *stream);
/* Host entry points (HOST pointers in the reference's X_loc layout; the copy
 * and the transpose happen inside, `staging` is a device scratch buffer of
 * n_t*M doubles).  They synchronise the stream before returning. */
int stk_block_upload_host(const double *host_rowmajor, int n_t, int M,
                          double *dst, int ld, double *staging, void *stream);
int stk_block_download_host(const double *src, int ld, int n_t, int M,
                            double *host_rowmajor, double *staging,
                            void *stream);

/* ---- BLAS-1 on blocks (n = M*ld doubles, pads are zero) -----------------
 * mpi_vector.py:84-122 (in-place ops), linalg.py:29-30,39-40. */
int stk_axpy(double a, const double *x, double *y, int64_t n, void *stream);
int stk_scale(double a, double *x, int64_t n, void *stream);
/* y = x + a*y   (linalg.py:39-40: p *= beta; p += z). */
int stk_xpay(const double *x, double a, double *y, int64_t n, void *stream);
/* w += a*p ; r -= a*t  in one pass (linalg.py:29-30). */
int stk_pcg_update(double a, const double *p, const double *t, double *w,
                   double *r, int64_t n, void *stream);
/* The same two updates with the scalar a = num_dev[0] / den_dev[0] read on the
 * device: alpha = r.z / p.t and beta = r.z / (r.z)_old of linalg.py:27,38 never
 * visit the host, so a PCG iteration has one read-back (the stop test). */
int stk_xpay_dev(const double *x, const double *num_dev, const double *den_dev,
                 double *y, int64_t n, void *stream);
int stk_pcg_update_dev(const double *num_dev, const double *den_dev,
                       const double *p, const double *t, double *w, double *r,
                       int64_t n, void *stream);
/* out_dev[0] = sum_k x[k]*y[k], single pass, deterministic: per-CTA
 * warp-shuffle partials, the last CTA adds them in index order
 * (mpi_vector.py:205-210 local np.dot; the allreduce is the caller's).
 * `ws` is a device workspace of STK_DOT_WS doubles, zero-initialised once. */
#define STK_DOT_WS 2048
int stk_dot(const double *x, const double *y, int64_t n, double *ws,
            double *out_dev, void *stream);

/* ---- space operator: batched CSR SpMM over all time slices --------------
 * mpi_kron.py:143-150 (IdentityKronMatMPI), linop.py:75-79, and the CSR
 * products inside multigrid.py:174,180.
 *   y[i,t] = alpha * sum_k coef_k[t] * sum_p vals_k[p] * x[indices[p], t]
 *            + beta * z[i,t]
 * K = 1: coef ignored (may be NULL), plain CSR.  K = 2: the matrix of slice t
 * is coef0[t]*vals0 + coef1[t]*vals1 on a shared sparsity pattern (e.g.
 * 2^j M_x + alpha A_x, heateq_mpi.py:97-98); coef arrays have ld entries.
 * z may be NULL when beta == 0 and may alias y; x must not alias y. */
int stk_space_spmm(int nrows, const int *indptr, const int *indices, int K,
                   const double *vals0, const double *vals1,
                   const double *coef0, const double *coef1, const double *x,
                   double alpha, double beta, const double *z, double *y,
                   int ld, void *stream);

/* Row schedule of a CSR structure for all stk_space_spmm* entry points (and the
 * SpMMs inside stk_mg_apply): `order` (device, nrows entries, a permutation of
 * the rows) is the order in which the kernels walk the rows of the matrix
 * whose row-pointer array is `indptr`; NULL removes it.  Results do not depend
 * on it -- rows are independent -- only DRAM traffic does: with a hierarchical
 * FE numbering the index order sweeps the mesh once per vertex class and
 * re-fetches every x row each time; a locality order (reverse Cuthill-McKee of
 * the pattern, computed at setup) walks the mesh once.  The reference has no
 * counterpart (scipy's csr_matvecs walks rows in index order,
 * mpi_kron.py:149); remove the order before freeing `indptr`. */
int stk_csr_set_row_order(const int *indptr, int nrows, const int *order);

/* ---- time operator: sparse matrix along the time axis -------------------
 * mpi_kron.py:186-201 (TridiagKronIdentityMPI), :285-317
 * (SparseKronIdentityMPI), :240-256 (MatKronIdentityMPI after permute).
 *   y[i,t] = alpha * sum_p vals[p] * X(i, indices[p]) + beta * y[i,t],
 *   t in [0, nrows_t); pads t in [nrows_t, ldy) are written as zero when
 *   beta == 0.
 * X(i,c) = x[i*ldx + c] for c < ncols_local, else the halo slice
 * xh[(c-ncols_local)*M + i] (slice-major, n_halo slices as received from
 * neighbour ranks, mpi_vector.py:140-203).  nnz = indptr[nrows_t].  Matrices
 * with many nonzeros per row (the multi-level wavelet transform) are staged in
 * shared memory together with a panel of time columns.  y must not alias x. */
int stk_time_apply(int M, int nrows_t, int nnz, const int *indptr,
                   const int *indices, const double *vals, const double *x,
                   int ldx, int ncols_local, const double *xh, int n_halo,
                   double alpha, double beta, double *y, int ldy,
                   void *stream);
/* Two-input form, y = alpha * [Ta Tb] [x0; x1] + beta * y: column c of the
 * stacked CSR addresses x0 (c < ncols_local), x1 (c < 2 ncols_local), the
 * n_halo0 halo slices of x0, then the halo slices of x1.  One pass computes
 * (A_t (x) I) Mx + (L_t (x) I) Ax of heateq_mpi.py:166-178.  y has pitch ldy
 * and wy <= ldy columns are written (two results side by side in one block
 * of pitch 2*ld go through the multigrid solve together). */
int stk_time_apply2(int M, int nrows_t, const int *indptr, const int *indices,
                    const double *vals, const double *x0, const double *x1,
                    int ldx, int ncols_local, const double *xh0, int n_halo0,
                    const double *xh1, double alpha, double beta, double *y,
                    int ldy, int wy, void *stream);
/* Both brackets of the regrouped Schur operator for TRIDIAGONAL time matrices
 * in one pass (the four TridiagKronIdentityMPI applies of heateq_mpi.py:166-178,
 * mpi_kron.py:186-201):
 *   y1 = (Ta (x) I) x0 + (Tb (x) I) x1,   y2 = (Tc (x) I) x0 + (Td (x) I) x1.
 * coef: 12 arrays of ld doubles (device): for local time row t the multipliers
 * of x0[t-1], x0[t], x0[t+1], x1[t-1], x1[t], x1[t+1] in y1[t], then the same
 * six for y2[t]; zero for t >= n and for neighbours that do not exist.
 * prev / next: the boundary slices of the neighbour ranks (M doubles each,
 * mpi_vector.py:140-187) or NULL.  x0, x1 have pitch ldx, y1, y2 pitch ldy; ld
 * columns are written (pads as zero). */
int stk_time_tridiag_pair(int M, int n, int ld, const double *coef,
                          const double *x0, const double *x1, int ldx,
                          const double *prev0, const double *next0,
                          const double *prev1, const double *next1, double *y1,
                          double *y2, int ldy, void *stream);
/* Two matrices on one sparsity pattern (M_x and A_x, heateq_mpi.py:166-178):
 * split: y0 = A0 x, y1 = A1 x (x and the pattern read once);
 * pair : y = alpha * (A0 x0 + A1 x1) + beta * z  (x0, x1 of pitch ldx). */
int stk_space_spmm_split(int nrows, const int *indptr, const int *indices,
                         const double *vals0, const double *vals1,
                         const double *x, double *y0, double *y1, int ld,
                         void *stream);
int stk_space_spmm_pair(int nrows, const int *indptr, const int *indices,
                        const double *vals0, const double *vals1,
                        const double *x0, const double *x1, int ldx,
                        double alpha, double beta, const double *z, double *y,
                        int ld, void *stream);
/* out[h*M + i] = x[i*ld + tidx[h]]   (pack time slices for a send). */
/* out[i, c_out + c] = in[i, idx ? idx[c] : c_in + c], c < n, i < M: column
 * gather / placement between blocks of pitches ld_in, ld_out (level-wise <->
 * node order of wavelets.py:109-117; the piece placement of permute,
 * mpi_vector.py:212-240).  in must not alias out. */
int stk_copy_cols(int M, int n, const double *in, int ld_in, int c_in,
                  const int *idx, double *out, int ld_out, int c_out,
                  void *stream);
/* x[i, t] = ux[i] * ut[t], t < ld (heateq_mpi.py:189-191: rhs = u0_t (x) u0_x). */
int stk_outer(int M, int ld, const double *ux, const double *ut, double *x,
              void *stream);
int stk_pack_slices(const double *x, int ld, int M, const int *tidx, int n,
                    double *out, void *stream);
/* x[i*ld + tidx[h]] = beta * x[...] + alpha * in[h*M + i]. */
int stk_unpack_slices(double *x, int ld, int M, const int *tidx, int n,
                      const double *in, double alpha, double beta,
                      void *stream);

/* ---- wavelet transform in time, in-place lifting ------------------------
 * wavelets.py:106-134 (WaveletTransformOp._matmat/_rmatmat, interleaved
 * ordering) for a block holding the whole time axis (n_t = 2^J + 1).
 * transpose = 0: x <- W src (levels 1..J); transpose = 1: x <- W^T src.
 * src may be x (in place) or another block of the same pitch. */
int stk_wavelet_lift(int M, int J, int transpose, const double *src,
                     double *x, int ld, void *stream);

/* A chain of nlev sparse level steps along the time axis of the extended
 * column [n_loc local slices | n_halo halo slices] of every space dof, in
 * shared memory (the lifting form of the wavelet transform on a time slab of a
 * P > 1 decomposition: wavelets.py:81-134 on the slices of mpi_vector.py:18-31
 * plus the remote slices of mpi_kron.py:281-283).  Level l changes the rows
 * trow[lev_ptr[l] .. lev_ptr[l+1]); row q is  sum_p tval[p] * col(tcol[p]),
 * p in [tptr[q], tptr[q+1]), evaluated on the values before the level.
 * xh: slice-major halo input (NULL = zeros); y: local result (pads zeroed);
 * yh_out: slice-major (n_halo, M) halo part of the result or NULL. */
int stk_time_chain(int M, int n_loc, int n_halo, int nlev, const int *lev_ptr,
                   const int *trow, const int *tptr, const int *tcol,
                   const double *tval, int R, int NNZ, const double *x,
                   int ldx, const double *xh, double *y, int ldy,
                   double *yh_out, void *stream);

/* ---- multigrid V-cycle, batched over all time slices --------------------
 * multigrid.py:130-197 (MultiGrid), :100-127 (PETSc MatSOR sweeps).
 * A handle describes the Galerkin hierarchies (levels 0..nlevels-1, the last
 * one finest) of G matrices on one sparsity pattern by device pointers the
 * caller keeps alive.  Every time slice t of a block belongs to one of the G
 * GROUPS (group[t]) and is preconditioned with that group's hierarchy: G = 1
 * is K_x = MG(A_x), G = J_time + 1 serves all C_j = MG(2^j M_x + alpha A_x) of
 * heateq_mpi.py:143-153 in one launch sequence. */
typedef struct stk_mg stk_mg;
stk_mg *stk_mg_create(int nlevels, int smoothsteps, int vcycles, int G);
void stk_mg_destroy(stk_mg *mg);
/* Level pattern, the G level matrices (vals: G x nnz in CSR order, diag:
 * G x nrows) and the Gauss-Seidel schedule: sched_rows (device) lists the rows
 * wavefront by wavefront, phase_ptr (HOST, nphases+1 entries) delimits the
 * wavefronts.  Rows inside a wavefront are mutually independent and every row
 * comes after all its lower-numbered neighbours, so running the wavefronts in
 * order is exactly the lexicographic sweep of multigrid.py:89-97 / MatSOR. */
int stk_mg_set_level(stk_mg *mg, int level, int nrows, int nnz,
                     const int *indptr, const int *indices, const double *vals,
                     const double *diag, const int *sched_rows,
                     const int *phase_ptr_host, int nphases);
/* Prolongation level-1 -> level (CSR, nrows(level) x nrows(level-1)) and its
 * transpose (multigrid.py:39-60). */
int stk_mg_set_transfer(stk_mg *mg, int level, const int *p_indptr,
                        const int *p_indices, const double *p_vals,
                        const int *r_indptr, const int *r_indices,
                        const double *r_vals);
/* Doubles of workspace stk_mg_apply needs for blocks of pitch ld. */
int64_t stk_mg_workspace(const stk_mg *mg, int ld);
/* x <- `vcycles` V(nu,nu)-cycles for A_{group[t]} x = b from x = 0, every slice
 * t of the block at once.  group: int[ld] (pads repeat the last slice's group)
 * or NULL when G = 1.  coarse_inv: G dense inverses (n0 x n0, row-major) of
 * the coarsest matrices. */
int stk_mg_apply(stk_mg *mg, const int *group, const double *coarse_inv,
                 const double *b, double *x, int ld, double *ws, void *stream);
/* One family of Gauss-Seidel sweeps on a level with the per-wavefront kernels
 * (exposed for tests): `nsweeps` forward (backward = 0) or backward sweeps. */
int stk_mg_smooth(stk_mg *mg, int level, int nsweeps, int backward,
                  const int *group, const double *f, double *u, int ld,
                  void *stream);

/* HOST helper (host pointers): wave[i] = wavefront of row i in the
 * lexicographic Gauss-Seidel dependency DAG of a symmetric-pattern CSR matrix
 * (0 for rows without lower-numbered neighbours).  Returns the number of
 * wavefronts.  Setup only; replaces nothing in the reference (PETSc sweeps
 * rows one by one). */
int stk_gs_wavefronts(int n, const int *indptr, const int *indices, int *wave);

/* ---- fused Gauss-Seidel smoother -----------------------------------------
 * `nsweeps` lexicographic sweeps of one level (multigrid.py:89-97, MatSOR
 * :113-127) in one pass over HBM.  A program (built on the host by
 * gs_program.compile_program; all array arguments are DEVICE pointers the
 * caller keeps alive, except none) lists, per spatial item and macro-step, the
 * shared-memory window loads and the row updates that may run concurrently;
 * see csrc/stk_gsfused.cu.  item_step / item_pass: nitems + 1 ints;
 * step_info: (end pass, end load) int pairs per macro-step; rec: npasses x
 * ngrp records of recw uint32 (row, prefetch hint, kind or CSR offset, window
 * slot | nnz << 16, then the uint16 window slots of the row's entries);
 * ld: (row, window slot) int pairs; ngrp = row updates a CTA runs side by
 * side (128 or 256: 512 or 1024 threads, 4 lanes per row). */
typedef struct stk_gs_prog stk_gs_prog;
stk_gs_prog *stk_gs_prog_create(int nitems, int nslots, int maxnnz, int generic,
                                int recw, int ngrp, const int *item_step,
                                const int *item_pass, const void *step_info,
                                const void *rec, const void *ld);
void stk_gs_prog_destroy(stk_gs_prog *prog);
/* u_out <- the program's sweeps applied to u_in (NULL: zero initial guess) with
 * right-hand side f; u_in must not alias u_out.  T = time values per CTA (8).
 * Matrix values of the G groups, entries in the program's order (diagonal
 * first): for programs compiled with row kinds the table
 * ktab[G][nkinds][maxnnz + 2] (bulk_kind = the kind most rows have: its values
 * are kept in registers), else cvals[G][vstride] (values of every row at its
 * CSR offset).  group: int[ld] or NULL when G = 1. */
int stk_gs_fused(const stk_gs_prog *prog, int G, int T, const double *ktab,
                 int nkinds, int bulk_kind, const double *cvals,
                 int64_t vstride, const int *group, const double *f,
                 const double *u_in, double *u_out, int ld, void *stream);
/* Attach the fused smoother programs (nu forward / nu backward sweeps) and this
 * handle's values in program order to a level; stk_mg_apply then runs each
 * smoothing phase of that level as one launch (and needs one more block of
 * workspace per fused level, see stk_mg_workspace).  Call before the first
 * stk_mg_apply. */
int stk_mg_set_fused(stk_mg *mg, int level, const stk_gs_prog *fwd,
                     const stk_gs_prog *bwd, const double *ktab, int nkinds,
                     int bulk_kind, const double *cvals, int T, int kstride,
                     const int *kind_of_row, const int *canon_indices);
/* With a kind table (kstride = maxnnz + 2 doubles per kind), kind_of_row[nrows]
 * and canon_indices[nnz] (the level's column indices in the table's entry
 * order, device arrays, may be NULL) also serve the level's grouped residual
 * A_{g(t)} u - f (multigrid.py:174): the G matrices are read from the table in
 * shared memory instead of G value arrays in HBM. */
/* A second pair of programs of the same level, compiled for wide blocks (a
 * CTA's run time does not depend on the slab width, so the best tiling does:
 * few long items when there are many time chunks, many short ones when there
 * are few): stk_mg_apply uses it for blocks of at least min_chunks chunks of T
 * time slices -- the two brackets of the Schur operator side by side
 * (heateq_mpi.py:166-178) -- and the first pair below that. */
int stk_mg_set_fused_wide(stk_mg *mg, int level, const stk_gs_prog *fwd,
                          const stk_gs_prog *bwd, int min_chunks);
/* HOST helper (host pointers): interval colouring of window lifetimes
 * [start[q], end[q]] (macro-steps); slot[q] out; returns the slots used. */
int stk_gs_alloc_slots(int n, const int *start, const int *end, int *slot);

#ifdef __cplusplus
}
#endif
#endif /* STK_H */
