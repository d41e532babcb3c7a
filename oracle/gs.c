/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Plain-C restatement of the reference's lexicographic Gauss-Seidel smoother.
 * The reference calls PETSc MatSOR through petsc4py
 * (/root/reference/source/multigrid.py:113-127, omega = 1, `its` sweeps,
 * non-zero initial guess) and defines the same update in pure Python in
 * `Smoother.PreSmooth/PostSmooth` (/root/reference/source/multigrid.py:89-97):
 *
 *     ax   = A[i,:] . u          (whole row, diagonal included)
 *     u[i] += (f[i] - ax) / a_ii
 *
 * ascending i for the pre-smoother, descending i for the post-smoother.
 * Compiled with -ffp-contract=off so that it is plain IEEE double arithmetic.
 */
#include <stddef.h>

static void sweep(int n, const int *indptr, const int *indices,
                  const double *data, const double *invdiag, const double *f,
                  double *u, int backward)
{
    for (int k = 0; k < n; ++k) {
        int i = backward ? n - 1 - k : k;
        double ax = 0.0;
        for (int p = indptr[i]; p < indptr[i + 1]; ++p)
            ax += data[p] * u[indices[p]];
        u[i] += invdiag[i] * (f[i] - ax);
    }
}

/* `its` sweeps on one vector.  invdiag[i] = 1/a_ii (multigrid.py:87). */
void gs_sweeps(int n, const int *indptr, const int *indices, const double *data,
               const double *invdiag, const double *f, double *u, int its,
               int backward)
{
    for (int s = 0; s < its; ++s)
        sweep(n, indptr, indices, data, invdiag, f, u, backward);
}

/* Same on `nvec` independent vectors stored one after the other (row-major
 * (nvec, n)), which is how the reference visits the time slices of a slab:
 * one MultiGrid._matvec per column. */
void gs_sweeps_multi(int nvec, int n, const int *indptr, const int *indices,
                     const double *data, const double *invdiag, const double *f,
                     double *u, int its, int backward)
{
    for (int v = 0; v < nvec; ++v)
        gs_sweeps(n, indptr, indices, data, invdiag, f + (size_t)v * n,
                  u + (size_t)v * n, its, backward);
}

/* y = A x  for CSR A (scipy's csr_matvec restated; multigrid.py:174,180). */
void csr_matvec(int nrows, const int *indptr, const int *indices,
                const double *data, const double *x, double *y)
{
    for (int i = 0; i < nrows; ++i) {
        double s = 0.0;
        for (int p = indptr[i]; p < indptr[i + 1]; ++p)
            s += data[p] * x[indices[p]];
        y[i] = s;
    }
}
