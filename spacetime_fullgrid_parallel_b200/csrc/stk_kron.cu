// Kronecker applies of libstk: space CSR SpMM batched over time slices, sparse
// time operator along the fast axis, slice pack/unpack for the halo exchange.
// Replaces scipy's csr_matvecs at the call sites of
// /root/reference/source/mpi_kron.py:131,149,194,199-200,250,307-315 and
// linop.py:75-79.
//
// Thread mapping (both kernels): one thread owns a pair of adjacent time
// values (a double2) of one row; consecutive threads walk the time axis first,
// so every load of x[j, :] and every store of y[i, :] is a contiguous
// 16-byte-per-lane access and the CSR row (indices, values) is a warp-uniform
// broadcast load that is read once for all time slices.
#include <stdlib.h>

#include "stk_common.cuh"

namespace stk {

// y[i,t] = alpha * (A x)[i,t] + beta * z[i,t]; K=2: A(t) = c0[t] A0 + c1[t] A1.
// Grid-stride over (row, double2 column) items with a resident grid.
template <int K, bool HAS_Z>
__global__ void __launch_bounds__(256)
    k_space_spmm(int nrows, const int *__restrict__ indptr, const int *__restrict__ indices,
                 const double *__restrict__ vals0, const double *__restrict__ vals1,
                 const double *__restrict__ coef0, const double *__restrict__ coef1,
                 const double *__restrict__ x, double alpha, double beta, const double *z,
                 double *y, int ld, unsigned ld2, const int *__restrict__ rows) {
    const unsigned total = (unsigned)nrows * ld2;
    const unsigned stride = gridDim.x * 256u;
    for (unsigned k = blockIdx.x * 256u + threadIdx.x; k < total; k += stride) {
        const unsigned r = k / ld2;
        unsigned c = (k - r * ld2) * 2u;
        const unsigned i = rows ? (unsigned)__ldg(rows + r) : r;
        int p0 = __ldg(indptr + i), p1 = __ldg(indptr + i + 1);
        double2 s0 = make_double2(0.0, 0.0), s1 = make_double2(0.0, 0.0);
        row_product<K>(p0, p1, indices, vals0, vals1, x, ld, c, s0, s1);
        if (K == 2) {
            double2 c0 = ldg2(coef0 + c), c1 = ldg2(coef1 + c);
            s0.x = fma(c0.x, s0.x, c1.x * s1.x);
            s0.y = fma(c0.y, s0.y, c1.y * s1.y);
        }
        size_t o = (size_t)i * ld + c;
        double2 out;
        if (HAS_Z) {
            double2 zv = ldv2(z + o);
            out.x = fma(alpha, s0.x, beta * zv.x);
            out.y = fma(alpha, s0.y, beta * zv.y);
        } else {
            out.x = alpha * s0.x;
            out.y = alpha * s0.y;
        }
        stv2(y + o, out);
    }
}

// Single-matrix form with FOUR time values per thread (256-bit loads and
// stores): the broadcast loads of the CSR row (index + value per entry) are
// shared by twice as many outputs, which is what bounds these kernels once the
// row schedule has removed the DRAM re-reads (L1 data path 72 % busy with two
// values per thread, profiles/r2c_ncu_summary.md).  Needs ld % 4 == 0 and
// 32-byte aligned blocks (the pitch rule; checked by the callers).
template <bool HAS_Z>
__global__ void __launch_bounds__(256)
    k_space_spmm4(int nrows, const int *__restrict__ indptr, const int *__restrict__ indices,
                  const double *__restrict__ vals, const double *__restrict__ x, double alpha,
                  double beta, const double *z, double *y, int ld, unsigned ld4,
                  const int *__restrict__ rows) {
    const unsigned total = (unsigned)nrows * ld4;
    const unsigned stride = gridDim.x * 256u;
    for (unsigned k = blockIdx.x * 256u + threadIdx.x; k < total; k += stride) {
        const unsigned r = k / ld4;
        const unsigned c = (k - r * ld4) * 4u;
        const unsigned i = rows ? (unsigned)__ldg(rows + r) : r;
        const int p1 = __ldg(indptr + i + 1);
        double4v s = {0.0, 0.0, 0.0, 0.0};
        for (int p = __ldg(indptr + i); p < p1; ++p) {
            const double4v xv = ldv4(x + (size_t)__ldg(indices + p) * ld + c);
            fma4(__ldg(vals + p), xv, s);
        }
        const size_t o = (size_t)i * ld + c;
        double4v out;
        if (HAS_Z) {
            const double4v zv = ldv4(z + o);
            out.x = fma(alpha, s.x, beta * zv.x);
            out.y = fma(alpha, s.y, beta * zv.y);
            out.z = fma(alpha, s.z, beta * zv.z);
            out.w = fma(alpha, s.w, beta * zv.w);
        } else {
            out.x = alpha * s.x;
            out.y = alpha * s.y;
            out.z = alpha * s.z;
            out.w = alpha * s.w;
        }
        stv4(y + o, out);
    }
}

// The same with one matrix per GROUP of time slices on a shared pattern:
// vals[g * vstride + p] is entry p of group g's matrix, grp[t] the group of
// time slice t (multigrid levels of the C_j = MG(2^j M_x + alpha A_x) family,
// heateq_mpi.py:143-153: every slice sees exactly the matrix the reference
// forms for it).  A thread's two time values may belong to two groups.
template <bool HAS_Z>
__global__ void __launch_bounds__(256)
    k_space_spmm_g(int nrows, const int *__restrict__ indptr, const int *__restrict__ indices,
                   const double *__restrict__ vals, size_t vstride, const int *__restrict__ grp,
                   const double *__restrict__ x, double alpha, double beta, const double *z,
                   double *y, int ld, unsigned ld2, const int *__restrict__ rows) {
    const unsigned total = (unsigned)nrows * ld2;
    const unsigned stride = gridDim.x * 256u;
    for (unsigned k = blockIdx.x * 256u + threadIdx.x; k < total; k += stride) {
        const unsigned r = k / ld2;
        unsigned c = (k - r * ld2) * 2u;
        const unsigned i = rows ? (unsigned)__ldg(rows + r) : r;
        const double *v0 = vals + (size_t)__ldg(grp + c) * vstride;
        const double *v1 = vals + (size_t)__ldg(grp + c + 1) * vstride;
        int p1 = __ldg(indptr + i + 1);
        double2 s = make_double2(0.0, 0.0);
        for (int p = __ldg(indptr + i); p < p1; ++p) {
            double2 xv = ldv2(x + (size_t)__ldg(indices + p) * ld + c);
            s.x = fma(__ldg(v0 + p), xv.x, s.x);
            s.y = fma(__ldg(v1 + p), xv.y, s.y);
        }
        size_t o = (size_t)i * ld + c;
        double2 out;
        if (HAS_Z) {
            double2 zv = ldv2(z + o);
            out.x = fma(alpha, s.x, beta * zv.x);
            out.y = fma(alpha, s.y, beta * zv.y);
        } else {
            out.x = alpha * s.x;
            out.y = alpha * s.y;
        }
        stv2(y + o, out);
    }
}

// Grouped SpMM through ROW KINDS (gs_program.row_kinds): on a uniformly refined
// mesh the rows of a level carry a handful of distinct value sequences, so the
// G group matrices are a table tab[g][kind][entry] of a few KB in shared
// memory instead of G value arrays in HBM.  k_space_spmm_g fetches one matrix
// value per (entry, time value) through L1 -- with one group per wavelet
// level, the 64 columns a warp covers hit ~9 arrays per entry and those loads,
// not x, bound the kernel (3.2 ms for the finest-level residual against 1.2 ms
// single-matrix).  Here a thread owns four time values (256-bit x loads), reads
// the row's kind once and its four groups' values from the table.  cidx lists
// the column indices in the table's entry order (diagonal first).
template <bool HAS_Z>
__global__ void __launch_bounds__(256)
    k_space_spmm_gk(int nrows, const int *__restrict__ indptr, const int *__restrict__ cidx,
                    const int *__restrict__ kind_of_row, const double *__restrict__ ktab, int G,
                    int nkinds, int kstride, const int *__restrict__ grp,
                    const double *__restrict__ x, double alpha, double beta, const double *z,
                    double *y, int ld, unsigned ld4, const int *__restrict__ rows) {
    extern __shared__ double s_tab[];
    const unsigned ks = (kstride & 1) ? kstride : kstride + 1;  // odd: banks spread
    for (int k = threadIdx.x; k < G * nkinds * kstride; k += 256) {
        const int row = k / kstride;
        s_tab[row * ks + (k - row * kstride)] = __ldg(ktab + k);
    }
    __syncthreads();
    const unsigned total = (unsigned)nrows * ld4;
    const unsigned stride = gridDim.x * 256u;
    for (unsigned k = blockIdx.x * 256u + threadIdx.x; k < total; k += stride) {
        const unsigned r = k / ld4;
        const unsigned c = (k - r * ld4) * 4u;
        const unsigned i = rows ? (unsigned)__ldg(rows + r) : r;
        const int4 g = __ldg(reinterpret_cast<const int4 *>(grp + c));
        const unsigned kind = (unsigned)__ldg(kind_of_row + i);
        const double *t0 = s_tab + ((unsigned)g.x * nkinds + kind) * ks;
        const double *t1 = s_tab + ((unsigned)g.y * nkinds + kind) * ks;
        const double *t2 = s_tab + ((unsigned)g.z * nkinds + kind) * ks;
        const double *t3 = s_tab + ((unsigned)g.w * nkinds + kind) * ks;
        const int p0 = __ldg(indptr + i), n = __ldg(indptr + i + 1) - p0;
        double4v s = {0.0, 0.0, 0.0, 0.0};
        for (int e = 0; e < n; ++e) {
            const double4v xv = ldv4(x + (size_t)__ldg(cidx + p0 + e) * ld + c);
            s.x = fma(t0[e], xv.x, s.x);
            s.y = fma(t1[e], xv.y, s.y);
            s.z = fma(t2[e], xv.z, s.z);
            s.w = fma(t3[e], xv.w, s.w);
        }
        const size_t o = (size_t)i * ld + c;
        double4v out;
        if (HAS_Z) {
            const double4v zv = ldv4(z + o);
            out.x = fma(alpha, s.x, beta * zv.x);
            out.y = fma(alpha, s.y, beta * zv.y);
            out.z = fma(alpha, s.z, beta * zv.z);
            out.w = fma(alpha, s.w, beta * zv.w);
        } else {
            out.x = alpha * s.x;
            out.y = alpha * s.y;
            out.z = alpha * s.z;
            out.w = alpha * s.w;
        }
        stv4(y + o, out);
    }
}

// One row of T against the time column of space dof i.
__device__ __forceinline__ double time_row(int t, const int *__restrict__ indptr,
                                           const int *__restrict__ indices,
                                           const double *__restrict__ vals,
                                           const double *__restrict__ xi, int ncols_local,
                                           const double *__restrict__ xh, int M, unsigned i) {
    double s = 0.0;
    int p1 = __ldg(indptr + t + 1);
    for (int p = __ldg(indptr + t); p < p1; ++p) {
        int c = __ldg(indices + p);
        double xv = (c < ncols_local) ? xi[c] : __ldg(xh + (size_t)(c - ncols_local) * M + i);
        s = fma(__ldg(vals + p), xv, s);
    }
    return s;
}

// y[i,t] = alpha * sum_p T[t,p] X(i, col_p) + beta * y[i,t].
// One thread per (i, pair of adjacent t); t fastest; grid-stride.  X(i, c)
// comes from the block for local columns and from the slice-major halo buffer
// otherwise.  Pads t >= nrows_t are written as zero (kept when accumulating).
template <bool ACC>
__global__ void __launch_bounds__(256)
    k_time_apply(int M, int nrows_t, const int *__restrict__ indptr,
                 const int *__restrict__ indices, const double *__restrict__ vals,
                 const double *__restrict__ x, int ldx, int ncols_local,
                 const double *__restrict__ xh, double alpha, double beta,
                 double *__restrict__ y, int ldy) {
    const unsigned ld2 = (unsigned)ldy / 2u;
    const unsigned total = (unsigned)M * ld2;
    const unsigned stride = gridDim.x * 256u;
    for (unsigned k = blockIdx.x * 256u + threadIdx.x; k < total; k += stride) {
        unsigned i = k / ld2;
        int t = (int)(k - i * ld2) * 2;
        size_t o = (size_t)i * ldy + t;
        const double *xi = x + (size_t)i * ldx;
        if (ACC) {  // rows of T without entries leave y untouched (G_t = e0 e0^T)
            bool e0 = t >= nrows_t || __ldg(indptr + t) == __ldg(indptr + t + 1);
            bool e1 = t + 1 >= nrows_t || __ldg(indptr + t + 1) == __ldg(indptr + t + 2);
            if (e0 && e1 && beta == 1.0) continue;
        }
        double2 out = make_double2(0.0, 0.0);
        if (t < nrows_t)
            out.x = alpha * time_row(t, indptr, indices, vals, xi, ncols_local, xh, M, i);
        if (t + 1 < nrows_t)
            out.y = alpha * time_row(t + 1, indptr, indices, vals, xi, ncols_local, xh, M, i);
        if (ACC) {
            double2 old = ldv2(y + o);
            out.x = fma(beta, old.x, out.x);
            out.y = fma(beta, old.y, out.y);
        }
        stv2(y + o, out);
    }
}

// Two-input form: y = alpha * [Ta Tb] [x0; x1] + beta * y.  Column c of the
// stacked matrix addresses x0 (c < n), x1 (c < 2n), the halo of x0, then the
// halo of x1.  Serves the brackets (A_t (x) M + L_t (x) A) x of the regrouped
// Schur operator in one pass over M x and A x.
__device__ __forceinline__ double time_row2(int t, const int *__restrict__ indptr,
                                            const int *__restrict__ indices,
                                            const double *__restrict__ vals,
                                            const double *__restrict__ x0i,
                                            const double *__restrict__ x1i, int n, int nh0,
                                            const double *__restrict__ xh0,
                                            const double *__restrict__ xh1, int M, unsigned i) {
    double s = 0.0;
    int p1 = __ldg(indptr + t + 1);
    for (int p = __ldg(indptr + t); p < p1; ++p) {
        int c = __ldg(indices + p);
        double xv;
        if (c < n) xv = x0i[c];
        else if (c < 2 * n) xv = x1i[c - n];
        else if (c < 2 * n + nh0) xv = __ldg(xh0 + (size_t)(c - 2 * n) * M + i);
        else xv = __ldg(xh1 + (size_t)(c - 2 * n - nh0) * M + i);
        s = fma(__ldg(vals + p), xv, s);
    }
    return s;
}

template <bool ACC>
__global__ void __launch_bounds__(256)
    k_time_apply2(int M, int nrows_t, const int *__restrict__ indptr,
                  const int *__restrict__ indices, const double *__restrict__ vals,
                  const double *__restrict__ x0, const double *__restrict__ x1, int ldx, int n,
                  const double *__restrict__ xh0, int nh0, const double *__restrict__ xh1,
                  double alpha, double beta, double *__restrict__ y, int ldy, int wy) {
    const unsigned ld2 = (unsigned)wy / 2u;  // columns written per row (<= pitch ldy)
    const unsigned total = (unsigned)M * ld2;
    const unsigned stride = gridDim.x * 256u;
    for (unsigned k = blockIdx.x * 256u + threadIdx.x; k < total; k += stride) {
        unsigned i = k / ld2;
        int t = (int)(k - i * ld2) * 2;
        size_t o = (size_t)i * ldy + t;
        const double *x0i = x0 + (size_t)i * ldx, *x1i = x1 + (size_t)i * ldx;
        double2 out = make_double2(0.0, 0.0);
        if (t < nrows_t)
            out.x = alpha * time_row2(t, indptr, indices, vals, x0i, x1i, n, nh0, xh0, xh1, M, i);
        if (t + 1 < nrows_t)
            out.y = alpha *
                    time_row2(t + 1, indptr, indices, vals, x0i, x1i, n, nh0, xh0, xh1, M, i);
        if (ACC) {
            double2 old = ldv2(y + o);
            out.x = fma(beta, old.x, out.x);
            out.y = fma(beta, old.y, out.y);
        }
        stv2(y + o, out);
    }
}

// Both brackets of the regrouped Schur operator in one pass
// (heateq_mpi.py:166-178):
//     y1 = (Ta (x) I) x0 + (Tb (x) I) x1,     y2 = (Tc (x) I) x0 + (Td (x) I) x1
// for four TRIDIAGONAL time matrices (A_t, L_t, L_t^T, M_t).  A thread owns one
// pair of adjacent time values for the whole kernel, so its 24 stencil
// coefficients live in registers and a CTA walks the space dofs: per dof a
// thread loads one double2 of x0 and of x1, gets the two neighbouring values
// from the adjacent lanes (shuffles; warp edges re-read them), and stores one
// double2 of each result: 32 B per dof of traffic and no matrix loads at all,
// where two passes of the general k_time_apply2 move 48 B and are bound by the
// L1 data path (18 broadcast loads of CSR entries per output).
// coef[k][t], k = 0..11: multipliers of x0[t-1], x0[t], x0[t+1], x1[t-1],
// x1[t], x1[t+1] for y1[t], then the same for y2[t]; zero for t >= n and for
// neighbours that do not exist.  prev/next: the neighbour ranks' boundary
// slices (M doubles each) or NULL.
template <int NT>
__global__ void __launch_bounds__(NT)
    k_time_tridiag_pair(int M, int n, int ld2, const double *__restrict__ coef, int ldc,
                        const double *__restrict__ x0, const double *__restrict__ x1, int ldx,
                        const double *__restrict__ prev0, const double *__restrict__ next0,
                        const double *__restrict__ prev1, const double *__restrict__ next1,
                        double *__restrict__ y1, double *__restrict__ y2, int ldy) {
    const int cp = threadIdx.x, lane = threadIdx.x & 31;
    const bool live = cp < ld2;
    const int t = 2 * cp;
    double2 c[12];
#pragma unroll
    for (int k = 0; k < 12; ++k)
        c[k] = live ? ldg2(coef + (size_t)k * ldc + t) : make_double2(0.0, 0.0);
    for (int i = blockIdx.x; i < M; i += gridDim.x) {
        const double *r0 = x0 + (size_t)i * ldx, *r1 = x1 + (size_t)i * ldx;
        double2 a = live ? ldv2(r0 + t) : make_double2(0.0, 0.0);
        double2 b = live ? ldv2(r1 + t) : make_double2(0.0, 0.0);
        if (t + 1 == n) {  // the slice after the last local one is the neighbour's
            a.y = next0 ? __ldg(next0 + i) : 0.0;
            b.y = next1 ? __ldg(next1 + i) : 0.0;
        }
        double al = __shfl_up_sync(0xffffffffu, a.y, 1), ar = __shfl_down_sync(0xffffffffu, a.x, 1);
        double bl = __shfl_up_sync(0xffffffffu, b.y, 1), br = __shfl_down_sync(0xffffffffu, b.x, 1);
        if (lane == 0) {
            al = cp > 0 ? (live ? r0[t - 1] : 0.0) : (prev0 ? __ldg(prev0 + i) : 0.0);
            bl = cp > 0 ? (live ? r1[t - 1] : 0.0) : (prev1 ? __ldg(prev1 + i) : 0.0);
        }
        if (t + 2 == n) {
            ar = next0 ? __ldg(next0 + i) : 0.0;
            br = next1 ? __ldg(next1 + i) : 0.0;
        } else if (lane == 31) {
            ar = (t + 2 < n) ? r0[t + 2] : 0.0;
            br = (t + 2 < n) ? r1[t + 2] : 0.0;
        }
        if (!live) continue;
        double2 o1, o2;
        o1.x = fma(c[5].x, b.y, fma(c[4].x, b.x, fma(c[3].x, bl,
               fma(c[2].x, a.y, fma(c[1].x, a.x, c[0].x * al)))));
        o1.y = fma(c[5].y, br, fma(c[4].y, b.y, fma(c[3].y, b.x,
               fma(c[2].y, ar, fma(c[1].y, a.y, c[0].y * a.x)))));
        o2.x = fma(c[11].x, b.y, fma(c[10].x, b.x, fma(c[9].x, bl,
               fma(c[8].x, a.y, fma(c[7].x, a.x, c[6].x * al)))));
        o2.y = fma(c[11].y, br, fma(c[10].y, b.y, fma(c[9].y, b.x,
               fma(c[8].y, ar, fma(c[7].y, a.y, c[6].y * a.x)))));
        stv2(y1 + (size_t)i * ldy + t, o1);
        stv2(y2 + (size_t)i * ldy + t, o2);
    }
}

// y0 = A0 x and y1 = A1 x for two matrices on one sparsity pattern: x and the
// pattern are read once (M_x x and A_x x of heateq_mpi.py:166-178).
__global__ void __launch_bounds__(256)
    k_space_spmm_split(int nrows, const int *__restrict__ indptr, const int *__restrict__ indices,
                       const double *__restrict__ vals0, const double *__restrict__ vals1,
                       const double *__restrict__ x, double *__restrict__ y0,
                       double *__restrict__ y1, int ld, unsigned ld2,
                       const int *__restrict__ rows) {
    const unsigned total = (unsigned)nrows * ld2;
    const unsigned stride = gridDim.x * 256u;
    for (unsigned k = blockIdx.x * 256u + threadIdx.x; k < total; k += stride) {
        const unsigned r = k / ld2;
        unsigned c = (k - r * ld2) * 2u;
        const unsigned i = rows ? (unsigned)__ldg(rows + r) : r;
        int p0 = __ldg(indptr + i), p1 = __ldg(indptr + i + 1);
        double2 s0 = make_double2(0.0, 0.0), s1 = make_double2(0.0, 0.0);
        row_product<2>(p0, p1, indices, vals0, vals1, x, ld, c, s0, s1);
        size_t o = (size_t)i * ld + c;
        stv2(y0 + o, s0);
        stv2(y1 + o, s1);
    }
}

// y = alpha * (A0 x0 + A1 x1) + beta * z on one sparsity pattern
// ((I (x) M) z1 + (I (x) A) z2 of the regrouped Schur operator).
template <bool HAS_Z>
__global__ void __launch_bounds__(256)
    k_space_spmm_pair(int nrows, const int *__restrict__ indptr, const int *__restrict__ indices,
                      const double *__restrict__ vals0, const double *__restrict__ vals1,
                      const double *__restrict__ x0, const double *__restrict__ x1, int ldx,
                      double alpha, double beta, const double *z, double *y, int ld,
                      unsigned ld2, const int *__restrict__ rows) {
    const unsigned total = (unsigned)nrows * ld2;
    const unsigned stride = gridDim.x * 256u;
    for (unsigned k = blockIdx.x * 256u + threadIdx.x; k < total; k += stride) {
        const unsigned r = k / ld2;
        unsigned c = (k - r * ld2) * 2u;
        const unsigned i = rows ? (unsigned)__ldg(rows + r) : r;
        int p1 = __ldg(indptr + i + 1);
        double2 s = make_double2(0.0, 0.0);
        for (int p = __ldg(indptr + i); p < p1; ++p) {
            size_t off = (size_t)__ldg(indices + p) * ldx + c;
            double2 a = ldv2(x0 + off), b = ldv2(x1 + off);
            double m = __ldg(vals0 + p), w = __ldg(vals1 + p);
            s.x = fma(m, a.x, fma(w, b.x, s.x));
            s.y = fma(m, a.y, fma(w, b.y, s.y));
        }
        size_t o = (size_t)i * ld + c;
        double2 out;
        if (HAS_Z) {
            double2 zv = ldv2(z + o);
            out.x = fma(alpha, s.x, beta * zv.x);
            out.y = fma(alpha, s.y, beta * zv.y);
        } else {
            out.x = alpha * s.x;
            out.y = alpha * s.y;
        }
        stv2(y + o, out);
    }
}

// Shared-memory form for time matrices with many nonzeros per row (the
// multi-level wavelet transform has ~2J+1): the whole CSR matrix and a panel
// of RB time columns (local slices + halo slices) are staged in shared memory,
// so the ~19 loads per output hit shared memory instead of L1/L2, and the
// panel is read from HBM once with coalesced loads.  Persistent CTAs walk the
// panels.  Shared layout: vals[nnz] | xs[RB][W] | indptr[nrows_t+1] | idx[nnz].
template <bool ACC>
__global__ void __launch_bounds__(256)
    k_time_apply_smem(int M, int nrows_t, int nnz, const int *__restrict__ indptr,
                      const int *__restrict__ indices, const double *__restrict__ vals,
                      const double *__restrict__ x, int ldx, int ncols_local,
                      const double *__restrict__ xh, int n_halo, double alpha, double beta,
                      double *__restrict__ y, int ldy, int RB, int W) {
    extern __shared__ double smem_d[];
    double *s_val = smem_d;
    double *xs = s_val + nnz;
    int *s_ptr = reinterpret_cast<int *>(xs + (size_t)RB * W);
    int *s_idx = s_ptr + nrows_t + 1;
    for (int k = threadIdx.x; k < nnz; k += 256) {
        s_val[k] = __ldg(vals + k);
        s_idx[k] = __ldg(indices + k);
    }
    for (int k = threadIdx.x; k <= nrows_t; k += 256) s_ptr[k] = __ldg(indptr + k);
    const int ntiles = (M + RB - 1) / RB;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i0 = tile * RB;
        const int rb = min(RB, M - i0);
        __syncthreads();  // matrix staged / previous panel consumed
        for (int e = threadIdx.x; e < rb * ncols_local; e += 256) {
            int r = e / ncols_local, c = e - r * ncols_local;
            xs[r * W + c] = x[(size_t)(i0 + r) * ldx + c];
        }
        for (int e = threadIdx.x; e < rb * n_halo; e += 256) {
            int h = e / rb, r = e - h * rb;
            xs[r * W + ncols_local + h] = __ldg(xh + (size_t)h * M + i0 + r);
        }
        __syncthreads();
        // each thread: one time row t for RR panel columns (the matrix entry
        // is read once from shared memory and reused RR times)
        constexpr int RR = 4;
        const int ngroups = (rb + RR - 1) / RR;
        for (int e = threadIdx.x; e < ngroups * ldy; e += 256) {
            int g = e / ldy, t = e - g * ldy;
            int r0 = g * RR;
            double acc[RR];
#pragma unroll
            for (int k = 0; k < RR; ++k) acc[k] = 0.0;
            if (t < nrows_t) {
                const double *xr = xs + r0 * W;
                for (int p = s_ptr[t]; p < s_ptr[t + 1]; ++p) {
                    double v = s_val[p];
                    int c = s_idx[p];
#pragma unroll
                    for (int k = 0; k < RR; ++k) acc[k] = fma(v, xr[k * W + c], acc[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < RR; ++k) {
                if (r0 + k >= rb) break;
                size_t o = (size_t)(i0 + r0 + k) * ldy + t;
                if (t >= nrows_t) {
                    if (!ACC) y[o] = 0.0;
                } else {
                    y[o] = ACC ? fma(alpha, acc[k], beta * y[o]) : alpha * acc[k];
                }
            }
        }
    }
}

// out[i, c_out + c] = in[i, src(c)] for c < n: src(c) = idx[c] (a gather along
// the time axis: level-wise <-> node order of wavelets.py:109-117) or c_in + c
// (the piece placement of the time <-> space re-sharding, mpi_vector.py:
// 212-240).  Threads walk c first: contiguous stores, near-contiguous loads.
__global__ void __launch_bounds__(256)
    k_copy_cols(int M, int n, const double *__restrict__ in, int ld_in, int c_in,
                const int *__restrict__ idx, double *__restrict__ out, int ld_out, int c_out) {
    const size_t total = (size_t)M * n;
    const size_t stride = (size_t)gridDim.x * 256u;
    for (size_t k = (size_t)blockIdx.x * 256u + threadIdx.x; k < total; k += stride) {
        const size_t i = k / n;
        const int c = (int)(k - i * n);
        const int src = idx ? __ldg(idx + c) : c_in + c;
        out[i * ld_out + c_out + c] = in[i * ld_in + src];
    }
}

// x[i, t] = ux[i] * ut[t] (t < ld; ut holds zeros on the pads): the Kronecker
// right-hand side of heateq_mpi.py:189-191.
__global__ void __launch_bounds__(256)
    k_outer(int M, int ld, const double *__restrict__ ux, const double *__restrict__ ut,
            double *__restrict__ x) {
    const size_t total = (size_t)M * ld;
    const size_t stride = (size_t)gridDim.x * 256u;
    for (size_t k = (size_t)blockIdx.x * 256u + threadIdx.x; k < total; k += stride) {
        const size_t i = k / ld;
        x[k] = __ldg(ux + i) * __ldg(ut + (k - i * ld));
    }
}

__global__ void __launch_bounds__(256)
    k_pack_slices(const double *__restrict__ x, int ld, int M, const int *__restrict__ tidx,
                  int n, double *__restrict__ out) {
    // threads walk i (coalesced store); the strided gather of x is a few
    // slices per block only.
    size_t k = (size_t)blockIdx.x * 256u + threadIdx.x;
    if (k >= (size_t)n * M) return;
    int h = (int)(k / M);
    int i = (int)(k - (size_t)h * M);
    out[k] = x[(size_t)i * ld + __ldg(tidx + h)];
}

__global__ void __launch_bounds__(256)
    k_unpack_slices(double *__restrict__ x, int ld, int M, const int *__restrict__ tidx, int n,
                    const double *__restrict__ in, double alpha, double beta) {
    size_t k = (size_t)blockIdx.x * 256u + threadIdx.x;
    if (k >= (size_t)n * M) return;
    int h = (int)(k / M);
    int i = (int)(k - (size_t)h * M);
    size_t o = (size_t)i * ld + __ldg(tidx + h);
    x[o] = (beta == 0.0) ? alpha * in[k] : fma(alpha, in[k], beta * x[o]);
}

// Row schedules (stk_csr_set_row_order): the space SpMMs are free to walk the
// rows of a matrix in any order; with a hierarchical FE numbering the index
// order visits the mesh several times over (one pass per vertex class), so
// every x row is fetched from DRAM once per pass.  A locality order registered
// for the CSR structure (keyed by its row-pointer array) makes the kernels
// walk the mesh once: neighbouring rows are re-read from L2.
struct RowOrder {
    int nrows;
    const int *order;
};
static std::unordered_map<const void *, RowOrder> g_row_orders;

const int *row_order_for(const int *indptr, int nrows) {
    if (g_row_orders.empty()) return nullptr;
    auto it = g_row_orders.find((const void *)indptr);
    return (it != g_row_orders.end() && it->second.nrows == nrows) ? it->second.order : nullptr;
}

int launch_space_spmm(int nrows, const int *indptr, const int *indices, int K,
                      const double *vals0, const double *vals1, const double *coef0,
                      const double *coef1, const double *x, double alpha, double beta,
                      const double *z, double *y, int ld, cudaStream_t s) {
    if (nrows == 0) return 0;
    unsigned ld2 = (unsigned)ld / 2u;
    int64_t work = (int64_t)nrows * ld2;
    if (work >= STK_MAX_ITEMS) return fail(-2, "stk_space_spmm: block too large for 32-bit grid");
    bool has_z = (beta != 0.0);
    if (has_z && z == nullptr) return fail(-1, "stk_space_spmm: beta != 0 needs z");
#define STK_SPMM(KK, ZZ)                                                                     \
    k_space_spmm<KK, ZZ><<<resident_grid(k_space_spmm<KK, ZZ>, 256, work), 256, 0, s>>>(     \
        nrows, indptr, indices, vals0, vals1, coef0, coef1, x, alpha, beta, z, y, ld, ld2, rows)
    const int *rows = row_order_for(indptr, nrows);
    static const bool wide = [] {  // STK_SPMM_WIDE=0: two values per thread
        const char *e = getenv("STK_SPMM_WIDE");
        return !(e && e[0] == '0');
    }();
    const bool aligned = (((uintptr_t)x | (uintptr_t)y | (uintptr_t)z) & 31u) == 0;
    if (K == 1 && wide && aligned && (ld & 3) == 0) {
        const unsigned ld4 = (unsigned)ld / 4u;
        const int64_t work4 = (int64_t)nrows * ld4;
        if (has_z)
            k_space_spmm4<true><<<resident_grid(k_space_spmm4<true>, 256, work4), 256, 0, s>>>(
                nrows, indptr, indices, vals0, x, alpha, beta, z, y, ld, ld4, rows);
        else
            k_space_spmm4<false><<<resident_grid(k_space_spmm4<false>, 256, work4), 256, 0, s>>>(
                nrows, indptr, indices, vals0, x, alpha, beta, z, y, ld, ld4, rows);
        return check_launch("k_space_spmm4");
    }
    if (K == 1) {
        if (has_z) STK_SPMM(1, true); else STK_SPMM(1, false);
    } else {
        if (has_z) STK_SPMM(2, true); else STK_SPMM(2, false);
    }
#undef STK_SPMM
    return check_launch("k_space_spmm");
}

int launch_space_spmm_grouped(int nrows, const int *indptr, const int *indices,
                              const double *vals, size_t vstride, const int *grp, const double *x,
                              double alpha, double beta, const double *z, double *y, int ld,
                              cudaStream_t s) {
    if (!grp)  // one group: the plain single-matrix kernel
        return launch_space_spmm(nrows, indptr, indices, 1, vals, nullptr, nullptr, nullptr, x,
                                 alpha, beta, z, y, ld, s);
    if (nrows == 0) return 0;
    unsigned ld2 = (unsigned)ld / 2u;
    int64_t work = (int64_t)nrows * ld2;
    if (work >= STK_MAX_ITEMS) return fail(-2, "stk_space_spmm: block too large for 32-bit grid");
    if (beta != 0.0 && z == nullptr) return fail(-1, "stk_space_spmm: beta != 0 needs z");
    const int *rows = row_order_for(indptr, nrows);
    if (beta != 0.0)
        k_space_spmm_g<true><<<resident_grid(k_space_spmm_g<true>, 256, work), 256, 0, s>>>(
            nrows, indptr, indices, vals, vstride, grp, x, alpha, beta, z, y, ld, ld2, rows);
    else
        k_space_spmm_g<false><<<resident_grid(k_space_spmm_g<false>, 256, work), 256, 0, s>>>(
            nrows, indptr, indices, vals, vstride, grp, x, alpha, beta, z, y, ld, ld2, rows);
    return check_launch("k_space_spmm_g");
}

// Grouped SpMM with the groups' values given as a row-kind table; returns 1 if
// the arguments do not allow it (the caller then uses the value arrays).
int launch_space_spmm_kinds(int nrows, const int *indptr, const int *cidx,
                            const int *kind_of_row, const double *ktab, int G, int nkinds,
                            int kstride, const int *grp, const double *x, double alpha,
                            double beta, const double *z, double *y, int ld, cudaStream_t s) {
    if (!grp || !cidx || !kind_of_row || !ktab || (ld & 3)) return 1;
    if ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)z) & 31u) || ((uintptr_t)grp & 15u)) return 1;
    const size_t smem = sizeof(double) * (size_t)G * nkinds * (kstride | 1);
    if (smem > 40 * 1024) return 1;
    if (nrows == 0) return 0;
    const unsigned ld4 = (unsigned)ld / 4u;
    const int64_t work = (int64_t)nrows * ld4;
    if (work >= STK_MAX_ITEMS) return fail(-2, "stk_space_spmm: block too large for 32-bit grid");
    if (beta != 0.0 && z == nullptr) return fail(-1, "stk_space_spmm: beta != 0 needs z");
    const int *rows = row_order_for(indptr, nrows);
    if (beta != 0.0)
        k_space_spmm_gk<true><<<resident_grid(k_space_spmm_gk<true>, 256, work), 256, smem, s>>>(
            nrows, indptr, cidx, kind_of_row, ktab, G, nkinds, kstride, grp, x, alpha, beta, z, y,
            ld, ld4, rows);
    else
        k_space_spmm_gk<false><<<resident_grid(k_space_spmm_gk<false>, 256, work), 256, smem, s>>>(
            nrows, indptr, cidx, kind_of_row, ktab, G, nkinds, kstride, grp, x, alpha, beta, z, y,
            ld, ld4, rows);
    return check_launch("k_space_spmm_gk");
}

}  // namespace stk

using namespace stk;

extern "C" {

int stk_csr_set_row_order(const int *indptr, int nrows, const int *order) {
    if (!indptr) return fail(-1, "stk_csr_set_row_order: null row pointer");
    if (!order)
        g_row_orders.erase((const void *)indptr);
    else
        g_row_orders[(const void *)indptr] = RowOrder{nrows, order};
    return 0;
}

int stk_space_spmm(int nrows, const int *indptr, const int *indices, int K, const double *vals0,
                   const double *vals1, const double *coef0, const double *coef1,
                   const double *x, double alpha, double beta, const double *z, double *y, int ld,
                   void *stream) {
    if (K != 1 && K != 2) return fail(-1, "stk_space_spmm: K must be 1 or 2");
    if (ld & 3) return fail(-1, "stk_space_spmm: pitch must be a multiple of 4");
    if (K == 2 && (!vals1 || !coef0 || !coef1))
        return fail(-1, "stk_space_spmm: K = 2 needs vals1, coef0, coef1");
    if (x == y) return fail(-1, "stk_space_spmm: x must not alias y");
    return launch_space_spmm(nrows, indptr, indices, K, vals0, vals1, coef0, coef1, x, alpha,
                             beta, z, y, ld, as_stream(stream));
}

int stk_time_apply(int M, int nrows_t, int nnz, const int *indptr, const int *indices,
                   const double *vals, const double *x, int ldx, int ncols_local,
                   const double *xh, int n_halo, double alpha, double beta, double *y, int ldy,
                   void *stream) {
    if (x == y) return fail(-1, "stk_time_apply: x must not alias y");
    if (nrows_t > ldy) return fail(-1, "stk_time_apply: nrows_t exceeds the pitch of y");
    if (M == 0) return 0;
    if (ldy & 1) return fail(-1, "stk_time_apply: pitch of y must be even");
    int64_t work = (int64_t)M * (ldy / 2);
    if (work >= STK_MAX_ITEMS) return fail(-2, "stk_time_apply: block too large");
    cudaStream_t s = as_stream(stream);
    // Dense-ish time matrices (> 4 nonzeros per row on average) go through the
    // shared-memory kernel when the matrix and a panel of >= 8 columns fit.
    const int W = (ncols_local + n_halo) | 1;  // odd pitch: conflict-free column walks
    const size_t mat_bytes = (size_t)nnz * 12 + (size_t)(nrows_t + 1) * 4 + 16;
    // 72 KB leaves room for 3 CTAs per SM; larger matrices take one CTA's 200 KB.
    size_t budget = 72 * 1024;
    int ctas_per_sm = 3;
    if (mat_bytes + (size_t)8 * W * 8 > budget) {
        budget = 200 * 1024;
        ctas_per_sm = 1;
    }
    if (nnz > 4 * (int64_t)nrows_t && mat_bytes + (size_t)8 * W * 8 <= budget) {
        int RB = (int)((budget - mat_bytes) / ((size_t)W * 8));
        if (RB > 64) RB = 64;
        RB &= ~3;  // panels are consumed four columns at a time
        size_t smem = mat_bytes + (size_t)RB * W * 8;
        int64_t ntiles = ((int64_t)M + RB - 1) / RB;
        if (beta != 0.0) {
            STK_TRY(check(cudaFuncSetAttribute(k_time_apply_smem<true>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)budget),
                          "stk_time_apply: smem attribute"));
            int64_t cap = (int64_t)sm_count() * ctas_per_sm;
            k_time_apply_smem<true><<<(unsigned)(ntiles < cap ? ntiles : cap), 256, smem, s>>>(
                M, nrows_t, nnz, indptr, indices, vals, x, ldx, ncols_local, xh, n_halo, alpha,
                beta, y, ldy, RB, W);
        } else {
            STK_TRY(check(cudaFuncSetAttribute(k_time_apply_smem<false>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)budget),
                          "stk_time_apply: smem attribute"));
            int64_t cap = (int64_t)sm_count() * ctas_per_sm;
            k_time_apply_smem<false><<<(unsigned)(ntiles < cap ? ntiles : cap), 256, smem, s>>>(
                M, nrows_t, nnz, indptr, indices, vals, x, ldx, ncols_local, xh, n_halo, alpha,
                beta, y, ldy, RB, W);
        }
        return check_launch("k_time_apply_smem");
    }
    if (beta != 0.0)
        k_time_apply<true><<<resident_grid(k_time_apply<true>, 256, work), 256, 0, s>>>(
            M, nrows_t, indptr, indices, vals, x, ldx, ncols_local, xh, alpha, beta, y, ldy);
    else
        k_time_apply<false><<<resident_grid(k_time_apply<false>, 256, work), 256, 0, s>>>(
            M, nrows_t, indptr, indices, vals, x, ldx, ncols_local, xh, alpha, beta, y, ldy);
    return check_launch("k_time_apply");
}

int stk_time_apply2(int M, int nrows_t, const int *indptr, const int *indices,
                    const double *vals, const double *x0, const double *x1, int ldx,
                    int ncols_local, const double *xh0, int n_halo0, const double *xh1,
                    double alpha, double beta, double *y, int ldy, int wy, void *stream) {
    if (x0 == y || x1 == y) return fail(-1, "stk_time_apply2: inputs must not alias y");
    if (nrows_t > wy || wy > ldy || (ldy & 1) || (wy & 1))
        return fail(-1, "stk_time_apply2: need nrows_t <= wy <= ldy, both even");
    if (M == 0) return 0;
    int64_t work = (int64_t)M * (wy / 2);
    if (work >= STK_MAX_ITEMS) return fail(-2, "stk_time_apply2: block too large");
    cudaStream_t s = as_stream(stream);
    if (beta != 0.0)
        k_time_apply2<true><<<resident_grid(k_time_apply2<true>, 256, work), 256, 0, s>>>(
            M, nrows_t, indptr, indices, vals, x0, x1, ldx, ncols_local, xh0, n_halo0, xh1, alpha,
            beta, y, ldy, wy);
    else
        k_time_apply2<false><<<resident_grid(k_time_apply2<false>, 256, work), 256, 0, s>>>(
            M, nrows_t, indptr, indices, vals, x0, x1, ldx, ncols_local, xh0, n_halo0, xh1, alpha,
            beta, y, ldy, wy);
    return check_launch("k_time_apply2");
}

int stk_time_tridiag_pair(int M, int n, int ld, const double *coef, const double *x0,
                          const double *x1, int ldx, const double *prev0, const double *next0,
                          const double *prev1, const double *next1, double *y1, double *y2,
                          int ldy, void *stream) {
    if ((ld & 1) || (ldx & 1) || (ldy & 1) || n > ld || ld > ldx || ld > 2048)
        return fail(-1, "stk_time_tridiag_pair: need even pitches, n <= ld <= ldx, ld <= 2048");
    if (x0 == y1 || x0 == y2 || x1 == y1 || x1 == y2 || y1 == y2)
        return fail(-1, "stk_time_tridiag_pair: aliasing");
    if (M == 0 || ld == 0) return 0;
    const int ld2 = ld / 2;
    const int threads = (ld2 + 31) / 32 * 32;
    int per_sm = 65536 / (threads * 88);  // ~82 registers per thread
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)sm_count() * per_sm;
    if (grid > M) grid = M;
    cudaStream_t s = as_stream(stream);
#define STK_TTP(NT)                                                                         \
    k_time_tridiag_pair<NT><<<(unsigned)grid, threads, 0, s>>>(M, n, ld2, coef, ld, x0, x1, ldx, \
                                                              prev0, next0, prev1, next1, y1, y2, ldy)
    if (threads <= 256) STK_TTP(256);
    else if (threads <= 512) STK_TTP(512);
    else STK_TTP(1024);  // 64 registers per thread: a few coefficients spill
#undef STK_TTP
    return check_launch("k_time_tridiag_pair");
}

int stk_space_spmm_split(int nrows, const int *indptr, const int *indices, const double *vals0,
                         const double *vals1, const double *x, double *y0, double *y1, int ld,
                         void *stream) {
    if (ld & 3) return fail(-1, "stk_space_spmm_split: pitch must be a multiple of 4");
    if (x == y0 || x == y1 || y0 == y1) return fail(-1, "stk_space_spmm_split: aliasing");
    if (nrows == 0) return 0;
    unsigned ld2 = (unsigned)ld / 2u;
    int64_t work = (int64_t)nrows * ld2;
    if (work >= STK_MAX_ITEMS) return fail(-2, "stk_space_spmm_split: block too large");
    k_space_spmm_split<<<resident_grid(k_space_spmm_split, 256, work), 256, 0,
                         as_stream(stream)>>>(nrows, indptr, indices, vals0, vals1, x, y0, y1, ld,
                                              ld2, row_order_for(indptr, nrows));
    return check_launch("k_space_spmm_split");
}

int stk_space_spmm_pair(int nrows, const int *indptr, const int *indices, const double *vals0,
                        const double *vals1, const double *x0, const double *x1, int ldx,
                        double alpha, double beta, const double *z, double *y, int ld,
                        void *stream) {
    if ((ld & 3) || (ldx & 1) || ldx < ld)
        return fail(-1, "stk_space_spmm_pair: need ld % 4 == 0 and an even ldx >= ld");
    if (x0 == y || x1 == y) return fail(-1, "stk_space_spmm_pair: inputs must not alias y");
    if (beta != 0.0 && !z) return fail(-1, "stk_space_spmm_pair: beta != 0 needs z");
    if (nrows == 0) return 0;
    unsigned ld2 = (unsigned)ld / 2u;
    int64_t work = (int64_t)nrows * ld2;
    if (work >= STK_MAX_ITEMS) return fail(-2, "stk_space_spmm_pair: block too large");
    cudaStream_t s = as_stream(stream);
    const int *rows = row_order_for(indptr, nrows);
    if (beta != 0.0)
        k_space_spmm_pair<true><<<resident_grid(k_space_spmm_pair<true>, 256, work), 256, 0, s>>>(
            nrows, indptr, indices, vals0, vals1, x0, x1, ldx, alpha, beta, z, y, ld, ld2, rows);
    else
        k_space_spmm_pair<false><<<resident_grid(k_space_spmm_pair<false>, 256, work), 256, 0,
                                   s>>>(nrows, indptr, indices, vals0, vals1, x0, x1, ldx, alpha,
                                        beta, z, y, ld, ld2, rows);
    return check_launch("k_space_spmm_pair");
}

int stk_copy_cols(int M, int n, const double *in, int ld_in, int c_in, const int *idx,
                  double *out, int ld_out, int c_out, void *stream) {
    if (M == 0 || n == 0) return 0;
    if (in == out) return fail(-1, "stk_copy_cols: in must not alias out");
    if (n < 0 || c_out < 0 || c_out + n > ld_out || (!idx && (c_in < 0 || c_in + n > ld_in)))
        return fail(-1, "stk_copy_cols: column range outside the pitch");
    k_copy_cols<<<resident_grid(k_copy_cols, 256, (int64_t)M * n), 256, 0, as_stream(stream)>>>(
        M, n, in, ld_in, c_in, idx, out, ld_out, c_out);
    return check_launch("k_copy_cols");
}

int stk_outer(int M, int ld, const double *ux, const double *ut, double *x, void *stream) {
    if (M == 0 || ld == 0) return 0;
    k_outer<<<resident_grid(k_outer, 256, (int64_t)M * ld), 256, 0, as_stream(stream)>>>(M, ld, ux,
                                                                                       ut, x);
    return check_launch("k_outer");
}

int stk_pack_slices(const double *x, int ld, int M, const int *tidx, int n, double *out,
                    void *stream) {
    if (n == 0 || M == 0) return 0;
    k_pack_slices<<<blocks_for((int64_t)n * M, 256), 256, 0, as_stream(stream)>>>(x, ld, M, tidx,
                                                                                 n, out);
    return check_launch("k_pack_slices");
}

int stk_unpack_slices(double *x, int ld, int M, const int *tidx, int n, const double *in,
                      double alpha, double beta, void *stream) {
    if (n == 0 || M == 0) return 0;
    k_unpack_slices<<<blocks_for((int64_t)n * M, 256), 256, 0, as_stream(stream)>>>(
        x, ld, M, tidx, n, in, alpha, beta);
    return check_launch("k_unpack_slices");
}

}  // extern "C"
