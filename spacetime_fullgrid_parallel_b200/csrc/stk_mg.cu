// Space multigrid V-cycle batched over all local time slices,
// /root/reference/source/multigrid.py:130-197 with the PETSc MatSOR sweeps of
// :100-127.  Every step is a sparse kernel over (row, time-pair) threads:
//   * Gauss-Seidel: the lexicographic sweep is run wavefront by wavefront (the
//     schedule is built on the host from the CSR dependency DAG); rows inside
//     a wavefront are independent, so each wavefront is one launch and the
//     iterates equal the sequential sweep's.
//   * residual, restriction, prolongation+correction: stk_space_spmm kernels.
//   * coarsest level: dense inverse per coefficient group.
// Every time slice t belongs to a GROUP g(t); the groups' level matrices share
// one sparsity pattern and are stored side by side (vals[g][nnz]).  That
// serves K_x = MG(A_x) (one group) and all C_j = MG(2^j M_x + alpha A_x) of
// heateq_mpi.py:143-153 (one group per j) with ONE launch sequence, and every
// slice sees exactly the level matrices the reference forms for it (Galerkin
// products of the combined matrix, multigrid.py:140-145).
#include <stdlib.h>

#include <unordered_map>
#include <vector>

#include "stk_common.cuh"

namespace stk {

int launch_space_spmm(int nrows, const int *indptr, const int *indices, int K,
                      const double *vals0, const double *vals1, const double *coef0,
                      const double *coef1, const double *x, double alpha, double beta,
                      const double *z, double *y, int ld, cudaStream_t s);
int launch_space_spmm_grouped(int nrows, const int *indptr, const int *indices,
                              const double *vals, size_t vstride, const int *grp, const double *x,
                              double alpha, double beta, const double *z, double *y, int ld,
                              cudaStream_t s);
int launch_space_spmm_kinds(int nrows, const int *indptr, const int *cidx,
                            const int *kind_of_row, const double *ktab, int G, int nkinds,
                            int kstride, const int *grp, const double *x, double alpha,
                            double beta, const double *z, double *y, int ld, cudaStream_t s);
}  // namespace stk
struct stk_gs_prog;
namespace stk {
int gs_fused_run(const stk_gs_prog *pg, int G, int T, const double *ktab, int nkinds,
                 int bulk_kind, const double *cvals, size_t vstride, const int *grp,
                 const double *f, const double *uin, double *uout, int ld, cudaStream_t s);

struct Level {
    int n = 0, nnz = 0;
    const int *indptr = nullptr, *indices = nullptr;
    const double *vals = nullptr;  // [G][nnz], CSR order
    const double *diag = nullptr;  // [G][n]
    const int *sched = nullptr;
    std::vector<int> phase_ptr;
    // transfer between this level and the next coarser one
    const int *p_indptr = nullptr, *p_indices = nullptr;
    const double *p_vals = nullptr;
    const int *r_indptr = nullptr, *r_indices = nullptr;
    const double *r_vals = nullptr;
    // fused smoother (stk_gsfused.cu): nu forward / nu backward sweeps as one
    // launch each, out of place; null = per-wavefront launches
    const stk_gs_prog *fused_fwd = nullptr, *fused_bwd = nullptr;
    // a second pair of programs, tiled for wide blocks (the two brackets of the
    // Schur operator side by side): used from `wide_min_chunks` time chunks on
    const stk_gs_prog *wide_fwd = nullptr, *wide_bwd = nullptr;
    int wide_min_chunks = 0;
    const double *ktab = nullptr;   // [G][nkinds][maxnnz + 2] (programs with kinds)
    const double *cvals = nullptr;  // [G][nnz], program entry order (generic programs)
    int nkinds = 0, bulk_kind = -1, fused_T = 8, kstride = 0;
    // row kinds of the level (programs with kinds): the grouped residual reads
    // the groups' values from ktab instead of vals
    const int *kind_of_row = nullptr, *canon_indices = nullptr;
};

// u_i += (f_i - A_i . u) / a_ii for the rows of one wavefront
// (multigrid.py:89-97: whole row including the diagonal, then the update).
// Grid-stride over (row, double2 column) items with a resident grid.
// GROUPED: the thread's two time values take their matrices from the groups
// grp[c], grp[c + 1].
// FIRST: first forward sweep from a zero initial guess = forward substitution
// with the lower triangle, u_i = (f_i - sum_{j<i} a_ij u_j) / a_ii: entries
// j >= i multiply zeros and are skipped, u is write-only (no memset needed).
// 8 CTAs per SM (32 registers, a few spilled bytes): the kernel is bound by
// memory latency x occupancy, full occupancy measured +10 % over 6 CTAs.
template <bool GROUPED, bool FIRST>
__global__ void __launch_bounds__(256, 8)
    k_gs_phase(const int *__restrict__ rows, int nrows, const int *__restrict__ indptr,
               const int *__restrict__ indices, const double *__restrict__ vals, size_t vstride,
               const double *__restrict__ diag, size_t dstride, const int *__restrict__ grp,
               const double *__restrict__ f, double *u, int ld, unsigned ld2) {
    const unsigned total = (unsigned)nrows * ld2;
    const unsigned stride = gridDim.x * 256u;
    for (unsigned k = blockIdx.x * 256u + threadIdx.x; k < total; k += stride) {
        unsigned r = k / ld2;
        unsigned c = (k - r * ld2) * 2u;
        int i = __ldg(rows + r);
        int p0 = __ldg(indptr + i), p1 = __ldg(indptr + i + 1);
        const double *v0 = vals, *v1 = vals, *e0 = diag, *e1 = diag;
        if (GROUPED) {
            const size_t g0 = (size_t)__ldg(grp + c), g1 = (size_t)__ldg(grp + c + 1);
            v0 += g0 * vstride;
            v1 += g1 * vstride;
            e0 += g0 * dstride;
            e1 += g1 * dstride;
        }
        double2 s = make_double2(0.0, 0.0);
        for (int p = p0; p < p1; ++p) {  // column indices are sorted
            int j = __ldg(indices + p);
            if (FIRST && j >= i) break;
            double2 xv = ldv2(u + (size_t)j * ld + c);
            double a0 = __ldg(v0 + p);
            s.x = fma(a0, xv.x, s.x);
            s.y = fma(GROUPED ? __ldg(v1 + p) : a0, xv.y, s.y);
        }
        size_t o = (size_t)i * ld + c;
        double2 fv = ldv2(f + o), uo = make_double2(0.0, 0.0);
        if (!FIRST) uo = ldv2(u + o);
        uo.x += (fv.x - s.x) / __ldg(e0 + i);
        uo.y += (fv.y - s.y) / __ldg(e1 + i);
        stv2(u + o, uo);
    }
}

// u[i,t] = sum_j inv[g(t)][i,j] f[j,t]   (multigrid.py:161-170, exact solve).
__global__ void __launch_bounds__(256)
    k_coarse_solve(int n0, const double *__restrict__ inv, const int *__restrict__ group,
                   const double *__restrict__ f, double *__restrict__ u, int ld) {
    unsigned k = blockIdx.x * 256u + threadIdx.x;
    unsigned i = k / (unsigned)ld;
    if (i >= (unsigned)n0) return;
    unsigned t = k - i * (unsigned)ld;
    const double *A = inv + (size_t)(group ? __ldg(group + t) : 0) * n0 * n0 + (size_t)i * n0;
    double s = 0.0;
    for (int j = 0; j < n0; ++j) s = fma(__ldg(A + j), f[(size_t)j * ld + t], s);
    u[(size_t)i * ld + t] = s;
}

// ---- the coarse levels of the V-cycle in ONE launch -----------------------
// Below some level every kernel of the cycle is a few microseconds of work and
// the cycle is a chain of ~30 dependent launches per level (one per Gauss-Seidel
// wavefront): launch latency, the same on every slab width -- what limits
// narrow time slabs (8 GPUs) first.  Time slices are independent, so one CTA
// takes a chunk of MGC_T slices through the WHOLE cycle below level `top`
// (multigrid.py:168-182 for j <= top, zero initial guess unless `u_in`):
// u and f of every level live in shared memory, a wavefront / residual /
// transfer is a block-synchronised phase instead of a launch, and the level
// matrices come through L1.  The arithmetic of every row is that of the
// per-level kernels (k_gs_phase, k_space_spmm, k_coarse_solve).
constexpr int MGC_T = 4;         // time values per CTA
constexpr int MGC_NT = 512;      // threads: 256 row lanes x 2 column pairs
constexpr int MGC_MAXLV = 8;

struct CoarseLevel {
    int n, nnz, nph;
    const int *indptr, *indices, *sched, *phase_ptr;  // phase_ptr: device copy
    const double *vals, *diag;
    const int *p_indptr, *p_indices, *r_indptr, *r_indices;
    const double *p_vals, *r_vals;
    int off_u, off_f;  // shared-memory offsets in rows of MGC_T doubles
};
struct CoarseArgs {
    CoarseLevel L[MGC_MAXLV];
    int top, nu, off_res, ld;
    const double *inv;  // [G][n0][n0]
    const int *grp;
    const double *f, *u_in;
    double *u;
};

template <bool GROUPED>
__global__ void __launch_bounds__(MGC_NT, 1) k_mg_coarse(const CoarseArgs a) {
    extern __shared__ __align__(16) double sm[];
    const int half = threadIdx.x & 1;    // which pair of the chunk's 4 columns
    const int lane = threadIdx.x >> 1;   // row lane
    constexpr int NL = MGC_NT / 2;
    const int c = blockIdx.x * MGC_T + 2 * half;
    const bool valid = c < a.ld;
    size_t g0 = 0, g1 = 0;
    if (GROUPED && valid) {
        g0 = (size_t)__ldg(a.grp + c);
        g1 = (size_t)__ldg(a.grp + c + 1);
    }
    auto at = [&](int off, int row) -> double2 * {
        return reinterpret_cast<double2 *>(sm + ((size_t)(off + row) * MGC_T + 2 * half));
    };
    const double2 zero2 = make_double2(0.0, 0.0);
    const int top = a.top;
    {   // right-hand side (and the initial guess) of the top level
        const CoarseLevel &lv = a.L[top];
        for (int i = lane; i < lv.n; i += NL) {
            *at(lv.off_f, i) = valid ? ldv2(a.f + (size_t)i * a.ld + c) : zero2;
            *at(lv.off_u, i) =
                (valid && a.u_in) ? ldv2(a.u_in + (size_t)i * a.ld + c) : zero2;
        }
    }
    __syncthreads();
    // one Gauss-Seidel sweep of level lv, wavefront by wavefront
    auto sweep = [&](const CoarseLevel &lv, bool backward) {
        const double *v0 = lv.vals + g0 * (size_t)lv.nnz, *v1 = lv.vals + g1 * (size_t)lv.nnz;
        const double *d0 = lv.diag + g0 * (size_t)lv.n, *d1 = lv.diag + g1 * (size_t)lv.n;
        for (int q = 0; q < lv.nph; ++q) {
            const int ph = backward ? lv.nph - 1 - q : q;
            const int r0 = __ldg(lv.phase_ptr + ph), r1 = __ldg(lv.phase_ptr + ph + 1);
            for (int r = r0 + lane; r < r1; r += NL) {
                const int i = __ldg(lv.sched + r);
                const int p1 = __ldg(lv.indptr + i + 1);
                double2 s = zero2;
                for (int p = __ldg(lv.indptr + i); p < p1; ++p) {
                    const double2 xv = *at(lv.off_u, __ldg(lv.indices + p));
                    const double a0 = __ldg(v0 + p);
                    s.x = fma(a0, xv.x, s.x);
                    s.y = fma(GROUPED ? __ldg(v1 + p) : a0, xv.y, s.y);
                }
                const double2 fv = *at(lv.off_f, i);
                double2 uo = *at(lv.off_u, i);
                uo.x += (fv.x - s.x) / __ldg(d0 + i);
                uo.y += (fv.y - s.y) / __ldg(d1 + i);
                *at(lv.off_u, i) = uo;
            }
            __syncthreads();
        }
    };
    // ---- down: smooth, residual, restrict ----
    for (int l = top; l >= 1; --l) {
        const CoarseLevel &lv = a.L[l];
        const CoarseLevel &lc = a.L[l - 1];
        for (int sw = 0; sw < a.nu; ++sw) sweep(lv, false);
        const double *v0 = lv.vals + g0 * (size_t)lv.nnz, *v1 = lv.vals + g1 * (size_t)lv.nnz;
        for (int i = lane; i < lv.n; i += NL) {  // res = A u - f
            const int p1 = __ldg(lv.indptr + i + 1);
            double2 s = zero2;
            for (int p = __ldg(lv.indptr + i); p < p1; ++p) {
                const double2 xv = *at(lv.off_u, __ldg(lv.indices + p));
                const double a0 = __ldg(v0 + p);
                s.x = fma(a0, xv.x, s.x);
                s.y = fma(GROUPED ? __ldg(v1 + p) : a0, xv.y, s.y);
            }
            const double2 fv = *at(lv.off_f, i);
            *at(a.off_res, i) = make_double2(s.x - fv.x, s.y - fv.y);
        }
        __syncthreads();
        for (int i = lane; i < lc.n; i += NL) {  // f_c = R res, u_c = 0
            const int p1 = __ldg(lv.r_indptr + i + 1);
            double2 s = zero2;
            for (int p = __ldg(lv.r_indptr + i); p < p1; ++p) {
                const double2 xv = *at(a.off_res, __ldg(lv.r_indices + p));
                const double w = __ldg(lv.r_vals + p);
                s.x = fma(w, xv.x, s.x);
                s.y = fma(w, xv.y, s.y);
            }
            *at(lc.off_f, i) = s;
            *at(lc.off_u, i) = zero2;
        }
        __syncthreads();
    }
    {   // level 0: exact solve with the dense inverse
        const CoarseLevel &l0 = a.L[0];
        const int n0 = l0.n;
        const double *i0 = a.inv + g0 * (size_t)n0 * n0, *i1 = a.inv + g1 * (size_t)n0 * n0;
        for (int i = lane; i < n0; i += NL) {
            double2 s = zero2;
            for (int j = 0; j < n0; ++j) {
                const double2 fv = *at(l0.off_f, j);
                s.x = fma(__ldg(i0 + (size_t)i * n0 + j), fv.x, s.x);
                s.y = fma(__ldg(i1 + (size_t)i * n0 + j), fv.y, s.y);
            }
            *at(l0.off_u, i) = s;
        }
        __syncthreads();
    }
    // ---- up: correct, smooth ----
    for (int l = 1; l <= top; ++l) {
        const CoarseLevel &lv = a.L[l];
        const CoarseLevel &lc = a.L[l - 1];
        for (int i = lane; i < lv.n; i += NL) {  // u -= P u_c
            const int p1 = __ldg(lv.p_indptr + i + 1);
            double2 s = zero2;
            for (int p = __ldg(lv.p_indptr + i); p < p1; ++p) {
                const double2 xv = *at(lc.off_u, __ldg(lv.p_indices + p));
                const double w = __ldg(lv.p_vals + p);
                s.x = fma(w, xv.x, s.x);
                s.y = fma(w, xv.y, s.y);
            }
            double2 uo = *at(lv.off_u, i);
            uo.x = fma(-1.0, s.x, uo.x);
            uo.y = fma(-1.0, s.y, uo.y);
            *at(lv.off_u, i) = uo;
        }
        __syncthreads();
        for (int sw = 0; sw < a.nu; ++sw) sweep(lv, true);
    }
    if (valid) {
        const CoarseLevel &lv = a.L[top];
        for (int i = lane; i < lv.n; i += NL) stv2(a.u + (size_t)i * a.ld + c, *at(lv.off_u, i));
    }
}

}  // namespace stk

using namespace stk;

// One captured V-cycle sequence: the launch arguments are baked in, so a
// graph is keyed by every pointer it was captured with.
struct GraphKey {
    const void *p[7];
    int ld;
    bool operator==(const GraphKey &o) const {
        for (int k = 0; k < 7; ++k)
            if (p[k] != o.p[k]) return false;
        return ld == o.ld;
    }
};
struct GraphKeyHash {
    size_t operator()(const GraphKey &k) const {
        size_t h = (size_t)k.ld * 1000003u;
        for (int i = 0; i < 7; ++i) h = h * 1099511628211ull ^ (size_t)k.p[i];
        return h;
    }
};
struct GraphEntry {
    cudaGraphExec_t exec = nullptr;  // null: key seen once, not captured yet
    int64_t launches = 0;
};

struct stk_mg {
    int nlevels, nu, vcycles, G;
    std::vector<Level> L;
    std::unordered_map<GraphKey, GraphEntry, GraphKeyHash> graphs;
    cudaStream_t capture_stream = nullptr;
    // coarse levels in one launch (k_mg_coarse): -1 = not decided yet, 0 = off
    int coarse_top = -1;
    int *coarse_phase_ptr = nullptr;  // device copies of the levels' phase_ptr
    std::vector<int> coarse_phase_off;
    size_t coarse_smem = 0;
    ~stk_mg() {
        for (auto &kv : graphs)
            if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        if (capture_stream) cudaStreamDestroy(capture_stream);
        if (coarse_phase_ptr) cudaFree(coarse_phase_ptr);
    }
};

// Decide once which levels the one-launch coarse cycle takes: the largest
// `top` whose u, f of all levels <= top plus one residual block fit the shared
// memory of a CTA.  STK_MG_COARSE=0 turns it off.
static int coarse_setup(stk_mg *mg) {
    if (mg->coarse_top >= 0) return 0;
    mg->coarse_top = 0;
    const char *e = getenv("STK_MG_COARSE");
    if (e && e[0] == '0') return 0;
    const size_t budget = 200 * 1024;
    int top = 0;
    for (int l = 1; l < mg->nlevels && l < MGC_MAXLV; ++l) {
        const Level &lv = mg->L[l];
        if (!lv.sched || !lv.p_indptr || !lv.r_indptr) break;
        size_t rows = (size_t)lv.n;
        for (int k = 0; k <= l; ++k) rows += 2 * (size_t)mg->L[k].n;
        if (rows * MGC_T * sizeof(double) > budget) break;
        top = l;
    }
    if (top < 1) return 0;
    std::vector<int> host;
    mg->coarse_phase_off.assign(top + 1, 0);
    for (int l = 1; l <= top; ++l) {
        mg->coarse_phase_off[l] = (int)host.size();
        host.insert(host.end(), mg->L[l].phase_ptr.begin(), mg->L[l].phase_ptr.end());
    }
    STK_TRY(check(cudaMalloc(&mg->coarse_phase_ptr, host.size() * sizeof(int)),
                  "stk_mg: coarse schedule"));
    STK_TRY(check(cudaMemcpy(mg->coarse_phase_ptr, host.data(), host.size() * sizeof(int),
                             cudaMemcpyHostToDevice),
                  "stk_mg: coarse schedule upload"));
    size_t rows = (size_t)mg->L[top].n;
    for (int k = 0; k <= top; ++k) rows += 2 * (size_t)mg->L[k].n;
    mg->coarse_smem = rows * MGC_T * sizeof(double);
    STK_TRY(check(cudaFuncSetAttribute(k_mg_coarse<true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)budget),
                  "stk_mg: coarse kernel attribute"));
    STK_TRY(check(cudaFuncSetAttribute(k_mg_coarse<false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)budget),
                  "stk_mg: coarse kernel attribute"));
    mg->coarse_top = top;
    return 0;
}

// The V-cycle below level l (l <= coarse_top) in one launch.
static int coarse_cycle(const stk_mg *mg, int l, const int *grp, const double *inv,
                        const double *f, double *u, int ld, cudaStream_t s, bool zero_guess) {
    CoarseArgs a;
    int off = 0;
    for (int k = 0; k <= l; ++k) {
        const Level &lv = mg->L[k];
        CoarseLevel &c = a.L[k];
        c.n = lv.n;
        c.nnz = lv.nnz;
        c.nph = k ? (int)lv.phase_ptr.size() - 1 : 0;
        c.indptr = lv.indptr;
        c.indices = lv.indices;
        c.sched = lv.sched;
        c.phase_ptr = k ? mg->coarse_phase_ptr + mg->coarse_phase_off[k] : nullptr;
        c.vals = lv.vals;
        c.diag = lv.diag;
        c.p_indptr = lv.p_indptr;
        c.p_indices = lv.p_indices;
        c.p_vals = lv.p_vals;
        c.r_indptr = lv.r_indptr;
        c.r_indices = lv.r_indices;
        c.r_vals = lv.r_vals;
        c.off_u = off;
        c.off_f = off + lv.n;
        off += 2 * lv.n;
    }
    a.top = l;
    a.nu = mg->nu;
    a.off_res = off;
    a.ld = ld;
    a.inv = inv;
    a.grp = grp;
    a.f = f;
    a.u_in = zero_guess ? nullptr : u;
    a.u = u;
    const size_t smem = (size_t)(off + mg->L[l].n) * MGC_T * sizeof(double);
    const unsigned grid = (unsigned)((ld + MGC_T - 1) / MGC_T);
    if (grp)
        k_mg_coarse<true><<<grid, MGC_NT, smem, s>>>(a);
    else
        k_mg_coarse<false><<<grid, MGC_NT, smem, s>>>(a);
    return check_launch("k_mg_coarse");
}

static int smooth(const stk_mg *mg, int l, int nsweeps, bool backward, const int *grp,
                  const double *f, double *u, int ld, cudaStream_t s, bool zero_guess = false) {
    const Level &lv = mg->L[l];
    const int nph = (int)lv.phase_ptr.size() - 1;
    const unsigned ld2 = (unsigned)ld / 2u;
    for (int sw = 0; sw < nsweeps; ++sw) {
        for (int q = 0; q < nph; ++q) {
            int ph = backward ? nph - 1 - q : q;
            int r0 = lv.phase_ptr[ph], nr = lv.phase_ptr[ph + 1] - r0;
            if (nr == 0) continue;
            if ((int64_t)nr * ld2 >= STK_MAX_ITEMS) return fail(-2, "stk_mg: block too large");
            const bool first = zero_guess && sw == 0 && !backward;
#define STK_GS(GG, FF)                                                                        \
    k_gs_phase<GG, FF><<<resident_grid(k_gs_phase<GG, FF>, 256, (int64_t)nr * ld2), 256, 0, s>>>( \
        lv.sched + r0, nr, lv.indptr, lv.indices, lv.vals, (size_t)lv.nnz, lv.diag, (size_t)lv.n, \
        grp, f, u, ld, ld2)
            if (grp) {
                if (first) STK_GS(true, true); else STK_GS(true, false);
            } else {
                if (first) STK_GS(false, true); else STK_GS(false, false);
            }
#undef STK_GS
            STK_TRY(check_launch("k_gs_phase"));
        }
    }
    return 0;
}

struct Workspace {
    std::vector<double *> u, f;  // per level below the finest
    std::vector<double *> alt;   // per fused level: output of the pre-smoother
    double *res;                 // residual of the current level
};

static int smooth_fused(const stk_mg *mg, int l, bool backward, const int *grp, const double *f,
                        const double *uin, double *uout, int ld, cudaStream_t s) {
    const Level &lv = mg->L[l];
    const bool wide = lv.wide_fwd && lv.wide_bwd &&
                      (ld + lv.fused_T - 1) / lv.fused_T >= lv.wide_min_chunks;
    const stk_gs_prog *pg = backward ? (wide ? lv.wide_bwd : lv.fused_bwd)
                                     : (wide ? lv.wide_fwd : lv.fused_fwd);
    return gs_fused_run(pg, mg->G, lv.fused_T, lv.ktab,
                        lv.nkinds, lv.bulk_kind, lv.cvals, (size_t)lv.nnz, grp, f, uin, uout, ld,
                        s);
}

static int coarse_solve(const stk_mg *mg, const double *inv, const int *group, const double *f,
                        double *u, int ld, cudaStream_t s) {
    int n0 = mg->L[0].n;
    k_coarse_solve<<<blocks_for((int64_t)n0 * ld, 256), 256, 0, s>>>(n0, inv, group, f, u, ld);
    return check_launch("k_coarse_solve");
}

// MGM(j, u_j, f_j) of multigrid.py:168-182.  `zero_guess`: u holds no data yet
// and stands for u = 0 (the reference passes np.zeros, multigrid.py:176,187).
static int cycle(const stk_mg *mg, int l, const int *grp, const double *inv, const double *f,
                 double *u, int ld, Workspace &ws, cudaStream_t s, bool zero_guess) {
    if (l >= 1 && l <= mg->coarse_top) return coarse_cycle(mg, l, grp, inv, f, u, ld, s, zero_guess);
    if (l == 0) return coarse_solve(mg, inv, grp, f, u, ld, s);
    const Level &lv = mg->L[l];
    const Level &lc = mg->L[l - 1];
    const bool fused = lv.fused_fwd && lv.fused_bwd && mg->nu > 0 && ws.alt[l];
    // `cur` holds the iterate between the two smoothers: the fused smoother
    // works out of place (items re-read rows of their neighbours' halo), so
    // the pre-smoother writes to the level's alternate block and the
    // post-smoother brings the result back to u.
    double *cur = u;
    if (fused) {
        cur = ws.alt[l];
        STK_TRY(smooth_fused(mg, l, false, grp, f, zero_guess ? nullptr : u, cur, ld, s));
    } else {
        if (zero_guess && mg->nu == 0)
            STK_TRY(check(cudaMemsetAsync(u, 0, sizeof(double) * (size_t)lv.n * ld, s),
                          "stk_mg: memset"));
        STK_TRY(smooth(mg, l, mg->nu, false, grp, f, u, ld, s, zero_guess));
    }
    // res = A u - f, f_c = R res (multigrid.py:174).  A fused kernel that
    // recomputes the fine residuals per coarse row was measured 2.6x slower
    // (6 block passes of DRAM reads instead of ~3): the residual block costs
    // less than the lost locality.
    int rc = 1;
    if (grp && lv.ktab && lv.kind_of_row && lv.canon_indices)
        rc = launch_space_spmm_kinds(lv.n, lv.indptr, lv.canon_indices, lv.kind_of_row, lv.ktab,
                                     mg->G, lv.nkinds, lv.kstride, grp, cur, 1.0, -1.0, f, ws.res,
                                     ld, s);
    if (rc == 1)
        rc = launch_space_spmm_grouped(lv.n, lv.indptr, lv.indices, lv.vals, (size_t)lv.nnz, grp,
                                       cur, 1.0, -1.0, f, ws.res, ld, s);
    STK_TRY(rc);
    STK_TRY(launch_space_spmm(lc.n, lv.r_indptr, lv.r_indices, 1, lv.r_vals, nullptr, nullptr,
                              nullptr, ws.res, 1.0, 0.0, nullptr, ws.f[l - 1], ld, s));
    STK_TRY(cycle(mg, l - 1, grp, inv, ws.f[l - 1], ws.u[l - 1], ld, ws, s, true));
    // u -= P u_c
    STK_TRY(launch_space_spmm(lv.n, lv.p_indptr, lv.p_indices, 1, lv.p_vals, nullptr, nullptr,
                              nullptr, ws.u[l - 1], -1.0, 1.0, cur, cur, ld, s));
    if (fused) return smooth_fused(mg, l, true, grp, f, cur, u, ld, s);
    return smooth(mg, l, mg->nu, true, grp, f, u, ld, s);
}

extern "C" {

stk_mg *stk_mg_create(int nlevels, int smoothsteps, int vcycles, int G) {
    if (nlevels < 1 || G < 1 || smoothsteps < 0 || vcycles < 0) {
        fail(-1, "stk_mg_create: bad arguments");
        return nullptr;
    }
    stk_mg *mg = new stk_mg;
    mg->nlevels = nlevels;
    mg->nu = smoothsteps;
    mg->vcycles = vcycles;
    mg->G = G;
    mg->L.resize(nlevels);
    return mg;
}

void stk_mg_destroy(stk_mg *mg) { delete mg; }

int stk_mg_set_level(stk_mg *mg, int level, int nrows, int nnz, const int *indptr,
                     const int *indices, const double *vals, const double *diag,
                     const int *sched_rows, const int *phase_ptr_host, int nphases) {
    if (!mg || level < 0 || level >= mg->nlevels) return fail(-1, "stk_mg_set_level: bad level");
    if (!vals || !diag) return fail(-1, "stk_mg_set_level: values missing");
    if (mg->coarse_top >= 0) {  // levels change: decide the coarse cycle again
        if (mg->coarse_phase_ptr) cudaFree(mg->coarse_phase_ptr);
        mg->coarse_phase_ptr = nullptr;
        mg->coarse_top = -1;
        for (auto &kv : mg->graphs)
            if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        mg->graphs.clear();
    }
    Level &lv = mg->L[level];
    lv.n = nrows;
    lv.nnz = nnz;
    lv.indptr = indptr;
    lv.indices = indices;
    lv.vals = vals;
    lv.diag = diag;
    lv.sched = sched_rows;
    lv.phase_ptr.assign(phase_ptr_host, phase_ptr_host + nphases + 1);
    if (lv.phase_ptr.front() != 0 || lv.phase_ptr.back() != nrows)
        return fail(-1, "stk_mg_set_level: schedule does not cover the rows");
    return 0;
}

int stk_mg_set_transfer(stk_mg *mg, int level, const int *p_indptr, const int *p_indices,
                        const double *p_vals, const int *r_indptr, const int *r_indices,
                        const double *r_vals) {
    if (!mg || level < 1 || level >= mg->nlevels)
        return fail(-1, "stk_mg_set_transfer: bad level");
    Level &lv = mg->L[level];
    lv.p_indptr = p_indptr;
    lv.p_indices = p_indices;
    lv.p_vals = p_vals;
    lv.r_indptr = r_indptr;
    lv.r_indices = r_indices;
    lv.r_vals = r_vals;
    return 0;
}

int64_t stk_mg_workspace(const stk_mg *mg, int ld) {
    int64_t rows = 0;
    for (int l = 0; l + 1 < mg->nlevels; ++l) rows += 2 * (int64_t)mg->L[l].n;
    rows += mg->L[mg->nlevels - 1].n;
    for (int l = 1; l < mg->nlevels; ++l)
        if (mg->L[l].fused_fwd && mg->L[l].fused_bwd) rows += mg->L[l].n;
    return rows * ld;
}

int stk_mg_set_fused(stk_mg *mg, int level, const stk_gs_prog *fwd, const stk_gs_prog *bwd,
                     const double *ktab, int nkinds, int bulk_kind, const double *cvals, int T,
                     int kstride, const int *kind_of_row, const int *canon_indices) {
    if (!mg || level < 1 || level >= mg->nlevels) return fail(-1, "stk_mg_set_fused: bad level");
    if (T != 8) return fail(-1, "stk_mg_set_fused: T must be 8");
    if (!ktab && !cvals) return fail(-1, "stk_mg_set_fused: values missing");
    Level &lv = mg->L[level];
    lv.fused_fwd = fwd;
    lv.fused_bwd = bwd;
    lv.ktab = ktab;
    lv.nkinds = nkinds;
    lv.bulk_kind = bulk_kind;
    lv.cvals = cvals;
    lv.fused_T = T;
    lv.kstride = kstride;
    lv.kind_of_row = ktab ? kind_of_row : nullptr;
    lv.canon_indices = ktab ? canon_indices : nullptr;
    for (auto &kv : mg->graphs)  // captured sequences are stale now
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    mg->graphs.clear();
    return 0;
}

int stk_mg_set_fused_wide(stk_mg *mg, int level, const stk_gs_prog *fwd, const stk_gs_prog *bwd,
                          int min_chunks) {
    if (!mg || level < 1 || level >= mg->nlevels)
        return fail(-1, "stk_mg_set_fused_wide: bad level");
    Level &lv = mg->L[level];
    if (!lv.fused_fwd || !lv.fused_bwd)
        return fail(-1, "stk_mg_set_fused_wide: attach the level's programs first");
    lv.wide_fwd = fwd;
    lv.wide_bwd = bwd;
    lv.wide_min_chunks = min_chunks;
    for (auto &kv : mg->graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    mg->graphs.clear();
    return 0;
}

int stk_mg_apply(stk_mg *mg, const int *group, const double *coarse_inv, const double *b,
                 double *x, int ld, double *wsbuf, void *stream) {
    if (!mg) return fail(-1, "stk_mg_apply: null handle");
    if (ld & 3) return fail(-1, "stk_mg_apply: pitch must be a multiple of 4");
    if (mg->G > 1 && !group) return fail(-1, "stk_mg_apply: several groups need the group table");
    if (b == x) return fail(-1, "stk_mg_apply: b must not alias x");
    cudaStream_t s = as_stream(stream);
    const int top = mg->nlevels - 1;
    STK_TRY(coarse_setup(mg));
    Workspace ws;
    ws.u.resize(mg->nlevels);
    ws.f.resize(mg->nlevels);
    double *q = wsbuf;
    for (int l = 0; l < top; ++l) {
        ws.u[l] = q;
        q += (size_t)mg->L[l].n * ld;
        ws.f[l] = q;
        q += (size_t)mg->L[l].n * ld;
    }
    ws.res = q;
    q += (size_t)mg->L[top].n * ld;
    ws.alt.assign(mg->nlevels, nullptr);
    for (int l = 1; l <= top; ++l)
        if (mg->L[l].fused_fwd && mg->L[l].fused_bwd) {
            ws.alt[l] = q;
            q += (size_t)mg->L[l].n * ld;
        }
    auto run = [&](cudaStream_t st) -> int {
        if (mg->vcycles == 0)
            STK_TRY(check(cudaMemsetAsync(x, 0, sizeof(double) * (size_t)mg->L[top].n * ld, st),
                          "stk_mg_apply: memset"));
        for (int v = 0; v < mg->vcycles; ++v)
            STK_TRY(cycle(mg, top, group, coarse_inv, b, x, ld, ws, st, v == 0));
        return 0;
    };
    // The apply is hundreds of launches, most of them on coarse levels where a
    // launch costs more than the kernel.  The second time the same buffers
    // come back (torch's caching allocator recycles addresses in a Krylov
    // loop) the sequence is captured into a CUDA graph; from then on it is
    // one graph launch.  STK_MG_GRAPH=0 disables this.
    static const bool use_graphs = [] {
        const char *e = getenv("STK_MG_GRAPH");
        return !(e && e[0] == '0');
    }();
    if (!use_graphs) return run(s);
    GraphKey key{{group, coarse_inv, b, x, wsbuf, nullptr, nullptr}, ld};
    auto it = mg->graphs.find(key);
    if (it == mg->graphs.end()) {
        if (mg->graphs.size() >= 64) {  // buffers are not recurring: start over
            for (auto &kv : mg->graphs)
                if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
            mg->graphs.clear();
        }
        mg->graphs.emplace(key, GraphEntry());
        return run(s);
    }
    if (!it->second.exec) {
        if (!mg->capture_stream)
            STK_TRY(check(cudaStreamCreateWithFlags(&mg->capture_stream, cudaStreamNonBlocking),
                          "stk_mg_apply: capture stream"));
        cudaGraph_t graph = nullptr;
        int64_t before = g_launches;
        STK_TRY(check(cudaStreamBeginCapture(mg->capture_stream, cudaStreamCaptureModeThreadLocal),
                      "stk_mg_apply: begin capture"));
        int rc = run(mg->capture_stream);
        cudaError_t e = cudaStreamEndCapture(mg->capture_stream, &graph);
        int64_t captured = g_launches - before;
        g_launches = before;
        if (rc != 0 || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            mg->graphs.erase(it);
            return run(s);  // capture unavailable: plain launches
        }
        cudaGraphExec_t exec = nullptr;
        e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            cudaGetLastError();
            mg->graphs.erase(it);
            return run(s);
        }
        it->second.exec = exec;
        it->second.launches = captured;
    }
    STK_TRY(check(cudaGraphLaunch(it->second.exec, s), "stk_mg_apply: graph launch"));
    g_launches += it->second.launches;
    return 0;
}

// Host helper (HOST pointers): wavefront number of every row of the
// lexicographic Gauss-Seidel dependency DAG, wave[i] = 1 + max wave[j] over
// the neighbours j < i (0 if none).  Returns the number of wavefronts.
int stk_gs_wavefronts(int n, const int *indptr, const int *indices, int *wave) {
    int depth = 0;
    for (int i = 0; i < n; ++i) {
        int w = 0;
        for (int p = indptr[i]; p < indptr[i + 1]; ++p) {
            int j = indices[p];
            if (j < i && wave[j] + 1 > w) w = wave[j] + 1;
        }
        wave[i] = w;
        if (w + 1 > depth) depth = w + 1;
    }
    return depth;
}

int stk_mg_smooth(stk_mg *mg, int level, int nsweeps, int backward, const int *group,
                  const double *f, double *u, int ld, void *stream) {
    if (!mg || level < 1 || level >= mg->nlevels) return fail(-1, "stk_mg_smooth: bad level");
    if (mg->G > 1 && !group) return fail(-1, "stk_mg_smooth: several groups need the group table");
    return smooth(mg, level, nsweeps, backward != 0, group, f, u, ld, as_stream(stream));
}

}  // extern "C"
