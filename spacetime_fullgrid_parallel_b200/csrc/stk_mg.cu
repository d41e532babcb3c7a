// Space multigrid V-cycle batched over all local time slices,
// /root/reference/source/multigrid.py:130-197 with the PETSc MatSOR sweeps of
// :100-127.  Every step is a sparse kernel over (row, time-pair) threads:
//   * Gauss-Seidel: the lexicographic sweep is run wavefront by wavefront (the
//     schedule is built on the host from the CSR dependency DAG); rows inside
//     a wavefront are independent, so each wavefront is one launch and the
//     iterates equal the sequential sweep's.
//   * residual, restriction, prolongation+correction: stk_space_spmm kernels.
//   * coarsest level: dense inverse per coefficient group.
// The matrix of slice t is c0[t]*A0 + c1[t]*A1 on a shared pattern (K = 2),
// which serves K_x = MG(A_x) and every C_j = MG(2^j M_x + alpha A_x) of
// heateq_mpi.py:143-153 with ONE hierarchy and one launch sequence.
#include <stdlib.h>

#include <unordered_map>
#include <vector>

#include "stk_common.cuh"

namespace stk {

int launch_space_spmm(int nrows, const int *indptr, const int *indices, int K,
                      const double *vals0, const double *vals1, const double *coef0,
                      const double *coef1, const double *x, double alpha, double beta,
                      const double *z, double *y, int ld, cudaStream_t s);
}  // namespace stk
struct stk_gs_prog;
namespace stk {
int gs_fused_run(const stk_gs_prog *pg, int K, int T, const double *ktab, int nkinds,
                 const double *v0, const double *v1, const double *d0, const double *d1,
                 const double *coef0, const double *coef1, const double *f, const double *uin,
                 double *uout, int ld, cudaStream_t s);

struct Level {
    int n = 0;
    const int *indptr = nullptr, *indices = nullptr;
    const double *v0 = nullptr, *v1 = nullptr, *d0 = nullptr, *d1 = nullptr;
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
    const double *ktab = nullptr;
    int nkinds = 0, fused_T = 8;
};

// u_i += (f_i - A_i . u) / a_ii for the rows of one wavefront
// (multigrid.py:89-97: whole row including the diagonal, then the update).
// Grid-stride over (row, double2 column) items with a resident grid.
// FIRST: first forward sweep from a zero initial guess = forward substitution
// with the lower triangle, u_i = (f_i - sum_{j<i} a_ij u_j) / a_ii: entries
// j >= i multiply zeros and are skipped, u is write-only (no memset needed).
// 8 CTAs per SM (32 registers, a few spilled bytes): the kernel is bound by
// memory latency x occupancy, full occupancy measured +10 % over 6 CTAs.
template <int K, bool FIRST>
__global__ void __launch_bounds__(256, 8)
    k_gs_phase(const int *__restrict__ rows, int nrows, const int *__restrict__ indptr,
               const int *__restrict__ indices, const double *__restrict__ v0,
               const double *__restrict__ v1, const double *__restrict__ d0,
               const double *__restrict__ d1, const double *__restrict__ coef0,
               const double *__restrict__ coef1, const double *__restrict__ f, double *u, int ld,
               unsigned ld2) {
    const unsigned total = (unsigned)nrows * ld2;
    const unsigned stride = gridDim.x * 256u;
    for (unsigned k = blockIdx.x * 256u + threadIdx.x; k < total; k += stride) {
        unsigned r = k / ld2;
        unsigned c = (k - r * ld2) * 2u;
        int i = __ldg(rows + r);
        int p0 = __ldg(indptr + i), p1 = __ldg(indptr + i + 1);
        double2 s0 = make_double2(0.0, 0.0), s1 = make_double2(0.0, 0.0);
        if (FIRST) {  // column indices are sorted: the lower triangle comes first
            for (int p = p0; p < p1; ++p) {
                int jj = __ldg(indices + p);
                if (jj >= i) break;
                double2 xx = ldv2(u + (size_t)jj * ld + c);
                double b0 = __ldg(v0 + p);
                s0.x = fma(b0, xx.x, s0.x);
                s0.y = fma(b0, xx.y, s0.y);
                if (K == 2) {
                    double b1 = __ldg(v1 + p);
                    s1.x = fma(b1, xx.x, s1.x);
                    s1.y = fma(b1, xx.y, s1.y);
                }
            }
        } else {
            row_product<K>(p0, p1, indices, v0, v1, u, ld, c, s0, s1);
        }
        double2 diag;
        if (K == 2) {
            double2 c0 = ldg2(coef0 + c), c1 = ldg2(coef1 + c);
            s0.x = fma(c0.x, s0.x, c1.x * s1.x);
            s0.y = fma(c0.y, s0.y, c1.y * s1.y);
            double e0 = __ldg(d0 + i), e1 = __ldg(d1 + i);
            diag.x = fma(c0.x, e0, c1.x * e1);
            diag.y = fma(c0.y, e0, c1.y * e1);
        } else {
            diag.x = diag.y = __ldg(d0 + i);
        }
        size_t o = (size_t)i * ld + c;
        double2 fv = ldv2(f + o), uo = make_double2(0.0, 0.0);
        if (!FIRST) uo = ldv2(u + o);
        uo.x += (fv.x - s0.x) / diag.x;
        uo.y += (fv.y - s0.y) / diag.y;
        stv2(u + o, uo);
    }
}

// The same wavefront update with FOUR time values per thread (256-bit loads and
// stores): twice the bytes in flight per warp for a latency-bound kernel.
template <int K, bool FIRST>
__global__ void __launch_bounds__(256)
    k_gs_phase4(const int *__restrict__ rows, int nrows, const int *__restrict__ indptr,
                const int *__restrict__ indices, const double *__restrict__ v0,
                const double *__restrict__ v1, const double *__restrict__ d0,
                const double *__restrict__ d1, const double *__restrict__ coef0,
                const double *__restrict__ coef1, const double *__restrict__ f, double *u, int ld,
                unsigned ld4) {
    const unsigned total = (unsigned)nrows * ld4;
    const unsigned stride = gridDim.x * 256u;
    for (unsigned k = blockIdx.x * 256u + threadIdx.x; k < total; k += stride) {
        unsigned r = k / ld4;
        unsigned c = (k - r * ld4) * 4u;
        int i = __ldg(rows + r);
        int p0 = __ldg(indptr + i), p1 = __ldg(indptr + i + 1);
        double4v s0 = {0.0, 0.0, 0.0, 0.0}, s1 = {0.0, 0.0, 0.0, 0.0};
        for (int p = p0; p < p1; ++p) {
            int j = __ldg(indices + p);
            if (FIRST && j >= i) break;
            double4v xv = ldv4(u + (size_t)j * ld + c);
            fma4(__ldg(v0 + p), xv, s0);
            if (K == 2) fma4(__ldg(v1 + p), xv, s1);
        }
        double4v diag;
        if (K == 2) {
            double4v c0 = ldv4(coef0 + c), c1 = ldv4(coef1 + c);
            s0.x = fma(c0.x, s0.x, c1.x * s1.x);
            s0.y = fma(c0.y, s0.y, c1.y * s1.y);
            s0.z = fma(c0.z, s0.z, c1.z * s1.z);
            s0.w = fma(c0.w, s0.w, c1.w * s1.w);
            double e0 = __ldg(d0 + i), e1 = __ldg(d1 + i);
            diag.x = fma(c0.x, e0, c1.x * e1);
            diag.y = fma(c0.y, e0, c1.y * e1);
            diag.z = fma(c0.z, e0, c1.z * e1);
            diag.w = fma(c0.w, e0, c1.w * e1);
        } else {
            diag.x = diag.y = diag.z = diag.w = __ldg(d0 + i);
        }
        size_t o = (size_t)i * ld + c;
        double4v fv = ldv4(f + o), uo = {0.0, 0.0, 0.0, 0.0};
        if (!FIRST) uo = ldv4(u + o);
        uo.x += (fv.x - s0.x) / diag.x;
        uo.y += (fv.y - s0.y) / diag.y;
        uo.z += (fv.z - s0.z) / diag.z;
        uo.w += (fv.w - s0.w) / diag.w;
        stv4(u + o, uo);
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
    int nlevels, nu, vcycles, K;
    std::vector<Level> L;
    std::unordered_map<GraphKey, GraphEntry, GraphKeyHash> graphs;
    cudaStream_t capture_stream = nullptr;
    ~stk_mg() {
        for (auto &kv : graphs)
            if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        if (capture_stream) cudaStreamDestroy(capture_stream);
    }
};

static int smooth(const stk_mg *mg, int l, int nsweeps, bool backward, const double *c0,
                  const double *c1, const double *f, double *u, int ld, cudaStream_t s,
                  bool zero_guess = false) {
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
#define STK_GS(KK, FF)                                                                        \
    k_gs_phase<KK, FF><<<resident_grid(k_gs_phase<KK, FF>, 256, (int64_t)nr * ld2), 256, 0, s>>>( \
        lv.sched + r0, nr, lv.indptr, lv.indices, lv.v0, lv.v1, lv.d0, lv.d1, c0, c1, f, u, ld,  \
        ld2)
#define STK_GS4(KK, FF)                                                                       \
    k_gs_phase4<KK, FF><<<resident_grid(k_gs_phase4<KK, FF>, 256, (int64_t)nr * (ld2 / 2)), 256,  \
                          0, s>>>(lv.sched + r0, nr, lv.indptr, lv.indices, lv.v0, lv.v1, lv.d0, \
                                  lv.d1, c0, c1, f, u, ld, ld2 / 2)
            // STK_GS_VEC4: 0 = never, 1 = always, default = for K = 2 only (the
            // per-slice-coefficient kernel gains ~5 % from 256-bit accesses, the
            // single-matrix one loses ~1.5 %: measured, profiles/r1_experiments.md)
            static const int vec4 = [] {
                const char *e = getenv("STK_GS_VEC4");
                return e ? atoi(e) : 2;
            }();
            if (vec4 == 1 || (vec4 == 2 && mg->K == 2)) {
                if (mg->K == 2) {
                    if (first) STK_GS4(2, true); else STK_GS4(2, false);
                } else {
                    if (first) STK_GS4(1, true); else STK_GS4(1, false);
                }
            } else
            if (mg->K == 2) {
                if (first) STK_GS(2, true); else STK_GS(2, false);
            } else {
                if (first) STK_GS(1, true); else STK_GS(1, false);
            }
#undef STK_GS
#undef STK_GS4
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

static int smooth_fused(const stk_mg *mg, int l, bool backward, const double *c0, const double *c1,
                        const double *f, const double *uin, double *uout, int ld,
                        cudaStream_t s) {
    const Level &lv = mg->L[l];
    return gs_fused_run(backward ? lv.fused_bwd : lv.fused_fwd, mg->K, lv.fused_T, lv.ktab,
                        lv.nkinds, lv.v0, lv.v1, lv.d0, lv.d1, c0, c1, f, uin, uout, ld, s);
}

static int coarse_solve(const stk_mg *mg, const double *inv, const int *group, const double *f,
                        double *u, int ld, cudaStream_t s) {
    int n0 = mg->L[0].n;
    k_coarse_solve<<<blocks_for((int64_t)n0 * ld, 256), 256, 0, s>>>(n0, inv, group, f, u, ld);
    return check_launch("k_coarse_solve");
}

// MGM(j, u_j, f_j) of multigrid.py:168-182.  `zero_guess`: u holds no data yet
// and stands for u = 0 (the reference passes np.zeros, multigrid.py:176,187).
static int cycle(const stk_mg *mg, int l, const double *c0, const double *c1, const double *inv,
                 const int *group, const double *f, double *u, int ld, Workspace &ws,
                 cudaStream_t s, bool zero_guess) {
    if (l == 0) return coarse_solve(mg, inv, group, f, u, ld, s);
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
        STK_TRY(smooth_fused(mg, l, false, c0, c1, f, zero_guess ? nullptr : u, cur, ld, s));
    } else {
        if (zero_guess && mg->nu == 0)
            STK_TRY(check(cudaMemsetAsync(u, 0, sizeof(double) * (size_t)lv.n * ld, s),
                          "stk_mg: memset"));
        STK_TRY(smooth(mg, l, mg->nu, false, c0, c1, f, u, ld, s, zero_guess));
    }
    // res = A u - f, f_c = R res (multigrid.py:174).  A fused kernel that
    // recomputes the fine residuals per coarse row was measured 2.6x slower
    // (6 block passes of DRAM reads instead of ~3): the residual block costs
    // less than the lost locality.
    STK_TRY(launch_space_spmm(lv.n, lv.indptr, lv.indices, mg->K, lv.v0, lv.v1, c0, c1, cur, 1.0,
                              -1.0, f, ws.res, ld, s));
    STK_TRY(launch_space_spmm(lc.n, lv.r_indptr, lv.r_indices, 1, lv.r_vals, nullptr, nullptr,
                              nullptr, ws.res, 1.0, 0.0, nullptr, ws.f[l - 1], ld, s));
    STK_TRY(cycle(mg, l - 1, c0, c1, inv, group, ws.f[l - 1], ws.u[l - 1], ld, ws, s, true));
    // u -= P u_c
    STK_TRY(launch_space_spmm(lv.n, lv.p_indptr, lv.p_indices, 1, lv.p_vals, nullptr, nullptr,
                              nullptr, ws.u[l - 1], -1.0, 1.0, cur, cur, ld, s));
    if (fused) return smooth_fused(mg, l, true, c0, c1, f, cur, u, ld, s);
    return smooth(mg, l, mg->nu, true, c0, c1, f, u, ld, s);
}

extern "C" {

stk_mg *stk_mg_create(int nlevels, int smoothsteps, int vcycles, int K) {
    if (nlevels < 1 || (K != 1 && K != 2) || smoothsteps < 0 || vcycles < 0) {
        fail(-1, "stk_mg_create: bad arguments");
        return nullptr;
    }
    stk_mg *mg = new stk_mg;
    mg->nlevels = nlevels;
    mg->nu = smoothsteps;
    mg->vcycles = vcycles;
    mg->K = K;
    mg->L.resize(nlevels);
    return mg;
}

void stk_mg_destroy(stk_mg *mg) { delete mg; }

int stk_mg_set_level(stk_mg *mg, int level, int nrows, const int *indptr, const int *indices,
                     const double *vals0, const double *vals1, const double *diag0,
                     const double *diag1, const int *sched_rows, const int *phase_ptr_host,
                     int nphases) {
    if (!mg || level < 0 || level >= mg->nlevels) return fail(-1, "stk_mg_set_level: bad level");
    if (mg->K == 2 && (!vals1 || !diag1)) return fail(-1, "stk_mg_set_level: K = 2 needs vals1");
    Level &lv = mg->L[level];
    lv.n = nrows;
    lv.indptr = indptr;
    lv.indices = indices;
    lv.v0 = vals0;
    lv.v1 = vals1;
    lv.d0 = diag0;
    lv.d1 = diag1;
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
                     const double *ktab, int nkinds, int T) {
    if (!mg || level < 1 || level >= mg->nlevels) return fail(-1, "stk_mg_set_fused: bad level");
    if (T != 8) return fail(-1, "stk_mg_set_fused: T must be 8");
    Level &lv = mg->L[level];
    lv.fused_fwd = fwd;
    lv.fused_bwd = bwd;
    lv.ktab = ktab;
    lv.nkinds = nkinds;
    lv.fused_T = T;
    for (auto &kv : mg->graphs)  // captured sequences are stale now
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    mg->graphs.clear();
    return 0;
}

int stk_mg_apply(stk_mg *mg, const double *coef0, const double *coef1, const double *coarse_inv,
                 const int *coarse_group, const double *b, double *x, int ld, double *wsbuf,
                 void *stream) {
    if (!mg) return fail(-1, "stk_mg_apply: null handle");
    if (ld & 3) return fail(-1, "stk_mg_apply: pitch must be a multiple of 4");
    if (mg->K == 2 && (!coef0 || !coef1)) return fail(-1, "stk_mg_apply: K = 2 needs coefs");
    if (b == x) return fail(-1, "stk_mg_apply: b must not alias x");
    cudaStream_t s = as_stream(stream);
    const int top = mg->nlevels - 1;
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
            STK_TRY(cycle(mg, top, coef0, coef1, coarse_inv, coarse_group, b, x, ld, ws, st,
                          v == 0));
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
    GraphKey key{{coef0, coef1, coarse_inv, coarse_group, b, x, wsbuf}, ld};
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

int stk_mg_smooth(stk_mg *mg, int level, int nsweeps, int backward, const double *coef0,
                  const double *coef1, const double *f, double *u, int ld, void *stream) {
    if (!mg || level < 1 || level >= mg->nlevels) return fail(-1, "stk_mg_smooth: bad level");
    return smooth(mg, level, nsweeps, backward != 0, coef0, coef1, f, u, ld, as_stream(stream));
}

}  // extern "C"
