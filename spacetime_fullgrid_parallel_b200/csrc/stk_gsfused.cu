// Fused Gauss-Seidel smoother: nu lexicographic sweeps of one multigrid level
// (/root/reference/source/multigrid.py:89-97, :113-127) in ONE pass over HBM.
//
// The host compiler (gs_program.py) cuts the level into spatial items and
// turns the nu * D wavefront stages of the sweeps into a skewed pipeline over
// a sliding window of rows.  One CTA = (item, chunk of T time slices): the
// window lives in shared memory (up to ~200 KB), the CTA walks the item's
// macro-steps, and per macro-step
//   1. issues the cp.async loads that bring rows into free window slots
//      LOOKAHEAD steps before their first use,
//   2. runs the step's row updates  u_i += (f_i - A_i . u) / a_ii  with every
//      operand u_j read from the window (LPO lanes per row, one double2 of
//      time values per lane), storing final values of the item's own rows,
//   3. waits for the loads that must have landed and synchronises.
// Items recompute the rows of their dependency closure, so CTAs never wait on
// each other.  Rows with bitwise identical matrix values share a "kind" whose
// values sit in shared memory (a uniformly refined mesh has a handful per
// level); otherwise the values come from the level's CSR arrays.
#include <algorithm>
#include <queue>
#include <vector>

#include "stk_common.cuh"

namespace stk {

constexpr int GS_LOOKAHEAD = 2;  // must equal gs_program.LOOKAHEAD
constexpr int GS_PREFETCH = 4;   // must equal gs_program.PREFETCH

struct GsArgs {
    const int *item_step, *item_pass;
    const int2 *step_info;  // per macro-step: (end pass, end load), global
    const uint4 *rec;       // recq uint4 per record (gs_program.GSProgram)
    int recq;
    const int2 *lds;  // window loads: (row, window slot)
    const double *ktab;  // [nkinds][K * maxnnz + 2]
    int nkinds, kstride, maxnnz, nslots;
    const double *v0, *v1, *d0, *d1;  // generic path: CSR values, diagonals
    const double *coef0, *coef1;
    const double *f, *uin;
    double *uout;
    int ld, nchunks;
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__host__ __device__ __forceinline__ unsigned gs_table_stride(int K, int kstride) {
    if (K == 2) return ((kstride / 2) & 1) ? kstride : kstride + 2;
    return (kstride & 1) ? kstride : kstride + 1;
}
__device__ __forceinline__ unsigned slot_of(const uint4 &w, int q) {
    const unsigned v = (q < 2) ? w.x : (q < 4) ? w.y : (q < 6) ? w.z : w.w;
    return (q & 1) ? (v >> 16) : (v & 0xffffu);
}

// Shared memory: window [nslots][T] | value table | reciprocal diagonals |
// record ring [PD][NGRP] x 32 B | f ring [PD][NGRP][T].
//   kinds, K = 1: the kind table as it is (values, diagonal, 1/diagonal)
//   kinds, K = 2: the kind table as it is ((v0, v1) pairs, one broadcast read
//                 per entry) and rdiag[kind][T] = 1 / (c0(t) e0 + c1(t) e1)
//                 for THIS chunk's time values.  (A per-chunk table of the
//                 combined values c0 v0 + c1 v1 costs one FMA less per entry
//                 but four shared-memory wavefronts more per warp, and the
//                 kernel is bound by shared-memory bandwidth: measured.)
//   generic: nothing (values come from the CSR arrays through L1/L2)
// Programs with row kinds list a row's entries diagonal first and padded to
// NNZ entries (7 or 8) with zero-valued ones, so the row product is a fixed,
// branch-free sequence; generic programs walk the CSR row (NNZ = 0).
// Every global operand of an op (its record, its right-hand side) is copied
// into the rings with cp.async PD passes before the op runs -- the program's
// static pass layout gives those addresses in advance -- and completion is
// tracked by cp.async groups (one per pass): register prefetch does not work
// here, the counting scoreboards make a wait on an old load wait for the
// newest one too (measured: profiles/r2_experiments.md).
template <int LPO, int K, bool GEN, int NT, int NNZ>
__global__ void __launch_bounds__(NT, 1) k_gs_fused(const GsArgs a) {
    constexpr int T = 2 * LPO;
    constexpr int NGRP = NT / LPO;
    constexpr int PD = GS_PREFETCH;
    extern __shared__ __align__(16) double smem[];
    double *win = smem;
    double *vtab = win + (size_t)a.nslots * T;
    // table strides padded so that different kinds fall into different banks
    // (K = 2: 16-byte pairs, an odd number of them per kind; K = 1: an odd
    // number of doubles; reciprocal diagonals: T + 2 doubles per kind)
    const unsigned ks = gs_table_stride(K, a.kstride);
    constexpr unsigned RS = T + 2;
    const unsigned tabsz = (unsigned)a.nkinds * ks + (K == 2 ? (unsigned)a.nkinds * RS : 0u);
    double *rdiag = vtab + (size_t)a.nkinds * ks;  // K == 2 only: [nkinds][RS]
    uint4 *hring = reinterpret_cast<uint4 *>(vtab + ((tabsz + 1u) & ~1u));  // [PD][NGRP] headers
    uint4 *nring = hring + PD * NGRP;                                        // [PD][NGRP] slots
    double *fring = reinterpret_cast<double *>(nring + PD * NGRP);           // [PD][NGRP][T]

    const int item = blockIdx.x / a.nchunks;
    const int chunk = blockIdx.x - item * a.nchunks;
    const int lane = threadIdx.x % LPO;
    const int grp = threadIdx.x / LPO;
    const int t = chunk * T + 2 * lane;
    const bool valid = t < a.ld;
    const int woff = 2 * lane;

    double2 c0 = make_double2(1.0, 1.0), c1 = make_double2(1.0, 1.0);
    if (K == 2 && valid) {
        c0 = ldg2(a.coef0 + t);
        c1 = ldg2(a.coef1 + t);
    }
    if (!GEN) {
        for (int k = threadIdx.x; k < a.nkinds * a.kstride; k += NT) {
            const int kind = k / a.kstride;
            vtab[kind * ks + (k - kind * a.kstride)] = __ldg(a.ktab + k);
        }
        __syncthreads();
        if (K == 1) {
            for (int k = threadIdx.x; k < a.nkinds; k += NT)
                vtab[k * ks + a.maxnnz + 1] = 1.0 / vtab[k * ks + a.maxnnz];
        } else {
            for (int k = grp; k < a.nkinds; k += NGRP) {
                const double e0 = vtab[k * ks + 2 * a.maxnnz];
                const double e1 = vtab[k * ks + 2 * a.maxnnz + 1];
                double2 r;
                r.x = valid ? 1.0 / fma(c0.x, e0, c1.x * e1) : 0.0;
                r.y = valid ? 1.0 / fma(c0.y, e0, c1.y * e1) : 0.0;
                *reinterpret_cast<double2 *>(rdiag + k * RS + woff) = r;
            }
        }
    }
    const int m0 = __ldg(a.item_step + item), m1 = __ldg(a.item_step + item + 1);
    const int p0 = __ldg(a.item_pass + item), p_last = __ldg(a.item_pass + item + 1);
    const bool zero_guess = (a.uin == nullptr);
    const size_t tcol = (size_t)t;
    const double *fcol = a.f + tcol;
    // this thread's places in the rings (slot r adds r * NGRP entries)
    uint4 *myh = hring + grp, *myn = nring + grp;
    double *myf = fring + grp * T + woff;

    // fetch the record of pass q and the f row `frow` into ring slot r
    uint4 *myrec = (lane ? myn : myh);  // lanes 0 / 1 copy the two halves of a record
    auto prefetch = [&](const uint4 *src, int r, unsigned frow) {
        if (lane < 2) cp_async16(myrec + r * NGRP, src);
        if (valid) cp_async16(myf + r * (NGRP * T), fcol + (size_t)frow * a.ld);
    };
    const size_t rec_pass = (size_t)NGRP * a.recq;  // uint4 per pass
    const uint4 *pf_src = a.rec + ((size_t)p0 * NGRP + grp) * a.recq + (lane & 1);
    // ---- prologue: the first PD passes, one cp.async group each ----
    for (int k = 0; k < PD; ++k) {
        if (p0 + k < p_last) {
            const unsigned frow =
                __ldg(reinterpret_cast<const unsigned *>(
                    a.rec + ((size_t)(p0 + k) * NGRP + grp) * a.recq)) & 0x7fffffffu;
            prefetch(pf_src, k, frow);
        }
        pf_src += rec_pass;
        cp_async_commit();
    }

    auto issue_load = [&](const int2 &e) {
        double *dst = win + (unsigned)e.y * T + woff;
        if (zero_guess || !valid)
            *reinterpret_cast<double2 *>(dst) = make_double2(0.0, 0.0);
        else
            cp_async16(dst, a.uin + (size_t)e.x * a.ld + tcol);
    };

    int ld_beg = (m0 > 0) ? __ldg(&a.step_info[m0 - 1].y) : 0;
    int2 info = __ldg(a.step_info + m0);
    int2 ldpf = make_int2(0, 0);
    if (ld_beg + grp < info.y) ldpf = __ldg(a.lds + ld_beg + grp);
    int p = p0, r = 0;
    __syncthreads();  // tables ready

    for (int m = m0; m < m1; ++m) {
        const int2 info_nx = (m + 1 < m1) ? __ldg(a.step_info + m + 1) : info;
        // 1. window loads of this step (visible LOOKAHEAD + 1 steps later); they
        //    join the cp.async group of the step's first pass
        if (ld_beg + grp < info.y) {
            issue_load(ldpf);
            for (int e = ld_beg + grp + NGRP; e < info.y; e += NGRP) issue_load(__ldg(a.lds + e));
        }
        if (info.y + grp < info_nx.y) ldpf = __ldg(a.lds + info.y + grp);
        if (p == info.x) cp_async_commit();  // no pass in this step: its own group
        // 2. the passes of this step
        for (; p < info.x; ++p) {
            cp_async_wait<PD - 1>();  // the group that fetched pass p has landed
            __syncwarp();
            const uint4 h = myh[r * NGRP];
            uint4 nb = myn[r * NGRP];
            const double2 fv = *reinterpret_cast<const double2 *>(myf + r * (NGRP * T));
            __syncwarp();  // everyone has read slot r before it is refilled
            if (p + PD < p_last) prefetch(pf_src, r, h.y);
            pf_src += rec_pass;
            cp_async_commit();
            r = (r + 1 == PD) ? 0 : r + 1;
            const int nnz = (int)(h.w >> 16);
            if (nnz == 0) continue;  // padding
            const unsigned row = h.x & 0x7fffffffu;
            const bool store = (h.x >> 31) != 0u;
            const unsigned self = h.w & 0xffffu;
            double2 s0 = make_double2(0.0, 0.0), s1 = make_double2(0.0, 0.0);
            const double *kv = GEN ? nullptr : vtab + (unsigned)h.z * ks;
            double2 uo = make_double2(0.0, 0.0);
            if (!GEN) {
                // fixed trip count, entry 0 = the diagonal (u_i itself)
#pragma unroll
                for (int q = 0; q < (NNZ ? NNZ : 8); ++q) {
                    if (NNZ == 0 && q >= a.maxnnz) break;
                    const unsigned sl = slot_of(nb, q);
                    const double2 xv = *reinterpret_cast<const double2 *>(win + sl * T + woff);
                    if (q == 0) uo = xv;
                    if (K == 2) {
                        const double2 av = *reinterpret_cast<const double2 *>(kv + 2 * q);
                        s0.x = fma(av.x, xv.x, s0.x);
                        s0.y = fma(av.x, xv.y, s0.y);
                        s1.x = fma(av.y, xv.x, s1.x);
                        s1.y = fma(av.y, xv.y, s1.y);
                    } else {
                        const double a0 = kv[q];
                        s0.x = fma(a0, xv.x, s0.x);
                        s0.y = fma(a0, xv.y, s0.y);
                    }
                }
            } else {
                const size_t voff = (size_t)h.z;
                for (int base = 0; base < nnz; base += 8) {
                    if (base)
                        nb = __ldg(a.rec + ((size_t)p * NGRP + grp) * a.recq + 1 + (base >> 3));
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int e = base + q;
                        if (e < nnz) {
                            const unsigned sl = slot_of(nb, q);
                            const double2 xv =
                                *reinterpret_cast<const double2 *>(win + sl * T + woff);
                            if (sl == self) uo = xv;  // the diagonal entry: u_i itself
                            const double a0 = __ldg(a.v0 + voff + e);
                            s0.x = fma(a0, xv.x, s0.x);
                            s0.y = fma(a0, xv.y, s0.y);
                            if (K == 2) {
                                const double a1 = __ldg(a.v1 + voff + e);
                                s1.x = fma(a1, xv.x, s1.x);
                                s1.y = fma(a1, xv.y, s1.y);
                            }
                        }
                    }
                }
            }
            double *up = win + self * T + woff;
            if (GEN) {
                double2 dg;
                if (K == 2) {
                    s0.x = fma(c0.x, s0.x, c1.x * s1.x);
                    s0.y = fma(c0.y, s0.y, c1.y * s1.y);
                    const double e0 = __ldg(a.d0 + row), e1 = __ldg(a.d1 + row);
                    dg.x = fma(c0.x, e0, c1.x * e1);
                    dg.y = fma(c0.y, e0, c1.y * e1);
                } else {
                    dg.x = dg.y = __ldg(a.d0 + row);
                }
                uo.x += (fv.x - s0.x) / dg.x;
                uo.y += (fv.y - s0.y) / dg.y;
            } else if (K == 2) {
                s0.x = fma(c0.x, s0.x, c1.x * s1.x);
                s0.y = fma(c0.y, s0.y, c1.y * s1.y);
                const double2 rd =
                    *reinterpret_cast<const double2 *>(rdiag + (unsigned)h.z * RS + woff);
                uo.x = fma(fv.x - s0.x, rd.x, uo.x);
                uo.y = fma(fv.y - s0.y, rd.y, uo.y);
            } else {
                const double rd = kv[a.maxnnz + 1];
                uo.x = fma(fv.x - s0.x, rd, uo.x);
                uo.y = fma(fv.y - s0.y, rd, uo.y);
            }
            *reinterpret_cast<double2 *>(up) = uo;
            if (store && valid) stv2(a.uout + (size_t)row * a.ld + tcol, uo);
        }
        // 3. every step commits >= 1 group, so all but the 2 newest groups
        //    include the window loads issued LOOKAHEAD steps ago
        cp_async_wait<GS_LOOKAHEAD>();
        __syncthreads();
        ld_beg = info.y;
        info = info_nx;
    }
    cp_async_wait<0>();
}

}  // namespace stk

using namespace stk;

struct stk_gs_prog {
    int nitems, nslots, maxnnz, generic, recw, ngrp;
    const int *item_step, *item_pass;
    const int2 *step_info, *ld;
    const uint4 *rec;
};

template <int LPO, int K, bool GEN, int NT, int NNZ>
static int launch_gs_fused(const GsArgs &a, int nitems, size_t smem, cudaStream_t s) {
    auto kern = k_gs_fused<LPO, K, GEN, NT, NNZ>;
    static thread_local size_t configured = 0;
    if (smem > configured) {
        STK_TRY(check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem),
                      "stk_gs_fused: shared memory attribute"));
        configured = smem;
    }
    kern<<<(unsigned)nitems * (unsigned)a.nchunks, NT, smem, s>>>(a);
    return check_launch("k_gs_fused");
}

namespace stk {
// shared with stk_mg.cu
int gs_fused_run(const stk_gs_prog *pg, int K, int T, const double *ktab, int nkinds,
                 const double *v0, const double *v1, const double *d0, const double *d1,
                 const double *coef0, const double *coef1, const double *f, const double *uin,
                 double *uout, int ld, cudaStream_t s) {
    if (!pg) return fail(-1, "stk_gs_fused: null program");
    if (K != 1 && K != 2) return fail(-1, "stk_gs_fused: K must be 1 or 2");
    if (T != 8) return fail(-1, "stk_gs_fused: T must be 8");
    if (ld & 3) return fail(-1, "stk_gs_fused: pitch must be a multiple of 4");
    if (uin == uout) return fail(-1, "stk_gs_fused: u_in must not alias u_out");
    if (K == 2 && (!coef0 || !coef1)) return fail(-1, "stk_gs_fused: K = 2 needs coefficients");
    if (pg->generic ? (!v0 || !d0 || (K == 2 && (!v1 || !d1))) : (!ktab || nkinds < 1))
        return fail(-1, "stk_gs_fused: matrix values missing");
    GsArgs a;
    a.item_step = pg->item_step;
    a.item_pass = pg->item_pass;
    a.step_info = pg->step_info;
    a.rec = pg->rec;
    a.recq = pg->recw / 4;
    a.lds = pg->ld;
    a.ktab = ktab;
    a.nkinds = pg->generic ? 0 : nkinds;
    a.maxnnz = pg->maxnnz;
    a.kstride = K * pg->maxnnz + 2;
    a.nslots = pg->nslots;
    a.v0 = v0;
    a.v1 = v1;
    a.d0 = d0;
    a.d1 = d1;
    a.coef0 = coef0;
    a.coef1 = coef1;
    a.f = f;
    a.uin = uin;
    a.uout = uout;
    a.ld = ld;
    a.nchunks = (ld + T - 1) / T;
    size_t tab = (size_t)a.nkinds * gs_table_stride(K, a.kstride) +
                 (K == 2 ? (size_t)a.nkinds * (T + 2) : 0);
    tab = (tab + 1) & ~(size_t)1;
    size_t smem = sizeof(double) * ((size_t)pg->nslots * T + tab) +
                  (size_t)GS_PREFETCH * pg->ngrp * (32 + 8 * T);
    if (smem > 227 * 1024) return fail(-1, "stk_gs_fused: window does not fit shared memory");
#define STK_GSF(KK, GG, NN) launch_gs_fused<4, KK, GG, 512, NN>(a, pg->nitems, smem, s)
    if (pg->ngrp != 128)
        return fail(-1, "stk_gs_fused: program compiled for an unsupported group count");
    if (pg->generic) return K == 1 ? STK_GSF(1, true, 0) : STK_GSF(2, true, 0);
    if (pg->maxnnz > 8) return fail(-1, "stk_gs_fused: row kinds need <= 8 entries per row");
    if (pg->maxnnz == 7) return K == 1 ? STK_GSF(1, false, 7) : STK_GSF(2, false, 7);
    if (pg->maxnnz == 8) return K == 1 ? STK_GSF(1, false, 8) : STK_GSF(2, false, 8);
    return K == 1 ? STK_GSF(1, false, 0) : STK_GSF(2, false, 0);
#undef STK_GSF
}
}  // namespace stk

extern "C" {

stk_gs_prog *stk_gs_prog_create(int nitems, int nslots, int maxnnz, int generic, int recw,
                                int ngrp, const int *item_step, const int *item_pass,
                                const void *step_info, const void *rec, const void *ld) {
    if (nitems < 1 || nslots < 1 || nslots > 65535 || maxnnz < 1 || recw < 8 || (recw & 3) ||
        !item_step || !item_pass || !step_info || !rec || !ld) {
        fail(-1, "stk_gs_prog_create: bad arguments");
        return nullptr;
    }
    stk_gs_prog *p = new stk_gs_prog;
    p->nitems = nitems;
    p->nslots = nslots;
    p->maxnnz = maxnnz;
    p->generic = generic;
    p->recw = recw;
    p->ngrp = ngrp;
    p->item_step = item_step;
    p->item_pass = item_pass;
    p->step_info = (const int2 *)step_info;
    p->rec = (const uint4 *)rec;
    p->ld = (const int2 *)ld;
    return p;
}

void stk_gs_prog_destroy(stk_gs_prog *p) { delete p; }

int stk_gs_fused(const stk_gs_prog *pg, int K, int T, const double *ktab, int nkinds,
                 const double *v0, const double *v1, const double *d0, const double *d1,
                 const double *coef0, const double *coef1, const double *f, const double *uin,
                 double *uout, int ld, void *stream) {
    return gs_fused_run(pg, K, T, ktab, nkinds, v0, v1, d0, d1, coef0, coef1, f, uin, uout, ld,
                        as_stream(stream));
}

// Host helper (HOST pointers): interval colouring of window lifetimes.  Row q
// occupies a slot during the macro-steps [start[q], end[q]]; a slot is reused
// only by a row that starts strictly after its previous occupant ended.
// Returns the number of slots used.
int stk_gs_alloc_slots(int n, const int *start, const int *end, int *slot) {
    std::vector<int> order(n);
    for (int q = 0; q < n; ++q) order[q] = q;
    std::stable_sort(order.begin(), order.end(),
                     [&](int x, int y) { return start[x] < start[y]; });
    typedef std::pair<int, int> EndSlot;  // (end, slot), earliest end on top
    std::priority_queue<EndSlot, std::vector<EndSlot>, std::greater<EndSlot>> busy;
    std::priority_queue<int, std::vector<int>, std::greater<int>> free_slots;
    int used = 0;
    for (int k = 0; k < n; ++k) {
        int q = order[k];
        while (!busy.empty() && busy.top().first < start[q]) {
            free_slots.push(busy.top().second);
            busy.pop();
        }
        int s;
        if (!free_slots.empty()) {
            s = free_slots.top();
            free_slots.pop();
        } else {
            s = used++;
        }
        slot[q] = s;
        busy.push(EndSlot(end[q], s));
    }
    return used;
}

}  // extern "C"
