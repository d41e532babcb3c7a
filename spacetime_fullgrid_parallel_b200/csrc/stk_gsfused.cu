// Fused Gauss-Seidel smoother: nu lexicographic sweeps of one multigrid level
// (/root/reference/source/multigrid.py:89-97, :113-127) in ONE pass over HBM.
//
// The host compiler (gs_program.py) cuts the level into spatial items and
// turns the nu * D wavefront stages of the sweeps into a skewed pipeline over
// a sliding window of rows.  One CTA = (item, chunk of T time slices): the
// window lives in shared memory (up to ~200 KB), the CTA walks the item's
// macro-steps, and per macro-step
//   1. issues the cp.async loads that bring rows into free window slots
//      LOOKAHEAD steps before their first use (the program itself streams in
//      through a ring of TMA bulk copies, one per pass of 128 row updates),
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
constexpr int GS_RING = 8;       // pass slots of the record ring (>= 2 * GS_MAXPASS + 2)
constexpr int GS_FRING = 4;      // passes between requesting an op's f and using it (= gs_program.PREFETCH)
constexpr int GS_MAXPASS = 3;    // passes per macro-step the host may emit (gs_program.MAX_PASSES)
static_assert(GS_RING >= 2 * GS_MAXPASS + 2, "record ring too small for the passes of two steps");

struct GsArgs {
    const int *item_step, *item_pass;
    const int2 *step_info;  // per macro-step: (end pass, end load), global
    const uint4 *rec;       // [npasses][recq][ngrp] uint4 (gs_program.GSProgram)
    int recq;
    const int2 *lds;  // window loads: (row, window slot)
    // matrix values: one matrix per GROUP of time slices, grp[t] = group of t
    const double *ktab;  // kinds: [G][nkinds][maxnnz + 2], entries in program order
    int G, nkinds, kstride, maxnnz, nslots, bulk_kind;
    const double *cvals;  // generic: [G][vstride] values in program entry order
    size_t vstride;
    const int *grp;
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
__device__ __forceinline__ unsigned smem_u32(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// TMA bulk copy global -> shared, completion counted on an mbarrier (UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes,
                                             void *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    unsigned done = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        if (++spins > (1u << 24)) __trap();  // a lost copy must not hang the GPU
    }
}
// shared-memory stride of a kind's value row: an odd number of doubles, so that
// different (group, kind) rows fall into different banks
__host__ __device__ __forceinline__ unsigned gs_table_stride(int kstride) {
    return (kstride & 1) ? kstride : kstride + 1;
}
__device__ __forceinline__ unsigned slot_of(const uint4 &w, int q) {
    const unsigned v = (q < 2) ? w.x : (q < 4) ? w.y : (q < 6) ? w.z : w.w;
    return (q & 1) ? (v >> 16) : (v & 0xffffu);
}

// Shared memory: window [nslots][T] | value table | reciprocal diagonals |
// record ring [GS_RING][recq][NGRP] x 16 B | GS_RING mbarriers | f ring
// [GS_FRING][NGRP][T].
// Matrix values.  Every time slice belongs to a group g(t) with its own matrix
// on the shared pattern (GROUPED; one group otherwise).  A program lists a
// row's entries in canonical order (diagonal first, then sorted by value), so
// rows of a uniformly refined mesh fall into a handful of KINDS with identical
// value sequences, and one of them -- the interior stencil -- covers almost
// every row:
//   kinds:   vtab[g][kind][entry] in shared memory; the BULK kind's values for
//            this lane's two time values live in registers (no shared-memory
//            read per entry for ~all ops); rdiag[kind][T] = 1 / diagonal
//   generic: cvals[g][csr offset + entry] through L1/L2, a division per update
// Programs with kinds pad rows to NNZ entries (7 or 8) with zero-valued ones:
// the row product is a fixed, branch-free sequence.
// The program is a stream: one thread feeds a ring of GS_RING pass slots with
// TMA bulk copies (one contiguous copy per pass of NGRP records, completion on
// the slot's mbarrier), refilling after each macro-step's barrier, when every
// warp is done with the passes before it.  The right-hand side f of an op is a
// scattered 64-byte piece: it comes through a per-group cp.async ring, requested
// GS_FRING passes ahead (the record names that pass's row) and tracked by
// cp.async groups, one per pass.  Register prefetch was measured and does not
// work here: one pass ahead does not cover the latency, and deeper prefetches
// share a counting scoreboard, so a wait on the oldest load waits for the
// newest one too (profiles/r2_experiments.md).
template <int LPO, bool GROUPED, bool GEN, int NT, int NNZ>
__global__ void __launch_bounds__(NT, 1) k_gs_fused(const GsArgs a) {
    constexpr int T = 2 * LPO;
    constexpr int NGRP = NT / LPO;
    constexpr bool BULK = !GEN && NNZ > 0;  // bulk kind's values in registers
    constexpr int NB = BULK ? NNZ : 1;
    extern __shared__ __align__(16) double smem[];
    double *win = smem;
    double *vtab = win + (size_t)a.nslots * T;
    const unsigned ks = gs_table_stride(a.kstride);
    constexpr unsigned RS = T + 2;  // odd number of 16-byte pieces per kind
    const unsigned ntab = GEN ? 0u : (unsigned)a.G * a.nkinds * ks;
    double *rdiag = vtab + ((ntab + 1u) & ~1u);  // [nkinds][RS]
    const unsigned nrd = GEN ? 0u : (unsigned)a.nkinds * RS;
    uint4 *ring = reinterpret_cast<uint4 *>(rdiag + nrd);  // [RING][recq][NGRP]
    const unsigned pass_q = (unsigned)NGRP * a.recq;        // uint4 per pass
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(ring + GS_RING * pass_q);
    double *fring = reinterpret_cast<double *>(mbar + GS_RING);  // [GS_FRING][NGRP][T]

    const int item = blockIdx.x / a.nchunks;
    const int chunk = blockIdx.x - item * a.nchunks;
    const int lane = threadIdx.x % LPO;
    const int grp = threadIdx.x / LPO;
    const int t = chunk * T + 2 * lane;
    const bool valid = t < a.ld;
    const int woff = 2 * lane;

    unsigned g0 = 0, g1 = 0;  // groups of this lane's two time values
    if (GROUPED && valid) {
        g0 = (unsigned)__ldg(a.grp + t);
        g1 = (unsigned)__ldg(a.grp + t + 1);
    }
    double b0[NB], b1[NB];  // bulk kind: entry values for t and t + 1
    double2 brd = make_double2(0.0, 0.0);
    if (!GEN) {
        for (int k = threadIdx.x; k < a.G * a.nkinds * a.kstride; k += NT) {
            const int row = k / a.kstride;
            vtab[row * ks + (k - row * a.kstride)] = __ldg(a.ktab + k);
        }
        __syncthreads();
        for (int k = grp; k < a.nkinds; k += NGRP) {  // 1 / diagonal (entry 0)
            double2 r;
            r.x = valid ? 1.0 / vtab[(g0 * a.nkinds + k) * ks] : 0.0;
            r.y = valid ? 1.0 / vtab[(g1 * a.nkinds + k) * ks] : 0.0;
            *reinterpret_cast<double2 *>(rdiag + k * RS + woff) = r;
        }
        if (BULK) {
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                b0[q] = vtab[(g0 * a.nkinds + a.bulk_kind) * ks + q];
                b1[q] = vtab[(g1 * a.nkinds + a.bulk_kind) * ks + q];
            }
            brd.x = valid ? 1.0 / b0[0] : 0.0;
            brd.y = valid ? 1.0 / b1[0] : 0.0;
        }
    }
    const int m0 = __ldg(a.item_step + item), m1 = __ldg(a.item_step + item + 1);
    const int p0 = __ldg(a.item_pass + item), p_last = __ldg(a.item_pass + item + 1);
    const bool zero_guess = (a.uin == nullptr);
    const size_t tcol = (size_t)t;
    const double *fcol = a.f + tcol;
    // ---- record ring: producer = thread 0 ----
    if (threadIdx.x == 0) {
        for (int r = 0; r < GS_RING; ++r) mbar_init(mbar + r, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int issued = p0;  // next pass to copy (thread 0)
    auto feed = [&](int consumed) {  // every pass < consumed has been read by all warps
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        while (issued < p_last && issued < consumed + GS_RING) {
            const int r = (issued - p0) & (GS_RING - 1);
            mbar_expect_tx(mbar + r, pass_q * 16u);
            tma_bulk_g2s(ring + r * pass_q, a.rec + (size_t)issued * pass_q, pass_q * 16u,
                         mbar + r);
            ++issued;
        }
    };
    if (threadIdx.x == 0) feed(p0);
    // ---- f ring: the first GS_FRING passes, one cp.async group each ----
    double *myf = fring + grp * T + woff;
    for (int k = 0; k < GS_FRING; ++k) {
        if (valid && p0 + k < p_last) {
            const unsigned frow = __ldg(&(a.rec + (size_t)(p0 + k) * pass_q + grp)->x) & 0x7fffffffu;
            cp_async16(myf + k * (NGRP * T), fcol + (size_t)frow * a.ld);
        }
        cp_async_commit();
    }
    int fr = 0;  // f ring slot of the current pass
    unsigned rr = 0, rpar = 0;  // record ring slot of the current pass and its phase parity

    auto issue_load = [&](const int2 &e) {
        double *dst = win + (unsigned)e.y * T + woff;
        // the row's right-hand side is first needed a few steps from now: have it
        // in L2 by then (one sector pair per row and chunk)
        if (valid && lane == 0)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(fcol + (size_t)e.x * a.ld));
        if (zero_guess || !valid)
            *reinterpret_cast<double2 *>(dst) = make_double2(0.0, 0.0);
        else
            cp_async16(dst, a.uin + (size_t)e.x * a.ld + tcol);
    };

    int ld_beg = (m0 > 0) ? __ldg(&a.step_info[m0 - 1].y) : 0;
    int2 info = __ldg(a.step_info + m0);
    // the step table is read two steps ahead: the next step's entry is needed
    // right after this step's loads are issued (no exposed L2 latency)
    int2 info_nx = (m0 + 1 < m1) ? __ldg(a.step_info + m0 + 1) : info;
    int2 ldpf = make_int2(0, 0);
    if (ld_beg + grp < info.y) ldpf = __ldg(a.lds + ld_beg + grp);
    int p = p0;
    __syncthreads();  // tables ready

    for (int m = m0; m < m1; ++m) {
        const int2 info_nn = (m + 2 < m1) ? __ldg(a.step_info + m + 2) : info_nx;
        // 1. window loads of this step (visible LOOKAHEAD + 1 steps later); they
        //    join the cp.async group of the step's first pass
        if (ld_beg + grp < info.y) {
            issue_load(ldpf);
            for (int e = ld_beg + grp + NGRP; e < info.y; e += NGRP) issue_load(__ldg(a.lds + e));
        }
        if (info.y + grp < info_nx.y) ldpf = __ldg(a.lds + info.y + grp);
        if (p == info.x) cp_async_commit();  // no pass in this step: its own group
        // 2. the passes of this step (window loads join the first pass's group)
        for (; p < info.x; ++p) {
            mbar_wait(mbar + rr, rpar);
            const uint4 *slot = ring + rr * pass_q + grp;
            rr = (rr + 1) & (GS_RING - 1);
            rpar ^= (rr == 0);
            const uint4 h = slot[0];
            uint4 nb = slot[NGRP];
            cp_async_wait<GS_FRING - 1>();  // this pass's f has landed (own copy)
            const double2 fv = *reinterpret_cast<const double2 *>(myf + fr * (NGRP * T));
            if (valid && p + GS_FRING < p_last)
                cp_async16(myf + fr * (NGRP * T), fcol + (size_t)h.y * a.ld);
            cp_async_commit();
            fr = (fr + 1 == GS_FRING) ? 0 : fr + 1;
            const int nnz = (int)(h.w >> 16);
            if (nnz == 0) continue;  // padding
            const unsigned row = h.x & 0x7fffffffu;
            const bool store = (h.x >> 31) != 0u;
            const unsigned self = h.w & 0xffffu;
            double2 s = make_double2(0.0, 0.0), uo = make_double2(0.0, 0.0), rd;
            if (!GEN) {
                // fixed trip count, entry 0 = the diagonal (u_i itself).  Almost
                // every row is of the bulk kind: its own branch keeps the table
                // reads and the register/table selects out of that path (rows of
                // one warp rarely mix kinds, so the branch seldom diverges).
                if (BULK && (int)h.z == a.bulk_kind) {
#pragma unroll
                    for (int e = 0; e < NB; ++e) {
                        const unsigned sl = slot_of(nb, e);
                        const double2 xv = *reinterpret_cast<const double2 *>(win + sl * T + woff);
                        if (e == 0) uo = xv;
                        s.x = fma(b0[e], xv.x, s.x);
                        s.y = fma(b1[e], xv.y, s.y);
                    }
                    rd = brd;
                } else {
                    const double *kv0 = vtab + (g0 * a.nkinds + h.z) * ks;
                    const double *kv1 = GROUPED ? vtab + (g1 * a.nkinds + h.z) * ks : kv0;
#pragma unroll
                    for (int e = 0; e < (NNZ ? NNZ : 8); ++e) {
                        if (NNZ == 0 && e >= a.maxnnz) break;
                        const unsigned sl = slot_of(nb, e);
                        const double2 xv = *reinterpret_cast<const double2 *>(win + sl * T + woff);
                        if (e == 0) uo = xv;
                        const double a0 = kv0[e];
                        const double a1 = GROUPED ? kv1[e] : a0;
                        s.x = fma(a0, xv.x, s.x);
                        s.y = fma(a1, xv.y, s.y);
                    }
                    rd = *reinterpret_cast<const double2 *>(rdiag + (unsigned)h.z * RS + woff);
                }
            } else {
                const double *c0 = a.cvals + (size_t)g0 * a.vstride + h.z;
                const double *c1 = a.cvals + (size_t)g1 * a.vstride + h.z;
                rd.x = __ldg(c0);  // the diagonal: entry 0
                rd.y = GROUPED ? __ldg(c1) : rd.x;
                for (int base = 0; base < nnz; base += 8) {
                    if (base) nb = slot[NGRP * (1 + (base >> 3))];
#pragma unroll
                    for (int qq = 0; qq < 8; ++qq) {
                        const int e = base + qq;
                        if (e < nnz) {
                            const unsigned sl = slot_of(nb, qq);
                            const double2 xv =
                                *reinterpret_cast<const double2 *>(win + sl * T + woff);
                            if (e == 0) uo = xv;
                            const double a0 = __ldg(c0 + e);
                            s.x = fma(a0, xv.x, s.x);
                            s.y = fma(GROUPED ? __ldg(c1 + e) : a0, xv.y, s.y);
                        }
                    }
                }
            }
            double *up = win + self * T + woff;
            if (GEN) {
                uo.x += (fv.x - s.x) / rd.x;
                uo.y += (fv.y - s.y) / rd.y;
            } else {
                uo.x = fma(fv.x - s.x, rd.x, uo.x);
                uo.y = fma(fv.y - s.y, rd.y, uo.y);
            }
            *reinterpret_cast<double2 *>(up) = uo;
            if (store && valid) stv2(a.uout + (size_t)row * a.ld + tcol, uo);
        }
        // 3. every step commits >= 1 group, so all but the 2 newest groups include
        //    the window loads issued LOOKAHEAD steps ago; publish the updates; the
        //    record ring slots of this step's passes are free again
        cp_async_wait<GS_LOOKAHEAD>();
        __syncthreads();
        if (threadIdx.x == 0) feed(p);
        ld_beg = info.y;
        info = info_nx;
        info_nx = info_nn;
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

template <int LPO, bool GROUPED, bool GEN, int NT, int NNZ>
static int launch_gs_fused(const GsArgs &a, int nitems, size_t smem, cudaStream_t s) {
    auto kern = k_gs_fused<LPO, GROUPED, GEN, NT, NNZ>;
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
int gs_fused_run(const stk_gs_prog *pg, int G, int T, const double *ktab, int nkinds,
                 int bulk_kind, const double *cvals, size_t vstride, const int *grp,
                 const double *f, const double *uin, double *uout, int ld, cudaStream_t s) {
    if (!pg) return fail(-1, "stk_gs_fused: null program");
    if (G < 1) return fail(-1, "stk_gs_fused: G must be >= 1");
    if (T != 8) return fail(-1, "stk_gs_fused: T must be 8");
    if (ld & 3) return fail(-1, "stk_gs_fused: pitch must be a multiple of 4");
    if (uin == uout) return fail(-1, "stk_gs_fused: u_in must not alias u_out");
    if (G > 1 && !grp) return fail(-1, "stk_gs_fused: several groups need the group table");
    if (pg->generic ? !cvals : (!ktab || nkinds < 1))
        return fail(-1, "stk_gs_fused: matrix values missing");
    GsArgs a;
    a.item_step = pg->item_step;
    a.item_pass = pg->item_pass;
    a.step_info = pg->step_info;
    a.rec = pg->rec;
    a.recq = pg->recw / 4;
    a.lds = pg->ld;
    a.ktab = ktab;
    a.G = G;
    a.nkinds = pg->generic ? 0 : nkinds;
    a.maxnnz = pg->maxnnz;
    a.kstride = pg->maxnnz + 2;
    a.nslots = pg->nslots;
    a.bulk_kind = bulk_kind;
    a.cvals = cvals;
    a.vstride = vstride;
    a.grp = grp;
    a.f = f;
    a.uin = uin;
    a.uout = uout;
    a.ld = ld;
    a.nchunks = (ld + T - 1) / T;
    size_t tab = 0;
    if (!pg->generic) {
        tab = (size_t)G * a.nkinds * gs_table_stride(a.kstride);
        tab = ((tab + 1) & ~(size_t)1) + (size_t)a.nkinds * (T + 2);
    }
    size_t smem = sizeof(double) * ((size_t)pg->nslots * T + tab) +
                  (size_t)GS_RING * pg->ngrp * pg->recw * 4 + GS_RING * 8 +
                  (size_t)GS_FRING * pg->ngrp * T * 8;
    if (smem > 227 * 1024) return fail(-1, "stk_gs_fused: window does not fit shared memory");
    if (pg->ngrp != 128)
        return fail(-1, "stk_gs_fused: program compiled for an unsupported group count");
    const bool grouped = (grp != nullptr);
#define STK_GSF(GG, GEN_, NN) launch_gs_fused<4, GG, GEN_, 512, NN>(a, pg->nitems, smem, s)
    if (pg->generic) return grouped ? STK_GSF(true, true, 0) : STK_GSF(false, true, 0);
    if (pg->maxnnz > 8) return fail(-1, "stk_gs_fused: row kinds need <= 8 entries per row");
    if (bulk_kind < 0 || bulk_kind >= nkinds) return fail(-1, "stk_gs_fused: bad bulk kind");
    if (pg->maxnnz == 7) return grouped ? STK_GSF(true, false, 7) : STK_GSF(false, false, 7);
    if (pg->maxnnz == 8) return grouped ? STK_GSF(true, false, 8) : STK_GSF(false, false, 8);
    return grouped ? STK_GSF(true, false, 0) : STK_GSF(false, false, 0);
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

int stk_gs_fused(const stk_gs_prog *pg, int G, int T, const double *ktab, int nkinds,
                 int bulk_kind, const double *cvals, int64_t vstride, const int *group,
                 const double *f, const double *uin, double *uout, int ld, void *stream) {
    return gs_fused_run(pg, G, T, ktab, nkinds, bulk_kind, cvals, (size_t)vstride, group, f, uin,
                        uout, ld, as_stream(stream));
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
