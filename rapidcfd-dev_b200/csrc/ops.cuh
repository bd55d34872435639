// ops.cuh -- engine operators (row-gather family) and generic elementwise / scalar
// kernels used by the solvers.  See engine.cuh for the Op concept.
#pragma once
#include "engine.cuh"

// device-resident solver state: alpha/beta and the convergence decision never leave
// the GPU inside the iteration loop (the reference synchronises the host three times
// per PCG iteration: PCG.C:142,166,195).
struct SolverScalars {
    double sum[8];  // latest (all-reduced) sums
    double wArA, wArAold, wApA, alpha, beta, omega;
    double rA0rA, rA0rAold, tAtA, tAsA;
    double normFactor, initialResidual, finalResidual, average, sumMag0;
    double tolerance, relTol, nCellsGlobal;
    int nIterations, converged, singular, stop;
    int maxIter, minIter, histCap, nSweeps;
    int bodies; // fused PCG: iteration bodies whose A.p product has been formed
    // fused PCG, deferred scalar steps (DeferredStep): steps completed (low 32 bits) | stop (bit 63), published by CTA 0 of
    // each sweep with one release store and read by every other CTA with one acquire load.  On its own 128-byte line so
    // that the waiting CTAs' polling does not slow down CTA 0's writes to the scalars above.
    alignas(128) unsigned long long gate;
    unsigned long long gatePad[15];
};

struct OpBase {
    const int *stop = nullptr;
    double *partials = nullptr;
    int waitHalo = 0; // 1: a peer-memory halo exchange of the staged vector is in flight
    // first thing every CTA of an engine kernel does; false => the CTA exits (the device has decided to stop)
    __device__ __forceinline__ bool prologue(const LayoutDev &) const { return !(stop && *stop); }
};

// ---- Amul / Tmul: out = diag*x + sum v*x[c]  (lduMatrixATmul.C:78-137) ----
// MODE 0 plain; 1: sum(out*x) [PCG wApA];  2: sum(aux*x), sum(out*x) [GAMG scale: num =
// <source,field>, den = <Acf,field>, GAMGSolverScale.C:80-125];  3: sum(out*aux) [PBiCG
// wApT, PBiCGStab rA0AyA];  4: sum(out*out), sum(out*aux) [PBiCGStab tAtA, tAsA]
template <int MODE>
struct AmulOp : OpBase {
    static constexpr int NVEC = 1;
    static constexpr int NRED = (MODE == 0) ? 0 : ((MODE == 1 || MODE == 3) ? 1 : 2);
    static constexpr bool LOCAL = false;
    const double *x, *diag, *aux;
    double *out;
    __device__ __forceinline__ void stage(int g, double &a, double &) const { a = x[g]; }
    __device__ __forceinline__ double pack_val(int row) const { return x[row]; }
    __device__ __forceinline__ void stage_own(int r, double2 &a, double2 &) const
    {
        a = *reinterpret_cast<const double2 *>(x + r);
    }
    __device__ __forceinline__ double init(int r, double a, double) const
    {
        return __dmul_rn(diag[r], a);
    }
    __device__ __forceinline__ double term(double acc, double v, double a, double) const
    {
        return __dadd_rn(acc, __dmul_rn(v, a));
    }
    __device__ __forceinline__ void finish(int r, double acc0, double acc1, double a0, double,
                                           double a1, double, double *red) const
    {
        *reinterpret_cast<double2 *>(out + r) = make_double2(acc0, acc1);
        if (MODE == 1) red[0] += acc0 * a0 + acc1 * a1;
        if (MODE == 2) {
            red[0] += aux[r] * a0 + aux[r + 1] * a1;
            red[1] += acc0 * a0 + acc1 * a1;
        }
        if (MODE == 3) red[0] += acc0 * aux[r] + acc1 * aux[r + 1];
        if (MODE == 4) {
            red[0] += acc0 * acc0 + acc1 * acc1;
            red[1] += acc0 * aux[r] + acc1 * aux[r + 1];
        }
    }
};

// ---- AINV ("DIC"/"DILU"): w = rD*(r - sum v*(rD[c]*r[c]))  (AINVPreconditionerF.H:42-99)
// The reference evaluates upper[f]*rD[nei]*r[nei] left to right; here the band stages
// t = rD*r once (one shared-memory tile, one gather per halo column and vector) and the
// sweep streams the matrix's own coefficients -- no scaled copy of the matrix, exactly the
// bytes of an Amul, and it works on the shared-coefficient layout.  The association differs
// from the reference's by one rounding per term: parity with the oracle is to ~1e-15
// relative per row, not bit-exact (tests state the tolerance).  Optional fused dot
// <w, dotv> (wArA with dotv = r; wArT with dotv = rT).
template <int NRED_>
struct AinvOp : OpBase {
    static constexpr int NVEC = 1, NRED = NRED_;
    static constexpr bool LOCAL = true;
    const double *r, *rD, *dotv;
    double *out;
    __device__ __forceinline__ void stage(int g, double &a, double &) const { a = __dmul_rn(rD[g], r[g]); }
    __device__ __forceinline__ double pack_val(int) const { return 0.0; }
    __device__ __forceinline__ void stage_own(int row, double2 &a, double2 &) const
    {
        double2 rr = *reinterpret_cast<const double2 *>(r + row);
        double2 dd = *reinterpret_cast<const double2 *>(rD + row);
        a = make_double2(__dmul_rn(dd.x, rr.x), __dmul_rn(dd.y, rr.y));
    }
    __device__ __forceinline__ double init(int, double, double) const { return 0.0; }
    __device__ __forceinline__ double term(double acc, double v, double t, double) const
    {
        return __dadd_rn(acc, __dmul_rn(v, t));
    }
    __device__ __forceinline__ void finish(int row, double acc0, double acc1, double, double, double, double,
                                           double *red) const
    {
        double2 d = *reinterpret_cast<const double2 *>(rD + row);
        double2 rr = *reinterpret_cast<const double2 *>(r + row);
        double w0 = __dmul_rn(d.x, __dsub_rn(rr.x, acc0));
        double w1 = __dmul_rn(d.y, __dsub_rn(rr.y, acc1));
        *reinterpret_cast<double2 *>(out + row) = make_double2(w0, w1);
        if (NRED == 1) {
            double d0 = dotv ? dotv[row] : rr.x, d1 = dotv ? dotv[row + 1] : rr.y;
            red[0] += w0 * d0 + w1 * d1;
        }
    }
};

// ---- Jacobi sweep (JacobiSmootherF.H:51-109; omega-damped, old psi everywhere) ----
struct JacobiOp : OpBase {
    static constexpr int NVEC = 1, NRED = 0;
    static constexpr bool LOCAL = false;
    const double *x, *diag, *b;
    double *out;
    double omega;
    int nCells;
    __device__ __forceinline__ void stage(int g, double &a, double &) const { a = x[g]; }
    __device__ __forceinline__ double pack_val(int row) const { return x[row]; }
    __device__ __forceinline__ void stage_own(int r, double2 &a, double2 &) const
    {
        a = *reinterpret_cast<const double2 *>(x + r);
    }
    __device__ __forceinline__ double init(int, double, double) const { return 0.0; }
    __device__ __forceinline__ double term(double acc, double v, double a, double) const
    {
        return __dadd_rn(acc, __dmul_rn(v, a));
    }
    __device__ __forceinline__ double row(int r, double acc, double a) const
    {
        double rD = __ddiv_rn(1.0, diag[r]);
        double w = __dmul_rn(omega, rD);
        double extra = __dadd_rn(__dmul_rn(1 - omega, a), __dmul_rn(w, b[r]));
        return __dsub_rn(extra, __dmul_rn(w, acc));
    }
    __device__ __forceinline__ void finish(int r, double acc0, double acc1, double a0, double,
                                           double a1, double, double *) const
    {
        *reinterpret_cast<double2 *>(out + r) = make_double2(row(r, acc0, a0), row(r + 1, acc1, a1));
    }
};

// ---- residual: rA = (b - diag*x) - sum v*x ; fused sum|rA|  (lduMatrixATmul.C:397-496)
template <int NRED_>
struct ResidualOp : OpBase {
    static constexpr int NVEC = 1, NRED = NRED_;
    static constexpr bool LOCAL = false;
    const double *x, *diag, *b;
    double *out;
    __device__ __forceinline__ void stage(int g, double &a, double &) const { a = x[g]; }
    __device__ __forceinline__ double pack_val(int row) const { return x[row]; }
    __device__ __forceinline__ void stage_own(int r, double2 &a, double2 &) const
    {
        a = *reinterpret_cast<const double2 *>(x + r);
    }
    __device__ __forceinline__ double init(int r, double a, double) const
    {
        return __dsub_rn(b[r], __dmul_rn(diag[r], a));
    }
    __device__ __forceinline__ double term(double acc, double v, double a, double) const
    {
        return __dsub_rn(acc, __dmul_rn(v, a));
    }
    __device__ __forceinline__ void finish(int r, double acc0, double acc1, double, double, double,
                                           double, double *red) const
    {
        *reinterpret_cast<double2 *>(out + r) = make_double2(acc0, acc1);
        if (NRED == 1) red[0] += fabs(acc0) + fabs(acc1);
    }
};

// ---- coefficient-only row sums: sumA (all entries, + diag), H1 (local, negated) ----
template <bool LOCAL_, bool NEG_>
struct CoeffSumOp : OpBase {
    static constexpr int NVEC = 0, NRED = 0;
    static constexpr bool LOCAL = LOCAL_;
    const double *diag; // nullptr => start from 0
    double *out;
    __device__ __forceinline__ void stage(int, double &, double &) const {}
    __device__ __forceinline__ double pack_val(int) const { return 0.0; }
    __device__ __forceinline__ void stage_own(int, double2 &, double2 &) const {}
    __device__ __forceinline__ double init(int r, double, double) const
    {
        return diag ? diag[r] : 0.0;
    }
    __device__ __forceinline__ double term(double acc, double v, double, double) const
    {
        return NEG_ ? __dsub_rn(acc, v) : __dadd_rn(acc, v);
    }
    __device__ __forceinline__ void finish(int r, double acc0, double acc1, double, double, double,
                                           double, double *) const
    {
        *reinterpret_cast<double2 *>(out + r) = make_double2(acc0, acc1);
    }
};

// ---- off-diagonal product sums: H = -(sum v*x) local only (lduMatrixOperations.C:107-155);
// GAMG interpolate: psi' = -(sum v*x)/diag over all entries (GAMGSolverInterpolate.C:45-110)
template <bool LOCAL_, bool DIVDIAG_>
struct OffDiagOp : OpBase {
    static constexpr int NVEC = 1, NRED = 0;
    static constexpr bool LOCAL = LOCAL_;
    const double *x, *diag;
    double *out;
    __device__ __forceinline__ void stage(int g, double &a, double &) const { a = x[g]; }
    __device__ __forceinline__ double pack_val(int row) const { return x[row]; }
    __device__ __forceinline__ void stage_own(int r, double2 &a, double2 &) const
    {
        a = *reinterpret_cast<const double2 *>(x + r);
    }
    __device__ __forceinline__ double init(int, double, double) const { return 0.0; }
    __device__ __forceinline__ double term(double acc, double v, double a, double) const
    {
        return DIVDIAG_ ? __dadd_rn(acc, __dmul_rn(v, a)) : __dsub_rn(acc, __dmul_rn(v, a));
    }
    __device__ __forceinline__ void finish(int r, double acc0, double acc1, double, double, double,
                                           double, double *) const
    {
        if (DIVDIAG_) {
            acc0 = __ddiv_rn(-acc0, diag[r]);
            acc1 = __ddiv_rn(-acc1, diag[r + 1]);
        }
        *reinterpret_cast<double2 *>(out + r) = make_double2(acc0, acc1);
    }
};

// ---------------------------------------------------------------------------
// elementwise kernels with fused reductions.  f(i2, red) handles elements 2*i2, 2*i2+1
// (vectors are 16-byte aligned and of even length => 128-bit loads/stores).
// ---------------------------------------------------------------------------
constexpr int EW_THREADS = 256;

template <int NRED, class F>
__global__ void __launch_bounds__(EW_THREADS) ew_kernel(int n2, const int *stop, double *partials, F f)
{
    if (stop && *stop) return;
    double red[NRED > 0 ? NRED : 1];
#pragma unroll
    for (int k = 0; k < (NRED > 0 ? NRED : 1); k++) red[k] = 0;
    for (int i = blockIdx.x * EW_THREADS + threadIdx.x; i < n2; i += gridDim.x * EW_THREADS) f(i, red);
    if (NRED > 0) block_reduce_store<(NRED > 0 ? NRED : 1), EW_THREADS>(red, partials, blockIdx.x);
}

inline int ew_grid(const b200ldu_ctx *ctx, int n2)
{
    int want = (n2 + EW_THREADS - 1) / EW_THREADS;
    int cap = ctx->smCount * 8;
    return want < 1 ? 1 : (want < cap ? want : cap);
}

template <int NRED, class F>
int ew_launch(b200ldu_ctx *ctx, int n2, const int *stop, double *partials, int *nPartialsOut, F f)
{
    int g = ew_grid(ctx, n2);
    ew_kernel<NRED, F><<<g, EW_THREADS, 0, ctx->stream>>>(n2, stop, partials, f);
    ctx->launches++;
    if (nPartialsOut) *nPartialsOut = g;
    KERNEL_CHECK();
    return B200LDU_OK;
}

// element-wise kernel that first runs a deferred scalar step (DeferredStep below) in its prologue
template <int NRED, class F, class Pre>
__global__ void __launch_bounds__(EW_THREADS) ew_kernel_pre(int n2, const int *stop, double *partials, F f, Pre pre)
{
    if (stop && *stop) return;
    if (!pre.run()) return;
    double red[NRED > 0 ? NRED : 1];
#pragma unroll
    for (int k = 0; k < (NRED > 0 ? NRED : 1); k++) red[k] = 0;
    for (int i = blockIdx.x * EW_THREADS + threadIdx.x; i < n2; i += gridDim.x * EW_THREADS) f(i, red);
    if (NRED > 0) block_reduce_store<(NRED > 0 ? NRED : 1), EW_THREADS>(red, partials, blockIdx.x);
}

template <int NRED, class F, class Pre>
int ew_launch(b200ldu_ctx *ctx, int n2, const int *stop, double *partials, int *nPartialsOut, F f, Pre pre)
{
    int g = ew_grid(ctx, n2);
    ew_kernel_pre<NRED, F, Pre><<<g, EW_THREADS, 0, ctx->stream>>>(n2, stop, partials, f, pre);
    ctx->launches++;
    if (nPartialsOut) *nPartialsOut = g;
    KERNEL_CHECK();
    return B200LDU_OK;
}

// sums NRED interleaved partial streams in fixed order, then thread 0 runs the scalar
// logic g(sc).  One CTA.  With more than one rank the per-rank sums are combined either
//  * over peer memory (p2p.nRanks > 1): thread 0 stores its sums + a sequence flag into
//    every rank's mailbox (NVLink peer stores), waits until every rank's flag for this
//    sequence number has arrived in its own mailbox and adds the contributions in rank
//    order -- one kernel, no NCCL call, bit-identical result on every rank; or
//  * through ncclAllReduce between a sum-only and a logic-only launch (scalar_step_on).
template <int NRED, bool RUN_LOGIC, class G>
__device__ __forceinline__ void scalar_body(const double *partials, int nPartials, SolverScalars *sc,
                                            G &g, const P2PRed &p2p)
{
    if (NRED > 0) {
        // thread t adds partials t, t+256, ... in that order (fixed => deterministic); the loads of eight of them are
        // issued together, all NRED sums of a partial at once: the step is serial time for the whole GPU
        double red[NRED > 0 ? NRED : 1];
#pragma unroll
        for (int k = 0; k < (NRED > 0 ? NRED : 1); k++) red[k] = 0;
        for (int base = threadIdx.x; base < nPartials; base += 256 * 8) {
            double v[8][NRED > 0 ? NRED : 1];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = base + u * 256;
                if (i < nPartials)
#pragma unroll
                    for (int k = 0; k < NRED; k++) v[u][k] = __ldcg(partials + (size_t)i * NRED + k);
            }
#pragma unroll
            for (int u = 0; u < 8; u++)
                if (base + u * 256 < nPartials)
#pragma unroll
                    for (int k = 0; k < NRED; k++) red[k] += v[u][k];
        }
        __shared__ double tot[NRED > 0 ? NRED : 1];
        block_reduce_store<(NRED > 0 ? NRED : 1), 256>(red, tot, 0);
        __syncthreads();
        if (p2p.nRanks > 1) {
            // one thread per peer: the R peer stores, flags and waits proceed concurrently, so the
            // exchange costs one NVLink round trip instead of R
            __shared__ double got[P2P_MAXR][NRED > 0 ? NRED : 1];
            const unsigned long long seq = *p2p.seq + 1;
            const int par = (int)(seq & 1);
            if (threadIdx.x < p2p.nRanks) {
                const int r = threadIdx.x;
                double *dst = p2p.mail[r] + ((size_t)(par * P2P_MAXR + p2p.rank) * 8);
                for (int k = 0; k < NRED; k++) dst[k] = tot[k];
                __threadfence_system();
                unsigned long long *f = p2p.flag[r] + (par * P2P_MAXR + p2p.rank);
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(seq) : "memory");
                spin_until(p2p.flag[p2p.rank] + (par * P2P_MAXR + r), seq, p2p.seq + 7);
                const double *src = p2p.mail[p2p.rank] + ((size_t)(par * P2P_MAXR + r) * 8);
                for (int k = 0; k < NRED; k++) got[r][k] = __ldcg(src + k);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int k = 0; k < NRED; k++) {
                    double acc = 0;
                    for (int r = 0; r < p2p.nRanks; r++) acc += got[r][k]; // rank order => same bits everywhere
                    sc->sum[k] = acc;
                }
                *p2p.seq = seq;
            }
        } else if (threadIdx.x == 0) {
            for (int k = 0; k < NRED; k++) sc->sum[k] = tot[k];
        }
    }
    if (RUN_LOGIC && threadIdx.x == 0) g(sc);
}

template <int NRED, bool RUN_LOGIC, class G>
__global__ void __launch_bounds__(256) scalar_kernel(const double *partials, int nPartials,
                                                     SolverScalars *sc, G g, P2PRed p2p)
{
    if (sc->stop) return;
    scalar_body<NRED, RUN_LOGIC>(partials, nPartials, sc, g, p2p);
}

// ---------------------------------------------------------------------------
// Deferred scalar step: the step that closes sweep n (fixed-order sum of its per-band partials, cross-rank
// all-reduce over the peer mailboxes, alpha / beta / convergence logic) runs in the PROLOGUE of sweep n+1's
// kernel instead of a one-CTA launch of its own: CTA 0 does the work (the kernel boundary has made the
// partials visible) and releases a generation counter; every other CTA waits for it before it touches
// anything that depends on the scalars.  Saves two launches and two kernel boundaries per PCG iteration.
// `want` is baked into the launch (graph-replayable): position of this sweep in the chunk of iterations.
// ---------------------------------------------------------------------------
template <int NRED, class G>
struct DeferredStep {
    const double *partials;
    int nPartials;
    SolverScalars *sc;
    G g;
    P2PRed p2p;
    unsigned want, mod; // wait until stepGen % mod == want
    int active;         // 0: nothing to close (first sweep of a solve)
    __device__ __forceinline__ bool run() const
    {
        if (!active) return *reinterpret_cast<volatile const int *>(&sc->stop) == 0;
        __shared__ unsigned long long gateSeen;
        if (blockIdx.x == 0) {
            if (*reinterpret_cast<volatile const int *>(&sc->stop)) {
                // stopped before this sweep (e.g. converged at the initial residual): tell the waiting CTAs
                if (threadIdx.x == 0) {
                    const unsigned long long v = sc->gate | (1ull << 63);
                    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&sc->gate), "l"(v) : "memory");
                }
                return false;
            }
            G gg = g;
            scalar_body<NRED, true>(partials, nPartials, sc, gg, p2p);
            __syncthreads();
            if (threadIdx.x == 0) {
                const unsigned long long next = ((sc->gate & 0xffffffffull) + 1) | ((unsigned long long)(sc->stop != 0) << 63);
                asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&sc->gate), "l"(next) : "memory");
                gateSeen = next;
            }
        } else {
            if (threadIdx.x == 0) {
                unsigned long long v, t0 = 0;
                unsigned spins = 0;
                for (;;) {
                    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(&sc->gate) : "memory");
                    // this sweep's step is done, or an earlier sweep already stopped the solve (then nothing advances)
                    if ((unsigned)(v & 0xffffffffull) % mod == want || (v >> 63)) break;
                    __nanosleep(64);
                    if ((++spins & 0x3ffu) == 0) { // bounded: CTA 0 is dispatched first in practice, not by contract
                        unsigned long long now;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                        if (!t0) t0 = now;
                        if (now - t0 > 20000000000ull) {
                            if (p2p.seq) atomicExch(p2p.seq + 7, 2ull);
                            v |= 1ull << 63;
                            break;
                        }
                    }
                }
                gateSeen = v;
            }
        }
        __syncthreads();
        return (gateSeen >> 63) == 0;
    }
};

// ---- fused PCG kernels (PCG.C:131-205 regrouped into two matrix sweeps per iteration) ----
// K_A: applies the solution/residual update of the PREVIOUS body (psi += alpha p,
// r -= alpha w; alpha from the device scalars) while staging r, then preconditions:
// z = rD*(r - sum v*(rD*r)[c]); fused sums <z,r> and sum|r|.  r is ping-ponged (other bands
// read the old halo values), psi is updated in place (own rows only).
template <class Pre>
struct PcgAinvOp : OpBase {
    static constexpr int NVEC = 1, NRED = 2;
    static constexpr bool LOCAL = true;
    const double *rOld, *w, *p, *rD;
    double *rNew, *psi, *z;
    const SolverScalars *sc;
    Pre pre; // scalar step of the previous sweep, run in this kernel's prologue
    __device__ __forceinline__ bool prologue(const LayoutDev &) const { return pre.run(); }
    __device__ __forceinline__ void stage(int g, double &a, double &) const
    {
        double r = sc->bodies > 0 ? fma(-sc->alpha, w[g], rOld[g]) : rOld[g];
        a = __dmul_rn(rD[g], r);
    }
    __device__ __forceinline__ double pack_val(int) const { return 0.0; }
    __device__ __forceinline__ void stage_own(int row, double2 &a, double2 &) const
    {
        double2 r = *reinterpret_cast<const double2 *>(rOld + row);
        if (sc->bodies > 0) {
            const double alpha = sc->alpha;
            double2 ww = *reinterpret_cast<const double2 *>(w + row);
            double2 pp = *reinterpret_cast<const double2 *>(p + row);
            double2 x = *reinterpret_cast<const double2 *>(psi + row);
            r.x = fma(-alpha, ww.x, r.x);
            r.y = fma(-alpha, ww.y, r.y);
            x.x = fma(alpha, pp.x, x.x);
            x.y = fma(alpha, pp.y, x.y);
            *reinterpret_cast<double2 *>(psi + row) = x;
        }
        *reinterpret_cast<double2 *>(rNew + row) = r;
        double2 dd = *reinterpret_cast<const double2 *>(rD + row);
        a = make_double2(__dmul_rn(dd.x, r.x), __dmul_rn(dd.y, r.y));
    }
    __device__ __forceinline__ double init(int, double, double) const { return 0.0; }
    __device__ __forceinline__ double term(double acc, double v, double t, double) const
    {
        return __dadd_rn(acc, __dmul_rn(v, t));
    }
    __device__ __forceinline__ void finish(int row, double acc0, double acc1, double, double, double, double,
                                           double *red) const
    {
        // rNew[row] was written by this CTA in phase 1 (visible after the barrier)
        double2 d = *reinterpret_cast<const double2 *>(rD + row);
        double2 r = *reinterpret_cast<const double2 *>(rNew + row);
        double z0 = __dmul_rn(d.x, __dsub_rn(r.x, acc0));
        double z1 = __dmul_rn(d.y, __dsub_rn(r.y, acc1));
        *reinterpret_cast<double2 *>(z + row) = make_double2(z0, z1);
        red[0] += z0 * r.x + z1 * r.y;
        red[1] += fabs(r.x) + fabs(r.y);
    }
};

// K_B: forms the new search direction while staging (p = z on the first body, else
// z + beta p; beta from the device scalars; p ping-ponged), w = A p, fused <w,p>.  The halo
// send (fused pack CTAs) evaluates the same expression at the patch face cells.
template <class Pre>
struct PcgAmulOp : OpBase {
    static constexpr int NVEC = 1, NRED = 1;
    static constexpr bool LOCAL = false;
    const double *z, *pOld, *diag;
    double *pNew, *out;
    const SolverScalars *sc;
    Pre pre; // scalar step of the previous sweep, run in this kernel's prologue
    __device__ __forceinline__ bool prologue(const LayoutDev &) const { return pre.run(); }
    __device__ __forceinline__ double pval(int g) const
    {
        return sc->bodies == 0 ? z[g] : fma(sc->beta, pOld[g], z[g]);
    }
    __device__ __forceinline__ void stage(int g, double &a, double &) const { a = pval(g); }
    __device__ __forceinline__ double pack_val(int row) const { return pval(row); }
    __device__ __forceinline__ void stage_own(int row, double2 &a, double2 &) const
    {
        double2 zz = *reinterpret_cast<const double2 *>(z + row);
        if (sc->bodies > 0) {
            const double beta = sc->beta;
            double2 po = *reinterpret_cast<const double2 *>(pOld + row);
            zz.x = fma(beta, po.x, zz.x);
            zz.y = fma(beta, po.y, zz.y);
        }
        *reinterpret_cast<double2 *>(pNew + row) = zz;
        a = zz;
    }
    __device__ __forceinline__ double init(int r, double a, double) const { return __dmul_rn(diag[r], a); }
    __device__ __forceinline__ double term(double acc, double v, double a, double) const
    {
        return __dadd_rn(acc, __dmul_rn(v, a));
    }
    __device__ __forceinline__ void finish(int r, double acc0, double acc1, double a0, double, double a1,
                                           double, double *red) const
    {
        *reinterpret_cast<double2 *>(out + r) = make_double2(acc0, acc1);
        red[0] += acc0 * a0 + acc1 * a1;
    }
};
