// pcg_fused2.cuh -- PCG iteration as TWO launches: the scalar step of each sweep lives in the prologue
// of the NEXT sweep's kernel.
//
// Reference loop: LDU/solvers/PCG/PCG.C:131-205 -- 8 Thrust launches + 3 host synchronisations (+3
// MPI_Allreduce when decomposed) per iteration.  Round 1 of this repo: 2 matrix sweeps + 2 one-CTA scalar
// kernels (which also carried the cross-rank all-reduce).  Here the scalar kernels are gone:
//   * every sweep kernel runs a fixed grid (one CTA per SM slot), each CTA walks its bands in a fixed
//     order and leaves ONE partial per fused sum;
//   * the prologue of the next sweep's kernel -- every CTA, redundantly -- adds those G partials in fixed
//     order, (several ranks) exchanges the rank's sums through the peer mailboxes over NVLink (CTA 0
//     publishes, every CTA waits for the R flags and adds in rank order), then evaluates alpha / beta and
//     the convergence test (PCG.C:170,190-205).  Identical sums => identical decisions in every CTA of
//     every rank; the solver state travels between launches in a double-buffered PcgCarry block.
// => 2 launches and 2 kernel boundaries per iteration (was 4 and 4), no NCCL call, no host in the loop
// (the host replays chunks of iterations as a CUDA graph and polls the stop flag, solvers.cu).
// Row arithmetic is the engine's (bit-comparable with the oracle); the global sums have a fixed order
// (bands of a CTA in order, CTAs by index, ranks by index): run-to-run deterministic.
#pragma once
#include "solver_steps.cuh"

// loads of vectors the previous launch wrote: plain loads are fine across a kernel boundary; L2-only keeps the
// streaming vectors out of L1
__device__ __forceinline__ double2 ld_cg2(const double *p) { return __ldcg(reinterpret_cast<const double2 *>(p)); }

// solver state handed from launch to launch (global memory, two copies: a launch reads one, CTA 0 writes the other)
struct PcgCarry {
    double wArA, wArAold, wApA, alpha, beta, finalResidual;
    unsigned long long rseq; // all-reduce sequence number (peer mailboxes)
    int nIterations, bodies, converged, singular, stop, pad;
};

// sweep A operator (body k): applies psi += alpha p, r -= alpha w of body k-1 while staging r, then
// z = rD (r - sum v (rD r)[c]); sums <z,r>, sum|r|
struct PAinvOp : OpBase {
    static constexpr int NVEC = 1, NRED = 2;
    static constexpr bool LOCAL = true;
    const double *rOld, *w, *p, *rD;
    double *rNew, *psi, *z;
    double alpha;
    int bodies;
    __device__ __forceinline__ void stage(int g, double &a, double &) const
    {
        double r = bodies > 0 ? fma(-alpha, w[g], rOld[g]) : rOld[g];
        a = __dmul_rn(rD[g], r);
    }
    __device__ __forceinline__ double pack_val(int) const { return 0.0; }
    __device__ __forceinline__ void stage_own(int row, double2 &a, double2 &) const
    {
        double2 r = ld_cg2(rOld + row);
        if (bodies > 0) {
            double2 ww = ld_cg2(w + row), pp = ld_cg2(p + row), x = ld_cg2(psi + row);
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
        double2 r = ld_cg2(rNew + row);
        double z0 = __dmul_rn(d.x, __dsub_rn(r.x, acc0));
        double z1 = __dmul_rn(d.y, __dsub_rn(r.y, acc1));
        *reinterpret_cast<double2 *>(z + row) = make_double2(z0, z1);
        red[0] += z0 * r.x + z1 * r.y;
        red[1] += fabs(r.x) + fabs(r.y);
    }
};

// sweep B operator: p = z (first body) | z + beta p while staging (p ping-ponged), w = A p, sum <w,p>;
// the fused halo send evaluates the same expression at the patch face cells
struct PAmulOp : OpBase {
    static constexpr int NVEC = 1, NRED = 1;
    static constexpr bool LOCAL = false;
    const double *z, *pOld, *diag;
    double *pNew, *out;
    double beta;
    int bodies;
    __device__ __forceinline__ double pval(int g) const { return bodies == 0 ? z[g] : fma(beta, pOld[g], z[g]); }
    __device__ __forceinline__ void stage(int g, double &a, double &) const { a = pval(g); }
    __device__ __forceinline__ double pack_val(int row) const { return pval(row); }
    __device__ __forceinline__ void stage_own(int row, double2 &a, double2 &) const
    {
        double2 zz = ld_cg2(z + row);
        if (bodies > 0) {
            double2 po = ld_cg2(pOld + row);
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

struct Pcg2Args {
    const PcgCarry *cin;   // state this launch starts from
    PcgCarry *cout;        // state it leaves (written by CTA 0)
    SolverScalars *sc;     // normFactor, tolerances; receives the result when the loop ends
    double *hist;
    const double *pin;     // partials of the previous sweep [gridPrev][2]
    double *pout;          // partials of this sweep [grid][2]
    int nIn;               // CTAs of the previous sweep (0: nothing to close -- first sweep of the solve)
    P2PRed p2p;
    unsigned long long *err; // ctx->d_seq + 7 (peer wait timed out) or null
    const double *rD;      // sweep A of the diagonal / unpreconditioned variants
    int pk;
};

// Prologue of a sweep: close the previous sweep.  SWEEP 0 (A of body k): alpha of body k-1 from <w,p>.
// SWEEP 1 (B of body k): residual and convergence of body k-1, beta of body k, from <z,r> and sum|r|.
// Every thread returns the same state; returns false when the solve is over.
template <int SWEEP>
__device__ __forceinline__ bool pcg2_prologue(const Pcg2Args &A, PcgCarry &c)
{
    constexpr int NRED = SWEEP == 0 ? 1 : 2; // sums left by the previous sweep
    __shared__ double tot[2];
    __shared__ double got[P2P_MAXR][2];
    __shared__ PcgCarry sh;
    const int tid = threadIdx.x;
    if (tid == 0) sh = *A.cin;
    __syncthreads();
    if (sh.stop) return false;
    if (A.nIn > 0) {
        double s[NRED];
#pragma unroll
        for (int k = 0; k < NRED; k++) {
            double t = 0;
            for (int i = tid; i < A.nIn; i += ENGINE_THREADS) t += __ldcg(A.pin + (size_t)i * 2 + k);
            s[k] = t;
        }
        block_reduce_store<NRED, ENGINE_THREADS>(s, tot, 0);
        __syncthreads();
        if (A.p2p.nRanks > 1) {
            // all-reduce over peer memory: CTA 0 stores this rank's sums + a sequence flag into every rank's
            // mailbox (its own included); every CTA waits for the R flags in the local mailbox
            const unsigned long long rseq = sh.rseq + 1;
            const int par = (int)(rseq & 1);
            if (blockIdx.x == 0 && tid < A.p2p.nRanks) {
                double *dst = A.p2p.mail[tid] + ((size_t)(par * P2P_MAXR + A.p2p.rank) * 8);
                for (int k = 0; k < NRED; k++) dst[k] = tot[k];
                __threadfence_system();
                unsigned long long *f = A.p2p.flag[tid] + (par * P2P_MAXR + A.p2p.rank);
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(rseq) : "memory");
            }
            if (tid < A.p2p.nRanks) {
                spin_until(A.p2p.flag[A.p2p.rank] + (par * P2P_MAXR + tid), rseq, A.err);
                const double *src = A.p2p.mail[A.p2p.rank] + ((size_t)(par * P2P_MAXR + tid) * 8);
                for (int k = 0; k < NRED; k++) {
                    double v;
                    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(src + k) : "memory");
                    got[tid][k] = v;
                }
            }
            __syncthreads();
            if (tid == 0) {
                for (int k = 0; k < NRED; k++) {
                    double acc = 0;
                    for (int r = 0; r < A.p2p.nRanks; r++) acc += got[r][k]; // rank order => same bits everywhere
                    tot[k] = acc;
                }
                sh.rseq = rseq;
            }
            __syncthreads();
        }
        if (tid == 0) {
            const SolverScalars *sc = A.sc;
            if (SWEEP == 0) { // scalar step B of the previous body: alpha (PCG.C:166-175)
                sh.wApA = tot[0];
                if (!(fabs(sh.wApA) / sc->normFactor > VSMALL_)) { // checkSingularity
                    sh.singular = 1;
                    sh.stop = 1;
                } else {
                    sh.alpha = sh.wArA / sh.wApA;
                    sh.bodies++;
                }
            } else { // scalar step A: closes body k-1 (residual, convergence: PCG.C:190-205), then beta of body k
                if (sh.bodies > 0) {
                    sh.finalResidual = tot[1] / sc->normFactor;
                    if (blockIdx.x == 0 && A.hist && sh.nIterations + 1 < sc->histCap) A.hist[sh.nIterations + 1] = sh.finalResidual;
                    sh.converged = (sh.finalResidual < sc->tolerance ||
                                    (sc->relTol > SMALL_ && sh.finalResidual < sc->relTol * sc->initialResidual)) ? 1 : 0;
                    const int n = sh.nIterations;
                    sh.nIterations = n + 1;
                    if (!((n < sc->maxIter && !sh.converged) || (n + 1 < sc->minIter))) sh.stop = 1;
                }
                if (!sh.stop) {
                    sh.wArAold = sh.wArA;
                    sh.wArA = tot[0];
                    sh.beta = sh.wArA / sh.wArAold;
                }
            }
        }
        __syncthreads();
    }
    c = sh;
    if (blockIdx.x == 0 && tid == 0) {
        *A.cout = sh;
        if (sh.stop) { // the loop is over: publish the result (read back by the host)
            SolverScalars *s = A.sc;
            s->wArA = sh.wArA, s->wArAold = sh.wArAold, s->wApA = sh.wApA, s->alpha = sh.alpha, s->beta = sh.beta;
            s->finalResidual = sh.finalResidual;
            s->nIterations = sh.nIterations;
            s->converged = sh.converged;
            s->singular = sh.singular;
            s->bodies = sh.bodies;
            s->stop = 1;
            if (A.p2p.nRanks > 1) *A.p2p.seq = sh.rseq;
        }
    }
    return !sh.stop;
}

// per-thread sums of one band -> acc[] in shared memory (no accumulator stays in registers across bands)
template <int NRED>
__device__ __forceinline__ void pcg2_band_sums(double *acc, double *tot, double (&red)[NRED])
{
    block_reduce_store<NRED, ENGINE_THREADS>(red, tot, 0);
    __syncthreads();
    if (threadIdx.x < NRED) acc[threadIdx.x] += tot[threadIdx.x];
    __syncthreads(); // also separates this band's tile from the next band's staging
}

// One sweep: fixed grid, CTA c takes bands c, c + G, ... (and, sweep B with a peer-memory halo, the pack jobs first)
template <class Op, int SWEEP>
__global__ void __launch_bounds__(ENGINE_THREADS, ENGINE_MINB) pcg2_kernel(const LayoutDev L, const double *__restrict__ val,
                                                                          const Op op0, const Pcg2Args A)
{
    extern __shared__ double smem[];
    __shared__ double acc[2], tot2[2];
    PcgCarry c;
    if (!pcg2_prologue<SWEEP>(A, c)) return;
    Op op = op0;
    if constexpr (SWEEP == 0)
        op.alpha = c.alpha;
    else
        op.beta = c.beta;
    op.bodies = c.bodies;
    if (threadIdx.x < 2) acc[threadIdx.x] = 0;
    const int G = gridDim.x, cta = blockIdx.x;
    HaloWait hw;
    if (SWEEP == 1 && L.nPackChunks > 0) {
        hw.seq = L.seqs[1] + 1; // advanced by the last CTA of this launch (below)
        hw.remoteTail = (hw.seq & 1) ? L.tail1 : L.tail0;
        for (int ch = cta; ch < L.nPackChunks; ch += G) {
            engine_pack_chunk(L, op, ch, hw.seq);
            __syncthreads();
        }
    }
    bool elementwise = false;
    if constexpr (SWEEP == 0) elementwise = A.pk != 2;
    if (elementwise) {
        // diagonal / no preconditioner: sweep A is element-wise
        const PAinvOp &o = reinterpret_cast<const PAinvOp &>(op);
        const int n2 = L.bandRows >> 1;
        for (int band = cta; band < L.nBands; band += G) {
            const size_t r0 = (size_t)band * L.bandRows;
            double red[2] = {0, 0};
            for (int i = threadIdx.x; i < n2; i += ENGINE_THREADS) {
                const size_t e = r0 + 2 * (size_t)i;
                double2 r = ld_cg2(o.rOld + e);
                if (o.bodies > 0) {
                    double2 ww = ld_cg2(o.w + e), pp = ld_cg2(o.p + e), x = ld_cg2(o.psi + e);
                    r.x = fma(-o.alpha, ww.x, r.x);
                    r.y = fma(-o.alpha, ww.y, r.y);
                    *reinterpret_cast<double2 *>(o.psi + e) = make_double2(fma(o.alpha, pp.x, x.x), fma(o.alpha, pp.y, x.y));
                }
                *reinterpret_cast<double2 *>(o.rNew + e) = r;
                double2 zz = r;
                if (A.pk == 1) {
                    double2 d = *reinterpret_cast<const double2 *>(o.rD + e);
                    zz = make_double2(__dmul_rn(d.x, r.x), __dmul_rn(d.y, r.y));
                }
                *reinterpret_cast<double2 *>(o.z + e) = zz;
                red[0] += zz.x * r.x + zz.y * r.y;
                red[1] += fabs(r.x) + fabs(r.y);
            }
            pcg2_band_sums<2>(acc, tot2, red);
        }
    } else {
        for (int band = cta; band < L.nBands; band += G) {
            double red[Op::NRED];
#pragma unroll
            for (int k = 0; k < Op::NRED; k++) red[k] = 0;
            engine_band(L, val, op, band, smem, red, hw);
            pcg2_band_sums<Op::NRED>(acc, tot2, red);
        }
    }
    __syncthreads();
    if (threadIdx.x < 2) A.pout[(size_t)cta * 2 + threadIdx.x] = acc[threadIdx.x];
    if (SWEEP == 1 && L.nPackChunks > 0 && threadIdx.x == 0) {
        if (atomicAdd(&L.seqs[6], 1ull) + 1 == (unsigned long long)gridDim.x) {
            L.seqs[6] = 0;
            __threadfence();
            L.seqs[1] = hw.seq;
        }
    }
}
