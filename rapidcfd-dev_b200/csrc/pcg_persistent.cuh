// pcg_persistent.cuh -- the whole PCG iteration loop as ONE persistent cooperative kernel.
//
// Reference loop: LDU/solvers/PCG/PCG.C:131-205 -- per iteration 8 Thrust launches, 3 host
// synchronisations for the reductions and (decomposed) 3 MPI_Allreduce + one halo exchange.
// Here: one CTA per SM slot stays resident for the whole solve and walks a fixed list of
// bands; an iteration is two matrix sweeps separated by two device-wide barriers, and
// everything that used to be a kernel boundary, a scalar-step launch or a collective call
// happens inside those barriers:
//   sweep A  applies psi += alpha p, r -= alpha w of the previous body while staging r, then
//            z = M^-1 r (AINV sweep, or element-wise for diagonal / none); sums <z,r>, sum|r|
//   barrier  every CTA stores ONE partial per sum (its bands accumulated in fixed order),
//            arrives (release), waits (acquire), then adds the G partials in fixed order;
//            with several ranks CTA 0 also pushes the rank's sums into every peer's mailbox
//            over NVLink and all CTAs add the R contributions in rank order
//            => every CTA of every rank holds bit-identical sums and takes the same
//            convergence decision; no scalar kernel, no NCCL call, no host
//   sweep B  p = z + beta p while staging (halo of p packed into the neighbours' receive
//            buffers by the first work items), w = A p; sum <w,p>
//   barrier  alpha
// The host launches once per solve and reads the scalars back at the end.  Row arithmetic
// is the engine's (engine.cuh), so results per row are the oracle's bit for bit; the global
// sums have a fixed, run-to-run deterministic order (per CTA, then over CTAs, then ranks).
#pragma once
#include "solver_steps.cuh"

// loads of data written earlier in the SAME kernel by other CTAs: L2 (never a stale L1 line)
__device__ __forceinline__ double2 ld_cg2(const double *p) { return __ldcg(reinterpret_cast<const double2 *>(p)); }

// sweep A operator: see PcgAinvOp history in ops.cuh (same arithmetic), scalars by value
struct PAinvOp : OpBase {
    static constexpr int NVEC = 1, NRED = 2;
    static constexpr bool LOCAL = true;
    const double *rOld, *w, *p, *rD;
    double *rNew, *psi, *z;
    double alpha;
    int bodies;
    __device__ __forceinline__ void stage(int g, double &a, double &) const
    {
        double r = bodies > 0 ? fma(-alpha, __ldcg(w + g), __ldcg(rOld + g)) : __ldcg(rOld + g);
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

// sweep B operator
struct PAmulOp : OpBase {
    static constexpr int NVEC = 1, NRED = 1;
    static constexpr bool LOCAL = false;
    const double *z, *pOld, *diag;
    double *pNew, *out;
    double beta;
    int bodies;
    __device__ __forceinline__ double pval(int g) const
    {
        return bodies == 0 ? __ldcg(z + g) : fma(beta, __ldcg(pOld + g), __ldcg(z + g));
    }
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

struct PcgArgs {
    LayoutDev L;
    const double *val, *diag, *rD;
    double *psi, *rb[2], *pb[2], *w, *z;
    SolverScalars *sc;
    double *hist;
    double *cpart;   // [2][gridDim][2] per-CTA partial sums, one slot per sweep kind
    unsigned *bar;   // [0] arrivals, [1] generation
    P2PRed p2p;      // peer mailboxes (nRanks > 1)
    unsigned long long *seqs; // ctx->d_seq (null on one rank): [0] reduction seq, [1] halo seq, [7] error word
    int pk;          // 0 none, 1 diagonal, 2 AINV
    long long maxBodies;
};

// device-wide barrier: all global writes of every CTA before it are visible to every CTA after it
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned &gen)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        gen++;
        unsigned prev;
        asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(bar) : "memory");
        if (prev == gridDim.x - 1) {
            bar[0] = 0; // nobody touches the counter again before it has seen the new generation
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 1), "r"(gen) : "memory");
        } else {
            unsigned g;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(bar + 1) : "memory");
            } while (g != gen);
        }
        __threadfence();
    }
    __syncthreads();
}

// sums of NRED per-thread values over the whole grid and all ranks, identical in every thread of every CTA
// of every rank.  Order: lanes (butterfly), warps, this CTA's bands (caller), CTAs by index, ranks by index.
template <int NRED>
__device__ __forceinline__ void grid_sums(const PcgArgs &A, int slot, double (&red)[NRED], unsigned &gen,
                                          unsigned long long &rseq, double (&out)[NRED])
{
    __shared__ double tot[2];
    __shared__ double got[P2P_MAXR][2];
    const int tid = threadIdx.x;
    block_reduce_store<NRED, ENGINE_THREADS>(red, tot, 0);
    __syncthreads();
    if (tid < NRED) A.cpart[((size_t)slot * gridDim.x + blockIdx.x) * 2 + tid] = tot[tid];
    grid_barrier(A.bar, gen);
    double s[NRED];
#pragma unroll
    for (int k = 0; k < NRED; k++) {
        double t = 0;
        for (int i = tid; i < (int)gridDim.x; i += ENGINE_THREADS)
            t += __ldcg(A.cpart + ((size_t)slot * gridDim.x + i) * 2 + k);
        s[k] = t;
    }
    block_reduce_store<NRED, ENGINE_THREADS>(s, tot, 0);
    __syncthreads();
    if (A.p2p.nRanks > 1) {
        // all-reduce over peer memory: CTA 0 stores this rank's sums + a sequence flag into every rank's
        // mailbox (its own included); every CTA waits for the R flags in the local mailbox
        rseq++;
        const int par = (int)(rseq & 1);
        if (blockIdx.x == 0 && tid < A.p2p.nRanks) {
            double *dst = A.p2p.mail[tid] + ((size_t)(par * P2P_MAXR + A.p2p.rank) * 8);
            for (int k = 0; k < NRED; k++) dst[k] = tot[k];
            __threadfence_system();
            unsigned long long *f = A.p2p.flag[tid] + (par * P2P_MAXR + A.p2p.rank);
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(rseq) : "memory");
        }
        if (tid < A.p2p.nRanks) {
            spin_until(A.p2p.flag[A.p2p.rank] + (par * P2P_MAXR + tid), rseq, A.seqs + 7);
            const double *src = A.p2p.mail[A.p2p.rank] + ((size_t)(par * P2P_MAXR + tid) * 8);
            for (int k = 0; k < NRED; k++) {
                double v;
                asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(src + k) : "memory");
                got[tid][k] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NRED; k++) {
            double acc = 0;
            for (int r = 0; r < A.p2p.nRanks; r++) acc += got[r][k]; // rank order => same bits everywhere
            out[k] = acc;
        }
    } else {
#pragma unroll
        for (int k = 0; k < NRED; k++) out[k] = tot[k];
    }
}

__global__ void __launch_bounds__(ENGINE_THREADS, ENGINE_MINB) pcg_persistent_kernel(const PcgArgs A)
{
    extern __shared__ double smem[];
    const LayoutDev &L = A.L;
    const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
    // solver state, replicated in every thread (identical arithmetic on identical sums)
    const SolverScalars *sc0 = A.sc;
    if (__ldcg(&sc0->stop)) return; // converged before the first iteration (PCG.C:120-128)
    const double normFactor = __ldcg(&sc0->normFactor), initialResidual = __ldcg(&sc0->initialResidual);
    const double tolerance = __ldcg(&sc0->tolerance), relTol = __ldcg(&sc0->relTol);
    const int maxIter = __ldcg(&sc0->maxIter), minIter = __ldcg(&sc0->minIter), histCap = __ldcg(&sc0->histCap);
    double wArA = GREAT_, wArAold = GREAT_, wApA = 0, alpha = 0, beta = 0, finalResidual = initialResidual;
    int nIterations = 0, bodies = 0, converged = 0, singular = 0;
    unsigned gen = *reinterpret_cast<volatile unsigned *>(A.bar + 1);
    unsigned long long rseq = A.seqs ? *reinterpret_cast<volatile unsigned long long *>(A.seqs + 0) : 0;
    unsigned long long hseq = A.seqs ? *reinterpret_cast<volatile unsigned long long *>(A.seqs + 1) : 0;
    const bool peerHalo = L.nPackChunks > 0 || (L.haloFlags && L.nRecv > 0);
    const int n2 = L.bandRows >> 1;

    for (long long k = 0; k < A.maxBodies; k++) {
        const double *rOld = A.rb[k & 1], *pPrev = A.pb[k & 1];
        double *rNew = A.rb[(k + 1) & 1], *pNew = A.pb[(k + 1) & 1];
        // ---------------- sweep A ----------------
        double redA[2] = {0, 0};
        if (A.pk == 2) {
            PAinvOp op;
            op.rOld = rOld, op.rNew = rNew, op.w = A.w, op.p = pPrev, op.psi = A.psi, op.z = A.z, op.rD = A.rD;
            op.alpha = alpha, op.bodies = bodies;
            for (int band = cta; band < L.nBands; band += G) {
                engine_band(L, A.val, op, band, smem, redA, HaloWait());
                __syncthreads();
            }
        } else {
            for (int band = cta; band < L.nBands; band += G) {
                const size_t r0 = (size_t)band * L.bandRows;
                for (int i = tid; i < n2; i += ENGINE_THREADS) {
                    const size_t e = r0 + 2 * (size_t)i;
                    double2 r = ld_cg2(rOld + e);
                    if (bodies > 0) {
                        double2 ww = ld_cg2(A.w + e), pp = ld_cg2(pPrev + e), x = ld_cg2(A.psi + e);
                        r.x = fma(-alpha, ww.x, r.x);
                        r.y = fma(-alpha, ww.y, r.y);
                        *reinterpret_cast<double2 *>(A.psi + e) = make_double2(fma(alpha, pp.x, x.x), fma(alpha, pp.y, x.y));
                    }
                    *reinterpret_cast<double2 *>(rNew + e) = r;
                    double2 zz = r;
                    if (A.pk == 1) {
                        double2 d = *reinterpret_cast<const double2 *>(A.rD + e);
                        zz = make_double2(__dmul_rn(d.x, r.x), __dmul_rn(d.y, r.y));
                    }
                    *reinterpret_cast<double2 *>(A.z + e) = zz;
                    redA[0] += zz.x * r.x + zz.y * r.y;
                    redA[1] += fabs(r.x) + fabs(r.y);
                }
            }
        }
        double sA[2];
        grid_sums<2>(A, 0, redA, gen, rseq, sA);
        // scalar step A: closes body k-1 (residual, convergence: PCG.C:190-205), then beta of body k
        bool stop = false;
        if (bodies > 0) {
            finalResidual = sA[1] / normFactor;
            if (cta == 0 && tid == 0 && A.hist && nIterations + 1 < histCap) A.hist[nIterations + 1] = finalResidual;
            converged = (finalResidual < tolerance || (relTol > SMALL_ && finalResidual < relTol * initialResidual)) ? 1 : 0;
            const int n = nIterations;
            nIterations = n + 1;
            stop = !((n < maxIter && !converged) || (n + 1 < minIter));
        }
        if (stop) break;
        wArAold = wArA;
        wArA = sA[0];
        beta = wArA / wArAold;

        // ---------------- sweep B ----------------
        double redB[1] = {0};
        {
            PAmulOp op;
            op.z = A.z, op.pOld = pPrev, op.pNew = pNew, op.out = A.w, op.diag = A.diag;
            op.beta = beta, op.bodies = bodies;
            HaloWait hw;
            if (peerHalo) {
                hseq++;
                hw.seq = hseq;
                hw.remoteTail = (hseq & 1) ? L.tail1 : L.tail0;
                // the halo of p goes out first: work items [0, nPackChunks) are pack jobs
                for (int c = cta; c < L.nPackChunks; c += G) {
                    engine_pack_chunk(L, op, c, hseq);
                    __syncthreads();
                }
            }
            for (int band = cta; band < L.nBands; band += G) {
                engine_band(L, A.val, op, band, smem, redB, hw);
                __syncthreads();
            }
        }
        double sB[1];
        grid_sums<1>(A, 1, redB, gen, rseq, sB);
        wApA = sB[0];
        if (!(fabs(wApA) / normFactor > VSMALL_)) { // checkSingularity, PCG.C:170
            singular = 1;
            break;
        }
        alpha = wArA / wApA;
        bodies++;
    }
    if (cta == 0 && tid == 0) {
        SolverScalars *s = A.sc;
        s->wArA = wArA, s->wArAold = wArAold, s->wApA = wApA, s->alpha = alpha, s->beta = beta;
        s->finalResidual = finalResidual;
        s->nIterations = nIterations;
        s->converged = converged;
        s->singular = singular;
        s->bodies = bodies;
        s->stop = 1;
        if (A.seqs) {
            A.seqs[0] = rseq;
            A.seqs[1] = hseq;
        }
    }
}
