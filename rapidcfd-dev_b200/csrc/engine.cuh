// engine.cuh -- the banded row-gather kernel engine (sm_100a).
//
// One CTA per band.  Phase 1 stages the band's input vector(s) into shared memory:
// the band's own rows with coalesced 128-bit loads, then the band's halo columns
// (rows of neighbouring bands or values received from coupled patches) with a sorted
// gather.  Phase 2: each warp walks its slices of 64 rows; lane k owns rows 2k, 2k+1 and
// reads slot j of both rows with one 128-bit coefficient load (double2); the columns come
// from the slice's compressed column blob (layout.cu 3b), copied asynchronously (cp.async)
// into a per-warp double buffer one slice ahead; the gathers hit shared memory only.  Row sums follow the
// reference's order exactly (diag, owner faces, neighbour faces, interface faces; products
// rounded separately: __dmul_rn/__dadd_rn) so results are bit-comparable with the oracle.
// Fused reductions (dot products needed by the Krylov solvers) are reduced with warp
// shuffles, then across the CTA through shared memory, and written as one partial per
// band -- summed in fixed order by the scalar-step kernel (deterministic, no atomics).
//
// Replaces the reference's Thrust functors: matrixMultiplyFunctor
// (LDU/lduMatrix/lduMatrixATmul.C:42-138), lduAddressingFunctor family
// (LDU/lduAddressing/lduAddressingFunctors.H:10-185), AINVPreconditionerFunctor
// (LDU/preconditioners/AINVPreconditioner/AINVPreconditionerF.H:5-100),
// JacobiSmootherFunctor (LDU/smoothers/Jacobi/JacobiSmootherF.H:8-110).
#pragma once
#include "internal.h"

__device__ __forceinline__ double2 ldg_stream2(const double *p)
{
    // coefficients and columns are read exactly once per kernel: keep them out of L1
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(r.x), "=d"(r.y)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t ldg_stream_u32(const void *p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

__device__ __forceinline__ void cp_async16(void *smemDst, const void *gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sums of NRED per-thread values -> out[blockIdx.x * NRED + k].
// Fixed combination order: lanes by xor-butterfly, warps in index order.
template <int NRED, int THREADS>
__device__ __forceinline__ void block_reduce_store(double (&red)[NRED], double *partials,
                                                   int slot)
{
    __shared__ double wsum[NRED][THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NRED; k++) {
        double v = warp_sum(red[k]);
        if (lane == 0) wsum[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < NRED) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; w++) s += wsum[threadIdx.x][w];
        partials[(size_t)slot * NRED + threadIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------
// Op concept
//   static constexpr int  NVEC   1 or 2 vectors staged in shared memory
//   static constexpr int  NRED   number of fused reductions (0..3)
//   static constexpr bool LOCAL  true: owner/neighbour entries only (no interfaces)
//   const int *stop              device flag; kernel exits at once when *stop != 0
//   double *partials             [nBands*NRED]
//   void   stage(int g, double &a, double &b)        value(s) of extended index g
//   void   stage_own(int r, double2 &a, double2 &b)  rows r, r+1 of the own band (may
//                                                    also write a fused vector update)
//   double init(int r, double a, double b)           start of the row sum
//   double term(double acc, double v, double a, double b)   acc (+) v*...
//   void   finish(int r, double acc0, double acc1, a0,b0,a1,b1, double *red)
//                                                    writes rows r, r+1; adds reductions
// ---------------------------------------------------------------------------
// what a band needs to know about a peer-memory halo exchange in flight
struct HaloWait {
    const double *remoteTail = nullptr; // receive buffer of this exchange (null: values are in the vector's tail)
    unsigned long long seq = 0;         // arrival flags must reach this value
};

__device__ __forceinline__ bool spin_until(const unsigned long long *f, unsigned long long seq, unsigned long long *err)
{
    // bounded wait on a flag written by a peer GPU (ld.acquire.sys): a peer that died or never entered the
    // matching kernel must not hang this GPU for ever -- after ~20 s the error word is set and the wait ends
    unsigned long long v, t0 = 0;
    unsigned spins = 0;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
        if (v >= seq) return true;
        if ((++spins & 0x3ffu) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (!t0) t0 = now;
            if (now - t0 > 20000000000ull) {
                if (err) atomicExch(err, 1ull);
                return false;
            }
        }
    }
}

// issue the asynchronous copy of one slice's column blob into a per-warp staging buffer
__device__ __forceinline__ void blob_issue(const LayoutDev &L, int s, char *buf, int lane)
{
    const int c0 = L.cStart[s], n = L.cStart[s + 1] - c0;
    const uint4 *g = L.cblob + c0;
    for (int i = lane; i < n; i += 32) cp_async16(buf + 16 * i, g + i);
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// One band: phase 1 stages the vector tile(s), phase 2 forms the row sums of the band's slices.
// `red` accumulates the Op's fused reductions (per thread).  The caller provides the CTA barrier
// between successive bands of a persistent kernel.
template <class Op>
__device__ __forceinline__ void engine_band(const LayoutDev &L, const double *__restrict__ val, const Op &op,
                                            const int band, double *smem, double *red, const HaloWait hw)
{
    const int rowBase = band * L.bandRows;
    const int stride = L.tileLen; // even: keeps the second tile 16-byte aligned
    double *xs = smem;
    double *ys = smem + ((Op::NVEC > 1) ? stride : 0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = ENGINE_THREADS / 32;
    char *wb0 = reinterpret_cast<char *>(smem + (size_t)stride * Op::NVEC) + (size_t)warp * 2 * L.wbufBytes;

    // the columns of this warp's first slice travel while the tile is being staged
    if (warp < L.slicesPerBand) blob_issue(L, band * L.slicesPerBand + warp, wb0, lane);

    // ---- phase 1: stage the band's vector tile + halo ----
    if (Op::NVEC > 0) {
        const uint32_t *rowPos = L.rowPos + (rowBase >> 1);
        for (int i = tid; i < (L.bandRows >> 1); i += ENGINE_THREADS) {
            double2 a, b = make_double2(0, 0);
            op.stage_own(rowBase + 2 * i, a, b);
            const uint32_t pp = __ldg(rowPos + i); // tile positions of rows 2i, 2i+1
            xs[pp & 0xffffu] = a.x;
            xs[pp >> 16] = a.y;
            if (Op::NVEC > 1) {
                ys[pp & 0xffffu] = b.x;
                ys[pp >> 16] = b.y;
            }
        }
        const int hs = L.haloStart[band], hn = L.haloStart[band + 1] - hs;
        // peer-memory halo: bands that reference received values wait for the neighbours' arrival
        // flags here, so interior bands overlap with the exchange
        if (hw.remoteTail && hn > 0 && __ldg(L.haloIdx + hs + hn - 1) >= L.nPad) {
            if (tid < L.nNbr) spin_until(L.haloFlags + L.nbr[tid], hw.seq, L.seqs ? L.seqs + 7 : nullptr);
            __syncthreads();
        }
        for (int i = tid; i < hn; i += ENGINE_THREADS) {
            int g = __ldg(L.haloIdx + hs + i);
            double a, b = 0;
            if (hw.remoteTail && g >= L.nPad)
                a = __ldcg(hw.remoteTail + (g - L.nPad));
            else
                op.stage(g, a, b);
            const int hp = __ldg(L.haloPos + hs + i);
            xs[hp] = a;
            if (Op::NVEC > 1) ys[hp] = b;
        }
        __syncthreads();
    }

    // ---- phase 2: row sums ----
    int cur = 0;
    for (int sl = warp; sl < L.slicesPerBand; sl += NW) {
        const int s = band * L.slicesPerBand + sl;
        char *buf = wb0 + (size_t)cur * L.wbufBytes;
        if (sl + NW < L.slicesPerBand) {
            blob_issue(L, s + NW, wb0 + (size_t)(cur ^ 1) * L.wbufBytes, lane);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        const long long base = L.sliceStart[s];
        const int Wall = L.sliceW[s];
        const int W = Op::LOCAL ? L.sliceWL[s] : Wall;
        const uint4 *meta = reinterpret_cast<const uint4 *>(buf);
        const uint32_t posw = reinterpret_cast<const uint32_t *>(buf + 16 * Wall)[lane];
        const int p0 = posw & 0xffffu, p1 = posw >> 16; // tile positions of this lane's two rows
        const uint16_t *exc = reinterpret_cast<const uint16_t *>(buf + 16 * Wall + 2 * SLICE_ROWS);
        const int lr = sl * SLICE_ROWS + 2 * lane; // local row of this lane's first row
        double acc0, acc1;
        acc0 = op.init(rowBase + lr, (Op::NVEC > 0) ? xs[p0] : 0, (Op::NVEC > 1) ? ys[p0] : 0);
        acc1 = op.init(rowBase + lr + 1, (Op::NVEC > 0) ? xs[p1] : 0, (Op::NVEC > 1) ? ys[p1] : 0);
        const double *vp = val + base + 2 * lane;
        const unsigned sh = (2u * lane) & 31u;
        const unsigned lowMask = (1u << sh) - 1u;
        // columns of slot j for this lane's two rows: regular rows sit at (the row's tile position + delta), the others
        // read their 16-bit column from the slice's exception list
#define SLOT_COLS(J, C0, C1)                                                              \
    {                                                                                     \
        const uint4 md_ = meta[J];                                                        \
        const unsigned word_ = lane < 16 ? md_.x : md_.y;                                 \
        C0 = p0 + (int)md_.z;                                                             \
        C1 = p1 + (int)md_.z;                                                             \
        if (md_.x | md_.y) {                                                              \
            const unsigned two_ = (word_ >> sh) & 3u;                                     \
            if (two_) {                                                                   \
                unsigned below_ = __popc(word_ & lowMask) + (lane < 16 ? 0u : __popc(md_.x)); \
                const uint16_t *e_ = exc + md_.w + below_;                                \
                if (two_ & 1u) C0 = *e_++;                                                \
                if (two_ & 2u) C1 = *e_;                                                  \
            }                                                                             \
        }                                                                                 \
    }
        // slots in groups of ENGINE_GROUP: all coefficient loads of a group are issued before any use (the
        // kernel lives on bytes in flight: one 512-byte row of loads per slot and warp)
#if ENGINE_COLMODE == 1
        const uint16_t *cp = L.col + base + 2 * lane;
#endif
        for (int j = 0; j < W; j += ENGINE_GROUP) {
            double2 v[ENGINE_GROUP];
#if ENGINE_COLMODE == 1
            uint32_t cc[ENGINE_GROUP];
#endif
#pragma unroll
            for (int k = 0; k < ENGINE_GROUP; k++)
                if (j + k < W) {
                    v[k] = ldg_stream2(vp + (size_t)(j + k) * SLICE_ROWS);
#if ENGINE_COLMODE == 1
                    cc[k] = ldg_stream_u32(cp + (size_t)(j + k) * SLICE_ROWS);
#endif
                }
#pragma unroll
            for (int k = 0; k < ENGINE_GROUP; k++)
                if (j + k < W) {
                    int c0, c1;
#if ENGINE_COLMODE == 1
                    c0 = cc[k] & 0xffffu, c1 = cc[k] >> 16;
#else
                    SLOT_COLS(j + k, c0, c1)
#endif
                    acc0 = op.term(acc0, v[k].x, (Op::NVEC > 0) ? xs[c0] : 0, (Op::NVEC > 1) ? ys[c0] : 0);
                    acc1 = op.term(acc1, v[k].y, (Op::NVEC > 0) ? xs[c1] : 0, (Op::NVEC > 1) ? ys[c1] : 0);
                }
        }
#undef SLOT_COLS
        {
            // the lane's own tile values again (not kept in registers across the slot loop)
            const volatile double *xv = xs, *yv = ys;
            const double a0 = (Op::NVEC > 0) ? xv[p0] : 0, a1 = (Op::NVEC > 0) ? xv[p1] : 0;
            const double b0 = (Op::NVEC > 1) ? yv[p0] : 0, b1 = (Op::NVEC > 1) ? yv[p1] : 0;
            op.finish(rowBase + lr, acc0, acc1, a0, b0, a1, b1, red);
        }
        __syncwarp(); // the buffer is overwritten by the copy issued in the next iteration
        cur ^= 1;
    }
}

// fused halo send (peer-memory path): gather the staged vector at the processor-patch face cells of one
// chunk, store it straight into the neighbour's receive buffer over NVLink and, when the patch is
// complete, release the neighbour's arrival flag
template <class Op>
__device__ __forceinline__ void engine_pack_chunk(const LayoutDev &L, const Op &op, int chunk, unsigned long long seq)
{
    const PackChunk c = L.packChunks[chunk];
    const PackPatch P = L.packPatches[c.patch];
    double *dst = P.dst[seq & 1];
    for (int i = c.begin + threadIdx.x; i < c.end; i += ENGINE_THREADS)
        dst[i - P.start] = op.pack_val(__ldg(L.sendRows + i));
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long done = atomicAdd(&L.seqs[8 + c.patch], 1ull) + 1;
        if (done == (unsigned long long)P.nChunks) {
            L.seqs[8 + c.patch] = 0;
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(P.flag), "l"(seq) : "memory");
        }
    }
}

template <class Op>
__global__ void __launch_bounds__(ENGINE_THREADS, ENGINE_MINB) engine_kernel(const LayoutDev L, const double *__restrict__ val,
                                                                  Op op)
{
    extern __shared__ double smem[];
    if (op.stop && *op.stop) return;
    // Fused halo send (peer-memory path): the first nPackChunks CTAs of the grid are pack jobs; all other
    // CTAs are SpMV bands (those that reference received values wait for the neighbours' flags).
    const bool fusedPack = op.waitHalo && L.nPackChunks > 0;
    HaloWait hw;
    if (fusedPack) {
        hw.seq = L.seqs[1] + 1; // seqs[1] only advances when the whole grid has finished
        if ((int)blockIdx.x < L.nPackChunks) {
            engine_pack_chunk(L, op, blockIdx.x, hw.seq);
            if (threadIdx.x == 0 && atomicAdd(&L.seqs[6], 1ull) + 1 == (unsigned long long)gridDim.x) {
                L.seqs[6] = 0;
                __threadfence();
                L.seqs[1] = hw.seq;
            }
            return;
        }
    } else if (op.waitHalo && L.haloFlags) {
        hw.seq = *L.haloSeq;
    }
    if (op.waitHalo && L.haloFlags) hw.remoteTail = (hw.seq & 1) ? L.tail1 : L.tail0;
    const int band = blockIdx.x - (fusedPack ? L.nPackChunks : 0);

    double red[Op::NRED > 0 ? Op::NRED : 1];
#pragma unroll
    for (int k = 0; k < (Op::NRED > 0 ? Op::NRED : 1); k++) red[k] = 0;
    engine_band(L, val, op, band, smem, red, hw);
    if (Op::NRED > 0) block_reduce_store<(Op::NRED > 0 ? Op::NRED : 1), ENGINE_THREADS>(red, op.partials, band);
    if (fusedPack && threadIdx.x == 0) {
        if (atomicAdd(&L.seqs[6], 1ull) + 1 == (unsigned long long)gridDim.x) {
            L.seqs[6] = 0;
            __threadfence();
            L.seqs[1] = hw.seq;
        }
    }
}

inline size_t engine_smem_bytes(const LayoutDev &L, int nvec)
{
    return sizeof(double) * (size_t)L.tileLen * (size_t)nvec +
           (size_t)(ENGINE_THREADS / 32) * 2 * (size_t)L.wbufBytes;
}

template <class Op>
int engine_launch(b200ldu_addr *a, const double *val, const Op &op)
{
    const LayoutDev &L = a->L;
    const size_t smem = engine_smem_bytes(L, Op::NVEC);
    // the kernel also owns a little static shared memory (reduction scratch): opt in to large dynamic
    // shared memory well before the 48 KB default limit.  The attribute is per device and instantiation.
    static size_t configured[64] = {0};
    const int dev = a->ctx->device & 63;
    if (smem > 40 * 1024 && smem > configured[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(engine_kernel<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = smem;
    }
    int grid = L.nBands + ((op.waitHalo && L.nPackChunks > 0) ? L.nPackChunks : 0);
    engine_kernel<Op><<<grid, ENGINE_THREADS, smem, a->ctx->stream>>>(L, val, op);
    a->ctx->launches++;
    KERNEL_CHECK();
    return B200LDU_OK;
}

// matrix sweep with A's coefficients, or A^T's
template <class Op>
int engine_launch_m(b200ldu_matrix *m, bool transpose, const Op &op)
{
    return engine_launch(m->a, transpose ? m->d_valT : m->d_val, op);
}
