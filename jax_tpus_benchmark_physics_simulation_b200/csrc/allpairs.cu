// allpairs.cu — dense O(N^2) path: the reference's own formulation (MD:50-75) on B200.
//
// One PERSISTENT cooperative kernel runs a whole equilibrate_fn / production_fn call
// (MD:77-106) without returning to the host:
//
//   per step   [b] partial forces  : the flattened (i-block x j) work is cut into gridDim.x
//                                    equal contiguous ranges (stream-K style), so every SM
//                                    sub-partition gets the same number of pair evaluations
//                                    whatever N is; j positions are staged through shared
//                                    memory, i positions and accumulators live in registers.
//              grid barrier
//              [d] reduce+integrate: the owner thread of particle g sums the partials of its
//                                    i-block in a fixed order (deterministic, atomic-free),
//                                    finishes the velocity-Verlet step, writes the sample /
//                                    energies, and drifts the particle into the other
//                                    position buffer (ping-pong) for the next step.
//              grid barrier
//
// F(R_new) of step n is carried to step n+1 (the reference recomputes it, bit-identically:
// SURVEY.md §0), so there is one O(N^2) evaluation per step.  Each particle's state is read
// and written once per step.
#include "ljmd_device.cuh"

#include <algorithm>
#include <cstdio>

namespace ljmd {

namespace {

constexpr int AP_THREADS = 128;   // 4 warps: one per SM sub-partition
#ifndef AP_UNROLL
#define AP_UNROLL 8
#endif
#ifndef AP_MINBLOCKS
#define AP_MINBLOCKS 4
#endif
constexpr int kApUnroll = AP_UNROLL;   // j particles per unrolled inner-loop body
constexpr int J_UNIT     = 8;     // granularity of the j split (particles)
constexpr int TILE_J     = 512;   // j particles staged per shared-memory tile (8 KB)
constexpr int RED_LANES  = 8;     // lanes that cooperate on one particle's partial sum
constexpr float SENT_J   = 1.0e18f;    // padding particles: far away, contribute exactly 0
constexpr float SENT_I   = -1.0e18f;

struct ApArgs {
    PairConsts pc;
    int   N, G, NJu, nI, maxseg, ppc;   // ppc = particles owned per CTA in phase [d]
    int   i_lo, Nloc;                   // this rank owns particles [i_lo, i_lo + Nloc) (atom decomposition)
    int   rank, P;                      // P > 1: positions are pushed to every peer over NVLink
    unsigned xepoch0;                   // cross-GPU barrier epochs already consumed by earlier launches
    float2*   peerR0[LJMD_MAX_RANKS];   // every rank's ping-pong position buffers (peer-mapped)
    float2*   peerR1[LJMD_MAX_RANKS];
    unsigned* peer_flags[LJMD_MAX_RANKS];   // every rank's arrival words [P]; [rank] is the local one
    float dt;
    const long long* cta_start;   // [G+1] flat (i-block * NJu + j-unit) range owned by each CTA
    const int*       cta_ib0;     // [G]   first i-block a CTA touches
    const int2*      iblk_ctas;   // [nI]  first / last CTA contributing to an i-block
    const float2*    R_in;        // positions of the state the call starts from
    float2*          Rbuf0;       // ping-pong position buffers
    float2*          Rbuf1;
    float2*          Vh;          // velocities (half-step between kernels' phases)
    float2*          Ftmp;        // forces held across the thermostat barrier
    float2*          part;        // [G*maxseg*BLOCK_I] partial forces
    // Newton's-third-law tile mode (IPT == 3): upper-triangular patches of Pt x Pt tiles of 64 x 64
    int              Pt, q, npr;  // patch edge in tiles, tiles per warp edge (Pt = 2q), patch rows
    const int*       cta_pstart;  // (unused: patches are handed out dynamically)
    int              npatch, npe; // number of patches; number of energy partials per parity
    int*             sched;       // [2] patch counters (by step parity)
    const int2*      patches;     // (pa, pb) patch coordinates, pa <= pb
    float2*          rowpart;     // [npatch][Pt*64] partial force on the patch's row (i) side
    float2*          colpart;     // [npatch][Pt*64] partial force on the patch's column (j) side
    float*           pe_part;     // [2*G] per-CTA partial potential energy (by step parity)
    float*           ke_part;     // [2*G]
    unsigned*        bar;
    int*             err;
    long long*       prof;        // optional [G][4] phase clocks (debug: LJMD_AP_PROF=1)
    long long        s_begin, s_end;   // steps [s_begin, s_end); s = -1 is the prologue force
    RunCtl           rc;
    float2*          R_out;
    float2*          V_out;
    float2*          F_out;
    float*           pe_out;
};

// ---- phase [b]: one segment = (i-block ib) x (j units [ju0, ju0+julen)) ---------------------
// Shared-memory tile layout: one float4 per j particle = (-xj, -xj, -yj, -yj): a single
// broadcast LDS.128 feeds the packed (two i per thread) pair evaluation.
template <int IPT, bool CUTOFF, bool PE>
__device__ __forceinline__ void ap_segment(const ApArgs& a, const float2* __restrict__ Rcur,
                                           int ib, int ju0, int julen, float4* sj,
                                           float (&fx)[IPT], float (&fy)[IPT], float& pe_acc) {
    constexpr int BI = AP_THREADS * IPT;
    const int tid = threadIdx.x;
    const PairConsts pc = a.pc;
    const PairConsts2 pc2 = make_pair_consts2(pc);
    float xi[IPT], yi[IPT];
    int   ii[IPT];
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        ii[k] = a.i_lo + ib * BI + k * AP_THREADS + tid;
        if (ii[k] < a.i_lo + a.Nloc) {
            const float* rp = reinterpret_cast<const float*>(Rcur + ii[k]);
            xi[k] = ldcg_f32(rp); yi[k] = ldcg_f32(rp + 1);
        } else { xi[k] = SENT_I; yi[k] = SENT_I; }
        fx[k] = 0.0f; fy[k] = 0.0f;
    }
    const int j0 = ju0 * J_UNIT, jend = (ju0 + julen) * J_UNIT;
    const int i_lo = a.i_lo + ib * BI, i_hi = i_lo + BI;

    for (int jt = j0; jt < jend; jt += TILE_J) {
        const int cnt = min(TILE_J, jend - jt);           // multiple of J_UNIT
        __syncthreads();                                   // previous tile fully consumed
        for (int q = tid; q < cnt; q += AP_THREADS) {
            const int j = jt + q;
            float2 r = (j < a.N) ? __ldcg(&Rcur[j]) : make_float2(SENT_J, SENT_J);
            sj[q] = make_float4(-r.x, -r.x, -r.y, -r.y);
        }
        __syncthreads();
        // the only j that can equal one of this CTA's i lie in [i_lo, i_hi): split the tile into
        // (before | diagonal | after) so the index test is paid on the overlap only
        const int qd0 = min(max(i_lo - jt, 0), cnt), qd1 = min(max(i_hi - jt, 0), cnt);
        if constexpr (IPT == 2) {
            const float2 xi2 = make_float2(xi[0], xi[1]), yi2 = make_float2(yi[0], yi[1]);
            float2 tfx = make_float2(0.0f, 0.0f), tfy = tfx, tpe = tfx;   // per-tile sums
#pragma unroll 1
            for (int part = 0; part < 3; ++part) {
                const int qa = (part == 0) ? 0 : (part == 1 ? qd0 : qd1);
                const int qb = (part == 0) ? qd0 : (part == 1 ? qd1 : cnt);
                if (part != 1) {
#pragma unroll kApUnroll
                    for (int q = qa; q < qb; ++q) {
                        const float4 v = sj[q];            // one j particle, warp broadcast
                        pair2_accum<CUTOFF, PE, false>(xi2, yi2, make_float2(v.x, v.y), make_float2(v.z, v.w),
                                                       true, true, pc, pc2, tfx, tfy, tpe);
                    }
                } else {
#pragma unroll kApUnroll
                    for (int q = qa; q < qb; ++q) {
                        const float4 v = sj[q];
                        const int j = jt + q;
                        pair2_accum<CUTOFF, PE, true>(xi2, yi2, make_float2(v.x, v.y), make_float2(v.z, v.w),
                                                      j != ii[0], j != ii[1], pc, pc2, tfx, tfy, tpe);
                    }
                }
            }
            fx[0] += tfx.x; fx[1] += tfx.y; fy[0] += tfy.x; fy[1] += tfy.y;
            if (PE) pe_acc += tpe.x + tpe.y;
        } else {
            float tfx = 0.0f, tfy = 0.0f, tpe = 0.0f;
#pragma unroll 1
            for (int part = 0; part < 3; ++part) {
                const int qa = (part == 0) ? 0 : (part == 1 ? qd0 : qd1);
                const int qb = (part == 0) ? qd0 : (part == 1 ? qd1 : cnt);
                if (part != 1) {
#pragma unroll kApUnroll
                    for (int q = qa; q < qb; ++q) {
                        const float4 v = sj[q];
                        pair_accum<CUTOFF, PE, false>(xi[0], yi[0], -v.x, -v.z, true, pc, tfx, tfy, tpe);
                    }
                } else {
#pragma unroll kApUnroll
                    for (int q = qa; q < qb; ++q) {
                        const float4 v = sj[q];
                        pair_accum<CUTOFF, PE, true>(xi[0], yi[0], -v.x, -v.z, (jt + q) != ii[0], pc, tfx, tfy, tpe);
                    }
                }
            }
            fx[0] += tfx; fy[0] += tfy;
            if (PE) pe_acc += tpe;
        }
    }
}

// ---- Newton's-third-law tiles -------------------------------------------------------------------
// A tile is 64 i-particles (two per lane, packed) x 32 j-particles (one per lane).  The j particle
// and the force accumulated ON it travel around the warp by shuffle: after 32 steps every j has met
// every lane's two i and is back in its home lane, so each unordered pair is evaluated once and
// applied to both sides with no atomics and a fixed summation order.
constexpr int T3_BLK = 64;

__device__ __forceinline__ int tri_base(int pa, int npr) { return pa * npr - (pa * (pa - 1)) / 2; }

template <bool CUTOFF, bool PE, bool N3L>
__device__ __forceinline__ void tile_half(const PairConsts& pc, const PairConsts2& pc2, float2 xi2,
                                          float2 yi2, float xj, float yj, bool self0, bool self1,
                                          float2& fx2, float2& fy2, float2& pe2, float& fjx, float& fjy) {
    const int src = (threadIdx.x + 1) & 31;
    // the reaction on the travelling j is accumulated as ONE scalar per component (the two i of this
    // lane are added in a fixed order), so only two values per component... per step two shuffles
    // move (xj, yj) and two move the accumulators
    float ajx = 0.0f, ajy = 0.0f;
    // step 0 peeled: the only step at which j can be one of this lane's own i (diagonal tiles)
    {
        float2 f, dx, dy;
        if (N3L) pair2_eval<CUTOFF, PE, false>(xi2, yi2, xj, yj, true, true, pc, pc2, f, dx, dy, pe2);
        else     pair2_eval<CUTOFF, PE, true >(xi2, yi2, xj, yj, !self0, !self1, pc, pc2, f, dx, dy, pe2);
        fx2 = __ffma2_rn(f, dx, fx2);
        fy2 = __ffma2_rn(f, dy, fy2);
        if (N3L) {
            ajx = fmaf(f.y, dx.y, __fmul_rn(f.x, dx.x));
            ajy = fmaf(f.y, dy.y, __fmul_rn(f.x, dy.x));
        }
        xj = __shfl_sync(0xffffffffu, xj, src);
        yj = __shfl_sync(0xffffffffu, yj, src);
        if (N3L) { ajx = __shfl_sync(0xffffffffu, ajx, src); ajy = __shfl_sync(0xffffffffu, ajy, src); }
    }
#pragma unroll 4
    for (int s = 1; s < 32; ++s) {
        float2 f, dx, dy;
        pair2_eval<CUTOFF, PE, false>(xi2, yi2, xj, yj, true, true, pc, pc2, f, dx, dy, pe2);
        fx2 = __ffma2_rn(f, dx, fx2);
        fy2 = __ffma2_rn(f, dy, fy2);
        if (N3L) {
            ajx = fmaf(f.y, dx.y, fmaf(f.x, dx.x, ajx));
            ajy = fmaf(f.y, dy.y, fmaf(f.x, dy.x, ajy));
        }
        xj = __shfl_sync(0xffffffffu, xj, src);
        yj = __shfl_sync(0xffffffffu, yj, src);
        if (N3L) { ajx = __shfl_sync(0xffffffffu, ajx, src); ajy = __shfl_sync(0xffffffffu, ajy, src); }
    }
    // 32 rotations: the travelling j (and its accumulator) is home again
    fjx = -ajx;
    fjy = -ajy;
}

__device__ __forceinline__ float2 ld_pos(const float2* __restrict__ R, int j, int N, float sent) {
    return (j < N) ? __ldcg(&R[j]) : make_float2(sent, sent);
}

template <bool CUTOFF, bool PE>
__device__ __forceinline__ void ap3_phase_forces(const ApArgs& a, const float2* __restrict__ Rcur,
                                                 float2* sacc /* 4 warps x 2 x q x 64 */, float* sred,
                                                 int par) {
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int wr = w >> 1, wc = w & 1, q = a.q, Pt = a.Pt;
    const PairConsts pc = a.pc;
    const PairConsts2 pc2 = make_pair_consts2(pc);
    float2* rowacc = sacc + (size_t)w * 2 * q * T3_BLK;        // this warp's [q][64] row-side sums
    float2* colacc = rowacc + (size_t)q * T3_BLK;              //             [q][64] column-side sums
    // CTA c takes patch c, then draws further patches from a per-step counter: rows of the triangle
    // differ in cost (diagonal patches, padding) and so do the SMs' shares of the L2; results and
    // partial sums are indexed by patch, so nothing depends on which CTA ran it.
    __shared__ int s_pi;
    for (int it = 0;; ++it) {
        if (tid == 0) s_pi = (it == 0) ? c : a.G + atomicAdd(&a.sched[par], 1);
        __syncthreads();
        const int pi = s_pi;
        if (pi >= a.npatch) break;
        float pe_thread = 0.0f;
        const int2 pp = a.patches[pi];
        const int pid = tri_base(pp.x, a.npr) + (pp.y - pp.x);
        for (int k = lane; k < q * T3_BLK; k += 32) colacc[k] = make_float2(0.0f, 0.0f);
        __syncwarp();
        for (int r = 0; r < q; ++r) {
            const int ba = pp.x * Pt + wr * q + r;              // i block
            const int i0 = ba * T3_BLK + lane, i1 = i0 + 32;
            const float2 p0 = ld_pos(Rcur, i0, a.N, SENT_I), p1 = ld_pos(Rcur, i1, a.N, SENT_I);
            const float2 xi2 = make_float2(p0.x, p1.x), yi2 = make_float2(p0.y, p1.y);
            float2 fx2 = make_float2(0.0f, 0.0f), fy2 = fx2;
            for (int cc = 0; cc < q; ++cc) {
                const int bb = pp.y * Pt + wc * q + cc;         // j block
                if (ba > bb) continue;                          // lower triangle (diagonal patches only)
                float2 pe2 = make_float2(0.0f, 0.0f);
                float2 tfx = make_float2(0.0f, 0.0f), tfy = tfx;   // per-tile sums (bounded chains)
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    const int j = bb * T3_BLK + h * 32 + lane;
                    const float2 pj = ld_pos(Rcur, j, a.N, SENT_J);
                    float fjx, fjy;
                    if (ba < bb) {
                        tile_half<CUTOFF, PE, true>(pc, pc2, xi2, yi2, pj.x, pj.y, false, false, tfx, tfy, pe2, fjx, fjy);
                        float2 acc = colacc[cc * T3_BLK + h * 32 + lane];
                        acc.x += fjx; acc.y += fjy;
                        colacc[cc * T3_BLK + h * 32 + lane] = acc;
                    } else {
                        tile_half<CUTOFF, PE, false>(pc, pc2, xi2, yi2, pj.x, pj.y, h == 0, h == 1, tfx, tfy, pe2, fjx, fjy);
                    }
                }
                fx2.x += tfx.x; fx2.y += tfx.y; fy2.x += tfy.x; fy2.y += tfy.y;
                // energy bookkeeping in the ORDERED-pair convention of the caller (0.5 * sum):
                // an unordered pair of an off-diagonal tile counts twice, a diagonal tile is ordered
                if (PE) pe_thread += (ba < bb ? 2.0f : 1.0f) * (pe2.x + pe2.y);
            }
            rowacc[r * T3_BLK + lane]      = make_float2(fx2.x, fy2.x);
            rowacc[r * T3_BLK + 32 + lane] = make_float2(fx2.y, fy2.y);
        }
        __syncthreads();
        // combine the two warps that share a patch row / column (fixed order) -> global partials
        float2* rp = a.rowpart + (size_t)pid * Pt * T3_BLK;
        float2* cp = a.colpart + (size_t)pid * Pt * T3_BLK;
        for (int k = tid; k < Pt * T3_BLK; k += AP_THREADS) {
            const int blk = k / T3_BLK, off = k - blk * T3_BLK;
            const int g2 = blk / q, rr = blk - g2 * q;          // warp-grid coordinate, tile within warp
            const float2* ra0 = sacc + (size_t)(g2 * 2 + 0) * 2 * q * T3_BLK + rr * T3_BLK + off;          // wr=g2, wc=0
            const float2* ra1 = sacc + (size_t)(g2 * 2 + 1) * 2 * q * T3_BLK + rr * T3_BLK + off;          // wr=g2, wc=1
            const float2* ca0 = sacc + (size_t)(0 * 2 + g2) * 2 * q * T3_BLK + (q + rr) * T3_BLK + off;    // wr=0, wc=g2
            const float2* ca1 = sacc + (size_t)(1 * 2 + g2) * 2 * q * T3_BLK + (q + rr) * T3_BLK + off;    // wr=1, wc=g2
            __stcg(&rp[k], make_float2(ra0->x + ra1->x, ra0->y + ra1->y));
            __stcg(&cp[k], make_float2(ca0->x + ca1->x, ca0->y + ca1->y));
        }
        if (PE) {                                    // per-patch energy partial (fixed reduction tree)
            float t = block_sum<AP_THREADS>(pe_thread, sred);
            if (tid == 0) __stcg(&a.pe_part[par * a.npe + pid], t);
        }
        __syncthreads();
    }
}

template <int IPT, bool CUTOFF, bool PE>
__device__ __forceinline__ void ap_phase_forces(const ApArgs& a, const float2* __restrict__ Rcur,
                                                float4* sj, float* sred, int par) {
    constexpr int BI = AP_THREADS * IPT;
    const int c = blockIdx.x, tid = threadIdx.x;
    long long w = a.cta_start[c];
    const long long w1 = a.cta_start[c + 1];
    int seg = 0;
    float pe_thread = 0.0f;
    while (w < w1) {
        const int ib = (int)(w / a.NJu);
        const int ju0 = (int)(w - (long long)ib * a.NJu);
        const int len = (int)min((long long)(a.NJu - ju0), w1 - w);
        float fx[IPT], fy[IPT];
        ap_segment<IPT, CUTOFF, PE>(a, Rcur, ib, ju0, len, sj, fx, fy, pe_thread);
        float2* dst = a.part + ((size_t)c * a.maxseg + seg) * BI;
#pragma unroll
        for (int k = 0; k < IPT; ++k) __stcg(&dst[k * AP_THREADS + tid], make_float2(fx[k], fy[k]));
        w += len;
        ++seg;
    }
    if (PE) {
        float t = block_sum<AP_THREADS>(pe_thread, sred);
        if (tid == 0) __stcg(&a.pe_part[par * a.npe + c], t);
    }
}

// fixed-order sum of a per-CTA partial array by one warp, in double
__device__ __forceinline__ double warp_sum_array(const float* p, int n) {
    double s = 0.0;
    for (int k = threadIdx.x & 31; k < n; k += 32) s += (double)__ldcg(&p[k]);
    return warp_sum(s);
}

template <int IPT, bool CUTOFF>
__global__ void __launch_bounds__(AP_THREADS, AP_MINBLOCKS)
ap_persistent_kernel(const ApArgs a) {
    constexpr int BI = AP_THREADS * (IPT == 3 ? 1 : IPT);
    constexpr bool V3 = (IPT == 3);
    __shared__ float4 smem_buf[V3 ? 2048 : TILE_J];          // v2: j tile; v3: per-warp tile sums (32 KB)
    float4* sj = smem_buf;
    __shared__ float  sred[AP_THREADS / 32];
    __shared__ float  s_lambda;
    const int c = blockIdx.x, tid = threadIdx.x;
    const RunCtl rc = a.rc;
    unsigned epoch = 0;
    __shared__ long long pt[5];                              // debug phase clocks (thread 0 only)
    const bool prof = (a.prof != nullptr) && tid == 0;
    if (prof) { pt[0] = pt[1] = pt[2] = pt[3] = 0; }

    for (long long s = a.s_begin; s < a.s_end; ++s) {
        const float2* Rcur  = (s < 0) ? a.R_in : ((s & 1) ? a.Rbuf1 : a.Rbuf0);
        float2*       Rnext = ((s + 1) & 1) ? a.Rbuf1 : a.Rbuf0;
        const int  par     = (int)((s + 1) & 1);
        const bool kick1   = (s >= 0);                     // prologue only evaluates F(R_in)
        const bool final   = (s == rc.nsteps - 1);
        const bool want_e  = kick1 && rc.energy_every > 0 && (s % rc.energy_every == 0);
        const bool want_pe = want_e || (rc.nsteps == 0 && a.pe_out != nullptr);
        const bool thermo  = kick1 && rc.thermo_every > 0 && rc.thermo_kT > 0.0f &&
                             ((s + 1) % rc.thermo_every == 0);
        const bool sample  = kick1 && rc.sample_every > 0 && (s % rc.sample_every == 0) &&
                             (s / rc.sample_every < rc.S);   // MD:93-100

        if (prof) pt[4] = clock64();
        // ---- [b] partial forces of R_cur ------------------------------------------------------
        if constexpr (V3) {
            if (c == 0 && tid == 0) __stcg(&a.sched[par ^ 1], 0);   // the other parity's patch counter
            if (want_pe) ap3_phase_forces<CUTOFF, true >(a, Rcur, reinterpret_cast<float2*>(smem_buf), sred, par);
            else         ap3_phase_forces<CUTOFF, false>(a, Rcur, reinterpret_cast<float2*>(smem_buf), sred, par);
        } else {
            if (want_pe) ap_phase_forces<IPT, CUTOFF, true >(a, Rcur, sj, sred, par);
            else         ap_phase_forces<IPT, CUTOFF, false>(a, Rcur, sj, sred, par);
        }
        if (prof) { long long t = clock64(); pt[0] += t - pt[4]; pt[4] = t; }
        grid_barrier(a.bar, (++epoch) * (unsigned)a.G, a.err);
        if (prof) { long long t = clock64(); pt[1] += t - pt[4]; pt[4] = t; }

        // ---- [d] reduce partials, finish the velocity-Verlet step -----------------------------
        // CTA c owns particles [c*ppc, (c+1)*ppc); a group of RED_LANES lanes sums the partials of
        // one particle (strided over the contributing CTAs, then a fixed xor-shuffle tree), and
        // lane 0 of the group integrates it.  Fixed order => bit-reproducible, no atomics.
        float ke_thread = 0.0f;
        const int grp = tid / RED_LANES, gl = tid % RED_LANES;
        const int g_end = min(a.Nloc, (c + 1) * a.ppc);                          // slab-local indices
        for (int gb = c * a.ppc; gb < g_end; gb += AP_THREADS / RED_LANES) {   // block-uniform trip count
            const int gloc = gb + grp;
            const int g = a.i_lo + gloc;                                       // global particle index
            const bool live = gloc < g_end;
            float Fx = 0.0f, Fy = 0.0f;
            float2 r = make_float2(0.0f, 0.0f), v = make_float2(0.0f, 0.0f);
            if (live && gl == 0) {                          // issued early: overlaps the partial loads
                r = __ldcg(&Rcur[g]);
                if (rc.nsteps > 0) v = a.Vh[g];
            }
            if (live) {
                if constexpr (V3) {
                    // row-side partials of patches (pr, pb >= pr), then column-side of (pa <= pr, pr)
                    const int blk = g / T3_BLK, pr = blk / a.Pt;
                    const size_t off = (size_t)(blk - pr * a.Pt) * T3_BLK + (g - blk * T3_BLK);
                    const size_t pstride = (size_t)a.Pt * T3_BLK;
                    const int n1 = a.npr - pr, ntot = n1 + pr + 1;
#pragma unroll 4
                    for (int t = gl; t < ntot; t += RED_LANES) {
                        float2 p;
                        if (t < n1) p = __ldcg(&a.rowpart[(size_t)(tri_base(pr, a.npr) + t) * pstride + off]);
                        else { const int pa = t - n1; p = __ldcg(&a.colpart[(size_t)(tri_base(pa, a.npr) + (pr - pa)) * pstride + off]); }
                        Fx += p.x; Fy += p.y;
                    }
                } else {
                    const int ib = gloc / BI, il = gloc - ib * BI;
                    const int2 cc = a.iblk_ctas[ib];
#pragma unroll 4
                    for (int c2 = cc.x + gl; c2 <= cc.y; c2 += RED_LANES) {
                        const int seg = ib - a.cta_ib0[c2];
                        const float2 p = __ldcg(&a.part[((size_t)c2 * a.maxseg + seg) * BI + il]);
                        Fx += p.x; Fy += p.y;
                    }
                }
            }
#pragma unroll
            for (int o = RED_LANES / 2; o > 0; o >>= 1) {
                Fx += __shfl_xor_sync(0xffffffffu, Fx, o);
                Fy += __shfl_xor_sync(0xffffffffu, Fy, o);
            }
            if (!live || gl != 0) continue;
            if (kick1) { v.x = kick(v.x, Fx, a.dt); v.y = kick(v.y, Fy, a.dt); }   // MD:74
            if (want_e || thermo) ke_thread += v.x * v.x + v.y * v.y;
            if (sample) rc.traj[(size_t)(s / rc.sample_every) * a.N + g] = r;
            if (thermo) {                                   // finish after the KE barrier
                a.Vh[g] = v;
                a.Ftmp[g] = make_float2(Fx, Fy);
                continue;
            }
            if (final) {
                if (a.R_out) a.R_out[g] = r;
                if (a.V_out) a.V_out[g] = v;
                if (a.F_out) a.F_out[g] = make_float2(Fx, Fy);
            } else {
                v.x = kick(v.x, Fx, a.dt); v.y = kick(v.y, Fy, a.dt);                // MD:70
                a.Vh[g] = v;
                const float2 rn = make_float2(drift(r.x, v.x, a.dt, a.pc.box),      // MD:71-72
                                              drift(r.y, v.y, a.dt, a.pc.box));
                __stcg(&Rnext[g], rn);
                if (a.P > 1) {                               // position "all-gather": direct NVLink stores
                    const int nb = (int)((s + 1) & 1);
                    for (int q = 0; q < a.P; ++q)
                        if (q != a.rank) (nb ? a.peerR1[q] : a.peerR0[q])[g] = rn;
                }
            }
        }
        if (a.P > 1) __threadfence_system();                 // peer stores visible before the arrival
        if (want_e || thermo) {
            float t = block_sum<AP_THREADS>(ke_thread, sred);
            if (tid == 0) __stcg(&a.ke_part[par * a.G + c], t);
        }
        if (thermo) {
            // velocity rescale: V *= sqrt(kT_target / (KE/N)), KE = 0.5*sum|V|^2
            grid_barrier(a.bar, (++epoch) * (unsigned)a.G, a.err);
            if (tid < 32) {
                double ke2 = warp_sum_array(a.ke_part + par * a.G, a.G);
                if (tid == 0) {
                    float ke = (float)(0.5 * ke2);
                    s_lambda = sqrtf(rc.thermo_kT / (ke / (float)a.N));   // (single-GPU only)
                }
            }
            __syncthreads();
            const float lam = s_lambda;
            for (int gloc = c * a.ppc + tid; gloc < g_end; gloc += AP_THREADS) {
                const int g = a.i_lo + gloc;
                const float2 r = __ldcg(&Rcur[g]);
                const float2 F = a.Ftmp[g];
                float2 v = a.Vh[g];
                v.x *= lam; v.y *= lam;
                if (final) {
                    if (a.R_out) a.R_out[g] = r;
                    if (a.V_out) a.V_out[g] = v;
                    if (a.F_out) a.F_out[g] = F;
                } else {
                    v.x = kick(v.x, F.x, a.dt); v.y = kick(v.y, F.y, a.dt);
                    a.Vh[g] = v;
                    __stcg(&Rnext[g], make_float2(drift(r.x, v.x, a.dt, a.pc.box),
                                                  drift(r.y, v.y, a.dt, a.pc.box)));
                }
            }
        }
        if (prof) { long long t = clock64(); pt[2] += t - pt[4]; pt[4] = t; }
        grid_barrier(a.bar, (++epoch) * (unsigned)a.G, a.err);
        if (a.P > 1 && !final) {
            // every rank has pushed its slab into our next-position buffer once all P arrival words
            // carry this epoch: one NVLink round trip per step, no NCCL on the step path
            const unsigned xe = a.xepoch0 + (unsigned)(s - a.s_begin) + 1u;
            if (c == 0 && tid < a.P && tid != a.rank) {
                __threadfence_system();
                volatile unsigned* f = a.peer_flags[tid] + a.rank;
                *f = xe;
            }
            if (tid == 0) {
                const volatile unsigned* mine = a.peer_flags[a.rank];
                long long t0 = clock64();
                for (int q = 0; q < a.P; ++q) {
                    if (q == a.rank) continue;
                    while ((int)(mine[q] - xe) < 0) {
                        if (clock64() - t0 > (1ll << 33)) { atomicExch(a.err, 2); break; }
                    }
                }
                __threadfence_system();
            }
            __syncthreads();
        }
        if (prof) { long long t = clock64(); pt[3] += t - pt[4]; pt[4] = t; }

        // ---- energies of the post-step state (one warp, fixed order, double combine) ----------
        if (c == 0 && tid < 32 && want_pe) {
            double pe2 = warp_sum_array(a.pe_part + par * a.npe, a.npe);
            double ke2 = want_e ? warp_sum_array(a.ke_part + par * a.G, a.G) : 0.0;
            if (tid == 0) {
                if (want_e) {
                    float* o = rc.ke_pe + 2 * (s / rc.energy_every);
                    o[0] = (float)(0.5 * ke2);
                    o[1] = (float)(0.5 * pe2);             // MD:61  0.5 * sum over ordered pairs
                } else {
                    a.pe_out[0] = (float)(0.5 * pe2);
                }
            }
        }
    }
    if (prof) {
#pragma unroll
        for (int k = 0; k < 4; ++k) a.prof[c * 4 + k] = pt[k];
    }
}

using ApKernel = void (*)(const ApArgs);

ApKernel pick_kernel(int ipt, bool cutoff) {
    if (ipt == 3) return cutoff ? ap_persistent_kernel<3, true> : ap_persistent_kernel<3, false>;
    if (ipt == 1) return cutoff ? ap_persistent_kernel<1, true> : ap_persistent_kernel<1, false>;
    return cutoff ? ap_persistent_kernel<2, true> : ap_persistent_kernel<2, false>;
}

// ---- g(r): pair-distance histogram (get_histogram, MD:117-124) -------------------------------
constexpr int GR_THREADS = 256;
constexpr int GR_TILE    = 256;
constexpr int GR_MAXBINS = 8192;

__global__ void __launch_bounds__(GR_THREADS)
gr_hist_kernel(const float2* __restrict__ Rh, int N, float box, float timg, int nbins,
               const float* __restrict__ edges, unsigned long long* __restrict__ counts,
               int ntile) {
    // blockIdx.x -> (tile_i <= tile_j) upper-triangular tile pair; blockIdx.y -> snapshot
    extern __shared__ unsigned char smem_raw[];
    float2*   sj     = reinterpret_cast<float2*>(smem_raw);                 // GR_TILE (8-byte aligned)
    float*    sedges = reinterpret_cast<float*>(sj + GR_TILE);              // nbins+1
    unsigned* shist  = reinterpret_cast<unsigned*>(sedges + nbins + 1);     // nbins
    const int tid = threadIdx.x;
    int t = blockIdx.x, ti = 0;
    while (t >= ntile - ti) { t -= ntile - ti; ++ti; }
    const int tj = ti + t;
    const float2* R = Rh + (size_t)blockIdx.y * N;
    for (int k = tid; k <= nbins; k += GR_THREADS) sedges[k] = edges[k];
    for (int k = tid; k < nbins; k += GR_THREADS) shist[k] = 0u;
    {
        const int j = tj * GR_TILE + tid;
        if (tid < GR_TILE) sj[tid] = (j < N) ? R[j] : make_float2(SENT_J, SENT_J);
    }
    __syncthreads();
    const int i = ti * GR_TILE + tid;
    if (i < N) {
        const float2 ri = R[i];
        const float lo = sedges[0], hi = sedges[nbins];
        const float scale = (float)nbins / (hi - lo);
        for (int q = 0; q < GR_TILE; ++q) {
            const int j = tj * GR_TILE + q;
            if (j <= i || j >= N) continue;                                  // triu, k=1
            const float dx = min_image(__fsub_rn(ri.x, sj[q].x), box, timg);
            const float dy = min_image(__fsub_rn(ri.y, sj[q].y), box, timg);
            const float r = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
            if (!(r >= lo && r <= hi)) continue;                             // numpy drops outliers
            int b = (int)((r - lo) * scale);
            b = max(0, min(b, nbins - 1));
            while (b > 0 && r < sedges[b]) --b;                              // exact against edges
            while (b < nbins - 1 && r >= sedges[b + 1]) ++b;
            atomicAdd(&shist[b], 1u);
        }
    }
    __syncthreads();
    unsigned long long* out = counts + (size_t)blockIdx.y * nbins;
    for (int k = tid; k < nbins; k += GR_THREADS)
        if (shist[k]) atomicAdd(&out[k], (unsigned long long)shist[k]);
}

}  // namespace

// ----------------------------------------------------------------------------------------------
struct AllPairs {
    int ipt = 1, G = 0, NJu = 0, nI = 0, maxseg = 0;
    int Pt = 0, q = 0, npr = 0, npatch = 0, npe = 0;   // Newton's-third-law tile mode (ipt == 3)
    int* sched = nullptr;
    int*  d_cta_pstart = nullptr;
    int2* d_patches = nullptr;
    float2 *rowpart = nullptr, *colpart = nullptr;
    long long* d_cta_start = nullptr;
    int*       d_cta_ib0 = nullptr;
    int2*      d_iblk = nullptr;
    float2 *Rbuf0 = nullptr, *Rbuf1 = nullptr, *Vh = nullptr, *Ftmp = nullptr, *part = nullptr;
    float *pe_part = nullptr, *ke_part = nullptr;
    int i_lo = 0, Nloc = 0;
    void* shared = nullptr;             // one allocation (IPC-shareable): Rbuf0 | Rbuf1 | arrival words
    unsigned* xflags = nullptr;
    unsigned  xepoch = 0;
    float2*   peerR0[LJMD_MAX_RANKS] = {};
    float2*   peerR1[LJMD_MAX_RANKS] = {};
    unsigned* peer_flags[LJMD_MAX_RANKS] = {};
    unsigned* bar = nullptr;
    int* err = nullptr;
    long long* prof = nullptr;
    ApKernel kernel = nullptr;
};

int ap_mode(ljmd_handle* h) { return h->ap ? h->ap->ipt : 0; }

int ap_create(ljmd_handle* h) {
    AllPairs* ap = new AllPairs();
    h->ap = ap;
    const long long N = h->p.N;
    const int P = std::max(1, h->nranks);
    // ipt 1 / 2: ordered pairs, one / two i per thread (stream-K split, any rank count);
    // ipt 3: Newton's-third-law tiles (each unordered pair once, single GPU)
    ap->ipt = (N >= 2048) ? ((P == 1) ? 3 : 2) : 1;
    if (const char* e = getenv("LJMD_AP_IPT")) { const int v = atoi(e); if (v >= 1 && v <= 3 && (v != 3 || P == 1)) ap->ipt = v; }
    const int BI = AP_THREADS * (ap->ipt == 3 ? 1 : ap->ipt);
    if (N % P != 0) { set_error("all-pairs atom decomposition needs N divisible by the rank count"); return LJMD_E_INVALID; }
    ap->Nloc = (int)(N / P);
    ap->i_lo = h->rank * ap->Nloc;
    ap->nI  = (ap->Nloc + BI - 1) / BI;             // i-blocks of THIS rank's slab
    ap->NJu = (int)((N + J_UNIT - 1) / J_UNIT);
    ap->kernel = pick_kernel(ap->ipt, h->pc.cutoff != 0);

    int per_sm = 0;
    LJ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ap->kernel, AP_THREADS, 0));
    int want = 4;
    if (const char* e = getenv("LJMD_AP_CTAS_PER_SM")) want = std::max(1, atoi(e));
    per_sm = std::min(per_sm, want);
    if (per_sm < 1) { set_error("all-pairs kernel does not fit on an SM"); return LJMD_E_STATE; }
    const long long W = (long long)ap->nI * ap->NJu;
    // at least ~64 j per thread per CTA so tiny systems do not pay for a wide barrier
    long long g_work = std::max<long long>(1, W / (64 / J_UNIT));
    ap->G = (int)std::min<long long>((long long)per_sm * h->num_sms, std::min(g_work, W));
    if (const char* e = getenv("LJMD_AP_GRID")) ap->G = std::max(1, std::min(atoi(e), ap->G));
    std::vector<int> pstart;
    std::vector<int2> patches;
    if (ap->ipt == 3) {
        // upper-triangular patches of Pt x Pt tiles (tile = 64 x 64 particles); a CTA's 2 x 2 warps take
        // q x q tiles each (Pt = 2q).  A diagonal patch does half the work of an off-diagonal one but
        // takes the same time (its busiest warp has q*q tiles), so patches are dealt out evenly.
        const int blocks = (int)((N + T3_BLK - 1) / T3_BLK);
        const long long slots = (long long)per_sm * h->num_sms;
        int q = 8;
        for (; q > 1; q >>= 1) {
            const long long npr = (blocks + 2 * q - 1) / (2 * q);
            if (npr * (npr + 1) / 2 >= 4 * slots) break;          // enough patches to balance
        }
        if (const char* e = getenv("LJMD_AP_Q")) q = std::max(1, std::min(8, atoi(e)));
        ap->q = q; ap->Pt = 2 * q;
        ap->npr = (blocks + ap->Pt - 1) / ap->Pt;
        for (int pa = 0; pa < ap->npr; ++pa)
            for (int pb = pa; pb < ap->npr; ++pb) patches.push_back(make_int2(pa, pb));
        const long long npatch = (long long)patches.size();
        ap->npatch = (int)npatch;
        ap->G = (int)std::min<long long>(slots, npatch);
        if (const char* e = getenv("LJMD_AP_GRID")) ap->G = std::max(1, std::min(atoi(e), ap->G));
        pstart.resize(ap->G + 1);
        for (int c = 0; c <= ap->G; ++c) pstart[c] = (int)(npatch * c / ap->G);
    }

    // stream-K split of the flattened (i-block, j-unit) space, by COST: a j unit that overlaps
    // its own i-block runs the index-tested loop (~16% more instructions), so it weighs more.
    const long long W_PLAIN = 100, W_DIAG = (ap->ipt == 2) ? 116 : 110;
    const int du = BI / J_UNIT;                       // diagonal units per row
    auto row_cost = [&](int b, long long u) {         // cost of units [0, u) of row b
        const long long d0 = std::min<long long>((long long)ap->i_lo / J_UNIT + (long long)b * du, ap->NJu);
        const long long d1 = std::min<long long>(d0 + du, ap->NJu);
        const long long nd = std::max<long long>(0, std::min(u, d1) - d0);
        return u * W_PLAIN + nd * (W_DIAG - W_PLAIN);
    };
    std::vector<long long> rowsum(ap->nI + 1, 0);
    for (int b = 0; b < ap->nI; ++b) rowsum[b + 1] = rowsum[b] + row_cost(b, ap->NJu);
    const long long total_cost = rowsum[ap->nI];
    std::vector<long long> start(ap->G + 1);
    {
        int b = 0;
        for (int c = 0; c <= ap->G; ++c) {
            const long long target = (long long)(((__int128)total_cost * c) / ap->G);
            while (b + 1 < ap->nI && rowsum[b + 1] <= target) ++b;
            long long lo = 0, hi = ap->NJu;          // smallest u with rowsum[b] + row_cost(b,u) >= target
            while (lo < hi) {
                const long long mid = (lo + hi) / 2;
                if (rowsum[b] + row_cost(b, mid) >= target) hi = mid; else lo = mid + 1;
            }
            start[c] = (long long)b * ap->NJu + lo;
        }
        start[0] = 0;
        start[ap->G] = W;
        for (int c = 1; c <= ap->G; ++c) start[c] = std::max(start[c], start[c - 1]);
    }
    std::vector<int> ib0(ap->G);
    int maxseg = 1;
    for (int c = 0; c < ap->G; ++c) {
        ib0[c] = (int)(start[c] / ap->NJu);
        int ibl = (int)((start[c + 1] - 1) / ap->NJu);
        maxseg = std::max(maxseg, ibl - ib0[c] + 1);
    }
    ap->maxseg = maxseg;
    std::vector<int2> iblk(ap->nI);
    {
        int c = 0;
        for (int b = 0; b < ap->nI; ++b) {
            const long long lo = (long long)b * ap->NJu, hi = lo + ap->NJu;
            while (start[c + 1] <= lo) ++c;
            int cl = c;
            while (cl + 1 < ap->G && start[cl + 1] < hi) ++cl;
            iblk[b] = make_int2(c, cl);
        }
    }
    LJ_CUDA(cudaMalloc(&ap->d_cta_start, sizeof(long long) * (ap->G + 1)));
    LJ_CUDA(cudaMalloc(&ap->d_cta_ib0, sizeof(int) * ap->G));
    LJ_CUDA(cudaMalloc(&ap->d_iblk, sizeof(int2) * ap->nI));
    LJ_CUDA(cudaMemcpy(ap->d_cta_start, start.data(), sizeof(long long) * (ap->G + 1), cudaMemcpyHostToDevice));
    LJ_CUDA(cudaMemcpy(ap->d_cta_ib0, ib0.data(), sizeof(int) * ap->G, cudaMemcpyHostToDevice));
    LJ_CUDA(cudaMemcpy(ap->d_iblk, iblk.data(), sizeof(int2) * ap->nI, cudaMemcpyHostToDevice));
    {   // one >= 2 MiB allocation so that its IPC handle maps exactly this region on the peers
        const size_t rb = (sizeof(float2) * (size_t)N + 255) / 256 * 256;
        size_t bytes = 2 * rb + 256;
        bytes = std::max<size_t>((bytes + (2u << 20) - 1) / (2u << 20) * (2u << 20), 4u << 20);
        LJ_CUDA(cudaMalloc(&ap->shared, bytes));
        LJ_CUDA(cudaMemset(ap->shared, 0, bytes));
        ap->Rbuf0 = reinterpret_cast<float2*>(ap->shared);
        ap->Rbuf1 = reinterpret_cast<float2*>(reinterpret_cast<char*>(ap->shared) + rb);
        ap->xflags = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ap->shared) + 2 * rb);
        ap->peerR0[h->rank] = ap->Rbuf0; ap->peerR1[h->rank] = ap->Rbuf1; ap->peer_flags[h->rank] = ap->xflags;
        if (P > 1) {
            void* peers[LJMD_MAX_RANKS];
            int r = dist_share(h, ap->shared, peers);
            if (r) return r;
            for (int q = 0; q < P; ++q) {
                char* base = reinterpret_cast<char*>(peers[q]);
                ap->peerR0[q] = reinterpret_cast<float2*>(base);
                ap->peerR1[q] = reinterpret_cast<float2*>(base + rb);
                ap->peer_flags[q] = reinterpret_cast<unsigned*>(base + 2 * rb);
            }
        }
    }
    LJ_CUDA(cudaMalloc(&ap->Vh, sizeof(float2) * N));
    LJ_CUDA(cudaMalloc(&ap->Ftmp, sizeof(float2) * N));
    LJ_CUDA(cudaMalloc(&ap->part, sizeof(float2) * (size_t)ap->G * ap->maxseg * BI));
    if (ap->ipt == 3) {
        const size_t np = patches.size(), ps = (size_t)ap->Pt * T3_BLK;
        LJ_CUDA(cudaMalloc(&ap->d_cta_pstart, sizeof(int) * pstart.size()));
        LJ_CUDA(cudaMalloc(&ap->d_patches, sizeof(int2) * np));
        LJ_CUDA(cudaMemcpy(ap->d_cta_pstart, pstart.data(), sizeof(int) * pstart.size(), cudaMemcpyHostToDevice));
        LJ_CUDA(cudaMemcpy(ap->d_patches, patches.data(), sizeof(int2) * np, cudaMemcpyHostToDevice));
        LJ_CUDA(cudaMalloc(&ap->rowpart, sizeof(float2) * np * ps));
        LJ_CUDA(cudaMalloc(&ap->colpart, sizeof(float2) * np * ps));
    }
    ap->npe = (ap->ipt == 3) ? (int)patches.size() : ap->G;
    LJ_CUDA(cudaMalloc(&ap->pe_part, sizeof(float) * 2 * ap->npe));
    LJ_CUDA(cudaMalloc(&ap->sched, sizeof(int) * 2));
    LJ_CUDA(cudaMalloc(&ap->ke_part, sizeof(float) * 2 * ap->G));
    LJ_CUDA(cudaMalloc(&ap->bar, sizeof(unsigned)));
    LJ_CUDA(cudaMalloc(&ap->err, sizeof(int)));
    LJ_CUDA(cudaMemset(ap->err, 0, sizeof(int)));
    if (getenv("LJMD_AP_PROF")) LJ_CUDA(cudaMalloc(&ap->prof, sizeof(long long) * 4 * ap->G));
    return 0;
}

void ap_destroy(ljmd_handle* h) {
    AllPairs* ap = h->ap;
    if (!ap) return;
    cudaFree(ap->d_cta_start); cudaFree(ap->d_cta_ib0); cudaFree(ap->d_iblk);
    cudaFree(ap->shared); cudaFree(ap->Vh); cudaFree(ap->Ftmp);
    cudaFree(ap->part); cudaFree(ap->pe_part); cudaFree(ap->ke_part);
    cudaFree(ap->d_cta_pstart); cudaFree(ap->d_patches); cudaFree(ap->rowpart); cudaFree(ap->colpart);
    cudaFree(ap->sched);
    cudaFree(ap->bar); cudaFree(ap->err);
    delete ap;
    h->ap = nullptr;
}

int ap_run(ljmd_handle* h, const float2* R_in, const float2* V_in, float2* R_out, float2* V_out,
           float2* F_out, float* pe_out, const RunCtl& rc) {
    AllPairs* ap = h->ap;
    const long long N = h->p.N;
    cudaStream_t st = h->stream;
    if (rc.nsteps > 0) {
        LJ_CUDA(cudaMemcpyAsync(ap->Vh, V_in, sizeof(float2) * N, cudaMemcpyDeviceToDevice, st));
        if (rc.traj && rc.S > 0)
            LJ_CUDA(cudaMemsetAsync(rc.traj, 0, sizeof(float2) * N * rc.S, st));   // MD:89
    }
    ApArgs a{};
    a.pc = h->pc;
    a.ppc = (ap->Nloc + ap->G - 1) / ap->G;
    a.i_lo = ap->i_lo; a.Nloc = ap->Nloc; a.rank = h->rank; a.P = std::max(1, h->nranks);
    for (int q = 0; q < LJMD_MAX_RANKS; ++q) { a.peerR0[q] = ap->peerR0[q]; a.peerR1[q] = ap->peerR1[q]; a.peer_flags[q] = ap->peer_flags[q]; }
    if (a.P > 1 && rc.thermo_every > 0) { set_error("the rescale thermostat is single-GPU only"); return LJMD_E_UNSUPPORTED; }
    a.N = (int)N; a.G = ap->G; a.NJu = ap->NJu; a.nI = ap->nI; a.maxseg = ap->maxseg;
    a.dt = h->p.dt;
    a.cta_start = ap->d_cta_start; a.cta_ib0 = ap->d_cta_ib0; a.iblk_ctas = ap->d_iblk;
    a.R_in = R_in; a.Rbuf0 = ap->Rbuf0; a.Rbuf1 = ap->Rbuf1; a.Vh = ap->Vh; a.Ftmp = ap->Ftmp;
    a.part = ap->part; a.pe_part = ap->pe_part; a.ke_part = ap->ke_part;
    a.Pt = ap->Pt; a.q = ap->q; a.npr = ap->npr; a.npatch = ap->npatch; a.npe = ap->npe; a.sched = ap->sched;
    a.cta_pstart = ap->d_cta_pstart; a.patches = ap->d_patches; a.rowpart = ap->rowpart; a.colpart = ap->colpart;
    a.bar = ap->bar; a.err = ap->err; a.prof = ap->prof;
    a.rc = rc;
    a.R_out = R_out; a.V_out = V_out; a.F_out = F_out; a.pe_out = pe_out;

    // bound the duration of a single launch (~0.5 s at a conservative 2e11 pairs/s)
    const double est_step_s = (double)N * (double)N / 2.0e11 + 3.0e-6;
    long long chunk = (long long)std::max(1.0, 0.5 / est_step_s);
    if (const char* e = getenv("LJMD_AP_CHUNK")) chunk = std::max(1, atoi(e));

    if (h->timed) LJ_CUDA(cudaEventRecord(h->ev0, st));
    long long s = -1;
    const long long s_last = rc.nsteps;     // exclusive
    while (s < s_last) {
        const long long e = std::min(s_last, s + chunk);
        a.s_begin = s; a.s_end = e;
        a.xepoch0 = ap->xepoch;
        if (a.P > 1) { int rb = dist_barrier(h); if (rb) return rb; }   // the ranks enter the kernel together
        LJ_CUDA(cudaMemsetAsync(ap->bar, 0, sizeof(unsigned), st));
        LJ_CUDA(cudaMemsetAsync(ap->sched, 0, sizeof(int) * 2, st));
        void* args[] = {(void*)&a};
        LJ_CUDA(cudaLaunchCooperativeKernel((void*)ap->kernel, dim3(ap->G), dim3(AP_THREADS), args, 0, st));
        h->launches++;
        ap->xepoch += (unsigned)(e - s);             // one cross-GPU epoch per step of the launch
        s = e;
    }
    if (a.P > 1) {
        // replicated out: slabs of the final state -> every rank (NCCL, once per call, not per step)
        const size_t slab = sizeof(float2) * (size_t)ap->Nloc;
        int r = 0;
        if (R_out && (r = dist_allgather(h, R_out, slab))) return r;
        if (V_out && (r = dist_allgather(h, V_out, slab))) return r;
        if (F_out && (r = dist_allgather(h, F_out, slab))) return r;
        for (long long k = 0; rc.traj && k < rc.S; ++k)
            if ((r = dist_allgather(h, rc.traj + (size_t)k * N, slab))) return r;
        if (pe_out && (r = dist_allreduce_f32(h, pe_out, 1))) return r;
        if (rc.ke_pe && rc.energy_every > 0) {       // "energies use one NCCL all-reduce"
            const long long ne = (rc.nsteps + rc.energy_every - 1) / rc.energy_every;
            if ((r = dist_allreduce_f32(h, rc.ke_pe, (size_t)(2 * ne)))) return r;
        }
    }
    if (h->timed) LJ_CUDA(cudaEventRecord(h->ev1, st));
    if (ap->prof) {   // debug: mean / max clocks per phase per step of the last launch
        LJ_CUDA(cudaStreamSynchronize(st));
        std::vector<long long> pv(4 * ap->G);
        LJ_CUDA(cudaMemcpy(pv.data(), ap->prof, sizeof(long long) * pv.size(), cudaMemcpyDeviceToHost));
        const double nst = (double)std::max<long long>(1, a.s_end - a.s_begin);
        const char* nm[4] = {"forces", "barrierA", "integrate", "barrierB"};
        for (int k = 0; k < 4; ++k) {
            double mean = 0, mx = 0;
            for (int c = 0; c < ap->G; ++c) { mean += pv[c * 4 + k]; mx = std::max<double>(mx, (double)pv[c * 4 + k]); }
            fprintf(stderr, "[ljmd prof] %-10s mean %9.0f  max %9.0f clocks/step (G=%d)\n", nm[k],
                    mean / ap->G / nst, mx / nst, ap->G);
        }
        if (getenv("LJMD_AP_PROF_CTAS")) {
            fprintf(stderr, "[ljmd prof] forces clocks/step by CTA:");
            for (int c = 0; c < ap->G; ++c) fprintf(stderr, "%s%d:%.0f", (c % 16) ? " " : "\n  ", c, pv[c * 4] / nst);
            fprintf(stderr, "\n");
        }
    }
    return 0;
}

int ap_check_error(ljmd_handle* h) {
    AllPairs* ap = h->ap;
    if (!ap) return 0;
    int e = 0;
    LJ_CUDA(cudaMemcpy(&e, ap->err, sizeof(int), cudaMemcpyDeviceToHost));
    if (e) {
        set_error("all-pairs persistent kernel: grid barrier timed out (device error flag %d)", e);
        return LJMD_E_STATE;
    }
    return 0;
}

int ap_gr_hist(ljmd_handle* h, const float2* R_hist, long long S, int nbins, const float* edges,
               long long* counts) {
    const int N = (int)h->p.N;
    if (nbins < 1 || nbins > GR_MAXBINS) { set_error("g(r): nbins must be in [1,%d]", GR_MAXBINS); return LJMD_E_INVALID; }
    cudaStream_t st = h->stream;
    LJ_CUDA(cudaMemsetAsync(counts, 0, sizeof(long long) * S * nbins, st));
    if (S == 0) return 0;
    const int ntile = (N + GR_TILE - 1) / GR_TILE;
    const long long npairs = (long long)ntile * (ntile + 1) / 2;
    const size_t smem = sizeof(float) * (nbins + 1) + sizeof(unsigned) * nbins + sizeof(float2) * GR_TILE;
    LJ_CUDA(cudaFuncSetAttribute(gr_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // grid.y is limited to 65535 snapshots per launch
    for (long long s0 = 0; s0 < S; s0 += 65535) {
        const int ns = (int)std::min<long long>(65535, S - s0);
        gr_hist_kernel<<<dim3((unsigned)npairs, ns), GR_THREADS, smem, st>>>(
            R_hist + (size_t)s0 * N, N, h->pc.box, h->pc.timg, nbins, edges,
            reinterpret_cast<unsigned long long*>(counts) + (size_t)s0 * nbins, ntile);
        LJ_CUDA(cudaGetLastError());
        h->launches++;
    }
    return 0;
}

}  // namespace ljmd
