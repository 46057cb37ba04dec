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
constexpr int J_UNIT     = 8;     // granularity of the j split (particles)
constexpr int TILE_J     = 512;   // j particles staged per shared-memory tile (4 KB)
constexpr float SENT_J   = 1.0e18f;    // padding particles: far away, contribute exactly 0
constexpr float SENT_I   = -1.0e18f;

struct ApArgs {
    PairConsts pc;
    int   N, G, NJu, nI, maxseg;
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
    float*           pe_part;     // [2*G] per-CTA partial potential energy (by step parity)
    float*           ke_part;     // [2*G]
    unsigned*        bar;
    int*             err;
    long long        s_begin, s_end;   // steps [s_begin, s_end); s = -1 is the prologue force
    RunCtl           rc;
    float2*          R_out;
    float2*          V_out;
    float2*          F_out;
    float*           pe_out;
};

// ---- phase [b]: one segment = (i-block ib) x (j units [ju0, ju0+julen)) ---------------------
template <int IPT, bool CUTOFF, bool PE>
__device__ __forceinline__ void ap_segment(const ApArgs& a, const float2* __restrict__ Rcur,
                                           int ib, int ju0, int julen, float4* sj,
                                           float (&fx)[IPT], float (&fy)[IPT], float& pe_acc) {
    constexpr int BI = AP_THREADS * IPT;
    const int tid = threadIdx.x;
    const PairConsts pc = a.pc;
    float xi[IPT], yi[IPT];
    int   ii[IPT];
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        ii[k] = ib * BI + k * AP_THREADS + tid;
        if (ii[k] < a.N) { float2 r = __ldcg(&Rcur[ii[k]]); xi[k] = r.x; yi[k] = r.y; }
        else             { xi[k] = SENT_I; yi[k] = SENT_I; }
        fx[k] = 0.0f; fy[k] = 0.0f;
    }
    const int j0 = ju0 * J_UNIT, jend = (ju0 + julen) * J_UNIT;
    const int i_lo = ib * BI, i_hi = i_lo + BI;

    for (int jt = j0; jt < jend; jt += TILE_J) {
        const int cnt = min(TILE_J, jend - jt);           // multiple of J_UNIT
        __syncthreads();                                   // previous tile fully consumed
        for (int q = tid; q < cnt; q += AP_THREADS) {
            const int j = jt + q;
            float2 r = (j < a.N) ? __ldcg(&Rcur[j]) : make_float2(SENT_J, SENT_J);
            reinterpret_cast<float2*>(sj)[q] = r;
        }
        __syncthreads();
        float tfx[IPT], tfy[IPT], tpe[IPT];               // per-tile sums (bounded chain length)
#pragma unroll
        for (int k = 0; k < IPT; ++k) { tfx[k] = 0.0f; tfy[k] = 0.0f; tpe[k] = 0.0f; }
        const bool diag = (jt < i_hi) && (jt + cnt > i_lo);   // tile contains some i == j
        if (!diag) {
#pragma unroll 4
            for (int q = 0; q < cnt / 2; ++q) {
                const float4 v = sj[q];                    // two j particles, warp broadcast
#pragma unroll
                for (int k = 0; k < IPT; ++k) {
                    pair_accum<CUTOFF, PE, false>(xi[k], yi[k], v.x, v.y, true, pc, tfx[k], tfy[k], tpe[k]);
                    pair_accum<CUTOFF, PE, false>(xi[k], yi[k], v.z, v.w, true, pc, tfx[k], tfy[k], tpe[k]);
                }
            }
        } else {
#pragma unroll 2
            for (int q = 0; q < cnt / 2; ++q) {
                const float4 v = sj[q];
                const int j = jt + 2 * q;
#pragma unroll
                for (int k = 0; k < IPT; ++k) {
                    pair_accum<CUTOFF, PE, true>(xi[k], yi[k], v.x, v.y, j != ii[k], pc, tfx[k], tfy[k], tpe[k]);
                    pair_accum<CUTOFF, PE, true>(xi[k], yi[k], v.z, v.w, (j + 1) != ii[k], pc, tfx[k], tfy[k], tpe[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            fx[k] += tfx[k]; fy[k] += tfy[k];
            if (PE) pe_acc += tpe[k];
        }
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
        if (tid == 0) __stcg(&a.pe_part[par * a.G + c], t);
    }
}

// fixed-order sum of a per-CTA partial array by one warp, in double
__device__ __forceinline__ double warp_sum_array(const float* p, int n) {
    double s = 0.0;
    for (int k = threadIdx.x & 31; k < n; k += 32) s += (double)__ldcg(&p[k]);
    return warp_sum(s);
}

template <int IPT, bool CUTOFF>
__global__ void __launch_bounds__(AP_THREADS)
ap_persistent_kernel(const ApArgs a) {
    constexpr int BI = AP_THREADS * IPT;
    __shared__ float4 sj[TILE_J / 2];
    __shared__ float  sred[AP_THREADS / 32];
    __shared__ float  s_lambda;
    const int c = blockIdx.x, tid = threadIdx.x;
    const RunCtl rc = a.rc;
    unsigned epoch = 0;

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

        // ---- [b] partial forces of R_cur ------------------------------------------------------
        if (want_pe) ap_phase_forces<IPT, CUTOFF, true >(a, Rcur, sj, sred, par);
        else         ap_phase_forces<IPT, CUTOFF, false>(a, Rcur, sj, sred, par);
        grid_barrier(a.bar, (++epoch) * (unsigned)a.G, a.err);

        // ---- [d] reduce partials, finish the velocity-Verlet step -----------------------------
        float ke_thread = 0.0f;
        for (int g = c * AP_THREADS + tid; g < a.N; g += a.G * AP_THREADS) {
            const int ib = g / BI, il = g - ib * BI;
            const int2 cc = a.iblk_ctas[ib];
            float Fx = 0.0f, Fy = 0.0f;
#pragma unroll 4
            for (int c2 = cc.x; c2 <= cc.y; ++c2) {
                const int seg = ib - a.cta_ib0[c2];
                const float2 p = __ldcg(&a.part[((size_t)c2 * a.maxseg + seg) * BI + il]);
                Fx += p.x; Fy += p.y;
            }
            const float2 r = __ldcg(&Rcur[g]);
            float2 v = (rc.nsteps > 0) ? a.Vh[g] : make_float2(0.0f, 0.0f);
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
                __stcg(&Rnext[g], make_float2(drift(r.x, v.x, a.dt, a.pc.box),      // MD:71-72
                                              drift(r.y, v.y, a.dt, a.pc.box)));
            }
        }
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
                    s_lambda = sqrtf(rc.thermo_kT / (ke / (float)a.N));
                }
            }
            __syncthreads();
            const float lam = s_lambda;
            for (int g = c * AP_THREADS + tid; g < a.N; g += a.G * AP_THREADS) {
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
        grid_barrier(a.bar, (++epoch) * (unsigned)a.G, a.err);

        // ---- energies of the post-step state (one warp, fixed order, double combine) ----------
        if (c == 0 && tid < 32 && want_pe) {
            double pe2 = warp_sum_array(a.pe_part + par * a.G, a.G);
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
}

using ApKernel = void (*)(const ApArgs);

ApKernel pick_kernel(int ipt, bool cutoff) {
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
    float*    sedges = reinterpret_cast<float*>(smem_raw);                  // nbins+1
    unsigned* shist  = reinterpret_cast<unsigned*>(sedges + nbins + 1);     // nbins
    float2*   sj     = reinterpret_cast<float2*>(shist + nbins);            // GR_TILE
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
    long long* d_cta_start = nullptr;
    int*       d_cta_ib0 = nullptr;
    int2*      d_iblk = nullptr;
    float2 *Rbuf0 = nullptr, *Rbuf1 = nullptr, *Vh = nullptr, *Ftmp = nullptr, *part = nullptr;
    float *pe_part = nullptr, *ke_part = nullptr;
    unsigned* bar = nullptr;
    int* err = nullptr;
    ApKernel kernel = nullptr;
};

int ap_create(ljmd_handle* h) {
    AllPairs* ap = new AllPairs();
    h->ap = ap;
    const long long N = h->p.N;
    ap->ipt = (N >= 2048) ? 2 : 1;
    if (const char* e = getenv("LJMD_AP_IPT")) ap->ipt = (atoi(e) == 2) ? 2 : 1;
    const int BI = AP_THREADS * ap->ipt;
    ap->nI  = (int)((N + BI - 1) / BI);
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
    long long g_own  = (N + AP_THREADS - 1) / AP_THREADS;       // enough threads to own particles once
    (void)g_own;
    ap->G = (int)std::min<long long>((long long)per_sm * h->num_sms, std::min(g_work, W));
    if (const char* e = getenv("LJMD_AP_GRID")) ap->G = std::max(1, std::min(atoi(e), ap->G));

    // stream-K split of the flattened (i-block, j-unit) space
    std::vector<long long> start(ap->G + 1);
    for (int c = 0; c <= ap->G; ++c) start[c] = (long long)(((__int128)W * c) / ap->G);
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
    LJ_CUDA(cudaMalloc(&ap->Rbuf0, sizeof(float2) * N));
    LJ_CUDA(cudaMalloc(&ap->Rbuf1, sizeof(float2) * N));
    LJ_CUDA(cudaMalloc(&ap->Vh, sizeof(float2) * N));
    LJ_CUDA(cudaMalloc(&ap->Ftmp, sizeof(float2) * N));
    LJ_CUDA(cudaMalloc(&ap->part, sizeof(float2) * (size_t)ap->G * ap->maxseg * BI));
    LJ_CUDA(cudaMalloc(&ap->pe_part, sizeof(float) * 2 * ap->G));
    LJ_CUDA(cudaMalloc(&ap->ke_part, sizeof(float) * 2 * ap->G));
    LJ_CUDA(cudaMalloc(&ap->bar, sizeof(unsigned)));
    LJ_CUDA(cudaMalloc(&ap->err, sizeof(int)));
    LJ_CUDA(cudaMemset(ap->err, 0, sizeof(int)));
    return 0;
}

void ap_destroy(ljmd_handle* h) {
    AllPairs* ap = h->ap;
    if (!ap) return;
    cudaFree(ap->d_cta_start); cudaFree(ap->d_cta_ib0); cudaFree(ap->d_iblk);
    cudaFree(ap->Rbuf0); cudaFree(ap->Rbuf1); cudaFree(ap->Vh); cudaFree(ap->Ftmp);
    cudaFree(ap->part); cudaFree(ap->pe_part); cudaFree(ap->ke_part);
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
    a.N = (int)N; a.G = ap->G; a.NJu = ap->NJu; a.nI = ap->nI; a.maxseg = ap->maxseg;
    a.dt = h->p.dt;
    a.cta_start = ap->d_cta_start; a.cta_ib0 = ap->d_cta_ib0; a.iblk_ctas = ap->d_iblk;
    a.R_in = R_in; a.Rbuf0 = ap->Rbuf0; a.Rbuf1 = ap->Rbuf1; a.Vh = ap->Vh; a.Ftmp = ap->Ftmp;
    a.part = ap->part; a.pe_part = ap->pe_part; a.ke_part = ap->ke_part;
    a.bar = ap->bar; a.err = ap->err;
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
        LJ_CUDA(cudaMemsetAsync(ap->bar, 0, sizeof(unsigned), st));
        void* args[] = {(void*)&a};
        LJ_CUDA(cudaLaunchCooperativeKernel((void*)ap->kernel, dim3(ap->G), dim3(AP_THREADS), args, 0, st));
        h->launches++;
        s = e;
    }
    if (h->timed) LJ_CUDA(cudaEventRecord(h->ev1, st));
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
