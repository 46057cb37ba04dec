// pairlaw.cu — the dense all-pairs acceleration kernel with the pair law as a template functor:
// the other two broadcast-pairwise kernels of the reference repo, gravity r^-3 in
// an open (non-periodic) plane:
//   LAW_NBODY   pairwise_forces of nbody_bh_merger_sim_single-host_workload.py (NBODY:54-67):
//               r_vec = pos[j] - pos[i]; r = sqrt(dx^2 + dy^2);
//               a_i += where(r >= 1e-6, (G m_j) / (r r r), 0) * r_vec,   j = 0..n-1 in order, j != i
//   LAW_EM3     the gravity term of acceleration() in three_particles_em_nonuni_single-host_workload.py
//               (EM3:25-38): r2 = dx^2 + dy^2 (+1 on the diagonal), clamped from below at 1e-12;
//               a_i = sum_j G m_j r_vec r2^(-3/2)   (the j = i term is r_vec = 0)
// The Lennard-Jones kernels (allpairs.cu) keep their own specialised evaluation (minimum image,
// packed FP32x2, Newton's third law); this kernel shares only the tiling: thread = particle i, the j
// particles stream through shared memory as (x, y, G m) tiles, accumulation in registers in j order
// (fixed order: bit-reproducible, and for the reference's n <= 5 exactly its summation order).
#include "ljmd_internal.cuh"

namespace ljmd {
namespace {

constexpr int PL_THREADS = 128;

struct LawNbody {   // NBODY:59-64
    static __device__ __forceinline__ void add(float dx, float dy, float gm, bool self, float& ax, float& ay) {
        const float r = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        const float r3 = __fmul_rn(__fmul_rn(r, r), r);
        const float mag = (!self && r >= 1.0e-6f) ? __fdiv_rn(gm, r3) : 0.0f;
        ax = __fadd_rn(ax, __fmul_rn(mag, dx));
        ay = __fadd_rn(ay, __fmul_rn(mag, dy));
    }
};
struct LawEm3 {     // EM3:25-38
    static __device__ __forceinline__ void add(float dx, float dy, float gm, bool self, float& ax, float& ay) {
        float r2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        if (self) r2 = __fadd_rn(r2, 1.0f);                      // + eye
        r2 = (r2 < 1.0e-12f) ? 1.0e-12f : r2;
        const float inv3 = __fdiv_rn(1.0f, __fmul_rn(r2, sqrtf(r2)));   // r2^(-3/2)
        ax = __fadd_rn(ax, __fmul_rn(__fmul_rn(gm, dx), inv3));
        ay = __fadd_rn(ay, __fmul_rn(__fmul_rn(gm, dy), inv3));
    }
};

template <class Law>
__global__ void __launch_bounds__(PL_THREADS)
pairlaw_accel_kernel(const float2* __restrict__ pos, const float* __restrict__ mass, int n, float G,
                     float2* __restrict__ acc) {
    __shared__ float sx[PL_THREADS], sy[PL_THREADS], sgm[PL_THREADS];
    const int i = blockIdx.x * PL_THREADS + threadIdx.x;
    float2 pi = make_float2(0.0f, 0.0f);
    if (i < n) pi = pos[i];
    float ax = 0.0f, ay = 0.0f;
    for (int j0 = 0; j0 < n; j0 += PL_THREADS) {
        const int j = j0 + threadIdx.x;
        __syncthreads();
        if (j < n) {
            const float2 pj = pos[j];
            sx[threadIdx.x] = pj.x; sy[threadIdx.x] = pj.y;
            sgm[threadIdx.x] = __fmul_rn(G, mass[j]);            // G * masses[j] first (NBODY:63, EM3:34)
        }
        __syncthreads();
        const int m = min(PL_THREADS, n - j0);
        if (i < n) {
#pragma unroll 4
            for (int t = 0; t < m; ++t)
                Law::add(__fsub_rn(sx[t], pi.x), __fsub_rn(sy[t], pi.y), sgm[t], j0 + t == i, ax, ay);
        }
    }
    if (i < n) acc[i] = make_float2(ax, ay);
}

}  // namespace

int pairlaw_accel(int law, const float2* pos, const float* mass, long long n, float G, float2* acc,
                  cudaStream_t stream) {
    if (!pos || !mass || !acc || n < 1 || n > (1ll << 30)) { set_error("pair-law acceleration: bad arguments"); return LJMD_E_INVALID; }
    const int grid = (int)((n + PL_THREADS - 1) / PL_THREADS);
    if (law == LJMD_LAW_GRAVITY_NBODY)
        pairlaw_accel_kernel<LawNbody><<<grid, PL_THREADS, 0, stream>>>(pos, mass, (int)n, G, acc);
    else if (law == LJMD_LAW_GRAVITY_EM3)
        pairlaw_accel_kernel<LawEm3><<<grid, PL_THREADS, 0, stream>>>(pos, mass, (int)n, G, acc);
    else { set_error("pair-law acceleration: unknown law %d", law); return LJMD_E_INVALID; }
    LJ_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ljmd
