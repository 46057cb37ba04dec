// cells.cu — sorted strip-cell path for large N (BASELINE configs 4-5).
//
// Not in the reference (it only has the dense O(N^2) form, MD:51): the pair arithmetic is the
// same as the all-pairs path (subtract, exact min-image, unfused r2, r2 < rc^2), so on identical
// fp32 inputs it evaluates exactly the pair set of the all-pairs+cutoff oracle.
//
// Geometry.  The box is cut into `nrows` horizontal rows of height >= rc + skin and every row into
// `nbx` narrow bins of width >= (rc + skin) / K  (K = 4).  Cells (row, bin) are numbered row-major
// and the particle state is kept SORTED by cell.  Everything within rc + skin of a particle of cell
// (r, b) lies in bins [b-K, b+K] of rows r-1, r, r+1, i.e. in THREE CONTIGUOUS RANGES of the sorted
// arrays: no neighbour list is stored, built or gathered; the per-step pass streams contiguous,
// vectorised position loads (9 narrow bins per row = 42 candidates/particle at rho 0.8, against 57
// for square 3x3 cells) and the rebuild is only a counting sort.
//
// Data layout in HBM (all in cell-sorted "slot" order; every row starts on a multiple of 4 slots and
// is padded with sentinel slots, so slots pair up as aligned float2 and bulk copies stay 16-byte aligned):
//   X[2], Y[2]  float   positions, structure-of-arrays, ping-pong by step (the buffer being read is
//                        never written in a step).  One 64-bit load = the x of two adjacent slots
//                        = one operand of the packed FP32x2 pipe.
//   V[2]        float2  velocities (buffers swap at a rebuild)
//   orig[2]     int32   original particle index of each slot (-1 = pad)
//   Rb          float2  positions at the last rebuild (max-displacement test against skin/2; also
//                        gives the slot's cell, hence its three candidate ranges)
//   cell_start  int32   prefix-sum cell index over nrows*nbx cells (+1)
//
// One PERSISTENT cooperative kernel runs a whole ljmd_run().  Per step ONE pass over the state and
// one grid barrier.  A row is cut into units of 64 slots = one warp (each thread owns two adjacent
// slots).  Every warp is its own producer and consumer: before it evaluates unit u it issues the TMA
// bulk copies (cp.async.bulk, mbarrier complete_tx) of everything unit u+1 needs — the x / y windows
// of rows r-1, r, r+1, the matching slices of the cell index and the unit's own Rb — into its private
// 2-stage shared-memory ring, from a per-unit copy plan that the rebuild precomputes.  The pair loop
// reads shared memory only (no load in it can miss), then velocity-Verlet + energies + displacement
// test and the new state is written: each particle's state is read once and written once per step.
// Warps never wait for each other inside a step.  Energies are per-unit partials reduced in unit
// order (deterministic).
// The rebuild (bin + histogram with rank capture -> row-structured prefix sum -> scatter ->
// deterministic in-cell order by original index + gather -> copy plans) runs inside the same kernel
// under a grid-uniform condition: no host round trip (MD:82,103).
#include "ljmd_device.cuh"

#include <algorithm>
#include <cstdio>

namespace ljmd {

namespace {

#ifndef LJMD_CELLS_UNROLL
#define LJMD_CELLS_UNROLL 2
#endif
#ifndef LJMD_CELLS_MINBLOCKS
#define LJMD_CELLS_MINBLOCKS 2
#endif
constexpr int   CL_UNROLL  = LJMD_CELLS_UNROLL;
constexpr int   CL_THREADS = 256;                  // 8 warps, each with a private TMA ring
constexpr int   CL_WARPS   = CL_THREADS / 32;
constexpr int   CL_UNIT    = 64;                   // slots per unit (one warp, two slots per thread)
constexpr int   CL_NST     = 2;                    // stages of a warp's ring
constexpr int   CL_WMAX    = 160;                  // staged window capacity per row (slots): a lattice
                                                   // row of 3 lattice lines beside one of 2 needs 1.5 units
constexpr int   CL_K       = 4;        // bins per (rc + skin)
constexpr int   CL_ORDER_MAX = 64;     // cells denser than this keep arrival order (see B5)
// Pad slots sit far outside any box, each at its OWN place (pad_x): two pads must never coincide,
// because a zero r2 would poison its partner in the shared-reciprocal evaluation (0 * inf).
// Squares and products of two squared distances stay far inside the fp32 range.
constexpr float CL_SENT    = 1.0e8f;

#ifndef LJMD_CELLS_RCP_PRODUCT
#define LJMD_CELLS_RCP_PRODUCT 0       // 1: one MUFU.RCP per TWO pairs (1/(a*b) trick), see eval_set
#endif

enum { ST_PR = 0, ST_PV = 1, ST_REBUILDS = 2, ST_ERR = 3, ST_FLAG = 4, ST_NTOT = 5, ST_NUNITS = 6,
       ST_WORDS = 8 };
enum { CERR_BARRIER = 1 };

// copy plan of one unit, precomputed by the rebuild (32 bytes)
struct __align__(16) UnitPlan {
    int row, slot0, n, direct;     // row, first slot, number of slots (even); direct = 1: windows too
                                   // large for a stage, the unit reads global memory
    int ws[3];                     // first staged slot of each window   (multiple of 4)
    int nw01;                      // staged slots of windows 0 and 1    (multiples of 4, 16 bits each)
    // (the size of window 2 travels in the upper half of `n`)
};
// candidate ranges of one thread (= two adjacent slots), precomputed by the rebuild: for each stencil
// row the slot-PAIR range [m0, m1) relative to the unit's staged window, one byte each, plus flags
//   .x = m0_0 | m1_0 << 8 | m0_1 << 16 | m1_1 << 24     .y = m0_2 | m1_2 << 8 | flags << 16
enum { PR_LIVE0 = 1, PR_LIVE1 = 2, PR_EDGE = 4 };
// one stage of a warp's ring (all 16-byte aligned for the bulk copies)
struct __align__(16) Stage {
    float    x[3][CL_WMAX];        // x window of rows r-1, r, r+1
    float    y[3][CL_WMAX];
    uint2    pr[32];               // the 32 threads' candidate ranges
    UnitPlan plan;                 // written by lane 0 before it issues the copies
};

struct CellsArgs {
    PairConsts pc;
    int   N, Nalloc, G;
    int   nrows, nbx, ncells;
    float inv_hy, inv_wx, half_skin2, dt;
    float*  X[2];
    float*  Y[2];
    float2* V[2];
    int*    orig[2];
    float2* Rb;
    float2* Fs;                     // forces held across the thermostat barrier
    int *key, *rank, *tmp, *cell_count, *cell_start, *row_tot;
    int2*     unit_tab;             // (row, first slot) of every unit
    UnitPlan* plans;                // copy plan of every unit
    int*      sched;                // [2] dynamic unit counters (by step parity)
    uint2*    pranges;              // candidate ranges of every slot pair
    int       maxunits;
    float  *pe_part, *ke_part;      // [2*maxunits] per-unit partials (by step parity)
    int*      state;                // ST_* words
    unsigned* bar;
    const float2* R_in;
    const float2* V_in;
    long long s_begin, s_end;
    RunCtl rc;
    float2 *R_out, *V_out, *F_out;
    float*  pe_out;
    int*    count_out;              // count mode: neighbour counts in original order
    int     mode;                   // 0 = run, 1 = build + count only
    long long* prof;                // optional [G][12] phase clocks (debug: LJMD_CELLS_PROF=1)
    float   count_r2;
};

__device__ __forceinline__ float pad_x(int row, int nrows) {
    return CL_SENT * (1.0f + (float)row / (float)nrows);
}

// one fp32 multiply then truncation (x >= 0); the clamp handles x == box (MD:72 closed range)
__device__ __forceinline__ int strip_coord(float x, float inv, int n) {
    const int c = (int)(x * inv);
    return max(0, min(c, n - 1));
}

// block-wide exclusive scan of one int per thread; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int block_exscan(int v, int* swarp /* CL_THREADS/32 + 1 */, int* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) swarp[w] = inc;
    __syncthreads();
    if (w == 0) {
        int t = (lane < CL_THREADS / 32) ? swarp[lane] : 0;
        int ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += u;
        }
        if (lane < CL_THREADS / 32) swarp[lane] = ti - t;   // exclusive warp offsets
        if (lane == 31) swarp[CL_THREADS / 32] = ti;
    }
    __syncthreads();
    *total = swarp[CL_THREADS / 32];
    return inc - v + swarp[w];
}

struct Ctx {
    unsigned epoch;
    unsigned uses;      // this warp's ring: units consumed so far (stage = uses % CL_NST)
    unsigned phase;     // bit s = parity of the next completion of stage s's mbarrier
    int pr, pv, ntot, nunits;
    long long* pt;      // shared-memory phase clocks (thread 0), or nullptr
};

#define CL_PROF(k)                                                          \
    do {                                                                    \
        if (ctx.pt && threadIdx.x == 0) {                                   \
            long long _t = clock64();                                       \
            ctx.pt[k] += _t - ctx.pt[11];                                   \
            ctx.pt[11] = _t;                                                \
        }                                                                   \
    } while (0)

#define CL_BARRIER() grid_barrier(a.bar, (++ctx.epoch) * (unsigned)a.G, a.state + ST_ERR)

// ---- rebuild: counting sort by cell, row-structured index, deterministic in-cell order -----------
__device__ void cells_rebuild(const CellsArgs& a, Ctx& ctx, int* sscan) {
    const int tid = threadIdx.x, gtid = blockIdx.x * CL_THREADS + tid, gsz = a.G * CL_THREADS;
    const float* __restrict__ Xc = a.X[ctx.pr];
    const float* __restrict__ Yc = a.Y[ctx.pr];
    const int* __restrict__ orig_old = a.orig[ctx.pv];
    const int ntot_old = ctx.ntot;
    CL_PROF(0);
    // B1: cell of every live slot + histogram; the atomic's return value is the slot's (arbitrary)
    //     arrival rank inside its cell, so the scatter needs no second atomic pass.
    //     (cell_count is all-zero on entry: cleared at create and again by B3 of every rebuild.)
    for (int k = gtid; k < ntot_old; k += gsz) {
        int c = -1;
        if (orig_old[k] >= 0) {
            c = strip_coord(Yc[k], a.inv_hy, a.nrows) * a.nbx + strip_coord(Xc[k], a.inv_wx, a.nbx);
            a.rank[k] = atomicAdd(&a.cell_count[c], 1);
        }
        a.key[k] = c;
    }
    CL_BARRIER();
    CL_PROF(2);
    // B2: population of every row
    for (int r = blockIdx.x; r < a.nrows; r += a.G) {
        int s = 0;
        for (int b = tid; b < a.nbx; b += CL_THREADS) s += a.cell_count[r * a.nbx + b];
        int tot;
        (void)block_exscan(s, sscan, &tot);
        if (tid == 0) a.row_tot[r] = tot;
    }
    CL_BARRIER();
    CL_PROF(3);
    // B3: row offsets (every row starts on a multiple of 4 slots) + in-row exclusive scan ->
    //     cell_start; the row is padded with sentinel slots; its counters are cleared for the next rebuild.
    {
        float* Xn = a.X[ctx.pr ^ 1];
        float* Yn = a.Y[ctx.pr ^ 1];
        int* on = a.orig[ctx.pv ^ 1];
        for (int r = blockIdx.x; r < a.nrows; r += a.G) {
            int s = 0, sc = 0;
            for (int q = tid; q < r; q += CL_THREADS) {
                const int len = (a.row_tot[q] + 3) & ~3;
                s += len;
                sc += (len + CL_UNIT - 1) / CL_UNIT;
            }
            int carry, chunk0;
            (void)block_exscan(s, sscan, &carry);
            (void)block_exscan(sc, sscan, &chunk0);
            const int rtot = a.row_tot[r];
            const int row0 = carry;
            // the row's units: consecutive runs of CL_UNIT slots, never straddling a row
            {
                const int len = (rtot + 3) & ~3, nch = (len + CL_UNIT - 1) / CL_UNIT;
                for (int j = tid; j < nch; j += CL_THREADS) a.unit_tab[chunk0 + j] = make_int2(r, row0 + j * CL_UNIT);
                if (r == a.nrows - 1 && tid == 0) a.state[ST_NUNITS] = chunk0 + nch;
            }
            for (int bb = 0; bb < a.nbx; bb += CL_THREADS) {
                const int b = bb + tid;
                int v = 0;
                if (b < a.nbx) { v = a.cell_count[r * a.nbx + b]; a.cell_count[r * a.nbx + b] = 0; }
                int tot;
                const int ex = block_exscan(v, sscan, &tot);
                if (b < a.nbx) a.cell_start[r * a.nbx + b] = carry + ex;
                carry += tot;
            }
            // pad the row to a multiple of 4 slots with sentinel slots, each at its own place
            if (tid < ((rtot + 3) & ~3) - rtot) {
                const int pad = row0 + rtot + tid;
                a.tmp[pad] = -1; on[pad] = -1;
                Xn[pad] = pad_x(r, a.nrows); Yn[pad] = CL_SENT * (1.0f + 0.125f * (float)tid);
                a.Rb[pad] = make_float2(CL_SENT, CL_SENT);
                a.V[ctx.pv ^ 1][pad] = make_float2(0.0f, 0.0f);
            }
            if (r == a.nrows - 1 && tid == 0) a.cell_start[a.ncells] = row0 + ((rtot + 3) & ~3);
        }
    }
    CL_BARRIER();
    CL_PROF(4);
    // B4: scatter source slots into their cell's range (arrival order inside a cell)
    for (int k = gtid; k < ntot_old; k += gsz) {
        const int c = a.key[k];
        if (c >= 0) a.tmp[a.cell_start[c] + a.rank[k]] = k;
    }
    CL_BARRIER();
    CL_PROF(5);
    // B5: each new slot picks the member of its cell whose ORIGINAL index has the slot's rank, so
    //     the sorted order (cell, orig) is a pure function of the positions (bit-reproducible
    //     summation order downstream), then gathers that member's state (coalesced writes).
    //     (A bin holds ~1.6 particles at liquid density; a bin with more than CL_ORDER_MAX members
    //     keeps its arrival order: still correct, no longer run-to-run bit-reproducible.)
    {
        const int ntot_new = a.cell_start[a.ncells];
        float* Xn = a.X[ctx.pr ^ 1];
        float* Yn = a.Y[ctx.pr ^ 1];
        const float2* Vo = a.V[ctx.pv];
        float2* Vn = a.V[ctx.pv ^ 1];
        int* on = a.orig[ctx.pv ^ 1];
        for (int d = gtid; d < ntot_new; d += gsz) {
            const int k = a.tmp[d];
            if (k < 0) continue;                                  // pad slot (initialised in B3)
            const int c = a.key[k];
            const int b = a.cell_start[c], n = a.cell_start[c + 1] - b, p = d - b;
            int ksel = k;
            if (n > 1 && n <= CL_ORDER_MAX) {
                for (int m = 0; m < n; ++m) {
                    const int km = a.tmp[b + m];
                    if (km < 0) continue;                         // the row's pad sits in its last bin
                    const int om = orig_old[km];
                    int rk = 0;
                    for (int q = 0; q < n; ++q) {
                        const int kq = a.tmp[b + q];
                        rk += (kq >= 0 && orig_old[kq] < om);
                    }
                    if (rk == p) { ksel = km; break; }
                }
            }
            const float x = Xc[ksel], y = Yc[ksel];
            Xn[d] = x; Yn[d] = y;
            a.Rb[d] = make_float2(x, y);
            Vn[d] = Vo[ksel];
            on[d] = orig_old[ksel];
        }
        ctx.ntot = ntot_new;
        ctx.nunits = __ldcg(a.state + ST_NUNITS);
    }
    ctx.pr ^= 1;
    ctx.pv ^= 1;
    if (blockIdx.x == 0 && tid == 0) a.state[ST_REBUILDS] += 1;
    CL_BARRIER();
    CL_PROF(6);
    // B6: per unit (one warp each) the copy plan, per thread of the unit its candidate ranges: both
    //     depend only on the build-time structure, so the per-step pass does no index arithmetic.
    //     Window k = bins [bf-K, bl+K] of row r+k-1 (periodic in r, clipped in x), aligned outwards
    //     to 4 slots; a thread's range k = bins [b0-K, b1+K] of that row, as slot pairs relative to
    //     the window.  Edge threads (wrapped ranges) and oversized units are flagged instead.
    {
        const int* __restrict__ cs = a.cell_start;
        const int lane = tid & 31;
        const int W = a.G * CL_WARPS, gw = blockIdx.x * CL_WARPS + (tid >> 5);
        for (int u = gw; u < ctx.nunits; u += W) {
            const int2 tab = a.unit_tab[u];
            const int row = tab.x, slot0 = tab.y;
            const int rend = cs[(row + 1) * a.nbx];                    // next row's first slot (even)
            const int n = min(CL_UNIT, rend - slot0);
            const int bf = strip_coord(a.Rb[slot0].x, a.inv_wx, a.nbx);
            const int bl = strip_coord(a.Rb[slot0 + n - 1].x, a.inv_wx, a.nbx);   // a pad clamps to the last bin
            const int lo = max(bf - CL_K, 0), hi = min(bl + CL_K, a.nbx - 1);
            int ws[3], nw[3];
            bool direct = false;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                int rr = row + k - 1;
                rr += (rr < 0) ? a.nrows : 0;
                rr -= (rr >= a.nrows) ? a.nrows : 0;
                ws[k] = cs[rr * a.nbx + lo] & ~3;
                nw[k] = ((cs[rr * a.nbx + hi + 1] + 3) & ~3) - ws[k];
                direct |= (nw[k] > CL_WMAX);
            }
            if (lane == 0) {
                UnitPlan pl;
                pl.row = row; pl.slot0 = slot0; pl.direct = direct ? 1 : 0;
                pl.n = n | (direct ? 0 : nw[2] << 16);
                pl.ws[0] = ws[0]; pl.ws[1] = ws[1]; pl.ws[2] = ws[2];
                pl.nw01 = direct ? 0 : (nw[0] | nw[1] << 16);
                a.plans[u] = pl;
            }
            // this lane's slot pair
            const int p = (slot0 >> 1) + lane;
            if (2 * lane < n) {
                const float4 rb = reinterpret_cast<const float4*>(a.Rb)[p];
                const bool live0 = rb.x < 0.5f * CL_SENT, live1 = rb.z < 0.5f * CL_SENT;
                unsigned flags = (live0 ? PR_LIVE0 : 0) | (live1 ? PR_LIVE1 : 0);
                uint2 w = make_uint2(0u, 0u);
                if (live0) {
                    const int b0 = strip_coord(rb.x, a.inv_wx, a.nbx);
                    const int b1 = live1 ? strip_coord(rb.z, a.inv_wx, a.nbx) : b0;
                    const bool edge = (row == 0) | (row == a.nrows - 1) | (b0 < CL_K) | (b1 > a.nbx - 1 - CL_K);
                    if (edge) {
                        flags |= PR_EDGE;
                    } else if (!direct) {
                        unsigned m[6];
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const int s = cs[(row + k - 1) * a.nbx + b0 - CL_K];
                            const int e = cs[(row + k - 1) * a.nbx + b1 + CL_K + 1];
                            m[2 * k]     = (unsigned)((s >> 1) - (ws[k] >> 1));
                            m[2 * k + 1] = (unsigned)(((e + 1) >> 1) - (ws[k] >> 1));
                        }
                        w.x = m[0] | m[1] << 8 | m[2] << 16 | m[3] << 24;
                        w.y = m[4] | m[5] << 8;
                    }
                }
                w.y |= flags << 16;
                a.pranges[p] = w;
            }
        }
    }
    CL_BARRIER();
    CL_PROF(7);
}

// ---- packed pair evaluation ------------------------------------------------------------------------
// A thread owns the adjacent slots i0 = 2p, i1 = 2p+1 (packed as xi2 = (x_i0, x_i1)).  One iteration
// loads the x and y of the two adjacent candidate slots j0 = 2m, j1 = 2m+1 as two 64-bit loads and
// evaluates FOUR pairs on the packed FP32x2 pipe with no register shuffling:
//   "straight" set (i0,j0),(i1,j1) from xi2, "swapped" set (i1,j0),(i0,j1) from the swapped copy
//   (x_i1, x_i0) (an operand swizzle of FADD2, not a register move).
// d is formed as xj - xi (= -(xi - xj) exactly), so the accumulators hold -F bit for bit.
//   r2 = dx*dx + dy*dy unfused (fma(t, 1, u) with a run-time 1: see pair2_accum) => the pair set
//   {r2 < rc2} is the oracle's.
//   LJMD_CELLS_RCP_PRODUCT: 1/r2 of the two pairs of a set from ONE MUFU.RCP of the product
//   (ir2_a = r2_b * rcp(r2_a * r2_b)): MUFU issues at 1/8 rate and would otherwise co-limit the loop.
// KEEP:   per-pair keep predicates are applied (self pair of the own row; exact bounds of edge ranges).
// MINIMG: exact minimum image on every displacement (edge warps).
// The cutoff is a select on the ALU pipe (FSETP + SEL): the FMA pipe is the one that saturates (a
// packed FP32x2 instruction occupies it for two cycles), so nothing that can run elsewhere is put on
// it; the select also discards the NaN of a masked 0 * inf.
template <bool PE, bool MINIMG, bool KEEP>
__device__ __forceinline__ void eval_set(const PairConsts& pc, const PairConsts2& c2, float2 nxi,
                                         float2 nyi, float2 xj, float2 yj, bool keepx, bool keepy,
                                         float2& ax, float2& ay, float2& pe2) {
    float2 dx = __fadd2_rn(xj, nxi);
    float2 dy = __fadd2_rn(yj, nyi);
    if (MINIMG) {
        dx.x = min_image(dx.x, pc.box, pc.timg); dx.y = min_image(dx.y, pc.box, pc.timg);
        dy.x = min_image(dy.x, pc.box, pc.timg); dy.y = min_image(dy.y, pc.box, pc.timg);
    }
    const float2 r2 = __ffma2_rn(__fmul2_rn(dx, dx), c2.one, __fmul2_rn(dy, dy));
    float2 ir2;
#if LJMD_CELLS_RCP_PRODUCT
    const float rp = rcp_approx(__fmul_rn(r2.x, r2.y));
    ir2 = __fmul2_rn(make_float2(r2.y, r2.x), make_float2(rp, rp));
#else
    ir2 = make_float2(rcp_approx(r2.x), rcp_approx(r2.y));
#endif
    if (KEEP) {
        ir2.x = (keepx & (r2.x < pc.rc2)) ? ir2.x : 0.0f;
        ir2.y = (keepy & (r2.y < pc.rc2)) ? ir2.y : 0.0f;
    } else {
        ir2.x = (r2.x < pc.rc2) ? ir2.x : 0.0f;
        ir2.y = (r2.y < pc.rc2) ? ir2.y : 0.0f;
    }
    const float2 ir6 = __fmul2_rn(__fmul2_rn(ir2, ir2), ir2);
    const float2 f = __fmul2_rn(__ffma2_rn(ir6, c2.c12, c2.nc6), __fmul2_rn(ir6, ir2));
    ax = __ffma2_rn(f, dx, ax);
    ay = __ffma2_rn(f, dy, ay);
    if (PE) pe2 = __ffma2_rn(ir6, __ffma2_rn(ir6, c2.d12, c2.nd6), pe2);
}

struct PairAcc {
    float2 axA, ayA, axB, ayB, peA, peB;
};

__device__ __forceinline__ void prefetch_l1(const void* p) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// interior: slots [2*m0, 2*m1) (range aligned outwards to slot pairs: the extra slots belong to the
// same row and to bins outside the stencil, hence beyond rc, and to no other range of this thread).
// pself = the thread's own slot pair if this is its own row, else -1.  ONE instance of this loop
// serves all three rows (it is the hot code: it has to stay inside the instruction cache).
template <bool PE>
__device__ __forceinline__ void row_range(const PairConsts& pc, const PairConsts2& c2,
                                          const float2* __restrict__ X2, const float2* __restrict__ Y2,
                                          int m0, int m1, int pself, float2 nxi, float2 nyi, float2 nxs,
                                          float2 nys, PairAcc& acc) {
#pragma unroll CL_UNROLL
    for (int m = m0; m < m1; ++m) {
        const float2 xj = X2[m], yj = Y2[m];
        const bool k = (m != pself);
        eval_set<PE, false, true >(pc, c2, nxi, nyi, xj, yj, k, k, acc.axA, acc.ayA, acc.peA);
        eval_set<PE, false, false>(pc, c2, nxs, nys, xj, yj, true, true, acc.axB, acc.ayB, acc.peB);
    }
}

// edge: slots [s, e) exactly, minimum image, any candidate may be one of the thread's own slots
// (rare path: kept out of line, generic pointers — the window may be shared or global memory)
template <bool PE>
__device__ __noinline__ void edge_range(const PairConsts& pc, const PairConsts2& c2,
                                        const float2* X2, const float2* Y2,
                                        int s, int e, int p, float2 nxi, float2 nyi, float2 nxs,
                                        float2 nys, PairAcc& acc) {
    for (int m = s >> 1; m < ((e + 1) >> 1); ++m) {
        const float2 xj = X2[m], yj = Y2[m];
        const bool v0 = (2 * m >= s), v1 = (2 * m + 1 < e), own = (m == p);
        eval_set<PE, true, true>(pc, c2, nxi, nyi, xj, yj, v0 & !own, v1 & !own, acc.axA, acc.ayA, acc.peA);
        eval_set<PE, true, true>(pc, c2, nxs, nys, xj, yj, v0, v1, acc.axB, acc.ayB, acc.peB);
    }
}

// ---- scalar path (count mode only) ---------------------------------------------------------------
// One particle against bins [b-K, b+K] of rows r-1, r, r+1 with periodic wrap of both indices: up to
// two pieces per row, exact bounds, scalar arithmetic with the exact minimum image.
template <bool PE, bool COUNT>
__device__ __forceinline__ void generic_particle(const CellsArgs& a, const float* __restrict__ X,
                                                 const float* __restrict__ Y, int i, float xi, float yi,
                                                 int r, int b, float lim2, float& fx, float& fy,
                                                 float& pe, int& cnt) {
    const PairConsts pc = a.pc;
    const int* __restrict__ cs = a.cell_start;
    for (int dr = -1; dr <= 1; ++dr) {
        int rr = r + dr;
        rr += (rr < 0) ? a.nrows : 0;
        rr -= (rr >= a.nrows) ? a.nrows : 0;
        const int lo = b - CL_K, hi = b + CL_K;
        for (int piece = 0; piece < 3; ++piece) {
            int bl, bh;
            if (piece == 0)      { bl = max(lo, 0); bh = min(hi, a.nbx - 1); }
            else if (piece == 1) { if (lo >= 0) continue; bl = lo + a.nbx; bh = a.nbx - 1; }
            else                 { if (hi < a.nbx) continue; bl = 0; bh = hi - a.nbx; }
            const int s = cs[rr * a.nbx + bl], e = cs[rr * a.nbx + bh + 1];
            for (int j = s; j < e; ++j) {
                if (j == i) continue;
                const float xj = X[j], yj = Y[j];
                if (COUNT) {
                    const float dx = min_image(__fsub_rn(xi, xj), pc.box, pc.timg);
                    const float dy = min_image(__fsub_rn(yi, yj), pc.box, pc.timg);
                    const float r2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                    cnt += (r2 < lim2);
                } else {
                    pair_accum<true, PE, false>(xi, yi, xj, yj, true, pc, fx, fy, pe);
                }
            }
        }
    }
}

// velocity-Verlet epilogue of one particle (MD:70-74) + outputs; returns "left the skin/2 ball"
struct StepFlags {
    bool kick1, final, want_e, thermo, sample;
    long long s;
};

__device__ __forceinline__ bool finish_particle(const CellsArgs& a, const StepFlags& f, int slot, int o,
                                                float rx, float ry, float Fx, float Fy, float2& v,
                                                float2 rb, float& xn, float& yn, float& ke_thread) {
    const RunCtl& rc = a.rc;
    xn = rx; yn = ry;
    if (f.kick1) { v.x = kick(v.x, Fx, a.dt); v.y = kick(v.y, Fy, a.dt); }                      // MD:74
    if (f.want_e || f.thermo) ke_thread += v.x * v.x + v.y * v.y;
    if (f.sample) rc.traj[(size_t)(f.s / rc.sample_every) * a.N + o] = make_float2(rx, ry);      // MD:93-100
    if (f.thermo) {
        a.Fs[slot] = make_float2(Fx, Fy);
        return false;
    }
    if (f.final) {
        if (a.R_out) a.R_out[o] = make_float2(rx, ry);
        if (a.V_out) a.V_out[o] = v;
        if (a.F_out) a.F_out[o] = make_float2(Fx, Fy);
        return false;
    }
    v.x = kick(v.x, Fx, a.dt); v.y = kick(v.y, Fy, a.dt);                                        // MD:70
    xn = drift(rx, v.x, a.dt, a.pc.box);                                                         // MD:71-72
    yn = drift(ry, v.y, a.dt, a.pc.box);
    const float ddx = min_image(__fsub_rn(xn, rb.x), a.pc.box, a.pc.timg);
    const float ddy = min_image(__fsub_rn(yn, rb.y), a.pc.box, a.pc.timg);
    return ddx * ddx + ddy * ddy > a.half_skin2;
}

// ---- mbarrier / TMA bulk-copy primitives (sm_90+ PTX; SASS: SYNCS.*, UBLKCP) ------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (bytes and both addresses multiples of 16), completion on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// order generic-proxy accesses (the previous step's st.global, made visible by the grid barrier)
// against the async proxy (the bulk copies that read them)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// sum of n floats in fixed order by the whole CTA, in double; result valid in every thread
__device__ __forceinline__ double block_sum_array(const float* p, int n, double* sdbl /* CL_THREADS/32 */) {
    double t = 0.0;
    for (int k = threadIdx.x; k < n; k += CL_THREADS) t += (double)__ldcg(p + k);
    t = warp_sum(t);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sdbl[threadIdx.x >> 5] = t;
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < CL_THREADS / 32; ++k) tot += sdbl[k];
    return tot;
}

// ---- lane 0: issue the bulk copies of one unit into a stage of the warp's ring ---------------------
struct PlanRegs { int4 q0, q1; };
__device__ __forceinline__ PlanRegs load_plan(const CellsArgs& a, int u) {
    const int4* __restrict__ pg = reinterpret_cast<const int4*>(a.plans + u);
    PlanRegs q;
    q.q0 = pg[0]; q.q1 = pg[1];
    return q;
}
__device__ __forceinline__ void issue_unit(const CellsArgs& a, const Ctx& ctx, const PlanRegs& q, Stage& st,
                                           unsigned long long* full) {
    int4* ps = reinterpret_cast<int4*>(&st.plan);
    ps[0] = q.q0; ps[1] = q.q1;
    if (q.q0.w) return;                                  // direct unit: nothing staged
    const float* __restrict__ X = a.X[ctx.pr];
    const float* __restrict__ Y = a.Y[ctx.pr];
    const int ws[3] = {q.q1.x, q.q1.y, q.q1.z};
    const int nw[3] = {q.q1.w & 0xffff, (int)((unsigned)q.q1.w >> 16), (int)((unsigned)q.q0.z >> 16)};
    const int slot0 = q.q0.y;
    unsigned bytes = 32u * 8u;
#pragma unroll
    for (int k = 0; k < 3; ++k) bytes += (unsigned)nw[k] * 8u;
    mbar_arrive_expect_tx(full, bytes);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (nw[k] > 0) {
            bulk_g2s(st.x[k], X + ws[k], (unsigned)nw[k] * 4u, full);
            bulk_g2s(st.y[k], Y + ws[k], (unsigned)nw[k] * 4u, full);
        }
    }
    bulk_g2s(st.pr, a.pranges + (slot0 >> 1), 32u * 8u, full);
}

// dynamic schedule: lane 0 draws the next unit from a per-step counter (one atomic per unit, issued
// two units ahead of its use); -1 = no more units
__device__ __forceinline__ int next_unit(const CellsArgs& a, const Ctx& ctx, int par) {
    const int u = atomicAdd(&a.sched[par], 1);
    return (u < ctx.nunits) ? u : -1;
}

// ---- forces + integrate of one unit (one warp) ------------------------------------------------------
template <bool PE, bool STAGED>
__device__ __forceinline__ void unit_compute(const CellsArgs& a, const Ctx& ctx, const StepFlags& fl,
                                             const Stage& st, float& pe_thread, float& ke_thread,
                                             int& moved) {
    const UnitPlan& m = st.plan;                       // (in shared memory)
    const int t = threadIdx.x & 31;
    const PairConsts pc = a.pc;
    const PairConsts2 c2 = make_pair_consts2(pc);
    const float2* __restrict__ X2g = reinterpret_cast<const float2*>(a.X[ctx.pr]);
    const float2* __restrict__ Y2g = reinterpret_cast<const float2*>(a.Y[ctx.pr]);
    float2* Xn2 = reinterpret_cast<float2*>(a.X[ctx.pr ^ 1]);
    float2* Yn2 = reinterpret_cast<float2*>(a.Y[ctx.pr ^ 1]);
    float4* V4 = reinterpret_cast<float4*>(a.V[ctx.pv]);
    const int* __restrict__ csg = a.cell_start;
    const int  slot0 = m.slot0, nslots = m.n & 0xffff, r = m.row;
    const int  p = (slot0 >> 1) + t;                   // global slot-pair index
    const bool act = 2 * t < nslots;
    float2 xi = make_float2(CL_SENT, CL_SENT), yi = xi;
    float4 rb = make_float4(CL_SENT, CL_SENT, CL_SENT, CL_SENT);
    float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    int2 o = make_int2(-1, -1);
    uint2 pr = make_uint2(0u, 0u);
    if (act) {
        // build positions, velocities and original indices are only needed by the epilogue: plain
        // coalesced loads, issued now, consumed after the pair loop
        rb = reinterpret_cast<const float4*>(a.Rb)[p];
        o = reinterpret_cast<const int2*>(a.orig[ctx.pv])[p];
        if (a.rc.nsteps > 0) v = V4[p];
        if (STAGED) {
            const int q = p - (m.ws[1] >> 1);
            pr = st.pr[t];
            xi = reinterpret_cast<const float2*>(st.x[1])[q];
            yi = reinterpret_cast<const float2*>(st.y[1])[q];
        } else {
            pr = a.pranges[p];
            xi = X2g[p]; yi = Y2g[p];
        }
    }
    const unsigned flags = pr.y >> 16;
    const bool live0 = (flags & PR_LIVE0) != 0, live1 = (flags & PR_LIVE1) != 0;
    const bool wedge = __any_sync(0xffffffffu, (flags & PR_EDGE) != 0);
    PairAcc acc;
    acc.axA = acc.ayA = acc.axB = acc.ayB = acc.peA = acc.peB = make_float2(0.0f, 0.0f);
    const float2 nxi = make_float2(-xi.x, -xi.y), nyi = make_float2(-yi.x, -yi.y);
    const float2 nxs = make_float2(-xi.y, -xi.x), nys = make_float2(-yi.y, -yi.x);
    if (STAGED && !wedge) {
        // the common case: three precomputed ranges, straight out of the staged windows
        if (live0) {
            const unsigned long long w = (unsigned long long)pr.x | ((unsigned long long)(pr.y & 0xffffu) << 32);
#pragma unroll 1
            for (int k = 0; k < 3; ++k) {       // rows r-1, r, r+1: one contiguous range each
                const int m0 = (int)((w >> (16 * k)) & 0xffu), m1 = (int)((w >> (16 * k + 8)) & 0xffu);
                const int pself = (k == 1) ? p - (m.ws[1] >> 1) : -1;
                row_range<PE>(pc, c2, reinterpret_cast<const float2*>(st.x[k]),
                              reinterpret_cast<const float2*>(st.y[k]), m0, m1, pself, nxi, nyi, nxs, nys, acc);
            }
        }
    } else if (live0) {
        // edge warps and oversized units: ranges from the global cell index, positions from global
        const int b0 = strip_coord(rb.x, a.inv_wx, a.nbx);
        const int b1 = live1 ? strip_coord(rb.z, a.inv_wx, a.nbx) : b0;
        if (!wedge) {
#pragma unroll 1
            for (int k = 0; k < 3; ++k) {
                const int s = csg[(r + k - 1) * a.nbx + b0 - CL_K], e = csg[(r + k - 1) * a.nbx + b1 + CL_K + 1];
                row_range<PE>(pc, c2, X2g, Y2g, s >> 1, (e + 1) >> 1, (k == 1) ? p : -1, nxi, nyi, nxs, nys, acc);
            }
        } else {
            // bins [b0-K, b1+K] of rows r-1, r, r+1 with periodic wrap of both indices: one piece plus
            // up to two wrapped pieces per row (the whole row once if the union would overlap itself)
            int lo = b0 - CL_K, hi = b1 + CL_K;
            if (hi - lo + 1 > a.nbx) { lo = 0; hi = a.nbx - 1; }
#pragma unroll 1
            for (int k = 0; k < 3; ++k) {
                int rr = r + k - 1;
                rr += (rr < 0) ? a.nrows : 0;
                rr -= (rr >= a.nrows) ? a.nrows : 0;
                const int* __restrict__ csr = csg + rr * a.nbx;
                edge_range<PE>(pc, c2, X2g, Y2g, csr[max(lo, 0)], csr[min(hi, a.nbx - 1) + 1], p, nxi, nyi, nxs, nys, acc);
                if (lo < 0)      edge_range<PE>(pc, c2, X2g, Y2g, csr[lo + a.nbx], csr[a.nbx], p, nxi, nyi, nxs, nys, acc);
                if (hi >= a.nbx) edge_range<PE>(pc, c2, X2g, Y2g, csr[0], csr[hi - a.nbx + 1], p, nxi, nyi, nxs, nys, acc);
            }
        }
    }
    // accumulators hold -F:  i0 <- straight.x + swapped.y,  i1 <- straight.y + swapped.x
    const float F0x = -(acc.axA.x + acc.axB.y), F0y = -(acc.ayA.x + acc.ayB.y);
    const float F1x = -(acc.axA.y + acc.axB.x), F1y = -(acc.ayA.y + acc.ayB.x);
    if (PE) pe_thread += (acc.peA.x + acc.peB.y) + (acc.peA.y + acc.peB.x);
    if (!act) return;
    // the epilogue operands were requested before the pair loop; this keeps the compiler from
    // hoisting their first use (e.g. the sign extension of an index) up to the load
    asm volatile("" : "+r"(o.x), "+r"(o.y), "+f"(rb.x), "+f"(v.x) :: "memory");
    float2 v0 = make_float2(v.x, v.y), v1 = make_float2(v.z, v.w);
    float2 xn = xi, yn = yi;
    if (live0) moved |= finish_particle(a, fl, 2 * p, o.x, xi.x, yi.x, F0x, F0y, v0, make_float2(rb.x, rb.y), xn.x, yn.x, ke_thread);
    if (live1) moved |= finish_particle(a, fl, 2 * p + 1, o.y, xi.y, yi.y, F1x, F1y, v1, make_float2(rb.z, rb.w), xn.y, yn.y, ke_thread);
    if (a.rc.nsteps > 0 && !(fl.final && !fl.thermo)) {
        V4[p] = make_float4(v0.x, v0.y, v1.x, v1.y);
        if (!fl.thermo) { Xn2[p] = xn; Yn2[p] = yn; }
    }
}

// ---- one step's pass of one warp --------------------------------------------------------------------
// Software pipeline, three units deep, driven by lane 0: while unit u is evaluated, the bulk copies of
// the next unit are in flight, the copy plan of the one after is being loaded, and the index of the
// one after that is being drawn from the step's counter.  Which warp evaluates which unit does not
// affect any result (energies are per-unit partials).
template <bool PE>
__device__ __forceinline__ void warp_force_pass(const CellsArgs& a, Ctx& ctx, const StepFlags& fl,
                                                Stage* ring /* this warp's CL_NST stages */,
                                                unsigned long long* full /* this warp's CL_NST barriers */,
                                                int par, bool want_ke, int& moved) {
    const int lane = threadIdx.x & 31;
    int u = -1, un = -1, unn = -1, unnn = -1;            // meaningful in lane 0 only
    PlanRegs qn;
    qn.q0 = qn.q1 = make_int4(0, 0, 0, 0);
    if (lane == 0) {
        fence_proxy_async();
        u = next_unit(a, ctx, par);
        un = (u >= 0) ? next_unit(a, ctx, par) : -1;
        unn = (un >= 0) ? next_unit(a, ctx, par) : -1;
        if (u >= 0) {
            const PlanRegs q = load_plan(a, u);
            issue_unit(a, ctx, q, ring[ctx.uses % CL_NST], &full[ctx.uses % CL_NST]);
        }
        if (un >= 0) qn = load_plan(a, un);
    }
    for (;;) {
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u < 0) break;
        const int s = (int)(ctx.uses % CL_NST);
        ++ctx.uses;
        __syncwarp();                                    // every lane is done with the other stage
        if (lane == 0) {
            if (unn >= 0) unnn = next_unit(a, ctx, par); // consumed two units from now
            if (un >= 0) issue_unit(a, ctx, qn, ring[s ^ 1], &full[s ^ 1]);
            if (unn >= 0) qn = load_plan(a, unn);        // consumed one unit from now
        }
        __syncwarp();                                    // plan of unit u (written by lane 0) is visible
        const Stage& st = ring[s];
        const bool direct = st.plan.direct != 0;
        float pe_thread = 0.0f, ke_thread = 0.0f;
        if (direct) {
            unit_compute<true, false>(a, ctx, fl, st, pe_thread, ke_thread, moved);
        } else {
            mbar_wait(&full[s], (ctx.phase >> s) & 1u);
            ctx.phase ^= 1u << s;
            unit_compute<PE, true>(a, ctx, fl, st, pe_thread, ke_thread, moved);
        }
        // per-unit energy partials (fixed shuffle tree), reduced in unit order after the barrier
        if (PE) {
            const float tsum = warp_sum(pe_thread);
            if (lane == 0) __stcg(&a.pe_part[par * a.maxunits + u], tsum);
        }
        if (want_ke) {
            const float tsum = warp_sum(ke_thread);
            if (lane == 0) __stcg(&a.ke_part[par * a.maxunits + u], tsum);
        }
        u = un; un = unn; unn = unnn; unnn = -1;
    }
}

extern __shared__ __align__(128) unsigned char cells_smem[];

__global__ void __launch_bounds__(CL_THREADS, LJMD_CELLS_MINBLOCKS)
cells_persistent_kernel(const CellsArgs a) {
    __shared__ int    sscan[CL_THREADS / 32 + 1];
    __shared__ double sdbl[CL_THREADS / 32];
    __shared__ float  s_lambda;
    __shared__ long long s_pt[12];
    __shared__ __align__(8) unsigned long long s_full[CL_WARPS][CL_NST];
    Stage* stages = reinterpret_cast<Stage*>(cells_smem);
    const int tid = threadIdx.x, gtid = blockIdx.x * CL_THREADS + tid, gsz = a.G * CL_THREADS;
    const int warp = tid >> 5;
    const RunCtl rc = a.rc;
    Ctx ctx;
    ctx.epoch = 0;
    ctx.uses = 0;
    ctx.phase = 0;
    ctx.pt = a.prof ? s_pt : nullptr;
    if (tid == 0) {
        if (ctx.pt) { for (int k = 0; k < 11; ++k) s_pt[k] = 0; s_pt[11] = clock64(); }
        for (int w = 0; w < CL_WARPS; ++w)
            for (int k = 0; k < CL_NST; ++k) mbar_init(&s_full[w][k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    ctx.pr = a.state[ST_PR];
    ctx.pv = a.state[ST_PV];
    ctx.ntot = a.state[ST_NTOT];
    ctx.nunits = a.state[ST_NUNITS];

    if (a.s_begin < 0) {
        // load the caller's state (original order, no pads) and sort it
        for (int i = gtid; i < a.N; i += gsz) {
            const float2 r = a.R_in[i];
            a.X[ctx.pr][i] = r.x;
            a.Y[ctx.pr][i] = r.y;
            a.V[ctx.pv][i] = a.V_in ? a.V_in[i] : make_float2(0.0f, 0.0f);
            a.orig[ctx.pv][i] = i;
        }
        ctx.ntot = a.N;
        CL_BARRIER();
        cells_rebuild(a, ctx, sscan);
        if (a.mode == 1) {
            // neighbour recount through the same ranges (scalar path), original order out
            const float* X = a.X[ctx.pr];
            const float* Y = a.Y[ctx.pr];
            const int* og = a.orig[ctx.pv];
            for (int i = gtid; i < ctx.ntot; i += gsz) {
                const int o = og[i];
                if (o < 0) continue;
                const float xi = X[i], yi = Y[i];
                float fx = 0.0f, fy = 0.0f, pe = 0.0f;
                int cnt = 0;
                generic_particle<false, true>(a, X, Y, i, xi, yi, strip_coord(yi, a.inv_hy, a.nrows),
                                              strip_coord(xi, a.inv_wx, a.nbx), a.count_r2, fx, fy, pe, cnt);
                a.count_out[o] = cnt;
            }
            if (gtid == 0) { a.state[ST_PR] = ctx.pr; a.state[ST_PV] = ctx.pv; a.state[ST_NTOT] = ctx.ntot; }
            return;
        }
    }

    for (long long s = a.s_begin; s < a.s_end; ++s) {
        const int par = (int)((s + 1) & 1);
        StepFlags fl;
        fl.s      = s;
        fl.kick1  = (s >= 0);
        fl.final  = (s == rc.nsteps - 1);
        fl.want_e = fl.kick1 && rc.energy_every > 0 && (s % rc.energy_every == 0);
        const bool want_pe = fl.want_e || (rc.nsteps == 0 && a.pe_out != nullptr);
        fl.thermo = fl.kick1 && rc.thermo_every > 0 && rc.thermo_kT > 0.0f &&
                    ((s + 1) % rc.thermo_every == 0);
        fl.sample = fl.kick1 && rc.sample_every > 0 && (s % rc.sample_every == 0) &&
                    (s / rc.sample_every < rc.S);
        // rebuild requested by the previous step's displacement test?
        if (s > a.s_begin || a.s_begin >= 0) {
            if (__ldcg(a.state + ST_FLAG) == (int)(s + 1)) cells_rebuild(a, ctx, sscan);
        }
        // the other parity's unit counter is idle during this step: clear it for the next one
        if (gtid == 0) __stcg(&a.sched[par ^ 1], 0);
        int moved = 0;
        {
            const bool want_ke = fl.want_e || fl.thermo;
            Stage* ring = stages + warp * CL_NST;
            if (want_pe) warp_force_pass<true >(a, ctx, fl, ring, s_full[warp], par, want_ke, moved);
            else         warp_force_pass<false>(a, ctx, fl, ring, s_full[warp], par, want_ke, moved);
        }
        if (fl.thermo) {
            CL_BARRIER();
            const double ke2 = block_sum_array(a.ke_part + par * a.maxunits, ctx.nunits, sdbl);
            if (tid == 0) s_lambda = sqrtf(rc.thermo_kT / ((float)(0.5 * ke2) / (float)a.N));
            __syncthreads();
            const float lam = s_lambda;
            const float* X = a.X[ctx.pr];
            const float* Y = a.Y[ctx.pr];
            float* Xn = a.X[ctx.pr ^ 1];
            float* Yn = a.Y[ctx.pr ^ 1];
            float2* V = a.V[ctx.pv];
            const int* og = a.orig[ctx.pv];
            for (int i = gtid; i < ctx.ntot; i += gsz) {
                const int o = og[i];
                if (o < 0) { if (!fl.final) { Xn[i] = X[i]; Yn[i] = Y[i]; } continue; }   // pad stays put
                const float rx = X[i], ry = Y[i];
                const float2 F = a.Fs[i];
                float2 v = V[i];
                v.x *= lam; v.y *= lam;
                if (fl.final) {
                    if (a.R_out) a.R_out[o] = make_float2(rx, ry);
                    if (a.V_out) a.V_out[o] = v;
                    if (a.F_out) a.F_out[o] = F;
                } else {
                    v.x = kick(v.x, F.x, a.dt); v.y = kick(v.y, F.y, a.dt);
                    V[i] = v;
                    const float xn = drift(rx, v.x, a.dt, a.pc.box), yn = drift(ry, v.y, a.dt, a.pc.box);
                    Xn[i] = xn; Yn[i] = yn;
                    const float2 rb = a.Rb[i];
                    const float ddx = min_image(__fsub_rn(xn, rb.x), a.pc.box, a.pc.timg);
                    const float ddy = min_image(__fsub_rn(yn, rb.y), a.pc.box, a.pc.timg);
                    moved |= (ddx * ddx + ddy * ddy > a.half_skin2);
                }
            }
        }
        // a particle left the skin/2 ball: ask for a rebuild before the next force evaluation.
        // The flag carries the step stamp, so it never needs clearing (no reset race).
        if (__syncthreads_or(moved) && tid == 0) __stcg(a.state + ST_FLAG, (int)(s + 2));
        if (!fl.final) ctx.pr ^= 1;
        CL_PROF(0);
        CL_BARRIER();
        CL_PROF(1);

        if (blockIdx.x == 0 && want_pe) {
            const double pe2 = block_sum_array(a.pe_part + par * a.maxunits, ctx.nunits, sdbl);
            double ke2 = 0.0;
            if (fl.want_e) ke2 = block_sum_array(a.ke_part + par * a.maxunits, ctx.nunits, sdbl);
            if (tid == 0) {
                if (fl.want_e) {
                    float* o = rc.ke_pe + 2 * (s / rc.energy_every);
                    o[0] = (float)(0.5 * ke2);
                    o[1] = (float)(0.5 * pe2);
                } else {
                    a.pe_out[0] = (float)(0.5 * pe2);
                }
            }
        }
    }
    if (gtid == 0) {
        a.state[ST_PR] = ctx.pr; a.state[ST_PV] = ctx.pv; a.state[ST_NTOT] = ctx.ntot;
    }
    if (ctx.pt && tid == 0)
        for (int k = 0; k < 11; ++k) a.prof[blockIdx.x * 12 + k] = s_pt[k];
}

// ---- stand-alone binning for ljmd_cell_assign (original order) ------------------------------------
__global__ void cell_assign_kernel(const float2* __restrict__ R, int N, int nrows, int nbx, float inv_hy,
                                   float inv_wx, int* __restrict__ cell_id, int* __restrict__ cell_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float2 r = R[i];
    const int c = strip_coord(r.y, inv_hy, nrows) * nbx + strip_coord(r.x, inv_wx, nbx);
    if (cell_id) cell_id[i] = c;
    if (cell_count) atomicAdd(&cell_count[c], 1);
}

}  // namespace

// ----------------------------------------------------------------------------------------------------
struct Cells {
    int G = 0, nrows = 0, nbx = 0, ncells = 0, Nalloc = 0;
    float inv_hy = 0, inv_wx = 0, hy = 0, wx = 0, rlist = 0;
    float *X[2] = {nullptr, nullptr}, *Y[2] = {nullptr, nullptr};
    float2 *V[2] = {nullptr, nullptr}, *Rb = nullptr, *Fs = nullptr;
    int *orig[2] = {nullptr, nullptr};
    int *key = nullptr, *rank = nullptr, *tmp = nullptr, *cell_count = nullptr, *cell_start = nullptr,
        *row_tot = nullptr, *sched = nullptr;
    int2* unit_tab = nullptr;
    UnitPlan* plans = nullptr;
    uint2* pranges = nullptr;
    int maxunits = 0;
    size_t smem = 0;
    float *pe_part = nullptr, *ke_part = nullptr;
    int* state = nullptr;
    unsigned* bar = nullptr;
    long long* prof = nullptr;
};

int cells_create(ljmd_handle* h) {
    Cells* cl = new Cells();
    h->cells = cl;
    const long long N = h->p.N;
    if (N > (1ll << 30)) { set_error("N too large for 32-bit particle indices"); return LJMD_E_INVALID; }
    const float rc = h->p.rc, skin = h->p.skin;
    cl->rlist = rc + skin;
    // rows at least rc + skin high, bins at least (rc + skin) / K wide, each with a few ulp(box) of
    // margin for the fp32 binning (one multiply + truncation)
    const double ulp = (double)(nextafterf(h->p.box, 2.0f * h->p.box) - h->p.box);
    cl->nrows = (int)floor((double)h->p.box / ((double)cl->rlist + 8.0 * ulp));
    cl->nbx   = (int)floor((double)h->p.box / ((double)cl->rlist / CL_K + 4.0 * ulp));
    if (cl->nrows < 3 || cl->nbx < 2 * CL_K + 1) {
        set_error("box %.3f is smaller than 3 rows of height rc+skin=%.3f: use the all-pairs path",
                  h->p.box, cl->rlist);
        return LJMD_E_INVALID;
    }
    if ((long long)cl->nrows * cl->nbx > (1ll << 30)) { set_error("cell index too large"); return LJMD_E_INVALID; }
    cl->ncells = cl->nrows * cl->nbx;
    cl->hy = h->p.box / (float)cl->nrows;
    cl->wx = h->p.box / (float)cl->nbx;
    cl->inv_hy = (float)cl->nrows / h->p.box;
    cl->inv_wx = (float)cl->nbx / h->p.box;
    cl->Nalloc = (int)(((N + 3 * (long long)cl->nrows + 63) / 64) * 64 + 64);

    cl->maxunits = (int)(N / CL_UNIT + cl->nrows + 2);
    cl->smem = sizeof(Stage) * CL_NST * CL_WARPS;
    LJ_CUDA(cudaFuncSetAttribute(cells_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cl->smem));
    int per_sm = 0;
    LJ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cells_persistent_kernel, CL_THREADS, cl->smem));
    if (per_sm < 1) { set_error("cell-list kernel does not fit on an SM"); return LJMD_E_STATE; }
    if (const char* e = getenv("LJMD_CELLS_CTAS_PER_SM")) per_sm = std::min(per_sm, std::max(1, atoi(e)));
    long long g = (long long)per_sm * h->num_sms;
    g = std::min<long long>(g, std::max<long long>(1, (N + CL_UNIT * CL_WARPS - 1) / (CL_UNIT * CL_WARPS)));
    cl->G = (int)g;

    const size_t na = (size_t)cl->Nalloc;
    for (int k = 0; k < 2; ++k) {
        LJ_CUDA(cudaMalloc(&cl->X[k], sizeof(float) * na));
        LJ_CUDA(cudaMalloc(&cl->Y[k], sizeof(float) * na));
        LJ_CUDA(cudaMalloc(&cl->V[k], sizeof(float2) * na));
        LJ_CUDA(cudaMalloc(&cl->orig[k], sizeof(int) * na));
    }
    LJ_CUDA(cudaMalloc(&cl->Rb, sizeof(float2) * na));
    LJ_CUDA(cudaMalloc(&cl->Fs, sizeof(float2) * na));
    LJ_CUDA(cudaMalloc(&cl->key, sizeof(int) * na));
    LJ_CUDA(cudaMalloc(&cl->rank, sizeof(int) * na));
    LJ_CUDA(cudaMalloc(&cl->tmp, sizeof(int) * na));
    LJ_CUDA(cudaMalloc(&cl->cell_count, sizeof(int) * (size_t)cl->ncells));
    LJ_CUDA(cudaMemset(cl->cell_count, 0, sizeof(int) * (size_t)cl->ncells));
    LJ_CUDA(cudaMalloc(&cl->cell_start, sizeof(int) * ((size_t)cl->ncells + 1 + 8)));   // + bulk-copy round-up
    LJ_CUDA(cudaMemset(cl->cell_start, 0, sizeof(int) * ((size_t)cl->ncells + 1 + 8)));
    LJ_CUDA(cudaMalloc(&cl->row_tot, sizeof(int) * (size_t)cl->nrows));
    LJ_CUDA(cudaMalloc(&cl->unit_tab, sizeof(int2) * (size_t)cl->maxunits));
    LJ_CUDA(cudaMalloc(&cl->plans, sizeof(UnitPlan) * (size_t)cl->maxunits));
    LJ_CUDA(cudaMalloc(&cl->sched, sizeof(int) * 2));
    LJ_CUDA(cudaMalloc(&cl->pranges, sizeof(uint2) * (na / 2 + 64)));
    LJ_CUDA(cudaMemset(cl->pranges, 0, sizeof(uint2) * (na / 2 + 64)));
    LJ_CUDA(cudaMalloc(&cl->pe_part, sizeof(float) * 2 * cl->maxunits));
    LJ_CUDA(cudaMalloc(&cl->ke_part, sizeof(float) * 2 * cl->maxunits));
    LJ_CUDA(cudaMalloc(&cl->state, sizeof(int) * ST_WORDS));
    LJ_CUDA(cudaMemset(cl->state, 0, sizeof(int) * ST_WORDS));
    LJ_CUDA(cudaMalloc(&cl->bar, sizeof(unsigned)));
    if (getenv("LJMD_CELLS_PROF")) LJ_CUDA(cudaMalloc(&cl->prof, sizeof(long long) * 12 * cl->G));
    return 0;
}

void cells_destroy(ljmd_handle* h) {
    Cells* cl = h->cells;
    if (!cl) return;
    for (int k = 0; k < 2; ++k) { cudaFree(cl->X[k]); cudaFree(cl->Y[k]); cudaFree(cl->V[k]); cudaFree(cl->orig[k]); }
    cudaFree(cl->Rb); cudaFree(cl->Fs); cudaFree(cl->key); cudaFree(cl->rank); cudaFree(cl->tmp);
    cudaFree(cl->cell_count); cudaFree(cl->cell_start); cudaFree(cl->row_tot);
    cudaFree(cl->unit_tab); cudaFree(cl->plans); cudaFree(cl->pranges); cudaFree(cl->sched);
    cudaFree(cl->pe_part); cudaFree(cl->ke_part);
    cudaFree(cl->state); cudaFree(cl->bar); cudaFree(cl->prof);
    delete cl;
    h->cells = nullptr;
}

static void fill_args(ljmd_handle* h, CellsArgs& a) {
    Cells* cl = h->cells;
    a.pc = h->pc;
    a.N = (int)h->p.N; a.Nalloc = cl->Nalloc; a.G = cl->G;
    a.nrows = cl->nrows; a.nbx = cl->nbx; a.ncells = cl->ncells;
    a.inv_hy = cl->inv_hy; a.inv_wx = cl->inv_wx;
    a.half_skin2 = (0.5f * h->p.skin) * (0.5f * h->p.skin);
    a.dt = h->p.dt;
    for (int k = 0; k < 2; ++k) { a.X[k] = cl->X[k]; a.Y[k] = cl->Y[k]; a.V[k] = cl->V[k]; a.orig[k] = cl->orig[k]; }
    a.Rb = cl->Rb; a.Fs = cl->Fs;
    a.key = cl->key; a.rank = cl->rank; a.tmp = cl->tmp;
    a.cell_count = cl->cell_count; a.cell_start = cl->cell_start; a.row_tot = cl->row_tot;
    a.sched = cl->sched;
    a.unit_tab = cl->unit_tab; a.plans = cl->plans; a.pranges = cl->pranges; a.maxunits = cl->maxunits;
    a.pe_part = cl->pe_part; a.ke_part = cl->ke_part;
    a.state = cl->state; a.bar = cl->bar; a.prof = cl->prof;
}

static int launch(ljmd_handle* h, CellsArgs& a) {
    Cells* cl = h->cells;
    LJ_CUDA(cudaMemsetAsync(cl->bar, 0, sizeof(unsigned), h->stream));
    LJ_CUDA(cudaMemsetAsync(cl->sched, 0, sizeof(int) * 2, h->stream));
    void* args[] = {(void*)&a};
    LJ_CUDA(cudaLaunchCooperativeKernel((void*)cells_persistent_kernel, dim3(cl->G), dim3(CL_THREADS),
                                        args, cl->smem, h->stream));
    h->launches++;
    return 0;
}

int cells_run(ljmd_handle* h, const float2* R_in, const float2* V_in, float2* R_out, float2* V_out,
              float2* F_out, float* pe_out, const RunCtl& rc) {
    Cells* cl = h->cells;
    const long long N = h->p.N;
    cudaStream_t st = h->stream;
    if (rc.nsteps > 0 && rc.traj && rc.S > 0)
        LJ_CUDA(cudaMemsetAsync(rc.traj, 0, sizeof(float2) * N * rc.S, st));      // MD:89
    // fresh call: parities 0, rebuild counter 0, flag 0 (the error word is sticky)
    LJ_CUDA(cudaMemsetAsync(cl->state, 0, sizeof(int) * 3, st));
    LJ_CUDA(cudaMemsetAsync(cl->state + ST_FLAG, 0, sizeof(int) * 3, st));
    CellsArgs a{};
    fill_args(h, a);
    a.R_in = R_in; a.V_in = V_in;
    a.rc = rc;
    a.R_out = R_out; a.V_out = V_out; a.F_out = F_out; a.pe_out = pe_out;
    a.mode = 0;
    // bound one launch to roughly half a second (conservative 2e10 particle-steps/s)
    long long chunk = std::max<long long>(1, (long long)(1.0e10 / (double)N));
    if (const char* e = getenv("LJMD_CELLS_CHUNK")) chunk = std::max(1, atoi(e));
    if (h->timed) LJ_CUDA(cudaEventRecord(h->ev0, st));
    long long s = -1;
    while (s < rc.nsteps) {
        const long long e = std::min(rc.nsteps, s + chunk);
        a.s_begin = s; a.s_end = std::max<long long>(e, 0);
        int r = launch(h, a);
        if (r) return r;
        s = a.s_end;
        if (rc.nsteps == 0) break;
    }
    if (h->timed) LJ_CUDA(cudaEventRecord(h->ev1, st));
    if (cl->prof) {   // debug: mean clocks per phase of the LAST launch, summed over its steps
        LJ_CUDA(cudaStreamSynchronize(st));
        std::vector<long long> pv(12 * cl->G);
        LJ_CUDA(cudaMemcpy(pv.data(), cl->prof, sizeof(long long) * pv.size(), cudaMemcpyDeviceToHost));
        const char* nm[8] = {"force+integrate", "step barrier", "B1 bin+hist", "B2 row totals",
                             "B3 scan", "B4 scatter", "B5 order+gather", "B6 copy plans"};
        for (int k = 0; k < 8; ++k) {
            double mean = 0, mx = 0;
            for (int c = 0; c < cl->G; ++c) { mean += pv[c * 12 + k]; mx = std::max<double>(mx, (double)pv[c * 12 + k]); }
            fprintf(stderr, "[ljmd cells prof] %-16s mean %12.0f  max %12.0f clocks (launch total)\n", nm[k], mean / cl->G, mx);
        }
    }
    return 0;
}

int cells_geometry(ljmd_handle* h, int* nrows, int* nbx, int* kbins, float* inv_hy, float* inv_wx) {
    Cells* cl = h->cells;
    if (nrows) *nrows = cl->nrows;
    if (nbx) *nbx = cl->nbx;
    if (kbins) *kbins = CL_K;
    if (inv_hy) *inv_hy = cl->inv_hy;
    if (inv_wx) *inv_wx = cl->inv_wx;
    return 0;
}

int cells_assign(ljmd_handle* h, const float2* R, int* cell_id, int* cell_count) {
    Cells* cl = h->cells;
    const int N = (int)h->p.N;
    if (cell_count) LJ_CUDA(cudaMemsetAsync(cell_count, 0, sizeof(int) * (size_t)cl->ncells, h->stream));
    cell_assign_kernel<<<(N + 255) / 256, 256, 0, h->stream>>>(R, N, cl->nrows, cl->nbx, cl->inv_hy,
                                                               cl->inv_wx, cell_id, cell_count);
    LJ_CUDA(cudaGetLastError());
    h->launches++;
    return 0;
}

int cells_neighbor_count(ljmd_handle* h, const float2* R, float radius, int* nbr_count) {
    Cells* cl = h->cells;
    if (!(radius > 0.0f) || radius > cl->rlist) {
        set_error("neighbor_count radius %.4f must be in (0, rc + skin = %.4f]", radius, cl->rlist);
        return LJMD_E_INVALID;
    }
    LJ_CUDA(cudaMemsetAsync(cl->state, 0, sizeof(int) * 3, h->stream));
    LJ_CUDA(cudaMemsetAsync(cl->state + ST_FLAG, 0, sizeof(int) * 3, h->stream));
    CellsArgs a{};
    fill_args(h, a);
    a.R_in = R; a.V_in = nullptr;
    a.mode = 1;
    a.count_r2 = radius * radius;
    a.count_out = nbr_count;
    a.s_begin = -1; a.s_end = 0;
    return launch(h, a);
}

long long cells_last_rebuilds(ljmd_handle* h) {
    Cells* cl = h->cells;
    int v = 0;
    cudaStreamSynchronize(h->stream);
    cudaMemcpy(&v, cl->state + ST_REBUILDS, sizeof(int), cudaMemcpyDeviceToHost);
    return v;
}

int cells_check_error(ljmd_handle* h) {
    Cells* cl = h->cells;
    if (!cl) return 0;
    int e = 0;
    LJ_CUDA(cudaMemcpy(&e, cl->state + ST_ERR, sizeof(int), cudaMemcpyDeviceToHost));
    if (e) { set_error("cell-list persistent kernel: grid barrier timed out (flag %d)", e); return LJMD_E_STATE; }
    return 0;
}

}  // namespace ljmd
