// cells.cu — sorted cell-list / compressed Verlet-list path for large N (BASELINE configs 4-5).
//
// Not in the reference (it only has the dense O(N^2) form, MD:51): the pair arithmetic is the
// same as the all-pairs path (subtract, exact min-image, unfused r2, r2 < rc^2), so on identical
// fp32 inputs it evaluates exactly the pair set of the all-pairs+cutoff oracle.
//
// Data layout in HBM (all in CELL-SORTED order, rebuilt only when the skin is exhausted):
//   Rs[2]    float2  positions, ping-pong by step (the read buffer is never written in a step)
//   Vh[2]    float2  velocities (buffers swap at a rebuild)
//   orig[2]  int32   original particle index of each sorted slot
//   Rbuild   float2  positions at the last rebuild (max-displacement test against skin/2)
//   meta     uint32  neighbour count | edge flag << 8 (particle of a boundary cell)
//   nbr2     uint32  compressed Verlet list, 2 neighbours per word, ELL layout
//                    nbr2[(s / 2) * Npad + i]; each neighbour is the int16 index distance j - i
//                    in sorted order (folded modulo N for the periodic top/bottom rows): sorted
//                    row-major cell order keeps every neighbour within +-(one cell row + 64)
//                    slots, so 2 bytes replace a 4-byte index and decode is one add
//   cell_start int32 prefix-sum cell index over ncell^2 cells (cells in row-major order)
//
// One PERSISTENT cooperative kernel runs a whole ljmd_run(): per step ONE pass over the state
// (force from the list + velocity-Verlet + energies + displacement test: each particle's state
// is read once and written once) and one grid barrier; the rebuild (single-digit radix =
// counting sort of cell ids, prefix sum, deterministic in-cell ordering, gather, list build)
// runs inside the same kernel under a grid-uniform condition: no host round trip (MD:82,103).
#include "ljmd_device.cuh"

#include <algorithm>
#include <cstdio>

namespace ljmd {

namespace {

constexpr int CL_THREADS  = 512;
constexpr int CL_MAXN     = 64;    // list capacity per particle; 19.7 neighbours expected at rho 0.8
constexpr int CL_BATCH    = 4;     // list words (= 8 neighbours) decoded and gathered together
constexpr int CL_PREFETCH = 16;    // list words requested up-front per particle (32 neighbours)
constexpr int CL_CELLCAP  = 16;    // in-register ordering of a cell's members (slow path beyond)

enum { ST_PR = 0, ST_PV = 1, ST_REBUILDS = 2, ST_ERR = 3, ST_FLAG = 4, ST_WORDS = 8 };
enum { CERR_BARRIER = 1, CERR_LIST_OVERFLOW = 2 };

struct CellsArgs {
    PairConsts pc;
    int   N, Npad, G, ncell, ncells;
    float inv_cell, cell_size, rlist2, half_skin2, dt;
    float2* Rs[2];
    float2* Vh[2];
    int*    orig[2];
    float2* Rbuild;
    int *key, *slot_src, *cell_count, *cell_start, *fill, *chunk_tot;
    unsigned* meta;
    unsigned* nbr2;
    float  *pe_part, *ke_part;      // [2*G]
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

__device__ __forceinline__ int cell_coord(float x, float inv_cell, int ncell) {
    // one fp32 multiply then truncation (x >= 0); the clamp handles x == box (MD:72 closed range)
    int c = (int)(x * inv_cell);
    return max(0, min(c, ncell - 1));
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
    int pr, pv;
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

// The candidates of stencil row yy for a particle in cell column cx are the members of cells
// cx-1, cx, cx+1 of that row: contiguous in sorted order, except that at the two edge columns
// the wrapped cell sits at the other end of the row.  The window is therefore expressed in a
// VIRTUAL index space (the row extended by one periodic copy on each side); to_real() folds a
// virtual index back into the row.  Interior columns never need the fold.
struct RowWin { int base, len, rstart, rlen; };

__device__ __forceinline__ RowWin row_window(const int* __restrict__ cs, int nc, int cx, int yy) {
    const int row0 = yy * nc;
    RowWin w;
    w.rstart = cs[row0];
    w.rlen = cs[row0 + nc] - w.rstart;
    int lo, hi;
    if (cx == 0)           { lo = cs[row0 + nc - 1] - w.rlen; hi = cs[row0 + 2]; }
    else if (cx == nc - 1) { lo = cs[row0 + nc - 2];          hi = cs[row0 + 1] + w.rlen; }
    else                   { lo = cs[row0 + cx - 1];          hi = cs[row0 + cx + 2]; }
    w.base = lo;
    w.len = hi - lo;
    return w;
}
__device__ __forceinline__ int to_real(int jv, int rstart, int rlen) {
    jv += (jv < rstart) ? rlen : 0;
    jv -= (jv >= rstart + rlen) ? rlen : 0;
    return jv;
}
__device__ __forceinline__ int wrap_row(int y, int nc) {
    y += (y < 0) ? nc : 0;
    y -= (y >= nc) ? nc : 0;
    return y;
}

// ---- rebuild: bin -> counting sort (radix 2^k single digit) -> prefix sum -> gather -> list ----
__device__ void cells_rebuild(const CellsArgs& a, Ctx& ctx, int* sscan, float rl2) {
    const int tid = threadIdx.x, gtid = blockIdx.x * CL_THREADS + tid, gsz = a.G * CL_THREADS;
    const float2* Rcur = a.Rs[ctx.pr];
    CL_PROF(0);
    // R1: clear per-cell counters
    for (int c = gtid; c < a.ncells; c += gsz) { a.cell_count[c] = 0; a.fill[c] = 0; }
    CL_BARRIER();
    CL_PROF(2);
    // R2: cell id of every particle (current order) + histogram
    for (int k = gtid; k < a.N; k += gsz) {
        const float2 r = Rcur[k];
        const int c = cell_coord(r.y, a.inv_cell, a.ncell) * a.ncell + cell_coord(r.x, a.inv_cell, a.ncell);
        a.key[k] = c;
        atomicAdd(&a.cell_count[c], 1);
    }
    CL_BARRIER();
    CL_PROF(3);
    // R3: prefix sum over cells.  pass 1: per-CTA chunk totals
    const int chunk = (a.ncells + a.G - 1) / a.G;
    const int c0 = min(a.ncells, blockIdx.x * chunk), c1 = min(a.ncells, c0 + chunk);
    {
        int s = 0;
        for (int c = c0 + tid; c < c1; c += CL_THREADS) s += a.cell_count[c];
        int tot;
        (void)block_exscan(s, sscan, &tot);
        if (tid == 0) a.chunk_tot[blockIdx.x] = tot;
    }
    CL_BARRIER();
    CL_PROF(4);
    //     pass 2: chunk offset + in-chunk exclusive scan -> cell_start
    {
        int s = 0;
        for (int c = tid; c < blockIdx.x; c += CL_THREADS) s += a.chunk_tot[c];
        int carry;
        (void)block_exscan(s, sscan, &carry);
        for (int cb = c0; cb < c1; cb += CL_THREADS) {
            const int c = cb + tid;
            const int v = (c < c1) ? a.cell_count[c] : 0;
            int tot;
            const int ex = block_exscan(v, sscan, &tot);
            if (c < c1) a.cell_start[c] = carry + ex;
            carry += tot;
        }
        if (blockIdx.x == a.G - 1 && tid == 0) a.cell_start[a.ncells] = a.N;
    }
    CL_BARRIER();
    CL_PROF(5);
    // R5: scatter source indices into their cell's slot range (arbitrary order inside a cell)
    for (int k = gtid; k < a.N; k += gsz) {
        const int c = a.key[k];
        const int slot = a.cell_start[c] + atomicAdd(&a.fill[c], 1);
        a.slot_src[slot] = k;
    }
    CL_BARRIER();
    CL_PROF(6);
    // R6a: order every cell's members by ORIGINAL particle index => the sorted order (cell, orig)
    //      is a pure function of the positions: bit-reproducible summation order downstream.
    //      Rank by counting in registers (no data-dependent control flow) for cells up to
    //      CL_CELLCAP members; insertion sort for the rare denser cell.
    const int* orig_old = a.orig[ctx.pv];
    for (int c = gtid; c < a.ncells; c += gsz) {
        const int b = a.cell_start[c], e = a.cell_start[c + 1], n = e - b;
        if (n <= 1) continue;
        if (n <= CL_CELLCAP) {
            int k[CL_CELLCAP], o[CL_CELLCAP];
#pragma unroll
            for (int p = 0; p < CL_CELLCAP; ++p) {
                k[p] = (p < n) ? a.slot_src[b + p] : 0;
                o[p] = (p < n) ? orig_old[k[p]] : 0x7fffffff;
            }
#pragma unroll
            for (int p = 0; p < CL_CELLCAP; ++p) {
                int rank = 0;
#pragma unroll
                for (int q = 0; q < CL_CELLCAP; ++q) rank += (o[q] < o[p]);
                if (p < n) a.slot_src[b + rank] = k[p];
            }
        } else {
            for (int p = b + 1; p < e; ++p) {
                const int kp = a.slot_src[p], op = orig_old[kp];
                int q = p - 1;
                while (q >= b) {
                    const int kq = a.slot_src[q];
                    if (orig_old[kq] <= op) break;
                    a.slot_src[q + 1] = kq;
                    --q;
                }
                a.slot_src[q + 1] = kp;
            }
        }
    }
    CL_BARRIER();
    CL_PROF(7);
    // R6b: gather the state into the new order (coalesced writes)
    {
        float2* Rn = a.Rs[ctx.pr ^ 1];
        const float2* Vo = a.Vh[ctx.pv];
        float2* Vn = a.Vh[ctx.pv ^ 1];
        int* on = a.orig[ctx.pv ^ 1];
        for (int i = gtid; i < a.N; i += gsz) {
            const int k = a.slot_src[i];
            const float2 r = Rcur[k];
            Rn[i] = r;
            a.Rbuild[i] = r;
            Vn[i] = Vo[k];
            on[i] = orig_old[k];
            a.meta[i] = (unsigned)a.key[k];                 // cell id, replaced by count|edge in R7
        }
    }
    ctx.pr ^= 1;
    ctx.pv ^= 1;
    CL_BARRIER();
    CL_PROF(8);
    // R7: compressed Verlet list from the 3x3 stencil: r2 < rl2, j != i, with r2 the same unfused
    //     fp32 expression as the oracle (bit-exact neighbour counts).  Warps whose particles all
    //     sit in interior cells skip the min-image and the window fold (both are identities there).
    {
        const float2* __restrict__ R = a.Rs[ctx.pr];
        const int* __restrict__ cs = a.cell_start;
        const PairConsts pc = a.pc;
        const int nc = a.ncell, N = a.N, halfN = a.N >> 1;
        for (int base = blockIdx.x * CL_THREADS; base < a.Npad; base += gsz) {
            const int  i = base + tid;
            const bool live = i < N;
            float2 ri = make_float2(0.0f, 0.0f);
            int cx = 1, cy = 1;
            if (live) {
                ri = R[i];
                const int c = (int)a.meta[i];
                cy = c / nc; cx = c - cy * nc;
            }
            const bool edge = (cx == 0) | (cx == nc - 1) | (cy == 0) | (cy == nc - 1);
            const bool wedge = __any_sync(0xffffffffu, edge);
            int n = 0;
            unsigned word = 0;
            // append neighbour j (hit) to the packed list
            auto push = [&](int dj) {
                if (a.mode == 0 && n < CL_MAXN) {
                    word |= ((unsigned)dj & 0xffffu) << (16 * (n & 1));
                    if (n & 1) { a.nbr2[(size_t)(n >> 1) * a.Npad + i] = word; word = 0; }
                }
                ++n;
            };
            if (live) {
#pragma unroll 1
                for (int r = 0; r < 3 && wedge; ++r) {
                    const int yy = wrap_row(cy + r - 1, nc);
                    {
                        const RowWin w = row_window(cs, nc, cx, yy);
#pragma unroll 1
                        for (int k0 = 0; k0 < w.len; k0 += 4) {
                            int jj[4]; float2 rj[4];
#pragma unroll
                            for (int t = 0; t < 4; ++t) {          // 4 independent loads in flight
                                jj[t] = to_real(w.base + min(k0 + t, w.len - 1), w.rstart, w.rlen);
                                rj[t] = R[jj[t]];
                            }
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const float dx = min_image(__fsub_rn(ri.x, rj[t].x), pc.box, pc.timg);
                                const float dy = min_image(__fsub_rn(ri.y, rj[t].y), pc.box, pc.timg);
                                const float r2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                                if (k0 + t < w.len && jj[t] != i && r2 < rl2) {
                                    int dj = jj[t] - i;                    // fold modulo N into +-N/2
                                    dj += (dj < -halfN) ? N : 0;
                                    dj -= (dj > halfN) ? N : 0;
                                    push(dj);
                                }
                            }
                        }
                    }
                }
                if (!wedge) {
                    // interior: the three row windows are plain contiguous ranges; walk them as one
                    // flattened candidate sequence with 8 position loads in flight
                    int jlo[3], len[3];
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const int row0 = (cy + r - 1) * nc + cx;
                        jlo[r] = cs[row0 - 1];
                        len[r] = cs[row0 + 2] - jlo[r];
                    }
                    const int e0 = len[0], e1 = len[0] + len[1], tot = e1 + len[2];
                    const int d1 = jlo[1] - e0, d2 = jlo[2] - e1;
#pragma unroll 1
                    for (int cb = 0; cb < tot; cb += 8) {
                        int jj[8]; float2 rj[8];
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            const int c = min(cb + t, tot - 1);
                            jj[t] = c + ((c < e0) ? jlo[0] : ((c < e1) ? d1 : d2));
                            rj[t] = R[jj[t]];
                        }
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            const float dx = __fsub_rn(ri.x, rj[t].x), dy = __fsub_rn(ri.y, rj[t].y);
                            const float r2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                            if (cb + t < tot && jj[t] != i && r2 < rl2) push(jj[t] - i);
                        }
                    }
                }
            }
            if (a.mode == 0) {
                // the force pass walks every lane to the warp's longest list, in batches of
                // CL_BATCH words: pad the shorter lists with dj = 0 (the particle itself, which
                // the keep predicate masks) so that decode needs no bounds handling.
                int nmax = min(n, CL_MAXN);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
                if (live) {
                    const int nw_pad = min(CL_MAXN / 2, (((nmax + 1) >> 1) + CL_BATCH - 1) / CL_BATCH * CL_BATCH);
                    int wdone = min(n, CL_MAXN) >> 1;
                    if ((n & 1) && n < CL_MAXN) { a.nbr2[(size_t)wdone * a.Npad + i] = word; ++wdone; }
                    for (int u = wdone; u < nw_pad; ++u) a.nbr2[(size_t)u * a.Npad + i] = 0u;
                    if (n > CL_MAXN) atomicOr(a.state + ST_ERR, CERR_LIST_OVERFLOW);
                    a.meta[i] = (unsigned)min(n, CL_MAXN) | (edge ? 0x100u : 0u);
                }
            } else if (live) {
                a.key[i] = n;                   // count mode: full count, nothing stored
            }
        }
    }
    if (blockIdx.x == 0 && tid == 0) a.state[ST_REBUILDS] += 1;
    CL_BARRIER();
    CL_PROF(9);
}

// ---- per-step pass: list force for one particle ----------------------------------------------------
// EDGE: the warp contains a particle of a boundary cell => min-image and the modulo-N index fold
// are applied; interior warps skip both (|d| << box/2 for every listed neighbour).
// The particle's list words are fetched 8 at a time (16 neighbours in flight) before any gather.
template <bool PE, bool EDGE>
__device__ __forceinline__ void pair_list(const PairConsts& pc, float2 ri, float2 rj, bool keep,
                                          float& fx, float& fy, float& pe) {
    float dx = __fsub_rn(ri.x, rj.x), dy = __fsub_rn(ri.y, rj.y);
    if (EDGE) { dx = min_image(dx, pc.box, pc.timg); dy = min_image(dy, pc.box, pc.timg); }
    const float r2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    float ir2 = rcp_approx(r2);
    ir2 = (keep & (r2 < pc.rc2)) ? ir2 : 0.0f;
    const float ir6 = ir2 * ir2 * ir2;
    const float f = fmaf(ir6, pc.c12, -pc.c6) * (ir6 * ir2);
    fx = fmaf(f, dx, fx);
    fy = fmaf(f, dy, fy);
    if (PE) pe = fmaf(ir6, fmaf(ir6, pc.d12, -pc.d6), pe);
}

template <bool PE, bool EDGE>
__device__ __forceinline__ void list_batch(const PairConsts& pc, const float2* __restrict__ R, int i,
                                           int N, float2 ri, int n, int s0, const unsigned (&w)[CL_BATCH],
                                           float& fx, float& fy, float& pe) {
    int    j[2 * CL_BATCH];
    float2 rj[2 * CL_BATCH];
#pragma unroll
    for (int u = 0; u < CL_BATCH; ++u) {
        j[2 * u]     = i + (int)(short)(w[u] & 0xffffu);
        j[2 * u + 1] = i + ((int)w[u] >> 16);
    }
#pragma unroll
    for (int t = 0; t < 2 * CL_BATCH; ++t) {
        if (EDGE) { j[t] += (j[t] < 0) ? N : 0; j[t] -= (j[t] >= N) ? N : 0; }
        rj[t] = R[j[t]];                            // all gathers of the batch in flight together
    }
#pragma unroll
    for (int t = 0; t < 2 * CL_BATCH; ++t)
        pair_list<PE, EDGE>(pc, ri, rj[t], (s0 + t) < n, fx, fy, pe);
}

// All list words of the particle (up to CL_PREFETCH) are requested in ONE burst before any of
// them is decoded: the pass is latency-bound otherwise (one HBM round trip per batch).
template <bool PE, bool EDGE>
__device__ __forceinline__ void list_force(const CellsArgs& a, const float2* __restrict__ R, int i,
                                           float2 ri, int n, int nmax, float& fx, float& fy,
                                           float& pe) {
    const PairConsts pc = a.pc;
    const unsigned* __restrict__ np = a.nbr2 + i;
    const size_t stride = (size_t)a.Npad;
    const int nw = (nmax + 1) >> 1;                 // warp-uniform; lists are zero-padded to a batch
    const int N = a.N;
    unsigned w[CL_PREFETCH];
#pragma unroll
    for (int u = 0; u < CL_PREFETCH; ++u) w[u] = (u < nw) ? np[(size_t)u * stride] : 0u;
#pragma unroll
    for (int u0 = 0; u0 < CL_PREFETCH; u0 += CL_BATCH) {
        if (u0 < nw) {                              // warp-uniform
            unsigned wb[CL_BATCH];
#pragma unroll
            for (int u = 0; u < CL_BATCH; ++u) wb[u] = w[u0 + u];
            list_batch<PE, EDGE>(pc, R, i, N, ri, n, 2 * u0, wb, fx, fy, pe);
        }
    }
    for (int u0 = CL_PREFETCH; u0 < nw; u0 += CL_BATCH) {     // rare: more than 2*CL_PREFETCH neighbours
        unsigned wb[CL_BATCH];
#pragma unroll
        for (int u = 0; u < CL_BATCH; ++u) wb[u] = np[(size_t)(u0 + u) * stride];
        list_batch<PE, EDGE>(pc, R, i, N, ri, n, 2 * u0, wb, fx, fy, pe);
    }
}

__global__ void __launch_bounds__(CL_THREADS, 2)
cells_persistent_kernel(const CellsArgs a) {
    __shared__ int   sscan[CL_THREADS / 32 + 1];
    __shared__ float sred[CL_THREADS / 32];
    __shared__ float s_lambda;
    __shared__ long long s_pt[12];
    const int tid = threadIdx.x, gtid = blockIdx.x * CL_THREADS + tid, gsz = a.G * CL_THREADS;
    const RunCtl rc = a.rc;
    Ctx ctx;
    ctx.epoch = 0;
    ctx.pt = a.prof ? s_pt : nullptr;
    if (ctx.pt && tid == 0) { for (int k = 0; k < 11; ++k) s_pt[k] = 0; s_pt[11] = clock64(); }
    ctx.pr = a.state[ST_PR];
    ctx.pv = a.state[ST_PV];

    if (a.s_begin < 0) {
        // load the caller's state (original order) and build the first list
        for (int i = gtid; i < a.N; i += gsz) {
            a.Rs[ctx.pr][i] = a.R_in[i];
            a.Vh[ctx.pv][i] = a.V_in ? a.V_in[i] : make_float2(0.0f, 0.0f);
            a.orig[ctx.pv][i] = i;
        }
        CL_BARRIER();
        cells_rebuild(a, ctx, sscan, (a.mode == 1) ? a.count_r2 : a.rlist2);
        if (a.mode == 1) {
            const int* og = a.orig[ctx.pv];
            for (int i = gtid; i < a.N; i += gsz) a.count_out[og[i]] = a.key[i];
            if (gtid == 0) { a.state[ST_PR] = ctx.pr; a.state[ST_PV] = ctx.pv; }
            return;
        }
    }

    for (long long s = a.s_begin; s < a.s_end; ++s) {
        const int  par     = (int)((s + 1) & 1);
        const bool kick1   = (s >= 0);
        const bool final   = (s == rc.nsteps - 1);
        const bool want_e  = kick1 && rc.energy_every > 0 && (s % rc.energy_every == 0);
        const bool want_pe = want_e || (rc.nsteps == 0 && a.pe_out != nullptr);
        const bool thermo  = kick1 && rc.thermo_every > 0 && rc.thermo_kT > 0.0f &&
                             ((s + 1) % rc.thermo_every == 0);
        const bool sample  = kick1 && rc.sample_every > 0 && (s % rc.sample_every == 0) &&
                             (s / rc.sample_every < rc.S);
        // rebuild requested by the previous step's displacement test?
        if (s > a.s_begin || a.s_begin >= 0) {
            if (__ldcg(a.state + ST_FLAG) == (int)(s + 1)) cells_rebuild(a, ctx, sscan, a.rlist2);
        }
        const float2* __restrict__ R = a.Rs[ctx.pr];
        float2*       Rnext = a.Rs[ctx.pr ^ 1];
        float2*       Vh    = a.Vh[ctx.pv];
        const int*    og    = a.orig[ctx.pv];

        float ke_thread = 0.0f, pe_thread = 0.0f;
        int   moved = 0;
        for (int base = blockIdx.x * CL_THREADS; base < a.Npad; base += gsz) {
            const int  i    = base + tid;
            const bool live = i < a.N;
            float2 ri = make_float2(0.0f, 0.0f);
            unsigned meta = 0u;                         // dead lanes: interior, no neighbours
            if (live) { ri = R[i]; meta = a.meta[i]; }
            const int n = meta & 0xff;
            int nmax = n;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
            const bool wedge = __any_sync(0xffffffffu, (meta & 0x100u) != 0u);
            float Fx = 0.0f, Fy = 0.0f, pe = 0.0f;
            const int ii = live ? i : 0;
            // epilogue operands requested now, consumed after the force loop
            float2 v = make_float2(0.0f, 0.0f), rb = make_float2(0.0f, 0.0f);
            if (live && rc.nsteps > 0) { v = Vh[i]; rb = a.Rbuild[i]; }
            if (want_pe) {
                if (wedge) list_force<true, true >(a, R, ii, ri, n, nmax, Fx, Fy, pe);
                else       list_force<true, false>(a, R, ii, ri, n, nmax, Fx, Fy, pe);
            } else {
                if (wedge) list_force<false, true >(a, R, ii, ri, n, nmax, Fx, Fy, pe);
                else       list_force<false, false>(a, R, ii, ri, n, nmax, Fx, Fy, pe);
            }
            if (!live) continue;
            pe_thread += pe;
            if (kick1) { v.x = kick(v.x, Fx, a.dt); v.y = kick(v.y, Fy, a.dt); }            // MD:74
            if (want_e || thermo) ke_thread += v.x * v.x + v.y * v.y;
            const int o = (sample || final) ? og[i] : 0;
            if (sample) rc.traj[(size_t)(s / rc.sample_every) * a.N + o] = ri;              // MD:93-100
            if (thermo) {
                Vh[i] = v;
                a.Rbuild[a.Npad + i] = make_float2(Fx, Fy);     // scratch half of Rbuild: F across barrier
                continue;
            }
            if (final) {
                if (a.R_out) a.R_out[o] = ri;
                if (a.V_out) a.V_out[o] = v;
                if (a.F_out) a.F_out[o] = make_float2(Fx, Fy);
            } else {
                v.x = kick(v.x, Fx, a.dt); v.y = kick(v.y, Fy, a.dt);                         // MD:70
                Vh[i] = v;
                const float2 rn = make_float2(drift(ri.x, v.x, a.dt, a.pc.box),               // MD:71-72
                                              drift(ri.y, v.y, a.dt, a.pc.box));
                Rnext[i] = rn;
                const float ddx = min_image(__fsub_rn(rn.x, rb.x), a.pc.box, a.pc.timg);
                const float ddy = min_image(__fsub_rn(rn.y, rb.y), a.pc.box, a.pc.timg);
                moved |= (ddx * ddx + ddy * ddy > a.half_skin2);
            }
        }
        if (want_pe) {
            float t = block_sum<CL_THREADS>(pe_thread, sred);
            if (tid == 0) __stcg(&a.pe_part[par * a.G + blockIdx.x], t);
        }
        if (want_e || thermo) {
            float t = block_sum<CL_THREADS>(ke_thread, sred);
            if (tid == 0) __stcg(&a.ke_part[par * a.G + blockIdx.x], t);
        }
        if (thermo) {
            CL_BARRIER();
            if (tid < 32) {
                double ke2 = 0.0;
                for (int k = tid; k < a.G; k += 32) ke2 += (double)__ldcg(&a.ke_part[par * a.G + k]);
                ke2 = warp_sum(ke2);
                if (tid == 0) s_lambda = sqrtf(rc.thermo_kT / ((float)(0.5 * ke2) / (float)a.N));
            }
            __syncthreads();
            const float lam = s_lambda;
            for (int i = gtid; i < a.N; i += gsz) {
                const float2 ri = R[i];
                const float2 F = a.Rbuild[a.Npad + i];
                float2 v = Vh[i];
                v.x *= lam; v.y *= lam;
                if (final) {
                    const int o = og[i];
                    if (a.R_out) a.R_out[o] = ri;
                    if (a.V_out) a.V_out[o] = v;
                    if (a.F_out) a.F_out[o] = F;
                } else {
                    v.x = kick(v.x, F.x, a.dt); v.y = kick(v.y, F.y, a.dt);
                    Vh[i] = v;
                    const float2 rn = make_float2(drift(ri.x, v.x, a.dt, a.pc.box),
                                                  drift(ri.y, v.y, a.dt, a.pc.box));
                    Rnext[i] = rn;
                    const float2 rb = a.Rbuild[i];
                    const float ddx = min_image(__fsub_rn(rn.x, rb.x), a.pc.box, a.pc.timg);
                    const float ddy = min_image(__fsub_rn(rn.y, rb.y), a.pc.box, a.pc.timg);
                    moved |= (ddx * ddx + ddy * ddy > a.half_skin2);
                }
            }
        }
        // a particle left the skin/2 ball: ask for a rebuild before the next force evaluation.
        // The flag carries the step stamp, so it never needs clearing (no reset race).
        if (__syncthreads_or(moved) && tid == 0) __stcg(a.state + ST_FLAG, (int)(s + 2));
        if (!final) ctx.pr ^= 1;
        CL_PROF(0);
        CL_BARRIER();
        CL_PROF(1);

        if (blockIdx.x == 0 && tid < 32 && want_pe) {
            double pe2 = 0.0, ke2 = 0.0;
            for (int k = tid; k < a.G; k += 32) {
                pe2 += (double)__ldcg(&a.pe_part[par * a.G + k]);
                if (want_e) ke2 += (double)__ldcg(&a.ke_part[par * a.G + k]);
            }
            pe2 = warp_sum(pe2);
            ke2 = warp_sum(ke2);
            if (tid == 0) {
                if (want_e) {
                    float* o = rc.ke_pe + 2 * (s / rc.energy_every);
                    o[0] = (float)(0.5 * ke2);
                    o[1] = (float)(0.5 * pe2);
                } else {
                    a.pe_out[0] = (float)(0.5 * pe2);
                }
            }
        }
    }
    if (gtid == 0) { a.state[ST_PR] = ctx.pr; a.state[ST_PV] = ctx.pv; }
    if (ctx.pt && tid == 0)
        for (int k = 0; k < 11; ++k) a.prof[blockIdx.x * 12 + k] = s_pt[k];
}

// ---- stand-alone binning for ljmd_cell_assign (original order) ------------------------------------
__global__ void cell_assign_kernel(const float2* __restrict__ R, int N, int ncell, float inv_cell,
                                   int* __restrict__ cell_id, int* __restrict__ cell_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float2 r = R[i];
    const int c = cell_coord(r.y, inv_cell, ncell) * ncell + cell_coord(r.x, inv_cell, ncell);
    if (cell_id) cell_id[i] = c;
    if (cell_count) atomicAdd(&cell_count[c], 1);
}

}  // namespace

// ----------------------------------------------------------------------------------------------------
struct Cells {
    int G = 0, ncell = 0, ncells = 0, Npad = 0;
    float inv_cell = 0, cell_size = 0, rlist = 0;
    float2 *Rs[2] = {nullptr, nullptr}, *Vh[2] = {nullptr, nullptr}, *Rbuild = nullptr;
    int *orig[2] = {nullptr, nullptr};
    int *key = nullptr, *slot_src = nullptr, *cell_count = nullptr, *cell_start = nullptr,
        *fill = nullptr, *chunk_tot = nullptr;
    unsigned *meta = nullptr, *nbr2 = nullptr;
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
    // cells must be at least rc + skin wide, with a few ulp(box) of margin for the fp32 binning
    const double margin = 8.0 * (double)(nextafterf(h->p.box, 2.0f * h->p.box) - h->p.box);
    cl->ncell = (int)floor((double)h->p.box / ((double)cl->rlist + margin));
    if (cl->ncell < 3) {
        set_error("box %.3f is smaller than 3 cells of width rc+skin=%.3f: use the all-pairs path",
                  h->p.box, cl->rlist);
        return LJMD_E_INVALID;
    }
    if (cl->ncell > 32768) cl->ncell = 32768;
    cl->ncells = cl->ncell * cl->ncell;
    cl->cell_size = h->p.box / (float)cl->ncell;
    cl->inv_cell = (float)cl->ncell / h->p.box;
    cl->Npad = (int)((N + 31) / 32 * 32);
    // int16 index distances: one cell row (+ a 3-cell window) must stay below 32767 slots
    if ((double)N / cl->ncell * 1.5 + 256.0 > 32767.0) {
        set_error("cell-list: N/ncell = %.0f particles per cell row exceeds the int16 list encoding", (double)N / cl->ncell);
        return LJMD_E_UNSUPPORTED;
    }

    int per_sm = 0;
    LJ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cells_persistent_kernel, CL_THREADS, 0));
    if (per_sm < 1) { set_error("cell-list kernel does not fit on an SM"); return LJMD_E_STATE; }
    int want = 2;
    if (const char* e = getenv("LJMD_CELLS_CTAS_PER_SM")) want = std::max(1, atoi(e));
    per_sm = std::min(per_sm, want);
    long long g = (long long)per_sm * h->num_sms;
    g = std::min<long long>(g, std::max<long long>(1, (N + CL_THREADS - 1) / CL_THREADS));
    cl->G = (int)g;

    const size_t nb2 = sizeof(float2) * (size_t)cl->Npad;
    for (int k = 0; k < 2; ++k) {
        LJ_CUDA(cudaMalloc(&cl->Rs[k], nb2));
        LJ_CUDA(cudaMalloc(&cl->Vh[k], nb2));
        LJ_CUDA(cudaMalloc(&cl->orig[k], sizeof(int) * (size_t)cl->Npad));
    }
    LJ_CUDA(cudaMalloc(&cl->Rbuild, 2 * nb2));      // second half: force scratch for the thermostat
    LJ_CUDA(cudaMalloc(&cl->key, sizeof(int) * (size_t)cl->Npad));
    LJ_CUDA(cudaMalloc(&cl->slot_src, sizeof(int) * (size_t)cl->Npad));
    LJ_CUDA(cudaMalloc(&cl->meta, sizeof(unsigned) * (size_t)cl->Npad));
    LJ_CUDA(cudaMalloc(&cl->nbr2, sizeof(unsigned) * (size_t)cl->Npad * (CL_MAXN / 2)));
    LJ_CUDA(cudaMalloc(&cl->cell_count, sizeof(int) * (size_t)cl->ncells));
    LJ_CUDA(cudaMalloc(&cl->fill, sizeof(int) * (size_t)cl->ncells));
    LJ_CUDA(cudaMalloc(&cl->cell_start, sizeof(int) * ((size_t)cl->ncells + 1)));
    LJ_CUDA(cudaMalloc(&cl->chunk_tot, sizeof(int) * cl->G));
    LJ_CUDA(cudaMalloc(&cl->pe_part, sizeof(float) * 2 * cl->G));
    LJ_CUDA(cudaMalloc(&cl->ke_part, sizeof(float) * 2 * cl->G));
    LJ_CUDA(cudaMalloc(&cl->state, sizeof(int) * ST_WORDS));
    LJ_CUDA(cudaMemset(cl->state, 0, sizeof(int) * ST_WORDS));
    LJ_CUDA(cudaMalloc(&cl->bar, sizeof(unsigned)));
    if (getenv("LJMD_CELLS_PROF")) LJ_CUDA(cudaMalloc(&cl->prof, sizeof(long long) * 12 * cl->G));
    return 0;
}

void cells_destroy(ljmd_handle* h) {
    Cells* cl = h->cells;
    if (!cl) return;
    for (int k = 0; k < 2; ++k) { cudaFree(cl->Rs[k]); cudaFree(cl->Vh[k]); cudaFree(cl->orig[k]); }
    cudaFree(cl->Rbuild); cudaFree(cl->key); cudaFree(cl->slot_src); cudaFree(cl->meta);
    cudaFree(cl->nbr2); cudaFree(cl->cell_count); cudaFree(cl->fill);
    cudaFree(cl->cell_start); cudaFree(cl->chunk_tot); cudaFree(cl->pe_part); cudaFree(cl->ke_part);
    cudaFree(cl->state); cudaFree(cl->bar); cudaFree(cl->prof);
    delete cl;
    h->cells = nullptr;
}

static void fill_args(ljmd_handle* h, CellsArgs& a) {
    Cells* cl = h->cells;
    a.pc = h->pc;
    a.N = (int)h->p.N; a.Npad = cl->Npad; a.G = cl->G;
    a.ncell = cl->ncell; a.ncells = cl->ncells;
    a.inv_cell = cl->inv_cell; a.cell_size = cl->cell_size;
    a.rlist2 = cl->rlist * cl->rlist;
    a.half_skin2 = (0.5f * h->p.skin) * (0.5f * h->p.skin);
    a.dt = h->p.dt;
    for (int k = 0; k < 2; ++k) { a.Rs[k] = cl->Rs[k]; a.Vh[k] = cl->Vh[k]; a.orig[k] = cl->orig[k]; }
    a.Rbuild = cl->Rbuild;
    a.key = cl->key; a.slot_src = cl->slot_src; a.cell_count = cl->cell_count;
    a.cell_start = cl->cell_start; a.fill = cl->fill; a.chunk_tot = cl->chunk_tot;
    a.meta = cl->meta; a.nbr2 = cl->nbr2;
    a.pe_part = cl->pe_part; a.ke_part = cl->ke_part;
    a.state = cl->state; a.bar = cl->bar; a.prof = cl->prof;
}

static int launch(ljmd_handle* h, CellsArgs& a) {
    Cells* cl = h->cells;
    LJ_CUDA(cudaMemsetAsync(cl->bar, 0, sizeof(unsigned), h->stream));
    void* args[] = {(void*)&a};
    LJ_CUDA(cudaLaunchCooperativeKernel((void*)cells_persistent_kernel, dim3(cl->G), dim3(CL_THREADS),
                                        args, 0, h->stream));
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
    LJ_CUDA(cudaMemsetAsync(cl->state + ST_FLAG, 0, sizeof(int), st));
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
        const char* nm[10] = {"force+integrate", "step barrier", "R1 clear", "R2 bin+hist", "R3 scan pass1",
                              "R4 scan pass2", "R5 scatter", "R6a cell sort", "R6b gather", "R7 list build"};
        for (int k = 0; k < 10; ++k) {
            double mean = 0, mx = 0;
            for (int c = 0; c < cl->G; ++c) { mean += pv[c * 12 + k]; mx = std::max<double>(mx, (double)pv[c * 12 + k]); }
            fprintf(stderr, "[ljmd cells prof] %-16s mean %12.0f  max %12.0f clocks (launch total)\n", nm[k], mean / cl->G, mx);
        }
    }
    return 0;
}

int cells_geometry(ljmd_handle* h, int* ncell, float* cell, float* inv_cell) {
    Cells* cl = h->cells;
    if (ncell) *ncell = cl->ncell;
    if (cell) *cell = cl->cell_size;
    if (inv_cell) *inv_cell = cl->inv_cell;
    return 0;
}

int cells_assign(ljmd_handle* h, const float2* R, int* cell_id, int* cell_count) {
    Cells* cl = h->cells;
    const int N = (int)h->p.N;
    if (cell_count) LJ_CUDA(cudaMemsetAsync(cell_count, 0, sizeof(int) * (size_t)cl->ncells, h->stream));
    cell_assign_kernel<<<(N + 255) / 256, 256, 0, h->stream>>>(R, N, cl->ncell, cl->inv_cell, cell_id, cell_count);
    LJ_CUDA(cudaGetLastError());
    h->launches++;
    return 0;
}

int cells_neighbor_count(ljmd_handle* h, const float2* R, float radius, int* nbr_count) {
    Cells* cl = h->cells;
    if (!(radius > 0.0f) || radius > cl->cell_size) {
        set_error("neighbor_count radius %.4f must be in (0, cell size %.4f]", radius, cl->cell_size);
        return LJMD_E_INVALID;
    }
    LJ_CUDA(cudaMemsetAsync(cl->state, 0, sizeof(int) * 3, h->stream));
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
    if (e & CERR_LIST_OVERFLOW) {
        set_error("cell-list: a particle has more than %d neighbours within rc+skin", CL_MAXN);
        return LJMD_E_STATE;
    }
    if (e) { set_error("cell-list persistent kernel: grid barrier timed out (flag %d)", e); return LJMD_E_STATE; }
    return 0;
}

}  // namespace ljmd
