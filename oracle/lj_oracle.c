/*
 * lj_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C, strict IEEE fp32)
 * of the hot path of the reference script
 *     /root/reference/molecular_dynamics_jax_single-host_workload.py   (MD:<line>)
 *
 * PARITY PIN: the reference ships no tests, golden vectors or fixtures, and JAX / XLA are not
 * installable in this image, so this restatement cannot be checked against the
 * reference running on JAX.  It is pinned (tests/test_reference_golden.py) to vectors produced by
 * the reference's OWN SOURCE FILE executed unmodified on a torch facade of the jax API
 * (tests/golden/jax_facade.py, make_reference_golden.py), to the independent torch-autodiff
 * restatement (oracle/lj_oracle.py) and to analytic known answers (tests/test_oracle.py).  Not
 * covered by that pin: XLA's own reduction order and pow lowering.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (jax_tpus_benchmark_physics_simulation_b200/) never does.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp).
 * -ffp-contract=off matters: every fp32 operation below must round exactly once, like the
 * elementwise jnp ops of the reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* MD:46-48  periodic_displacement(dr, box) = dr - box * round(dr / box)
 * jnp.round is round-half-to-even == rintf under the default rounding mode. */
static inline float minimg(float d, float box) {
    return d - box * rintf(d / box);
}

float orc_periodic_displacement(float d, float box) { return minimg(d, box); }

/* jnp.mod(x, box) (MD:72): result takes the sign of the divisor; can return exactly `box`
 * for tiny negative x in fp32.  Same algorithm as numpy's npy_divmod. */
static inline float pymodf(float x, float b) {
    float m = fmodf(x, b);
    if (m != 0.0f) {
        if ((b < 0.0f) != (m < 0.0f)) m += b;
    } else {
        m = copysignf(0.0f, b);
    }
    return m;
}
float orc_mod(float x, float b) { return pymodf(x, b); }

/* One ordered pair (i,j), i != j, reference formulation MD:51-59 + closed-form gradient
 * (SURVEY.md §8a4).  sigma = epsilon = 1 folded in by the caller through sig2 / eps. */
typedef struct { float fx, fy, e; int inside; } pair_t;

static inline pair_t pair_term(float xi, float yi, float xj, float yj, float box,
                               float sig2, float eps, float rc2) {
    pair_t o;
    float dx = minimg(xi - xj, box);          /* MD:51-52: subtract first, then wrap */
    float dy = minimg(yi - yj, box);
    float r2 = dx * dx + dy * dy;              /* MD:53 */
    float s2 = sig2 / r2;                      /* MD:56 */
    float s6 = (s2 * s2) * s2;                 /* MD:57  x**3 -> (x*x)*x */
    float s12 = s6 * s6;                       /* MD:58 */
    o.inside = (r2 < rc2);                     /* NOT in reference: plain truncation */
    o.e = (4.0f * eps) * (s12 - s6);           /* MD:59 */
    float fs = ((24.0f * eps) * (2.0f * s12 - s6)) * (s2 / sig2); /* -dE/dr2 * 2 ; sig2==1 -> s2 */
    o.fx = fs * dx;
    o.fy = fs * dy;
    return o;
}

/* force_fn (MD:64) + total_energy_fn (MD:50-62), all-pairs O(N^2).
 *   R (N,2) fp32; F (N,2) fp32 out (may be NULL); pe out (may be NULL).
 *   rc2 = INFINITY reproduces the reference (no cutoff).
 *   acc_double != 0: per-particle sums are accumulated in double and rounded once
 *     (isolates per-pair fp32 arithmetic from summation order);
 *   acc_double == 0: naive sequential fp32 accumulation in j order.
 *   rows [i0, i1) only (for subset checks at large N). */
void orc_forces_rows(int64_t N, const float* R, float box, float sigma, float eps, float rc2,
                     int64_t i0, int64_t i1, float* F, double* pe_out, int acc_double) {
    float sig2 = sigma * sigma;
    double pe_tot = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : pe_tot)
    for (int64_t i = i0; i < i1; ++i) {
        float xi = R[2 * i], yi = R[2 * i + 1];
        double fxd = 0.0, fyd = 0.0, ed = 0.0;
        float fxs = 0.0f, fys = 0.0f;
        for (int64_t j = 0; j < N; ++j) {
            if (j == i) continue;              /* MD:54-55,60 diagonal mask */
            pair_t p = pair_term(xi, yi, R[2 * j], R[2 * j + 1], box, sig2, eps, rc2);
            if (!p.inside) continue;
            if (acc_double) { fxd += p.fx; fyd += p.fy; }
            else            { fxs += p.fx; fys += p.fy; }
            ed += p.e;
        }
        if (F) {
            F[2 * (i - i0)]     = acc_double ? (float)fxd : fxs;
            F[2 * (i - i0) + 1] = acc_double ? (float)fyd : fys;
        }
        pe_tot += ed;
    }
    if (pe_out) *pe_out = 0.5 * pe_tot;        /* MD:61 */
}

void orc_forces(int64_t N, const float* R, float box, float sigma, float eps, float rc2,
                float* F, double* pe_out, int acc_double) {
    orc_forces_rows(N, R, box, sigma, eps, rc2, 0, N, F, pe_out, acc_double);
}

double orc_kinetic(int64_t N, const float* V) {
    double ke = 0.0;
    for (int64_t i = 0; i < 2 * N; ++i) ke += (double)(V[i] * V[i]);
    return 0.5 * ke;
}

/* verlet_step MD:66-75, with F carried (F(R_new) of step n == F(R) of step n+1 bit for bit,
 * SURVEY.md §0).  In-place on R, V, F.  fp32 op order exactly as written in the reference:
 *   V_half = V + (0.5*F)*dt ; R_new = mod(R + V_half*dt, box) ; V_new = V_half + (0.5*F_new)*dt */
void orc_step(int64_t N, float* R, float* V, float* F, float box, float sigma, float eps,
              float rc2, float dt, double* pe_out, int acc_double) {
    for (int64_t k = 0; k < 2 * N; ++k) {
        float vh = V[k] + (0.5f * F[k]) * dt;
        V[k] = vh;
        R[k] = pymodf(R[k] + vh * dt, box);
    }
    orc_forces(N, R, box, sigma, eps, rc2, F, pe_out, acc_double);
    for (int64_t k = 0; k < 2 * N; ++k) V[k] = V[k] + (0.5f * F[k]) * dt;
}

/* equilibrate_fn MD:77-83 / production_fn MD:85-106 (sample rule MD:93-100).
 * traj (S,N,2) may be NULL; ke_pe (ceil(nsteps/energy_every),2) doubles may be NULL.
 * thermostat_kT > 0: V *= sqrt(kT / (KE/N)) after every thermostat_every-th step (new). */
void orc_run(int64_t N, float* R, float* V, float box, float sigma, float eps, float rc2,
             float dt, int64_t nsteps, int64_t sample_every, float* traj,
             int64_t energy_every, double* ke_pe, float thermostat_kT,
             int64_t thermostat_every, int acc_double) {
    float* F = (float*)malloc(sizeof(float) * 2 * N);
    int64_t S = (sample_every > 0) ? nsteps / sample_every : 0;
    if (traj && S > 0) memset(traj, 0, sizeof(float) * 2 * N * S);
    orc_forces(N, R, box, sigma, eps, rc2, F, NULL, acc_double);
    for (int64_t i = 0; i < nsteps; ++i) {
        double pe = 0.0;
        orc_step(N, R, V, F, box, sigma, eps, rc2, dt, &pe, acc_double);
        if (traj && sample_every > 0 && i % sample_every == 0 && i / sample_every < S)
            memcpy(traj + (i / sample_every) * 2 * N, R, sizeof(float) * 2 * N);
        if (ke_pe && energy_every > 0 && i % energy_every == 0) {
            ke_pe[2 * (i / energy_every)] = orc_kinetic(N, V);
            ke_pe[2 * (i / energy_every) + 1] = pe;
        }
        if (thermostat_kT > 0.0f && thermostat_every > 0 && (i + 1) % thermostat_every == 0) {
            float ke = (float)orc_kinetic(N, V);
            float lam = sqrtf(thermostat_kT / (ke / (float)N));
            for (int64_t k = 0; k < 2 * N; ++k) V[k] *= lam;
        }
    }
    free(F);
}

/* ---- cell-list recount (new functionality) -------------------------
 * strip cells: row = min((int)(y * inv_hy), nrows-1), bin = min((int)(x * inv_wx), nbx-1),
 * cell = row * nbx + bin; one fp32 multiply then truncation (coordinates >= 0 so truncation
 * == floor); the min handles a coordinate == box (MD:72 closed interval).                */
void orc_cell_assign(int64_t N, const float* R, int32_t nrows, int32_t nbx, float inv_hy,
                     float inv_wx, int32_t* cell_id, int32_t* cell_count) {
    if (cell_count) memset(cell_count, 0, sizeof(int32_t) * (size_t)nrows * nbx);
    for (int64_t i = 0; i < N; ++i) {
        int32_t cx = (int32_t)(R[2 * i] * inv_wx);
        int32_t cy = (int32_t)(R[2 * i + 1] * inv_hy);
        if (cx > nbx - 1) cx = nbx - 1;
        if (cy > nrows - 1) cy = nrows - 1;
        if (cx < 0) cx = 0;
        if (cy < 0) cy = 0;
        int32_t c = cy * nbx + cx;
        if (cell_id) cell_id[i] = c;
        if (cell_count) cell_count[c]++;
    }
}

/* neighbour recount: nbr_count[i] = #{ j != i : r2_minimg(i,j) < radius^2 }, with r2 computed
 * exactly as pair_term does (subtract, wrap, dx*dx + dy*dy, each rounded once).
 * Brute force over a 3x3 (or wider if needed) stencil of a CPU cell grid built here with
 * cell edge >= radius; independent of the GPU's binning. */
void orc_neighbor_count(int64_t N, const float* R, float box, float radius, int32_t* nbr_count) {
    int32_t nc = (int32_t)floor((double)box / (double)radius);
    if (nc < 1) nc = 1;
    if (nc > 4096) nc = 4096;
    double cell = (double)box / nc;
    int32_t* head = (int32_t*)malloc(sizeof(int32_t) * ((size_t)nc * nc + 1));
    int32_t* cid = (int32_t*)malloc(sizeof(int32_t) * N);
    int32_t* order = (int32_t*)malloc(sizeof(int32_t) * N);
    memset(head, 0, sizeof(int32_t) * ((size_t)nc * nc + 1));
    for (int64_t i = 0; i < N; ++i) {
        int32_t cx = (int32_t)floor((double)R[2 * i] / cell), cy = (int32_t)floor((double)R[2 * i + 1] / cell);
        if (cx >= nc) cx = nc - 1;
        if (cy >= nc) cy = nc - 1;
        if (cx < 0) cx = 0;
        if (cy < 0) cy = 0;
        cid[i] = cy * nc + cx;
        head[cid[i] + 1]++;
    }
    for (int64_t c = 0; c < (int64_t)nc * nc; ++c) head[c + 1] += head[c];
    int32_t* fill = (int32_t*)malloc(sizeof(int32_t) * (size_t)nc * nc);
    memcpy(fill, head, sizeof(int32_t) * (size_t)nc * nc);
    for (int64_t i = 0; i < N; ++i) order[fill[cid[i]]++] = (int32_t)i;
    float r2max = radius * radius;
    /* a stencil of +-2 cells is used (not +-1) so that the recount does not depend on the
     * fp32-vs-double binning of particles sitting exactly on a cell edge */
    int span = (nc >= 5) ? 2 : (nc - 1) / 2;
    int full = (nc < 5);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < N; ++i) {
        float xi = R[2 * i], yi = R[2 * i + 1];
        int32_t cx = cid[i] % nc, cy = cid[i] / nc;
        int32_t cnt = 0;
        if (full) {
            for (int64_t j = 0; j < N; ++j) {
                if (j == i) continue;
                float dx = minimg(xi - R[2 * j], box), dy = minimg(yi - R[2 * j + 1], box);
                float r2 = dx * dx + dy * dy;
                cnt += (r2 < r2max);
            }
        } else {
            for (int oy = -span; oy <= span; ++oy)
                for (int ox = -span; ox <= span; ++ox) {
                    int32_t c = ((cy + oy + nc) % nc) * nc + ((cx + ox + nc) % nc);
                    for (int32_t k = head[c]; k < head[c + 1]; ++k) {
                        int32_t j = order[k];
                        if (j == i) continue;
                        float dx = minimg(xi - R[2 * j], box), dy = minimg(yi - R[2 * j + 1], box);
                        float r2 = dx * dx + dy * dy;
                        cnt += (r2 < r2max);
                    }
                }
        }
        nbr_count[i] = cnt;
    }
    free(head); free(cid); free(order); free(fill);
}

/* Forces with cutoff through a CPU cell grid: same per-pair arithmetic as orc_forces, pairs
 * visited in (cell, original-index) order; per-particle sums in double when acc_double.
 * Used as the oracle for N too large for the O(N^2) loop (configs 4-5). */
void orc_forces_cells(int64_t N, const float* R, float box, float sigma, float eps, float rc,
                      float* F, double* pe_out) {
    int32_t nc = (int32_t)floor((double)box / (double)rc);
    if (nc < 5) { orc_forces(N, R, box, sigma, eps, rc * rc, F, pe_out, 1); return; }
    double cell = (double)box / nc;
    size_t ncc = (size_t)nc * nc;
    int32_t* head = (int32_t*)calloc(ncc + 1, sizeof(int32_t));
    int32_t* cid = (int32_t*)malloc(sizeof(int32_t) * N);
    int32_t* order = (int32_t*)malloc(sizeof(int32_t) * N);
    for (int64_t i = 0; i < N; ++i) {
        int32_t cx = (int32_t)floor((double)R[2 * i] / cell), cy = (int32_t)floor((double)R[2 * i + 1] / cell);
        if (cx >= nc) cx = nc - 1;
        if (cy >= nc) cy = nc - 1;
        if (cx < 0) cx = 0;
        if (cy < 0) cy = 0;
        cid[i] = cy * nc + cx;
        head[cid[i] + 1]++;
    }
    for (size_t c = 0; c < ncc; ++c) head[c + 1] += head[c];
    int32_t* fill = (int32_t*)malloc(sizeof(int32_t) * ncc);
    memcpy(fill, head, sizeof(int32_t) * ncc);
    for (int64_t i = 0; i < N; ++i) order[fill[cid[i]]++] = (int32_t)i;
    float sig2 = sigma * sigma, rc2 = rc * rc;
    double pe_tot = 0.0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : pe_tot)
    for (int64_t i = 0; i < N; ++i) {
        float xi = R[2 * i], yi = R[2 * i + 1];
        int32_t cx = cid[i] % nc, cy = cid[i] / nc;
        double fxd = 0.0, fyd = 0.0, ed = 0.0;
        for (int oy = -2; oy <= 2; ++oy)
            for (int ox = -2; ox <= 2; ++ox) {
                int32_t c = ((cy + oy + nc) % nc) * nc + ((cx + ox + nc) % nc);
                for (int32_t k = head[c]; k < head[c + 1]; ++k) {
                    int32_t j = order[k];
                    if (j == i) continue;
                    pair_t p = pair_term(xi, yi, R[2 * j], R[2 * j + 1], box, sig2, eps, rc2);
                    if (!p.inside) continue;
                    fxd += p.fx; fyd += p.fy; ed += p.e;
                }
            }
        F[2 * i] = (float)fxd;
        F[2 * i + 1] = (float)fyd;
        pe_tot += ed;
    }
    if (pe_out) *pe_out = 0.5 * pe_tot;
    free(head); free(cid); free(order); free(fill);
}

/* g(r) histogram stage, get_histogram MD:117-124: unordered pairs i<j, minimum-image
 * distance sqrt(r2) binned with numpy.histogram semantics on edges = linspace(0, r_max, nbins+1)
 * (uniform bins: index from the scaled value, corrected against the fp32 edges; last bin
 * right-closed; values outside [0, r_max] dropped).  edges: float[nbins+1] supplied by caller. */
void orc_gr_hist(int64_t N, const float* R, float box, int32_t nbins, const float* edges,
                 int64_t* counts) {
    memset(counts, 0, sizeof(int64_t) * nbins);
    float lo = edges[0], hi = edges[nbins];
    for (int64_t i = 0; i < N; ++i)
        for (int64_t j = i + 1; j < N; ++j) {
            float dx = minimg(R[2 * i] - R[2 * j], box), dy = minimg(R[2 * i + 1] - R[2 * j + 1], box);
            float r = sqrtf(dx * dx + dy * dy);
            if (!(r >= lo && r <= hi)) continue;
            /* binary search: largest k with edges[k] <= r ; r == hi -> last bin */
            int32_t a = 0, b = nbins;
            while (b - a > 1) { int32_t m = (a + b) >> 1; if (edges[m] <= r) a = m; else b = m; }
            counts[a]++;
        }
}

/* ---- the other dense pairwise kernels of the reference repo ----------------
 * pairwise_forces(positions, masses), nbody_bh_merger_sim_single-host_workload.py NBODY:54-67:
 *   for i, for j != i (in order): r_vec = pos[j] - pos[i]; r = ||r_vec||;
 *   acc[i] += where(r >= 1e-6, G * m[j] / r**3, 0) * r_vec          (fp32, r**3 = r*r*r)            */
void orc_gravity_nbody(int64_t n, const float* pos, const float* mass, float G, float* acc) {
    for (int64_t i = 0; i < n; ++i) {
        float ax = 0.0f, ay = 0.0f;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i) continue;
            float dx = pos[2 * j] - pos[2 * i], dy = pos[2 * j + 1] - pos[2 * i + 1];
            float r = sqrtf(dx * dx + dy * dy);
            float mag = (r >= 1.0e-6f) ? (G * mass[j]) / ((r * r) * r) : 0.0f;
            ax = ax + mag * dx;
            ay = ay + mag * dy;
        }
        acc[2 * i] = ax; acc[2 * i + 1] = ay;
    }
}
/* gravity term of acceleration(pos, vel, masses, charges), three_particles_em_nonuni EM3:25-38:
 *   r_diff = pos[j] - pos[i]; r2 = sum(r_diff^2) + eye; r2 = where(r2 < 1e-12, 1e-12, r2);
 *   acc[i] = sum_j G * m[j] * r_diff * r2**(-1.5)                                                   */
void orc_gravity_em3(int64_t n, const float* pos, const float* mass, float G, float* acc) {
    for (int64_t i = 0; i < n; ++i) {
        float ax = 0.0f, ay = 0.0f;
        for (int64_t j = 0; j < n; ++j) {
            float dx = pos[2 * j] - pos[2 * i], dy = pos[2 * j + 1] - pos[2 * i + 1];
            float r2 = dx * dx + dy * dy + (i == j ? 1.0f : 0.0f);
            if (r2 < 1.0e-12f) r2 = 1.0e-12f;
            float inv3 = powf(r2, -1.5f);
            ax = ax + ((G * mass[j]) * dx) * inv3;
            ay = ay + ((G * mass[j]) * dy) * inv3;
        }
        acc[2 * i] = ax; acc[2 * i + 1] = ay;
    }
}
