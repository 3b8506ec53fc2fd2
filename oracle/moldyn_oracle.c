/*
 * moldyn_oracle.c — CPU restatement of AndrewChe7/moldyn's `solve` step loop.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under moldyn_b200/ (the product) may link,
 * import or execute this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * The reference is Rust and cannot be built in this image (no cargo/rustc, no
 * network), so this is a restatement, operation by operation, of the cited
 * reference lines (paths relative to /root/reference).  It is pinned against
 * every golden value of the reference's own tests (tests/golden/reference_kats.json,
 * tests/test_oracle_golden.py).  Many-body forces, thermostat/barostat
 * trajectories, FCC init and non-default r_cut are NOT pinned by any enabled
 * reference test ("parity unpinned" for those).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared
 *   (no FMA contraction and no reassociation: Rust never contracts or
 *    reassociates f64 arithmetic).
 *
 * Data model: arrays are xyz-interleaved: pos[3*i+{0,1,2}] like Vec<Vector3<f64>>.  The orc_* functions of the first part
 * take one particle type; the orc_*_multi functions at the end take State.particles flattened type by type
 * (start[t] .. start[t+1]) and restate what the reference does with several types, quirks included (see there).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* core/src/lib.rs:15 */
static const double K_B = 1.380648528;

typedef struct {
    double sigma, eps, r_cut, u_cut;
} orc_lj;

/* ------------------------------------------------------------------------- */
/* Potential::get_potential_and_force — solver/src/solver/potential.rs:57-70  */
/* sigma_r.pow(6) is f64::powi(6) (num_traits Pow<i32>): square-and-multiply  */
/* x^2 * x^4 (compiler-rt __powidf2 / LLVM powi expansion give this order).   */
ORC_API void orc_lj_potential_and_force(const orc_lj *p, double r, double *u, double *f)
{
    if (r > p->r_cut) {
        *u = 0.0;
        *f = 0.0;
        return;
    }
    double sigma_r = p->sigma / r;
    double x2 = sigma_r * sigma_r;
    double x4 = x2 * x2;
    double sigma_r_6 = x2 * x4;
    double sigma_r_12 = sigma_r_6 * sigma_r_6;
    *u = 4.0 * p->eps * (sigma_r_12 - sigma_r_6) - p->u_cut;
    *f = (24.0 * p->eps / r) * (sigma_r_6 - 2.0 * sigma_r_12);
}

/* Potential::new_lennard_jones — potential.rs:27-55: r_cut = sigma*2.5, u_cut = U(r_cut) with u_cut=0 */
ORC_API void orc_lj_new(double sigma, double eps, orc_lj *out)
{
    out->sigma = sigma;
    out->eps = eps;
    out->r_cut = sigma * 2.5;
    out->u_cut = 0.0;
    double u, f;
    orc_lj_potential_and_force(out, out->r_cut, &u, &f);
    out->u_cut = u;
}

/* Inner body of update_force for one ordered pair — potential.rs:181-211.
 * Returns 1 if the pair is inside the (inclusive) cutoff. */
static inline int pair_term(const orc_lj *p, const double *pi, const double *pj, const double *bb,
                            double *fx, double *fy, double *fz, double *u, double *t)
{
    double rx = pj[0] - pi[0];
    double ry = pj[1] - pi[1];
    double rz = pj[2] - pi[2];
    if (rx < -bb[0] / 2.0) rx += bb[0]; else if (rx > bb[0] / 2.0) rx -= bb[0];
    if (ry < -bb[1] / 2.0) ry += bb[1]; else if (ry > bb[1] / 2.0) ry -= bb[1];
    if (rz < -bb[2] / 2.0) rz += bb[2]; else if (rz > bb[2] / 2.0) rz -= bb[2];
    /* nalgebra Vector3::norm(): sqrt(x*x + y*y + z*z), summed left to right */
    double r_abs = sqrt((rx * rx + ry * ry) + rz * rz);
    if (r_abs > p->r_cut) return 0;
    double pot, force;
    orc_lj_potential_and_force(p, r_abs, &pot, &force);
    /* force_vec = r / r_abs * force : component-wise divide, then component-wise multiply */
    double vx = rx / r_abs * force;
    double vy = ry / r_abs * force;
    double vz = rz / r_abs * force;
    *t = vx * rx + vy * ry + vz * rz;
    *fx = vx; *fy = vy; *fz = vz; *u = pot;
    return 1;
}

/* update_force — potential.rs:158-216, rows [i0,i1) only (i0=0,i1=n is the full call).
 * Parallel over i exactly like rayon's par_iter_mut (each i summed by one task in ascending j). */
ORC_API void orc_update_force_rows(const orc_lj *p, int64_t n, const double *pos, const double *bb,
                                   int64_t i0, int64_t i1, double *force, double *pot, double *vir)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = i0; i < i1; ++i) {
        double fx = 0.0, fy = 0.0, fz = 0.0, u = 0.0, w = 0.0;
        const double *pi = pos + 3 * i;
        for (int64_t j = 0; j < n; ++j) {
            if (i == j) continue;
            double vx, vy, vz, pu, pt;
            if (!pair_term(p, pi, pos + 3 * j, bb, &vx, &vy, &vz, &pu, &pt)) continue;
            fx += vx; fy += vy; fz += vz; u += pu; w += pt;
        }
        force[3 * i] = fx; force[3 * i + 1] = fy; force[3 * i + 2] = fz;
        pot[i] = u; vir[i] = w;
    }
}

ORC_API void orc_update_force(const orc_lj *p, int64_t n, const double *pos, const double *bb,
                              double *force, double *pot, double *vir)
{
    orc_update_force_rows(p, n, pos, bb, 0, n, force, pot, vir);
}

/* ------------------------------------------------------------------------- */
/* Θ(N) variant of update_force: candidates come from a cell grid, but each i  */
/* still sums its in-range partners in ASCENDING j with pair_term(), so the    */
/* result is bit-identical to the Θ(N²) scan (out-of-range j contribute        */
/* nothing there).  "Improved CPU algorithm — not the reference's"; validated  */
/* against orc_update_force in tests.  r_search >= r_cut selects candidates.   */
static int cmp_i64(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

static inline int64_t wrap_cell(double x, double L, int64_t nc)
{
    double s = x / L;
    s -= floor(s);
    int64_t c = (int64_t)(s * (double)nc);
    if (c >= nc) c = nc - 1;
    if (c < 0) c = 0;
    return c;
}

typedef struct {
    int64_t nc[3];
    int64_t *start; /* ncell+1 */
    int64_t *items; /* n, atom indices grouped by cell, ascending inside a cell */
} cellgrid;

static void cellgrid_build(cellgrid *g, int64_t n, const double *pos, const double *bb, double r_search)
{
    for (int d = 0; d < 3; ++d) {
        int64_t c = (int64_t)floor(bb[d] / r_search);
        if (c < 1) c = 1;
        /* keep the grid small for dilute systems */
        int64_t cap = (int64_t)ceil(cbrt((double)n)) + 1;
        if (c > cap) c = cap;
        g->nc[d] = c;
    }
    int64_t ncell = g->nc[0] * g->nc[1] * g->nc[2];
    g->start = (int64_t *)calloc((size_t)ncell + 1, sizeof(int64_t));
    g->items = (int64_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
    int64_t *cid = (int64_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) {
        int64_t cx = wrap_cell(pos[3 * i], bb[0], g->nc[0]);
        int64_t cy = wrap_cell(pos[3 * i + 1], bb[1], g->nc[1]);
        int64_t cz = wrap_cell(pos[3 * i + 2], bb[2], g->nc[2]);
        cid[i] = (cx * g->nc[1] + cy) * g->nc[2] + cz;
        g->start[cid[i] + 1]++;
    }
    for (int64_t c = 0; c < ncell; ++c) g->start[c + 1] += g->start[c];
    int64_t *fill = (int64_t *)malloc((size_t)ncell * sizeof(int64_t));
    memcpy(fill, g->start, (size_t)ncell * sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) g->items[fill[cid[i]]++] = i;
    free(fill);
    free(cid);
}

static void cellgrid_free(cellgrid *g)
{
    free(g->start);
    free(g->items);
}

/* Collect the (deduplicated) neighbour cells of the cell that holds `pi`. */
static int64_t neighbour_cells(const cellgrid *g, const double *pi, const double *bb, int64_t *out)
{
    int64_t c[3], lo[3], cnt[3];
    for (int d = 0; d < 3; ++d) {
        c[d] = wrap_cell(pi[d], bb[d], g->nc[d]);
        if (g->nc[d] >= 3) { lo[d] = c[d] - 1; cnt[d] = 3; }
        else { lo[d] = 0; cnt[d] = g->nc[d]; }
    }
    int64_t k = 0;
    for (int64_t a = 0; a < cnt[0]; ++a)
        for (int64_t b = 0; b < cnt[1]; ++b)
            for (int64_t e = 0; e < cnt[2]; ++e) {
                int64_t cx = (lo[0] + a + g->nc[0]) % g->nc[0];
                int64_t cy = (lo[1] + b + g->nc[1]) % g->nc[1];
                int64_t cz = (lo[2] + e + g->nc[2]) % g->nc[2];
                out[k++] = (cx * g->nc[1] + cy) * g->nc[2] + cz;
            }
    return k;
}

ORC_API void orc_update_force_cells(const orc_lj *p, int64_t n, const double *pos, const double *bb,
                                    double *force, double *pot, double *vir)
{
    cellgrid g;
    cellgrid_build(&g, n, pos, bb, p->r_cut);
#pragma omp parallel
    {
        int64_t cap = 1024, *cand = (int64_t *)malloc((size_t)cap * sizeof(int64_t));
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < n; ++i) {
            int64_t cells[27];
            int64_t ncells = neighbour_cells(&g, pos + 3 * i, bb, cells);
            int64_t m = 0;
            for (int64_t q = 0; q < ncells; ++q) {
                int64_t s = g.start[cells[q]], e = g.start[cells[q] + 1];
                if (m + (e - s) > cap) {
                    while (m + (e - s) > cap) cap *= 2;
                    cand = (int64_t *)realloc(cand, (size_t)cap * sizeof(int64_t));
                }
                for (int64_t t = s; t < e; ++t) cand[m++] = g.items[t];
            }
            qsort(cand, (size_t)m, sizeof(int64_t), cmp_i64);
            double fx = 0.0, fy = 0.0, fz = 0.0, u = 0.0, w = 0.0;
            for (int64_t q = 0; q < m; ++q) {
                int64_t j = cand[q];
                if (j == i) continue;
                double vx, vy, vz, pu, pt;
                if (!pair_term(p, pos + 3 * i, pos + 3 * j, bb, &vx, &vy, &vz, &pu, &pt)) continue;
                fx += vx; fy += vy; fz += vz; u += pu; w += pt;
            }
            force[3 * i] = fx; force[3 * i + 1] = fy; force[3 * i + 2] = fz;
            pot[i] = u; vir[i] = w;
        }
        free(cand);
    }
    cellgrid_free(&g);
}

/* Neighbour-set oracle: for every i the ascending list of j != i whose
 * reference min-image distance (potential.rs:181-201) is <= r_list.
 * counts[i] receives the count; if nbr != NULL, nbr[offsets[i] + k] the partners
 * (offsets = exclusive prefix of counts, supplied by the caller on the 2nd call).
 * Cell-list accelerated; the predicate is the reference's exact arithmetic. */
ORC_API void orc_neighbour_sets(int64_t n, const double *pos, const double *bb, double r_list,
                                int64_t *counts, const int64_t *offsets, int64_t *nbr)
{
    orc_lj p = {1.0, 1.0, r_list, 0.0};
    cellgrid g;
    cellgrid_build(&g, n, pos, bb, r_list);
#pragma omp parallel
    {
        int64_t cap = 1024, *cand = (int64_t *)malloc((size_t)cap * sizeof(int64_t));
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < n; ++i) {
            int64_t cells[27];
            int64_t ncells = neighbour_cells(&g, pos + 3 * i, bb, cells);
            int64_t m = 0;
            for (int64_t q = 0; q < ncells; ++q) {
                int64_t s = g.start[cells[q]], e = g.start[cells[q] + 1];
                if (m + (e - s) > cap) {
                    while (m + (e - s) > cap) cap *= 2;
                    cand = (int64_t *)realloc(cand, (size_t)cap * sizeof(int64_t));
                }
                for (int64_t t = s; t < e; ++t) cand[m++] = g.items[t];
            }
            qsort(cand, (size_t)m, sizeof(int64_t), cmp_i64);
            int64_t c = 0;
            for (int64_t q = 0; q < m; ++q) {
                int64_t j = cand[q];
                if (j == i) continue;
                double vx, vy, vz, pu, pt;
                if (!pair_term(&p, pos + 3 * i, pos + 3 * j, bb, &vx, &vy, &vz, &pu, &pt)) continue;
                if (nbr) nbr[offsets[i] + c] = j;
                ++c;
            }
            counts[i] = c;
        }
        free(cand);
    }
    cellgrid_free(&g);
}

/* ------------------------------------------------------------------------- */
/* State::apply_boundary_conditions — core/src/particle.rs:120-142             */
ORC_API void orc_apply_boundary_conditions(int64_t n, double *pos, const double *bb)
{
    for (int64_t i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) {
            double *x = &pos[3 * i + d];
            if (*x < 0.0) *x += bb[d];
            else if (*x >= bb[d]) *x -= bb[d];
        }
}

/* get_center_of_mass_velocity — solver/src/macro_parameters/mod.rs:12-25      */
ORC_API void orc_center_of_mass_velocity(int64_t n, const double *vel, double mass, double *out)
{
    double sx = 0.0, sy = 0.0, sz = 0.0, sw = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        sx += vel[3 * i] * mass;
        sy += vel[3 * i + 1] * mass;
        sz += vel[3 * i + 2] * mass;
        sw += 1.0 * mass;
    }
    out[0] = sx / sw; out[1] = sy / sw; out[2] = sz / sw;
}

/* get_momentum_of_system — mod.rs:28-34 */
ORC_API void orc_momentum(int64_t n, const double *vel, double mass, double *out)
{
    double px = 0.0, py = 0.0, pz = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        px += vel[3 * i] * mass; py += vel[3 * i + 1] * mass; pz += vel[3 * i + 2] * mass;
    }
    out[0] = px; out[1] = py; out[2] = pz;
}

/* get_kinetic_energy — macro_parameters/energy.rs:4-22 (nalgebra dot: (x*x + y*y) + z*z) */
ORC_API double orc_kinetic_energy(int64_t n, const double *vel, double mass)
{
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        const double *v = vel + 3 * i;
        s += mass * ((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]) / 2.0;
    }
    return s;
}

/* get_thermal_energy — energy.rs:8-37 */
ORC_API double orc_thermal_energy(int64_t n, const double *vel, double mass, const double *vcom)
{
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double a = vel[3 * i] - vcom[0], b = vel[3 * i + 1] - vcom[1], c = vel[3 * i + 2] - vcom[2];
        s += mass * ((a * a + b * b) + c * c) / 2.0;
    }
    return s;
}

/* get_potential_energy — energy.rs:40-49 */
ORC_API double orc_potential_energy(int64_t n, const double *pot)
{
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += pot[i];
    return s / 2.0;
}

/* get_temperature — macro_parameters/temperature.rs:4-7 (Kelvin) */
ORC_API double orc_temperature(double thermal_energy, int64_t n)
{
    double t = (2.0 * thermal_energy) / (3.0 * (double)n * K_B);
    return t * 100.0;
}

/* get_pressure — macro_parameters/pressure.rs:5-20 */
ORC_API double orc_pressure(int64_t n, const double *vel, const double *vir, double mass,
                            const double *bb, const double *vcom)
{
    double volume = bb[0] * bb[1] * bb[2];
    double result1 = 0.0, result2 = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double dx = vel[3 * i] - vcom[0], dy = vel[3 * i + 1] - vcom[1], dz = vel[3 * i + 2] - vcom[2];
        result1 += mass * dx * dx;
        result1 += mass * dy * dy;
        result1 += mass * dz * dz;
        result2 -= vir[i];
    }
    return (result1 + result2 * 0.5) / volume / 3.0;
}

/* ------------------------------------------------------------------------- */
/* Integrator::calculate (VerletMethod) — solver/src/solver/integrator.rs:14-59 */
/* thermostat kinds: 0 none, 1 Berendsen (thermostat.rs:31-34,54-58),          */
/*                   2 Nose-Hoover (thermostat.rs:35-39,59-65)                 */
/* barostat kinds:   0 none, 1 Berendsen (barostat.rs:21-49)                   */
typedef struct {
    int kind;
    double tau, target, lambda, psi;
} orc_thermostat;

typedef struct {
    int kind;
    double beta, tau, target, myu;
} orc_barostat;

typedef struct {
    int64_t n;
    double mass;
    double *pos, *vel, *force, *pot, *vir; /* caller-owned */
    double bb[3];
} orc_state;

static double state_temperature(const orc_state *s)
{
    double mv[3];
    orc_center_of_mass_velocity(s->n, s->vel, s->mass, mv);
    double e = orc_thermal_energy(s->n, s->vel, s->mass, mv);
    return orc_temperature(e, s->n);
}

/* force_mode: 0 = Θ(N²) reference scan, 1 = cell-list (bit-identical sums) */
ORC_API void orc_step(const orc_lj *p, orc_state *s, double dt, orc_thermostat *th, orc_barostat *ba,
                      int force_mode)
{
    int64_t n = s->n;
    if (ba && ba->kind == 1) { /* barostat.rs:21-31 */
        double mv[3];
        orc_center_of_mass_velocity(n, s->vel, s->mass, mv);
        double pressure = orc_pressure(n, s->vel, s->vir, s->mass, s->bb, mv);
        double myu_cubed = 1.0 + dt * ba->beta / ba->tau * (pressure - ba->target);
        ba->myu = cbrt(myu_cubed);
    }
    if (th && th->kind) { /* thermostat.rs:24-44 */
        double temperature = state_temperature(s);
        if (th->kind == 1) {
            double lambda_squared = 1.0 + dt / th->tau * (th->target / temperature - 1.0);
            th->lambda = sqrt(lambda_squared);
        } else {
            double psi_dot = -((th->target / temperature) - 1.0) / th->tau;
            th->psi += psi_dot * (dt / 2.0);
            th->lambda = exp(-th->psi * dt / 2.0);
        }
    }
    double temp = dt / (2.0 * s->mass); /* integrator.rs:29-30 */
    for (int64_t k = 0; k < 3 * n; ++k) s->vel[k] = s->vel[k] + s->force[k] * temp;
    if (th && th->kind) { /* thermostat.rs:47-73 */
        double temperature = state_temperature(s); /* recomputed from the kicked state */
        for (int64_t k = 0; k < 3 * n; ++k) s->vel[k] *= th->lambda;
        if (th->kind == 2) {
            double psi_dot = -((th->target / temperature) - 1.0) / th->tau;
            th->psi += psi_dot * (dt / 2.0);
        }
    }
    for (int64_t k = 0; k < 3 * n; ++k) s->pos[k] += s->vel[k] * dt; /* integrator.rs:40-44 */
    orc_apply_boundary_conditions(n, s->pos, s->bb);
    if (force_mode == 0) orc_update_force(p, n, s->pos, s->bb, s->force, s->pot, s->vir);
    else orc_update_force_cells(p, n, s->pos, s->bb, s->force, s->pot, s->vir);
    for (int64_t k = 0; k < 3 * n; ++k) s->vel[k] += s->force[k] * temp; /* integrator.rs:47-53 */
    if (ba && ba->kind == 1) { /* barostat.rs:39-49 (one particle type → one scaling) */
        s->bb[0] *= ba->myu; s->bb[1] *= ba->myu; s->bb[2] *= ba->myu;
        for (int64_t k = 0; k < 3 * n; ++k) s->pos[k] *= ba->myu;
    }
}

/* ------------------------------------------------------------------------- */
/* Several particle types.  State.particles is Vec<Vec<Particle>> indexed by type id (core/src/particle.rs:24-32); here it is
 * flattened type by type: type t owns atoms [start[t], start[t+1]), mass[t] is Particle.mass of that type, and table[t1*T+t2]
 * is PotentialsDatabase::get_potential(t1, t2) = the entry keyed (min, max) (potential.rs:147-155), so table is symmetric.
 *
 * What the reference does with T > 1, restated literally:
 *   update_force (potential.rs:158-216): `for t1 in 0..T { for t2 in t1..T {` — atoms of type t1 accumulate the terms of
 *     their type-t2 partners, j ascending, on top of what the earlier t2 left in particle.force/potential/temp; atoms of type
 *     t2 > t1 receive NOTHING from type t1 (no second pass, no Newton's third law).  `symmetric` = 1 lets t2 run over 0..T
 *     instead (every atom accumulates from every type in ascending flattened order): the physically meaningful variant.
 *   Integrator::calculate (integrator.rs:14-59): calculate_myu and calculate_lambda are called once per type and overwrite the
 *     same myu / lambda (Nose-Hoover: psi accumulates), so the coefficients of the LAST type are the ones applied — to every
 *     type; the kicks use the type's own mass (particle_type[0].mass); barostat.update runs once per type and scales the
 *     box every time: box *= myu^T while every position is scaled once.
 * A type without atoms makes the reference panic (index 0 of an empty Vec, integrator.rs:29): callers must not pass one. */
ORC_API void orc_update_force_multi(int T, const int64_t *start, const orc_lj *table, const double *pos, const double *bb,
                                    int symmetric, double *force, double *pot, double *vir)
{
    for (int t1 = 0; t1 < T; ++t1) {
#pragma omp parallel for schedule(dynamic, 64)
        for (int64_t i = start[t1]; i < start[t1 + 1]; ++i) {
            double fx = 0.0, fy = 0.0, fz = 0.0, u = 0.0, w = 0.0;
            const double *pi = pos + 3 * i;
            for (int t2 = symmetric ? 0 : t1; t2 < T; ++t2) {
                const orc_lj *p = &table[t1 * T + t2];
                for (int64_t j = start[t2]; j < start[t2 + 1]; ++j) {
                    if (i == j) continue; /* potential.rs:178-180 (same type, same index) */
                    double vx, vy, vz, pu, pt;
                    if (!pair_term(p, pi, pos + 3 * j, bb, &vx, &vy, &vz, &pu, &pt)) continue;
                    fx += vx; fy += vy; fz += vz; u += pu; w += pt;
                }
            }
            force[3 * i] = fx; force[3 * i + 1] = fy; force[3 * i + 2] = fz;
            pot[i] = u; vir[i] = w;
        }
    }
}

typedef struct {
    int T;
    const int64_t *start;  /* T + 1 offsets */
    const double *mass;    /* T */
    double *pos, *vel, *force, *pot, *vir; /* caller-owned, start[T] atoms */
    double bb[3];
} orc_state_multi;

static double type_temperature(const orc_state_multi *s, int t)
{
    int64_t a = s->start[t], n = s->start[t + 1] - a;
    double mv[3];
    orc_center_of_mass_velocity(n, s->vel + 3 * a, s->mass[t], mv);
    return orc_temperature(orc_thermal_energy(n, s->vel + 3 * a, s->mass[t], mv), n);
}

/* per-type macro parameters (the reference's functions all take a particle_type_id) */
ORC_API void orc_macro_type(const orc_state_multi *s, int t, double *out /* ke, thermal, pe, T, P, vcom[3] */)
{
    int64_t a = s->start[t], n = s->start[t + 1] - a;
    double mv[3];
    orc_center_of_mass_velocity(n, s->vel + 3 * a, s->mass[t], mv);
    out[0] = orc_kinetic_energy(n, s->vel + 3 * a, s->mass[t]);
    out[1] = orc_thermal_energy(n, s->vel + 3 * a, s->mass[t], mv);
    out[2] = orc_potential_energy(n, s->pot + a);
    out[3] = orc_temperature(out[1], n);
    out[4] = orc_pressure(n, s->vel + 3 * a, s->vir + a, s->mass[t], s->bb, mv);
    out[5] = mv[0]; out[6] = mv[1]; out[7] = mv[2];
}

ORC_API void orc_step_multi(const orc_lj *table, orc_state_multi *s, double dt, orc_thermostat *th, orc_barostat *ba,
                            int symmetric)
{
    const int T = s->T;
    const int64_t n = s->start[T];
    if (ba && ba->kind == 1) /* integrator.rs:18-22: once per type, the last one stays */
        for (int t = 0; t < T; ++t) {
            int64_t a = s->start[t], nt = s->start[t + 1] - a;
            double mv[3];
            orc_center_of_mass_velocity(nt, s->vel + 3 * a, s->mass[t], mv);
            double pressure = orc_pressure(nt, s->vel + 3 * a, s->vir + a, s->mass[t], s->bb, mv);
            double myu_cubed = 1.0 + dt * ba->beta / ba->tau * (pressure - ba->target);
            ba->myu = cbrt(myu_cubed);
        }
    if (th && th->kind) /* integrator.rs:23-27 */
        for (int t = 0; t < T; ++t) {
            double temperature = type_temperature(s, t);
            if (th->kind == 1) {
                double lambda_squared = 1.0 + dt / th->tau * (th->target / temperature - 1.0);
                th->lambda = sqrt(lambda_squared);
            } else {
                double psi_dot = -((th->target / temperature) - 1.0) / th->tau;
                th->psi += psi_dot * (dt / 2.0);
                th->lambda = exp(-th->psi * dt / 2.0);
            }
        }
    for (int t = 0; t < T; ++t) { /* integrator.rs:28-34 */
        double temp = dt / (2.0 * s->mass[t]);
        for (int64_t k = 3 * s->start[t]; k < 3 * s->start[t + 1]; ++k) s->vel[k] = s->vel[k] + s->force[k] * temp;
    }
    if (th && th->kind) /* integrator.rs:35-39 → thermostat.rs:47-73, type by type */
        for (int t = 0; t < T; ++t) {
            double temperature = type_temperature(s, t);
            for (int64_t k = 3 * s->start[t]; k < 3 * s->start[t + 1]; ++k) s->vel[k] *= th->lambda;
            if (th->kind == 2) {
                double psi_dot = -((th->target / temperature) - 1.0) / th->tau;
                th->psi += psi_dot * (dt / 2.0);
            }
        }
    for (int64_t k = 0; k < 3 * n; ++k) s->pos[k] += s->vel[k] * dt;
    orc_apply_boundary_conditions(n, s->pos, s->bb);
    orc_update_force_multi(T, s->start, table, s->pos, s->bb, symmetric, s->force, s->pot, s->vir);
    for (int t = 0; t < T; ++t) { /* integrator.rs:47-53 */
        double temp = dt / (2.0 * s->mass[t]);
        for (int64_t k = 3 * s->start[t]; k < 3 * s->start[t + 1]; ++k) s->vel[k] += s->force[k] * temp;
    }
    if (ba && ba->kind == 1) /* integrator.rs:54-58 → barostat.rs:39-49: the box is scaled once PER TYPE */
        for (int t = 0; t < T; ++t) {
            s->bb[0] *= ba->myu; s->bb[1] *= ba->myu; s->bb[2] *= ba->myu;
            for (int64_t k = 3 * s->start[t]; k < 3 * s->start[t + 1]; ++k) s->pos[k] *= ba->myu;
        }
}

/* ------------------------------------------------------------------------- */
/* Input generators (solver/src/initializer/position.rs:24-104, velocity.rs:6-29) */
/* kind 0 = UnitCell::U (index = x*sy*sz + y*sz + z), 1 = UnitCell::FCC        */
ORC_API int64_t orc_init_positions(int kind, int64_t sx, int64_t sy, int64_t sz, double cell,
                                   const double *start, double *pos)
{
    int64_t k = 0;
    for (int64_t x = 0; x < sx; ++x)
        for (int64_t y = 0; y < sy; ++y)
            for (int64_t z = 0; z < sz; ++z) {
                double fx = (double)x, fy = (double)y, fz = (double)z;
                if (kind == 0) {
                    pos[3 * k] = start[0] + fx * cell;
                    pos[3 * k + 1] = start[1] + fy * cell;
                    pos[3 * k + 2] = start[2] + fz * cell;
                    ++k;
                } else {
                    pos[3 * k] = start[0] + fx * cell;
                    pos[3 * k + 1] = start[1] + fy * cell;
                    pos[3 * k + 2] = start[2] + fz * cell;
                    ++k;
                    pos[3 * k] = start[0] + fx * cell;
                    pos[3 * k + 1] = start[1] + (fy + 0.5) * cell;
                    pos[3 * k + 2] = start[2] + (fz + 0.5) * cell;
                    ++k;
                    pos[3 * k] = start[0] + (fx + 0.5) * cell;
                    pos[3 * k + 1] = start[1] + fy * cell;
                    pos[3 * k + 2] = start[2] + (fz + 0.5) * cell;
                    ++k;
                    pos[3 * k] = start[0] + (fx + 0.5) * cell;
                    pos[3 * k + 1] = start[1] + (fy + 0.5) * cell;
                    pos[3 * k + 2] = start[2] + fz * cell;
                    ++k;
                }
            }
    return k;
}

/* The reference draws from an unseeded thread_rng (velocity.rs:7), so its values are
 * not reproducible; the oracle uses its own seeded generator (splitmix64 → xoshiro256++,
 * Marsaglia polar normals) and keeps the reference's structure: sigma_v = sqrt(K_B*T*0.01/m),
 * first N/2 drawn, second N/2 the negated copy (velocity.rs:12-28). */
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

typedef struct { uint64_t s[4]; } xo_rng;

static void xo_seed(xo_rng *r, uint64_t seed)
{
    for (int i = 0; i < 4; ++i) {
        uint64_t z = (seed += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        r->s[i] = z ^ (z >> 31);
    }
}

static uint64_t xo_next(xo_rng *r)
{
    uint64_t *s = r->s, result = rotl64(s[0] + s[3], 23) + s[0], t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
    return result;
}

static double xo_uniform(xo_rng *r) { return (double)(xo_next(r) >> 11) * (1.0 / 9007199254740992.0); }

static double xo_normal(xo_rng *r)
{
    for (;;) {
        double u = 2.0 * xo_uniform(r) - 1.0, v = 2.0 * xo_uniform(r) - 1.0, q = u * u + v * v;
        if (q > 0.0 && q < 1.0) return u * sqrt(-2.0 * log(q) / q);
    }
}

ORC_API void orc_init_velocities(int64_t n, double temperature_kelvin, double mass, uint64_t seed, double *vel)
{
    xo_rng r;
    xo_seed(&r, seed);
    double temperature = temperature_kelvin * 0.01;
    double sigma = sqrt(K_B * temperature / mass);
    int64_t half = n / 2;
    for (int64_t i = 0; i < half; ++i) {
        double x = sigma * xo_normal(&r), y = sigma * xo_normal(&r), z = sigma * xo_normal(&r);
        vel[3 * i] = x; vel[3 * i + 1] = y; vel[3 * i + 2] = z;
        vel[3 * (i + half)] = -x; vel[3 * (i + half) + 1] = -y; vel[3 * (i + half) + 2] = -z;
    }
}

/* randomize_positions analogue (position.rs:140-151; the reference's StdRng stream is
 * not reproducible outside the rand crate, so this is shape-only: uniform in the box). */
ORC_API void orc_random_positions(int64_t n, const double *bb, uint64_t seed, double *pos)
{
    xo_rng r;
    xo_seed(&r, seed);
    for (int64_t i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) pos[3 * i + d] = xo_uniform(&r) * bb[d];
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int t)
{
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}
