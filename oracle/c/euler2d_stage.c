/*
 * ORACLE -- test infrastructure, not product code.
 *
 * Plain C (OpenMP) restatement of the reference's INVISCID per-stage 2D Euler DFR right-hand
 * side and SSP-RK(5,4) step, organised the way the reference runs it on the CPU: the same
 * phases with a barrier between them (StepWorker, model_problems/Euler2D/euler.go:420-653),
 * materialised Q_Face / EdgeStore / F_RT_DOF / RHSQ arrays, operator applications as
 * [m x n]·[n x K] products, threads standing in for the goroutine partitions.  It consumes the
 * same flat dfr2d_problem as the device library (include/dfr2d.h).
 *
 * Uses: (1) second, independently written checker next to oracle/euler2d_oracle.py (the two
 * are compared in tests/test_c_oracle.py); (2) the CPU baseline bench.py times on all host
 * cores ("port": the Go toolchain is absent, this is NOT the Go solver).
 * Parity status: pinned only through the numpy oracle and the reference KATs it passes
 * (tests/test_oracle_kats.py); see DESIGN.md section 5.  PerssonC0 dissipation is not restated
 * here (ora_create refuses it) -- the numpy oracle covers that path.
 *
 * Reference citations are file:line under model_problems/Euler2D/ unless noted.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "dfr2d.h"

typedef struct ora {
    dfr2d_problem p;
    int N, NpInt, NpEdge, NpFlux;
    long K, NE;
    /* owned copies of the problem arrays */
    double *FEI, *DivInt, *Jdet, *Jinv, *nxL, *nyL, *IInII, *hK, *bpx, *bpy;
    int *kL, *kR, *numL, *numR, *nconn, *bc, *etoe, *bp_of_edge;
    /* state (euler.go:320-406) */
    double *Q[5], *Residual, *RHSQ, *Q_Face, *F_RT_DOF, *DT;
    double *EdgeFlux, *EdgeQ, *Agg; /* EdgeStore: [4][NE][NpEdge] x2, Aggregates [NE] */
    double Time, GlobalDT;
    long StepCount;
} ora;

static double *dcopy(const double *s, long n) {
    double *d = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (s) memcpy(d, s, sizeof(double) * (size_t)n);
    return d;
}
static int *icopy(const int32_t *s, long n) {
    int *d = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (long i = 0; i < n; i++) d[i] = s[i];
    return d;
}

/* SSP54 coefficients, euler.go:511-563 */
static const double RK_A = 0.391752226571890;
static const double RK1[3] = {0.444370493651235, 0.555629506348765, 0.368410593050371};
static const double RK2[3] = {0.620101851488403, 0.379898148511597, 0.251891774271694};
static const double RK3[3] = {0.178079954393132, 0.821920045606868, 0.544974750228521};
static const double RK4[5] = {0.517231671970585, 0.096059710526146, 0.386708617503269, 0.063692468666290,
                              0.226007483236906};

/* fluids.go:289-320 */
static inline double pressure(double g, const double *q) {
    double oorho = 1.0 / q[0];
    double u = q[1] * oorho, v = q[2] * oorho;
    double u2 = u * u + v * v;
    double qq = 0.5 * q[0] * u2;
    return (g - 1.0) * (q[3] - qq);
}
static inline double sound_speed(double g, const double *q) {
    return sqrt(fabs(g * pressure(g, q) * (1.0 / q[0])));
}
/* fluxes.go:76-87 FluxCalcBase */
static inline void flux_calc(double g, const double *q, double *fx, double *fy) {
    double oorho = 1.0 / q[0];
    double u = q[1] * oorho, v = q[2] * oorho;
    double p = pressure(g, q);
    fx[0] = q[1]; fx[1] = q[1] * u + p; fx[2] = q[1] * v; fx[3] = u * (q[3] + p);
    fy[0] = q[2]; fy[1] = q[2] * u; fy[2] = q[2] * v + p; fy[3] = v * (q[3] + p);
}

/* isentropic_vortex/analytic_vortex.go:31-82 */
static void ivortex_state(const dfr2d_vortex *vx, double t, double x, double y, double *q) {
    double beta = vx->Beta, x0 = vx->X0, y0 = vx->Y0, gamma = vx->Gamma, ufs = vx->Ufs;
    double oo2pi = 0.5 * (1.0 / M_PI);
    double gm1 = gamma - 1.0, oogm1 = 1.0 / gm1;
    double fac = 16.0 * gamma * (M_PI * M_PI);
    double beta2 = beta * beta;
    double u = ufs, v = 0.0;
    double xmut = x - u * t, ymvt = y - v * t;
    double r2 = (xmut - x0) * (xmut - x0) + (ymvt - y0) * (ymvt - y0);
    double ex1r = exp(1.0 - r2);
    double tv1 = 1.0 - (gm1 * beta2 * exp(2.0 * (1.0 - r2)) / fac);
    u = u - beta * ex1r * (ymvt - y0) * oo2pi;
    v = v + beta * ex1r * (xmut - x0) * oo2pi;
    double rho = pow(tv1, oogm1);
    double p = pow(rho, gamma);
    double qq = 0.5 * rho * (u * u + v * v);
    q[0] = rho; q[1] = rho * u; q[2] = rho * v; q[3] = p * (1.0 / (gamma - 1.0)) + qq;
}

/* bcs.go:70-133 RiemannBC; q is the interior state on entry, the boundary state on exit */
static void riemann_bc(const dfr2d_freestream *fs, double *q, const double *qinf, double nx, double ny) {
    double gamma = fs->Gamma, p_inf = fs->Pinf, c_inf = fs->Cinf, minf = fs->Minf;
    double rho_int = q[0], u_int = q[1] / q[0], v_int = q[2] / q[0];
    double p_int = pressure(gamma, q), c_int = sound_speed(gamma, q);
    double gm1 = gamma - 1.0, oogm1 = 1.0 / gm1;
    double rho_inf = qinf[0], u_inf = qinf[1] / qinf[0], v_inf = qinf[2] / qinf[0];
    double tx = -ny, ty = nx;
    double vn_int = nx * u_int + ny * v_int;
    if (minf <= 1.0) {
        double vn_inf = nx * u_inf + ny * v_inf;
        double rinf = vn_inf - 2.0 * c_inf * oogm1;
        double rint = vn_int + 2.0 * c_int * oogm1;
        double vnorm = 0.5 * (rint + rinf);
        double c = 0.25 * gm1 * (rint - rinf);
        double vtang = 0.0, beta = 0.0; /* a NaN normal velocity matches neither case */
        if (vn_int < 0) {
            vtang = tx * u_inf + ty * v_inf;
            beta = p_inf / pow(rho_inf, gamma);
        } else if (vn_int >= 0) {
            vtang = tx * u_int + ty * v_int;
            beta = p_int / pow(rho_int, gamma);
        }
        double u = vnorm * nx + vtang * tx, v = vnorm * ny + vtang * ty;
        double rho = pow(c * c / (gamma * beta), oogm1);
        double p = beta * pow(rho, gamma);
        q[0] = rho; q[1] = rho * u; q[2] = rho * v; q[3] = p * oogm1 + 0.5 * rho * (u * u + v * v);
    } else if (vn_int < 0) {
        q[0] = qinf[0]; q[1] = qinf[1]; q[2] = qinf[2]; q[3] = qinf[3];
    }
}

/* fluxes.go:284-413 */
static void roe_flux(double gamma, const double *ql, const double *qr, double nx, double ny, double *f) {
    double gm1 = gamma - 1.0;
    double rho_ulr = ql[1] * nx + ql[2] * ny, rho_vlr = ql[1] * (-ny) + ql[2] * nx;
    double rho_urr = qr[1] * nx + qr[2] * ny, rho_vrr = qr[1] * (-ny) + qr[2] * nx;
    double rho_l = ql[0], u_l = rho_ulr / ql[0], v_l = rho_vlr / ql[0];
    double rho_r = qr[0], u_r = rho_urr / qr[0], v_r = rho_vrr / qr[0];
    double p_l = pressure(gamma, ql), p_r = pressure(gamma, qr);
    double h_l = (ql[3] + p_l) / rho_l, h_r = (qr[3] + p_r) / rho_r;
    double rho_ls = sqrt(rho_l), rho_rs = sqrt(rho_r);
    double rho_lsrs = rho_ls + rho_rs;
    double rho = rho_ls * rho_rs;
    double u = (rho_ls * u_l + rho_rs * u_r) / rho_lsrs;
    double v = (rho_ls * v_l + rho_rs * v_r) / rho_lsrs;
    double h = (rho_ls * h_l + rho_rs * h_r) / rho_lsrs;
    double c2 = gm1 * (h - 0.5 * (u * u + v * v));
    double c = sqrt(c2);
    double dw1 = -0.5 * (rho * (u_r - u_l)) / c + 0.5 * (p_r - p_l) / c2;
    double dw2 = (rho_r - rho_l) - (p_r - p_l) / c2;
    double dw3 = rho * (v_r - v_l);
    double dw4 = 0.5 * (rho * (u_r - u_l)) / c + 0.5 * (p_r - p_l) / c2;
    dw1 = fabs(u - c) * dw1;
    dw2 = fabs(u) * dw2;
    dw3 = fabs(u) * dw3;
    dw4 = fabs(u + c) * dw4;
    double f0 = 0.5 * (rho_ulr + rho_urr);
    double f1 = 0.5 * (rho_ulr * u_l + rho_urr * u_r + p_l + p_r);
    double f2 = 0.5 * (rho_vlr * u_l + rho_vrr * u_r);
    double f3 = 0.5 * ((p_l + ql[3]) * u_l + (p_r + qr[3]) * u_r);
    f0 = f0 - 0.5 * (dw1 + dw2 + dw4);
    f1 = f1 - 0.5 * (dw1 * (u - c) + dw2 * u + dw4 * (u + c));
    f2 = f2 - 0.5 * (dw1 * v + dw2 * v + dw3 + dw4 * v);
    f3 = f3 - 0.5 * (dw1 * (h - u * c) + 0.5 * dw2 * (u * u + v * v) + dw3 * v + dw4 * (h + u * c));
    f[0] = f0; f[1] = nx * f1 - ny * f2; f[2] = ny * f1 + nx * f2; f[3] = f3;
}

/* fluxes.go:161-190 */
static void lax_flux(double gamma, const double *ql, const double *qr, double nx, double ny, double *f) {
    double rho_l = ql[0], rho_r = qr[0];
    double u_l = ql[1] / rho_l, v_l = ql[2] / rho_l, u_r = qr[1] / rho_r, v_r = qr[2] / rho_r;
    double p_l = pressure(gamma, ql), p_r = pressure(gamma, qr);
    double c_l = sound_speed(gamma, ql), c_r = sound_speed(gamma, qr);
    double max_v = fmax(sqrt(u_l * u_l + v_l * v_l) + c_l, sqrt(u_r * u_r + v_r * v_r) + c_r);
    f[0] = 0.5 * (nx * (ql[1] + qr[1]) + ny * (ql[2] + qr[2]));
    f[1] = 0.5 * (nx * (ql[1] * u_l + qr[1] * u_r + p_l + p_r) + ny * (ql[1] * v_l + qr[1] * v_r));
    f[2] = 0.5 * (nx * (ql[2] * u_l + qr[2] * u_r) + ny * (ql[2] * v_l + qr[2] * v_r + p_l + p_r));
    f[3] = 0.5 * (nx * ((p_l + ql[3]) * u_l + (p_r + qr[3]) * u_r) + ny * ((p_l + ql[3]) * v_l + (p_r + qr[3]) * v_r));
    for (int n = 0; n < 4; n++) f[n] = f[n] + 0.5 * max_v * (ql[n] - qr[n]);
}

/* fluxes.go:135-159 */
static void avg_flux(double gamma, const double *ql, const double *qr, double nx, double ny, double *f) {
    double fxl[4], fyl[4], fxr[4], fyr[4];
    flux_calc(gamma, ql, fxl, fyl);
    flux_calc(gamma, qr, fxr, fyr);
    for (int n = 0; n < 4; n++) f[n] = nx * (0.5 * (fxl[n] + fxr[n])) + ny * (0.5 * (fyl[n] + fyr[n]));
}

/* fluxes.go:415-503, reproduced as written */
static void roe_er_flux(double gamma, const double *ql, const double *qr, double nx, double ny, double *f) {
    double gm1 = gamma - 1.0;
    double rho_l = ql[0], rho_r = qr[0];
    double oorho_l = 1.0 / rho_l, oorho_r = 1.0 / rho_r;
    double rho_ls = sqrt(rho_l), rho_rs = sqrt(rho_r);
    double u_l = ql[1] * oorho_l, v_l = ql[2] * oorho_l, u_r = qr[1] * oorho_r, v_r = qr[2] * oorho_r;
    double e_l = ql[3], e_r = qr[3];
    double uu_l = nx * u_l + ny * v_l, uu_r = nx * u_r + ny * v_r;
    double p_l = pressure(gamma, ql), p_r = pressure(gamma, qr);
    double hh_l = e_l + p_l, hh_r = e_r + p_r;
    double oors = 1.0 / (rho_ls + rho_rs);
    double u = (rho_ls * u_l + rho_rs * u_r) * oors, v = (rho_ls * v_l + rho_rs * v_r) * oors, h = (hh_l + hh_r) * oors;
    double rho = rho_ls * rho_rs;
    double hh = h * rho;
    double uu = nx * u + ny * v;
    double c2 = gm1 * (h - 0.5 * (u * u + v * v));
    double c = sqrt(c2), ooc = 1.0 / c;
    double uabs = fabs(uu);
    f[0] = 0.5 * (uu_l * rho_l + uu_r * rho_r);
    f[1] = 0.5 * (uu_l * rho_l * u_l + p_l * nx + uu_r * rho_r * u_r + p_r * nx);
    f[2] = 0.5 * (uu_l * rho_l * v_l + p_l * ny + uu_r * rho_r * v_r + p_r * ny);
    f[3] = 0.5 * (uu_l * hh_l + uu_r * hh_r);
    double uef = 0.05 * c;
    double du = u_r - u_l, dv = v_r - v_l;
    double delta_v2 = du * du + dv * dv;
    double oovmag = 1.0 / sqrt(u * u + v * v);
    double n1x, n1y;
    if (delta_v2 < 0.01 * c2) { n1x = nx; n1y = ny; } else { n1x = oovmag * du; n1y = oovmag * dv; }
    double n2x = n1y * (nx * n1y - n1x * ny), n2y = -n1x * (nx * n1y - n1x * ny);
    double alp1 = nx * n1x + ny * n1y, alp2 = nx * n2x + ny * n2y;
    double u1x = n1x * u, u1y = n1y * v, u2x = n2x * u, u2y = n2y * v;
    double urot = sqrt(alp1 * alp1 * (u1x * u1x + u1y * u1y)) + sqrt(alp2 * alp2 * (u2x * u2x + u2y * u2y));
    double sigma = fmax(uabs, fmin(uef, urot));
    double uabs_prime = uabs - 0.25 * fmax(0.0, uu_r - uu_l) * (fabs(uu + c) - fabs(uu - c));
    double d_u = uu_r - uu_l, d_p = p_r - p_l, d_rho = rho_r - rho_l;
    double d_rho_u = rho_r * u_r - rho_l * u_l, d_rho_v = rho_r * v_r - rho_l * v_l, d_e = e_r - e_l;
    double d_pu = rho * d_u * fmax(0.0, c - uabs_prime);
    double swt = fabs(uu) * fmin(uabs_prime, c);
    double d_pp = swt * d_p * ooc, d_uu = swt * d_u * ooc;
    f[0] = f[0] - 0.5 * (sigma * d_rho + (d_pu + d_pp) * 0 + d_uu * rho);
    f[1] = f[1] - 0.5 * (sigma * d_rho_u + (d_pu + d_pp) * nx + d_uu * rho * u);
    f[2] = f[2] - 0.5 * (sigma * d_rho_v + (d_pu + d_pp) * ny + d_uu * rho * v);
    f[3] = f[3] - 0.5 * (sigma * d_e + (d_pu + d_pp) * ny + d_uu * hh);
}

ora *ora_create(const dfr2d_problem *p) {
    if (p->dissipation) return NULL;
    ora *o = (ora *)calloc(1, sizeof(ora));
    o->p = *p;
    int N = p->N;
    o->N = N; o->NpInt = (N + 1) * (N + 2) / 2; o->NpEdge = N + 2; o->NpFlux = (N + 2) * (N + 4);
    long K = p->K, NE = p->NE;
    o->K = K; o->NE = NE;
    int ni = o->NpInt, ne = o->NpEdge, nf = o->NpFlux;
    o->FEI = dcopy(p->FluxEdgeInterp, 3L * ne * ni);
    o->DivInt = dcopy(p->DivInt, (long)ni * nf);
    o->Jdet = dcopy(p->Jdet, K);
    o->Jinv = dcopy(p->Jinv, 4 * K);
    o->IInII = dcopy(p->IInII, 3 * K);
    o->hK = dcopy(NULL, K);
    for (long k = 0; k < K; k++) o->hK[k] = p->EdgeLenMax[k] / (double)((N + 1) * (N + 1));
    o->kL = icopy(p->edge_kL, NE); o->kR = icopy(p->edge_kR, NE);
    o->numL = icopy(p->edge_numL, NE); o->numR = icopy(p->edge_numR, NE);
    o->nconn = icopy(p->edge_nconn, NE); o->bc = icopy(p->edge_bc, NE);
    o->etoe = icopy(p->EtoEdge, 3 * K);
    o->nxL = dcopy(NULL, NE); o->nyL = dcopy(NULL, NE);
    for (long e = 0; e < NE; e++) {
        o->nxL[e] = p->FaceNormX[o->kL[e] + K * o->numL[e]];
        o->nyL[e] = p->FaceNormY[o->kL[e] + K * o->numL[e]];
    }
    o->bp_of_edge = (int *)malloc(sizeof(int) * (size_t)(NE > 0 ? NE : 1));
    for (long e = 0; e < NE; e++) o->bp_of_edge[e] = -1;
    for (long b = 0; b < p->NBP; b++) o->bp_of_edge[p->bp_edge[b]] = (int)b;
    o->bpx = dcopy(p->bp_x, p->NBP * ne); o->bpy = dcopy(p->bp_y, p->NBP * ne);
    long reg = 4L * ni * K;
    for (int r = 0; r < 5; r++) o->Q[r] = (double *)calloc((size_t)reg, sizeof(double));
    o->Residual = (double *)calloc((size_t)reg, sizeof(double));
    o->RHSQ = (double *)calloc((size_t)reg, sizeof(double));
    o->Q_Face = (double *)calloc((size_t)(12L * ne * K), sizeof(double));
    o->F_RT_DOF = (double *)calloc((size_t)(4L * nf * K), sizeof(double));
    o->DT = (double *)calloc((size_t)K, sizeof(double));
    o->EdgeFlux = (double *)calloc((size_t)(4L * NE * ne), sizeof(double));
    o->EdgeQ = (double *)calloc((size_t)(4L * NE * ne), sizeof(double));
    o->Agg = (double *)calloc((size_t)(NE > 0 ? NE : 1), sizeof(double));
    return o;
}

void ora_destroy(ora *o) {
    if (!o) return;
    free(o->FEI); free(o->DivInt); free(o->Jdet); free(o->Jinv); free(o->IInII); free(o->hK);
    free(o->kL); free(o->kR); free(o->numL); free(o->numR); free(o->nconn); free(o->bc); free(o->etoe);
    free(o->nxL); free(o->nyL); free(o->bp_of_edge); free(o->bpx); free(o->bpy);
    for (int r = 0; r < 5; r++) free(o->Q[r]);
    free(o->Residual); free(o->RHSQ); free(o->Q_Face); free(o->F_RT_DOF); free(o->DT);
    free(o->EdgeFlux); free(o->EdgeQ); free(o->Agg);
    free(o);
}

int ora_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void ora_set_state(ora *o, const double *Q) { memcpy(o->Q[0], Q, sizeof(double) * 4 * (size_t)o->NpInt * (size_t)o->K); }
void ora_get_state(ora *o, double *Q) { memcpy(Q, o->Q[0], sizeof(double) * 4 * (size_t)o->NpInt * (size_t)o->K); }
void ora_residual(ora *o, double *maxR) {
    long n1 = (long)o->NpInt * o->K;
    for (int n = 0; n < 4; n++) {
        double m = -DBL_MAX;
        for (long i = 0; i < n1; i++) m = fmax(m, o->Residual[n * n1 + i]);
        maxR[n] = m;
    }
}

/* Y[m x K] = A[m x n] . X[n x K] on the column range [k0,k1) -- the role dgemm plays in the
 * reference (edges.go:485-491, euler.go:665-699) */
static void op_apply(const double *A, int m, int n, const double *X, double *Y, long K, long k0, long k1) {
    for (int i = 0; i < m; i++) {
        double *y = Y + (long)i * K;
        const double *a = A + (long)i * n;
        for (long k = k0; k < k1; k++) y[k] = a[0] * X[k];
        for (int j = 1; j < n; j++) {
            const double *x = X + (long)j * K;
            double aij = a[j];
            for (long k = k0; k < k1; k++) y[k] += aij * x[k];
        }
    }
}

#define BLK 256

/* one RK stage: StepWorker euler.go:420-653 (inviscid branches) */
static void stage(ora *o, int rk) {
    const dfr2d_problem *p = &o->p;
    const long K = o->K, NE = o->NE;
    const int ni = o->NpInt, ne = o->NpEdge, nf = o->NpFlux;
    const double gamma = p->FSFar.Gamma;
    double *qqq = o->Q[rk];
    const long nblk = (K + BLK - 1) / BLK;

    /* InterpolateSolutionToEdges, edges.go:485-491 */
#pragma omp parallel for schedule(static)
    for (long b = 0; b < nblk; b++) {
        long k0 = b * BLK, k1 = k0 + BLK < K ? k0 + BLK : K;
        for (int n = 0; n < 4; n++)
            op_apply(o->FEI, 3 * ne, ni, qqq + (long)n * ni * K, o->Q_Face + (long)n * 3 * ne * K, K, k0, k1);
    }

    /* CalculateEdgeEulerFlux, edges.go:325-452 (+ bcs.go) and StoreEdgeAggregates, edges.go:246-289 */
#pragma omp parallel for schedule(static)
    for (long e = 0; e < NE; e++) {
        int kl = o->kL[e], kr = o->kR[e];
        double nx = o->nxL[e], ny = o->nyL[e];
        int bc = o->bc[e];
        double amax = -DBL_MAX;
        double oohk = 1.0 / o->hK[kl];
        for (int i = 0; i < ne; i++) {
            long rowL = (long)o->numL[e] * ne + i;
            double ql[4], qr[4], f[4] = {0, 0, 0, 0};
            for (int n = 0; n < 4; n++) {
                ql[n] = o->Q_Face[((long)n * 3 * ne + rowL) * K + kl];
                o->EdgeQ[((long)n * NE + e) * ne + i] = ql[n]; /* before any BC overwrite, edges.go:344-350 */
            }
            if (o->nconn[e] == 2) {
                long rowR = (long)o->numR[e] * ne + (ne - 1 - i);
                for (int n = 0; n < 4; n++) qr[n] = o->Q_Face[((long)n * 3 * ne + rowR) * K + kr];
                switch (p->flux_type) {
                case DFR2D_FLUX_Average: avg_flux(gamma, ql, qr, nx, ny, f); break;
                case DFR2D_FLUX_LaxFriedrichs: lax_flux(gamma, ql, qr, nx, ny, f); break;
                case DFR2D_FLUX_Roe: roe_flux(gamma, ql, qr, nx, ny, f); break;
                default: roe_er_flux(gamma, ql, qr, nx, ny, f); break;
                }
            } else if (bc == DFR2D_BC_Periodic || bc == DFR2D_BC_PeriodicReversed) {
                /* no flux is computed for an unpaired periodic edge */
            } else if (bc == DFR2D_BC_Wall || bc == DFR2D_BC_Cyl) {
                double pw = pressure(gamma, ql); /* bcs.go:11-23 */
                f[1] = nx * pw; f[2] = ny * pw;
            } else {
                if (bc == DFR2D_BC_Far || bc == DFR2D_BC_In || bc == DFR2D_BC_Out) {
                    const dfr2d_freestream *fs = bc == DFR2D_BC_Far ? &p->FSFar : (bc == DFR2D_BC_In ? &p->FSIn : &p->FSOut);
                    riemann_bc(fs, ql, fs->Qinf, nx, ny);
                } else if (bc == DFR2D_BC_IVortex) {
                    double qex[4];
                    long b = o->bp_of_edge[e];
                    ivortex_state(&p->vortex, o->Time, o->bpx[b * ne + i], o->bpy[b * ne + i], qex);
                    riemann_bc(&p->FSFar, ql, qex, nx, ny);
                }
                if (bc != DFR2D_BC_None)
                    for (int n = 0; n < 4; n++) o->Q_Face[((long)n * 3 * ne + rowL) * K + kl] = ql[n]; /* bcs.go:47-50 */
                double fx[4], fy[4];
                flux_calc(gamma, ql, fx, fy);
                for (int n = 0; n < 4; n++) f[n] = nx * fx[n] + ny * fy[n];
            }
            for (int n = 0; n < 4; n++) o->EdgeFlux[((long)n * NE + e) * ne + i] = f[n];
            /* aggregate on the post-BC L state */
            double c = sound_speed(gamma, ql);
            double oor = 1.0 / ql[0];
            double u = ql[1] * oor, v = ql[2] * oor;
            double w = oohk * (sqrt(u * u + v * v) + c);
            if (w > amax) amax = w;
        }
        o->Agg[e] = amax;
    }

    /* CalcElementMaxWaveSpeed edges.go:291-323, InitializeDT euler.go:655-663 */
    double gmax = -DBL_MAX;
#pragma omp parallel for schedule(static) reduction(max : gmax)
    for (long k = 0; k < K; k++) {
        double dt = rk == 0 ? -100.0 : o->DT[k];
        for (int e = 0; e < 3; e++) {
            double a = o->Agg[o->etoe[3 * k + e]];
            if (a > dt) dt = a;
            if (a > gmax) gmax = a;
        }
        o->DT[k] = dt;
    }
    /* calculateGlobalDT euler.go:945-971 / CalculateLocalDT :973-1002 */
    if (!p->local_time_stepping) {
        o->GlobalDT = p->CFL / fmax(0.0, gmax);
        if (o->Time + o->GlobalDT > p->FinalTime) o->GlobalDT = p->FinalTime - o->Time;
    }
    const double gdt = o->GlobalDT;
    const int local = p->local_time_stepping;
    const double cfl = p->CFL;

    /* rkAdvance euler.go:502-565: SetRTFluxInternal :701-726, SetRTFluxOnEdges edges.go:454-483,
     * RHSInternalPoints euler.go:665-699, SSP54 combination */
    double *q0 = o->Q[0], *q1 = o->Q[1], *q2 = o->Q[2], *q3 = o->Q[3], *q4 = o->Q[4];
    const long reg1 = (long)ni * K;
#pragma omp parallel for schedule(static)
    for (long b = 0; b < nblk; b++) {
        long k0 = b * BLK, k1 = k0 + BLK < K ? k0 + BLK : K;
        for (long k = k0; k < k1; k++) o->DT[k] = local ? cfl / o->DT[k] : gdt;
        for (int i = 0; i < ni; i++)
            for (long k = k0; k < k1; k++) {
                double q[4], fx[4], fy[4];
                for (int n = 0; n < 4; n++) q[n] = qqq[n * reg1 + (long)i * K + k];
                flux_calc(gamma, q, fx, fy);
                double jd = o->Jdet[k];
                const double *ji = o->Jinv + 4 * k;
                for (int n = 0; n < 4; n++) {
                    double *F = o->F_RT_DOF + (long)n * nf * K;
                    F[(long)i * K + k] = jd * (ji[0] * fx[n] + ji[1] * fy[n]);
                    F[(long)(i + ni) * K + k] = jd * (ji[2] * fx[n] + ji[3] * fy[n]);
                }
            }
        for (int e3 = 0; e3 < 3; e3++)
            for (long k = k0; k < k1; k++) {
                long e = o->etoe[3 * k + e3];
                int owner = o->kL[e] == k;
                double iin = o->IInII[(long)e3 * K + k];
                for (int n = 0; n < 4; n++) {
                    const double *ef = o->EdgeFlux + ((long)n * NE + e) * ne;
                    double *F = o->F_RT_DOF + (long)n * nf * K + (long)(2 * ni + e3 * ne) * K + k;
                    for (int i = 0; i < ne; i++) F[(long)i * K] = (owner ? ef[i] : -ef[ne - 1 - i]) * iin;
                }
            }
        for (int n = 0; n < 4; n++) {
            double *rhs = o->RHSQ + n * reg1;
            op_apply(o->DivInt, ni, nf, o->F_RT_DOF + (long)n * nf * K, rhs, K, k0, k1);
            for (int i = 0; i < ni; i++)
                for (long k = k0; k < k1; k++) {
                    long x = n * reg1 + (long)i * K + k;
                    double r = o->RHSQ[x] * (-(1.0 / o->Jdet[k]));
                    o->RHSQ[x] = r;
                    double dt_rhs = o->DT[k] * r;
                    switch (rk) {
                    case 0: q1[x] = q0[x] + RK_A * dt_rhs; break;
                    case 1: q2[x] = RK1[0] * q0[x] + RK1[1] * q1[x] + RK1[2] * dt_rhs; break;
                    case 2: q3[x] = RK2[0] * q0[x] + RK2[1] * q2[x] + RK2[2] * dt_rhs; break;
                    case 3:
                        o->Residual[x] = r;
                        q4[x] = RK3[0] * q0[x] + RK3[1] * q3[x] + RK3[2] * dt_rhs;
                        break;
                    default: {
                        double dt_r3 = o->DT[k] * o->Residual[x];
                        double rr = -q0[x] + RK4[0] * q2[x] + RK4[1] * q3[x] + RK4[2] * q4[x] + RK4[3] * dt_r3 + RK4[4] * dt_rhs;
                        o->Residual[x] = rr;
                        q0[x] = q0[x] + rr;
                    }
                    }
                }
        }
    }
}

int ora_step(ora *o, int nsteps, dfr2d_step_info *info) {
    int finished = 0;
    for (int s = 0; s < nsteps; s++) {
        for (int rk = 0; rk < 5; rk++) stage(o, rk);
        o->Time += o->GlobalDT;
        o->StepCount++;
        finished = o->Time >= o->p.FinalTime || o->StepCount >= o->p.max_iterations;
        if (finished) break;
    }
    if (info) {
        info->time = o->Time; info->dt = o->GlobalDT; info->steps = o->StepCount;
        info->finished = finished; info->nan_found = 0;
    }
    return 0;
}

/* RHSQ of stage rk on the current register rk, without advancing (test hook) */
int ora_rhs(ora *o, int rk, double *out) {
    size_t reg = sizeof(double) * 4 * (size_t)o->NpInt * (size_t)o->K;
    double *save[5], *sres = (double *)malloc(reg), *sdt = (double *)malloc(sizeof(double) * (size_t)o->K);
    for (int r = 0; r < 5; r++) { save[r] = (double *)malloc(reg); memcpy(save[r], o->Q[r], reg); }
    memcpy(sres, o->Residual, reg);
    memcpy(sdt, o->DT, sizeof(double) * (size_t)o->K);
    double gdt = o->GlobalDT;
    stage(o, rk);
    memcpy(out, o->RHSQ, reg);
    for (int r = 0; r < 5; r++) { memcpy(o->Q[r], save[r], reg); free(save[r]); }
    memcpy(o->Residual, sres, reg); memcpy(o->DT, sdt, sizeof(double) * (size_t)o->K);
    o->GlobalDT = gdt;
    free(sres); free(sdt);
    return 0;
}
