/*
 * ORACLE -- test infrastructure, not product code.
 *
 * Plain C (OpenMP) restatement of the reference's per-stage 2D Euler DFR right-hand side and
 * SSP-RK(5,4) step -- the inviscid path and the PerssonC0 sensor / artificial-dissipation path --
 * organised the way the reference runs it on the CPU: the same
 * phases with a barrier between them (StepWorker, model_problems/Euler2D/euler.go:420-653),
 * materialised Q_Face / EdgeStore / F_RT_DOF / RHSQ arrays, operator applications as
 * [m x n]·[n x K] products, threads standing in for the goroutine partitions.  It consumes the
 * same flat dfr2d_problem as the device library (include/dfr2d.h).
 *
 * Uses: (1) second, independently written checker next to oracle/euler2d_oracle.py (the two
 * are compared in tests/test_c_oracle.py); (2) the CPU baseline bench.py times on all host
 * cores ("port": the Go toolchain is absent, this is NOT the Go solver).
 * Parity status: pinned only through the numpy oracle and the reference KATs it passes
 * (tests/test_oracle_kats.py); see DESIGN.md section 5.  The dissipation phases follow the Go
 * data flow literally: materialised DXMetric/DYMetric, DOFX/DOFY, GradX/GradY, Epsilon, DissX/DissY,
 * the viscous edge store and the DissDOF/DissDiv scratch (dissipation.go, edges.go:151-244).
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
    /* PerssonC0 dissipation (dissipation.go:90-217, dfr_shock_capturing.go:71-107) */
    int diss;
    long NV;
    double *Div, *V, *Vinv, *Mass, *D, *P, *mf, *Bary, *edge_len, *nxk, *nyk;
    int *EToV;
    double *DXMetric, *DYMetric;            /* [NpFlux x K]  dfr_startup.go:213-254 */
    double *Se, *SigmaScalar, *EpsilonScalar, *SigmaVertex, *EpsVertex, *Epsilon;
    double *DissX, *DissY, *EdgeVisc, *AggV, *DTVisc;
    double *LS0, *LS1, *LS2, *DOFX, *DOFY;  /* LScratch [NpInt x K] x3, DissDOF/DissDOF2 [NpFlux x K] */
    double Kappa, Eps0;
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
    o->diss = p->dissipation != 0;
    if (o->diss) {
        long NV = p->NV;
        o->NV = NV;
        o->Div = dcopy(p->Div, (long)nf * nf);
        o->V = dcopy(p->V, (long)ni * ni); o->Vinv = dcopy(p->Vinv, (long)ni * ni);
        o->Mass = dcopy(p->MassMatrix, (long)ni * ni); o->D = dcopy(p->D, (long)ni * ni); o->P = dcopy(p->P, (long)ni * ni);
        o->mf = dcopy(p->ModeFilter, ni);
        o->Bary = dcopy(p->Bary, 3L * nf);
        o->edge_len = dcopy(p->edge_len, NE);
        o->nxk = dcopy(p->FaceNormX, 3 * K); o->nyk = dcopy(p->FaceNormY, 3 * K);
        o->EToV = icopy(p->EToV, 3 * K);
        /* NewScalarDissipation: Kappa 5 unless given, Eps0 fixed before the override (dissipation.go:140-147) */
        o->Kappa = 5.0;
        o->Eps0 = o->Kappa / 1.5;
        if (p->Kappa != 0.0) o->Kappa = p->Kappa;
        /* CalculateRTBasedDerivativeMetrics, DG2D/dfr_startup.go:213-254 */
        o->DXMetric = dcopy(NULL, (long)nf * K); o->DYMetric = dcopy(NULL, (long)nf * K);
        for (long k = 0; k < K; k++) {
            const double *ji = o->Jinv + 4 * k;
            double oojd = 1.0 / o->Jdet[k];
            for (int i = 0; i < ni; i++) {
                o->DXMetric[(long)i * K + k] = ji[0]; o->DYMetric[(long)i * K + k] = ji[1];
                o->DXMetric[(long)(i + ni) * K + k] = ji[2]; o->DYMetric[(long)(i + ni) * K + k] = ji[3];
            }
            for (int fn = 0; fn < 3; fn++) {
                double iin = o->IInII[(long)fn * K + k];
                for (int i = 0; i < ne; i++) {
                    long r = 2 * ni + fn * ne + i;
                    o->DXMetric[r * K + k] = oojd * o->nxk[(long)fn * K + k] * iin;
                    o->DYMetric[r * K + k] = oojd * o->nyk[(long)fn * K + k] * iin;
                }
            }
        }
        o->Se = (double *)calloc((size_t)K, sizeof(double));
        o->SigmaScalar = (double *)calloc((size_t)K, sizeof(double));
        o->EpsilonScalar = (double *)calloc((size_t)K, sizeof(double));
        o->SigmaVertex = (double *)calloc((size_t)(NV > 0 ? NV : 1), sizeof(double));
        o->EpsVertex = (double *)calloc((size_t)(NV > 0 ? NV : 1), sizeof(double));
        o->Epsilon = (double *)calloc((size_t)((long)nf * K), sizeof(double));
        o->DissX = (double *)calloc((size_t)(4L * nf * K), sizeof(double));
        o->DissY = (double *)calloc((size_t)(4L * nf * K), sizeof(double));
        o->EdgeVisc = (double *)calloc((size_t)(4L * NE * ne), sizeof(double));
        o->AggV = (double *)calloc((size_t)(NE > 0 ? NE : 1), sizeof(double));
        o->DTVisc = (double *)calloc((size_t)K, sizeof(double));
        o->LS0 = (double *)calloc((size_t)((long)ni * K), sizeof(double));
        o->LS1 = (double *)calloc((size_t)((long)ni * K), sizeof(double));
        o->LS2 = (double *)calloc((size_t)((long)ni * K), sizeof(double));
        o->DOFX = (double *)calloc((size_t)((long)nf * K), sizeof(double));
        o->DOFY = (double *)calloc((size_t)((long)nf * K), sizeof(double));
    }
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
    if (o->diss) {
        free(o->Div); free(o->V); free(o->Vinv); free(o->Mass); free(o->D); free(o->P); free(o->mf); free(o->Bary);
        free(o->edge_len); free(o->nxk); free(o->nyk); free(o->EToV); free(o->DXMetric); free(o->DYMetric);
        free(o->Se); free(o->SigmaScalar); free(o->EpsilonScalar); free(o->SigmaVertex); free(o->EpsVertex);
        free(o->Epsilon); free(o->DissX); free(o->DissY); free(o->EdgeVisc); free(o->AggV); free(o->DTVisc);
        free(o->LS0); free(o->LS1); free(o->LS2); free(o->DOFX); free(o->DOFY);
    }
    free(o);
}

int ora_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Explicit thread count for the CPU arm of bench.py: launchers such as torchrun export OMP_NUM_THREADS=1, which would
 * silently turn the "all host cores" baseline into a single-thread one.  n <= 0 leaves the OpenMP default. */
int ora_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) {
        omp_set_dynamic(0);
        omp_set_num_threads(n);
    }
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

void ora_set_state(ora *o, const double *Q) { memcpy(o->Q[0], Q, sizeof(double) * 4 * (size_t)o->NpInt * (size_t)o->K); }
void ora_set_register(ora *o, int reg, const double *Q) {
    if (reg >= 0 && reg < 5) memcpy(o->Q[reg], Q, sizeof(double) * 4 * (size_t)o->NpInt * (size_t)o->K);
}
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

/* ---- PerssonC0 dissipation phases --------------------------------------------------------------------------- */

/* ModeAliasShockFinder.UpdateSeMoment, DG2D/dfr_shock_capturing.go:142-166 */
static void update_se_moment(ora *o, const double *rho) {
    const long K = o->K;
    const int ni = o->NpInt;
#pragma omp parallel for schedule(static)
    for (long b = 0; b < (K + BLK - 1) / BLK; b++) {
        long k0 = b * BLK, k1 = k0 + BLK < K ? k0 + BLK : K;
        op_apply(o->P, ni, ni, rho, o->LS0, K, k0, k1);
        op_apply(o->Mass, ni, ni, rho, o->LS1, K, k0, k1);
        op_apply(o->D, ni, ni, rho, o->LS2, K, k0, k1);
    }
#pragma omp parallel for schedule(static)
    for (long k = 0; k < K; k++) {
        double num = 0.0, den = 0.0;
        for (int i = 0; i < ni; i++) {
            long ind = k + (long)i * K;
            double di = o->LS2[ind];
            num += di * o->LS0[ind];
            den += rho[ind] * o->LS1[ind];
        }
        o->Se[k] = log10(num / den);
    }
}

/* UpdateShockFinderSigma dissipation.go:491-518 (sigma lives outside the loop, as there) and
 * CalculateElementViscosity :397-412 */
static void update_sigma_and_viscosity(ora *o) {
    const long K = o->K;
    const double kappa = o->Kappa;
    const double S0 = 4.0 / pow((double)(o->N + 1), 4.0);
    const double left = S0 - kappa, right = S0 + kappa;
    const double ookappa = 0.5 / kappa;
    double sigma = 0.0;
    for (long k = 0; k < K; k++) {
        double se = o->Se[k];
        if (se < left) sigma = 0.0;
        else if (se >= left && se <= right) sigma = 0.5 * (1.0 + sin(M_PI * ookappa * (se - S0)));
        else if (se > right) sigma = 1.0;
        o->SigmaScalar[k] = sigma;
    }
    for (long k = 0; k < K; k++) o->EpsilonScalar[k] = o->Eps0 * o->hK[k] * o->SigmaScalar[k];
}

/* MergeElementScalarToVertices with Max, euler.go:1048-1065: the sorted (vertex, element) sweep resets a vertex at
 * its first element and maxes over the rest -- i.e. the max over the incident elements; untouched vertices keep
 * their value */
static void merge_to_vertices(ora *o, const double *elem, double *vert, char *seen) {
    memset(seen, 0, (size_t)(o->NV > 0 ? o->NV : 1));
    for (long k = 0; k < o->K; k++)
        for (int v = 0; v < 3; v++) {
            int vv = o->EToV[3 * k + v];
            if (!seen[vv]) { vert[vv] = elem[k]; seen[vv] = 1; }
            else if (elem[k] > vert[vv]) vert[vv] = elem[k];
        }
}

/* limitAndFilterSolution dissipation.go:606-622 on the columns [k0,k1) of one variable */
static void limit_and_filter(ora *o, double *U, double *scratch, long k0, long k1) {
    const long K = o->K;
    const int ni = o->NpInt;
    op_apply(o->Vinv, ni, ni, U, scratch, K, k0, k1);
    for (int i = 1; i < ni; i++)
        for (long k = k0; k < k1; k++) {
            double alpha = sin(0.5 * M_PI * o->SigmaScalar[k]);
            scratch[k + K * i] *= o->mf[i] * (1.0 - alpha);
        }
    op_apply(o->V, ni, ni, scratch, U, K, k0, k1);
}

/* EdgeStore.GetEdgeValues edges.go:93-113: slice of edge `e` as element k sees it and the sign (+1 owner, -1 not) */
static inline int edge_sign(const ora *o, long e, long k) { return o->kL[e] == k ? 1 : -1; }

/* CalculateEpsilonGradient dissipation.go:244-272 (C0) over GetSolutionGradientUsingRTElement euler.go:864-918 */
static void calculate_epsilon_gradient(ora *o, const double *qqq) {
    const long K = o->K, NE = o->NE;
    const int ni = o->NpInt, ne = o->NpEdge, nf = o->NpFlux;
    for (int n = 0; n < 4; n++) {
        const double *Q = qqq + (long)n * ni * K;
#pragma omp parallel for schedule(static)
        for (long k = 0; k < K; k++)
            for (int i = 0; i < ni; i++) {
                long ind = k + (long)i * K, ind2 = k + (long)(i + ni) * K;
                double Un = Q[ind];
                o->DOFX[ind] = o->DXMetric[ind] * Un; o->DOFY[ind] = o->DYMetric[ind] * Un;
                o->DOFX[ind2] = o->DXMetric[ind2] * Un; o->DOFY[ind2] = o->DYMetric[ind2] * Un;
            }
#pragma omp parallel for schedule(static)
        for (long k = 0; k < K; k++)
            for (int edgeNum = 0; edgeNum < 3; edgeNum++) {
                long e = o->etoe[3 * k + edgeNum];
                const double *edgeVals = o->EdgeQ + ((long)n * NE + e) * ne;
                int sign = edge_sign(o, e, k);
                int shift = edgeNum * ne;
                for (int i = 0; i < ne; i++) {
                    int ii = sign < 0 ? ne - 1 - i : i;
                    long ind = k + (long)(2 * ni + i + shift) * K;
                    double Un = edgeVals[ii];
                    o->DOFX[ind] = o->DXMetric[ind] * Un;
                    o->DOFY[ind] = o->DYMetric[ind] * Un;
                }
            }
        double *GX = o->DissX + (long)n * nf * K, *GY = o->DissY + (long)n * nf * K;
#pragma omp parallel for schedule(static)
        for (long b = 0; b < (K + BLK - 1) / BLK; b++) {
            long k0 = b * BLK, k1 = k0 + BLK < K ? k0 + BLK : K;
            op_apply(o->Div, nf, nf, o->DOFX, GX, K, k0, k1);
            op_apply(o->Div, nf, nf, o->DOFY, GY, K, k0, k1);
        }
#pragma omp parallel for schedule(static)
        for (long x = 0; x < (long)nf * K; x++) { GX[x] *= o->Epsilon[x]; GY[x] *= o->Epsilon[x]; }
    }
}

/* StoreEdgeViscousFlux edges.go:151-244 and the viscous half of StoreEdgeAggregates edges.go:274-286 */
static void store_edge_viscous_flux(ora *o) {
    const long K = o->K, NE = o->NE;
    const int ni = o->NpInt, ne = o->NpEdge, nf = o->NpFlux;
    const double Omega = 1.0 * (double)(o->N * o->N);
#pragma omp parallel for schedule(static)
    for (long e = 0; e < NE; e++) {
        long kL = o->kL[e];
        long kR = o->nconn[e] == 2 ? o->kR[e] : 0;   /* ConnectedTris[1] is 0 for one-tri edges (unused there) */
        int shiftL = ne * o->numL[e], shiftR = ne * (o->nconn[e] == 2 ? o->numR[e] : 0);
        double nxL = o->nxL[e], nyL = o->nyL[e];
        double nxR = nxL, nyR = nyL;                 /* normalR := GetFaceNormal(kLGlobal, edgeNumberL), edges.go:175 */
        double ooedgeLength = 1.0 / o->edge_len[e];
        for (int n = 0; n < 4; n++) {
            const double *DX = o->DissX + (long)n * nf * K, *DY = o->DissY + (long)n * nf * K;
            double *vf = o->EdgeVisc + ((long)n * NE + e) * ne;
            if (o->nconn[e] == 1) {
                for (int i = 0; i < ne; i++) {
                    long indL = kL + (long)(2 * ni + shiftL + i) * K;
                    vf[i] = nxL * DX[indL] + nyL * DY[indL];
                }
            } else {
                /* both GetEdgeValues calls of edges.go:225-228 return the one stored (owner-side) slice */
                const double *edgeQL = o->EdgeQ + ((long)n * NE + e) * ne, *edgeQR = edgeQL;
                for (int i = 0; i < ne; i++) {
                    long indL = kL + (long)(2 * ni + shiftL + i) * K;
                    long indR = kR + (long)(2 * ni + shiftR + ne - 1 - i) * K;
                    double vFL = nxL * DX[indL] + nyL * DY[indL];
                    double vFR = nxR * DX[indR] + nyR * DY[indR];
                    vf[i] = 0.5 * (vFL + vFR);
                    double LambdaAvg = 0.5 * (o->Epsilon[indL] + o->Epsilon[indR]);
                    int ii = ne - 1 - i;
                    vf[i] -= (Omega * LambdaAvg * ooedgeLength) * (edgeQL[i] - edgeQR[ii]);
                }
            }
        }
        double oohKVisc = 1.0 / o->hK[kL];
        double m = -DBL_MAX;
        for (int i = 0; i < ne; i++) {
            double v = oohKVisc * oohKVisc * o->Epsilon[kL + (long)(2 * ni + shiftL + i) * K];
            m = v > m ? v : m;
        }
        o->AggV[e] = m;
    }
}

/* AddDissipation dissipation.go:274-346 for variable n on the columns [k0,k1): RHSQ += (1/Jdet) DivInt . DOF */
static void add_dissipation(ora *o, int n, double *rhs, long k0, long k1) {
    const long K = o->K, NE = o->NE;
    const int ni = o->NpInt, ne = o->NpEdge, nf = o->NpFlux;
    const double *DiX = o->DissX + (long)n * nf * K, *DiY = o->DissY + (long)n * nf * K;
    double *DOF = o->DOFX, *DIV = o->LS0;        /* DissDOF / DissDiv scratch (column ranges are disjoint) */
    for (long k = k0; k < k1; k++) {
        double jd = o->Jdet[k];
        const double *ji = o->Jinv + 4 * k;
        for (int i = 0; i < ni; i++) {
            long ind = k + K * i, ind2 = k + K * (i + ni);
            DOF[ind] = jd * (ji[0] * DiX[ind] + ji[1] * DiY[ind]);
            DOF[ind2] = jd * (ji[2] * DiX[ind] + ji[3] * DiY[ind]);
        }
        for (int edgeNum = 0; edgeNum < 3; edgeNum++) {
            double IInII = o->IInII[k + K * edgeNum];
            int shift = ne * edgeNum;
            long e = o->etoe[3 * k + edgeNum];
            const double *edgeFlux = o->EdgeVisc + ((long)n * NE + e) * ne;
            int sign = edge_sign(o, e, k);
            for (int i = 0; i < ne; i++) {
                int ii = sign == -1 ? ne - 1 - i : i;
                DOF[k + (long)(2 * ni + i + shift) * K] = edgeFlux[ii] * IInII * (double)sign;
            }
        }
    }
    op_apply(o->DivInt, ni, nf, DOF, DIV, K, k0, k1);
    for (long k = k0; k < k1; k++) {
        double oojd = 1.0 / o->Jdet[k];
        for (int i = 0; i < ni; i++) rhs[k + (long)i * K] += oojd * DIV[k + (long)i * K];
    }
}

/* one RK stage: StepWorker euler.go:420-653 */
static void stage(ora *o, int rk) {
    const dfr2d_problem *p = &o->p;
    const long K = o->K, NE = o->NE;
    const int ni = o->NpInt, ne = o->NpEdge, nf = o->NpFlux;
    const double gamma = p->FSFar.Gamma;
    double *qqq = o->Q[rk];
    const long nblk = (K + BLK - 1) / BLK;

    if (o->diss) {
        /* euler.go:574-616: sensor, element viscosity, element->vertex max, vertex->element mean,
         * InterpolateEpsilonSigma (dissipation.go:219-242), stage-2 limiter of the stage input */
        update_se_moment(o, qqq);
        update_sigma_and_viscosity(o);
        char *seen = (char *)malloc((size_t)(o->NV > 0 ? o->NV : 1));
        merge_to_vertices(o, o->SigmaScalar, o->SigmaVertex, seen);
        merge_to_vertices(o, o->EpsilonScalar, o->EpsVertex, seen);
        free(seen);
#pragma omp parallel for schedule(static)
        for (long k = 0; k < K; k++) {          /* MergeVertexScalarToElement(sum, div3), euler.go:1067-1086 */
            double acc = 0.0;
            for (int v = 0; v < 3; v++) acc = acc + o->SigmaVertex[o->EToV[v + 3 * k]];
            o->SigmaScalar[k] = acc / 3.0;
        }
#pragma omp parallel for schedule(static)
        for (long k = 0; k < K; k++) {
            double v3[3] = {o->EpsVertex[o->EToV[3 * k]], o->EpsVertex[o->EToV[3 * k + 1]], o->EpsVertex[o->EToV[3 * k + 2]]};
            for (int i = 0; i < nf; i++) {
                const double *b = o->Bary + 3 * i;
                double acc = b[0] * v3[0];
                acc += b[1] * v3[1];
                acc += b[2] * v3[2];
                o->Epsilon[k + (long)i * K] = acc;
            }
        }
        if (rk == 2) {
#pragma omp parallel for schedule(static)
            for (long b = 0; b < nblk; b++) {
                long k0 = b * BLK, k1 = k0 + BLK < K ? k0 + BLK : K;
                for (int n = 0; n < 4; n++) limit_and_filter(o, qqq + (long)n * ni * K, o->LS2, k0, k1);
            }
        }
    }

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

    if (o->diss) {
        calculate_epsilon_gradient(o, qqq);   /* euler.go:624-628 */
        store_edge_viscous_flux(o);           /* euler.go:629-635 */
    }

    /* CalcElementMaxWaveSpeed edges.go:291-323, InitializeDT euler.go:655-663 */
    double gmax = -DBL_MAX, gmaxv = -DBL_MAX;
    const int diss = o->diss;
#pragma omp parallel for schedule(static) reduction(max : gmax, gmaxv)
    for (long k = 0; k < K; k++) {
        double dt = rk == 0 ? -100.0 : o->DT[k];
        for (int e = 0; e < 3; e++) {
            double a = o->Agg[o->etoe[3 * k + e]];
            if (a > dt) dt = a;
            if (a > gmax) gmax = a;
            if (diss) {
                double av = o->AggV[o->etoe[3 * k + e]];
                if (av > gmaxv) gmaxv = av;
                o->DTVisc[k] = fmax(o->DTVisc[k], av);
            }
        }
        o->DT[k] = dt;
    }
    /* calculateGlobalDT euler.go:945-971 / CalculateLocalDT :973-1002 */
    const double NP12 = (double)((o->N + 1) * (o->N + 1));
    const double C_diff = 1.0 / NP12;
    if (!p->local_time_stepping) {
        o->GlobalDT = p->CFL / fmax(0.0, gmax);
        if (diss) o->GlobalDT = fmin(o->GlobalDT, C_diff / fmax(0.0, gmaxv));
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
        if (local && diss)
            for (long k = k0; k < k1; k++)
                if (o->DTVisc[k] > 1.e-9) {
                    o->DTVisc[k] = C_diff / o->DTVisc[k];
                    o->DT[k] = fmin(o->DT[k], o->DTVisc[k]);
                }
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
                for (long k = k0; k < k1; k++) rhs[(long)i * K + k] *= -(1.0 / o->Jdet[k]);
            if (diss) {   /* euler.go:496-501 */
                add_dissipation(o, n, rhs, k0, k1);
                limit_and_filter(o, rhs, o->LS2, k0, k1);
            }
            for (int i = 0; i < ni; i++)
                for (long k = k0; k < k1; k++) {
                    long x = n * reg1 + (long)i * K + k;
                    double r = o->RHSQ[x];
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
    double *sdv = NULL;
    if (o->diss) { sdv = (double *)malloc(sizeof(double) * (size_t)o->K); memcpy(sdv, o->DTVisc, sizeof(double) * (size_t)o->K); }
    stage(o, rk);
    memcpy(out, o->RHSQ, reg);
    if (sdv) { memcpy(o->DTVisc, sdv, sizeof(double) * (size_t)o->K); free(sdv); }
    for (int r = 0; r < 5; r++) { memcpy(o->Q[r], save[r], reg); free(save[r]); }
    memcpy(o->Residual, sres, reg); memcpy(o->DT, sdt, sizeof(double) * (size_t)o->K);
    o->GlobalDT = gdt;
    free(sres); free(sdt);
    return 0;
}
