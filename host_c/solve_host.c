/*
 * solve_host.c -- a compiled, non-Python host that drives include/dfr2d.h with exactly the call sequence of the cgo shim
 * (go/dfr2d/dfr2d.go: New -> SetState -> Step ... -> Residual -> GetState -> Close, and NewMulti -> multi_*), i.e. the
 * body of Euler.Solve's time loop (model_problems/Euler2D/euler.go:171-190) with `c.RK.Step(c)` replaced by the library.
 *
 * The Go toolchain is absent from this image, so the shim itself cannot be compiled here.  This program is what cgo
 * would generate reduced to C: it includes the header as C (not C++), fills dfr2d_problem field by field, and passes
 * plain host pointers.  tests/test_c_host.py builds it with gcc, feeds it a problem pack and compares what it writes
 * with the oracle.  It holds no numerics of its own and never touches oracle/.
 *
 * Problem pack (little endian), written by tests/c_host_pack.py:
 *   "DFR2DPK1"                                   8 bytes
 *   int64  scalar_bytes ; scalar_bytes bytes     = dfr2d_problem up to (not including) FluxEdgeInterp, as laid out by
 *                                                  the writer -- checked against offsetof() here, so a struct-layout
 *                                                  disagreement between the two sides fails loudly
 *   28 x { int64 nbytes ; data }                 the pointer members in declaration order
 *   int64 nbytes ; data                          initial state Q  [4][NpInt][K]
 *
 * usage: solve_host <pack> <out> <nsteps> [n_parts [n_devices [migrate]]]
 *   n_parts = 1 : dfr2d_create / dfr2d_set_state / nsteps x dfr2d_step(1) / dfr2d_residual / dfr2d_get_state
 *   n_parts > 1 : the MultiSolver sequence (dfr2d_multi_create / dfr2d_multi_set_state / dfr2d_multi_step /
 *                 dfr2d_multi_residual / dfr2d_multi_get_state / dfr2d_multi_destroy), partition g on device g % n_devices
 *   migrate = 1 : every step call is made from a freshly created OS thread -- a goroutine that is not locked to its
 *                 thread (no runtime.LockOSThread) migrates between OS threads from one cgo call to the next, so the
 *                 library must select its device inside every entry point (SURVEY.md 8b "Threading")
 * out: "DFR2DOUT" ; dfr2d_step_info ; double maxR[4] ; int64 n ; double Q[n]
 */
#include <pthread.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dfr2d.h"

#define POINTER_MEMBERS(X) \
    X(FluxEdgeInterp) X(DivInt) X(Div) X(V) X(Vinv) X(MassMatrix) X(D) X(P) X(ModeFilter) X(Bary) \
    X(Jdet) X(Jinv) X(FaceNormX) X(FaceNormY) X(IInII) X(EdgeLenMax) \
    X(EToV) X(edge_kL) X(edge_kR) X(edge_numL) X(edge_numR) X(edge_nconn) X(edge_bc) X(edge_len) X(EtoEdge) \
    X(bp_edge) X(bp_x) X(bp_y)

static void die(const char *what, const char *detail) {
    fprintf(stderr, "solve_host: %s%s%s\n", what, detail ? ": " : "", detail ? detail : "");
    exit(1);
}

static void *read_block(FILE *f, int64_t *nbytes) {
    if (fread(nbytes, sizeof *nbytes, 1, f) != 1 || *nbytes < 0) die("truncated pack (block length)", NULL);
    void *p = malloc(*nbytes > 0 ? (size_t)*nbytes : 1);
    if (!p) die("out of memory", NULL);
    if (*nbytes > 0 && fread(p, 1, (size_t)*nbytes, f) != (size_t)*nbytes) die("truncated pack (block data)", NULL);
    return p;
}

/* one step call, possibly on another OS thread */
typedef struct step_call {
    dfr2d_handle **hs;
    int n_parts;
    dfr2d_step_info *info;
    double *maxR; /* single partition: dfr2d_residual after the step */
    int rc;
    const char *what;
} step_call;

static void *do_step(void *arg) {
    step_call *c = (step_call *)arg;
    if (c->n_parts == 1) {
        c->what = "dfr2d_step";
        c->rc = dfr2d_step(c->hs[0], 1, c->info);
        if (!c->rc) {
            c->what = "dfr2d_residual";
            c->rc = dfr2d_residual(c->hs[0], c->maxR);
        }
    } else {
        c->what = "dfr2d_multi_step";
        c->rc = dfr2d_multi_step(c->hs, c->n_parts, 1, c->info);
    }
    return NULL;
}

static void step_once(step_call *c, int migrate) {
    if (migrate) {
        pthread_t t;
        if (pthread_create(&t, NULL, do_step, c)) die("pthread_create", NULL);
        pthread_join(t, NULL);
    } else {
        do_step(c);
    }
    if (c->rc) die(c->what, dfr2d_last_error(c->hs[0]));
    if (c->info->nan_found) die("NAN found", NULL);
}

int main(int argc, char **argv) {
    if (argc < 4) die("usage: solve_host <pack> <out> <nsteps> [n_parts [n_devices [migrate]]]", NULL);
    int migrate = argc > 6 ? atoi(argv[6]) : 0;
    int nsteps = atoi(argv[3]);
    int n_parts = argc > 4 ? atoi(argv[4]) : 1;
    int n_devices = argc > 5 ? atoi(argv[5]) : 1;
    if (nsteps < 0 || n_parts < 1 || n_devices < 1) die("bad arguments", NULL);

    FILE *f = fopen(argv[1], "rb");
    if (!f) die("cannot open pack", argv[1]);
    char magic[8];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "DFR2DPK1", 8)) die("not a problem pack", argv[1]);

    dfr2d_problem p;
    memset(&p, 0, sizeof p);
    int64_t nb;
    void *scalars = read_block(f, &nb);
    if (nb != (int64_t)offsetof(dfr2d_problem, FluxEdgeInterp))
        die("struct layout mismatch: the writer's scalar section is not offsetof(dfr2d_problem, FluxEdgeInterp)", NULL);
    memcpy(&p, scalars, (size_t)nb);
    free(scalars);
#define X(name) { void *a = read_block(f, &nb); memcpy(&p.name, &a, sizeof a); }
    POINTER_MEMBERS(X)
#undef X
    int64_t qbytes;
    double *Q = (double *)read_block(f, &qbytes);
    fclose(f);
    int64_t NpInt = (int64_t)(p.N + 1) * (p.N + 2) / 2;
    if (qbytes != 4 * NpInt * p.K * (int64_t)sizeof(double)) die("state block has the wrong size", NULL);

    dfr2d_handle **hs = (dfr2d_handle **)calloc((size_t)n_parts, sizeof *hs);
    if (n_parts == 1) {
        if (dfr2d_create(&p, 1, 0, 0, &hs[0])) die("dfr2d_create", dfr2d_last_error(NULL));
    } else {
        int *devices = (int *)malloc((size_t)n_parts * sizeof *devices);
        for (int g = 0; g < n_parts; g++) devices[g] = g % n_devices;
        if (dfr2d_multi_create(&p, n_parts, devices, hs)) die("dfr2d_multi_create", dfr2d_last_error(NULL));
        free(devices);
    }

    dfr2d_step_info info;
    memset(&info, 0, sizeof info);
    double maxR[4] = {0, 0, 0, 0};
    step_call call = {hs, n_parts, &info, maxR, 0, ""};
    if (n_parts == 1) {
        if (dfr2d_set_state(hs[0], Q)) die("dfr2d_set_state", dfr2d_last_error(hs[0]));
        /* for !finished { c.RK.Step(c); steps++; ... PrintUpdate every n }  (euler.go:175-186) */
        for (int s = 0; s < nsteps && !info.finished; s++) step_once(&call, migrate);
        memset(Q, 0, (size_t)qbytes);
        if (dfr2d_get_state(hs[0], Q)) die("dfr2d_get_state", dfr2d_last_error(hs[0]));
    } else {
        if (dfr2d_multi_set_state(hs, n_parts, Q)) die("dfr2d_multi_set_state", dfr2d_last_error(hs[0]));
        for (int s = 0; s < nsteps && !info.finished; s++) step_once(&call, migrate);
        /* MultiSolver.Residual: signed max over the partitions' own maxima */
        if (dfr2d_multi_residual(hs, n_parts, maxR)) die("dfr2d_multi_residual", dfr2d_last_error(hs[0]));
        memset(Q, 0, (size_t)qbytes);
        if (dfr2d_multi_get_state(hs, n_parts, Q)) die("dfr2d_multi_get_state", dfr2d_last_error(hs[0]));
    }
    int64_t launches = 0;
    for (int g = 0; g < n_parts; g++) launches += dfr2d_launch_count(hs[g]);
    dfr2d_multi_destroy(hs, n_parts);

    f = fopen(argv[2], "wb");
    if (!f) die("cannot open output", argv[2]);
    int64_t n = qbytes / (int64_t)sizeof(double);
    fwrite("DFR2DOUT", 1, 8, f);
    fwrite(&info, sizeof info, 1, f);
    fwrite(maxR, sizeof maxR, 1, f);
    fwrite(&n, sizeof n, 1, f);
    fwrite(Q, sizeof(double), (size_t)n, f);
    fclose(f);
    printf("solve_host: %d partition(s), %lld step(s), time %.17g, %lld kernel launches\n", n_parts,
           (long long)info.steps, info.time, (long long)launches);
    return 0;
}
