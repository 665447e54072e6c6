/*
 * ba_oracle.c -- CPU ORACLE for the MegBA LM / Schur-PCG hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It may be imported / linked /
 * executed only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs, and there only as the checker or the timed CPU arm.
 * The product (megba_b200/, include/) never calls into it and has no CPU fallback.
 *
 * What it is: a plain-C (+OpenMP) restatement of the arithmetic of
 *   src/geo/analytical_derivatives.cu:16-285      residual + analytical Jacobian
 *   src/edge/build_linear_system.cu:47-239        Hpp/Hll/Hpl/g assembly (+ Omega)
 *   src/linear_system/schur_LM_linear_system.cu:114-218   damping, backup, update
 *   src/solver/schur_pcg_solver.cu:61-639         inverse, MakeV, PCG, SolveW
 *   src/solver/implicit_schur_pcg_solver.cu:21-90 matrix-free operators
 *   src/algo/lm_algo.cu:26-223                    LM loop and its scalars
 * of MegviiRobot/MegBA, each function citing the lines it follows.
 *
 * Pinning: upstream ships no tests, golden vectors or fixtures for this path and has no CPU execution path, so it
 * cannot be run here as a whole.  The oracle is pinned by
 *   (i)  the reference's own CUDA sources for rows a2-a8, a10, a11 (analytical_derivatives.cu, build_linear_system.cu,
 *        schur_pcg_solver.cu, schur_LM_linear_system.cu, lm_algo.cu/.cpp -- i.e. its whole LM loop) compiled
 *        UNMODIFIED against shim headers into oracle/_ref/ (oracle/shim/README.md) and run on the GPU box by
 *        tests/test_gpu_reference_pin.py against this oracle and against the product;
 *   (ii) independent mathematics in tests/test_oracle.py (central finite differences on J, dense normal-equation
 *        solve vs PCG, independent projection, cost monotonicity).
 * Nothing on the path is left pinned by restatement alone except the trivial per-edge parameter update (a9).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int64_t nC, nP, nE;
  int32_t *cam_idx, *pt_idx; /* [nE] copies */
  int64_t *pt_ptr, *pt_edges; /* edges grouped by point  (Hlp rows)  */
  int64_t *cam_ptr, *cam_edges; /* edges grouped by camera (Hpl rows)  */
} orc_graph;

typedef struct {
  double cost;     /* F = sum r^2 after this iteration (printed error = F/2) */
  int accepted;    /* 1 accept, 0 reject ("Iter k failed"), -1 step-size stop */
  int pcg_iters;   /* n of schur_pcg_solver.cu:405 for this iteration */
  double region;   /* trust region after the update */
  double dx_norm;  /* ||dx||_2 */
  double rho;      /* gain ratio */
} orc_lm_trace;

/* Edge -> vertex incidence lists.  Stands in for HessianEntrance /
 * buildPositionContainer (src/problem/base_problem.cpp:183-214,
 * src/edge/base_edge.cpp:224-262): only the grouping matters to the arithmetic. */
orc_graph *orc_graph_create(int64_t nC, int64_t nP, int64_t nE, const int32_t *cam_idx,
                            const int32_t *pt_idx) {
  orc_graph *G = (orc_graph *)calloc(1, sizeof(orc_graph));
  G->nC = nC; G->nP = nP; G->nE = nE;
  G->cam_idx = (int32_t *)malloc(sizeof(int32_t) * (nE ? nE : 1));
  G->pt_idx = (int32_t *)malloc(sizeof(int32_t) * (nE ? nE : 1));
  memcpy(G->cam_idx, cam_idx, sizeof(int32_t) * nE);
  memcpy(G->pt_idx, pt_idx, sizeof(int32_t) * nE);
  G->pt_ptr = (int64_t *)calloc(nP + 1, sizeof(int64_t));
  G->cam_ptr = (int64_t *)calloc(nC + 1, sizeof(int64_t));
  G->pt_edges = (int64_t *)malloc(sizeof(int64_t) * (nE ? nE : 1));
  G->cam_edges = (int64_t *)malloc(sizeof(int64_t) * (nE ? nE : 1));
  for (int64_t e = 0; e < nE; ++e) { G->pt_ptr[pt_idx[e] + 1]++; G->cam_ptr[cam_idx[e] + 1]++; }
  for (int64_t p = 0; p < nP; ++p) G->pt_ptr[p + 1] += G->pt_ptr[p];
  for (int64_t c = 0; c < nC; ++c) G->cam_ptr[c + 1] += G->cam_ptr[c];
  int64_t *fp = (int64_t *)malloc(sizeof(int64_t) * (nP + 1));
  int64_t *fc = (int64_t *)malloc(sizeof(int64_t) * (nC + 1));
  memcpy(fp, G->pt_ptr, sizeof(int64_t) * (nP + 1));
  memcpy(fc, G->cam_ptr, sizeof(int64_t) * (nC + 1));
  for (int64_t e = 0; e < nE; ++e) {
    G->pt_edges[fp[pt_idx[e]]++] = e;
    G->cam_edges[fc[cam_idx[e]]++] = e;
  }
  free(fp); free(fc);
  return G;
}

void orc_graph_destroy(orc_graph *G) {
  if (!G) return;
  free(G->cam_idx); free(G->pt_idx); free(G->pt_ptr); free(G->pt_edges);
  free(G->cam_ptr); free(G->cam_edges); free(G);
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---- fp64 instantiation ---- */
#define R double
#define SFX _f64
#define R_EPS 2.220446049250313e-16 /* std::numeric_limits<double>::epsilon() */
#define R_SQRT sqrt
#define R_SIN sin
#define R_COS cos
#define R_ABS fabs
#include "ba_oracle_impl.h"
#undef R
#undef SFX
#undef R_EPS
#undef R_SQRT
#undef R_SIN
#undef R_COS
#undef R_ABS

/* ---- fp32 instantiation ---- */
#define R float
#define SFX _f32
#define R_EPS 1.1920928955078125e-07f
#define R_SQRT sqrtf
#define R_SIN sinf
#define R_COS cosf
#define R_ABS fabsf
#include "ba_oracle_impl.h"
