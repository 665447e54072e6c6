/*
 * ba_oracle_impl.h -- TEST INFRASTRUCTURE ONLY (see ba_oracle.c header).
 *
 * Type-generic body of the CPU oracle.  Included twice by ba_oracle.c with
 *   R   = double / float      (the reference's template parameter T)
 *   SFX = _f64   / _f32
 * Every function cites the reference file:line whose arithmetic it restates
 * (paths relative to the MegBA checkout, e.g. /root/reference).
 *
 * Layouts used by the oracle (plain AoS, chosen for readability, not speed):
 *   cam[nC][9]  = {angle-axis(3), t(3), f, k1, k2}      pt[nP][3]
 *   obs[nE][2], cam_idx[nE], pt_idx[nE]
 *   r[nE][2], J[nE][2][12]   (12 = 9 camera columns then 3 point columns)
 *   Hpp[nC][9][9] row-major, Hll[nP][3][3] row-major, W[nE][9][3] (= Hpl block)
 *   g[9nC + 3nP]  (g = -J^T r)
 */

#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)
#define FN(name) CAT(name, SFX)

/* ------------------------------------------------------------------------- */
/* a2: residual + analytical Jacobian of ONE observation.
 * Restates src/geo/analytical_derivatives.cu:16-159 (AngleAxisRotatePoint with
 * derivatives; first-order branch when theta^2 <= eps, :24,131-158) and
 * :186-284 (projection, radial distortion, the three "error factors").      */
static void FN(edge_rj)(const R *cam, const R *X, const R *ob, R *res, R *Jout) {
  const R *aa = cam;
  R P[3], dP_daa[3][3], dP_dX[3][3];
  const R theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (theta2 > R_EPS) {
    const R theta = R_SQRT(theta2);
    const R s = R_SIN(theta), c = R_COS(theta);
    const R ti = (R)1.0 / theta;
    const R nti2 = -ti * ti;
    const R omc = (R)1.0 - c;
    R w[3], wxp[3], dth[3], dw[3][3], dwxp_daa[3][3], dtmp_daa[3];
    for (int i = 0; i < 3; ++i) w[i] = aa[i] * ti;
    wxp[0] = w[1] * X[2] - w[2] * X[1];
    wxp[1] = w[2] * X[0] - w[0] * X[2];
    wxp[2] = w[0] * X[1] - w[1] * X[0];
    const R wdotx = w[0] * X[0] + w[1] * X[1] + w[2] * X[2];
    const R tmp = wdotx * omc;
    for (int i = 0; i < 3; ++i) P[i] = X[i] * c + wxp[i] * s + w[i] * tmp;

    for (int j = 0; j < 3; ++j) dth[j] = ti * aa[j];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        dw[i][j] = (i == j ? ti : (R)0) + nti2 * aa[i] * dth[j];
    /* d(w x X)/dX is the skew matrix of w */
    const R skw[3][3] = {{0, -w[2], w[1]}, {w[2], 0, -w[0]}, {-w[1], w[0], 0}};
    for (int j = 0; j < 3; ++j) {
      dwxp_daa[0][j] = X[2] * dw[1][j] - X[1] * dw[2][j];
      dwxp_daa[1][j] = -X[2] * dw[0][j] + X[0] * dw[2][j];
      dwxp_daa[2][j] = X[1] * dw[0][j] - X[0] * dw[1][j];
      dtmp_daa[j] = s * wdotx * dth[j] +
                    omc * (X[0] * dw[0][j] + X[1] * dw[1][j] + X[2] * dw[2][j]);
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        const R dc = -s * dth[j], ds = c * dth[j];
        dP_daa[i][j] = X[i] * dc + wxp[i] * ds + s * dwxp_daa[i][j] +
                       w[i] * dtmp_daa[j] + tmp * dw[i][j];
        dP_dX[i][j] = (i == j ? c : (R)0) + s * skw[i][j] + w[i] * (omc * w[j]);
      }
  } else {
    /* analytical_derivatives.cu:131-158 */
    P[0] = X[0] + (aa[1] * X[2] - aa[2] * X[1]);
    P[1] = X[1] + (aa[2] * X[0] - aa[0] * X[2]);
    P[2] = X[2] + (aa[0] * X[1] - aa[1] * X[0]);
    const R a[3][3] = {{0, X[2], -X[1]}, {-X[2], 0, X[0]}, {X[1], -X[0], 0}};
    const R b[3][3] = {{1, -aa[2], aa[1]}, {aa[2], 1, -aa[0]}, {-aa[1], aa[0], 1}};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        dP_daa[i][j] = a[i][j];
        dP_dX[i][j] = b[i][j];
      }
  }
  /* analytical_derivatives.cu:186-203 */
  for (int i = 0; i < 3; ++i) P[i] += cam[3 + i];
  const R f = cam[6], k1 = cam[7], k2 = cam[8];
  const R B[2] = {-P[0] / P[2], -P[1] / P[2]};
  const R l2 = B[0] * B[0] + B[1] * B[1];
  const R rp = (R)1. + k1 * l2 + k2 * l2 * l2;
  const R fr = f * rp;
  res[0] = fr * B[0] - ob[0];
  res[1] = fr * B[1] - ob[1];
  const R tmpB[2] = {f * B[0] * (k1 + 2 * k2 * l2), f * B[1] * (k1 + 2 * k2 * l2)};
  const R nrp2 = -1 / P[2];
  const R p0d = P[0] / (P[2] * P[2]);
  const R p1d = P[1] / (P[2] * P[2]);
  for (int row = 0; row < 2; ++row) {
    R ef[3];
    if (row == 0) { /* :205-208 */
      ef[0] = nrp2 * (fr + 2 * tmpB[0] * B[0]);
      ef[1] = nrp2 * 2 * tmpB[0] * B[1];
      ef[2] = p0d * (fr + 2 * tmpB[0] * B[0]) + p1d * 2 * tmpB[0] * B[1];
    } else { /* :246-249 */
      ef[0] = nrp2 * 2 * tmpB[1] * B[0];
      ef[1] = nrp2 * (fr + 2 * tmpB[1] * B[1]);
      ef[2] = p0d * 2 * tmpB[1] * B[0] + p1d * (fr + 2 * tmpB[1] * B[1]);
    }
    R *Jr = Jout + 12 * row;
    for (int j = 0; j < 3; ++j) {
      Jr[j] = ef[0] * dP_daa[0][j] + ef[1] * dP_daa[1][j] + ef[2] * dP_daa[2][j];
      Jr[3 + j] = ef[j];
      Jr[9 + j] = ef[0] * dP_dX[0][j] + ef[1] * dP_dX[1][j] + ef[2] * dP_dX[2][j];
    }
    Jr[6] = rp * B[row];
    Jr[7] = f * l2 * B[row];
    Jr[8] = f * l2 * l2 * B[row];
  }
}

/* residual only (same arithmetic as edge_rj's value path) */
static void FN(edge_r)(const R *cam, const R *X, const R *ob, R *res) {
  R J[24];
  FN(edge_rj)(cam, X, ob, res, J);
}

/* a2 over the whole batch: EdgeVector::forward -> AnalyticalDerivativesKernelMatrix
 * (src/edge/base_edge.cpp:160-163, analytical_derivatives.cu:289-322).        */
void FN(orc_forward)(int64_t nE, const int32_t *cam_idx, const int32_t *pt_idx,
                     const R *obs, const R *cam, const R *pt, R *r, R *J) {
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < nE; ++e) {
    R Jl[24], rl[2];
    FN(edge_rj)(cam + 9 * (int64_t)cam_idx[e], pt + 3 * (int64_t)pt_idx[e],
                obs + 2 * e, rl, Jl);
    r[2 * e] = rl[0];
    r[2 * e + 1] = rl[1];
    if (J) memcpy(J + 24 * e, Jl, sizeof(Jl));
  }
}

/* a10: F = sum r^2  (computeResidualNorm, src/algo/lm_algo.cu:26-51) */
double FN(orc_cost)(int64_t nE, const R *r) {
  double acc = 0.;
#pragma omp parallel for reduction(+ : acc) schedule(static)
  for (int64_t i = 0; i < 2 * nE; ++i) acc += (double)(r[i] * r[i]);
  return acc;
}

/* a3 / a3': Hessian + gradient assembly.
 * Restates makeHSchur (src/edge/build_linear_system.cu:88-146) with the device
 * helpers makeHpp :48-53, makeHpl :56-65, makeHll :79-85; sign g = -J^T r :135.
 * Layouts :119,127-128.  `info` (per-edge 2x2, row-major, may be NULL) selects
 * the JMulInfo/makeHSchurWithInfo path (:149-239): left factor = Omega*J where
 * left_i = sum_j J_j * Omega[i + 2j] (:167-174).  W may be NULL (implicit, a3').
 * Hlp is W transposed and is not stored separately.                            */
void FN(orc_build)(const orc_graph *G, const R *r, const R *J, const R *info,
                   R *Hpp, R *Hll, R *W, R *g) {
  const int64_t nC = G->nC, nP = G->nP;
  /* camera side: parallel over cameras, edges visited in list order */
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t c = 0; c < nC; ++c) {
    R H[81], gc[9];
    for (int i = 0; i < 81; ++i) H[i] = 0;
    for (int i = 0; i < 9; ++i) gc[i] = 0;
    for (int64_t k = G->cam_ptr[c]; k < G->cam_ptr[c + 1]; ++k) {
      const int64_t e = G->cam_edges[k];
      const R *Je = J + 24 * e;
      for (int row = 0; row < 2; ++row) {
        R L[12];
        for (int y = 0; y < 12; ++y)
          L[y] = info ? Je[y] * info[4 * e + row] + Je[12 + y] * info[4 * e + row + 2]
                      : Je[12 * row + y];
        const R *Rr = Je + 12 * row;
        for (int y = 0; y < 9; ++y) {
          for (int i = 0; i < 9; ++i) H[9 * y + i] += L[y] * Rr[i];
          gc[y] += -L[y] * r[2 * e + row];
        }
      }
    }
    memcpy(Hpp + 81 * c, H, sizeof(H));
    memcpy(g + 9 * c, gc, sizeof(gc));
  }
  /* point side + coupling block */
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t p = 0; p < nP; ++p) {
    R H[9], gp[3];
    for (int i = 0; i < 9; ++i) H[i] = 0;
    for (int i = 0; i < 3; ++i) gp[i] = 0;
    for (int64_t k = G->pt_ptr[p]; k < G->pt_ptr[p + 1]; ++k) {
      const int64_t e = G->pt_edges[k];
      const R *Je = J + 24 * e;
      R We[27];
      for (int i = 0; i < 27; ++i) We[i] = 0;
      for (int row = 0; row < 2; ++row) {
        R L[12];
        for (int y = 0; y < 12; ++y)
          L[y] = info ? Je[y] * info[4 * e + row] + Je[12 + y] * info[4 * e + row + 2]
                      : Je[12 * row + y];
        const R *Rr = Je + 12 * row;
        for (int y = 0; y < 3; ++y) {
          for (int i = 0; i < 3; ++i) H[3 * y + i] += L[9 + y] * Rr[9 + i];
          gp[y] += -L[9 + y] * r[2 * e + row];
        }
        for (int y = 0; y < 9; ++y)
          for (int i = 0; i < 3; ++i) We[3 * y + i] += L[y] * Rr[9 + i];
      }
      if (W) memcpy(W + 27 * e, We, sizeof(We));
    }
    memcpy(Hll + 9 * p, H, sizeof(H));
    memcpy(g + 9 * nC + 3 * p, gp, sizeof(gp));
  }
}

/* a4: Marquardt damping of the block diagonals.
 * Restates ExtractOldAndApplyNewDiagKernel / RecoverDiagKernel
 * (src/linear_system/schur_LM_linear_system.cu:114-150) and processDiag :162-185:
 * diag <- (1/region + 1) * diag, old diagonal saved / reused on `recover`.      */
void FN(orc_process_diag)(int64_t num, int dim, double region, int recover,
                          R *H, R *diag) {
  const R a = (R)(1. / region);
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < num; ++b)
    for (int i = 0; i < dim; ++i) {
      R *h = H + b * dim * dim + i * dim + i;
      if (recover) {
        *h = (a + 1) * diag[b * dim + i];
      } else {
        const R d = *h;
        diag[b * dim + i] = d;
        *h = (a + 1) * d;
      }
    }
}

/* a5: batched dense inverse of n x n blocks (n = 3 or 9).
 * The reference calls cublas{S,D}matinvBatched (src/solver/schur_pcg_solver.cu:
 * 61-97); that routine is closed-source, so this is Gauss-Jordan with partial
 * pivoting -- the published textbook algorithm for the same quantity.  Blocks are
 * symmetric, so row-/column-major interpretation coincide (:116).               */
void FN(orc_block_inverse)(int64_t num, int n, const R *A, R *Ainv) {
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < num; ++b) {
    R M[9][18];
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        M[i][j] = A[b * n * n + i * n + j];
        M[i][n + j] = (i == j) ? (R)1 : (R)0;
      }
    for (int col = 0; col < n; ++col) {
      int piv = col;
      for (int i = col + 1; i < n; ++i)
        if (R_ABS(M[i][col]) > R_ABS(M[piv][col])) piv = i;
      if (piv != col)
        for (int j = 0; j < 2 * n; ++j) {
          const R t = M[col][j];
          M[col][j] = M[piv][j];
          M[piv][j] = t;
        }
      const R inv = (R)1 / M[col][col];
      for (int j = 0; j < 2 * n; ++j) M[col][j] *= inv;
      for (int i = 0; i < n; ++i) {
        if (i == col) continue;
        const R fct = M[i][col];
        if (fct != (R)0)
          for (int j = 0; j < 2 * n; ++j) M[i][j] -= fct * M[col][j];
      }
    }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) Ainv[b * n * n + i * n + j] = M[i][n + j];
  }
}

/* y_b = B_b x_b for block-diagonal B (oursGgemvBatched, schur_pcg_solver.cu:100-121) */
static void FN(block_gemv)(int64_t num, int n, const R *B, const R *x, R *y) {
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < num; ++b) {
    R out[9];
    for (int i = 0; i < n; ++i) {
      R s = 0;
      for (int j = 0; j < n; ++j) s += B[b * n * n + i * n + j] * x[b * n + j];
      out[i] = s;
    }
    for (int i = 0; i < n; ++i) y[b * n + i] = out[i];
  }
}

/* t = Hlp * xc  (3nP <- 9nC).  Explicit: cusparseSpMV on Hlp, schur_pcg_solver.cu:
 * 206,320,563 (Hlp block = W^T, build_linear_system.cu:68-76).  Implicit (W==NULL):
 * implicitETMulx, src/solver/implicit_schur_pcg_solver.cu:57-90.                */
void FN(orc_hlp_mul)(const orc_graph *G, const R *W, const R *J, const R *xc, R *t) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t p = 0; p < G->nP; ++p) {
    R acc[3] = {0, 0, 0};
    for (int64_t k = G->pt_ptr[p]; k < G->pt_ptr[p + 1]; ++k) {
      const int64_t e = G->pt_edges[k];
      const R *x = xc + 9 * (int64_t)G->cam_idx[e];
      if (W) {
        for (int i = 0; i < 3; ++i)
          for (int y = 0; y < 9; ++y) acc[i] += W[27 * e + 3 * y + i] * x[y];
      } else {
        for (int row = 0; row < 2; ++row) {
          const R *Jr = J + 24 * e + 12 * row;
          R jcx = 0;
          for (int y = 0; y < 9; ++y) jcx += Jr[y] * x[y];
          for (int i = 0; i < 3; ++i) acc[i] += Jr[9 + i] * jcx;
        }
      }
    }
    for (int i = 0; i < 3; ++i) t[3 * p + i] = acc[i];
  }
}

/* q = Hpl * yp  (9nC <- 3nP).  Explicit: cusparseSpMV on Hpl, schur_pcg_solver.cu:
 * 230,344,490.  Implicit: implicitEMulx, implicit_schur_pcg_solver.cu:21-54.    */
void FN(orc_hpl_mul)(const orc_graph *G, const R *W, const R *J, const R *yp, R *q) {
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t c = 0; c < G->nC; ++c) {
    R acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t k = G->cam_ptr[c]; k < G->cam_ptr[c + 1]; ++k) {
      const int64_t e = G->cam_edges[k];
      const R *y = yp + 3 * (int64_t)G->pt_idx[e];
      if (W) {
        for (int yy = 0; yy < 9; ++yy)
          for (int i = 0; i < 3; ++i) acc[yy] += W[27 * e + 3 * yy + i] * y[i];
      } else {
        for (int row = 0; row < 2; ++row) {
          const R *Jr = J + 24 * e + 12 * row;
          R jpx = 0;
          for (int i = 0; i < 3; ++i) jpx += Jr[9 + i] * y[i];
          for (int yy = 0; yy < 9; ++yy) acc[yy] += Jr[yy] * jpx;
        }
      }
    }
    for (int i = 0; i < 9; ++i) q[9 * c + i] = acc[i];
  }
}

/* Camera-major copies of the coupling data.  The reference keeps TWO value arrays, Hpl (camera-major rows) and Hlp
 * (point-major rows), so that both SpMVs stream their matrix (include/linear_system/schur_linear_system.h:22-27);
 * the oracle does the same for its timed solve: Wcm[k] / Jcm[k] is the block of edge cam_edges[k].               */
static R *FN(camera_major_copy)(const orc_graph *G, const R *src, int width) {
  R *dst = (R *)malloc(sizeof(R) * (size_t)width * (size_t)(G->nE ? G->nE : 1));
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < G->nE; ++k) memcpy(dst + (size_t)width * k, src + (size_t)width * G->cam_edges[k], sizeof(R) * width);
  return dst;
}

static void FN(hpl_mul_cm)(const orc_graph *G, const R *Wcm, const R *Jcm, const R *yp, R *q) {
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t c = 0; c < G->nC; ++c) {
    R acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t k = G->cam_ptr[c]; k < G->cam_ptr[c + 1]; ++k) {
      const R *y = yp + 3 * (int64_t)G->pt_idx[G->cam_edges[k]];
      if (Wcm) {
        const R *We = Wcm + 27 * k;
        for (int yy = 0; yy < 9; ++yy)
          for (int i = 0; i < 3; ++i) acc[yy] += We[3 * yy + i] * y[i];
      } else {
        for (int row = 0; row < 2; ++row) {
          const R *Jr = Jcm + 24 * k + 12 * row;
          R jpx = 0;
          for (int i = 0; i < 3; ++i) jpx += Jr[9 + i] * y[i];
          for (int yy = 0; yy < 9; ++yy) acc[yy] += Jr[yy] * jpx;
        }
      }
    }
    for (int i = 0; i < 9; ++i) q[9 * c + i] = acc[i];
  }
}

/* q = S p = Hpp p - Hpl Hll^-1 Hlp p  (one PCG operator application,
 * schur_pcg_solver.cu:315-366).  tmp: 3nP scratch.                              */
void FN(orc_schur_mul)(const orc_graph *G, const R *HppD, const R *HllInv, const R *W,
                       const R *J, const R *p, R *tmp, R *q) {
  const int64_t nC = G->nC, nP = G->nP;
  FN(orc_hlp_mul)(G, W, J, p, tmp);
  FN(block_gemv)(nP, 3, HllInv, tmp, tmp);
  FN(orc_hpl_mul)(G, W, J, tmp, q);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nC; ++c)
    for (int i = 0; i < 9; ++i) {
      R s = 0;
      for (int j = 0; j < 9; ++j) s += HppD[81 * c + 9 * i + j] * p[9 * c + j];
      q[9 * c + i] = s - q[9 * c + i]; /* oursGgemvBatched<T,1,-1> :363-365 */
    }
}

/* same operator as orc_schur_mul, with the Hpl product streaming the camera-major copy */
static void FN(schur_mul_cm)(const orc_graph *G, const R *HppD, const R *HllInv, const R *W, const R *J, const R *Wcm,
                             const R *Jcm, const R *p, R *tmp, R *q) {
  const int64_t nC = G->nC, nP = G->nP;
  FN(orc_hlp_mul)(G, W, J, p, tmp);
  FN(block_gemv)(nP, 3, HllInv, tmp, tmp);
  FN(hpl_mul_cm)(G, Wcm, Jcm, tmp, q);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nC; ++c)
    for (int i = 0; i < 9; ++i) {
      R s = 0;
      for (int j = 0; j < 9; ++j) s += HppD[81 * c + 9 * i + j] * p[9 * c + j];
      q[9 * c + i] = s - q[9 * c + i];
    }
}

static double FN(dotd)(int64_t n, const R *a, const R *b) {
  /* cublas dot returns T; products formed in T, summed here in double so the
   * oracle's own summation noise stays below the parity tolerances */
  double acc = 0;
#pragma omp parallel for reduction(+ : acc) schedule(static)
  for (int64_t i = 0; i < n; ++i) acc += (double)(a[i] * b[i]);
  return acc;
}

/* a5-a8: SchurPCGSolverDistributed (schur_pcg_solver.cu:599-639) =
 *   Hll^-1 (:614-620) -> schurMakeVDistributed (:430-510) ->
 *   schurPCGSolverDistributedCUDA (:124-427) -> schurSolveWDistributed (:513-596).
 * HppD / HllD are the DAMPED blocks.  g is NOT destroyed here (the reference
 * overwrites g_l with w and g_c with v; callers never read them again before the
 * next rebuild/rollback, so keeping g intact is observationally identical).
 * dx[0..9nC) enters as the warm start x0 (:193,206) and leaves as the camera
 * step; dx[9nC..) leaves as the point step.  Returns the reference's `done`.   */
int FN(orc_schur_pcg)(const orc_graph *G, const R *HppD, const R *HllD, const R *W,
                      const R *J, const R *g, int max_iter, double tol,
                      double refuse_ratio, R *dx, int *iters_out, double *rho_out) {
  const int64_t nC = G->nC, nP = G->nP, nc9 = 9 * nC, np3 = 3 * nP;
  R *HllInv = (R *)malloc(sizeof(R) * 9 * nP);
  R *HppInv = (R *)malloc(sizeof(R) * 81 * nC);
  R *w = (R *)malloc(sizeof(R) * np3);
  R *tmp = (R *)malloc(sizeof(R) * np3);
  R *v = (R *)malloc(sizeof(R) * nc9);
  R *rN = (R *)malloc(sizeof(R) * nc9);
  R *pN = (R *)malloc(sizeof(R) * nc9);
  R *ax = (R *)malloc(sizeof(R) * nc9);
  R *xbk = (R *)malloc(sizeof(R) * nc9);
  R *Wcm = W ? FN(camera_major_copy)(G, W, 27) : NULL;
  R *Jcm = W ? NULL : FN(camera_major_copy)(G, J, 24);
  FN(orc_block_inverse)(nP, 3, HllD, HllInv);
  /* MakeV: w = Hll^-1 g_l ; v = g_c - Hpl w   (:467-493, beta = 1/worldSize = 1) */
  FN(block_gemv)(nP, 3, HllInv, g + nc9, w);
  FN(hpl_mul_cm)(G, Wcm, Jcm, w, v);
  for (int64_t i = 0; i < nc9; ++i) v[i] = g[i] - v[i];
  /* PCG */
  FN(orc_block_inverse)(nC, 9, HppD, HppInv); /* :199 */
  FN(schur_mul_cm)(G, HppD, HllInv, W, J, Wcm, Jcm, dx, tmp, ax); /* :202-251 with x0 */
  for (int64_t i = 0; i < nc9; ++i) rN[i] = v[i] - ax[i]; /* :182,256 */
  int n = 0, done = 0;
  R rhoN = 0, rhoNm1 = 0;
  R rhoMin = (R)INFINITY;
  memcpy(xbk, dx, sizeof(R) * nc9); /* reference leaves this uninitialised */
  do {
    FN(block_gemv)(nC, 9, HppInv, rN, ax);     /* z = M^-1 r :273-275 */
    rhoN = (R)FN(dotd)(nc9, rN, ax);           /* :277-287 */
    if (rhoN > (R)refuse_ratio * rhoMin) {     /* :288-295 */
      memcpy(dx, xbk, sizeof(R) * nc9);
      break;
    }
    rhoMin = rhoN < rhoMin ? rhoN : rhoMin;    /* :296 */
    if (n >= 1) {
      const R beta = rhoN / rhoNm1;            /* :299-306 */
      for (int64_t i = 0; i < nc9; ++i) pN[i] = ax[i] + beta * pN[i];
    } else {
      memcpy(pN, ax, sizeof(R) * nc9);         /* :310 */
    }
    FN(schur_mul_cm)(G, HppD, HllInv, W, J, Wcm, Jcm, pN, tmp, ax); /* q :315-366 */
    const double dot_sum = FN(dotd)(nc9, pN, ax);          /* :368-384 */
    const R alpha = (R)(rhoN / dot_sum);                   /* :385 */
    memcpy(xbk, dx, sizeof(R) * nc9);                      /* :389 */
    for (int64_t i = 0; i < nc9; ++i) {
      dx[i] += alpha * pN[i];                              /* :391 */
      rN[i] += -alpha * ax[i];                             /* :400 */
    }
    rhoNm1 = rhoN;
    ++n;
    done = R_ABS(rhoN) < (R)tol;                           /* :406 */
  } while (!done && n < max_iter);
  /* SolveW: dx_l = w - Hll^-1 (Hlp dx_c)   (:559-585) */
  FN(orc_hlp_mul)(G, W, J, dx, tmp);
  FN(block_gemv)(nP, 3, HllInv, tmp, tmp);
  for (int64_t i = 0; i < np3; ++i) dx[nc9 + i] = w[i] - tmp[i];
  if (iters_out) *iters_out = n;
  if (rho_out) *rho_out = (double)rhoN;
  free(HllInv); free(HppInv); free(w); free(tmp); free(v);
  free(rN); free(pN); free(ax); free(xbk); free(Wcm); free(Jcm);
  return done;
}

/* a10: rho denominator  sum (J dx + r)^2  (JdxpF, src/algo/lm_algo.cu:61-79;
 * computeRhoDenominator :82-126 -- the caller subtracts F).                   */
double FN(orc_jdx_plus_r_sq)(const orc_graph *G, const R *J, const R *r, const R *dx) {
  const int64_t nC = G->nC;
  double acc = 0;
#pragma omp parallel for reduction(+ : acc) schedule(static)
  for (int64_t e = 0; e < G->nE; ++e) {
    const R *dc = dx + 9 * (int64_t)G->cam_idx[e];
    const R *dp = dx + 9 * nC + 3 * (int64_t)G->pt_idx[e];
    for (int row = 0; row < 2; ++row) {
      const R *Jr = J + 24 * e + 12 * row;
      R s = 0;
      for (int i = 0; i < 9; ++i) s += Jr[i] * dc[i];
      for (int i = 0; i < 3; ++i) s += Jr[9 + i] * dp[i];
      acc += (double)((s + r[2 * e + row]) * (s + r[2 * e + row]));
    }
  }
  return acc;
}

static double FN(l2sq)(int64_t n, const R *a) { return FN(dotd)(n, a, a); }

static double FN(linf)(int64_t n, const R *a) {
  double m = 0;
#pragma omp parallel for reduction(max : m) schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const double v = fabs((double)a[i]);
    if (v > m) m = v;
  }
  return m;
}

/* a11: the LM control law, LMAlgo::solveCUDA (src/algo/lm_algo.cu:139-223), with
 * EdgeVector::update (src/edge/update.cu:14-41: trial params = x + dx),
 * backup/rollback (schur_LM_linear_system.cu:188-209) and applyUpdate (:212-218).
 * x = [cam(9nC) | pt(3nP)] is updated in place.  `implicit` selects a3'/a7-implicit.
 * trace (may be NULL) receives one orc_lm_trace row per iteration, row 0 = start. */
int FN(orc_lm_solve)(const orc_graph *G, const R *obs, R *x, int implicit,
                     int lm_max_iter, double initial_region, double eps1, double eps2,
                     int pcg_max_iter, double pcg_tol, double pcg_refuse,
                     orc_lm_trace *trace, int *n_trace) {
  const int64_t nC = G->nC, nP = G->nP, nE = G->nE, nX = 9 * nC + 3 * nP;
  R *r = (R *)malloc(sizeof(R) * 2 * nE), *J = (R *)malloc(sizeof(R) * 24 * nE);
  R *r2 = (R *)malloc(sizeof(R) * 2 * nE), *J2 = (R *)malloc(sizeof(R) * 24 * nE);
  R *W = implicit ? NULL : (R *)malloc(sizeof(R) * 27 * nE);
  R *Hpp = (R *)malloc(sizeof(R) * 81 * nC), *Hll = (R *)malloc(sizeof(R) * 9 * nP);
  R *dgc = (R *)malloc(sizeof(R) * 9 * nC), *dgp = (R *)malloc(sizeof(R) * 3 * nP);
  R *g = (R *)malloc(sizeof(R) * nX), *gbk = (R *)malloc(sizeof(R) * nX);
  R *dx = (R *)calloc(nX, sizeof(R)), *dxbk = (R *)malloc(sizeof(R) * nX);
  R *xt = (R *)malloc(sizeof(R) * nX);
  double region = initial_region; /* lm_algo.cpp:15 */
  int recover = 0, nt = 0;
  FN(orc_forward)(nE, G->cam_idx, G->pt_idx, obs, x, x + 9 * nC, r, J); /* :145 */
  FN(orc_build)(G, r, J, NULL, Hpp, Hll, W, g);                          /* :147 */
  double Fnew = FN(orc_cost)(nE, r), F = Fnew;                           /* :148 */
  if (trace) {
    orc_lm_trace t0 = {Fnew, 1, 0, region, 0, 0};
    trace[nt++] = t0;
  }
  int stop = 0, k = 0;
  double v = 2.;
  memcpy(dxbk, dx, sizeof(R) * nX); /* :155 */
  memcpy(gbk, g, sizeof(R) * nX);
  while (!stop && k < lm_max_iter) { /* :163 */
    ++k;
    FN(orc_process_diag)(nC, 9, region, recover, Hpp, dgc); /* :165 */
    FN(orc_process_diag)(nP, 3, region, recover, Hll, dgp);
    int pcg_it = 0;
    double rho_last = 0;
    FN(orc_schur_pcg)(G, Hpp, Hll, W, implicit ? J : NULL, g, pcg_max_iter, pcg_tol,
                      pcg_refuse, dx, &pcg_it, &rho_last); /* :166-169 */
    const double dxl2 = sqrt(FN(l2sq)(nX, dx)), xl2 = sqrt(FN(l2sq)(nX, x));
    if (dxl2 <= eps2 * (xl2 + eps1)) { /* :176-179 */
      if (trace) {
        orc_lm_trace tb = {Fnew, -1, pcg_it, region, dxl2, 0};
        trace[nt++] = tb;
      }
      break;
    }
    for (int64_t i = 0; i < nX; ++i) xt[i] = x[i] + dx[i]; /* :180 */
    const double rho_den = FN(orc_jdx_plus_r_sq)(G, J, r, dx) - Fnew; /* :181-182 */
    F = Fnew;
    FN(orc_forward)(nE, G->cam_idx, G->pt_idx, obs, xt, xt + 9 * nC, r2, J2); /* :184 */
    Fnew = FN(orc_cost)(nE, r2);
    const double rho = -(F - Fnew) / rho_den; /* :186 */
    int accepted;
    if (F > Fnew) { /* :187 */
      R *sw;
      sw = r; r = r2; r2 = sw;
      sw = J; J = J2; J2 = sw;
      FN(orc_build)(G, r, J, NULL, Hpp, Hll, W, g); /* :189 */
      memcpy(dxbk, dx, sizeof(R) * nX);             /* :192 */
      memcpy(gbk, g, sizeof(R) * nX);
      for (int64_t i = 0; i < nX; ++i) x[i] += dx[i]; /* :194 */
      const double m = 1 - pow(2 * rho - 1, 3);
      region /= (1. / 3. > m ? 1. / 3. : m); /* :197-198 */
      v = 2.;
      recover = 0;
      stop = FN(linf)(nX, g) <= eps1; /* :203-205 */
      accepted = 1;
    } else {
      memcpy(dx, dxbk, sizeof(R) * nX); /* :208 */
      memcpy(g, gbk, sizeof(R) * nX);
      Fnew = F;     /* :210 */
      region /= v;  /* :211 */
      v *= 2;
      recover = 1;
      accepted = 0;
    }
    if (trace) {
      orc_lm_trace ti = {Fnew, accepted, pcg_it, region, dxl2, rho};
      trace[nt++] = ti;
    }
  }
  if (n_trace) *n_trace = nt;
  free(r); free(J); free(r2); free(J2); free(W); free(Hpp); free(Hll);
  free(dgc); free(dgp); free(g); free(gbk); free(dx); free(dxbk); free(xt);
  return k;
}

#undef FN
#undef CAT
#undef CAT2
