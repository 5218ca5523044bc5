// L-BFGS-B for a handful of variables, in reverse communication, usable on the host and inside a kernel.
//
// The reference optimises the GP hyperparameters of every latent with scipy.optimize.minimize(method="L-BFGS-B")
// (vlgp/gp.py:100-123): three variables log(sigma^2, omega, eps), all boxed, gradient masked to [0, 1, 0], scipy's
// defaults (m = 10, factr = ftol / eps = 1e7, pgtol = 1e-5, maxls = 20).  scipy's routine is a translation of
// L-BFGS-B 3.0 (Byrd, Lu, Nocedal, Zhu; Morales, Nocedal 2011) and is not part of /root/reference (scipy is unpinned
// there, requirements.txt); this file restates the published algorithm -- projected gradient, generalised Cauchy
// point, subspace minimisation with the 2011 projection step, More'-Thuente line search (MINPACK-2 dcsrch / dcstep),
// compact limited-memory BFGS matrices -- with the same control flow, so that the H-step round loop can run without
// the Python interpreter (host loop in capi.cu) or without the host at all (device loop in hstep_opt.cu).
// Parity is anchored on the reference's call site: tests/test_lbfgsb_port.py drives this code and scipy's setulb on
// the same objectives (smooth test functions in 1-4 variables with active bounds, and the H-step's own inconsistent
// objective / gradient pair) and compares the iterates evaluation by evaluation.
//
// The objective of the H-step is NOT differentiable consistently with its gradient (the reference's gradient treats
// the posterior covariances as constants, vlgp/gp.py:36-41), so its line searches typically collapse onto their
// starting point: a dozen evaluations within 1e-12 of each other until maxls is hit, then the failure path (restart
// from steepest descent, or ABNORMAL termination), or until the relative-width test of dcsrch accepts one of them.
// `collapse_tol` > 0 cuts such a line search short as soon as its whole bracket lies within collapse_tol of the
// starting point in x (every point it could still return does, too): with the best step still 0 the search is declared
// failed at once and the failure path is taken; with a best step > 0 that step is accepted the way dcsrch's own
// "XTOL TEST SATISFIED" exit accepts it (the relative test made absolute).  0 = follow the full sequence.
#pragma once

#ifdef __CUDACC__
#define LB_HD __host__ __device__
#else
#define LB_HD
#endif

#include <math.h>

namespace lbfgsb {

constexpr int NMAX = 4;          // variables (the H-step has 3)
constexpr int M = 10;            // limited-memory pairs (scipy's default maxcor)
constexpr int M2 = 2 * M;

enum Task : int {
    T_START = 0,
    T_FG = 1,                    // evaluate f, g at x and call again
    T_CONV_PGTOL = 2,            // CONVERGENCE: NORM OF PROJECTED GRADIENT <= PGTOL
    T_CONV_FACTR = 3,            // CONVERGENCE: RELATIVE REDUCTION OF F <= FACTR*EPSMCH
    T_ABNORMAL = 4,              // ABNORMAL TERMINATION IN LNSRCH
    T_STOP_MAXITER = 5,
    T_STOP_MAXFUN = 6,
    T_ERROR = 7,
};

struct State {
    // problem
    int n;
    double x[NMAX], l[NMAX], u[NMAX];
    int nbd[NMAX];
    double f, g[NMAX];
    double factr, pgtol, collapse_tol;
    int maxls, maxiter, maxfun;
    // limited-memory matrices (column-major, Fortran shapes ws(n,m), sy(m,m), wn(2m,2m) ...)
    double ws[NMAX * M], wy[NMAX * M], sy[M * M], ss[M * M], wt[M * M], wn[M2 * M2], snd[M2 * M2];
    double z[NMAX], r[NMAX], d[NMAX], t[NMAX], xp[NMAX], wa[8 * M];
    int index[NMAX], iwhere[NMAX], indx2[NMAX];
    // scalars of mainlb
    int prjctd, cnstnd, boxed, updatd;
    int nfree, nact, ileave, nenter, col, head, itail, iter, iupdat, iback, ifun, nfgv, info, nseg, nskip, iword, wrk;
    double theta, fold, tol, dnorm, epsmch, gd, stpmx, sbgnrm, stp, gdold, dtd;
    // dcsrch
    int ls_started, brackt, stage;
    double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
    // driver
    int phase, task, nit, n_collapsed;
};

#define LB_WS(i, j) s.ws[((j) - 1) * NMAX + (i) - 1]
#define LB_WY(i, j) s.wy[((j) - 1) * NMAX + (i) - 1]
#define LB_SY(i, j) s.sy[((j) - 1) * M + (i) - 1]
#define LB_SS(i, j) s.ss[((j) - 1) * M + (i) - 1]
#define LB_WT(i, j) s.wt[((j) - 1) * M + (i) - 1]
#define LB_WN(i, j) s.wn[((j) - 1) * M2 + (i) - 1]
#define LB_WN1(i, j) s.snd[((j) - 1) * M2 + (i) - 1]

// ---- LINPACK dpofa / dtrsl on a column-major matrix with leading dimension ld (1-based accessors) -------------------
LB_HD inline int dpofa(double *a, int ld, int n) {
#define A_(i, j) a[((j) - 1) * ld + (i) - 1]
    for (int j = 1; j <= n; ++j) {
        double sacc = 0.0;
        for (int k = 1; k <= j - 1; ++k) {
            double dot = 0.0;
            for (int q = 1; q <= k - 1; ++q) dot += A_(q, k) * A_(q, j);
            double tt = A_(k, j) - dot;
            tt = tt / A_(k, k);
            A_(k, j) = tt;
            sacc += tt * tt;
        }
        sacc = A_(j, j) - sacc;
        if (sacc <= 0.0) return j;
        A_(j, j) = sqrt(sacc);
    }
    return 0;
#undef A_
}

// job: 00 T x = b lower, 01 T x = b upper, 10 T' x = b lower, 11 T' x = b upper
LB_HD inline int dtrsl(const double *tm, int ld, int n, double *b, int job) {
#define T_(i, j) tm[((j) - 1) * ld + (i) - 1]
    for (int i = 1; i <= n; ++i)
        if (T_(i, i) == 0.0) return i;
    int kase = (job % 10 == 0) ? 1 : 2;
    if ((job % 100) / 10 != 0) kase += 2;
    if (kase == 1) {
        b[0] = b[0] / T_(1, 1);
        for (int j = 2; j <= n; ++j) {
            const double temp = -b[j - 2];
            for (int q = 0; q < n - j + 1; ++q) b[j - 1 + q] += temp * T_(j + q, j - 1);
            b[j - 1] = b[j - 1] / T_(j, j);
        }
    } else if (kase == 2) {
        b[n - 1] = b[n - 1] / T_(n, n);
        for (int jj = 2; jj <= n; ++jj) {
            const int j = n - jj + 1;
            const double temp = -b[j];
            for (int q = 0; q < j; ++q) b[q] += temp * T_(1 + q, j + 1);
            b[j - 1] = b[j - 1] / T_(j, j);
        }
    } else if (kase == 3) {
        b[n - 1] = b[n - 1] / T_(n, n);
        for (int jj = 2; jj <= n; ++jj) {
            const int j = n - jj + 1;
            double dot = 0.0;
            for (int q = 0; q < jj - 1; ++q) dot += T_(j + 1 + q, j) * b[j + q];
            b[j - 1] = b[j - 1] - dot;
            b[j - 1] = b[j - 1] / T_(j, j);
        }
    } else {
        b[0] = b[0] / T_(1, 1);
        for (int j = 2; j <= n; ++j) {
            double dot = 0.0;
            for (int q = 0; q < j - 1; ++q) dot += T_(1 + q, j) * b[q];
            b[j - 1] = b[j - 1] - dot;
            b[j - 1] = b[j - 1] / T_(j, j);
        }
    }
    return 0;
#undef T_
}

// ---- projected-gradient norm ----------------------------------------------------------------------------------------
LB_HD inline double projgr(const State &s) {
    double nrm = 0.0;
    for (int i = 0; i < s.n; ++i) {
        double gi = s.g[i];
        if (s.nbd[i] != 0) {
            if (gi < 0.0) {
                if (s.nbd[i] >= 2) gi = fmax(s.x[i] - s.u[i], gi);
            } else {
                if (s.nbd[i] <= 2) gi = fmin(s.x[i] - s.l[i], gi);
            }
        }
        nrm = fmax(nrm, fabs(gi));
    }
    return nrm;
}

// ---- product of the 2m x 2m middle matrix of the compact L-BFGS formula with v ---------------------------------------
LB_HD inline int bmv(const State &s, const double *v, double *p) {
    const int col = s.col;
    if (col == 0) return 0;
    p[col] = v[col];
    for (int i = 2; i <= col; ++i) {
        const int i2 = col + i;
        double sum = 0.0;
        for (int k = 1; k <= i - 1; ++k) sum += LB_SY(i, k) * v[k - 1] / LB_SY(k, k);
        p[i2 - 1] = v[i2 - 1] + sum;
    }
    int info = dtrsl(s.wt, M, col, p + col, 11);
    if (info != 0) return info;
    for (int i = 1; i <= col; ++i) p[i - 1] = v[i - 1] / sqrt(LB_SY(i, i));
    info = dtrsl(s.wt, M, col, p + col, 1);
    if (info != 0) return info;
    for (int i = 1; i <= col; ++i) p[i - 1] = -p[i - 1] / sqrt(LB_SY(i, i));
    for (int i = 1; i <= col; ++i) {
        double sum = 0.0;
        for (int k = i + 1; k <= col; ++k) sum += LB_SY(k, i) * p[col + k - 1] / LB_SY(i, i);
        p[i - 1] += sum;
    }
    return 0;
}

// ---- heap of breakpoints ----------------------------------------------------------------------------------------------
LB_HD inline void hpsolb(int n, double *t, int *iorder, int iheap) {
    if (iheap == 0) {
        for (int k = 2; k <= n; ++k) {
            const double ddum = t[k - 1];
            const int indxin = iorder[k - 1];
            int i = k;
            while (i > 1) {
                const int j = i / 2;
                if (ddum < t[j - 1]) {
                    t[i - 1] = t[j - 1];
                    iorder[i - 1] = iorder[j - 1];
                    i = j;
                } else {
                    break;
                }
            }
            t[i - 1] = ddum;
            iorder[i - 1] = indxin;
        }
    }
    if (n > 1) {
        int i = 1;
        const double out = t[0];
        const int indxou = iorder[0];
        const double ddum = t[n - 1];
        const int indxin = iorder[n - 1];
        for (;;) {
            int j = i + i;
            if (j <= n - 1) {
                if (t[j] < t[j - 1]) j = j + 1;
                if (t[j - 1] < ddum) {
                    t[i - 1] = t[j - 1];
                    iorder[i - 1] = iorder[j - 1];
                    i = j;
                    continue;
                }
            }
            break;
        }
        t[i - 1] = ddum;
        iorder[i - 1] = indxin;
        t[n - 1] = out;
        iorder[n - 1] = indxou;
    }
}

// ---- generalised Cauchy point -----------------------------------------------------------------------------------------
// iorder = s.indx2, breakpoints t = s.t, xcp = s.z; p = wa[0..2m), c = wa[2m..4m), wbp = wa[4m..6m), v = wa[6m..8m)
LB_HD inline int cauchy(State &s) {
    const int n = s.n, col = s.col, col2 = 2 * s.col;
    double *p = s.wa, *c = s.wa + 2 * M, *wbp = s.wa + 4 * M, *v = s.wa + 6 * M;
    double *xcp = s.z, *tt = s.t, *d = s.d;
    int *iorder = s.indx2;
    const double theta = s.theta;
    if (s.sbgnrm <= 0.0) {
        for (int i = 0; i < n; ++i) xcp[i] = s.x[i];
        return 0;
    }
    bool bnded = true;
    int nfree = n + 1, nbreak = 0, ibkmin = 0;
    double bkmin = 0.0, f1 = 0.0;
    for (int i = 0; i < col2; ++i) p[i] = 0.0;
    for (int i = 1; i <= n; ++i) {
        const double neggi = -s.g[i - 1];
        double tl = 0.0, tu = 0.0;
        if (s.iwhere[i - 1] != 3 && s.iwhere[i - 1] != -1) {
            if (s.nbd[i - 1] <= 2) tl = s.x[i - 1] - s.l[i - 1];
            if (s.nbd[i - 1] >= 2) tu = s.u[i - 1] - s.x[i - 1];
            const bool xlower = s.nbd[i - 1] <= 2 && tl <= 0.0;
            const bool xupper = s.nbd[i - 1] >= 2 && tu <= 0.0;
            s.iwhere[i - 1] = 0;
            if (xlower) {
                if (neggi <= 0.0) s.iwhere[i - 1] = 1;
            } else if (xupper) {
                if (neggi >= 0.0) s.iwhere[i - 1] = 2;
            } else {
                if (fabs(neggi) <= 0.0) s.iwhere[i - 1] = -3;
            }
        }
        int pointr = s.head;
        if (s.iwhere[i - 1] != 0 && s.iwhere[i - 1] != -1) {
            d[i - 1] = 0.0;
        } else {
            d[i - 1] = neggi;
            f1 -= neggi * neggi;
            for (int j = 1; j <= col; ++j) {
                p[j - 1] += LB_WY(i, pointr) * neggi;
                p[col + j - 1] += LB_WS(i, pointr) * neggi;
                pointr = pointr % M + 1;
            }
            if (s.nbd[i - 1] <= 2 && s.nbd[i - 1] != 0 && neggi < 0.0) {
                nbreak++;
                iorder[nbreak - 1] = i;
                tt[nbreak - 1] = tl / (-neggi);
                if (nbreak == 1 || tt[nbreak - 1] < bkmin) {
                    bkmin = tt[nbreak - 1];
                    ibkmin = nbreak;
                }
            } else if (s.nbd[i - 1] >= 2 && neggi > 0.0) {
                nbreak++;
                iorder[nbreak - 1] = i;
                tt[nbreak - 1] = tu / neggi;
                if (nbreak == 1 || tt[nbreak - 1] < bkmin) {
                    bkmin = tt[nbreak - 1];
                    ibkmin = nbreak;
                }
            } else {
                nfree--;
                iorder[nfree - 1] = i;
                if (fabs(neggi) > 0.0) bnded = false;
            }
        }
    }
    if (theta != 1.0)
        for (int j = 0; j < col; ++j) p[col + j] *= theta;
    for (int i = 0; i < n; ++i) xcp[i] = s.x[i];
    if (nbreak == 0 && nfree == n + 1) return 0;
    for (int j = 0; j < col2; ++j) c[j] = 0.0;
    double f2 = -theta * f1;
    const double f2_org = f2;
    if (col > 0) {
        const int info = bmv(s, p, v);
        if (info != 0) return info;
        double dot = 0.0;
        for (int j = 0; j < col2; ++j) dot += v[j] * p[j];
        f2 -= dot;
    }
    double dtm = -f1 / f2;
    double tsum = 0.0;
    s.nseg = 1;
    bool all_fixed = false;
    if (nbreak != 0) {
        int nleft = nbreak, iter = 1;
        double tj = 0.0;
        for (;;) {
            const double tj0 = tj;
            int ibp;
            if (iter == 1) {
                tj = bkmin;
                ibp = iorder[ibkmin - 1];
            } else {
                if (iter == 2) {
                    if (ibkmin != nbreak) {
                        tt[ibkmin - 1] = tt[nbreak - 1];
                        iorder[ibkmin - 1] = iorder[nbreak - 1];
                    }
                }
                hpsolb(nleft, tt, iorder, iter - 2);
                tj = tt[nleft - 1];
                ibp = iorder[nleft - 1];
            }
            const double dt = tj - tj0;
            if (dtm < dt) break;
            tsum += dt;
            nleft--;
            iter++;
            const double dibp = d[ibp - 1];
            d[ibp - 1] = 0.0;
            double zibp;
            if (dibp > 0.0) {
                zibp = s.u[ibp - 1] - s.x[ibp - 1];
                xcp[ibp - 1] = s.u[ibp - 1];
                s.iwhere[ibp - 1] = 2;
            } else {
                zibp = s.l[ibp - 1] - s.x[ibp - 1];
                xcp[ibp - 1] = s.l[ibp - 1];
                s.iwhere[ibp - 1] = 1;
            }
            if (nleft == 0 && nbreak == n) {
                dtm = dt;
                all_fixed = true;
                break;
            }
            s.nseg++;
            const double dibp2 = dibp * dibp;
            f1 = f1 + dt * f2 + dibp2 - theta * dibp * zibp;
            f2 = f2 - theta * dibp2;
            if (col > 0) {
                for (int j = 0; j < col2; ++j) c[j] += dt * p[j];
                int pointr = s.head;
                for (int j = 1; j <= col; ++j) {
                    wbp[j - 1] = LB_WY(ibp, pointr);
                    wbp[col + j - 1] = theta * LB_WS(ibp, pointr);
                    pointr = pointr % M + 1;
                }
                const int info = bmv(s, wbp, v);
                if (info != 0) return info;
                double wmc = 0.0, wmp = 0.0, wmw = 0.0;
                for (int j = 0; j < col2; ++j) wmc += c[j] * v[j];
                for (int j = 0; j < col2; ++j) wmp += p[j] * v[j];
                for (int j = 0; j < col2; ++j) wmw += wbp[j] * v[j];
                for (int j = 0; j < col2; ++j) p[j] += -dibp * wbp[j];
                f1 += dibp * wmc;
                f2 += 2.0 * dibp * wmp - dibp2 * wmw;
            }
            f2 = fmax(s.epsmch * f2_org, f2);
            if (nleft > 0) {
                dtm = -f1 / f2;
                continue;
            } else if (bnded) {
                f1 = 0.0;
                f2 = 0.0;
                dtm = 0.0;
            } else {
                dtm = -f1 / f2;
            }
            break;
        }
    }
    if (!all_fixed) {
        if (dtm <= 0.0) dtm = 0.0;
        tsum += dtm;
        for (int i = 0; i < n; ++i) xcp[i] += tsum * d[i];
    }
    if (col > 0)
        for (int j = 0; j < col2; ++j) c[j] += dtm * p[j];
    return 0;
}

// ---- free / active variable bookkeeping ------------------------------------------------------------------------------
LB_HD inline void freev(State &s) {
    const int n = s.n;
    s.nenter = 0;
    s.ileave = n + 1;
    if (s.iter > 0 && s.cnstnd) {
        for (int i = 1; i <= s.nfree; ++i) {
            const int k = s.index[i - 1];
            if (s.iwhere[k - 1] > 0) {
                s.ileave--;
                s.indx2[s.ileave - 1] = k;
            }
        }
        for (int i = 1 + s.nfree; i <= n; ++i) {
            const int k = s.index[i - 1];
            if (s.iwhere[k - 1] <= 0) {
                s.nenter++;
                s.indx2[s.nenter - 1] = k;
            }
        }
    }
    s.wrk = (s.ileave < n + 1) || (s.nenter > 0) || s.updatd;
    s.nfree = 0;
    int iact = n + 1;
    for (int i = 1; i <= n; ++i) {
        if (s.iwhere[i - 1] <= 0) {
            s.nfree++;
            s.index[s.nfree - 1] = i;
        } else {
            iact--;
            s.index[iact - 1] = i;
        }
    }
}

// ---- LEL' factorisation of the indefinite K matrix of the subspace problem -------------------------------------------
LB_HD inline int formk(State &s) {
    const int n = s.n, col = s.col, nsub = s.nfree;
    const int *ind = s.index, *indx2 = s.indx2;
    int upcl;
    if (s.updatd) {
        if (s.iupdat > M) {
            for (int jy = 1; jy <= M - 1; ++jy) {
                const int js = M + jy;
                for (int q = 0; q < M - jy; ++q) LB_WN1(jy + q, jy) = LB_WN1(jy + 1 + q, jy + 1);
                for (int q = 0; q < M - jy; ++q) LB_WN1(js + q, js) = LB_WN1(js + 1 + q, js + 1);
                for (int q = 0; q < M - 1; ++q) LB_WN1(M + 1 + q, jy) = LB_WN1(M + 2 + q, jy + 1);
            }
        }
        const int pbegin = 1, pend = nsub, dbegin = nsub + 1, dend = n;
        const int iy = col, is = M + col;
        int ipntr = s.head + col - 1;
        if (ipntr > M) ipntr -= M;
        int jpntr = s.head;
        for (int jy = 1; jy <= col; ++jy) {
            const int js = M + jy;
            double temp1 = 0.0, temp2 = 0.0, temp3 = 0.0;
            for (int k = pbegin; k <= pend; ++k) {
                const int k1 = ind[k - 1];
                temp1 += LB_WY(k1, ipntr) * LB_WY(k1, jpntr);
            }
            for (int k = dbegin; k <= dend; ++k) {
                const int k1 = ind[k - 1];
                temp2 += LB_WS(k1, ipntr) * LB_WS(k1, jpntr);
                temp3 += LB_WS(k1, ipntr) * LB_WY(k1, jpntr);
            }
            LB_WN1(iy, jy) = temp1;
            LB_WN1(is, js) = temp2;
            LB_WN1(is, jy) = temp3;
            jpntr = jpntr % M + 1;
        }
        const int jy = col;
        jpntr = s.head + col - 1;
        if (jpntr > M) jpntr -= M;
        ipntr = s.head;
        for (int i = 1; i <= col; ++i) {
            const int is2 = M + i;
            double temp3 = 0.0;
            for (int k = pbegin; k <= pend; ++k) {
                const int k1 = ind[k - 1];
                temp3 += LB_WS(k1, ipntr) * LB_WY(k1, jpntr);
            }
            ipntr = ipntr % M + 1;
            LB_WN1(is2, jy) = temp3;
        }
        upcl = col - 1;
    } else {
        upcl = col;
    }
    int ipntr = s.head;
    for (int iy = 1; iy <= upcl; ++iy) {
        const int is = M + iy;
        int jpntr = s.head;
        for (int jy = 1; jy <= iy; ++jy) {
            const int js = M + jy;
            double temp1 = 0.0, temp2 = 0.0, temp3 = 0.0, temp4 = 0.0;
            for (int k = 1; k <= s.nenter; ++k) {
                const int k1 = indx2[k - 1];
                temp1 += LB_WY(k1, ipntr) * LB_WY(k1, jpntr);
                temp2 += LB_WS(k1, ipntr) * LB_WS(k1, jpntr);
            }
            for (int k = s.ileave; k <= n; ++k) {
                const int k1 = indx2[k - 1];
                temp3 += LB_WY(k1, ipntr) * LB_WY(k1, jpntr);
                temp4 += LB_WS(k1, ipntr) * LB_WS(k1, jpntr);
            }
            LB_WN1(iy, jy) = LB_WN1(iy, jy) + temp1 - temp3;
            LB_WN1(is, js) = LB_WN1(is, js) - temp2 + temp4;
            jpntr = jpntr % M + 1;
        }
        ipntr = ipntr % M + 1;
    }
    ipntr = s.head;
    for (int is = M + 1; is <= M + upcl; ++is) {
        int jpntr = s.head;
        for (int jy = 1; jy <= upcl; ++jy) {
            double temp1 = 0.0, temp3 = 0.0;
            for (int k = 1; k <= s.nenter; ++k) {
                const int k1 = indx2[k - 1];
                temp1 += LB_WS(k1, ipntr) * LB_WY(k1, jpntr);
            }
            for (int k = s.ileave; k <= n; ++k) {
                const int k1 = indx2[k - 1];
                temp3 += LB_WS(k1, ipntr) * LB_WY(k1, jpntr);
            }
            if (is <= jy + M)
                LB_WN1(is, jy) = LB_WN1(is, jy) + temp1 - temp3;
            else
                LB_WN1(is, jy) = LB_WN1(is, jy) - temp1 + temp3;
            jpntr = jpntr % M + 1;
        }
        ipntr = ipntr % M + 1;
    }
    const double theta = s.theta;
    for (int iy = 1; iy <= col; ++iy) {
        const int is = col + iy, is1 = M + iy;
        for (int jy = 1; jy <= iy; ++jy) {
            const int js = col + jy, js1 = M + jy;
            LB_WN(jy, iy) = LB_WN1(iy, jy) / theta;
            LB_WN(js, is) = LB_WN1(is1, js1) * theta;
        }
        for (int jy = 1; jy <= iy - 1; ++jy) LB_WN(jy, is) = -LB_WN1(is1, jy);
        for (int jy = iy; jy <= col; ++jy) LB_WN(jy, is) = LB_WN1(is1, jy);
        LB_WN(iy, iy) = LB_WN(iy, iy) + LB_SY(iy, iy);
    }
    int info = dpofa(s.wn, M2, col);
    if (info != 0) return -1;
    const int col2 = 2 * col;
    for (int js = col + 1; js <= col2; ++js) {
        info = dtrsl(s.wn, M2, col, &LB_WN(1, js), 11);
        (void)info;
    }
    for (int is = col + 1; is <= col2; ++is)
        for (int js = is; js <= col2; ++js) {
            double dot = 0.0;
            for (int q = 1; q <= col; ++q) dot += LB_WN(q, is) * LB_WN(q, js);
            LB_WN(is, js) = LB_WN(is, js) + dot;
        }
    info = dpofa(&LB_WN(col + 1, col + 1), M2, col);
    if (info != 0) return -2;
    return 0;
}

// ---- reduced gradient at the Cauchy point -------------------------------------------------------------------------------
LB_HD inline int cmprlb(State &s) {
    const int col = s.col;
    if (!s.cnstnd && col > 0) {
        for (int i = 0; i < s.n; ++i) s.r[i] = -s.g[i];
        return 0;
    }
    for (int i = 1; i <= s.nfree; ++i) {
        const int k = s.index[i - 1];
        s.r[i - 1] = -s.theta * (s.z[k - 1] - s.x[k - 1]) - s.g[k - 1];
    }
    const int info = bmv(s, s.wa + 2 * M, s.wa);
    if (info != 0) return -8;
    int pointr = s.head;
    for (int j = 1; j <= col; ++j) {
        const double a1 = s.wa[j - 1], a2 = s.theta * s.wa[col + j - 1];
        for (int i = 1; i <= s.nfree; ++i) {
            const int k = s.index[i - 1];
            s.r[i - 1] += LB_WY(k, pointr) * a1 + LB_WS(k, pointr) * a2;
        }
        pointr = pointr % M + 1;
    }
    return 0;
}

// ---- subspace minimisation (with the projection step of L-BFGS-B 3.0) -------------------------------------------------
LB_HD inline int subsm(State &s) {
    const int n = s.n, nsub = s.nfree, col = s.col;
    const int *ind = s.index;
    double *x = s.z, *d = s.r, *xp = s.xp, *wv = s.wa;
    const double *xx = s.x, *gg = s.g;
    const double theta = s.theta;
    if (nsub <= 0) return 0;
    int pointr = s.head;
    for (int i = 1; i <= col; ++i) {
        double temp1 = 0.0, temp2 = 0.0;
        for (int j = 1; j <= nsub; ++j) {
            const int k = ind[j - 1];
            temp1 += LB_WY(k, pointr) * d[j - 1];
            temp2 += LB_WS(k, pointr) * d[j - 1];
        }
        wv[i - 1] = temp1;
        wv[col + i - 1] = theta * temp2;
        pointr = pointr % M + 1;
    }
    const int col2 = 2 * col;
    int info = dtrsl(s.wn, M2, col2, wv, 11);
    if (info != 0) return info;
    for (int i = 0; i < col; ++i) wv[i] = -wv[i];
    info = dtrsl(s.wn, M2, col2, wv, 1);
    if (info != 0) return info;
    pointr = s.head;
    for (int jy = 1; jy <= col; ++jy) {
        const int js = col + jy;
        for (int i = 1; i <= nsub; ++i) {
            const int k = ind[i - 1];
            d[i - 1] = d[i - 1] + LB_WY(k, pointr) * wv[jy - 1] / theta + LB_WS(k, pointr) * wv[js - 1];
        }
        pointr = pointr % M + 1;
    }
    for (int i = 0; i < nsub; ++i) d[i] *= 1.0 / theta;
    s.iword = 0;
    for (int i = 0; i < n; ++i) xp[i] = x[i];
    for (int i = 1; i <= nsub; ++i) {
        const int k = ind[i - 1];
        const double dk = d[i - 1];
        double xk = x[k - 1];
        if (s.nbd[k - 1] != 0) {
            if (s.nbd[k - 1] == 1) {
                x[k - 1] = fmax(s.l[k - 1], xk + dk);
                if (x[k - 1] == s.l[k - 1]) s.iword = 1;
            } else if (s.nbd[k - 1] == 2) {
                xk = fmax(s.l[k - 1], xk + dk);
                x[k - 1] = fmin(s.u[k - 1], xk);
                if (x[k - 1] == s.l[k - 1] || x[k - 1] == s.u[k - 1]) s.iword = 1;
            } else if (s.nbd[k - 1] == 3) {
                x[k - 1] = fmin(s.u[k - 1], xk + dk);
                if (x[k - 1] == s.u[k - 1]) s.iword = 1;
            }
        } else {
            x[k - 1] = xk + dk;
        }
    }
    if (s.iword == 0) return 0;
    double dd_p = 0.0;
    for (int i = 0; i < n; ++i) dd_p += (x[i] - xx[i]) * gg[i];
    if (dd_p > 0.0) {
        for (int i = 0; i < n; ++i) x[i] = xp[i];
        double alpha = 1.0, temp1 = alpha;
        int ibd = 0;
        for (int i = 1; i <= nsub; ++i) {
            const int k = ind[i - 1];
            const double dk = d[i - 1];
            if (s.nbd[k - 1] != 0) {
                if (dk < 0.0 && s.nbd[k - 1] <= 2) {
                    const double temp2 = s.l[k - 1] - x[k - 1];
                    if (temp2 >= 0.0)
                        temp1 = 0.0;
                    else if (dk * alpha < temp2)
                        temp1 = temp2 / dk;
                } else if (dk > 0.0 && s.nbd[k - 1] >= 2) {
                    const double temp2 = s.u[k - 1] - x[k - 1];
                    if (temp2 <= 0.0)
                        temp1 = 0.0;
                    else if (dk * alpha > temp2)
                        temp1 = temp2 / dk;
                }
                if (temp1 < alpha) {
                    alpha = temp1;
                    ibd = i;
                }
            }
        }
        if (alpha < 1.0) {
            const double dk = d[ibd - 1];
            const int k = ind[ibd - 1];
            if (dk > 0.0) {
                x[k - 1] = s.u[k - 1];
                d[ibd - 1] = 0.0;
            } else if (dk < 0.0) {
                x[k - 1] = s.l[k - 1];
                d[ibd - 1] = 0.0;
            }
        }
        for (int i = 1; i <= nsub; ++i) {
            const int k = ind[i - 1];
            x[k - 1] = x[k - 1] + alpha * d[i - 1];
        }
    }
    return 0;
}

// ---- limited-memory matrix updates ---------------------------------------------------------------------------------------
LB_HD inline void matupd(State &s, double rr, double dr) {
    const int n = s.n;
    if (s.iupdat <= M) {
        s.col = s.iupdat;
        s.itail = (s.head + s.iupdat - 2) % M + 1;
    } else {
        s.itail = s.itail % M + 1;
        s.head = s.head % M + 1;
    }
    for (int i = 1; i <= n; ++i) {
        LB_WS(i, s.itail) = s.d[i - 1];
        LB_WY(i, s.itail) = s.r[i - 1];
    }
    s.theta = rr / dr;
    const int col = s.col;
    if (s.iupdat > M) {
        for (int j = 1; j <= col - 1; ++j) {
            for (int q = 0; q < j; ++q) LB_SS(1 + q, j) = LB_SS(2 + q, j + 1);
            for (int q = 0; q < col - j; ++q) LB_SY(j + q, j) = LB_SY(j + 1 + q, j + 1);
        }
    }
    int pointr = s.head;
    for (int j = 1; j <= col - 1; ++j) {
        double a = 0.0, b = 0.0;
        for (int i = 1; i <= n; ++i) a += s.d[i - 1] * LB_WY(i, pointr);
        for (int i = 1; i <= n; ++i) b += LB_WS(i, pointr) * s.d[i - 1];
        LB_SY(col, j) = a;
        LB_SS(j, col) = b;
        pointr = pointr % M + 1;
    }
    if (s.stp == 1.0)
        LB_SS(col, col) = s.dtd;
    else
        LB_SS(col, col) = s.stp * s.stp * s.dtd;
    LB_SY(col, col) = dr;
}

LB_HD inline int formt(State &s) {
    const int col = s.col;
    for (int j = 1; j <= col; ++j) LB_WT(1, j) = s.theta * LB_SS(1, j);
    for (int i = 2; i <= col; ++i)
        for (int j = i; j <= col; ++j) {
            const int k1 = (i < j ? i : j) - 1;
            double ddum = 0.0;
            for (int k = 1; k <= k1; ++k) ddum += LB_SY(i, k) * LB_SY(j, k) / LB_SY(k, k);
            LB_WT(i, j) = ddum + s.theta * LB_SS(i, j);
        }
    const int info = dpofa(s.wt, M, col);
    return info != 0 ? -3 : 0;
}

// ---- More'-Thuente step (MINPACK-2 dcstep) ---------------------------------------------------------------------------------
LB_HD inline void dcstep(double &stx, double &fx, double &dx, double &sty, double &fy, double &dy, double &stp, double fp,
                         double dp, int &brackt, double stpmin, double stpmax) {
    double gamma, p, q, r, sq, stpc, stpf, stpq, theta;
    const double sgnd = dp * (dx / fabs(dx));
    if (fp > fx) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        sq = fmax(fmax(fabs(theta), fabs(dx)), fabs(dp));
        gamma = sq * sqrt((theta / sq) * (theta / sq) - (dx / sq) * (dp / sq));
        if (stp < stx) gamma = -gamma;
        p = (gamma - dx) + theta;
        q = ((gamma - dx) + gamma) + dp;
        r = p / q;
        stpc = stx + r * (stp - stx);
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
        if (fabs(stpc - stx) < fabs(stpq - stx))
            stpf = stpc;
        else
            stpf = stpc + (stpq - stpc) / 2.0;
        brackt = 1;
    } else if (sgnd < 0.0) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        sq = fmax(fmax(fabs(theta), fabs(dx)), fabs(dp));
        gamma = sq * sqrt((theta / sq) * (theta / sq) - (dx / sq) * (dp / sq));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = ((gamma - dp) + gamma) + dx;
        r = p / q;
        stpc = stp + r * (stx - stp);
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (fabs(stpc - stp) > fabs(stpq - stp))
            stpf = stpc;
        else
            stpf = stpq;
        brackt = 1;
    } else if (fabs(dp) < fabs(dx)) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        sq = fmax(fmax(fabs(theta), fabs(dx)), fabs(dp));
        gamma = sq * sqrt(fmax(0.0, (theta / sq) * (theta / sq) - (dx / sq) * (dp / sq)));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = (gamma + (dx - dp)) + gamma;
        r = p / q;
        if (r < 0.0 && gamma != 0.0)
            stpc = stp + r * (stx - stp);
        else if (stp > stx)
            stpc = stpmax;
        else
            stpc = stpmin;
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (brackt) {
            if (fabs(stpc - stp) < fabs(stpq - stp))
                stpf = stpc;
            else
                stpf = stpq;
            if (stp > stx)
                stpf = fmin(stp + 0.66 * (sty - stp), stpf);
            else
                stpf = fmax(stp + 0.66 * (sty - stp), stpf);
        } else {
            if (fabs(stpc - stp) > fabs(stpq - stp))
                stpf = stpc;
            else
                stpf = stpq;
            stpf = fmin(stpmax, stpf);
            stpf = fmax(stpmin, stpf);
        }
    } else {
        if (brackt) {
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
            sq = fmax(fmax(fabs(theta), fabs(dy)), fabs(dp));
            gamma = sq * sqrt((theta / sq) * (theta / sq) - (dy / sq) * (dp / sq));
            if (stp > sty) gamma = -gamma;
            p = (gamma - dp) + theta;
            q = ((gamma - dp) + gamma) + dy;
            r = p / q;
            stpc = stp + r * (sty - stp);
            stpf = stpc;
        } else if (stp > stx) {
            stpf = stpmax;
        } else {
            stpf = stpmin;
        }
    }
    if (fp > fx) {
        sty = stp;
        fy = fp;
        dy = dp;
    } else {
        if (sgnd < 0.0) {
            sty = stx;
            fy = fx;
            dy = dx;
        }
        stx = stp;
        fx = fp;
        dx = dp;
    }
    stp = stpf;
}

// returns 0: evaluate at the new stp (FG), 1: CONVERGENCE, 2: WARNING, 3: ERROR (stp unchanged)
LB_HD inline int dcsrch(State &s, double f, double g, double &stp, double ftol, double gtol, double xtol, double stpmin,
                        double stpmax) {
    const double xtrapl = 1.1, xtrapu = 4.0;
    if (!s.ls_started) {
        if (stp < stpmin) return 3;
        if (stp > stpmax) return 3;
        if (g >= 0.0) return 3;
        s.ls_started = 1;
        s.brackt = 0;
        s.stage = 1;
        s.finit = f;
        s.ginit = g;
        s.gtest = ftol * s.ginit;
        s.width = stpmax - stpmin;
        s.width1 = s.width / 0.5;
        s.stx = 0.0;
        s.fx = s.finit;
        s.gx = s.ginit;
        s.sty = 0.0;
        s.fy = s.finit;
        s.gy = s.ginit;
        s.stmin = 0.0;
        s.stmax = stp + xtrapu * stp;
        return 0;
    }
    const double ftest = s.finit + stp * s.gtest;
    if (s.stage == 1 && f <= ftest && g >= 0.0) s.stage = 2;
    int res = 0;
    if (s.brackt && (stp <= s.stmin || stp >= s.stmax)) res = 2;
    if (s.brackt && s.stmax - s.stmin <= xtol * s.stmax) res = 2;
    if (s.collapse_tol > 0.0 && s.brackt && s.stx > 0.0 && s.stmax * s.dnorm <= s.collapse_tol) res = 2;   // absolute xtol
    if (stp == stpmax && f <= ftest && g <= s.gtest) res = 2;
    if (stp == stpmin && (f > ftest || g >= s.gtest)) res = 2;
    if (f <= ftest && fabs(g) <= gtol * (-s.ginit)) res = 1;
    if (res != 0) return res;
    if (s.stage == 1 && f <= s.fx && f > ftest) {
        const double fm = f - stp * s.gtest;
        double fxm = s.fx - s.stx * s.gtest, fym = s.fy - s.sty * s.gtest;
        const double gm = g - s.gtest;
        double gxm = s.gx - s.gtest, gym = s.gy - s.gtest;
        dcstep(s.stx, fxm, gxm, s.sty, fym, gym, stp, fm, gm, s.brackt, s.stmin, s.stmax);
        s.fx = fxm + s.stx * s.gtest;
        s.fy = fym + s.sty * s.gtest;
        s.gx = gxm + s.gtest;
        s.gy = gym + s.gtest;
    } else {
        dcstep(s.stx, s.fx, s.gx, s.sty, s.fy, s.gy, stp, f, g, s.brackt, s.stmin, s.stmax);
    }
    if (s.brackt) {
        if (fabs(s.sty - s.stx) >= 0.66 * s.width1) stp = s.stx + 0.5 * (s.sty - s.stx);
        s.width1 = s.width;
        s.width = fabs(s.sty - s.stx);
    }
    if (s.brackt) {
        s.stmin = fmin(s.stx, s.sty);
        s.stmax = fmax(s.stx, s.sty);
    } else {
        s.stmin = stp + xtrapl * (stp - s.stx);
        s.stmax = stp + xtrapu * (stp - s.stx);
    }
    stp = fmax(stp, stpmin);
    stp = fmin(stp, stpmax);
    if ((s.brackt && (stp <= s.stmin || stp >= s.stmax)) || (s.brackt && s.stmax - s.stmin <= xtol * s.stmax)) stp = s.stx;
    if (s.collapse_tol > 0.0 && s.brackt && s.stx > 0.0 && s.stmax * s.dnorm <= s.collapse_tol) stp = s.stx;
    return 0;
}

// ---- driver ------------------------------------------------------------------------------------------------------------------
enum Phase : int { PH_START = 0, PH_FG_START = 1, PH_FG_LN = 2, PH_DONE = 3 };

LB_HD inline void init(State &s, int n, const double *x0, const double *lo, const double *up, double factr = 1e7,
                       double pgtol = 1e-5, int maxls = 20, int maxiter = 15000, int maxfun = 15000,
                       double collapse_tol = 0.0) {
    s.n = n;
    for (int i = 0; i < n; ++i) {
        s.l[i] = lo[i];
        s.u[i] = up[i];
        s.nbd[i] = 2;
        s.x[i] = fmin(fmax(x0[i], lo[i]), up[i]);          // scipy's driver clips x0 to the bounds before setulb
        s.g[i] = 0.0;
    }
    s.f = 0.0;
    s.factr = factr;
    s.pgtol = pgtol;
    s.maxls = maxls;
    s.maxiter = maxiter;
    s.maxfun = maxfun;
    s.collapse_tol = collapse_tol;
    s.phase = PH_START;
    s.task = T_START;
    s.nit = 0;
    s.n_collapsed = 0;
}

LB_HD inline void refresh_memory(State &s) {
    s.info = 0;
    s.col = 0;
    s.head = 1;
    s.theta = 1.0;
    s.iupdat = 0;
    s.updatd = 0;
}

// Advance until the optimiser needs f, g at s.x (returns 1; set s.f, s.g and call again) or has finished (returns 0;
// s.task says why, s.x / s.f hold the result).
LB_HD inline int advance(State &s) {
    const int n = s.n;
    if (s.phase == PH_DONE) return 0;
    if (s.phase == PH_START) {
        s.epsmch = 2.220446049250313e-16;
        s.col = 0;
        s.head = 1;
        s.theta = 1.0;
        s.iupdat = 0;
        s.updatd = 0;
        s.iback = 0;
        s.itail = 0;
        s.ifun = 0;
        s.iword = 0;
        s.nact = 0;
        s.ileave = 0;
        s.nenter = 0;
        s.tol = s.factr * s.epsmch;
        s.nfgv = 0;
        s.nskip = 0;
        s.nfree = n;
        s.iter = 0;
        s.info = 0;
        s.fold = 0.0;
        s.dnorm = 0.0;
        s.gd = 0.0;
        s.gdold = 0.0;
        s.stp = 0.0;
        s.stpmx = 0.0;
        s.sbgnrm = 0.0;
        s.dtd = 0.0;
        s.nseg = 0;
        s.wrk = 0;
        for (int i = 0; i < n; ++i)
            if (s.l[i] > s.u[i]) {
                s.task = T_ERROR;
                s.phase = PH_DONE;
                return 0;
            }
        // active(): project x onto the box, classify the variables
        s.prjctd = 0;
        s.cnstnd = 0;
        s.boxed = 1;
        for (int i = 0; i < n; ++i) {
            if (s.nbd[i] > 0) {
                if (s.nbd[i] <= 2 && s.x[i] <= s.l[i]) {
                    if (s.x[i] < s.l[i]) {
                        s.prjctd = 1;
                        s.x[i] = s.l[i];
                    }
                } else if (s.nbd[i] >= 2 && s.x[i] >= s.u[i]) {
                    if (s.x[i] > s.u[i]) {
                        s.prjctd = 1;
                        s.x[i] = s.u[i];
                    }
                }
            }
        }
        for (int i = 0; i < n; ++i) {
            if (s.nbd[i] != 2) s.boxed = 0;
            if (s.nbd[i] == 0) {
                s.iwhere[i] = -1;
            } else {
                s.cnstnd = 1;
                if (s.nbd[i] == 2 && s.u[i] - s.l[i] <= 0.0)
                    s.iwhere[i] = 3;
                else
                    s.iwhere[i] = 0;
            }
        }
        s.phase = PH_FG_START;
        s.task = T_FG;
        return 1;
    }
    if (s.phase == PH_FG_START) {
        s.nfgv = 1;
        s.sbgnrm = projgr(s);
        if (s.sbgnrm <= s.pgtol) {
            s.task = T_CONV_PGTOL;
            s.phase = PH_DONE;
            return 0;
        }
    }
    bool resume_ls = (s.phase == PH_FG_LN);
    for (;;) {                       // label 222 of mainlb
        if (!resume_ls) {
            s.iword = -1;
            bool skip_to_333 = false;
            if (!s.cnstnd && s.col > 0) {
                for (int i = 0; i < n; ++i) s.z[i] = s.x[i];
                s.wrk = s.updatd;
                s.nseg = 0;
                skip_to_333 = true;
            }
            if (!skip_to_333) {
                s.info = cauchy(s);
                if (s.info != 0) {
                    refresh_memory(s);
                    continue;
                }
                freev(s);
                s.nact = n - s.nfree;
            }
            if (s.nfree != 0 && s.col != 0) {
                if (s.wrk) s.info = formk(s);
                if (s.info != 0) {
                    refresh_memory(s);
                    continue;
                }
                s.info = cmprlb(s);
                if (s.info == 0) s.info = subsm(s);
                if (s.info != 0) {
                    refresh_memory(s);
                    continue;
                }
            }
            for (int i = 0; i < n; ++i) s.d[i] = s.z[i] - s.x[i];
            // ---- lnsrlb, first entry -------------------------------------------------------------------------
            double dtd = 0.0;
            for (int i = 0; i < n; ++i) dtd += s.d[i] * s.d[i];
            s.dtd = dtd;
            s.dnorm = sqrt(dtd);
            s.stpmx = 1e10;
            if (s.cnstnd) {
                if (s.iter == 0) {
                    s.stpmx = 1.0;
                } else {
                    for (int i = 0; i < n; ++i) {
                        const double a1 = s.d[i];
                        if (s.nbd[i] != 0) {
                            if (a1 < 0.0 && s.nbd[i] <= 2) {
                                const double a2 = s.l[i] - s.x[i];
                                if (a2 >= 0.0)
                                    s.stpmx = 0.0;
                                else if (a1 * s.stpmx < a2)
                                    s.stpmx = a2 / a1;
                            } else if (a1 > 0.0 && s.nbd[i] >= 2) {
                                const double a2 = s.u[i] - s.x[i];
                                if (a2 <= 0.0)
                                    s.stpmx = 0.0;
                                else if (a1 * s.stpmx > a2)
                                    s.stpmx = a2 / a1;
                            }
                        }
                    }
                }
            }
            if (s.iter == 0 && !s.boxed)
                s.stp = fmin(1.0 / s.dnorm, s.stpmx);
            else
                s.stp = 1.0;
            for (int i = 0; i < n; ++i) {
                s.t[i] = s.x[i];
                s.r[i] = s.g[i];
            }
            s.fold = s.f;
            s.ifun = 0;
            s.iback = 0;
            s.ls_started = 0;
        }
        resume_ls = false;
        // ---- label 556 of lnsrlb ---------------------------------------------------------------------------------
        bool ls_failed = false, need_fg = false;
        {
            double gd = 0.0;
            for (int i = 0; i < n; ++i) gd += s.g[i] * s.d[i];
            s.gd = gd;
            s.info = 0;
            if (s.ifun == 0) {
                s.gdold = gd;
                if (gd >= 0.0) s.info = -4;            // ascent direction in projection: line search impossible
            }
            if (s.info == 0) {
                const int res = dcsrch(s, s.f, s.gd, s.stp, 1e-3, 0.9, 0.1, 0.0, s.stpmx);
                if (res != 1 && res != 2) {
                    s.ifun++;
                    s.nfgv++;
                    s.iback = s.ifun - 1;
                    // a line search whose whole bracket [0, sty] lies within collapse_tol of its start cannot leave it
                    if (s.collapse_tol > 0.0 && s.brackt && s.stx == 0.0 && s.sty * s.dnorm <= s.collapse_tol &&
                        s.iback < s.maxls) {
                        s.n_collapsed++;
                        s.nfgv--;
                        s.ifun--;
                        s.iback = s.ifun - 1 < 0 ? 0 : s.ifun - 1;
                        s.info = -10;
                    } else {
                        if (s.stp == 1.0) {
                            for (int i = 0; i < n; ++i) s.x[i] = s.z[i];
                        } else {
                            for (int i = 0; i < n; ++i) s.x[i] = s.stp * s.d[i] + s.t[i];
                        }
                        need_fg = true;
                    }
                }
            }
            if (s.info != 0 || s.iback >= s.maxls) ls_failed = true;
        }
        if (ls_failed) {
            for (int i = 0; i < n; ++i) {
                s.x[i] = s.t[i];
                s.g[i] = s.r[i];
            }
            s.f = s.fold;
            if (s.col == 0) {
                if (s.info == 0) {
                    s.info = -9;
                    s.nfgv--;
                    s.ifun--;
                    s.iback--;
                }
                s.task = T_ABNORMAL;
                s.iter++;
                s.phase = PH_DONE;
                return 0;
            }
            if (s.info == 0) s.nfgv--;
            refresh_memory(s);
            continue;
        }
        if (need_fg) {
            s.phase = PH_FG_LN;
            s.task = T_FG;
            return 1;
        }
        // ---- new iterate --------------------------------------------------------------------------------------------
        s.iter++;
        s.sbgnrm = projgr(s);
        // scipy's driver between NEW_X and the next call: iteration / evaluation limits (_lbfgsb_py.py)
        s.nit++;
        if (s.nit >= s.maxiter) {
            s.task = T_STOP_MAXITER;
            s.phase = PH_DONE;
            return 0;
        }
        if (s.nfgv > s.maxfun) {
            s.task = T_STOP_MAXFUN;
            s.phase = PH_DONE;
            return 0;
        }
        if (s.sbgnrm <= s.pgtol) {
            s.task = T_CONV_PGTOL;
            s.phase = PH_DONE;
            return 0;
        }
        {
            const double ddum = fmax(fmax(fabs(s.fold), fabs(s.f)), 1.0);
            if ((s.fold - s.f) <= s.tol * ddum) {
                s.task = T_CONV_FACTR;
                if (s.iback >= 10) s.info = -5;
                s.phase = PH_DONE;
                return 0;
            }
        }
        double rr = 0.0, dr, ddum;
        for (int i = 0; i < n; ++i) s.r[i] = s.g[i] - s.r[i];
        for (int i = 0; i < n; ++i) rr += s.r[i] * s.r[i];
        if (s.stp == 1.0) {
            dr = s.gd - s.gdold;
            ddum = -s.gdold;
        } else {
            dr = (s.gd - s.gdold) * s.stp;
            for (int i = 0; i < n; ++i) s.d[i] *= s.stp;
            ddum = -s.gdold * s.stp;
        }
        if (dr <= s.epsmch * ddum) {
            s.nskip++;
            s.updatd = 0;
        } else {
            s.updatd = 1;
            s.iupdat++;
            matupd(s, rr, dr);
            s.info = formt(s);
            if (s.info != 0) refresh_memory(s);
        }
    }
}

}   // namespace lbfgsb
