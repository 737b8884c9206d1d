// r-pkg/src/bwgr_b200_shim.cpp -- the Rcpp side of the drop-in: one body per replaced `_bWGR_<fn>` symbol of the reference's
// generated glue (src/RcppExports.cpp), same symbol, same SEXP arguments in the same order, same returned list.  Each body
// forwards R's own column-major double memory to the C ABI of include/bwgr_b200.h and wraps the caller-allocated outputs.
// Add this file to the package's src/, delete the bodies of the same names from src/RcppExports.cpp (the CallEntries table
// and R/RcppExports.R stay as they are), and link libbwgr_b200.so (src/Makevars: PKG_LIBS = -lbwgr_b200).
// R / Rcpp do not exist in the build image of this repository, so this file is not compiled by the test-suite: the Python
// mirror (bwgr_b200/api.py) makes the SAME C calls with the same argument marshalling and is what tests/ exercises.
#include <Rcpp.h>

#include "bwgr_b200.h"

using Rcpp::List;
using Rcpp::Named;
using Rcpp::NumericMatrix;
using Rcpp::NumericVector;

namespace {

bwgr_handle* handle() {  // one handle (= one GPU) per R session
  static bwgr_handle* h = nullptr;
  if (!h && bwgr_create(0, &h) != BWGR_OK) Rcpp::stop(bwgr_last_error());  // no B200 -> R error; there is no CPU path
  return h;
}
void check(int rc) { if (rc != BWGR_OK) Rcpp::stop(bwgr_last_error()); }

// the store persists on the handle: the same matrix object (pointer, shape) fitted again is not packed again
bwgr_handle* load(SEXP genSEXP, int64_t* n, int64_t* p, bool centred_ok = false) {
  static const double* last = nullptr;
  static int64_t ln = 0, lp = 0;
  static bool lc = false;
  NumericMatrix gen(genSEXP);  // R's own memory, no copy
  *n = gen.nrow(); *p = gen.ncol();
  bwgr_handle* h = handle();
  if (gen.begin() != last || *n != ln || *p != lp || lc != centred_ok) {
    // exact integer codes (plus one constant per column for the solvers that centre the columns) go to the int8 store; any other
    // real-valued matrix -- NA cells imputed with column means (R/wgr.R:13-19), IMP() / CNT() output -- to the float32 store, the type the
    // reference's own glue narrows to (RcppExports.cpp:115-116), which the grid kernel family serves
    int rc = centred_ok ? bwgr_geno_load_f64_centred(h, gen.begin(), *n, *p, *n, BWGR_STORE_I8) : bwgr_geno_load_f64(h, gen.begin(), *n, *p, *n, BWGR_STORE_I8);
    if (rc == BWGR_ERR_ARG) rc = bwgr_geno_load_f64(h, gen.begin(), *n, *p, *n, BWGR_STORE_F32);
    check(rc);
    last = gen.begin(); ln = *n; lp = *p; lc = centred_ok;
  }
  return h;
}
uint64_t seed_from_R() {  // set.seed() stays in control: one draw from R's stream seeds the Philox counters
  Rcpp::RNGScope scope;
  return (uint64_t)(unif_rand() * 9007199254740992.0);
}

struct EmFit {
  NumericVector b, d, hat, vb;
  double mu = 0, scal[BWGR_NSCAL] = {0, 0, 0, 0, 0, 0};
  int its = 0;
};
EmFit em(int model, SEXP ySEXP, SEXP genSEXP, double df, double R2, double Pi, double alpha, const double* weights = nullptr) {
  int64_t n, p;
  bwgr_handle* h = load(genSEXP, &n, &p);
  NumericVector y(ySEXP);
  if (y.size() != n) Rcpp::stop("y and gen disagree on the number of individuals");
  EmFit f;
  f.b = NumericVector(p); f.d = NumericVector(p); f.hat = NumericVector(n); f.vb = NumericVector(p);
  bwgr_em_params par = {model, 1, -1, df, R2, Pi, alpha, nullptr, weights};
  bwgr_em_out out = {&f.mu, f.b.begin(), f.d.begin(), f.hat.begin(), f.vb.begin(), f.scal, &f.its};
  check(bwgr_em_fit(h, &par, y.begin(), &out));
  return f;
}

struct GibbsFit {
  NumericVector b, d, hat, vb;
  double mu = 0, scal[BWGR_NSCAL] = {0, 0, 0, 0, 0, 0};
};
GibbsFit gibbs(int model, SEXP ySEXP, SEXP XSEXP, double it, double bi, double pi, double df, double R2) {
  int64_t n, p;
  bwgr_handle* h = load(XSEXP, &n, &p);
  NumericVector y(ySEXP);
  if (y.size() != n) Rcpp::stop("y and X disagree on the number of individuals");
  GibbsFit f;
  f.b = NumericVector(p); f.d = NumericVector(p); f.hat = NumericVector(n); f.vb = NumericVector(p);
  bwgr_gibbs_params par = {model, 1, (int)it, (int)bi, pi, df, R2, seed_from_R()};  // it, bi arrive as floats and are cast (:611, :642)
  bwgr_gibbs_out out = {&f.mu, f.b.begin(), f.d.begin(), f.hat.begin(), f.vb.begin(), f.scal};
  check(bwgr_gibbs_fit(h, &par, y.begin(), &out));
  return f;
}

}  // namespace

// ---- univariate EM solvers (reference glue: src/RcppExports.cpp:53-163, :451-476; lists: Rcpp20260726ai.cpp:122-127 ...) ----
RcppExport SEXP _bWGR_emRR(SEXP ySEXP, SEXP genSEXP, SEXP dfSEXP, SEXP R2SEXP) {
BEGIN_RCPP
  EmFit f = em(BWGR_EM_RR, ySEXP, genSEXP, Rcpp::as<double>(dfSEXP), Rcpp::as<double>(R2SEXP), 0.75, 0.02);
  return List::create(Named("mu") = f.mu, Named("b") = f.b, Named("hat") = f.hat, Named("Va") = f.scal[0], Named("Ve") = f.scal[1],
                      Named("h2") = f.scal[2]);  // Rcpp20260726ai.cpp:348-353
END_RCPP
}
RcppExport SEXP _bWGR_emBA(SEXP ySEXP, SEXP genSEXP, SEXP dfSEXP, SEXP R2SEXP) {
BEGIN_RCPP
  EmFit f = em(BWGR_EM_BA, ySEXP, genSEXP, Rcpp::as<double>(dfSEXP), Rcpp::as<double>(R2SEXP), 0.75, 0.02);
  return List::create(Named("mu") = f.mu, Named("b") = f.b, Named("hat") = f.hat, Named("Vb") = f.vb, Named("Ve") = f.scal[1],
                      Named("h2") = f.scal[2]);  // :122-127
END_RCPP
}
RcppExport SEXP _bWGR_emBB(SEXP ySEXP, SEXP genSEXP, SEXP dfSEXP, SEXP R2SEXP, SEXP PiSEXP) {
BEGIN_RCPP
  EmFit f = em(BWGR_EM_BB, ySEXP, genSEXP, Rcpp::as<double>(dfSEXP), Rcpp::as<double>(R2SEXP), Rcpp::as<double>(PiSEXP), 0.02);
  return List::create(Named("mu") = f.mu, Named("b") = f.b, Named("d") = f.d, Named("hat") = f.hat, Named("Vb") = f.vb,
                      Named("Ve") = f.scal[1], Named("h2") = f.scal[2]);  // :180-186
END_RCPP
}
RcppExport SEXP _bWGR_emBC(SEXP ySEXP, SEXP genSEXP, SEXP dfSEXP, SEXP R2SEXP, SEXP PiSEXP) {
BEGIN_RCPP
  EmFit f = em(BWGR_EM_BC, ySEXP, genSEXP, Rcpp::as<double>(dfSEXP), Rcpp::as<double>(R2SEXP), Rcpp::as<double>(PiSEXP), 0.02);
  return List::create(Named("mu") = f.mu, Named("b") = f.b, Named("d") = f.d, Named("hat") = f.hat, Named("Vg") = f.scal[3],
                      Named("Va") = f.scal[0], Named("Ve") = f.scal[1], Named("h2") = f.scal[2]);  // :239-246
END_RCPP
}
RcppExport SEXP _bWGR_emDE(SEXP ySEXP, SEXP genSEXP, SEXP R2SEXP) {
BEGIN_RCPP
  EmFit f = em(BWGR_EM_DE, ySEXP, genSEXP, 10, Rcpp::as<double>(R2SEXP), 0.75, 0.02);
  return List::create(Named("mu") = f.mu, Named("b") = f.b, Named("hat") = f.hat, Named("Vb") = f.vb, Named("Ve") = f.scal[1],
                      Named("h2") = f.scal[2]);  // :299-305
END_RCPP
}
RcppExport SEXP _bWGR_emBL(SEXP ySEXP, SEXP genSEXP, SEXP R2SEXP, SEXP alphaSEXP) {
BEGIN_RCPP
  EmFit f = em(BWGR_EM_BL, ySEXP, genSEXP, 10, Rcpp::as<double>(R2SEXP), 0.75, Rcpp::as<double>(alphaSEXP));
  return List::create(Named("mu") = f.mu, Named("b") = f.b, Named("hat") = f.hat, Named("h2") = f.scal[2]);  // :393-396
END_RCPP
}
RcppExport SEXP _bWGR_emEN(SEXP ySEXP, SEXP genSEXP, SEXP R2SEXP, SEXP alphaSEXP) {
BEGIN_RCPP
  EmFit f = em(BWGR_EM_EN, ySEXP, genSEXP, 10, Rcpp::as<double>(R2SEXP), 0.75, Rcpp::as<double>(alphaSEXP));
  return List::create(Named("mu") = f.mu, Named("b") = f.b, Named("hat") = f.hat, Named("Va") = f.scal[0], Named("Ve") = f.scal[1],
                      Named("h2") = f.scal[2]);  // :454-459
END_RCPP
}
RcppExport SEXP _bWGR_emML(SEXP ySEXP, SEXP genSEXP, SEXP DSEXP) {
BEGIN_RCPP
  NumericVector D;  // optional marker weights (:471-475): the penalty of marker j becomes Lmb / D[j]
  if (!Rf_isNull(DSEXP)) D = NumericVector(DSEXP);
  EmFit f = em(BWGR_EM_ML, ySEXP, genSEXP, 10, 0.5, 0.75, 0.02, Rf_isNull(DSEXP) ? nullptr : D.begin());
  return List::create(Named("mu") = f.mu, Named("b") = f.b, Named("hat") = f.hat, Named("h2") = f.scal[2], Named("Vb") = f.scal[3],
                      Named("Va") = f.scal[0], Named("Ve") = f.scal[1]);  // :514-520
END_RCPP
}
RcppExport SEXP _bWGR_emBCpi(SEXP ySEXP, SEXP genSEXP, SEXP dfSEXP, SEXP R2SEXP, SEXP PiSEXP) {
BEGIN_RCPP
  EmFit f = em(BWGR_EM_BCPI, ySEXP, genSEXP, Rcpp::as<double>(dfSEXP), Rcpp::as<double>(R2SEXP), Rcpp::as<double>(PiSEXP), 0.02);
  return List::create(Named("mu") = f.mu, Named("b") = f.b, Named("d") = f.d, Named("pi") = f.scal[4], Named("hat") = f.hat,
                      Named("Vg") = f.scal[3], Named("Va") = f.scal[0], Named("Ve") = f.scal[1], Named("h2") = f.scal[2]);  // :1539-1546
END_RCPP
}
RcppExport SEXP _bWGR_lasso(SEXP ySEXP, SEXP genSEXP) {
BEGIN_RCPP
  EmFit f = em(BWGR_EM_LASSO, ySEXP, genSEXP, 10, 0.5, 0.75, 0.02);
  return List::create(Named("mu") = f.mu, Named("b") = f.b, Named("h2") = f.scal[2], Named("hat") = f.hat, Named("Lmb") = f.scal[5]);  // :1494-1498
END_RCPP
}

// ---- univariate Gibbs samplers (glue :177-289; lists :628-634, :690-698, :750-758, :802-808, :848-854, :912-920, :978-986) ----
#define BWGR_GIBBS_ARGS6 SEXP ySEXP, SEXP XSEXP, SEXP itSEXP, SEXP biSEXP, SEXP dfSEXP, SEXP R2SEXP
#define BWGR_GIBBS_ARGS7 SEXP ySEXP, SEXP XSEXP, SEXP itSEXP, SEXP biSEXP, SEXP piSEXP, SEXP dfSEXP, SEXP R2SEXP
RcppExport SEXP _bWGR_BayesRR(BWGR_GIBBS_ARGS6) {
BEGIN_RCPP
  GibbsFit f = gibbs(BWGR_GIBBS_RR, ySEXP, XSEXP, Rcpp::as<double>(itSEXP), Rcpp::as<double>(biSEXP), 0, Rcpp::as<double>(dfSEXP), Rcpp::as<double>(R2SEXP));
  return List::create(Named("mu") = f.mu, Named("b") = f.b, Named("hat") = f.hat, Named("vb") = f.scal[0], Named("ve") = f.scal[1],
                      Named("h2") = f.scal[2], Named("MSx") = f.scal[3]);
END_RCPP
}
RcppExport SEXP _bWGR_BayesA(BWGR_GIBBS_ARGS6) {
BEGIN_RCPP
  GibbsFit f = gibbs(BWGR_GIBBS_A, ySEXP, XSEXP, Rcpp::as<double>(itSEXP), Rcpp::as<double>(biSEXP), 0, Rcpp::as<double>(dfSEXP), Rcpp::as<double>(R2SEXP));
  return List::create(Named("mu") = f.mu, Named("b") = f.b, Named("hat") = f.hat, Named("vb") = f.vb, Named("ve") = f.scal[1],
                      Named("h2") = f.scal[2], Named("MSx") = f.scal[3]);
END_RCPP
}
RcppExport SEXP _bWGR_BayesB(BWGR_GIBBS_ARGS7) {
BEGIN_RCPP
  GibbsFit f = gibbs(BWGR_GIBBS_B, ySEXP, XSEXP, Rcpp::as<double>(itSEXP), Rcpp::as<double>(biSEXP), Rcpp::as<double>(piSEXP), Rcpp::as<double>(dfSEXP), Rcpp::as<double>(R2SEXP));
  return List::create(Named("mu") = f.mu, Named("b") = f.b, Named("d") = f.d, Named("hat") = f.hat, Named("vb") = f.vb,
                      Named("ve") = f.scal[1], Named("h2") = f.scal[2], Named("MSx") = f.scal[3]);
END_RCPP
}
RcppExport SEXP _bWGR_BayesC(BWGR_GIBBS_ARGS7) {
BEGIN_RCPP
  GibbsFit f = gibbs(BWGR_GIBBS_C, ySEXP, XSEXP, Rcpp::as<double>(itSEXP), Rcpp::as<double>(biSEXP), Rcpp::as<double>(piSEXP), Rcpp::as<double>(dfSEXP), Rcpp::as<double>(R2SEXP));
  return List::create(Named("mu") = f.mu, Named("b") = f.b, Named("d") = f.d, Named("hat") = f.hat, Named("vb") = f.scal[0],
                      Named("ve") = f.scal[1], Named("h2") = f.scal[2], Named("MSx") = f.scal[3]);
END_RCPP
}
RcppExport SEXP _bWGR_BayesL(BWGR_GIBBS_ARGS6) {
BEGIN_RCPP
  GibbsFit f = gibbs(BWGR_GIBBS_L, ySEXP, XSEXP, Rcpp::as<double>(itSEXP), Rcpp::as<double>(biSEXP), 0, Rcpp::as<double>(dfSEXP), Rcpp::as<double>(R2SEXP));
  return List::create(Named("mu") = f.mu, Named("b") = f.b, Named("hat") = f.hat, Named("vb") = f.vb, Named("ve") = f.scal[1],
                      Named("h2") = f.scal[2], Named("MSx") = f.scal[3]);
END_RCPP
}
static SEXP gibbs_pi_list(const GibbsFit& f, bool per_marker_vb) {  // BayesCpi / BayesDpi: pi and PVAL = -log(1 - d) in place of MSx
  NumericVector pval(f.d.size());
  for (R_xlen_t j = 0; j < f.d.size(); j++) pval[j] = -std::log(1.0 - f.d[j]);
  if (per_marker_vb)
    return List::create(Named("mu") = f.mu, Named("b") = f.b, Named("d") = f.d, Named("pi") = f.scal[4], Named("hat") = f.hat,
                        Named("h2") = f.scal[2], Named("vb") = f.vb, Named("ve") = f.scal[1], Named("PVAL") = pval);
  return List::create(Named("mu") = f.mu, Named("b") = f.b, Named("d") = f.d, Named("pi") = f.scal[4], Named("hat") = f.hat,
                      Named("h2") = f.scal[2], Named("vb") = f.scal[0], Named("ve") = f.scal[1], Named("PVAL") = pval);
}
RcppExport SEXP _bWGR_BayesCpi(BWGR_GIBBS_ARGS6) {
BEGIN_RCPP
  return gibbs_pi_list(gibbs(BWGR_GIBBS_CPI, ySEXP, XSEXP, Rcpp::as<double>(itSEXP), Rcpp::as<double>(biSEXP), 0, Rcpp::as<double>(dfSEXP), Rcpp::as<double>(R2SEXP)), false);
END_RCPP
}
RcppExport SEXP _bWGR_BayesDpi(BWGR_GIBBS_ARGS6) {
BEGIN_RCPP
  return gibbs_pi_list(gibbs(BWGR_GIBBS_DPI, ySEXP, XSEXP, Rcpp::as<double>(itSEXP), Rcpp::as<double>(biSEXP), 0, Rcpp::as<double>(dfSEXP), Rcpp::as<double>(R2SEXP)), true);
END_RCPP
}

// ---- KMUP: one Kuo-Mallick sweep (glue :16-31; Rcpp20260726ai.cpp:12-38).  wgr() in R/wgr.R calls this once per MCMC iteration;
// a package that wants the whole loop native calls bwgr_wgr_fit instead ----
RcppExport SEXP _bWGR_KMUP(SEXP XSEXP, SEXP bSEXP, SEXP dSEXP, SEXP xxSEXP, SEXP eSEXP, SEXP LSEXP, SEXP VeSEXP, SEXP piSEXP) {
BEGIN_RCPP
  int64_t n, p;
  bwgr_handle* h = load(XSEXP, &n, &p);
  NumericVector b = Rcpp::clone(NumericVector(bSEXP)), d = Rcpp::clone(NumericVector(dSEXP)), e = Rcpp::clone(NumericVector(eSEXP));
  NumericVector xx(xxSEXP), L(LSEXP);
  if (b.size() != p || d.size() != p || xx.size() != p || L.size() != p || e.size() != n) Rcpp::stop("KMUP: argument lengths disagree with X");
  check(bwgr_kmup_sweep(h, b.begin(), d.begin(), xx.begin(), e.begin(), L.begin(), Rcpp::as<double>(VeSEXP), Rcpp::as<double>(piSEXP), seed_from_R()));
  return List::create(Named("b") = b, Named("d") = d, Named("e") = e);
END_RCPP
}

// ---- KMUP2: the bagged sweep of wgr(bag != 1) (glue :34-50; Rcpp20260726ai.cpp:41-77).  Use = 0-based rows (R/wgr.R:68); the third list
// element is the residual of the rows in use.  The whole bagged loop native: bwgr_wgr_fit_bag ----
RcppExport SEXP _bWGR_KMUP2(SEXP XSEXP, SEXP UseSEXP, SEXP bSEXP, SEXP dSEXP, SEXP xxSEXP, SEXP ESEXP, SEXP LSEXP, SEXP VeSEXP, SEXP piSEXP) {
BEGIN_RCPP
  int64_t n, p;
  bwgr_handle* h = load(XSEXP, &n, &p);
  NumericVector b = Rcpp::clone(NumericVector(bSEXP)), d = Rcpp::clone(NumericVector(dSEXP));
  NumericVector Use(UseSEXP), xx(xxSEXP), E(ESEXP), L(LSEXP);
  if (b.size() != p || d.size() != p || xx.size() != p || L.size() != p || E.size() != n) Rcpp::stop("KMUP2: argument lengths disagree with X");
  NumericVector e(Use.size());
  check(bwgr_kmup2_sweep(h, Use.begin(), Use.size(), b.begin(), d.begin(), xx.begin(), E.begin(), e.begin(), L.begin(), Rcpp::as<double>(VeSEXP),
                         Rcpp::as<double>(piSEXP), seed_from_R()));
  return List::create(Named("b") = b, Named("d") = d, Named("e") = e);
END_RCPP
}

// ---- two-design solvers (glue :291-355; Rcpp20260726ai.cpp:990-1305): y = mu + X1 b1 + X2 b2 + e.  Two handles, one store each; the
// marker loops of the reference are bwgr_kmup_sweep on each store with the shared residual (BayesB2's step IS KMUP's, :1106-1118; the
// others are its pi = 0 case; emML2 its Ve -> 0 limit), everything around them is the reference's own driver arithmetic.  Same loops
// as bwgr_b200/api.py: _gibbs2 / emML2, which is what tests/ runs ----
namespace {
struct Design {
  bwgr_handle* h = nullptr;
  int64_t n = 0, p = 0;
  NumericVector xx;
  double MSx = 0;
};
Design design(int slot, SEXP XSEXP) {  // slot 0 / 1: one handle per design, stores cached by matrix pointer like load()
  static bwgr_handle* hs[2] = {nullptr, nullptr};
  static const double* last[2] = {nullptr, nullptr};
  static int64_t ln[2] = {0, 0}, lp[2] = {0, 0};
  NumericMatrix X(XSEXP);
  Design d;
  d.n = X.nrow(); d.p = X.ncol();
  if (!hs[slot] && bwgr_create(0, &hs[slot]) != BWGR_OK) Rcpp::stop(bwgr_last_error());
  d.h = hs[slot];
  if (X.begin() != last[slot] || d.n != ln[slot] || d.p != lp[slot]) {
    int rc = bwgr_geno_load_f64(d.h, X.begin(), d.n, d.p, d.n, BWGR_STORE_I8);
    if (rc == BWGR_ERR_ARG) rc = bwgr_geno_load_f64(d.h, X.begin(), d.n, d.p, d.n, BWGR_STORE_F32);
    check(rc);
    last[slot] = X.begin(); ln[slot] = d.n; lp[slot] = d.p;
  }
  d.xx = NumericVector(d.p);
  NumericVector sx(d.p);
  check(bwgr_geno_stats(d.h, d.xx.begin(), sx.begin()));
  for (int64_t j = 0; j < d.p; j++) d.MSx += (d.xx[j] - sx[j] * sx[j] / (double)d.n) / ((double)d.n - 1.0);  // sum of fvar(x_j)
  return d;
}
double var1(const NumericVector& y) { double m = Rcpp::mean(y), s = 0; for (double v : y) s += (v - m) * (v - m); return s / (y.size() - 1.0); }

SEXP gibbs2(int model /* 0 A2, 1 B2, 2 RR2 */, SEXP ySEXP, SEXP X1SEXP, SEXP X2SEXP, double it, double bi, double pi, double df, double R2) {
  Rcpp::RNGScope scope;
  NumericVector y(ySEXP);
  Design T[2] = {design(0, X1SEXP), design(1, X2SEXP)};
  const int64_t n = T[0].n;
  if (T[1].n != n || y.size() != n) Rcpp::stop("y, X1 and X2 disagree on the number of individuals");
  const int iit = (int)it, ibi = (int)bi;
  const double vy = var1(y), Se = (1 - R2) * df * vy;
  double Sb[2], vbs[2] = {0, 0}, VBs[2] = {0, 0}, mu = Rcpp::mean(y), ve = vy, MU = 0, VE = 0;
  NumericVector b[2], d[2], vb[2], L[2], B[2], D[2], VB[2], e = y - mu;
  for (int q = 0; q < 2; q++) {
    const int64_t p = T[q].p;
    Sb[q] = R2 * df * vy / T[q].MSx;
    b[q] = NumericVector(p); d[q] = NumericVector(p); B[q] = NumericVector(p); D[q] = NumericVector(p); VB[q] = NumericVector(p);
    vb[q] = NumericVector(p, Sb[q]);
    L[q] = model == 2 ? NumericVector(p, T[q].MSx) : NumericVector(ve / vb[q]);
  }
  for (int i = 0; i < iit; i++) {
    for (int q = 0; q < 2; q++) {
      check(bwgr_kmup_sweep(T[q].h, b[q].begin(), d[q].begin(), T[q].xx.begin(), e.begin(), L[q].begin(), ve, model == 1 ? pi : 0.0, seed_from_R()));
      if (model != 2) for (int64_t j = 0; j < T[q].p; j++) vb[q][j] = (Sb[q] + b[q][j] * b[q][j]) / R::rchisq(df + 1);
    }
    const double eM = R::rnorm(Rcpp::mean(e), std::sqrt(ve / n));
    mu += eM; e = e - eM;
    ve = (Rcpp::sum(e * e) + Se) / R::rchisq(n + df);
    for (int q = 0; q < 2; q++) {
      if (model == 2) { vbs[q] = (Sb[q] + Rcpp::sum(b[q] * b[q])) / R::rchisq(df + T[q].p); L[q] = NumericVector(T[q].p, ve / vbs[q]); }
      else L[q] = ve / vb[q];
    }
    if (i > ibi) { MU += mu; VE += ve; for (int q = 0; q < 2; q++) { B[q] += b[q]; D[q] += d[q]; VB[q] += vb[q]; VBs[q] += vbs[q]; } }
  }
  const double MCMC = it - bi;
  MU /= MCMC; VE /= MCMC;
  for (int q = 0; q < 2; q++) { B[q] = B[q] / MCMC; D[q] = D[q] / MCMC; VB[q] = VB[q] / MCMC; VBs[q] /= MCMC; }
  const double vg = model == 2 ? VBs[0] * T[0].MSx + VBs[1] * T[1].MSx : Rcpp::sum(VB[0]) + Rcpp::sum(VB[1]);
  NumericVector fit(n), u2(n);
  check(bwgr_fitted(T[0].h, B[0].begin(), MU, fit.begin()));
  check(bwgr_fitted(T[1].h, B[1].begin(), 0.0, u2.begin()));
  fit = fit + u2;
  if (model == 1)
    return List::create(Named("mu") = MU, Named("b1") = B[0], Named("b2") = B[1], Named("d1") = D[0], Named("d2") = D[1], Named("hat") = fit,
                        Named("vb1") = VB[0], Named("vb2") = VB[1], Named("ve") = VE, Named("h2") = vg / (vg + VE));
  if (model == 2)
    return List::create(Named("hat") = fit, Named("mu") = MU, Named("b1") = B[0], Named("b2") = B[1], Named("vb1") = VBs[0], Named("vb2") = VBs[1],
                        Named("ve") = VE, Named("h2") = vg / (vg + VE));
  return List::create(Named("hat") = fit, Named("mu") = MU, Named("b1") = B[0], Named("b2") = B[1], Named("vb1") = VB[0], Named("vb2") = VB[1],
                      Named("ve") = VE, Named("h2") = vg / (vg + VE));
}
}  // namespace
RcppExport SEXP _bWGR_BayesA2(SEXP ySEXP, SEXP X1SEXP, SEXP X2SEXP, SEXP itSEXP, SEXP biSEXP, SEXP dfSEXP, SEXP R2SEXP) {
BEGIN_RCPP
  return gibbs2(0, ySEXP, X1SEXP, X2SEXP, Rcpp::as<double>(itSEXP), Rcpp::as<double>(biSEXP), 0, Rcpp::as<double>(dfSEXP), Rcpp::as<double>(R2SEXP));
END_RCPP
}
RcppExport SEXP _bWGR_BayesB2(SEXP ySEXP, SEXP X1SEXP, SEXP X2SEXP, SEXP itSEXP, SEXP biSEXP, SEXP piSEXP, SEXP dfSEXP, SEXP R2SEXP) {
BEGIN_RCPP
  return gibbs2(1, ySEXP, X1SEXP, X2SEXP, Rcpp::as<double>(itSEXP), Rcpp::as<double>(biSEXP), Rcpp::as<double>(piSEXP), Rcpp::as<double>(dfSEXP),
                Rcpp::as<double>(R2SEXP));
END_RCPP
}
RcppExport SEXP _bWGR_BayesRR2(SEXP ySEXP, SEXP X1SEXP, SEXP X2SEXP, SEXP itSEXP, SEXP biSEXP, SEXP dfSEXP, SEXP R2SEXP) {
BEGIN_RCPP
  return gibbs2(2, ySEXP, X1SEXP, X2SEXP, Rcpp::as<double>(itSEXP), Rcpp::as<double>(biSEXP), 0, Rcpp::as<double>(dfSEXP), Rcpp::as<double>(R2SEXP));
END_RCPP
}
RcppExport SEXP _bWGR_emML2(SEXP ySEXP, SEXP X1SEXP, SEXP X2SEXP, SEXP D1SEXP, SEXP D2SEXP) {
BEGIN_RCPP
  NumericVector y(ySEXP);
  Design T[2] = {design(0, X1SEXP), design(1, X2SEXP)};
  const int64_t n = T[0].n;
  if (T[1].n != n || y.size() != n) Rcpp::stop("y, X1 and X2 disagree on the number of individuals");
  SEXP Ds[2] = {D1SEXP, D2SEXP};
  NumericVector b[2], bc[2], u[2], ones[2], L[2], e = y - Rcpp::mean(y), cY(n);
  double Lmb[2], vb[2] = {0, 0}, mu = Rcpp::mean(y), ve = 0;
  for (int q = 0; q < 2; q++) {
    b[q] = NumericVector(T[q].p); u[q] = NumericVector(n); ones[q] = NumericVector(T[q].p, 1.0); Lmb[q] = T[q].MSx;
    if (!Rf_isNull(Ds[q]) && NumericVector(Ds[q]).size() != T[q].p) Rcpp::stop("emML2: one weight per marker");
  }
  for (int numit = 0; numit < 350; numit++) {  // :1224, :1259-1296
    for (int q = 0; q < 2; q++) {
      bc[q] = Rcpp::clone(b[q]);
      L[q] = Rf_isNull(Ds[q]) ? NumericVector(T[q].p, Lmb[q]) : NumericVector(Lmb[q] / NumericVector(Ds[q]));
      check(bwgr_kmup_sweep(T[q].h, b[q].begin(), ones[q].begin(), T[q].xx.begin(), e.begin(), L[q].begin(), 1e-30, 0.0, 1));  // ridge step
    }
    for (int q = 0; q < 2; q++) check(bwgr_fitted(T[q].h, b[q].begin(), 0.0, u[q].begin()));  // u = X b (:1275-1276)
    const double eM = Rcpp::mean(e);
    mu += eM; e = e - eM;
    cY = u[0] + u[1] + e;
    ve = Rcpp::sum(e * cY) / n;
    for (int q = 0; q < 2; q++) { vb[q] = Rcpp::sum(u[q] * cY) / n / T[q].MSx; Lmb[q] = ve / vb[q]; }
    if (Rcpp::sum(Rcpp::abs(bc[0] - b[0])) + Rcpp::sum(Rcpp::abs(bc[1] - b[1])) < 10e-8) break;
  }
  NumericVector fit = u[0] + u[1] + mu;
  return List::create(Named("mu") = mu, Named("b1") = b[0], Named("b2") = b[1], Named("Vb1") = vb[0], Named("Vb2") = vb[1], Named("Ve") = ve,
                      Named("u1") = u[0], Named("u2") = u[1], Named("MSx1") = T[0].MSx, Named("MSx2") = T[1].MSx,
                      Named("h2") = 1 - ve / var1(y), Named("hat") = fit);
END_RCPP
}

// ---- GSRR / GSFLM: the warm-start solvers of mm() (glue :493-527; lists Rcpp20260726ai.cpp:1591-1593, :1625-1627) ----
static SEXP gs_call(int which, SEXP ySEXP, SEXP eSEXP, SEXP genSEXP, SEXP bSEXP, SEXP LmbSEXP, SEXP xxSEXP, SEXP cxxSEXP, SEXP maxitSEXP) {
  int64_t n, p;
  bwgr_handle* h = load(genSEXP, &n, &p);
  NumericVector y(ySEXP), xx(xxSEXP);
  NumericVector e = Rcpp::clone(NumericVector(eSEXP)), b = Rcpp::clone(NumericVector(bSEXP)), Lmb = Rcpp::clone(NumericVector(LmbSEXP));
  if (y.size() != n || e.size() != n || b.size() != p || Lmb.size() != p || xx.size() != p) Rcpp::stop("GSRR / GSFLM: argument lengths disagree with gen");
  NumericVector vb(p);
  double scal[4] = {0, 0, 0, 0};
  check(bwgr_gs_fit(h, which, y.begin(), e.begin(), b.begin(), Lmb.begin(), xx.begin(), Rcpp::as<double>(cxxSEXP), Rcpp::as<int>(maxitSEXP),
                    vb.begin(), scal));
  return List::create(Named("mu") = scal[0], Named("b") = b, Named("h2") = scal[1], Named("e") = e, Named("Lmb") = Lmb, Named("vb") = vb);
}
RcppExport SEXP _bWGR_GSRR(SEXP ySEXP, SEXP eSEXP, SEXP genSEXP, SEXP bSEXP, SEXP LmbSEXP, SEXP xxSEXP, SEXP cxxSEXP, SEXP maxitSEXP) {
BEGIN_RCPP
  return gs_call(0, ySEXP, eSEXP, genSEXP, bSEXP, LmbSEXP, xxSEXP, cxxSEXP, maxitSEXP);
END_RCPP
}
RcppExport SEXP _bWGR_GSFLM(SEXP ySEXP, SEXP eSEXP, SEXP genSEXP, SEXP bSEXP, SEXP LmbSEXP, SEXP xxSEXP, SEXP cxxSEXP, SEXP maxitSEXP) {
BEGIN_RCPP
  return gs_call(1, ySEXP, eSEXP, genSEXP, bSEXP, LmbSEXP, xxSEXP, cxxSEXP, maxitSEXP);
END_RCPP
}

// ---- MRR3 / MRR3F (glue :755-840; list RcppEigen20230423.cpp:687-700).  The 31 arguments after (Y, X) travel as one double array in
// the order of R/RcppExports.R:180; `verbose` stays on the R side ----
static SEXP mrr3_call(int f32_variant, SEXP YSEXP, SEXP XSEXP, const double* par) {
  int64_t n, p;
  bwgr_handle* h = load(XSEXP, &n, &p, /*centred_ok=*/true);  // MRR3 centres every column itself: mrr(Y, CNT(gen)) is accepted
  NumericMatrix Y(YSEXP);
  if (Y.nrow() != n) Rcpp::stop("Y and X disagree on the number of individuals");
  const int k = Y.ncol(), maxit = (int)par[0];
  NumericVector mu(k), h2(k), ve(k), MSx(k), cnv(3 * (size_t)maxit);
  NumericMatrix b(p, k), hat(n, k), GC(k, k), vb(k, k), W(p, k);
  int its = 0;
  check(bwgr_mrr3_fit(h, f32_variant, Y.begin(), k, par, mu.begin(), b.begin(), hat.begin(), h2.begin(), GC.begin(), vb.begin(), ve.begin(),
                      MSx.begin(), cnv.begin(), W.begin(), &its));
  NumericVector c1(its), c2(its), c3(its);
  for (int i = 0; i < its; i++) { c1[i] = cnv[i]; c2[i] = cnv[maxit + i]; c3[i] = cnv[2 * (size_t)maxit + i]; }
  return List::create(Named("mu") = mu, Named("b") = b, Named("hat") = hat, Named("h2") = h2, Named("GC") = GC, Named("vb") = vb,
                      Named("ve") = ve, Named("MSx") = MSx, Named("cnvB") = c1, Named("cnvH2") = c2, Named("cnvV") = c3,
                      Named("b_Weights") = W, Named("Its") = its);
}
#define BWGR_MRR3_ARGS                                                                                                              \
  SEXP YSEXP, SEXP XSEXP, SEXP maxitSEXP, SEXP tolSEXP, SEXP coresSEXP, SEXP THSEXP, SEXP NLfactorSEXP, SEXP InnerGSSEXP, SEXP NoInvSEXP, \
      SEXP HCSSEXP, SEXP XFASEXP, SEXP ACSSEXP, SEXP NumXFASEXP, SEXP R2SEXP, SEXP gc0SEXP, SEXP df0SEXP, SEXP updateMuSEXP,         \
      SEXP weight_prior_h2SEXP, SEXP weight_prior_gcSEXP, SEXP PenCorSEXP, SEXP MinCorSEXP, SEXP uncorH2belowSEXP,                   \
      SEXP roundGCupFromSEXP, SEXP roundGCupToSEXP, SEXP roundGCdownFromSEXP, SEXP roundGCdownToSEXP, SEXP bucketGCfromSEXP,         \
      SEXP bucketGCtoSEXP, SEXP DeflateMaxSEXP, SEXP DeflateBySEXP, SEXP OneVarBSEXP, SEXP OneVarESEXP, SEXP verboseSEXP
#define BWGR_MRR3_PAR                                                                                                               \
  {Rcpp::as<double>(maxitSEXP), Rcpp::as<double>(tolSEXP), Rcpp::as<double>(coresSEXP), Rcpp::as<double>(THSEXP),                    \
   Rcpp::as<double>(NLfactorSEXP), Rcpp::as<double>(InnerGSSEXP), Rcpp::as<double>(NoInvSEXP), Rcpp::as<double>(HCSSEXP),            \
   Rcpp::as<double>(XFASEXP), Rcpp::as<double>(ACSSEXP), Rcpp::as<double>(NumXFASEXP), Rcpp::as<double>(R2SEXP),                     \
   Rcpp::as<double>(gc0SEXP), Rcpp::as<double>(df0SEXP), Rcpp::as<double>(updateMuSEXP), Rcpp::as<double>(weight_prior_h2SEXP),      \
   Rcpp::as<double>(weight_prior_gcSEXP), Rcpp::as<double>(PenCorSEXP), Rcpp::as<double>(MinCorSEXP),                               \
   Rcpp::as<double>(uncorH2belowSEXP), Rcpp::as<double>(roundGCupFromSEXP), Rcpp::as<double>(roundGCupToSEXP),                      \
   Rcpp::as<double>(roundGCdownFromSEXP), Rcpp::as<double>(roundGCdownToSEXP), Rcpp::as<double>(bucketGCfromSEXP),                  \
   Rcpp::as<double>(bucketGCtoSEXP), Rcpp::as<double>(DeflateMaxSEXP), Rcpp::as<double>(DeflateBySEXP),                             \
   Rcpp::as<double>(OneVarBSEXP), Rcpp::as<double>(OneVarESEXP)}
RcppExport SEXP _bWGR_MRR3(BWGR_MRR3_ARGS) {
BEGIN_RCPP
  const double par[30] = BWGR_MRR3_PAR;
  return mrr3_call(0, YSEXP, XSEXP, par);
END_RCPP
}
RcppExport SEXP _bWGR_MRR3F(BWGR_MRR3_ARGS) {  // the float twin names its fifth tuning argument NonLinearFactor (R/RcppExports.R:184)
BEGIN_RCPP
  const double par[30] = BWGR_MRR3_PAR;
  return mrr3_call(1, YSEXP, XSEXP, par);
END_RCPP
}
