// ref_capi.cpp -- C entry points over the reference's OWN univariate solver source, compiled UNMODIFIED from where it lies
// (REF_SRC = /root/reference/src/Rcpp20260726ai.cpp, given by oracle/Makefile) against the stand-in headers in oracle/shim/.
// TEST INFRASTRUCTURE ONLY: oracle/_ref/libbwgr_ref.so exists to pin the hand-written oracle (bwgr_oracle.hpp) against the
// reference's text; the product never loads it.  No reference source is copied into this repo: the #include below is the
// only contact, and the library can only be built where /root/reference exists.
#include <RcppEigen.h>

#include REF_SRC

#include <cstring>

namespace {
using Rcpp::List;
using Rcpp::Value;

Eigen::MatrixXf mat_f(const float* X, int n, int p) { Eigen::MatrixXf m(n, p); std::memcpy(m.data(), X, sizeof(float) * (size_t)n * p); return m; }
Eigen::VectorXf vec_f(const float* x, int n) { Eigen::VectorXf v(n); std::memcpy(v.data(), x, sizeof(float) * (size_t)n); return v; }
void put(const List* l, const char* name, double* out) {
  if (!out) return;
  const Value* v = l->get(name);
  if (!v) return;
  for (size_t i = 0; i < v->v.size(); i++) out[i] = v->v[i];
}
double scal(const List* l, const char* name) { const Value* v = l->get(name); return v && !v->v.empty() ? v->v[0] : 0.0; }
}  // namespace

extern "C" {

void ref_set_seed(uint64_t s) { R::set_seed(s); }

// model ids as in oracle.py EM_MODELS.  The sweep counts are the reference's own hard-coded ones (it = 200, or maxit / tol).
// scal = {Va, Ve, h2, Vg, pi, Lmb}
int ref_em(int model, const float* y, const float* X, int n, int p, float df, float R2, float Pi, float alpha, double* mu, double* b,
           double* d, double* hat, double* vbv, double* scal_out) {
  Eigen::VectorXf yy = vec_f(y, n);
  Eigen::MatrixXf gen = mat_f(X, n, p);
  SEXP r = nullptr;
  switch (model) {
    case 0: r = emRR(yy, gen, df, R2); break;
    case 1: r = emBA(yy, gen, df, R2); break;
    case 2: r = emBB(yy, gen, df, R2, Pi); break;
    case 3: r = emBC(yy, gen, df, R2, Pi); break;
    case 4: r = emBL(yy, gen, R2, alpha); break;
    case 5: r = emEN(yy, gen, R2, alpha); break;
    case 6: r = emDE(yy, gen, R2); break;
    case 7: r = emML(yy, gen); break;
    case 8: r = emBCpi(yy, gen, df, R2, Pi); break;
    case 9: r = lasso(yy, gen); break;
    default: return -1;
  }
  List* l = (List*)r;
  *mu = scal(l, "mu");
  put(l, "b", b); put(l, "d", d); put(l, "hat", hat);
  if (l->get("Vb") && l->get("Vb")->v.size() > 1) put(l, "Vb", vbv);
  for (int i = 0; i < 6; i++) scal_out[i] = 0;
  scal_out[0] = scal(l, "Va"); scal_out[1] = scal(l, "Ve"); scal_out[2] = scal(l, "h2"); scal_out[3] = scal(l, "Vg");
  scal_out[4] = scal(l, "pi"); scal_out[5] = scal(l, "Lmb");
  if (model == 7) scal_out[3] = scal(l, "Vb");  // emML returns its scalar marker variance as "Vb"
  delete l;
  return 0;
}

// emML(y, gen, D) with marker weights; scal as ref_em
int ref_emml_weighted(const float* y, const float* X, int n, int p, const double* D, double* mu, double* b, double* hat, double* scal_out) {
  Rcpp::Nullable<Rcpp::NumericVector> Dn(Rcpp::NumericVector(D, (size_t)p));
  List* l = (List*)emML(vec_f(y, n), mat_f(X, n, p), Dn);
  *mu = scal(l, "mu");
  put(l, "b", b); put(l, "hat", hat);
  for (int i = 0; i < 6; i++) scal_out[i] = 0;
  scal_out[0] = scal(l, "Va"); scal_out[1] = scal(l, "Ve"); scal_out[2] = scal(l, "h2"); scal_out[3] = scal(l, "Vb");
  delete l;
  return 0;
}

// model ids as in oracle.py GIBBS_MODELS.  scal = {vb, ve, h2, MSx, pi}
int ref_gibbs(int model, const float* y, const float* X, int n, int p, float it, float bi, float pi, float df, float R2, uint64_t seed,
              double* mu, double* b, double* d, double* hat, double* vbv, double* scal_out) {
  Eigen::VectorXf yy = vec_f(y, n);
  Eigen::MatrixXf gen = mat_f(X, n, p);
  R::set_seed(seed);
  SEXP r = nullptr;
  switch (model) {
    case 0: r = BayesRR(yy, gen, it, bi, df, R2); break;
    case 1: r = BayesA(yy, gen, it, bi, df, R2); break;
    case 2: r = BayesB(yy, gen, it, bi, pi, df, R2); break;
    case 3: r = BayesC(yy, gen, it, bi, pi, df, R2); break;
    case 4: r = BayesL(yy, gen, it, bi, df, R2); break;
    case 5: r = BayesCpi(yy, gen, it, bi, df, R2); break;
    case 6: r = BayesDpi(yy, gen, it, bi, df, R2); break;
    default: return -1;
  }
  List* l = (List*)r;
  *mu = scal(l, "mu");
  put(l, "b", b); put(l, "d", d); put(l, "hat", hat);
  const Value* vb = l->get("vb");
  if (vb && vb->v.size() > 1) put(l, "vb", vbv);
  for (int i = 0; i < 5; i++) scal_out[i] = 0;
  scal_out[0] = vb && vb->v.size() == 1 ? vb->v[0] : 0.0; scal_out[1] = scal(l, "ve"); scal_out[2] = scal(l, "h2");
  scal_out[3] = scal(l, "MSx"); scal_out[4] = scal(l, "pi");
  delete l;
  return 0;
}

int ref_kmup(const float* X, int n, int p, float* b, float* d, const float* xx, float* e, const float* L, float Ve, float pi, uint64_t seed) {
  R::set_seed(seed);
  List* l = (List*)KMUP(mat_f(X, n, p), vec_f(b, p), vec_f(d, p), vec_f(xx, p), vec_f(e, n), vec_f(L, p), Ve, pi);
  const Value *vb = l->get("b"), *vd = l->get("d"), *ve = l->get("e");
  for (int j = 0; j < p; j++) { b[j] = (float)vb->v[j]; d[j] = (float)vd->v[j]; }
  for (int i = 0; i < n; i++) e[i] = (float)ve->v[i];
  delete l;
  return 0;
}

// KMUP2 (:41-77): Use = row indices (0-based, as floats) of the bagged sample, E = residuals of ALL rows; e_out has length nuse
int ref_kmup2(const float* X, int n, int p, const float* Use, int nuse, float* b, float* d, const float* xx, const float* E, float* e_out,
              const float* L, float Ve, float pi, uint64_t seed) {
  R::set_seed(seed);
  List* l = (List*)KMUP2(mat_f(X, n, p), vec_f(Use, nuse), vec_f(b, p), vec_f(d, p), vec_f(xx, p), vec_f(E, n), vec_f(L, p), Ve, pi);
  const Value *vb = l->get("b"), *vd = l->get("d"), *ve = l->get("e");
  for (int j = 0; j < p; j++) { b[j] = (float)vb->v[j]; d[j] = (float)vd->v[j]; }
  for (int i = 0; i < nuse; i++) e_out[i] = (float)ve->v[i];
  delete l;
  return 0;
}

// GSRR (:1597-1628) / GSFLM (:1564-1594): warm-start Gauss-Seidel solvers used by mm() (R/mix.R:890-892).
// which = 0 GSRR, 1 GSFLM.  In/out: b[p], e[n], Lmb[p] (GSRR) ; scal = {mu?, Ve, ...} -- see the Python wrapper.
int ref_gs(int which, const float* y, float* e, const float* X, int n, int p, float* b, float* Lmb, const float* xx, float cxx, int maxit,
           double* out_scal) {
  List* l = which == 0 ? (List*)GSRR(vec_f(y, n), vec_f(e, n), mat_f(X, n, p), vec_f(b, p), vec_f(Lmb, p), vec_f(xx, p), cxx, maxit)
                       : (List*)GSFLM(vec_f(y, n), vec_f(e, n), mat_f(X, n, p), vec_f(b, p), vec_f(Lmb, p), vec_f(xx, p), cxx, maxit);
  const Value *vb = l->get("b"), *ve = l->get("e"), *vl = l->get("Lmb");
  if (vb) for (int j = 0; j < p; j++) b[j] = (float)vb->v[j];
  if (ve) for (int i = 0; i < n; i++) e[i] = (float)ve->v[i];
  if (vl) for (size_t j = 0; j < vl->v.size() && j < (size_t)p; j++) Lmb[j] = (float)vl->v[j];
  if (out_scal) { out_scal[0] = scal(l, "mu"); out_scal[1] = scal(l, "ve"); out_scal[2] = scal(l, "vb"); out_scal[3] = scal(l, "h2"); }
  delete l;
  return 0;
}

// two-design solvers (:990-1305).  model: 0 BayesA2, 1 BayesB2, 2 BayesRR2, 3 emML2.  Same outputs as orc_two_design.
int ref_two_design(int model, const float* y, const float* X1, const float* X2, int n, int p1, int p2, float it, float bi, float pi, float df,
                   float R2, uint64_t seed, const double* D1, const double* D2, double* b1, double* b2, double* d1, double* d2, double* vb1,
                   double* vb2, double* hat, double* u1, double* u2, double* scal_out) {
  Eigen::VectorXf yy = vec_f(y, n);
  Eigen::MatrixXf A = mat_f(X1, n, p1), B = mat_f(X2, n, p2);
  R::set_seed(seed);
  SEXP r = nullptr;
  if (model == 0) r = BayesA2(yy, A, B, it, bi, df, R2);
  else if (model == 1) r = BayesB2(yy, A, B, it, bi, pi, df, R2);
  else if (model == 2) r = BayesRR2(yy, A, B, it, bi, df, R2);
  else if (model == 3) {
    Rcpp::Nullable<Rcpp::NumericVector> n1, n2;
    if (D1) n1 = Rcpp::Nullable<Rcpp::NumericVector>(Rcpp::NumericVector(D1, (size_t)p1));
    if (D2) n2 = Rcpp::Nullable<Rcpp::NumericVector>(Rcpp::NumericVector(D2, (size_t)p2));
    r = emML2(yy, A, B, n1, n2);
  } else return -1;
  List* l = (List*)r;
  put(l, "b1", b1); put(l, "b2", b2); put(l, "d1", d1); put(l, "d2", d2); put(l, "hat", hat); put(l, "u1", u1); put(l, "u2", u2);
  for (int i = 0; i < 8; i++) scal_out[i] = 0;
  scal_out[0] = scal(l, "mu"); scal_out[1] = model == 3 ? scal(l, "Ve") : scal(l, "ve"); scal_out[2] = scal(l, "h2");
  if (model == 3) { scal_out[3] = scal(l, "Vb1"); scal_out[4] = scal(l, "Vb2"); scal_out[5] = scal(l, "MSx1"); scal_out[6] = scal(l, "MSx2"); }
  else if (model == 2) { scal_out[3] = scal(l, "vb1"); scal_out[4] = scal(l, "vb2"); }
  else { put(l, "vb1", vb1); put(l, "vb2", vb2); }
  delete l;
  return 0;
}

// CNT (:1308) and IMP (:1316): column centring / mean imputation of the genotype matrix (in place, column-major)
int ref_cnt_imp(int which, float* X, int n, int p) {
  Eigen::MatrixXf o = which == 0 ? CNT(mat_f(X, n, p)) : IMP(mat_f(X, n, p));
  std::memcpy(X, o.data(), sizeof(float) * (size_t)n * p);
  return 0;
}

}  // extern "C"
