// ref_capi_mrr.cpp -- C entry point over the reference's MRR3 / MRR3F (src/RcppEigen20230423.cpp:317-1079).
// The rest of that file needs Eigen's sparse module, Map and BDCSVD call forms the stand-in headers do not carry, so
// oracle/Makefile extracts exactly those lines, verbatim, into oracle/_ref/mrr3_slice.cpp (git-ignored, never committed) and
// this file includes that slice.  TEST INFRASTRUCTURE ONLY (see ref_capi.cpp).
#include <RcppEigen.h>

#include "_ref/mrr3_slice.cpp"

#include <cstring>

extern "C" {

// Same contract as orc_mrr3 (oracle_capi.cpp): par[] = the 30 arguments after (Y, X) in the order of R/RcppExports.R:180.
int ref_mrr3(int f32_variant, const double* Y, const double* X, int n, int k, int p, const double* par, double* mu, double* b,
             double* hat, double* h2, double* GC, double* vb, double* ve, double* MSx, double* cnv, double* W, int* its) {
  int q = 0;
  const int maxit = (int)par[q++]; const double tol = par[q++]; const int cores = (int)par[q++]; const bool TH = par[q++] != 0;
  const double NLfactor = par[q++]; const bool InnerGS = par[q++] != 0, NoInv = par[q++] != 0, HCS = par[q++] != 0, XFA = par[q++] != 0,
               ACS = par[q++] != 0;
  const int NumXFA = (int)par[q++]; const double R2 = par[q++], gc0 = par[q++], df0 = par[q++]; const bool updateMu = par[q++] != 0;
  const double wph2 = par[q++], wpgc = par[q++], PenCor = par[q++], MinCor = par[q++], uncorH2below = par[q++], rUpFrom = par[q++],
               rUpTo = par[q++], rDownFrom = par[q++], rDownTo = par[q++], bFrom = par[q++], bTo = par[q++], DeflateMax = par[q++],
               DeflateBy = par[q++];
  const bool OneVarB = par[q++] != 0, OneVarE = par[q++] != 0;
  SEXP r;
  if (f32_variant) {
    Eigen::MatrixXf Yf(n, k), Xf(n, p);
    for (size_t i = 0; i < (size_t)n * k; i++) Yf.d[i] = (float)Y[i];
    for (size_t i = 0; i < (size_t)n * p; i++) Xf.d[i] = (float)X[i];
    r = MRR3F(Yf, Xf, maxit, (float)tol, cores, TH, (float)NLfactor, InnerGS, NoInv, HCS, XFA, ACS, NumXFA, (float)R2, (float)gc0, (float)df0,
              updateMu, (float)wph2, (float)wpgc, (float)PenCor, (float)MinCor, (float)uncorH2below, (float)rUpFrom, (float)rUpTo,
              (float)rDownFrom, (float)rDownTo, (float)bFrom, (float)bTo, (float)DeflateMax, (float)DeflateBy, OneVarB, OneVarE, false);
  } else {
    Eigen::MatrixXd Yd(n, k), Xd(n, p);
    std::memcpy(Yd.data(), Y, sizeof(double) * (size_t)n * k);
    std::memcpy(Xd.data(), X, sizeof(double) * (size_t)n * p);
    r = MRR3(Yd, Xd, maxit, tol, cores, TH, NLfactor, InnerGS, NoInv, HCS, XFA, ACS, NumXFA, R2, gc0, df0, updateMu, wph2, wpgc, PenCor,
             MinCor, uncorH2below, rUpFrom, rUpTo, rDownFrom, rDownTo, bFrom, bTo, DeflateMax, DeflateBy, OneVarB, OneVarE, false);
  }
  Rcpp::List* l = (Rcpp::List*)r;
  auto put = [&](const char* name, double* out) { const Rcpp::Value* v = l->get(name); if (v && out) for (size_t i = 0; i < v->v.size(); i++) out[i] = v->v[i]; };
  put("mu", mu); put("b", b); put("hat", hat); put("h2", h2); put("GC", GC); put("vb", vb); put("ve", ve); put("MSx", MSx); put("b_Weights", W);
  const Rcpp::Value* it = l->get("Its");
  const int nit = it ? (int)it->v[0] : 0;
  *its = nit;
  const char* names[3] = {"cnvB", "cnvH2", "cnvV"};
  for (int c = 0; c < 3; c++) { const Rcpp::Value* v = l->get(names[c]); if (v) for (size_t i = 0; i < v->v.size(); i++) cnv[(size_t)c * maxit + i] = v->v[i]; }
  delete l;
  return 0;
}

}  // extern "C"
