/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement (plain C99, libm, C complex) of the BISIP MCMC likelihood hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
 * load this library; nothing under bisip_b200/ links or calls it.
 *
 * Parity status: the forward / log-probability half is PINNED against the reference's own
 * Cython + NumPy code built into oracle/_ref/ (tests/test_oracle.py, golden vectors under
 * tests/golden/).  The sampler half restates emcee's StretchMove/RedBlueMove (emcee is a
 * third-party, unpinned dependency — reference requirements.txt:2 — whose source is absent
 * from /root/reference and from this image): "parity unpinned" for the sampler; it is
 * anchored on SURVEY.md App. B and on the Monte-Carlo-error comparison with
 * oracle/emcee_restatement.py (MT19937 draw order) driving the reference log-probability.
 *
 * Reference lines followed (paths relative to /root/reference/src/bisip/):
 *   C_ColeCole   cython_funcs.pyx:33-34     ColeCole_cyth   cython_funcs.pyx:49-62
 *   C_Dias       cython_funcs.pyx:36-40     Dias2000_cyth   cython_funcs.pyx:64-73
 *   C_Shin       cython_funcs.pyx:42-44     Shin2015_cyth   cython_funcs.pyx:96-108
 *   C_Debye      cython_funcs.pyx:46-47     Decomp_cyth     cython_funcs.pyx:75-94
 *   _log_likelihood models.py:59-62   _log_prior models.py:64-69   _log_probability models.py:71-76
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -std=c11 bisip_oracle.c -lm
 *        (-ffp-contract=off so that a*b+c is never fused: the proposal arithmetic must
 *        round exactly like the explicitly unfused CUDA code).
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef double complex cplx;

enum { BISIP_MODEL_COLECOLE = 0, BISIP_MODEL_DIAS = 1, BISIP_MODEL_SHIN = 2, BISIP_MODEL_DECOMP = 3 };

static inline cplx cx(double re) { return CMPLX(re, 0.0); }
static const cplx JAY = CMPLX(0.0, 1.0); /* cython_funcs.pyx:26-28 */

/* cython_funcs.pyx:33-34: m*(1 - 1/(1 + (j*w*exp(lt))**c)) */
static cplx c_colecole(double w, double m, double lt, double c) {
  cplx base = (JAY * cx(w)) * cx(exp(lt));
  cplx p = cpow(base, cx(c));
  return cx(m) * (cx(1.0) - cx(1.0) / (cx(1.0) + p));
}

/* cython_funcs.pyx:36-40 */
static cplx c_dias(double w, double R0, double m, double log_tau, double eta, double delta) {
  double tau_p = exp(log_tau) * (1 / delta - 1) / (1 - m);
  double tau_pp = pow(exp(log_tau), 2.0) * pow(eta, 2.0);
  cplx mu = (JAY * cx(w)) * cx(exp(log_tau)) + cpow((JAY * cx(w)) * cx(tau_pp), cx(0.5));
  cplx inner = cx(1.0) + ((JAY * cx(w)) * cx(tau_p)) * (cx(1.0) + cx(1.0) / mu);
  return cx(R0) * (cx(1.0) - cx(m) * (cx(1.0) - cx(1.0) / inner));
}

/* cython_funcs.pyx:42-44 */
static cplx c_shin(double w, double R, double log_Q, double n) {
  cplx z_cpe = cx(1.0) / (cx(exp(log_Q)) * cpow(JAY * cx(w), cx(n)));
  return cpow(cx(1.0) / z_cpe + cx(1.0 / R), cx(-1.0));
}

/* cython_funcs.pyx:46-47 */
static cplx c_debye(double w, double m, double tau, double c) {
  return cx(m) * (cx(1.0) - cx(1.0) / (cx(1.0) + cpow((JAY * cx(w)) * cx(tau), cx(c))));
}

/* ColeCole_cyth, cython_funcs.pyx:49-62.  Z is (2,N) row-major: [real; imag]. */
void oracle_forward_colecole(const double *w, int N, double R0, const double *m, const double *lt,
                             const double *c, int D, double *Z) {
  for (int j = 0; j < N; ++j) {
    cplx z = 0;
    for (int i = 0; i < D; ++i) z += c_colecole(w[j], m[i], lt[i], c[i]);
    z = cx(R0) * (cx(1.0) - z);
    Z[j] = creal(z);
    Z[N + j] = cimag(z);
  }
}

/* Dias2000_cyth, cython_funcs.pyx:64-73 */
void oracle_forward_dias(const double *w, int N, double R0, double m, double log_tau, double eta,
                         double delta, double *Z) {
  for (int j = 0; j < N; ++j) {
    cplx z = c_dias(w[j], R0, m, log_tau, eta, delta);
    Z[j] = creal(z);
    Z[N + j] = cimag(z);
  }
}

/* Shin2015_cyth, cython_funcs.pyx:96-108 */
void oracle_forward_shin(const double *w, int N, const double *R, const double *log_Q, const double *n,
                         int D, double *Z) {
  for (int j = 0; j < N; ++j) {
    cplx z = 0;
    for (int i = 0; i < D; ++i) z += c_shin(w[j], R[i], log_Q[i], n[i]);
    Z[j] = creal(z);
    Z[N + j] = cimag(z);
  }
}

/* Decomp_cyth, cython_funcs.pyx:75-94.  log_taus is (D, S) row-major (D = poly_deg+1). */
void oracle_forward_decomp(const double *w, int N, const double *taus, const double *log_taus, int S,
                           double c_exp, double R0, const double *a, int D, double *Z) {
  double *M = (double *)calloc((size_t)S, sizeof(double));
  for (int i = 0; i < D; ++i)
    for (int k = 0; k < S; ++k) M[k] = M[k] + a[i] * log_taus[(size_t)i * S + k];
  for (int j = 0; j < N; ++j) {
    cplx z = 0;
    for (int k = 0; k < S; ++k) z += c_debye(w[j], M[k], taus[k], c_exp);
    z = cx(R0) * (cx(1.0) - z);
    Z[j] = creal(z);
    Z[N + j] = cimag(z);
  }
  free(M);
}

/* ---- model description shared by the log-probability and the sampler ---- */
typedef struct {
  int model;         /* BISIP_MODEL_* */
  int ndim;
  int N;             /* frequencies */
  int n_modes;       /* ColeCole modes */
  int S;             /* decomposition: taus */
  int D;             /* decomposition: poly_deg+1 */
  double c_exp;      /* decomposition */
  const double *w;       /* (N) */
  const double *taus;    /* (S) */
  const double *log_taus;/* (D,S) */
  const double *y;       /* (2,N) zn */
  const double *yerr;    /* (2,N) zn_err */
  const double *bounds;  /* (2,ndim): row 0 lower, row 1 upper */
} oracle_problem;

void oracle_forward(const oracle_problem *p, const double *theta, double *Z) {
  switch (p->model) {
    case BISIP_MODEL_COLECOLE: { /* models.py:267-271 */
      int K = p->n_modes;
      oracle_forward_colecole(p->w, p->N, theta[0], theta + 1, theta + 1 + K, theta + 1 + 2 * K, K, Z);
    } break;
    case BISIP_MODEL_DIAS: /* models.py:305 */
      oracle_forward_dias(p->w, p->N, theta[0], theta[1], theta[2], theta[3], theta[4], Z);
      break;
    case BISIP_MODEL_SHIN: /* models.py:345-349 */
      oracle_forward_shin(p->w, p->N, theta, theta + 2, theta + 4, 2, Z);
      break;
    default: /* models.py:228-229 */
      oracle_forward_decomp(p->w, p->N, p->taus, p->log_taus, p->S, p->c_exp, theta[0], theta + 1, p->D, Z);
  }
}

/* models.py:64-69: strict inequalities; NaN theta => -inf */
double oracle_log_prior(const double *theta, const double *bounds, int ndim) {
  for (int d = 0; d < ndim; ++d)
    if (!(bounds[d] < theta[d])) return -INFINITY;
  for (int d = 0; d < ndim; ++d)
    if (!(theta[d] < bounds[ndim + d])) return -INFINITY;
  return 0.0;
}

/* models.py:59-62: -0.5*sum((y-f)**2/sigma2 + 2*log(sigma2)) over the (2,N) array.
 * NumPy sums pairwise; here the 2N terms are added in index order (agreement ~1e-15 rel). */
double oracle_log_likelihood(const oracle_problem *p, const double *theta, double *Zwork) {
  oracle_forward(p, theta, Zwork);
  double s = 0.0;
  for (int i = 0; i < 2 * p->N; ++i) {
    double sigma2 = p->yerr[i] * p->yerr[i];
    double r = p->y[i] - Zwork[i];
    s += r * r / sigma2 + 2 * log(sigma2);
  }
  return -0.5 * s;
}

/* models.py:71-76 */
double oracle_log_probability(const oracle_problem *p, const double *theta, double *Zwork) {
  double lp = oracle_log_prior(theta, p->bounds, p->ndim);
  if (!isfinite(lp)) return -INFINITY;
  return lp + oracle_log_likelihood(p, theta, Zwork);
}

void oracle_log_probability_many(const oracle_problem *p, const double *theta, int n, double *out) {
  double *Z = (double *)malloc(sizeof(double) * 2 * (size_t)p->N);
  for (int i = 0; i < n; ++i) out[i] = oracle_log_probability(p, theta + (size_t)i * p->ndim, Z);
  free(Z);
}

void oracle_forward_many(const oracle_problem *p, const double *theta, int n, double *Zout) {
  for (int i = 0; i < n; ++i) oracle_forward(p, theta + (size_t)i * p->ndim, Zout + (size_t)i * 2 * p->N);
}

/* ---- Philox4x32-10 (Salmon et al. 2011, Random123); KAT in tests/test_oracle.py ---- */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline double u53(uint32_t a, uint32_t b) { /* NumPy legacy random_sample bit recipe */
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

/*
 * Stretch-move ensemble sampler with the SAME Philox stream layout as the CUDA kernel
 * (bisip_b200/csrc/sampler.cuh), restating emcee's RedBlueMove/StretchMove:
 *   counter = (index, step, spectrum, purpose), key = (seed_lo, seed_hi)
 *   purpose 0: shuffle keys — walker i uses word (i&3) of counter index (i>>2) with its low
 *              ceil(log2 W) bits replaced by i (unique keys); walkers are ranked by key;
 *              ranks [0,H0) form split 0, [H0,W) split 1, H0=(W+1)/2
 *   purpose 1+s: proposal p of split s, ONE call for all its draws: stretch u = u53(x0,x1);
 *              partner = mulhi(x2, Nc); accept draw u2 = u53(x3, (x2 << 16) | 0x8000) — 43 random
 *              bits (x3 and the low half of x2, which the partner index does not depend on),
 *              centred so that u2 > 0
 * chain (nkeep,W,ndim), logp (nkeep,W): steps t with t>=discard+thin-1 and
 * (t-(discard+thin-1))%thin==0 (emcee backend.get_value slicing).  accepted (W) int32.
 * coords (W,ndim) in: p0, out: final ensemble.  Returns 0, or 1 if a NaN log-prob was seen.
 */
int oracle_ensemble_run(const oracle_problem *p, double *coords, int W, int nsteps, int step0,
                        uint64_t seed, uint32_t spectrum, double a, int discard, int thin,
                        double *chain, double *logp, int32_t *accepted, double *lp_final) {
  const int ndim = p->ndim;
  const int H0 = (W + 1) / 2;
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  double *Z = (double *)malloc(sizeof(double) * 2 * (size_t)p->N);
  double *lp = (double *)malloc(sizeof(double) * (size_t)W);
  double *q = (double *)malloc(sizeof(double) * (size_t)H0 * ndim);
  double *lpq = (double *)malloc(sizeof(double) * (size_t)H0);
  double *fac = (double *)malloc(sizeof(double) * (size_t)H0);
  double *lu2 = (double *)malloc(sizeof(double) * (size_t)H0);
  uint32_t *keys = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)W);
  int *list = (int *)malloc(sizeof(int) * (size_t)W);
  int nan_seen = 0;
  for (int i = 0; i < W; ++i) {
    lp[i] = oracle_log_probability(p, coords + (size_t)i * ndim, Z);
    if (isnan(lp[i])) nan_seen = 1;
    accepted[i] = 0;
  }
  const int first = discard + thin - 1;
  int kept = 0;
  uint32_t kmask = 1;
  while ((int)kmask < W) kmask <<= 1;
  kmask -= 1;
  for (int it = 0; it < nsteps; ++it) {
    const uint32_t t = (uint32_t)(step0 + it);
    for (int i = 0; i < W; ++i) {
      uint32_t ctr[4] = {(uint32_t)(i >> 2), t, spectrum, 0u}, out[4];
      oracle_philox4x32_10(ctr, key, out);
      keys[i] = (out[i & 3] & ~kmask) | (uint32_t)i;   /* unique: low bits carry the walker index */
    }
    for (int i = 0; i < W; ++i) {
      int rank = 0;
      for (int j = 0; j < W; ++j) rank += keys[j] < keys[i];
      list[rank] = i;
    }
    for (int s = 0; s < 2; ++s) {
      const int off = s ? H0 : 0, Hs = s ? W - H0 : H0;
      const int coff = s ? 0 : H0, Nc = W - Hs;
      for (int pp = 0; pp < Hs; ++pp) {
        uint32_t ctr[4] = {(uint32_t)pp, t, spectrum, (uint32_t)(1 + s)}, x[4];
        oracle_philox4x32_10(ctr, key, x);
        double u = u53(x[0], x[1]);
        lu2[pp] = log(u53(x[3], (x[2] << 16) | 0x8000u));
        double zr = (a - 1.0) * u + 1.0;
        double zz = zr * zr / a;
        int r = (int)(((uint64_t)x[2] * (uint64_t)Nc) >> 32);
        const double *cj = coords + (size_t)list[coff + r] * ndim;
        const double *sk = coords + (size_t)list[off + pp] * ndim;
        for (int d = 0; d < ndim; ++d) q[(size_t)pp * ndim + d] = cj[d] - (cj[d] - sk[d]) * zz;
        fac[pp] = (ndim - 1.0) * log(zz);
      }
      for (int pp = 0; pp < Hs; ++pp) {
        lpq[pp] = oracle_log_probability(p, q + (size_t)pp * ndim, Z);
        if (isnan(lpq[pp])) nan_seen = 1;
      }
      for (int pp = 0; pp < Hs; ++pp) {
        double lu = lu2[pp];
        int k = list[off + pp];
        double lnpdiff = fac[pp] + lpq[pp] - lp[k];
        if (lnpdiff > lu) {
          memcpy(coords + (size_t)k * ndim, q + (size_t)pp * ndim, sizeof(double) * ndim);
          lp[k] = lpq[pp];
          accepted[k] += 1;
        }
      }
    }
    if (it >= first && (it - first) % thin == 0) {
      if (chain) memcpy(chain + (size_t)kept * W * ndim, coords, sizeof(double) * (size_t)W * ndim);
      if (logp) memcpy(logp + (size_t)kept * W, lp, sizeof(double) * (size_t)W);
      ++kept;
    }
  }
  if (lp_final) memcpy(lp_final, lp, sizeof(double) * (size_t)W);
  free(Z); free(lp); free(q); free(lpq); free(fac); free(lu2); free(keys); free(list);
  return nan_seen;
}
