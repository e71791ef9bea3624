// rbvfit_b200 -- device-resident ensemble slice sampler (zeus's differential move), sm_100a.
//
// The reference runs zeus.EnsembleSampler(nwalkers, ndim, vfit.lnprob).run_mcmc(...) when sampler='zeus'
// (vfit_mcmc.py:425-440, 536-540).  zeus-mcmc is a third-party dependency that is not vendored in the reference
// (requirements.txt:11); its algorithm is restated from the published description (Karamanis & Beutler 2021,
// "Ensemble slice sampling", Algorithms 2-3; zeus 2.x defaults), exactly as rbvfit_b200/slice_sampler.py does on
// the host:
//
//   per step, for each half S of the ensemble (split as in rbv_sampler.cuh) with complement C, per walker k of S:
//       direction  eta_k = 2 mu (C_j - C_l),  j != l drawn from C                     (differential move)
//       slice level  y_k = lnp(X_k) - Exp(1)
//       stepping out: [L, R] = [-U, 1 - U]; L -= 1 while lnp(X_k + L eta_k) >= y_k (budget J), then the same for R
//       shrinking:    t ~ U(L, R); accept X_k + t eta_k when lnp >= y_k, else L = t (t < 0) or R = t
//   after the step: mu <- mu * 2 n_exp / (n_exp + n_con) while tuning is on (done by the host part of rbv_slice_run)
//
// Here every walker of the half is a small state machine (phase 0 = widening L, 1 = widening R, 2 = shrinking,
// 3 = finished) and one ITERATION advances every unfinished walker by one likelihood evaluation:
//   slice_candidate_kernel  ->  the lnprob launch over the half (finished rows are skipped through row_skip)
//                           ->  slice_update_kernel
// so the number of device batches per half-step is the longest chain of evaluations any single walker needs (the
// host sampler needs the sum of the three loops' longest chains).  Random numbers come from the same counter-based
// Philox streams as the stretch move: purpose 8 + split (partners, budget), 10 + split (slice level, bracket),
// 16 + 2 it + split (the shrink draw of iteration it).  All floating-point steps that decide the chain are written
// with explicitly rounded operations so that oracle/slice_replay.py reproduces them bit for bit in numpy.
#pragma once

#include "rbv_sampler.cuh"

namespace rbv {

struct SliceCounters {          // device, read back by the host after every iteration
  unsigned int remaining;       // rows of the half-step that are not finished after the last update
  unsigned int nexp, ncon;      // expansions / contractions of the current step
  unsigned int pad;
  unsigned long long ncall;     // likelihood rows evaluated in this run
};

struct SliceParams {
  double* coords;        // [W, ndim] current ensemble (in/out)
  double* lnp;           // [W]       its log-probabilities (in/out)
  double* cand;          // [h, ndim] candidates of the active half, h = ceil(W/2)
  double* lnp_cand;      // [h]
  double* dir;           // [h, ndim] directions
  double* z0;            // [h] slice levels
  double* lo;            // [h] bracket
  double* hi;            // [h]
  double* tcur;          // [h] the shrink draw that produced the current candidate
  int* jbudget;          // [h] expansions left on the left / right side
  int* kbudget;          // [h]
  int* phase;            // [h]
  int* skip;             // [h] row_skip of the lnprob launch (1 = finished)
  int* walker_of;        // [h] walker index of row k
  int* flag;             // bit 0: a candidate's lnprob was NaN
  SliceCounters* ctr;
  unsigned long long seed;
  double mu;
  int W, ndim, maxsteps;
};

// Start of a half-step: direction, slice level, bracket and budgets of every walker of the half (warp per row).
__global__ void __launch_bounds__(128) slice_begin_kernel(const SliceParams P, unsigned long long step, int split) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  int offS, nS, offC, nC;
  split_geometry(P.W, split, offS, nS, offC, nC);
  if (k >= nS) return;
  if (k == 0 && lane == 0 && split == 0) {
    P.ctr->nexp = 0u;
    P.ctr->ncon = 0u;
  }
  uint32_t pa, pb;
  sampler_perm(P.seed, P.W, step, pa, pb);
  const int i = walker_at(pa, pb, P.W, offS + k);
  const uint4 r = sampler_rand(P.seed, step, (uint32_t)i, 8u + (uint32_t)split);
  const int j = (int)(r.x % (uint32_t)nC);
  const int l = (int)(((uint32_t)j + 1u + r.y % (uint32_t)(nC - 1)) % (uint32_t)nC);   // l != j
  const double* cj = P.coords + (size_t)walker_at(pa, pb, P.W, offC + j) * P.ndim;
  const double* cl = P.coords + (size_t)walker_at(pa, pb, P.W, offC + l) * P.ndim;
  const double two_mu = 2.0 * P.mu;
  for (int d = lane; d < P.ndim; d += 32) P.dir[(size_t)k * P.ndim + d] = __dmul_rn(two_mu, __dsub_rn(cj[d], cl[d]));
  if (lane == 0) {
    const uint4 q = sampler_rand(P.seed, step, (uint32_t)i, 10u + (uint32_t)split);
    const int J = (int)__dmul_rn((double)P.maxsteps, u01(r.z, r.w));     // floor: the product is >= 0
    const double left = -u01(q.z, q.w);
    P.z0[k] = P.lnp[i] + log(u01(q.x, q.y));                             // lnp - Exp(1)
    P.lo[k] = left;
    P.hi[k] = __dadd_rn(left, 1.0);
    P.jbudget[k] = J;
    P.kbudget[k] = P.maxsteps - 1 - J;
    P.phase[k] = 0;
    P.walker_of[k] = i;
  }
}

// One candidate per unfinished walker: X_k + s eta_k with s = L (phase 0), R (phase 1) or a draw from (L, R).
__global__ void __launch_bounds__(128) slice_candidate_kernel(const SliceParams P, unsigned long long step, int split,
                                                              int it) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  int offS, nS, offC, nC;
  split_geometry(P.W, split, offS, nS, offC, nC);
  if (k >= nS) return;
  if (k == 0 && lane == 0) P.ctr->remaining = 0u;     // every block of this iteration's update kernel runs later
  const int ph = P.phase[k];
  if (lane == 0) P.skip[k] = (ph == 3);
  if (ph == 3) return;
  const int i = P.walker_of[k];
  double s;
  if (ph == 0) {
    s = P.lo[k];
  } else if (ph == 1) {
    s = P.hi[k];
  } else {
    const uint4 r = sampler_rand(P.seed, step, (uint32_t)i, 16u + 2u * (uint32_t)it + (uint32_t)split);
    const double L = P.lo[k], R = P.hi[k];
    s = __dadd_rn(L, __dmul_rn(u01(r.x, r.y), __dsub_rn(R, L)));
    if (lane == 0) P.tcur[k] = s;
  }
  const double* x = P.coords + (size_t)i * P.ndim;
  const double* e = P.dir + (size_t)k * P.ndim;
  for (int d = lane; d < P.ndim; d += 32) P.cand[(size_t)k * P.ndim + d] = __dadd_rn(x[d], __dmul_rn(s, e[d]));
}

// Advance every unfinished walker's state machine with the lnprob of its candidate.
__global__ void __launch_bounds__(128) slice_update_kernel(const SliceParams P, int split) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  int offS, nS, offC, nC;
  split_geometry(P.W, split, offS, nS, offC, nC);
  if (k >= nS) return;
  int ph = P.phase[k];
  if (ph == 3) return;
  const double zs = P.lnp_cand[k], z0 = P.z0[k];
  const bool inside = zs >= z0;                        // NaN: outside (and flagged)
  if (ph == 2 && inside) {                             // accepted: the candidate becomes the walker
    const int i = P.walker_of[k];
    for (int d = lane; d < P.ndim; d += 32) P.coords[(size_t)i * P.ndim + d] = P.cand[(size_t)k * P.ndim + d];
    if (lane == 0) P.lnp[i] = zs;
    ph = 3;
  } else if (lane == 0) {
    if (ph == 0) {
      const int b = P.jbudget[k];
      if (inside && b >= 1) {
        P.lo[k] = __dsub_rn(P.lo[k], 1.0);
        P.jbudget[k] = b - 1;
        atomicAdd(&P.ctr->nexp, 1u);
      } else {
        ph = 1;
      }
    } else if (ph == 1) {
      const int b = P.kbudget[k];
      if (inside && b >= 1) {
        P.hi[k] = __dadd_rn(P.hi[k], 1.0);
        P.kbudget[k] = b - 1;
        atomicAdd(&P.ctr->nexp, 1u);
      } else {
        ph = 2;
      }
    } else {
      const double t = P.tcur[k];
      if (t < 0.0) P.lo[k] = t;
      else P.hi[k] = t;
      atomicAdd(&P.ctr->ncon, 1u);
    }
  }
  if (lane == 0) {
    if (zs != zs) atomicOr(P.flag, 1);                 // zeus / emcee: "Probability function returned NaN"
    P.phase[k] = ph;
    if (ph != 3) atomicAdd(&P.ctr->remaining, 1u);
    atomicAdd(&P.ctr->ncall, 1ull);
  }
}

// End of a step: the ensemble and its lnprob become row s of the chain (warp per walker).
__global__ void __launch_bounds__(128) slice_record_kernel(const SliceParams P, double* __restrict__ chain_row,
                                                           double* __restrict__ lnp_row) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= P.W) return;
  if (chain_row)
    for (int d = lane; d < P.ndim; d += 32) chain_row[(size_t)w * P.ndim + d] = P.coords[(size_t)w * P.ndim + d];
  if (lnp_row && lane == 0) lnp_row[w] = P.lnp[w];
}

}  // namespace rbv
