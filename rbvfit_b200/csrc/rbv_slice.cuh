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
//   after the step: mu <- mu * 2 n_exp / (n_exp + n_con) while tuning is on (slice_record_kernel)
//
// Here every walker of the half is a small state machine (widening: either end of the bracket still open;
// shrinking; finished) and one ITERATION advances every unfinished walker by one step of it:
//   slice_candidate_kernel  ->  the lnprob launch over the half (masked rows are skipped through row_skip)
//                           ->  slice_update_kernel
// While a walker widens, BOTH ends of its bracket are evaluated in the same iteration (row k = X + L eta, row
// n_S + k = X + R eta; the two sides have independent budgets, so this is zeus's left-then-right loop run
// concurrently); while it shrinks only row k is live.  The number of device batches per half-step is therefore the
// longest chain  max(n_L, n_R) + 1 + n_shrink  over the walkers, where the host sampler pays the sum of the longest
// chains of its three loops.  Random numbers come from the same counter-based Philox streams as the stretch move:
// purpose 8 + split (partners, budget), 10 + split (slice level, bracket), 16 + 2 it + split (the shrink draw of
// iteration it).  All floating-point steps that decide the chain are written with explicitly rounded operations so
// that oracle/slice_replay.py reproduces them bit for bit in numpy.
//
// Who drives the loop.  Everything the loop needs -- step, iteration index, mu and its adaptation state, the count of
// unfinished walkers -- lives in one device struct (SliceCounters).  Graph mode: the iteration is the body of a CUDA
// graph WHILE node and the last warp of slice_update_kernel sets the loop condition (cudaGraphSetConditional), so a
// half-step is ONE graph launch and a whole run is enqueued without a single host synchronisation.  Polling mode
// (no graph): the host enqueues iterations one ahead of an asynchronous read-back of the counters.
#pragma once

#include "rbv_sampler.cuh"

namespace rbv {

struct SliceCounters {          // device; the loop state of rbv_slice_run
  unsigned int remaining;       // walkers of the half-step that are not finished after the last update
  unsigned int nexp, ncon;      // expansions / contractions of the current step
  unsigned int it;              // iteration index within the half-step
  unsigned int ticket;          // warps of slice_update_kernel that have finished the current iteration
  unsigned int guard;           // iterations started in this half-step (second, independent loop bound)
  int error;                    // 1 = a half-step needed more than maxiter iterations
  unsigned long long step;      // global index of the step being sampled
  unsigned long long ncall;     // likelihood rows evaluated in this run
  unsigned long long batches;   // lnprob launches in this run
  unsigned long long total_exp, total_con;   // expansions / contractions of the finished steps of this run
  double mu;                    // length scale of the directions
  int tune, good;               // adaptation state (zeus: tune, count of steps inside the tolerance)
};

// walker state: bit 0 = left end still open, bit 1 = right end still open (widening while either is set)
constexpr int kSliceLeft = 1, kSliceRight = 2, kSliceShrink = 4, kSliceDone = 8;

struct SliceParams {
  double* coords;        // [W, ndim] current ensemble (in/out)
  double* lnp;           // [W]       its log-probabilities (in/out)
  double* cand;          // [2 h, ndim] candidates of the active half, h = ceil(W/2): row k and row n_S + k
  double* lnp_cand;      // [2 h]
  double* dir;           // [h, ndim] directions
  double* z0;            // [h] slice levels
  double* lo;            // [h] bracket
  double* hi;            // [h]
  double* tcur;          // [h] the shrink draw that produced the current candidate
  int* jbudget;          // [h] expansions left on the left / right side
  int* kbudget;          // [h]
  int* phase;            // [h] kSlice* state
  int* skip;             // [2 h] row_skip of the lnprob launch (1 = no candidate in this row)
  int* walker_of;        // [h] walker index of row k
  int* flag;             // bit 0: a candidate's lnprob was NaN
  SliceCounters* ctr;
  unsigned long long seed;
  double tolerance;
  int W, ndim, maxsteps, maxiter, patience;
};

// Start of a half-step: direction, slice level, bracket and budgets of every walker of the half (warp per row).
__global__ void __launch_bounds__(128) slice_begin_kernel(const SliceParams P, unsigned long long step, int split) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  int offS, nS, offC, nC;
  split_geometry(P.W, split, offS, nS, offC, nC);
  if (k >= nS) return;
  if (k == 0 && lane == 0) {       // loop state of this half-step; the iteration kernels are later launches
    if (split == 0) {
      P.ctr->nexp = 0u;
      P.ctr->ncon = 0u;
    }
    P.ctr->step = step;
    P.ctr->it = 0u;
    P.ctr->ticket = 0u;
    P.ctr->guard = 0u;
  }
  uint32_t pa, pb;
  sampler_perm(P.seed, P.W, step, pa, pb);
  const int i = walker_at(pa, pb, P.W, offS + k);
  const uint4 r = sampler_rand(P.seed, step, (uint32_t)i, 8u + (uint32_t)split);
  const int j = (int)(r.x % (uint32_t)nC);
  const int l = (int)(((uint32_t)j + 1u + r.y % (uint32_t)(nC - 1)) % (uint32_t)nC);   // l != j
  const double* cj = P.coords + (size_t)walker_at(pa, pb, P.W, offC + j) * P.ndim;
  const double* cl = P.coords + (size_t)walker_at(pa, pb, P.W, offC + l) * P.ndim;
  const double two_mu = 2.0 * P.ctr->mu;                // written by the previous step's slice_record_kernel
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
    P.phase[k] = kSliceLeft | kSliceRight;
    P.walker_of[k] = i;
  }
}

// Candidates of the current iteration: widening walkers put X + L eta into row k and X + R eta into row n_S + k (open
// ends only), shrinking walkers X + t eta with t drawn from (L, R) into row k; every other row is masked.
// `loop` is the WHILE node's handle in graph mode (use_loop != 0): should slice_update_kernel ever fail to end the
// loop, the iteration count kept HERE ends it.
__global__ void __launch_bounds__(128) slice_candidate_kernel(const SliceParams P, int split,
                                                              cudaGraphConditionalHandle loop, int use_loop) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  int offS, nS, offC, nC;
  split_geometry(P.W, split, offS, nS, offC, nC);
  if (k >= nS) return;
  const unsigned long long step = P.ctr->step;
  const unsigned int it = P.ctr->it;                  // both written by earlier launches
  if (k == 0 && lane == 0) {      // every block of this iteration's update kernel runs later
    P.ctr->remaining = 0u;
    const unsigned int g = P.ctr->guard + 1u;
    P.ctr->guard = g;
    if (g > (unsigned)P.maxiter + 8u) {
      P.ctr->error = 1;
      if (use_loop) cudaGraphSetConditional(loop, 0u);
    }
  }
  const int ph = P.phase[k];
  const bool rowA = (ph & (kSliceLeft | kSliceShrink)) != 0, rowB = (ph & kSliceRight) != 0;
  if (lane == 0) {
    P.skip[k] = !rowA;
    P.skip[nS + k] = !rowB;
  }
  if (!rowA && !rowB) return;
  const int i = P.walker_of[k];
  const double* x = P.coords + (size_t)i * P.ndim;
  const double* e = P.dir + (size_t)k * P.ndim;
  if (rowA) {
    double s = P.lo[k];
    if (ph & kSliceShrink) {
      const uint4 r = sampler_rand(P.seed, step, (uint32_t)i, 16u + 2u * it + (uint32_t)split);
      s = __dadd_rn(s, __dmul_rn(u01(r.x, r.y), __dsub_rn(P.hi[k], s)));
      if (lane == 0) P.tcur[k] = s;
    }
    for (int d = lane; d < P.ndim; d += 32) P.cand[(size_t)k * P.ndim + d] = __dadd_rn(x[d], __dmul_rn(s, e[d]));
  }
  if (rowB) {
    const double s = P.hi[k];
    for (int d = lane; d < P.ndim; d += 32) P.cand[(size_t)(nS + k) * P.ndim + d] = __dadd_rn(x[d], __dmul_rn(s, e[d]));
  }
}

// Advance every unfinished walker's state machine with the lnprob of its candidate(s).  The last warp to finish
// closes the iteration: iteration index, batch count and -- in graph mode -- the WHILE node's condition.
__global__ void __launch_bounds__(128) slice_update_kernel(const SliceParams P, int split,
                                                           cudaGraphConditionalHandle loop, int use_loop) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  int offS, nS, offC, nC;
  split_geometry(P.W, split, offS, nS, offC, nC);
  if (k >= nS) return;
  int ph = P.phase[k];
  if (!(ph & kSliceDone)) {
    const double z0 = P.z0[k];
    if (ph & kSliceShrink) {
      const double zs = P.lnp_cand[k];
      if (zs >= z0) {                                    // accepted: the candidate becomes the walker
        const int i = P.walker_of[k];
        for (int d = lane; d < P.ndim; d += 32) P.coords[(size_t)i * P.ndim + d] = P.cand[(size_t)k * P.ndim + d];
        if (lane == 0) P.lnp[i] = zs;
        ph = kSliceDone;
      } else if (lane == 0) {                            // NaN: outside (and flagged)
        const double t = P.tcur[k];
        if (t < 0.0) P.lo[k] = t;
        else P.hi[k] = t;
        atomicAdd(&P.ctr->ncon, 1u);
      }
      if (lane == 0) {
        if (zs != zs) atomicOr(P.flag, 1);               // zeus / emcee: "Probability function returned NaN"
        atomicAdd(&P.ctr->ncall, 1ull);
      }
    } else if (lane == 0) {
      unsigned int widened = 0u, evaluated = 0u;
      if (ph & kSliceLeft) {
        const double zs = P.lnp_cand[k];
        const int b = P.jbudget[k];
        if (zs != zs) atomicOr(P.flag, 1);
        if (zs >= z0 && b >= 1) {
          P.lo[k] = __dsub_rn(P.lo[k], 1.0);
          P.jbudget[k] = b - 1;
          ++widened;
        } else {
          ph &= ~kSliceLeft;
        }
        ++evaluated;
      }
      if (ph & kSliceRight) {
        const double zs = P.lnp_cand[nS + k];
        const int b = P.kbudget[k];
        if (zs != zs) atomicOr(P.flag, 1);
        if (zs >= z0 && b >= 1) {
          P.hi[k] = __dadd_rn(P.hi[k], 1.0);
          P.kbudget[k] = b - 1;
          ++widened;
        } else {
          ph &= ~kSliceRight;
        }
        ++evaluated;
      }
      if (widened) atomicAdd(&P.ctr->nexp, widened);
      atomicAdd(&P.ctr->ncall, (unsigned long long)evaluated);
      if (ph == 0) ph = kSliceShrink;
    }
    if (lane == 0) {
      P.phase[k] = ph;
      if (!(ph & kSliceDone)) atomicAdd(&P.ctr->remaining, 1u);
    }
  }
  __syncwarp();
  if (lane == 0) {
    __threadfence();
    if (atomicAdd(&P.ctr->ticket, 1u) == (unsigned)nS - 1u) {      // every row of the half has been updated
      P.ctr->ticket = 0u;
      const unsigned int rem = atomicAdd(&P.ctr->remaining, 0u);
      const unsigned int it = P.ctr->it + 1u;
      P.ctr->it = it;
      P.ctr->batches += 1ull;
      bool go = rem > 0u;
      if (go && it > (unsigned)P.maxiter) {                         // zeus: "Number of contractions exceeded ..."
        P.ctr->error = 1;
        go = false;
      }
      if (use_loop) cudaGraphSetConditional(loop, go ? 1u : 0u);
      __threadfence();
    }
  }
}

// End of a step: the ensemble and its lnprob become row s of the chain (warp per walker); one thread adapts mu
// (zeus: stochastic approximation towards an expansion fraction of 1/2) for the next step.
__global__ void __launch_bounds__(128) slice_record_kernel(const SliceParams P, double* __restrict__ chain_row,
                                                           double* __restrict__ lnp_row, double* __restrict__ mu_out) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= P.W) return;
  if (chain_row)
    for (int d = lane; d < P.ndim; d += 32) chain_row[(size_t)w * P.ndim + d] = P.coords[(size_t)w * P.ndim + d];
  if (lnp_row && lane == 0) lnp_row[w] = P.lnp[w];
  if (w == 0 && lane == 0) {
    SliceCounters& c = *P.ctr;
    c.total_exp += c.nexp;
    c.total_con += c.ncon;
    if (c.tune) {
      const double ne = (double)max(c.nexp, 1u), tot = ne + (double)c.ncon;
      c.mu = __dmul_rn(c.mu, __ddiv_rn(__dmul_rn(2.0, ne), tot));
      if (fabs(__ddiv_rn(ne, tot) - 0.5) < P.tolerance) c.good += 1;
      if (c.good > P.patience) c.tune = 0;
    }
    if (mu_out) *mu_out = c.mu;
  }
}

}  // namespace rbv
