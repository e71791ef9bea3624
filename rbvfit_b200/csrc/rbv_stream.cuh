// rbvfit_b200 -- streaming form of the lnprob kernel for big batches (included by rbv_kernels.cu).
//
// voigt_tile_kernel gives a CTA one (walker, tile): phase 1 (flux of the tile into shared memory), a CTA-wide
// barrier, phase 2 (LSF + chi^2).  At the headline geometry (8192 walkers x 100 000 px) ncu attributed 1.56 of the
// 9.1 stall cycles per issued instruction to that barrier and another share to the per-CTA prologue / epilogue.
// Here a WARP owns a work item = (walker, contiguous range of output pixels of one instrument) and streams through
// it row by row (256 pixels = 8 per lane) with no CTA-wide synchronisation at all:
//
//   item start   line constants of the walker (written by prep_kernel) and the flipped taps -> the warp's private
//                shared memory; the K-1 flux values in front of the first row are evaluated directly (every line,
//                tier chosen per pixel) -- the only recomputation between neighbouring ranges
//   every 4 rows phase 0 of voigt_tile_kernel for the next 1024 pixels (prepare_super_chunk: tier lists + far-field
//                record, one lane per line)
//   every row    tau (far-field polynomial + listed lines, tau_wofz<8>) -> exp(-tau) -> private flux buffer
//                [carry K-1 | row 256];  __syncwarp;  LSF for the row's 256 outputs (8 consecutive outputs per lane,
//                sliding register window, same code as phase 2 of the tile kernel) -> chi^2 terms accumulated per
//                lane;  the last K-1 flux values move to the front of the buffer (the next row's carry)
//   item end     warp-shuffle sum of the lanes' chi^2 -> partials[walker, range]; finalize_kernel adds the ranges
//                in fixed order and writes lnprob (and applies the sampler's accept/reject when fused)
//
// Items are numbered range-major (all walkers of range 0, then range 1, ...) and handed out through one global
// counter, so the warps resident on an SM work on the same pixels of different walkers at about the same time and
// share the 1/lambda and (flux, inv_sigma2) rows through L1.  The range decomposition depends on the spectrum and
// the LSF only, not on the batch size: a walker's lnprob is bit-identical in every batch that takes this path.
#pragma once

namespace rbv {

#ifndef RBV_STREAM_LSF_UNROLL
#define RBV_STREAM_LSF_UNROLL 1
#endif
constexpr int kStreamLsfUnroll = RBV_STREAM_LSF_UNROLL;
constexpr int kStreamRow = 256;        // pixels per row (8 per lane)
constexpr int kStreamRowsPerRecord = kSuperPix / kStreamRow;

// per-warp shared-memory layout (offsets in doubles from the warp's base; all even)
struct StreamSmem {
  int lc, taps, flux, rec, lists, total, scratch;
};
__host__ __device__ inline StreamSmem stream_smem_layout(int L, int K, int Kpad) {
  StreamSmem s;
  s.lc = 0;
  s.taps = L * LC_STRIDE;
  s.flux = s.taps + Kpad;
  const int n_el = (K - 1) + kStreamRow + 16;                 // carry | row | slack of the register window
  const int slots = n_el + (n_el >> 3) + 2;
  s.scratch = s.flux + (((K - 1) + ((K - 1) >> 3) + 1) & ~1);  // phase-0 transpose scratch: the still empty row
  s.rec = s.flux + ((slots + 1) & ~1);
  s.lists = s.rec + SC_STRIDE;
  const int list_stride = (L + 3) & ~3;                       // u16 entries per list, 3 lists
  s.total = (s.lists + (3 * list_stride * 2 + 7) / 8 + 1) & ~1;
  return s;
}

// tau of ONE pixel with every line evaluated on its own, tier chosen per pixel (item start only)
__device__ __noinline__ double tau_direct_pixel(int lc_off, int L, bool fast, double u,
                                                const double* __restrict__ core_tab) {
  double tau = 0.0;
  for (int l = 0; l < L; ++l) {
    const int off = lc_off + l * LC_STRIDE;
    const double A = smem[off + LC_A], B = smem[off + LC_B], a2 = smem[off + LC_A2];
    const double x = fma(A, u, -B);
    if (fast) {
      tau = fma(smem[off + LC_COEF], tg_H(x, smem[off + LC_a], a2, smem[off + LC_AUX]), tau);
      continue;
    }
    const double d = fma(x, x, a2);
    double t;
    if (__double2hiint(a2) >= 0x3ff00000) t = smem[off + LC_COEF] * general_H(x, smem[off + LC_a], d);
    else if (d >= kDFar) t = asym_series<kNQFar>(smem + off + LC_Q, d);
    else if (d >= kDNear) t = asym_series<kNQMid>(smem + off + LC_Q, d);
    else if (d >= kDCore) t = asym_series<kNQNear>(smem + off + LC_Q, d);
    else t = smem[off + LC_COEF] * core_H(x, smem[off + LC_a], a2, core_tab);   // NaN lands here and propagates
    tau += t;
  }
  return tau;
}


// ---- phase 1 of one row, written for a small instruction footprint ---------------------------------------------
// The warps of an SM run this kernel out of step with each other, so the union of what they execute has to fit the
// 32 KB instruction cache of the SM (ncu on the first version, which inlined tau_wofz<8> with its per-tier unrolled
// bodies: 2.2 of 9 stall cycles per issued instruction were instruction fetch).  One loop body serves every directly
// evaluated line: the number of series coefficients is chosen per (line, row) from the row's smallest |z|^2 (a warp
// reduction of the high words -- a line that is 'core' for its 1024-pixel super-chunk is far-wing for most of its
// rows), pixels inside the core get rho = 0 in the series and the core evaluation in a second pass; rare paths
// (core tables, Weideman, a >= 1, exp with range reduction) are calls.
__device__ __noinline__ double core_H_call(double x, double a, double a2, const double* __restrict__ tab) {
  return core_H(x, a, a2, tab);
}

__device__ __noinline__ double general_H_call(double x, double a, double d) { return general_H(x, a, d); }

// a >= 1 (unphysical damping, correctness only); by value into the call: u / tau stay in registers
__device__ __forceinline__ void stream_general_line(int off, const double (&u)[8], double (&tau)[8]) {
  const double A = smem[off + LC_A], B = smem[off + LC_B], a2 = smem[off + LC_A2], a = smem[off + LC_a],
               coef = smem[off + LC_COEF];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double x = fma(A, u[j], -B);
    tau[j] = fma(coef, general_H_call(x, a, fma(x, x, a2)), tau[j]);
  }
}

__device__ __forceinline__ void stream_direct_line(int off, const double (&u)[8], double (&tau)[8],
                                                   const double* __restrict__ core_tab) {
  const double A = smem[off + LC_A], B = smem[off + LC_B], a2 = smem[off + LC_A2];
  double rho[8], s[8];
  int hm = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double x = fma(A, u[j], -B);
    s[j] = fma(x, x, a2);                                   // |z|^2
    hm = min(hm, __double2hiint(s[j]));
  }
  hm = __reduce_min_sync(0xffffffffu, hm);                  // NaN: huge (or negative -> core path), propagates
  const bool has_core = hm < kHiCore;
  if (!has_core) {
#pragma unroll
    for (int j = 0; j < 8; ++j) rho[j] = rcp_pos(s[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) rho[j] = (__double2hiint(s[j]) < kHiCore) ? 0.0 : rcp_pos(s[j]);
  }
  const int nq = (hm >= kHiFar) ? kNQFar : (hm >= kHiNear) ? kNQMid : kNQNear;
  const int qoff = off + LC_Q;
  {
    const double qtop = smem[qoff + nq - 1];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = qtop;
  }
#pragma unroll 1
  for (int p = nq - 2; p >= 0; --p) {
    const double q = smem[qoff + p];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = fma(s[j], rho[j], q);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) tau[j] = fma(s[j], rho[j], tau[j]);
  if (has_core) {
    const double a = smem[off + LC_a], coef = smem[off + LC_COEF];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const double x = fma(A, u[j], -B);
      if (__double2hiint(fma(x, x, a2)) < kHiCore) tau[j] = fma(coef, core_H_call(x, a, a2, core_tab), tau[j]);
    }
  }
}

__device__ __forceinline__ void stream_tau_row(int lc_off, int L, const unsigned short* __restrict__ list,
                                               int rec_off, const double (&u)[8], double (&tau)[8],
                                               const double* __restrict__ core_tab) {
  const int4 n = *reinterpret_cast<const int4*>(smem + rec_off + SC_COUNTS);   // n_far, n_other, -, n_farfield
  farfield_eval(rec_off, u, tau);       // the record always holds a polynomial (zero when no line qualified)
  const int n_direct = n.x + n.y;
#pragma unroll 1
  for (int k = 0; k < n_direct; ++k) {
    const int e = (k < n.x) ? (int)list[k] : (int)list[L - 1 - (k - n.x)];
    const int off = lc_off + (e & 0xfff) * LC_STRIDE;
    if ((e >> 12) == kTierGeneral) stream_general_line(off, u, tau);
    else stream_direct_line(off, u, tau, core_tab);
  }
}

__device__ __noinline__ void stream_exp_store(double* fo, int step, double t0, double t1, double t2, double t3,
                                              double t4, double t5, double t6, double t7) {
  fo[0] = exp_flux(-t0);
  fo[step] = exp_flux(-t1);
  fo[2 * step] = exp_flux(-t2);
  fo[3 * step] = exp_flux(-t3);
  fo[4 * step] = exp_flux(-t4);
  fo[5 * step] = exp_flux(-t5);
  fo[6 * step] = exp_flux(-t6);
  fo[7 * step] = exp_flux(-t7);
}

// phase 0 as a call (once per kStreamRowsPerRecord rows)
__device__ __noinline__ void stream_prepare(const double2* __restrict__ ublk, int L, int lc_off, int rec_off,
                                            unsigned short* list, int list_stride, float ff_eps, int plo, int phi,
                                            int scratch_off, int lane) {
  InstDev I;
  I.ublk = ublk;
  I.L = L;
  prepare_super_chunk(I, lc_off, rec_off, list, list + list_stride, list + 2 * list_stride, 0.0, ff_eps, plo, phi,
                      scratch_off, lane);
}

template <int LOGR>
__global__ void __launch_bounds__(kThreads, RBV_MIN_CTAS)
voigt_stream_kernel(const __grid_constant__ LaunchParams prm, const int warp_doubles) {
  constexpr int R = 1 << LOGR;
  static_assert(R == 8, "the packed observed-spectrum layout assumes 8 outputs per lane");
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int wbase = (threadIdx.x >> 5) * warp_doubles;
  const unsigned n_items = (unsigned)prm.W * (unsigned)prm.n_tiles;
  unsigned int* queue = prm.tickets;       // tickets[0]: zeroed by prep_kernel, unused otherwise on this path

  unsigned next = 0u;
  if (lane == 0) next = atomicAdd(queue, 1u);
  next = __shfl_sync(kFull, next, 0);
  while (next < n_items) {
    const unsigned id = next;
    if (lane == 0) next = atomicAdd(queue, 1u);     // the round trip overlaps this item
    const int slot = (int)(id / (unsigned)prm.W);
    const int w = (int)(id - (unsigned)slot * (unsigned)prm.W);
    if (!prm.oob[w]) {
      int k = 0;
      if (prm.wps > 0) {
        k = w / prm.wps;
      } else {
        while (k + 1 < prm.n_inst && slot >= prm.geom[k + 1].first_tile) ++k;
      }
      const int gk = prm.wps > 0 ? 0 : k;
      const int range_len = prm.geom[gk].tile, first_slot = prm.geom[gk].first_tile;
      InstDev I;
      if (prm.inst_in_params) I = prm.inst_v[k];
      else I = prm.inst[k];
      const bool fast = (I.method == RBV_VOIGT_FAST);
      const StreamSmem S = stream_smem_layout(I.L, I.K, I.Kpad);
      const int lc_off = wbase + S.lc, taps_off = wbase + S.taps, flux_off = wbase + S.flux,
                rec_off = wbase + S.rec;
      const int list_stride = (I.L + 3) & ~3;
      unsigned short* s_list = reinterpret_cast<unsigned short*>(smem + wbase + S.lists);
      const float ff_eps = prm.farfield ? (float)(kFFEps / (double)I.L) : 0.f;
      const int h = I.K >> 1, halo = I.K - 1;
      const int o_lo = (slot - first_slot) * range_len, o_hi = min(o_lo + range_len, I.P);

      // ---- item start: line constants, taps, slack, leading flux values
      {
        const double2* src = reinterpret_cast<const double2*>(
            prm.lc + ((size_t)w * prm.n_lines_total + (prm.wps > 0 ? 0 : I.line_base)) * LC_STRIDE);
        double2* dst = reinterpret_cast<double2*>(smem + lc_off);
        for (int i = lane; i < I.L * (LC_STRIDE / 2); i += 32) dst[i] = src[i];
        for (int i = lane; i < I.Kpad; i += 32) smem[taps_off + i] = __ldg(I.taps_rev + i);
        if (lane < 16) smem[flux_off + smem_pos(halo + kStreamRow + lane, LOGR)] = 0.0;
      }
      __syncwarp();
      for (int i = lane; i < halo; i += 32) {
        const int p = min(max(o_lo - h + i, 0), I.P - 1);
        const double tau = tau_direct_pixel(lc_off, I.L, fast, __ldg(I.inv_wave + p), prm.core_tab);
        smem[flux_off + smem_pos(i, LOGR)] = exp_flux(-tau);
      }

      double part = 0.0;
      const int n_rows = (o_hi - o_lo + kStreamRow - 1) / kStreamRow;
      double* fo = smem + flux_off + smem_pos(halo + lane, LOGR);     // slot(halo + 32 j + lane) = fo[36 j]
      constexpr int kRowStep = (R + 1) * (32 / R);
      for (int r = 0; r < n_rows; ++r) {
        const int pf = o_lo + h + r * kStreamRow;      // first pixel of the row's new flux values
        if (!fast && (r % kStreamRowsPerRecord) == 0) {
          const int plo = min(max(pf, 0), I.P - 1);
          const int phi = min(max(pf + kSuperPix - 1, 0), I.P - 1);
          stream_prepare(I.ublk, I.L, lc_off, rec_off, s_list, list_stride, ff_eps, plo, phi, wbase + S.scratch, lane);
          __syncwarp();
        }
        // ---- phase 1: 8 pixels per lane
        double u[8], tau[8];
        if (pf + kStreamRow <= I.P) {
          const double* up = I.inv_wave + pf + lane;
#pragma unroll
          for (int j = 0; j < 8; ++j) u[j] = __ldg(up + j * 32);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) u[j] = __ldg(I.inv_wave + min(pf + j * 32 + lane, I.P - 1));   // edge replication
        }
        if (fast) tau_fast<8>(lc_off, I.L, u, tau);
        else stream_tau_row(lc_off, I.L, s_list, rec_off, u, tau, prm.core_tab);
        unsigned hmax = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) hmax = max(hmax, (unsigned)__double2hiint(tau[j]));
        if (__reduce_max_sync(kFull, hmax) < 0x3F900000u) {
#pragma unroll
          for (int j = 0; j < 8; ++j) fo[j * kRowStep] = exp_small(-tau[j]);
        } else {
          stream_exp_store(fo, kRowStep, tau[0], tau[1], tau[2], tau[3], tau[4], tau[5], tau[6], tau[7]);
        }
        // observed spectrum of the lane's 8 outputs (block-transposed copy: lane-contiguous 16-byte loads)
        const int o_row = o_lo + r * kStreamRow;
        double obs[R], wgt[R];
        {
          const double2* src = I.obs_w + (size_t)(o_row >> 8) * 256 + lane;
#pragma unroll
          for (int q = 0; q < R; ++q) {
            const double2 v = __ldg(src + 32 * q);
            obs[q] = v.x;
            wgt[q] = v.y;
          }
        }
        __syncwarp();
        // ---- phase 2: M_p = sum_m taps_rev[m] * E[o + m] for the lane's outputs o = 8 lane .. 8 lane + 7
        double acc[R], win[2 * R - 1];
#pragma unroll
        for (int q = 0; q < R; ++q) acc[q] = 0.0;
        int fw = flux_off + (R + 1) * lane;
#pragma unroll
        for (int q = 0; q < R - 1; ++q) win[q] = smem[fw + q];
        const int n_blocks = I.Kpad >> LOGR;
        __builtin_assume(n_blocks >= 1);
#pragma unroll kStreamLsfUnroll
        for (int blk = 0; blk < n_blocks; ++blk, fw += R + 1) {
          const int m0 = blk << LOGR;
          win[R - 1] = smem[fw + R - 1];
#pragma unroll
          for (int q = 1; q < R; ++q) win[R - 1 + q] = smem[fw + R + q];
#pragma unroll
          for (int mm = 0; mm < R; mm += 2) {
            const double2 tap = *reinterpret_cast<const double2*>(smem + taps_off + m0 + mm);
#pragma unroll
            for (int q = 0; q < R; ++q) acc[q] = fma(tap.x, win[mm + q], acc[q]);
#pragma unroll
            for (int q = 0; q < R; ++q) acc[q] = fma(tap.y, win[mm + 1 + q], acc[q]);
          }
#pragma unroll
          for (int q = 0; q < R - 1; ++q) win[q] = win[R + q];
        }
        const int n_out = o_hi - o_row;     // outputs of this row (>= 256 except in the last row)
        const int e0 = lane << LOGR;
        if (e0 + R <= n_out) {
#pragma unroll
          for (int q = 0; q < R; ++q) {
            const double resid = obs[q] - acc[q];                   // vfit_mcmc.py:310
            part = fma(resid * resid, wgt[q], part);
          }
        } else {
#pragma unroll
          for (int q = 0; q < R; ++q) {
            if (e0 + q < n_out) {
              const double resid = obs[q] - acc[q];
              part = fma(resid * resid, wgt[q], part);
            }
          }
        }
        __syncwarp();
        // the row's last K-1 flux values become the next row's carry (element i <- i + 256: slot + 288)
        for (int i = lane; i < halo; i += 32) {
          const int sl = flux_off + smem_pos(i, LOGR);
          smem[sl] = smem[sl + kStreamRow + (kStreamRow >> LOGR)];
        }
        __syncwarp();
      }
      part = warp_sum(part);
      if (lane == 0) prm.partials[(size_t)w * prm.n_tiles + slot] = part;
    }
    next = __shfl_sync(kFull, next, 0);
  }
}

}  // namespace rbv
