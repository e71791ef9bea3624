// rbvfit_b200 -- streaming form of the lnprob kernel for big batches (included by rbv_kernels.cu).
//
// voigt_tile_kernel gives a CTA one (walker, tile): phase 1 (flux of the tile into shared memory), a CTA-wide
// barrier, phase 2 (LSF + chi^2).  At the headline geometry (8192 walkers x 100 000 px) ncu attributed 1.56 of the
// 9.1 stall cycles per issued instruction to that barrier and another share to the per-CTA prologue / epilogue.
// Here a WARP owns a work item = (walker, contiguous range of output pixels of one instrument) and streams through
// it row by row (256 pixels = 8 per lane) with no CTA-wide synchronisation at all:
//
//   item start   line constants of the walker (written by prep_kernel) and the flipped taps -> the warp's private
//                shared memory (asynchronous copies); the K-1 flux values in front of the first row are evaluated
//                directly only for the FIRST range of a spectrum (edge replication).  Any other range leaves the
//                K-1 outputs that need its predecessor's flux to finalize_stream_kernel: it stores the head of
//                its first row, its predecessor stores its last carry (boundary records), nothing is recomputed
//   every 4 rows phase 0 of voigt_tile_kernel for the next 1024 pixels (prepare_super_chunk: tier lists + far-field
//                record, one lane per line)
//   every row    tau (far-field polynomial + listed lines, tau_wofz<8>) -> exp(-tau) -> private flux buffer
//                [carry K-1 | row 256];  __syncwarp;  LSF for the row's 256 outputs (8 consecutive outputs per lane,
//                sliding register window, same code as phase 2 of the tile kernel) -> chi^2 terms accumulated per
//                lane;  the last K-1 flux values move to the front of the buffer (the next row's carry)
//   item end     warp-shuffle sum of the lanes' chi^2 -> partials[walker, range]; finalize_stream_kernel adds the
//                ranges and the boundary outputs in fixed order and writes lnprob (and applies the sampler's
//                accept/reject when fused)
//
// Items are numbered range-major (all walkers of range 0, then range 1, ...) and handed out through one global
// counter, so the warps resident on an SM work on the same pixels of different walkers at about the same time and
// share the 1/lambda and (flux, inv_sigma2) rows through L1.  The range decomposition depends on the spectra only,
// not on the batch size: a walker's lnprob is bit-identical in every batch that takes this path.
#pragma once

namespace rbv {

#ifndef RBV_STREAM_THREADS
#define RBV_STREAM_THREADS 256
#endif
#ifndef RBV_STREAM_MIN_CTAS
#define RBV_STREAM_MIN_CTAS 2
#endif
constexpr int kStreamThreads = RBV_STREAM_THREADS;
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
  const int list_stride = (L + 3) & ~3;                       // u16 entries per list, 2 lists
  s.total = (s.lists + (2 * list_stride * 2 + 7) / 8 + 1) & ~1;
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
template <int PPT>
__device__ __forceinline__ void stream_general_line(int off, const double (&u)[PPT], double (&tau)[PPT]) {
  const double A = smem[off + LC_A], B = smem[off + LC_B], a2 = smem[off + LC_A2], a = smem[off + LC_a],
               coef = smem[off + LC_COEF];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const double x = fma(A, u[j], -B);
    tau[j] = fma(coef, general_H_call(x, a, fma(x, x, a2)), tau[j]);
  }
}

template <int PPT>
__device__ __forceinline__ void stream_direct_line(int off, const double (&u)[PPT], double (&tau)[PPT],
                                                   const double* __restrict__ core_tab) {
  const double A = smem[off + LC_A], B = smem[off + LC_B], a2 = smem[off + LC_A2];
  double rho[PPT], s[PPT];
  int hm = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const double x = fma(A, u[j], -B);
    s[j] = fma(x, x, a2);                                   // |z|^2
    hm = min(hm, __double2hiint(s[j]));
  }
  hm = __reduce_min_sync(0xffffffffu, hm);                  // NaN: huge (or negative -> core path), propagates
  const bool has_core = hm < kHiCore;
  if (!has_core) {
#pragma unroll
    for (int j = 0; j < PPT; ++j) rho[j] = rcp_pos(s[j]);
  } else {
#pragma unroll
    for (int j = 0; j < PPT; ++j) rho[j] = (__double2hiint(s[j]) < kHiCore) ? 0.0 : rcp_pos(s[j]);
  }
  // series length from the row's smallest |z|^2: 4 coefficients from 200 Doppler widths on (one more than the far
  // tier of the tile kernel: the coefficient pairs then line up with 16-byte loads), 6 from 24, 13 below
  const int qoff = off + LC_Q;
  int p;                                                    // next coefficient pair: Q[p - 1], Q[p]
  if (hm >= kHiNear) {
    const int top = (hm >= kHiFar) ? 2 : 4;                 // pair (Q[top], Q[top + 1])
    const double2 q = *reinterpret_cast<const double2*>(smem + qoff + top);
#pragma unroll
    for (int j = 0; j < PPT; ++j) s[j] = fma(q.y, rho[j], q.x);
    p = top - 1;
  } else {
    const double qtop = smem[qoff + kNQNear - 1];
#pragma unroll
    for (int j = 0; j < PPT; ++j) s[j] = qtop;
    p = kNQNear - 2;
  }
#pragma unroll 1
  for (; p >= 1; p -= 2) {
    const double2 q = *reinterpret_cast<const double2*>(smem + qoff + p - 1);
#pragma unroll
    for (int j = 0; j < PPT; ++j) s[j] = fma(fma(s[j], rho[j], q.y), rho[j], q.x);
  }
#pragma unroll
  for (int j = 0; j < PPT; ++j) tau[j] = fma(s[j], rho[j], tau[j]);
  if (has_core) {
    const double a = smem[off + LC_a], coef = smem[off + LC_COEF];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const double x = fma(A, u[j], -B);
      if (__double2hiint(fma(x, x, a2)) < kHiCore) tau[j] = fma(coef, core_H_call(x, a, a2, core_tab), tau[j]);
    }
  }
}

template <int PPT>
__device__ __forceinline__ void stream_tau_row(int lc_off, const unsigned short* __restrict__ list, int rec_off,
                                               const double (&u)[PPT], double (&tau)[PPT],
                                               const double* __restrict__ core_tab) {
  const int n_direct = *reinterpret_cast<const int*>(smem + rec_off + SC_COUNTS);
  farfield_eval(rec_off, u, tau);       // the record always holds a polynomial (zero when no line qualified)
#pragma unroll 1
  for (int k = 0; k < n_direct; ++k) {
    const int e = list[k];
    const int off = lc_off + (e & 0x7fff) * LC_STRIDE;
    if (e & 0x8000) stream_general_line<PPT>(off, u, tau);
    else stream_direct_line<PPT>(off, u, tau, core_tab);
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

// Phase 0 of the streaming kernel, once per kStreamRowsPerRecord rows (a call): range of 1/lambda of the next
// super-chunk from the block table, then per line (lane = line) far field or direct -- the a-priori gate of
// classify_lines, rbv_kernels.cu, without its tiers (stream_direct_line takes the series length per row) -- then
// the far-field record.  list: directly evaluated lines (bit 15: a >= 1), listff: far-field lines.
__device__ __noinline__ void stream_prepare(const double2* __restrict__ useg, int seg, int L, int lc_off, int rec_off,
                                            unsigned short* __restrict__ list, unsigned short* __restrict__ listff,
                                            bool farfield, int scratch_off, int lane) {
  // range of 1/lambda over the super-chunk (set-up table), widened to the high words: 1/lambda > 0
  const double2 mm = __ldg(useg + seg);
  const double umin = __hiloint2double(__double2hiint(mm.x), 0),
               umax = __hiloint2double(__double2hiint(mm.y), (int)0xffffffff);
  const double du = umax - umin;
  // gate distance: xm >= G_l du^(m / (m + 2)) (fill_fp32_constants); the power through the FP32 special-function
  // unit, rounded up by more than its error
  const float dup = __powf((float)du, (float)RBV_FF_M / (float)(RBV_FF_M + 2)) * 1.001f;
  const unsigned lt = (1u << lane) - 1u;
  int n_dir = 0, n_ff = 0;
  for (int l0 = 0; l0 < L; l0 += 32) {
    const int l = l0 + lane;
    const bool valid = l < L;
    bool ff = false, general = false;
    if (valid) {
      const int off = lc_off + l * LC_STRIDE;
      const double A = smem[off + LC_A], B = smem[off + LC_B], a2 = smem[off + LC_A2];
      const double x1 = fma(A, umin, -B), x2 = fma(A, umax, -B);
      const bool crosses = ((__double2hiint(x1) ^ __double2hiint(x2)) < 0);   // line centre inside the super-chunk
      const int hmin = min(__double2hiint(fma(x1, x1, a2)), __double2hiint(fma(x2, x2, a2)));
      general = __double2hiint(a2) >= 0x3ff00000;                             // a >= 1
      if (farfield && !general && !crosses && hmin >= kHiNear) {               // |z|^2 >= 576: 6-term series exact
        const float xm = fminf(fabsf((float)x1), fabsf((float)x2)) * 0.99999f;     // rounded towards the line
        const float G = reinterpret_cast<const float2*>(smem + off + LC_F32B)->y;
        ff = xm >= G * dup;                        // NaN fails the comparison and stays on the direct path
      }
    }
    const unsigned ff_mask = __ballot_sync(0xffffffffu, valid && ff);
    const unsigned dir_mask = __ballot_sync(0xffffffffu, valid && !ff);
    if (valid) {
      if (ff) listff[n_ff + __popc(ff_mask & lt)] = (unsigned short)l;
      else list[n_dir + __popc(dir_mask & lt)] = (unsigned short)(l | (general ? 0x8000 : 0));
    }
    n_ff += __popc(ff_mask);
    n_dir += __popc(dir_mask);
  }
  __syncwarp();
  if (n_ff > 0) farfield_coefficients(lc_off, listff, n_ff, umin, umax, rec_off, scratch_off, lane);
  else if (lane < SC_COUNTS) smem[rec_off + lane] = 0.0;          // zero polynomial, t = 0
  if (lane == 0) *reinterpret_cast<int*>(smem + rec_off + SC_COUNTS) = n_dir;
}

// LSF of the lane's 8 consecutive outputs: acc[q] = sum_m taps_rev[m] * E[8 lane + q + m] with a sliding register
// window over the padded flux buffer (slot(8 g + j) = 9 g + j).  NB > 0: exactly NB blocks of 8 taps, fully unrolled;
// NB = 0: n_blocks at run time.
template <int NB>
__device__ __forceinline__ void stream_lsf(int fw, int taps_off, int n_blocks, double (&acc)[8]) {
  constexpr int R = 8;
  double win[2 * R - 1];
#pragma unroll
  for (int q = 0; q < R; ++q) acc[q] = 0.0;
#pragma unroll
  for (int q = 0; q < R - 1; ++q) win[q] = smem[fw + q];
  auto block = [&](int blk) {
    const int m0 = blk * R;
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
    fw += R + 1;
  };
  if (NB > 0) {
#pragma unroll
    for (int blk = 0; blk < NB; ++blk) block(blk);
  } else {
#pragma unroll 1
    for (int blk = 0; blk < n_blocks; ++blk) block(blk);
  }
}

#ifdef RBV_STREAM_TIMELINE
// experiments only (tools/stream_timeline.py): per warp {first ticket drawn, last item finished, items, segments}
__device__ unsigned long long g_stream_timeline[4 * 148 * 4 * 8];
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif

// BND = true: range boundaries through boundary records (above).  BND = false: every range evaluates the K-1 flux
// values in front of its first row itself and owns all its outputs -- cheaper when that evaluation is (few lines), and
// it keeps the row loop free of the boundary code for spectra that have a single range (a sightline batch): measured
// against one instantiation for everything, C5a_L4 3.77 -> 3.68 ms, C5b-1024 2.146 -> 2.12 ms.
template <int LOGR, bool BND>
__global__ void __launch_bounds__(kStreamThreads, RBV_STREAM_MIN_CTAS)
voigt_stream_kernel(const __grid_constant__ LaunchParams prm, const int warp_doubles) {
  constexpr int R = 1 << LOGR;
  static_assert(R == 8, "the packed observed-spectrum layout assumes 8 outputs per lane");
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int wbase = (threadIdx.x >> 5) * warp_doubles;
  const unsigned n_items = (unsigned)prm.W * (unsigned)prm.n_tiles;
  unsigned int* queue = prm.tickets + prm.W;   // zeroed by prep_kernel, like the per-walker tickets

#ifdef RBV_STREAM_TIMELINE
  const unsigned long long tl_t0 = global_ns();
  unsigned long long tl_items = 0, tl_segs = 0;
#endif
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

      InstDev I;
      if (prm.inst_in_params) I = prm.inst_v[k];
      else I = prm.inst[k];
      const bool fast = (I.method == RBV_VOIGT_FAST);
      const StreamSmem S = stream_smem_layout(I.L, I.K, I.Kpad);
      const int lc_off = wbase + S.lc, taps_off = wbase + S.taps, flux_off = wbase + S.flux,
                rec_off = wbase + S.rec;
      const int list_stride = (I.L + 3) & ~3;
      unsigned short* s_list = reinterpret_cast<unsigned short*>(smem + wbase + S.lists);
      const int h = I.K >> 1, halo = I.K - 1;
      const int o_lo = (int)prm.range_lo[slot] * kSuperPix, o_hi = min((int)prm.range_hi[slot] * kSuperPix, I.P);

      // ---- item start: line constants, taps, slack, leading flux values
      {
        const double2* src = reinterpret_cast<const double2*>(
            prm.lc + ((size_t)w * prm.n_lines_total + (prm.wps > 0 ? 0 : I.line_base)) * LC_STRIDE);
        // asynchronous copies (LDGSTS): every 16-byte piece is in flight at once -- a register-staged loop keeps one
        // load per lane in flight and cost 6 us of warp time per item (measured through the schedule: 8 more items
        // per walker = +3 % at C5a)
        const unsigned dst = (unsigned)__cvta_generic_to_shared(smem + lc_off);
        for (int i = lane; i < I.L * (LC_STRIDE / 2); i += 32)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * (unsigned)i), "l"(src + i) : "memory");
        const unsigned tdst = (unsigned)__cvta_generic_to_shared(smem + taps_off);
        for (int i = lane; i < I.Kpad; i += 32)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(tdst + 8u * (unsigned)i), "l"(I.taps_rev + i) : "memory");
        if (lane < 16) smem[flux_off + smem_pos(halo + kStreamRow + lane, LOGR)] = 0.0;
        asm volatile("cp.async.wait_all;" ::: "memory");
      }
      __syncwarp();
      // The K-1 flux values in front of the first row.  First range of a spectrum: pixels -h .. h-1 (edge
      // replicated), every line evaluated on its own.  Any other range: they are the previous range's last carry,
      // which another warp computes at some other time -- so the K-1 outputs that need them are left out here
      // (q_first) and added by finalize from the boundary record: the previous range's carry + this range's head.
      // (BND = false: every range is treated like a first one and evaluates them itself.)
      const bool first_range = !BND || (slot == prm.geom[prm.wps > 0 ? 0 : k].first_tile);
      if (first_range) {
        for (int i = lane; i < halo; i += 32) {
          const int p = min(max(o_lo - h + i, 0), I.P - 1);
          const double tau = tau_direct_pixel(lc_off, I.L, fast, __ldg(I.inv_wave + p), prm.core_tab);
          smem[flux_off + smem_pos(i, LOGR)] = exp_flux(-tau);
        }
      } else {
        for (int i = lane; i < halo; i += 32) smem[flux_off + smem_pos(i, LOGR)] = 0.0;
      }
      int q_first = first_range ? 0 : halo;      // outputs of the first row in front of this index: finalize's
      double part = 0.0;
      const int n_rows = (o_hi - o_lo + kStreamRow - 1) / kStreamRow;
      double* fo = smem + flux_off + smem_pos(halo + lane, LOGR);     // slot(halo + 32 j + lane) = fo[36 j]
      constexpr int kRowStep = (R + 1) * (32 / R);
      auto load_u = [&](int pf_row, double (&u)[8]) {
        if (pf_row + kStreamRow <= I.P) {
          const double* up = I.inv_wave + pf_row + lane;
#pragma unroll
          for (int j = 0; j < 8; ++j) u[j] = __ldg(up + j * 32);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) u[j] = __ldg(I.inv_wave + min(pf_row + j * 32 + lane, I.P - 1));   // edge replication
        }
      };
      for (int r = 0; r < n_rows; ++r) {
        const int pf = o_lo + h + r * kStreamRow;      // first pixel of the row's new flux values
        if (!fast && (r % kStreamRowsPerRecord) == 0) {
          stream_prepare(I.useg, (o_lo + r * kStreamRow) / kSuperPix, I.L, lc_off, rec_off, s_list, s_list + list_stride,
                         prm.farfield != 0, wbase + S.scratch, lane);
          __syncwarp();
        }
        // ---- phase 1: 8 pixels per lane
        double u[8], tau[8];
        load_u(pf, u);
        if (fast) tau_fast<8>(lc_off, I.L, u, tau);
        else stream_tau_row<8>(lc_off, s_list, rec_off, u, tau, prm.core_tab);
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
        if (BND && q_first > 0) {   // first row of a range that has a predecessor: its head -> boundary record
          double* bnd_rec = prm.bnd + ((size_t)w * prm.n_tiles + slot) * prm.bnd_stride;
          for (int i = lane; i < halo; i += 32) bnd_rec[halo + i] = smem[flux_off + smem_pos(halo + i, LOGR)];
        }
        // ---- phase 2: M_p = sum_m taps_rev[m] * E[o + m] for the lane's outputs o = 8 lane .. 8 lane + 7
        double acc[R];
        {
          const int fw = flux_off + (R + 1) * lane;
          const int n_blocks = I.Kpad >> LOGR;
          // the common LSF sizes (K <= 8, 16, 24 taps) fully unrolled: every load of the window is issued up front
          // and the window shift is register renaming; anything longer loops over 8-tap blocks
          if (n_blocks == 3) stream_lsf<3>(fw, taps_off, 3, acc);
          else if (n_blocks == 2) stream_lsf<2>(fw, taps_off, 2, acc);
          else if (n_blocks == 1) stream_lsf<1>(fw, taps_off, 1, acc);
          else stream_lsf<0>(fw, taps_off, n_blocks, acc);
        }
        const int n_out = o_hi - o_row;     // outputs of this row (>= 256 except in the last row)
        const int e0 = lane << LOGR;
        if (e0 + R <= n_out && (!BND || e0 >= q_first)) {
#pragma unroll
          for (int q = 0; q < R; ++q) {
            const double resid = obs[q] - acc[q];                   // vfit_mcmc.py:310
            part = fma(resid * resid, wgt[q], part);
          }
        } else {
#pragma unroll
          for (int q = 0; q < R; ++q) {
            if (e0 + q < n_out && (!BND || e0 + q >= q_first)) {
              const double resid = obs[q] - acc[q];
              part = fma(resid * resid, wgt[q], part);
            }
          }
        }
        if (BND) q_first = 0;
        __syncwarp();
        // the row's last K-1 flux values become the next row's carry (element i <- i + 256: slot + 288)
        for (int i = lane; i < halo; i += 32) {
          const int sl = flux_off + smem_pos(i, LOGR);
          smem[sl] = smem[sl + kStreamRow + (kStreamRow >> LOGR)];
        }
        __syncwarp();
      }
      if (BND && o_hi < I.P) {   // the last carry = the K-1 flux values in front of the next range's first output
        double* bnd_next = prm.bnd + ((size_t)w * prm.n_tiles + slot + 1) * prm.bnd_stride;
        for (int i = lane; i < halo; i += 32) bnd_next[i] = smem[flux_off + smem_pos(i, LOGR)];
      }
      part = warp_sum(part);
      if (lane == 0) prm.partials[(size_t)w * prm.n_tiles + slot] = part;
#ifdef RBV_STREAM_TIMELINE
      tl_items++;
      tl_segs += (unsigned long long)(prm.range_hi[slot] - prm.range_lo[slot]);
#endif
    }
    next = __shfl_sync(kFull, next, 0);
  }
#ifdef RBV_STREAM_TIMELINE
  if (lane == 0) {
    const unsigned gw = blockIdx.x * (kStreamThreads / 32) + (threadIdx.x >> 5);
    if (gw < 148 * 4 * 8) {
      g_stream_timeline[4 * gw] = tl_t0;
      g_stream_timeline[4 * gw + 1] = global_ns();
      g_stream_timeline[4 * gw + 2] = tl_items;
      g_stream_timeline[4 * gw + 3] = tl_segs;
    }
  }
#endif
}

}  // namespace rbv
