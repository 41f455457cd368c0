// The fused walker-tile kernel: psi, local energy, grad psi and the Metropolis step.
//
// One CTA owns TW walkers at a time (persistent, grid-stride over tiles) and stages the
// basis / MO / configuration tables in shared memory once.  Per tile:
//   P0  coalesced load of the walker coordinates (optionally + proposal) -> smem
//   P1  thread (walker, electron): Jastrow gradient/Laplacian terms + potentials
//   P2  thread (walker, MO block, electron): shells -> AO value/grad/lap in registers,
//       contracted on the fly against the MO columns some configuration occupies;
//       B_kin row assembled in registers; only mo and B_kin (or grad mo) go to smem
//   P3  thread (walker, spin, unique occupation): det, inverse / Tr(A^-1 B)
//   P4  thread (walker): CI sum, psi, E_kin, E_L  (or accept/reject + in-place update)
// Nothing but pos in and psi/E_L out touches HBM.
#include <cuda_runtime.h>

#include <cstdio>

#pragma once
#include <cstdlib>

#include "device.cuh"
#include "philox.cuh"

#include "fused_args.h"

template <int V>
struct IntTag { static constexpr int value = V; };

template <int NCH, int MB>
struct MoSink {
  double acc[NCH][MB];
  const double *w;
  int nmup;
  __device__ __forceinline__ void init(const double *w_, int nmup_) {
    w = w_; nmup = nmup_;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int j = 0; j < MB; ++j) acc[c][j] = 0.0;
  }
  __device__ __forceinline__ void emit(int ao, const double (&v)[NCH]) {
    const double *wr = w + ao * nmup;
    if (MB >= 2) {
      // two weights per LDS.128: the rows are 16-byte aligned (nmup and MB are even, o_mow is even)
      const double2 *wr2 = reinterpret_cast<const double2 *>(wr);
#pragma unroll
      for (int j = 0; j < MB / 2; ++j) {
        const double2 wj = wr2[j];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          acc[c][2 * j] = fma(v[c], wj.x, acc[c][2 * j]);
          acc[c][2 * j + 1] = fma(v[c], wj.y, acc[c][2 * j + 1]);
        }
      }
    } else {
      const double wj = wr[0];
#pragma unroll
      for (int c = 0; c < NCH; ++c) acc[c][0] = fma(v[c], wj, acc[c][0]);
    }
  }
};

// shared-memory plan per CTA (doubles), after the tables:
//   spos [TW][3Ne] | jv [TW][Ne][8] | mo [NCHS][TW][Ne][nmup] | dets [TW][nuu+nud] | trs [TW][nuu+nud]
//   | wsum [TW][4] | scratch
template <int MODE>
__host__ __device__ constexpr int nchs() { return MODE == MODE_ELOC ? 2 : (MODE == MODE_GRAD ? 4 : 1); }

// Spin blocks larger than this are factorised by one WARP each (warp_gauss_jordan); smaller ones
// by one THREAD each (many independent small blocks, e.g. CAS expansions).  Measured on B200:
// H2O cas(4,4) (n=5, 12 blocks/walker) 0.96 ms thread-per-block vs 3.4 ms warp-per-block;
// C4H6 (n=15, 2 blocks/walker) 4.4 ms vs 2.1 ms.
#define QMCB_WARP_LU_MIN 7
__host__ __device__ inline bool use_warp_lu(const DevSys &S) {
  return (S.nup > S.ndown ? S.nup : S.ndown) >= QMCB_WARP_LU_MIN;
}

__host__ __device__ inline int mo_row_stride(const DevSys &S) { return S.nmup | 1; }

__host__ __device__ inline int lu_scratch_per_item(const DevSys &S, int mode, bool warp_tiles) {
  const int n = S.nup > S.ndown ? S.nup : S.ndown;
  if (mode == MODE_GRAD) return n <= 3 ? n * n : 2 * n * n;   // inverse kept ([A|I] for n>3)
  if (n <= 3) return 0;                                      // closed forms
  if (!warp_tiles && n <= 16) return 0;                      // register LU (n <= 6) / register warp Gauss-Jordan
  return mode == MODE_ELOC ? 2 * n * n : n * n;            // [A|B] or A
}

// Work area of one tile (doubles): spos [TW][3Ne] | jv [TW][Ne][8] | mo [NCHS][TW][Ne][nmup]
// | dets [TW][nun] | trs [TW][nun] | wsum [TW][4] | LU scratch
__host__ __device__ inline size_t tile_doubles(const DevSys &S, int mode, int tw, int lu_conc, bool warp_tiles) {
  const int nchs_ = mode == MODE_ELOC ? 2 : (mode == MODE_GRAD ? 4 : 1);
  const int nun = S.nuu + S.nud;
  size_t d = (size_t)tw * 3 * S.nelec + (size_t)tw * S.nelec * 8 + (size_t)nchs_ * tw * S.nelec * mo_row_stride(S);
  d += 2 * (size_t)tw * nun + (size_t)tw * 4;
  d += (size_t)lu_conc * lu_scratch_per_item(S, mode, warp_tiles);
  return (d + 1) & ~(size_t)1;
}

// CTAs (256 threads) per SM the compiler must allow for warp-owned tiles.  Measured on B200,
// LiH E_L kernel: 1 -> 0.677 ms, 2 -> 0.466 ms (123 regs, no spills), 3 -> 0.502 ms (80 regs,
// spills), 4 -> 0.535 ms (64 regs).
// Later: 128-thread CTAs x 5 per SM (96 registers, no spills, 20 warps/SM) measured equal on the
// E_L kernel and 3-7 % faster on psi / grad / H2 than 256 x 2.
// private slice of one thread in THREAD tiles: pos | g,lap [4][Ne] (E_L only) | mo (,B) rows | dets | traces | 2
__host__ __device__ inline size_t thread_tile_doubles(const DevSys &S, int mode) {
  const int nchs_ = mode == MODE_ELOC ? 2 : 1;
  const size_t d = (size_t)3 * S.nelec + (mode == MODE_ELOC ? 4 * S.nelec : 0) + (size_t)nchs_ * S.nelec * S.nmup +
                   2 * (size_t)(S.nuu + S.nud) + 2;
  return d | 1;   // odd stride: lanes hit distinct banks
}

#ifndef QMCB_MINBLOCKS
#define QMCB_MINBLOCKS 5
#endif
#ifndef QMCB_WARP_CTA
#define QMCB_WARP_CTA 128      // threads per CTA for warp-owned tiles
#endif

// Tile ownership (template TILE):
//   0 CTA    : a tile of TW walkers belongs to the CTA; phases are separated by __syncthreads.
//   1 WARP   : a tile belongs to ONE WARP (Ne * NBLK divides 32); phases are separated by
//              __syncwarp only, warps never wait for each other.
//   2 THREAD : one walker per THREAD (small systems, closed-form determinants): no synchronisation
//              at all, every electron pair of the Jastrow factor is visited once instead of twice,
//              and the per-walker epilogue runs on all lanes.  Work area: a private, odd-strided
//              slice of shared memory per thread.
#define QMCB_THREAD_CTA 128
// CTA tiles with 16 MO columns per thread (psi / E_L of large systems): 32 accumulators + the shell
// state do not fit the 128 registers that 512 threads allow (ncu: LDL in the projection loop,
// long-scoreboard stalls 1.8 per issue); these run as two CTAs of at most QMCB_MB16_THREADS threads
#ifndef QMCB_MB16_THREADS
#define QMCB_MB16_THREADS 192
#endif
template <int MODE, int MB, int RT, int TILE>
__global__ void __launch_bounds__(TILE == 1 ? QMCB_WARP_CTA : (TILE == 2 ? QMCB_THREAD_CTA : (MB == 16 ? QMCB_MB16_THREADS : 512)),
                                  TILE == 1 ? QMCB_MINBLOCKS : (TILE == 2 ? 4 : (MB == 16 ? 2 : 1)))
    fused_kernel(const DevSys S, const FusedArgs a, const int TW, const int NBLK, const int lu_conc) {
  constexpr bool WARP = TILE == 1;
  constexpr bool THREAD = TILE == 2;
  // AO channels contracted against the MO columns: psi 1 (ao); grad psi 4 (ao + gradient); E_L 2 - ao
  // and the folded kinetic channel lap ao + 2 grad ln J . grad ao + (lap J / J) ao (device.cuh: FoldJ)
  constexpr int NCH = MODE == MODE_ELOC ? 2 : (MODE == MODE_GRAD ? 4 : 1);
  constexpr int NCHS = nchs<MODE>();
  extern __shared__ __align__(16) double smem[];
  Tab T;
  double *ws = stage_tables(S, smem, T);
  const int Ne = S.nelec, ne3 = 3 * Ne, nmup = S.nmup;
  // row stride of the mo / B_kin rows in the work area: odd for shared tiles (threads of a warp own
  // consecutive rows: conflict-free row writes and column reads); private slices keep nmup
  const int ldm = THREAD ? nmup : mo_row_stride(S);
  const int nun = S.nuu + S.nud;
  const int nthr = THREAD ? 1 : (WARP ? 32 : (int)blockDim.x);
  const int tid = THREAD ? 0 : (WARP ? (int)(threadIdx.x & 31) : (int)threadIdx.x);
  const int64_t unit = THREAD ? (int64_t)blockIdx.x * blockDim.x + threadIdx.x
                              : (WARP ? (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5) : blockIdx.x);
  const int64_t nunit = THREAD ? (int64_t)gridDim.x * blockDim.x
                               : (WARP ? (int64_t)gridDim.x * (blockDim.x >> 5) : gridDim.x);
  if (WARP) ws += (threadIdx.x >> 5) * tile_doubles(S, MODE, TW, lu_conc, true);
  if (THREAD) ws += threadIdx.x * thread_tile_doubles(S, MODE);
  double *spos = ws;
  double *jv = spos + TW * ne3;
  double *smo = jv + (THREAD ? (NCH > 1 ? 4 * Ne : 0) : TW * Ne * 8);
  double *sdet = smo + (size_t)NCHS * TW * Ne * ldm;
  double *str = sdet + TW * nun;
  double *wsum = str + TW * nun;
  double *scr = wsum + TW * 4;
  const int64_t ntile = (a.W + TW - 1) / TW;
  const size_t chs = (size_t)TW * Ne * ldm;
  const int jvs = TW * Ne;   // jv is stored [quantity][walker][electron]
  __syncthreads();
#define TILE_SYNC() do { if (THREAD) {} else if (WARP) __syncwarp(); else __syncthreads(); } while (0)

  for (int64_t tile = unit; tile < ntile; tile += nunit) {
    const int64_t w0 = tile * TW;
    const int tw = (int)((a.W - w0) < TW ? (a.W - w0) : TW);
    // ---- P0: coordinates (+ proposal)
    if (MODE == MODE_MH && !a.disp && a.proba_normal) {
      // in-kernel normal proposals: one Philox call yields the four normals of a GLOBAL element
      // quad (4q .. 4q+3), so the draw of an element does not depend on the tiling
      const int64_t g0 = w0 * ne3, g1 = g0 + (int64_t)tw * ne3;
      for (int64_t q = (g0 >> 2) + tid; 4 * q < g1; q += nthr) {
        double z[4];
        philox_normal4(a.seed, a.offset, (uint64_t)q, z);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int64_t g = 4 * q + h;
          if (g < g0 || g >= g1) continue;
          const int i = (int)(g - g0);
          const int wl = i / ne3, e = (i - wl * ne3) / 3;
          int me = a.move_elec;
          if (me == -2) {
            if (a.elec_index) me = a.elec_index[w0 + wl];
            else me = (int)(philox_u32(a.seed, a.offset, (uint64_t)(w0 + wl), 2u) % (unsigned)Ne);
          }
          double v = a.pos[g];
          if (me < 0 || me == e) v += a.scale * z[h];
          spos[i] = v;
        }
      }
    } else {
      for (int i = tid; i < tw * ne3; i += nthr) {
        double v = a.pos[w0 * ne3 + i];
        if (MODE == MODE_MH) {
          const int wl = i / ne3, c = i - wl * ne3, e = c / 3;
          int me = a.move_elec;
          if (me == -2) {
            if (a.elec_index) me = a.elec_index[w0 + wl];
            else me = (int)(philox_u32(a.seed, a.offset, (uint64_t)(w0 + wl), 2u) % (unsigned)Ne);
          }
          if (me < 0 || me == e) {
            double d;
            if (a.disp) d = a.disp[w0 * ne3 + i];
            else d = a.scale * (2.0 * philox_uniform(a.seed, a.offset, (uint64_t)(w0 * ne3 + i), 0u) - 1.0);
            v += d;
          }
        }
        spos[i] = v;
      }
    }
    TILE_SYNC();
    // ---- P1: Jastrow + potentials, thread (wl, e)
    double tks = 0.0, tven = 0.0, tvee = 0.0;   // THREAD tiles: walker totals stay in registers
    if (THREAD) {
      walker_terms<(NCH > 1), (MODE == MODE_ELOC)>(S, T, spos, jv, jvs, tks, tven, tvee);
    } else if (NCH > 1 && S.use_jee && S.een_nterm == 0 && Ne >= 6 && Ne <= 32 && TW <= (nthr >> 5) * (32 / Ne)) {
      // every e-e pair once: the electrons of a walker on consecutive lanes of one warp
      // (electron_terms_paired); one round covers the tile
      const int lane = tid & 31, per = 32 / Ne;
      const int sub = lane / Ne, e = lane - sub * Ne;
      const int wl = (tid >> 5) * per + sub;
      const bool act = sub < per && wl < tw;
      if ((tid >> 5) * per < tw) {           // warp-uniform: this warp owns at least one walker
        ElecTerms o;
        electron_terms_paired<(MODE == MODE_ELOC)>(S, T, spos + (act ? wl : 0) * ne3, act ? e : 0, sub * Ne, act, o);
        if (act) {
          double *q = jv + wl * Ne + e;
          q[0] = o.gx; q[jvs] = o.gy; q[2 * jvs] = o.gz; q[3 * jvs] = o.lap;
          q[4 * jvs] = o.ks; q[5 * jvs] = o.ven; q[6 * jvs] = o.vee;
        }
      }
    } else
    for (int it = tid; it < tw * Ne; it += nthr) {
      const int wl = it / Ne, e = it - wl * Ne;
      ElecTerms o;
      electron_terms<(NCH > 1), (MODE == MODE_ELOC)>(S, T, spos + wl * ne3, e, o);
      double *q = jv + it;
      if (NCH > 1) { q[0] = o.gx; q[jvs] = o.gy; q[2 * jvs] = o.gz; q[3 * jvs] = o.lap; }
      q[4 * jvs] = o.ks; q[5 * jvs] = o.ven; q[6 * jvs] = o.vee;
    }
    if (NCH > 1) TILE_SYNC();   // P2 reads jv only when it assembles B_kin / the gradient
    // ---- P2: AO -> MO rows, thread (wl, blk, e)
    for (int it = tid; it < tw * NBLK * Ne; it += nthr) {
      // (THREAD tiles: one walker, one column block - no integer divisions in the electron loop)
      const int wl = THREAD ? 0 : it / (NBLK * Ne), rem = it - wl * NBLK * Ne;
      const int blk = THREAD ? 0 : rem / Ne, e = rem - blk * Ne;
      const double *sp = spos + wl * ne3 + 3 * e;
      MoSink<NCH, MB> sink;
      sink.init(T.mow() + blk * MB, nmup);
      double *dst = smo + ((size_t)wl * Ne + e) * ldm + blk * MB;
      if constexpr (MODE == MODE_ELOC) {
        // B_kin = -1/2 (lap mo + 2 grad ln J . grad mo + (lap J / J) mo), folded per AO
        const double *q = jv + wl * Ne + e;
        const FoldJ fj{2.0 * q[0], 2.0 * q[jvs], 2.0 * q[2 * jvs], q[3 * jvs]};
        eval_aos<NCH, RT, true>(S, T, sp[0], sp[1], sp[2], sink, fj);
#pragma unroll
        for (int j = 0; j < MB; ++j) {
          dst[j] = sink.acc[0][j];
          dst[chs + j] = -0.5 * sink.acc[1][j];
        }
      } else {
        eval_aos<NCH, RT>(S, T, sp[0], sp[1], sp[2], sink);
        if constexpr (MODE == MODE_GRAD) {
#pragma unroll
          for (int j = 0; j < MB; ++j) {
            dst[j] = sink.acc[0][j];
            dst[chs + j] = sink.acc[1][j];
            dst[2 * chs + j] = sink.acc[2][j];
            dst[3 * chs + j] = sink.acc[3][j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < MB; ++j) dst[j] = sink.acc[0][j];
        }
      }
    }
    TILE_SYNC();
    // ---- P3: determinants (and traces / inverses) per (wl, unique occupation)
    {
      const int nitem = tw * nun;
      const int per = lu_scratch_per_item(S, MODE, TILE != 0);
      if (TILE == 0 && use_warp_lu(S)) {
        // CTA tiles with blocks larger than 3x3: ONE WARP per spin block (warp_gauss_jordan);
        // scratch is contiguous per slot: slot = item (GRAD keeps every inverse) or the warp
        const int warp = tid >> 5, lane = tid & 31, nwarp = nthr >> 5;
        if ((S.nup > S.ndown ? S.nup : S.ndown) <= 16) {
          // register-resident, TWO blocks per warp: lane hl of a half-warp owns row hl of [A | B]
          // (or [A | I]); see half_warp_gauss_jordan
          const int nmax = S.nup > S.ndown ? S.nup : S.ndown;
          auto run = [&](auto np_tag) {
            constexpr int NP = decltype(np_tag)::value;
            const int half = lane >> 4, hl = lane & 15;
            for (int it0 = 2 * warp; it0 < nitem; it0 += 2 * nwarp) {
              const int it = it0 + half;
              const bool act = it < nitem;
              const int wl = act ? it / nun : 0, u = act ? it - wl * nun : 0;
              const bool up = u < S.nuu;
              const int n = act ? (up ? S.nup : S.ndown) : 0;
              const int *cols = up ? T.ucu() + u * S.nup : T.ucd() + (u - S.nuu) * S.ndown;
              const double *A = smo + ((size_t)wl * Ne + (up ? 0 : S.nup) + hl) * ldm;
              double a[NP], r[NP];
#pragma unroll
              for (int j = 0; j < NP; ++j) {
                const bool in = hl < n && j < n;
                const int c = in ? cols[j] : 0;
                a[j] = in ? A[c] : 0.0;
                r[j] = MODE == MODE_ELOC ? (in ? A[chs + c] : 0.0) : (j == hl ? 1.0 : 0.0);
              }
              int kc;
              double ipiv;
              const double det =
                  half_warp_gauss_jordan<(MODE == MODE_ELOC || MODE == MODE_GRAD), NP>(n, nmax, a, r, hl, kc, ipiv);
              double tr = 0.0;
              if (MODE == MODE_ELOC) {
                // Tr(inv(A) B): the row that pivoted column k holds pivot * element (k, k) in r[k]
#pragma unroll
                for (int j = 0; j < NP; ++j) tr = (j == kc) ? r[j] : tr;
                tr *= ipiv;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, o, 16);
              }
              if (MODE == MODE_GRAD && kc < n) {
                // the gradient phase reads inv(A) from the scratch: right block of [A | I], row kc
                double *m = scr + (size_t)it * per + (size_t)kc * 2 * n + n;
#pragma unroll
                for (int j = 0; j < NP; ++j)
                  if (j < n) m[j] = r[j] * ipiv;
              }
              if (act && hl == 0) { sdet[wl * nun + u] = n > 0 ? det : 1.0; str[wl * nun + u] = tr; }
            }
          };
          if (nmax <= 8) run(IntTag<8>{});
          else if (nmax <= 12) run(IntTag<12>{});
          else run(IntTag<16>{});
          if (MODE == MODE_GRAD) __syncwarp();
        } else
        for (int it = warp; it < nitem; it += nwarp) {
          // blocks larger than 16x16: one warp per block on a shared-memory copy
          const int wl = it / nun, u = it - wl * nun;
          const bool up = u < S.nuu;
          const int n = up ? S.nup : S.ndown;
          const int *cols = up ? T.ucu() + u * S.nup : T.ucd() + (u - S.nuu) * S.ndown;
          const double *A = smo + ((size_t)wl * Ne + (up ? 0 : S.nup)) * ldm;
          double *m = scr + (size_t)(MODE == MODE_GRAD ? it : warp) * per;
          const int nr = (MODE == MODE_ELOC || MODE == MODE_GRAD) ? n : 0;
          const int ldw = n + nr;
          double det = 1.0, tr = 0.0;
          if (n > 0) {
            for (int idx = lane; idx < n * n; idx += 32) {
              const int i = idx / n, j = idx - i * n;
              m[i * ldw + j] = A[i * ldm + cols[j]];
              if (MODE == MODE_ELOC) m[i * ldw + n + j] = A[chs + i * ldm + cols[j]];
              if (MODE == MODE_GRAD) m[i * ldw + n + j] = i == j ? 1.0 : 0.0;
            }
            __syncwarp();
            det = warp_gauss_jordan(n, nr, m, lane);
            if (MODE == MODE_ELOC) {
              double v = 0.0;
              for (int i = lane; i < n; i += 32) v += m[i * ldw + n + i];
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
              tr = v;
            }
            __syncwarp();
          }
          if (lane == 0) { sdet[wl * nun + u] = det; str[wl * nun + u] = tr; }
        }
      } else {
      // scratch slot: GRAD keeps every inverse resident (slot = item); otherwise one slot
      // per participating thread, reused across rounds; elements interleaved (stride conc)
      const int conc = per ? lu_conc : nthr;
      const int stride = (MODE == MODE_GRAD) ? nthr : (conc < nthr ? conc : nthr);
      if (tid < stride) {
        for (int it = tid; it < nitem; it += stride) {
          const int wl = THREAD ? 0 : it / nun, u = it - wl * nun;
          const bool up = u < S.nuu;
          const int n = up ? S.nup : S.ndown;
          const int *cols = up ? T.ucu() + u * S.nup : T.ucd() + (u - S.nuu) * S.ndown;
          const double *A = smo + ((size_t)wl * Ne + (up ? 0 : S.nup)) * ldm;
          double det = 1.0, tr = 0.0;
          if (n == 0) {
            det = 1.0; tr = 0.0;
          } else if (MODE == MODE_GRAD) {
            double *m = scr + it;   // element stride = conc
            if (n <= 3) det = inverse_small(n, A, ldm, cols, m, conc);
            else if (TILE == 0 && n == 4) det = inverse_reg<4>(A, ldm, cols, m, 2 * n, conc);
            else if (TILE == 0 && n == 5) det = inverse_reg<5>(A, ldm, cols, m, 2 * n, conc);
            else if (TILE == 0 && n == 6) det = inverse_reg<6>(A, ldm, cols, m, 2 * n, conc);
            else {
              const int ldw = 2 * n;
              for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) {
                  m[(i * ldw + j) * conc] = A[i * ldm + cols[j]];
                  m[(i * ldw + n + j) * conc] = i == j ? 1.0 : 0.0;
                }
              det = gauss_jordan(n, n, m, conc);
            }
          } else if (n <= 3) {
            det_trace_small(n, A, A + chs, ldm, cols, MODE == MODE_ELOC, det, tr);
          } else if (TILE == 0 && n == 4) {   // CTA-tile kernels only: keeps calls out of the warp-tile kernels
            det_trace_reg<4, MODE == MODE_ELOC>(A, A + chs, ldm, cols, det, tr);
          } else if (TILE == 0 && n == 5) {
            det_trace_reg<5, MODE == MODE_ELOC>(A, A + chs, ldm, cols, det, tr);
          } else if (TILE == 0 && n == 6) {
            det_trace_reg<6, MODE == MODE_ELOC>(A, A + chs, ldm, cols, det, tr);
          } else {
            double *m = scr + tid;
            const int nr = MODE == MODE_ELOC ? n : 0;
            const int ldw = n + nr;
            for (int i = 0; i < n; ++i)
              for (int j = 0; j < n; ++j) {
                m[(i * ldw + j) * conc] = A[i * ldm + cols[j]];
                if (nr) m[(i * ldw + n + j) * conc] = A[chs + i * ldm + cols[j]];
              }
            det = gauss_jordan(n, nr, m, conc);
            if (nr)
              for (int i = 0; i < n; ++i) tr += m[(i * ldw + n + i) * conc];
          }
          sdet[wl * nun + u] = det;
          str[wl * nun + u] = tr;
        }
      }
      }
    }
    TILE_SYNC();
    // ---- P4: per-walker epilogue
    // long CI expansions: the sum over configurations is split over `parts` threads per walker
    // (strided configurations, partial sums added in part order: deterministic); the partials go
    // to the rows of jv that P2 has consumed (grad psi still needs them: it keeps the serial sum)
    int parts = 1;
    if (!THREAD && MODE != MODE_GRAD && S.nconf >= 8) {
      parts = nthr / TW;
      if (parts > 2 * Ne) parts = 2 * Ne;
      if (parts > 16) parts = 16;
      if (parts < 1) parts = 1;
    }
    if (parts > 1) {
      for (int it = tid; it < tw * parts; it += nthr) {
        const int wl = it / parts, pt = it - wl * parts;
        const double *dd = sdet + wl * nun, *tt = str + wl * nun;
        double sig = 0.0, ksig = 0.0;
        for (int c = pt; c < S.nconf; c += parts) {
          const int iu = T.ciu()[c], id = S.nuu + T.cid()[c];
          const double d = T.ci()[c] * dd[iu] * dd[id];
          sig += d;
          if (MODE == MODE_ELOC) ksig += d * (tt[iu] + tt[id]);
        }
        jv[2 * it] = sig; jv[2 * it + 1] = ksig;
      }
      TILE_SYNC();
    }
    for (int wl = tid; wl < tw; wl += nthr) {
      const double *dd = sdet + wl * nun, *tt = str + wl * nun;
      double sig = 0.0, ksig = 0.0;
      if (parts > 1) {
        for (int pt = 0; pt < parts; ++pt) { sig += jv[2 * (wl * parts + pt)]; ksig += jv[2 * (wl * parts + pt) + 1]; }
      } else
      for (int c = 0; c < S.nconf; ++c) {
        const int iu = T.ciu()[c], id = S.nuu + T.cid()[c];
        const double d = T.ci()[c] * dd[iu] * dd[id];
        sig += d;
        if (MODE == MODE_ELOC) ksig += d * (tt[iu] + tt[id]);
      }
      double ks = tks, ven = tven, vee = tvee;
      for (int e = 0; !THREAD && e < Ne; ++e) {
        const double *q = jv + wl * Ne + e;
        ks += q[4 * jvs]; ven += q[5 * jvs]; vee += q[6 * jvs];
      }
      const double J = (S.use_jee || S.use_jen || S.een_nterm > 0) ? exp_clamped(S, T.etab(), ks) : 1.0;
      const double psi = J * sig;
      if (MODE == MODE_PSI) {
        a.out0[w0 + wl] = psi;
      } else if (MODE == MODE_ELOC) {
        const double ekin = ksig / sig;
        a.out0[w0 + wl] = ekin + ven + vee + S.vnn;
        if (a.out1) a.out1[w0 + wl] = psi;
        if (a.out2) a.out2[w0 + wl] = ekin;
      } else if (MODE == MODE_MH) {
        double fxn = psi * psi;
        if (fxn == 0.0) fxn = a.eps;
        const double fx = a.out0[w0 + wl];
        double df = fxn / fx;
        if (df > 1.0) df = 1.0;
        const double tau = a.tau ? a.tau[w0 + wl]
                                 : philox_uniform(a.seed, a.offset, (uint64_t)(w0 + wl), 1u);
        const bool acc = (df - tau) >= 0.0;
        wsum[wl * 4] = acc ? 1.0 : 0.0;
        if (acc) a.out0[w0 + wl] = fxn;   // fxn is never 0 here
        if (a.accept) a.accept[w0 + wl] = acc ? 1 : 0;
      } else {
        wsum[wl * 4] = J;
        wsum[wl * 4 + 1] = sig;
      }
    }
    if (MODE == MODE_MH) {
      TILE_SYNC();
      int cnt = 0;
      for (int i = tid; i < tw * ne3; i += nthr) {
        const int wl = i / ne3;
        if (wsum[wl * 4] != 0.0) {
          a.pos_rw[w0 * ne3 + i] = spos[i];
          if (i == wl * ne3) ++cnt;
        }
      }
      if (a.naccept) {
        // (THREAD tiles: lanes leave the tile loop at different times)
        const unsigned mask = THREAD ? __activemask() : 0xffffffffu;
        cnt = __reduce_add_sync(mask, cnt);
        if ((int)(threadIdx.x & 31) == __ffs(mask) - 1 && cnt) atomicAdd(a.naccept, (unsigned long long)cnt);
      }
    }
    if (MODE == MODE_GRAD) {
      TILE_SYNC();
      // d psi / d r_{e,c} = J [ sum_u C_u sum_j inv_u[j][e] dmo_c[e][cols_u[j]] + g_{e,c} Sigma ]
      // (slater_jastrow.py:346-447), C_u = D_u * sum_{n: occ_s(n)=u} c_n D_other(n)
      const int conc = lu_conc;
      // C_u once per (walker, unique occupation) - not once per electron - into the trace slots,
      // which grad psi does not use
      for (int it = tid; it < tw * nun; it += nthr) {
        const int wl = it / nun, u = it - wl * nun;
        const bool up = u < S.nuu;
        const int us = up ? u : u - S.nuu;
        const double *dd = sdet + wl * nun;
        double cu = 0.0;
        for (int c = 0; c < S.nconf; ++c) {
          if ((up ? T.ciu()[c] : T.cid()[c]) != us) continue;
          cu += T.ci()[c] * dd[up ? S.nuu + T.cid()[c] : T.ciu()[c]];
        }
        str[it] = cu * dd[u];
      }
      TILE_SYNC();
      for (int it = tid; it < tw * Ne; it += nthr) {
        const int wl = it / Ne, e = it - wl * Ne;
        const bool up = e < S.nup;
        const int n = up ? S.nup : S.ndown;
        const int el = up ? e : e - S.nup;
        const double *row = smo + ((size_t)wl * Ne + e) * ldm;
        double gsx = 0, gsy = 0, gsz = 0;
        const int nu = up ? S.nuu : S.nud;
        for (int u = 0; u < nu; ++u) {
          const double cu = str[wl * nun + (up ? u : S.nuu + u)];
          if (cu == 0.0) continue;
          const int item = wl * nun + (up ? u : S.nuu + u);
          const bool contiguous = TILE == 0 && use_warp_lu(S);   // layout written by P3
          const double *inv = contiguous ? scr + (size_t)item * lu_scratch_per_item(S, MODE, TILE != 0) : scr + item;
          const int es = contiguous ? 1 : conc;
          const int *cols = up ? T.ucu() + u * S.nup : T.ucd() + u * S.ndown;
          double tx = 0, ty = 0, tz = 0;
          const int ild = (n <= 3 && !contiguous) ? n : 2 * n, ioff = (n <= 3 && !contiguous) ? 0 : n;
          for (int j = 0; j < n; ++j) {
            const double iv = inv[(j * ild + ioff + el) * es];
            tx += iv * row[chs + cols[j]];
            ty += iv * row[2 * chs + cols[j]];
            tz += iv * row[3 * chs + cols[j]];
          }
          gsx += cu * tx; gsy += cu * ty; gsz += cu * tz;
        }
        const double J = wsum[wl * 4], sig = wsum[wl * 4 + 1];
        const double *q = jv + it;
        double ox = J * (gsx + q[0] * sig), oy = J * (gsy + q[jvs] * sig), oz = J * (gsz + q[2 * jvs] * sig);
        if (a.pdf) {
          const double f = 2.0 * sig * J;
          ox *= f; oy *= f; oz *= f;
        }
        double *g = a.out0 + (w0 + wl) * ne3 + 3 * e;
        g[0] = ox; g[1] = oy; g[2] = oz;
      }
    }
    TILE_SYNC();
  }
#undef TILE_SYNC
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
#ifdef QMCB_FUSED_MAIN
static int choose(const qmcb_plan *p, int mode, LaunchCfg &c) {
  const DevSys &S = p->sys;
  int mb = 1;
  while (mb < S.nmu && mb < 8) mb *= 2;
  // psi / E_L / Metropolis contract one or two AO channels: 16 columns per thread fit the register
  // budget, and a thread that owns all columns evaluates the basis functions of its electron once
  if (mode != MODE_GRAD && S.nmu > 8 && S.nmup % 16 == 0 && S.nelec * (S.nmup / 16) <= QMCB_MB16_THREADS &&
      !getenv("QMCB_MB8"))
    mb = 16;
  c.mb = mb;
  c.nblk = S.nmup / mb;
  const int per_walker = S.nelec * c.nblk;
  if (per_walker > 512) {
    qmcb_set_error("qmcb: nelec * MO blocks exceeds one CTA");
    return QMCB_ESMEM;
  }
  const int nun = S.nuu + S.nud;
  const int budget = p->smem_optin - 1024;
  const size_t tab = (size_t)table_doubles(S) * sizeof(double);
  // ---- one walker per thread: small systems with closed-form determinants (not for grad psi)
  {
    const int nbig = S.nup > S.ndown ? S.nup : S.ndown;
    const char *env = getenv("QMCB_TILE");          // tuning experiments: force 0/1 (never 2 by force)
    const bool allow = !(env && (env[0] == '0' || env[0] == '1'));
    if (allow && mode != MODE_GRAD && c.nblk == 1 && mb <= 4 && nbig <= 3 && S.nelec <= 8 && nun <= 8 &&
        S.een_nterm == 0) {
      const size_t unit = thread_tile_doubles(S, mode) * sizeof(double);
      const size_t sm = tab + unit * QMCB_THREAD_CTA;
      if ((int)sm <= budget / 4) {
        c.warp = 2; c.tw = 1; c.threads = QMCB_THREAD_CTA; c.smem = (int)sm; c.lu_conc = 0;
        return 0;
      }
    }
  }
  // ---- warp-owned tiles when the threads of a walker tile a warp exactly
  if (32 % per_walker == 0 && !(getenv("QMCB_TILE") && getenv("QMCB_TILE")[0] == '0')) {
    const int tw = 32 / per_walker;
    const int per = lu_scratch_per_item(S, mode, true);
    int conc = 0;
    if (per) {
      conc = tw * nun;
      if (mode != MODE_GRAD && conc > 32) conc = 32;
    }
    const size_t unit = tile_doubles(S, mode, tw, conc, true) * sizeof(double);
    for (int warps = QMCB_WARP_CTA / 32; warps >= 1; warps /= 2) {
      const size_t sm = tab + unit * warps;
      if ((int)sm <= budget) {
        c.warp = 1; c.tw = tw; c.threads = warps * 32; c.smem = (int)sm; c.lu_conc = conc;
        return 0;
      }
    }
  }
  // ---- CTA-owned tiles
  // two 256-thread CTAs per SM overlap each other's barriers (measured, C4H6: E_L 1.59 -> 1.42 ms,
  // psi 0.92 -> 0.76 ms; H2O unchanged); grad psi keeps its inverses resident and prefers one large CTA
  int cta_max = mode == MODE_GRAD ? 512 : 256;
  if (const char *env = getenv("QMCB_CTA_THREADS")) cta_max = atoi(env) > 0 ? atoi(env) : cta_max;
  const int hard_max = mb == 16 ? QMCB_MB16_THREADS : 512;       // launch bounds of the instantiation
  if (cta_max > hard_max) cta_max = hard_max;
  if (cta_max < per_walker) cta_max = ((per_walker + 31) / 32) * 32;
  int tw = cta_max / per_walker;
  if (tw > 128) tw = 128;
  if (tw < 1) tw = 1;
  for (; tw >= 1; --tw) {
    int threads = ((tw * per_walker + 31) / 32) * 32;
    if (threads > hard_max) continue;
    const int per = lu_scratch_per_item(S, mode, false);
    int conc = 0;
    if (per) {
      conc = tw * nun;                                     // GRAD: every inverse stays resident
      // otherwise one slot per warp (warp-cooperative Gauss-Jordan) or per thread
      const int slots = use_warp_lu(S) ? threads / 32 : threads;
      if (mode != MODE_GRAD && conc > slots) conc = slots;
    }
    const size_t sm = tab + tile_doubles(S, mode, tw, conc, false) * sizeof(double);
    if ((int)sm <= budget) {
      c.warp = 0; c.tw = tw; c.threads = threads; c.smem = (int)sm; c.lu_conc = conc;
      return 0;
    }
  }
  qmcb_set_error("qmcb: system does not fit the shared-memory tiling");
  return QMCB_ESMEM;
}

int qmcb_choose_launch(qmcb_plan *p) {
  int rc = choose(p, MODE_PSI, p->cfg_psi);
  if (!rc) rc = choose(p, MODE_ELOC, p->cfg_eloc);
  if (!rc) rc = choose(p, MODE_GRAD, p->cfg_grad);
  return rc;
}
#endif  // QMCB_FUSED_MAIN

template <int MODE, int MB, int RT, int TILE>
static int launch_k(const qmcb_plan *p, const LaunchCfg &c, const FusedArgs &a, cudaStream_t st) {
  constexpr bool WARP = TILE == 1;
  auto k = fused_kernel<MODE, MB, RT, TILE>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, c.smem);
  if (e != cudaSuccess) return qmcb_cuda_rc((int)e, "fused_impl.cuh");
  const int64_t ntile = (a.W + c.tw - 1) / c.tw;
  const int64_t units_per_cta = TILE == 2 ? c.threads : (WARP ? c.threads / 32 : 1);
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, c.threads, c.smem);
  if (occ < 1) occ = 1;
  int64_t grid = (int64_t)p->sm_count * occ;
  const int64_t need = (ntile + units_per_cta - 1) / units_per_cta;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  k<<<(unsigned)grid, c.threads, c.smem, st>>>(p->sys, a, c.tw, c.nblk, c.lu_conc);
  return qmcb_cuda_rc((int)cudaGetLastError(), "fused_impl.cuh launch");
}

template <int MODE, int MB>
static int launch_t(const qmcb_plan *p, const LaunchCfg &c, const FusedArgs &a, cudaStream_t st) {
  const bool pure = p->sys.radial_type == QMCB_GTO_PURE;
  if (c.warp == 2) {
    // one walker per thread: instantiated for psi / E_L / Metropolis and narrow column blocks only
    if constexpr (MODE != MODE_GRAD && MB <= 4)
      return pure ? launch_k<MODE, MB, 0, 2>(p, c, a, st) : launch_k<MODE, MB, 1, 2>(p, c, a, st);
  }
  if (c.warp == 1) return pure ? launch_k<MODE, MB, 0, 1>(p, c, a, st) : launch_k<MODE, MB, 1, 1>(p, c, a, st);
  return pure ? launch_k<MODE, MB, 0, 0>(p, c, a, st) : launch_k<MODE, MB, 1, 0>(p, c, a, st);
}

template <int MODE>
static int launch(const qmcb_plan *p, const LaunchCfg &c, const FusedArgs &a, cudaStream_t st) {
  switch (c.mb) {
    case 1: return launch_t<MODE, 1>(p, c, a, st);
    case 2: return launch_t<MODE, 2>(p, c, a, st);
    case 4: return launch_t<MODE, 4>(p, c, a, st);
    case 16:
      if constexpr (MODE != MODE_GRAD) return launch_t<MODE, 16>(p, c, a, st);
    default: return launch_t<MODE, 8>(p, c, a, st);
  }
}


static inline int check(const qmcb_plan *p, const void *pos, int64_t W) {
  if (!p || !p->d_dbl) { qmcb_set_error("qmcb: plan has no device tables"); return QMCB_EINVAL; }
  if (W < 0 || (W > 0 && !pos)) { qmcb_set_error("qmcb: bad walker array"); return QMCB_EINVAL; }
  return 0;
}
