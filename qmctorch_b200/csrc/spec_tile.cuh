// Structure-specialised WARP-TILE kernels (NVRTC): psi, E_L (+ fused statistics) and the Metropolis
// step for mid-size and large wave functions - spin blocks up to 16 x 16, up to 32 electrons, up to 16
// occupied MO columns, CI expansions, three-body Jastrow: H2O CAS, CO2, C4H6 (BASELINE configs 4, 5).
//
// The generic CTA-tile kernel (fused_impl.cuh) interprets the basis between CTA barriers and sits at
// 37 % of the FP64 pipe, latency-bound (profiles/README.md).  Here
//   * ONE WARP owns PER = 32 / Ne walkers from the coordinate load to E_L: lane = (walker, electron),
//     the electrons of a walker on consecutive lanes.  There is no CTA barrier in the walker loop - only
//     __syncwarp - so the warps of an SM drift apart and the pipe-bound projection of one warp overlaps
//     the latency-bound exponentials, pivots and divisions of the others;
//   * the Jastrow terms of an electron (grad ln J, lap J / J) never leave its registers between the
//     pair phase and the basis walk that folds them into the kinetic channel (device.cuh: FoldJ);
//   * the shell walk + AO -> MO contraction is the GENERATED straight-line program of spec.cu (one
//     basic block per electron: ptxas interleaves the independent exponentials of an atom with the
//     projection FMAs of the previous shell); primitive constants are constant-bank operands, the MO
//     weights come from shared memory (two per LDS.128, broadcast) or the constant bank (small bases);
//   * every size - Ne, spin-block orders, occupied columns, unique determinants, configurations,
//     three-body terms - is a compile-time constant: the determinant phase is det_trace_reg<N> /
//     half_warp_gauss_jordan<NP> for exactly this structure, loops over atoms and terms unroll.
// Arithmetic = the generic kernel's (same device functions), same parity tests.
//
// Compiled with -DQMCB_SPEC -DSPEC_TILE after the generated prelude, fused_args.h, device.cuh,
// philox.cuh, spec_common.cuh; the generated spec_aos<MODE> is spliced in at the marker below.
#pragma once

// prelude (tile kernels) additionally defines: SPEC_NMU (= SPEC_NMUP: exact occupied columns), SPEC_NAO,
// SPEC_NCONF, SPEC_O_MOW SPEC_O_CI (double-table offsets), SPEC_O_UCU SPEC_O_UCD SPEC_O_CIU SPEC_O_CID
// (int-table offsets), SPEC_MWLD (row stride of the MO weights in the plan's table), SPEC_THREADS, SPEC_MINB

SPEC_GENERATED_CODE

#ifndef SPEC_TILE_PREFETCH
#define SPEC_TILE_PREFETCH 1
#endif
#ifndef SPEC_TILE_ALIGN
#define SPEC_TILE_ALIGN 0     // 1: the warps of a CTA start every tile together (instruction-cache sharing experiment)
#endif

struct TileTab {
  const SpecParams &P;
  __device__ __forceinline__ SpecVals atoms() const { return SpecVals{P, SPEC_OFF_ATOM}; }
};

template <int MODE>
__device__ __forceinline__ void spect_body(const SpecParams &P, const FusedArgs &a) {
  constexpr int NCH = spec_nch<MODE>();
  constexpr int Ne = SPEC_NE, ne3 = 3 * SPEC_NE, NM = SPEC_NMUP, NUU = SPEC_NUU, NUD = SPEC_NUD, NUN = NUU + NUD;
  constexpr int NUP = SPEC_NUP, NDN = SPEC_NDOWN, NBIG = NUP > NDN ? NUP : NDN, NCONF = SPEC_NCONF;
  constexpr int PER = 32 / Ne;                       // walkers per warp
  constexpr int LDM = NM | 1;                        // odd row stride: lanes own consecutive rows
  constexpr int NROW = MODE == MODE_ELOC ? 2 : 1;    // mo | B_kin
  constexpr int NWARP = SPEC_THREADS / 32;
  constexpr bool WB = MODE == MODE_ELOC;
  constexpr bool HASJ = SPEC_USE_JEE || SPEC_USE_JEN || SPEC_EEN_NTERM > 0;
  // ---- CTA tables: exp table | MO weights (optional) | CI coefficients | int tables
  constexpr int NMW = SPEC_MOW_SMEM ? SPEC_NAO * SPEC_MWLD : 0;
  constexpr int NCI = (NCONF + 1) & ~1;
  constexpr int NIT = (NUU * NUP + NUD * NDN + 2 * NCONF + 1) & ~1;
  // ---- per-warp work area (doubles)
  // spos [PER][3Ne]; PF: a second buffer receives the NEXT tile's coordinates (cp.async) while this one
  // is computed (psi / E_L: nothing is added to the coordinates)
  constexpr bool PF = SPEC_TILE_PREFETCH && MODE != MODE_MH;
  constexpr int NPOS = (PER * ne3 + 1) & ~1;
  constexpr int O_JV = NPOS * (PF ? 2 : 1);
  constexpr int O_MO = O_JV + 3 * 32;                // jv [3][32]: per-lane ks / ven / vee (later CI partials)
  constexpr int ROWS = PER * Ne * LDM;
  // the three-body factor tables (device.cuh: een_table_*) live where the mo / B_kin rows are written later
  constexpr int TLS = SPEC_EEN_NTERM > 0 ? een_table_doubles<SpecParams>() : 0;
  constexpr int NMO = NROW * ROWS > 32 * TLS ? NROW * ROWS : 32 * TLS;
  constexpr int O_DET = O_MO + NMO;
  constexpr int O_TR = O_DET + PER * NUN;
  constexpr int WS = (O_TR + PER * NUN + 1) & ~1;
  extern __shared__ __align__(16) double smem[];
  double *et = smem;
  double *mw = et + QMCB_ETAB;
  double *sci = mw + NMW;
  int *sit = reinterpret_cast<int *>(sci + NCI);
  // (the shuffle tells the compiler that the warp index - and every loop bound derived from it - is warp-uniform)
  const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  double *wsp = reinterpret_cast<double *>(sit + NIT) + (size_t)warp * WS;
  for (int i = threadIdx.x; i < QMCB_ETAB; i += blockDim.x) et[i] = P.etab_g[i];
  for (int i = threadIdx.x; i < NMW; i += blockDim.x) mw[i] = P.dtab_g[SPEC_O_MOW + i];
  for (int i = threadIdx.x; i < NCONF; i += blockDim.x) sci[i] = P.dtab_g[SPEC_O_CI + i];
  int *s_ucu = sit, *s_ucd = s_ucu + NUU * NUP, *s_ciu = s_ucd + NUD * NDN, *s_cid = s_ciu + NCONF;
  for (int i = threadIdx.x; i < NUU * NUP; i += blockDim.x) s_ucu[i] = P.itab_g[SPEC_O_UCU + i];
  for (int i = threadIdx.x; i < NUD * NDN; i += blockDim.x) s_ucd[i] = P.itab_g[SPEC_O_UCD + i];
  for (int i = threadIdx.x; i < NCONF; i += blockDim.x) { s_ciu[i] = P.itab_g[SPEC_O_CIU + i]; s_cid[i] = P.itab_g[SPEC_O_CID + i]; }
  double *spos = wsp, *jv = wsp + O_JV, *smo = wsp + O_MO, *sdet = wsp + O_DET, *str_ = wsp + O_TR;
  const TileTab T{P};
  __syncthreads();
  const unsigned full = 0xffffffffu;
  const int sub = lane / Ne, e = lane - sub * Ne;
  const int64_t nunit = (int64_t)gridDim.x * NWARP, unit = (int64_t)blockIdx.x * NWARP + warp;
  const int64_t ntile = (a.W + PER - 1) / PER;
  double st_s = 0.0, st_s2 = 0.0;     // fused energy statistics (E_L only), lanes with e == 0
  int st_nf = 0, st_nb = 0;
  // cp.async (LDGSTS): 8-byte copies global -> shared without registers; buffer `par` of the double buffer
  auto prefetch = [&](int64_t t, int par) {
    if (PF && t < ntile) {
      const int64_t w0n = t * PER;
      const int twn = (int)((a.W - w0n) < PER ? (a.W - w0n) : PER);
      const unsigned dst = (unsigned)__cvta_generic_to_shared(wsp + par * NPOS);
      const double *src = a.pos + w0n * ne3;
      for (int i = lane; i < twn * ne3; i += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8 * i), "l"(src + i) : "memory");
    }
    if (PF) asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int par = 0;
  prefetch(unit, 0);

  const int64_t nround = (ntile + nunit - 1) / nunit;     // the same trip count for every warp (SPEC_TILE_ALIGN)
  int64_t tile = unit;
  for (int64_t rnd = 0; SPEC_TILE_ALIGN ? rnd < nround : tile < ntile; ++rnd, tile += nunit, par ^= 1) {
    if (SPEC_TILE_ALIGN) {
      __syncthreads();
      if (tile >= ntile) continue;
    }
    const int64_t w0 = tile * PER;
    const int tw = (int)((a.W - w0) < PER ? (a.W - w0) : PER);
    const bool act = sub < tw;                       // (sub < PER follows: tw <= PER)
    if (PF) {
      spos = wsp + par * NPOS;
      prefetch(tile + nunit, par ^ 1);               // the other buffer was consumed one tile ago
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else
    // ---- P0: coordinates (+ proposal), coalesced over the warp's walkers
    if (MODE == MODE_MH && !a.disp && a.proba_normal) {
      // (MH never prefetches: PF is false)
      // in-kernel normal proposals: one Philox call yields the four normals of a GLOBAL element quad
      // (4q .. 4q+3), so the draw of an element does not depend on the tiling (same stream as the
      // generic and the one-walker-per-thread kernels)
      const int64_t g0 = w0 * ne3, g1 = g0 + (int64_t)tw * ne3;
      for (int64_t q = (g0 >> 2) + lane; 4 * q < g1; q += 32) {
        double z[4];
        philox_normal4(a.seed, a.offset, (uint64_t)q, z);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int64_t g = 4 * q + h;
          if (g < g0 || g >= g1) continue;
          const int i = (int)(g - g0);
          const int wl = i / ne3, el = (i - wl * ne3) / 3;
          int me = a.move_elec;
          if (me == -2) {
            if (a.elec_index) me = a.elec_index[w0 + wl];
            else me = (int)(philox_u32(a.seed, a.offset, (uint64_t)(w0 + wl), 2u) % (unsigned)Ne);
          }
          double v = a.pos[g];
          if (me < 0 || me == el) v += a.scale * z[h];
          spos[i] = v;
        }
      }
    } else {
      for (int i = lane; i < tw * ne3; i += 32) {
        double v = a.pos[w0 * ne3 + i];
        if (MODE == MODE_MH) {
          const int wl = i / ne3, el = (i - wl * ne3) / 3;
          int me = a.move_elec;
          if (me == -2) {
            if (a.elec_index) me = a.elec_index[w0 + wl];
            else me = (int)(philox_u32(a.seed, a.offset, (uint64_t)(w0 + wl), 2u) % (unsigned)Ne);
          }
          if (me < 0 || me == el) {
            double d;
            if (a.disp) d = a.disp[w0 * ne3 + i];
            else d = a.scale * (2.0 * philox_uniform(a.seed, a.offset, (uint64_t)(w0 * ne3 + i), 0u) - 1.0);
            v += d;
          }
        }
        spos[i] = v;
      }
    }
    __syncwarp();
    // ---- P1: Jastrow terms + e-e potential of this lane's electron (registers)
    const double *sp = spos + (act ? sub : 0) * ne3;
    ElecTerms o;
    o.gx = o.gy = o.gz = o.lap = o.ks = o.ven = o.vee = 0.0;
    if (HASJ || WB) {
      if (spec_deriv<MODE>() && Ne >= 2 && SPEC_USE_JEE) {
        // every e-e pair once; partners exchange their contributions by shuffles (all lanes take part)
        electron_terms_paired<WB, false, SPEC_EEN_NTERM == 0>(P, T, sp, act ? e : 0, sub * Ne, act, o);
      } else if (act) {
        electron_terms<spec_deriv<MODE>(), WB, false, false, SPEC_EEN_NTERM == 0>(P, T, sp, e, o);
      }
      if constexpr (SPEC_EEN_NTERM > 0) {
        // three-body term: tabulate this lane's electron, then the pair loop on the tables
        double *tab = smo + lane * TLS;
        if (act) een_table_fill(P, T, sp, e, tab);
        __syncwarp();
        if (act) {
          een_table_terms<spec_deriv<MODE>()>(P, T, sp, e, smo + sub * Ne * TLS, TLS, o.gx, o.gy, o.gz, o.lap, o.ks);
          if (spec_deriv<MODE>()) o.lap += o.gx * o.gx + o.gy * o.gy + o.gz * o.gz;
        }
        __syncwarp();       // the tables are dead: the rows of mo / B_kin reuse the space
      }
    }
    // ---- P2: generated shell walk + projection; rows of mo (and B_kin) -> shared memory
    if (act) {
      double acc[NCH][NM];
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int j = 0; j < NM; ++j) acc[c][j] = 0.0;
      FoldJ fj{0.0, 0.0, 0.0, 0.0};
      if (WB && HASJ) fj = FoldJ{2.0 * o.gx, 2.0 * o.gy, 2.0 * o.gz, o.lap};
      double ven = 0.0;
      spec_aos<MODE>(P, et, mw, sp[3 * e], sp[3 * e + 1], sp[3 * e + 2], fj, ven, acc);
      o.ven = ven;
      double *row = smo + (sub * Ne + e) * LDM;
#pragma unroll
      for (int j = 0; j < NM; ++j) {
        row[j] = acc[0][j];
        if (WB) row[ROWS + j] = -0.5 * acc[NCH - 1][j];
      }
    }
    jv[lane] = o.ks; jv[32 + lane] = o.ven; jv[64 + lane] = o.vee;
    __syncwarp();
    // ---- P3: determinants (and traces) of the unique spin occupations of the warp's walkers
    if constexpr (NBIG <= 6) {
      // one thread per spin block: closed forms (n <= 3) or register LU (n = 4..6)
      for (int it = lane; it < tw * NUN; it += 32) {
        const int wl = it / NUN, u = it - wl * NUN;
        const bool up = u < NUU;
        const int *cols = up ? s_ucu + u * NUP : s_ucd + (u - NUU) * NDN;
        const double *A = smo + (wl * Ne + (up ? 0 : NUP)) * LDM;
        double det = 1.0, tr = 0.0;
        if constexpr (NUP == NDN) {
          // equal spin blocks: ONE call for all lanes (two calls would run the up and the down lanes one
          // after the other)
          if constexpr (NUP >= 4) det_trace_reg<NUP, WB>(A, A + ROWS, LDM, cols, det, tr);
          else if constexpr (NUP >= 1) det_trace_small(NUP, A, A + ROWS, LDM, cols, WB, det, tr);
        } else if (up) {
          if constexpr (NUP >= 4) det_trace_reg<NUP, WB>(A, A + ROWS, LDM, cols, det, tr);
          else if constexpr (NUP >= 1) det_trace_small(NUP, A, A + ROWS, LDM, cols, WB, det, tr);
        } else {
          if constexpr (NDN >= 4) det_trace_reg<NDN, WB>(A, A + ROWS, LDM, cols, det, tr);
          else if constexpr (NDN >= 1) det_trace_small(NDN, A, A + ROWS, LDM, cols, WB, det, tr);
        }
        sdet[it] = det;
        str_[it] = tr;
      }
    } else {
      // two blocks per warp, one per half-warp, lane = ROW of [A | B] (device.cuh: half_warp_gauss_jordan)
      constexpr int NP = NBIG <= 8 ? 8 : (NBIG <= 12 ? 12 : 16);
      const int half = lane >> 4, hl = lane & 15;
      for (int it0 = 0; it0 < tw * NUN; it0 += 2) {
        const int it = it0 + half;
        const bool on = it < tw * NUN;
        const int wl = on ? it / NUN : 0, u = on ? it - wl * NUN : 0;
        const bool up = u < NUU;
        const int n = on ? (up ? NUP : NDN) : 0;
        const int *cols = up ? s_ucu + u * NUP : s_ucd + (u - NUU) * NDN;
        const double *A = smo + (wl * Ne + (up ? 0 : NUP) + hl) * LDM;
        double ar[NP], rr[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          const bool in = hl < n && j < n;
          const int c = in ? cols[j] : 0;
          ar[j] = in ? A[c] : 0.0;
          rr[j] = WB ? (in ? A[ROWS + c] : 0.0) : 0.0;
        }
        int kc;
        double ipiv;
        const double det = half_warp_gauss_jordan<WB, NP>(n, NBIG, ar, rr, hl, kc, ipiv);
        double tr = 0.0;
        if (WB) {
#pragma unroll
          for (int j = 0; j < NP; ++j) tr = (j == kc) ? rr[j] : tr;
          tr *= ipiv;
#pragma unroll
          for (int ofs = 8; ofs > 0; ofs >>= 1) tr += __shfl_xor_sync(full, tr, ofs, 16);
        }
        if (on && hl == 0) { sdet[it] = n > 0 ? det : 1.0; str_[it] = tr; }
      }
    }
    __syncwarp();
    // ---- P4: CI sum, psi, E_L / accept
    double sig = 0.0, ksig = 0.0;
    if constexpr (NCONF >= 8) {
      // configurations strided over the walker's lanes, partial sums added in lane order
      if (act) {
        const double *dd = sdet + sub * NUN, *tt = str_ + sub * NUN;
        double s0 = 0.0, k0 = 0.0;
        for (int c = e; c < NCONF; c += Ne) {
          const int iu = s_ciu[c], id = NUU + s_cid[c];
          const double d = sci[c] * dd[iu] * dd[id];
          s0 += d;
          if (WB) k0 += d * (tt[iu] + tt[id]);
        }
        // ks / ven / vee of this lane are still needed: the partials go to the rows of mo, which the
        // determinant phase has consumed
        smo[2 * lane] = s0; smo[2 * lane + 1] = k0;
      }
      __syncwarp();
      if (act && e == 0) {
        for (int k = 0; k < Ne; ++k) { sig += smo[2 * (lane + k)]; ksig += smo[2 * (lane + k) + 1]; }
      }
    } else if (act && e == 0) {
      const double *dd = sdet + sub * NUN, *tt = str_ + sub * NUN;
#pragma unroll
      for (int c = 0; c < NCONF; ++c) {
        const int iu = s_ciu[c], id = NUU + s_cid[c];
        const double d = sci[c] * dd[iu] * dd[id];
        sig += d;
        if (WB) ksig += d * (tt[iu] + tt[id]);
      }
    }
    bool accepted = false;
    if (act && e == 0) {
      double ks = 0.0, ven = 0.0, vee = 0.0;
#pragma unroll
      for (int k = 0; k < Ne; ++k) { ks += jv[lane + k]; ven += jv[32 + lane + k]; vee += jv[64 + lane + k]; }
      const double J = HASJ ? exp_clamped(P, et, ks) : 1.0;
      const double psi = J * sig;
      const int64_t w = w0 + sub;
      if (MODE == MODE_PSI) {
        a.out0[w] = psi;
      } else if (MODE == MODE_ELOC) {
        const double ekin = ksig / sig;
        const double el = ekin + ven + vee + P.vnn;
        a.out0[w] = el;
        if (isfinite(el)) { st_s += el; st_s2 = fma(el, el, st_s2); ++st_nf; } else ++st_nb;
        if (a.out1) a.out1[w] = psi;
        if (a.out2) a.out2[w] = ekin;
      } else if (MODE == MODE_MH) {
        double fxn = psi * psi;
        if (fxn == 0.0) fxn = a.eps;
        const double fx = a.out0[w];
        double df = fxn / fx;
        if (df > 1.0) df = 1.0;
        const double tau = a.tau ? a.tau[w] : philox_uniform(a.seed, a.offset, (uint64_t)w, 1u);
        accepted = (df - tau) >= 0.0;
        if (accepted) a.out0[w] = fxn;   // fxn is never 0 here
        if (a.accept) a.accept[w] = accepted ? 1 : 0;
      }
    }
    if (MODE == MODE_MH) {
      // the walker's lanes write back its accepted coordinates
      const int acc_w = __shfl_sync(full, accepted ? 1 : 0, (sub < PER ? sub : 0) * Ne);
      if (act && acc_w) {
        const double *s3 = spos + sub * ne3 + 3 * e;
        double *dst = a.pos_rw + (w0 + sub) * ne3 + 3 * e;
        dst[0] = s3[0]; dst[1] = s3[1]; dst[2] = s3[2];
      }
      if (a.naccept) {
        const int cnt = __reduce_add_sync(full, (act && e == 0 && accepted) ? 1 : 0);
        if (lane == 0 && cnt) atomicAdd(a.naccept, (unsigned long long)cnt);
      }
    }
    __syncwarp();
  }
  // ---- fused statistics (as spec_kernel.cuh): fixed-order reduction -> one partial per CTA; the CTA that
  // arrives last adds the partials in index order
  if (MODE == MODE_ELOC && a.stats_part) {
    double q[4] = {st_s, st_s2, (double)st_nf, (double)st_nb};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int ofs = 16; ofs > 0; ofs >>= 1) q[k] += __shfl_xor_sync(full, q[k], ofs);
    __syncthreads();
    double *red = reinterpret_cast<double *>(sit + NIT);
    if (lane == 0)
#pragma unroll
      for (int k = 0; k < 4; ++k) red[warp * 4 + k] = q[k];
    __syncthreads();
    if (threadIdx.x < 4) {
      double t = 0.0;
      for (int wq = 0; wq < NWARP; ++wq) t += red[wq * 4 + threadIdx.x];
      a.stats_part[blockIdx.x * 4 + threadIdx.x] = t;
    }
    if (a.stats_ticket && a.stats_out) {
      __shared__ int last;
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) last = atomicInc(a.stats_ticket, gridDim.x - 1) == gridDim.x - 1;
      __syncthreads();
      if (last) {
        __threadfence();
        for (int qi = warp; qi < 4; qi += NWARP) {
          double s = 0.0;
          for (int i = lane; i < (int)gridDim.x; i += 32) s += __ldcg(a.stats_part + i * 4 + qi);
#pragma unroll
          for (int ofs = 16; ofs > 0; ofs >>= 1) s += __shfl_xor_sync(full, s, ofs);
          if (lane == 0) a.stats_out[qi] = s;
        }
      }
    }
  }
}

extern "C" __global__ void __launch_bounds__(SPEC_THREADS, SPEC_MINB)
    spect_psi(const __grid_constant__ SpecParams P, const FusedArgs a) { spect_body<MODE_PSI>(P, a); }
extern "C" __global__ void __launch_bounds__(SPEC_THREADS, SPEC_MINB)
    spect_eloc(const __grid_constant__ SpecParams P, const FusedArgs a) { spect_body<MODE_ELOC>(P, a); }
extern "C" __global__ void __launch_bounds__(SPEC_THREADS, SPEC_MINB)
    spect_mh(const __grid_constant__ SpecParams P, const FusedArgs a) { spect_body<MODE_MH>(P, a); }
