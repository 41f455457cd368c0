// Operator-level entry points: the reference's nn.Module operators one by one, for
// sub-module parity and for users that call wf.ao / wf.mo / wf.jastrow / wf.pool directly.
// These write their (large) results to HBM by construction; the fused kernel in fused.cu
// is the hot path.
#include <cuda_runtime.h>

#include "device.cuh"

// ---- AtomicOrbitals.forward -----------------------------------------------------------
// MULTI: some AO is a sum of several cartesian monomials (real spherical harmonics of l = 2, plan.cu):
// its components accumulate into a zeroed row instead of storing
template <int NCH, bool MULTI>
struct AoStoreSink {
  double *ao, *dao, *d2ao;   // already offset to this (walker, electron) row
  __device__ __forceinline__ void emit(int a, const double (&v)[NCH]) {
    if (MULTI) {
      ao[a] += v[0];
      if (NCH > 1) {
        dao[3 * a] += v[1]; dao[3 * a + 1] += v[2]; dao[3 * a + 2] += v[3];
        d2ao[a] += v[4];
      }
    } else {
      ao[a] = v[0];
      if (NCH > 1) {
        dao[3 * a] = v[1]; dao[3 * a + 1] = v[2]; dao[3 * a + 2] = v[3];
        d2ao[a] = v[4];
      }
    }
  }
};

template <int NCH, bool MULTI>
__global__ void __launch_bounds__(256) ao_kernel(const DevSys S, const double *pos, int64_t rows, double *ao,
                                                 double *dao, double *d2ao) {
  extern __shared__ __align__(16) double smem[];
  Tab T;
  stage_tables(S, smem, T);
  __syncthreads();
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < rows;
       row += (int64_t)gridDim.x * blockDim.x) {
    AoStoreSink<NCH, MULTI> sink;
    sink.ao = ao + row * S.nao;
    sink.dao = NCH > 1 ? dao + row * S.nao * 3 : nullptr;
    sink.d2ao = NCH > 1 ? d2ao + row * S.nao : nullptr;
    if (MULTI) {
      for (int k = 0; k < S.nao; ++k) {
        sink.ao[k] = 0.0;
        if (NCH > 1) { sink.dao[3 * k] = 0.0; sink.dao[3 * k + 1] = 0.0; sink.dao[3 * k + 2] = 0.0; sink.d2ao[k] = 0.0; }
      }
    }
    eval_aos<NCH, 1>(S, T, pos[3 * row], pos[3 * row + 1], pos[3 * row + 2], sink);
  }
}

extern "C" int qmcb_ao(const qmcb_plan *p, const double *pos, int64_t W, int one_elec, double *ao,
                       double *dao, double *d2ao, void *stream) {
  if (!p || !p->d_dbl || !pos || !ao || W < 0 || ((dao == nullptr) != (d2ao == nullptr))) {
    qmcb_set_error("qmcb_ao: bad arguments");
    return QMCB_EINVAL;
  }
  if (W == 0) return 0;
  const int64_t rows = W * (one_elec ? 1 : p->sys.nelec);
  const int smem = table_doubles(p->sys) * 8;
  int64_t grid = (rows + 255) / 256;
  if (grid > (int64_t)p->sm_count * 8) grid = (int64_t)p->sm_count * 8;
  cudaStream_t st = (cudaStream_t)stream;
  auto run = [&](auto kern, double *d1, double *d2) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    kern<<<(unsigned)grid, 256, smem, st>>>(p->sys, pos, rows, ao, d1, d2);
  };
  if (dao) {
    if (p->multi_component) run(ao_kernel<5, true>, dao, d2ao); else run(ao_kernel<5, false>, dao, d2ao);
  } else {
    if (p->multi_component) run(ao_kernel<1, true>, nullptr, nullptr); else run(ao_kernel<1, false>, nullptr, nullptr);
  }
  return qmcb_cuda_rc((int)cudaGetLastError(), "operators.cu launch");
}

// ---- MolecularOrbitals.forward: [rows,nao] x [nao,nmo] ----------------------------------
__global__ void __launch_bounds__(256) mo_kernel(const double *x, const double *w, int64_t rows, int nao, int nmo,
                                                 double *out) {
  extern __shared__ double sw[];   // [nao][nmo]
  for (int i = threadIdx.x; i < nao * nmo; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int64_t total = rows * nmo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / nmo;
    const int m = (int)(i - r * nmo);
    const double *xr = x + r * nao;
    double acc = 0.0;
    for (int a = 0; a < nao; ++a) acc = fma(xr[a], sw[a * nmo + m], acc);
    out[i] = acc;
  }
}

extern "C" int qmcb_mo(const qmcb_plan *p, const double *x, int64_t rows, double *out, void *stream) {
  if (!p || !p->d_mo_full || !x || !out || rows < 0) {
    qmcb_set_error("qmcb_mo: bad arguments");
    return QMCB_EINVAL;
  }
  if (rows == 0) return 0;
  const int nao = p->sys.nao, nmo = p->sys.nmo;
  const int smem = nao * nmo * 8;
  if (smem > p->smem_optin) { qmcb_set_error("qmcb_mo: MO matrix exceeds shared memory"); return QMCB_ESMEM; }
  cudaFuncSetAttribute(mo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int64_t grid = (rows * nmo + 255) / 256;
  if (grid > (int64_t)p->sm_count * 8) grid = (int64_t)p->sm_count * 8;
  mo_kernel<<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(x, p->d_mo_full, rows, nao, nmo, out);
  return qmcb_cuda_rc((int)cudaGetLastError(), "operators.cu launch");
}

// ---- Jastrow factor + derivatives -------------------------------------------------------
__global__ void __launch_bounds__(256) jastrow_kernel(DevSys S, const double *pos, int64_t W, int deriv,
                                                      double *J, double *dJ, double *d2J) {
  extern __shared__ __align__(16) double smem[];
  Tab T;
  double *ws = stage_tables(S, smem, T);
  const int Ne = S.nelec, ne3 = 3 * Ne;
  const int TW = blockDim.x / Ne;          // walkers per CTA
  double *spos = ws;                       // [TW][3Ne]
  double *sks = spos + TW * ne3;           // [TW][Ne]
  __syncthreads();
  const int64_t ntile = (W + TW - 1) / TW;
  for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int64_t w0 = tile * TW;
    const int tw = (int)((W - w0) < TW ? (W - w0) : TW);
    for (int i = threadIdx.x; i < tw * ne3; i += blockDim.x) spos[i] = pos[w0 * ne3 + i];
    __syncthreads();
    ElecTerms o{};
    const int it = threadIdx.x;
    const int wl = it / Ne, e = it - wl * Ne;
    const bool act = it < tw * Ne;
    if (act) {
      if (deriv) electron_terms<true, false>(S, T, spos + wl * ne3, e, o);
      else electron_terms<false, false>(S, T, spos + wl * ne3, e, o);
      sks[it] = o.ks;
    }
    __syncthreads();
    if (act) {
      double ks = 0.0;
      for (int k = 0; k < Ne; ++k) ks += sks[wl * Ne + k];
      const double Jv = exp(ks);
      if (e == 0) J[w0 + wl] = Jv;
      if (deriv) {
        double *g = dJ + (w0 + wl) * ne3;   // [3][Ne]
        g[e] = o.gx * Jv; g[Ne + e] = o.gy * Jv; g[2 * Ne + e] = o.gz * Jv;
        d2J[(w0 + wl) * Ne + e] = o.lap * Jv;
      }
    }
    __syncthreads();
  }
}

extern "C" int qmcb_jastrow(const qmcb_plan *p, const double *pos, int64_t W, int which, double *J,
                            double *dJ, double *d2J, void *stream) {
  if (!p || !p->d_dbl || !pos || !J || W < 0 || ((dJ == nullptr) != (d2J == nullptr)) || which < 0 ||
      which > 3) {
    qmcb_set_error("qmcb_jastrow: bad arguments");
    return QMCB_EINVAL;
  }
  if (W == 0) return 0;
  DevSys S = p->sys;
  if (which == 1) { S.use_jen = 0; S.een_nterm = 0; }
  if (which == 2) { S.use_jee = 0; S.een_nterm = 0; }
  if (which == 3) { S.use_jee = 0; S.use_jen = 0; }
  const int Ne = S.nelec;
  if (Ne > 256) { qmcb_set_error("qmcb_jastrow: nelec > 256"); return QMCB_EINVAL; }
  const int tw = 256 / Ne;
  const int threads = ((tw * Ne + 31) / 32) * 32;
  const int smem = (table_doubles(S) + tw * 3 * Ne + tw * Ne) * 8;
  cudaFuncSetAttribute(jastrow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int64_t grid = (W + tw - 1) / tw;
  if (grid > (int64_t)p->sm_count * 8) grid = (int64_t)p->sm_count * 8;
  // blockDim must be a multiple of Ne for the (wl,e) mapping: launch tw*Ne threads rounded up,
  // the kernel derives TW from blockDim/Ne which is still tw.
  jastrow_kernel<<<(unsigned)grid, threads, smem, (cudaStream_t)stream>>>(S, pos, W, dJ != nullptr, J, dJ, d2J);
  return qmcb_cuda_rc((int)cudaGetLastError(), "operators.cu launch");
}

// ---- SlaterPooling.forward / .operator --------------------------------------------------
// one thread per (op, walker, unique occupation); generic Gauss-Jordan in per-thread local
// scratch is acceptable here (operator-level parity path, n <= 16).
#define QMCB_SLATER_NMAX 16
__global__ void __launch_bounds__(128) slater_kernel(DevSys S, const double *mo, const double *bop, int64_t nop,
                                                     int64_t W, double *dets, double *trace) {
  extern __shared__ __align__(16) double smem[];
  Tab T;
  stage_tables(S, smem, T);
  __syncthreads();
  const int nun = S.nuu + S.nud, Ne = S.nelec, nmo = S.nmo;
  const int64_t nops = bop ? nop : 1;
  const int64_t total = nops * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t op = i / W, w = i - op * W;
    double du[64], tu[64];   // nun <= 64 enforced on the host
    for (int u = 0; u < nun; ++u) {
      const bool up = u < S.nuu;
      const int n = up ? S.nup : S.ndown;
      const int *colsu = up ? T.ucu() + u * S.nup : T.ucd() + (u - S.nuu) * S.ndown;
      const double *A = mo + (w * Ne + (up ? 0 : S.nup)) * nmo;
      const double *B = bop ? bop + ((op * W + w) * Ne + (up ? 0 : S.nup)) * nmo : nullptr;
      double m[QMCB_SLATER_NMAX * 2 * QMCB_SLATER_NMAX];
      const int nr = B ? n : 0, ldw = n + nr;
      for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) {
          const int col = T.used()[colsu[c]];
          m[r * ldw + c] = A[r * nmo + col];
          if (nr) m[r * ldw + n + c] = B[r * nmo + col];
        }
      double det = n ? gauss_jordan(n, nr, m, 1) : 1.0, tr = 0.0;
      for (int r = 0; r < nr; ++r) tr += m[r * ldw + n + r];
      du[u] = det; tu[u] = tr;
    }
    for (int c = 0; c < S.nconf; ++c) {
      const int iu = T.ciu()[c], id = S.nuu + T.cid()[c];
      if (op == 0 && dets) dets[w * S.nconf + c] = du[iu] * du[id];
      if (trace && bop) trace[(op * W + w) * S.nconf + c] = tu[iu] + tu[id];
    }
  }
}

extern "C" int qmcb_slater(const qmcb_plan *p, const double *mo, const double *bop, int64_t nop, int64_t W,
                           double *dets, double *trace, void *stream) {
  if (!p || !p->d_dbl || !mo || W < 0 || (bop && !trace)) {
    qmcb_set_error("qmcb_slater: bad arguments");
    return QMCB_EINVAL;
  }
  if (W == 0) return 0;
  const DevSys &S = p->sys;
  const int n = S.nup > S.ndown ? S.nup : S.ndown;
  if (n > QMCB_SLATER_NMAX || S.nuu + S.nud > 64) {
    qmcb_set_error("qmcb_slater: operator-level path supports spin blocks <= 16 and <= 64 unique occupations");
    return QMCB_EINVAL;
  }
  const int smem = table_doubles(S) * 8;
  cudaFuncSetAttribute(slater_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int64_t total = (bop ? nop : 1) * W;
  int64_t grid = (total + 127) / 128;
  if (grid > (int64_t)p->sm_count * 8) grid = (int64_t)p->sm_count * 8;
  slater_kernel<<<(unsigned)grid, 128, smem, (cudaStream_t)stream>>>(S, mo, bop, nop, W, dets, trace);
  return qmcb_cuda_rc((int)cudaGetLastError(), "operators.cu launch");
}

// ---- energy statistics ------------------------------------------------------------------
// stage 1: per-CTA partial (sum, sum of squares about a pivot-free accumulation in FP64,
// finite count, non-finite count) in fixed order; stage 2: one CTA adds the partials in
// index order -> bitwise reproducible for a given W.
#define STATS_CTAS 296
__global__ void __launch_bounds__(256) stats_stage1(const double *e, int64_t W, double *part) {
  __shared__ double sh[4][256];
  double s = 0, s2 = 0, nf = 0, nb = 0;
  // contiguous chunk per CTA, strided inside the CTA: fixed order independent of scheduling
  const int64_t chunk = (W + gridDim.x - 1) / gridDim.x;
  const int64_t b = chunk * blockIdx.x;
  const int64_t en = b + chunk < W ? b + chunk : W;
  for (int64_t i = b + threadIdx.x; i < en; i += blockDim.x) {
    const double v = e[i];
    if (isfinite(v)) { s += v; s2 += v * v; nf += 1.0; }
    else nb += 1.0;
  }
  sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = s2; sh[2][threadIdx.x] = nf; sh[3][threadIdx.x] = nb;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off)
      for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x < 4) part[blockIdx.x * 4 + threadIdx.x] = sh[threadIdx.x][0];
}

// one warp per quantity: lane-strided partial sums, then a fixed-order butterfly
__global__ void stats_stage2(const double *part, int n, double *out4) {
  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s = 0;
  for (int i = lane; i < n; i += 32) s += part[i * 4 + q];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out4[q] = s;
}

extern "C" int64_t qmcb_stats_workspace_bytes(int64_t) {
  static_assert(STATS_CTAS <= QMCB_STATS_MAX_PARTIALS, "workspace");
  return (int64_t)QMCB_STATS_MAX_PARTIALS * 4 * 8;
}

int qmcb_stats_finish(const double *part, int n, double *out4, void *stream) {
  stats_stage2<<<1, 128, 0, (cudaStream_t)stream>>>(part, n, out4);
  return qmcb_cuda_rc((int)cudaGetLastError(), "operators.cu stats_stage2");
}

extern "C" int qmcb_energy_stats(const double *eloc, int64_t W, double *out4, void *workspace, void *stream) {
  if (!eloc || !out4 || !workspace || W < 0) {
    qmcb_set_error("qmcb_energy_stats: bad arguments");
    return QMCB_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  stats_stage1<<<STATS_CTAS, 256, 0, st>>>(eloc, W, (double *)workspace);
  stats_stage2<<<1, 128, 0, st>>>((const double *)workspace, STATS_CTAS, out4);
  return qmcb_cuda_rc((int)cudaGetLastError(), "operators.cu launch");
}

// ---- FP64 pipe probes ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_probe(int64_t iters, double *sink) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double b = 0.999999999, c = 1e-12;
  for (int64_t i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (r == 123.456) sink[0] = r;
}

__global__ void __launch_bounds__(256) dmma_probe(int64_t iters, double *sink) {
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  const double a = 1e-3 * (threadIdx.x & 7), b = 1e-3 * (threadIdx.x >> 3);
  for (int64_t i = 0; i < iters; ++i) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
  }
  const double r = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
  if (r == 123.456) sink[0] = r;
}

// both at once: does the DMMA sub-pipe run concurrently with the FP64 vector pipe?
__global__ void __launch_bounds__(256) mixed_probe(int64_t iters, double *sink) {
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double a = 1e-3 * (threadIdx.x & 7), b = 1e-3 * (threadIdx.x >> 3), bb = 0.999999999, cc = 1e-12;
  for (int64_t i = 0; i < iters; ++i) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
    a0 = fma(a0, bb, cc); a1 = fma(a1, bb, cc);
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
    a2 = fma(a2, bb, cc); a3 = fma(a3, bb, cc);
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
    a4 = fma(a4, bb, cc); a5 = fma(a5, bb, cc);
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
    a6 = fma(a6, bb, cc); a7 = fma(a7, bb, cc);
  }
  const double r = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1] + a0 + a1 + a2 + a3 + a4 + a5 +
                   a6 + a7;
  if (r == 123.456) sink[0] = r;
}

extern "C" int qmcb_fp64_probe(int kind, int64_t iters, double *sink, double *flops, void *stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ctas = sms * 8, threads = 256;
  if (kind == 0) {
    dfma_probe<<<ctas, threads, 0, (cudaStream_t)stream>>>(iters, sink);
    if (flops) *flops = 2.0 * 8.0 * (double)iters * ctas * threads;
  } else if (kind == 1) {
    dmma_probe<<<ctas, threads, 0, (cudaStream_t)stream>>>(iters, sink);
    // one m8n8k4 = 8*8*4 FMA = 512 flop per warp
    if (flops) *flops = 512.0 * 4.0 * (double)iters * ctas * (threads / 32);
  } else {
    mixed_probe<<<ctas, threads, 0, (cudaStream_t)stream>>>(iters, sink);
    // per iteration and warp: 4 DMMA (2048 flop) + 8 DFMA x 32 lanes (512 flop)
    if (flops) *flops = (2048.0 + 512.0) * (double)iters * ctas * (threads / 32);
  }
  return qmcb_cuda_rc((int)cudaGetLastError(), "operators.cu launch");
}
