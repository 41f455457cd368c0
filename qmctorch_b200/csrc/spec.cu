// Structure-specialised kernels: source generation, NVRTC compilation, module cache, launch.
//
// For wave functions that run one walker per thread (small systems, closed-form determinants,
// gto_pure radial functions) the basis walk, the AO -> MO contraction, the determinants and the CI
// sum are emitted as straight-line CUDA for that one structure and compiled with NVRTC for the
// device's architecture; see spec_kernel.cuh for the device side and the rationale.  Parameters
// stay run-time data (kernel-parameter block), so qmcb_plan_update() never recompiles.
//
// libnvrtc and libcuda are opened with dlopen at first use: libqmcb.so itself links against
// neither, loads on machines without a driver (CPU tests), and falls back to the generic CUDA
// kernels when NVRTC is not installed (QMCB_JIT=0 forces that; QMCB_JIT=2 makes a missing or
// failing NVRTC an error).  There is no CPU path here either.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "fused_args.h"
#include "plan.h"
#include "spec.h"

#include "_spec_embed.inc"   // SPEC_SRC_ARGS, SPEC_SRC_DEVICE, SPEC_SRC_PHILOX, SPEC_SRC_KERNEL (build.py)

namespace {

// CTA shape of the specialised kernels (launch bounds -> register budget).  QMCB_SPEC_THREADS /
// QMCB_SPEC_MINB override them for tuning experiments.
int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return (e && atoi(e) > 0) ? atoi(e) : dflt;
}
const int SPEC_THREADS = env_int("QMCB_SPEC_THREADS", 128);
const int SPEC_MINB = env_int("QMCB_SPEC_MINB", 4);
// warp-tile kernels (spec_tile.cuh): CTA shape and where the MO weights live
const int TILE_THREADS = env_int("QMCB_TILE_THREADS", 256);
const int TILE_MINB = env_int("QMCB_TILE_MINB", 0);          // 0: chosen from the accumulator count
const int TILE_MOW = env_int("QMCB_TILE_MOW", 0);            // 0: auto, 1: constant bank, 2: shared memory
constexpr int TILE_MAX_VALUES = 3800;  // doubles in the parameter block (32 KB of kernel parameters on sm_70+)
constexpr size_t TILE_SMEM_BUDGET = 110 * 1024;
// defaults of the kernel's tuning switches (must match the #ifndef defaults in spec_kernel.cuh)
constexpr int SPEC_DEFAULT_PREFETCH = 0;
constexpr int SPEC_DEFAULT_PREFETCH_ELOC = 1;
constexpr int SPEC_DEFAULT_MOW_SMEM = 0;
constexpr size_t SPEC_SMEM_BUDGET = 100 * 1024;   // per CTA: at least two CTAs per SM
constexpr int SPEC_MAX_VALUES = 960;   // doubles in the parameter block (8 KB; sm_70+ take 32 KB of kernel parameters)

// ---------------------------------------------------------------------------------------
// dynamic bindings
// ---------------------------------------------------------------------------------------
typedef struct _nvrtcProgram *nvrtcProgram;
typedef struct CUmod_st *CUmodule;
typedef struct CUfunc_st *CUfunction;
typedef struct CUstream_st *CUstream;

struct Dyn {
  bool tried = false, ok = false;
  std::string why;
  // nvrtc
  int (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *);
  int (*CompileProgram)(nvrtcProgram, int, const char *const *);
  int (*GetProgramLogSize)(nvrtcProgram, size_t *);
  int (*GetProgramLog)(nvrtcProgram, char *);
  int (*GetCUBINSize)(nvrtcProgram, size_t *);
  int (*GetCUBIN)(nvrtcProgram, char *);
  int (*DestroyProgram)(nvrtcProgram *);
  int (*GetPTXSize)(nvrtcProgram, size_t *) = nullptr;   // optional (QMCB_JIT_DUMP)
  int (*GetPTX)(nvrtcProgram, char *) = nullptr;
  // driver
  bool drv = false;
  int (*ModuleLoadData)(CUmodule *, const void *);
  int (*ModuleGetFunction)(CUfunction *, CUmodule, const char *);
  int (*FuncSetAttribute)(CUfunction, int, int);
  int (*OccupancyMaxActiveBlocks)(int *, CUfunction, int, size_t);
  int (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                      void **, void **);
  int (*GetErrorString)(int, const char **);
};

Dyn &dyn() {
  static Dyn d;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (d.tried) return d;
  d.tried = true;
  void *h = nullptr;
  const char *names[] = {getenv("QMCB_NVRTC"), "libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                         "/usr/local/cuda/lib64/libnvrtc.so"};
  for (const char *n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (h) break;
  }
  if (!h) { d.why = "libnvrtc not found (set QMCB_NVRTC=/path/to/libnvrtc.so)"; return d; }
#define BIND(field, sym) *(void **)(&d.field) = dlsym(h, sym); if (!d.field) { d.why = std::string("missing ") + sym; return d; }
  BIND(CreateProgram, "nvrtcCreateProgram") BIND(CompileProgram, "nvrtcCompileProgram")
  BIND(GetProgramLogSize, "nvrtcGetProgramLogSize") BIND(GetProgramLog, "nvrtcGetProgramLog")
  BIND(GetCUBINSize, "nvrtcGetCUBINSize") BIND(GetCUBIN, "nvrtcGetCUBIN") BIND(DestroyProgram, "nvrtcDestroyProgram")
#undef BIND
  *(void **)(&d.GetPTXSize) = dlsym(h, "nvrtcGetPTXSize");
  *(void **)(&d.GetPTX) = dlsym(h, "nvrtcGetPTX");
  d.ok = true;
  void *c = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
  if (c) {
#define BINDC(field, sym) *(void **)(&d.field) = dlsym(c, sym);
    BINDC(ModuleLoadData, "cuModuleLoadData") BINDC(ModuleGetFunction, "cuModuleGetFunction")
    BINDC(FuncSetAttribute, "cuFuncSetAttribute")
    BINDC(OccupancyMaxActiveBlocks, "cuOccupancyMaxActiveBlocksPerMultiprocessor")
    BINDC(LaunchKernel, "cuLaunchKernel") BINDC(GetErrorString, "cuGetErrorString")
#undef BINDC
    d.drv = d.ModuleLoadData && d.ModuleGetFunction && d.FuncSetAttribute && d.OccupancyMaxActiveBlocks &&
            d.LaunchKernel;
  }
  return d;
}

// QMCB_SPEC_DEFS="-DSPEC_PREFETCH=1 -DSPEC_EUNROLL=2": tuning switches of spec_kernel.cuh
std::vector<std::string> extra_defs() {
  std::vector<std::string> out;
  const char *e = getenv("QMCB_SPEC_DEFS");
  if (!e) return out;
  std::istringstream is(e);
  std::string tok;
  while (is >> tok) out.push_back(tok);
  return out;
}
int def_value(const char *name, int dflt) {
  for (auto &d : extra_defs()) {
    const std::string key = std::string("-D") + name + "=";
    if (d.compare(0, key.size(), key) == 0) return atoi(d.c_str() + key.size());
  }
  return dflt;
}

// replicas of the exp table in the one-walker-per-thread kernels (power of two <= 16; device.cuh: exp_core)
int spec_etab_rep() {
  int r = env_int("QMCB_SPEC_ETAB_REP", 1);   // measured (LiH E_L, ms): 1 -> 0.136, 4 / 8 -> 0.141, 16 -> 0.142: the conflicts of the unreplicated table are not the limiter
  int p2 = 1;
  while (p2 * 2 <= r && p2 < 16) p2 *= 2;
  return p2;
}

int jit_level() {
  const char *e = getenv("QMCB_JIT");
  return e ? atoi(e) : 1;
}

// ---------------------------------------------------------------------------------------
// the structure walk: emits the program text and/or collects the parameter values (same order)
// ---------------------------------------------------------------------------------------
struct Layout { int nv = 0, off_atom = 0, off_mow = 0, off_ci = 0; };

// kind of specialised kernel a plan gets: one walker per thread (small systems, spec_kernel.cuh) or
// warp-owned walker tiles (mid-size / large systems, spec_tile.cuh)
enum Kind { KIND_NONE = 0, KIND_THREAD = 1, KIND_TILE = 2 };

bool tile_mow_smem(const DevSys &S) {
  if (TILE_MOW == 1) return false;
  if (TILE_MOW == 2) return true;
  // constant-bank operands cost no instruction and no register; measured on C4H6 (94 x 15 weights, 12 KB):
  // E_L 0.374 ms per 2e4 walkers against 0.421 ms with LDS.128 broadcasts.  Shared memory only when the
  // weights would not fit the 32 KB kernel-parameter space next to the primitive constants.
  return S.nao * S.nmu + 6 * S.nprim + S.ncomp + 4 * S.natom > TILE_MAX_VALUES;
}
int tile_threads(const DevSys &S) {
  // the three-body factor tables need ~12 KB per warp: smaller CTAs keep 12 warps per SM resident
  if (!getenv("QMCB_TILE_THREADS") && S.een_nterm > 0) return 128;
  return TILE_THREADS;
}
int tile_minb(const DevSys &S) {
  if (TILE_MINB > 0) return TILE_MINB;
  // 256 threads x 2 = 16 warps per SM at 128 registers.  Measured (E_L, ms): C4H6 2e4 walkers 0.400 against
  // 0.442 for 128 x 3 (168 registers, 12 warps) and 0.479 for 128 x 5; H2O cas(4,4) 1e5 walkers 0.173 / 0.177
  return S.een_nterm > 0 ? 3 : 2;
}

bool eligible(const qmcb_plan *p, std::string *why) {
  const DevSys &S = p->sys;
  const int nbig = S.nup > S.ndown ? S.nup : S.ndown;
  const char *w = nullptr;
  if (S.een_nterm > 0) w = "three-body Jastrow";
  else if (S.nmu > 8) w = "more than 8 occupied MO columns";
  else if (nbig > 3) w = "spin block larger than 3x3";
  else if (S.nelec > 8 || S.nelec < 1) w = "more than 8 electrons";
  else if (S.nuu + S.nud > 16 || S.nconf > 64) w = "too many determinants";
  else if (((size_t)QMCB_ETAB * spec_etab_rep() + (size_t)SPEC_THREADS * ((10 * S.nelec + 1 + 2 * S.nelec * S.nmu) | 1)) * sizeof(double) > SPEC_SMEM_BUDGET)
    w = "per-thread slices exceed the shared-memory budget";
  if (w) { if (why) *why = w; return false; }
  return true;
}

size_t tile_smem_doubles(const DevSys &S, int mode);

bool eligible_tile(const qmcb_plan *p, std::string *why) {
  const DevSys &S = p->sys;
  const int nbig = S.nup > S.ndown ? S.nup : S.ndown;
  const char *w = nullptr;
  if (S.nelec > 32 || S.nelec < 1) w = "more than 32 electrons (a walker must fit one warp)";
  else if (S.nmu > 16) w = "more than 16 occupied MO columns";
  else if (nbig > 16) w = "spin block larger than 16x16";
  else if (S.nconf > 4096 || S.nuu + S.nud > 256) w = "too many determinants";
  else if (tile_smem_doubles(S, MODE_ELOC) * sizeof(double) > TILE_SMEM_BUDGET) w = "tables exceed the shared-memory budget";
  if (w) { if (why) *why = w; return false; }
  return true;
}

Kind choose_kind(const qmcb_plan *p, std::string *why) {
  const char *force = getenv("QMCB_SPEC_KIND");     // tuning / tests: "thread" | "tile"
  std::string w1, w2;
  const bool t1 = eligible(p, &w1), t2 = eligible_tile(p, &w2);
  if (force && !strcmp(force, "tile")) { if (!t2 && why) *why = w2; return t2 ? KIND_TILE : KIND_NONE; }
  if (force && !strcmp(force, "thread")) { if (!t1 && why) *why = w1; return t1 ? KIND_THREAD : KIND_NONE; }
  if (t1) return KIND_THREAD;
  if (t2) return KIND_TILE;
  if (why) *why = w1 + "; " + w2;
  return KIND_NONE;
}

bool walk(const qmcb_plan *p, std::string *code, std::vector<double> *vals, Layout &L, Kind kind = KIND_THREAD) {
  const DevSys &S = p->sys;
  const std::vector<double> &hd = p->hd;
  const std::vector<int> &hi = p->hi;
  std::ostringstream o;
  o.precision(17);
  int nv = 0;
  auto push = [&](double x) { if (vals) vals->push_back(x); return nv++; };
  L.off_atom = nv;
  for (int i = 0; i < 4 * S.natom; ++i) push(hd[S.o_atoms + i]);
  const double *rec = hd.data() + S.o_stream;
  auto ints = [](double d, int &lo, int &hi2) {
    union { double d; int i[2]; } u;
    u.d = d; lo = u.i[0]; hi2 = u.i[1];
  };
  o << "template <int MODE, int NCH>\n__device__ __forceinline__ void spec_aos(const SpecParams &P, const double *et, "
       "const double *mw, double ex, double ey, double ez, const FoldJ fj, double &ven, double (&acc)[NCH][SPEC_NMUP]) {\n";
  int nexp = 0;
  // spec_aos_grad (MODE_BWD_ALL): second walk over the primitives for the basis-parameter gradients
  std::ostringstream og;
  og.precision(17);
  og << "template <int MODE>\n__device__ __forceinline__ void spec_aos_grad(const SpecParams &P, const double *et, "
        "const double *gao, double ex, double ey, double ez, double (&aC)[SPEC_NBAS], double (&aE)[SPEC_NBAS]) {\n";
  int sidx = 0;      // running shell index (plan tables: o_spo, o_sco, o_pfo)
  auto mono = [](int kx, int ky, int kz) {
    std::string t;
    const char *nm[3] = {"x", "y", "z"};
    const int k[3] = {kx, ky, kz};
    for (int c = 0; c < 3; ++c)
      for (int i = 0; i < k[c]; ++i) t += (t.empty() ? "" : " * ") + std::string(nm[c]);
    return t.empty() ? std::string("1.0") : t;
  };
  for (int A = 0; A < S.natom; ++A) {
    const int ns = hi[S.o_ash + A + 1] - hi[S.o_ash + A];
    const int oa = L.off_atom + 4 * A;
    const int rt = S.radial_type;
    if (ns == 0) {   // no basis functions on this atom: only its potential
      o << "  if (MODE == MODE_ELOC) {  // atom " << A << "\n    const double x = ex - spec_pv<MODE, " << oa
        << ">(), y = ey - spec_pv<MODE, " << oa + 1 << ">(), z = ez - spec_pv<MODE, " << oa + 2
        << ">();\n    spec_ven<MODE, " << oa + 3 << ", 0>(x * x + y * y + z * z, 0.0, ven);\n  }\n";
      continue;
    }
    o << "  {  // atom " << A << "\n    const double x = ex - spec_pv<MODE, " << oa << ">(), y = ey - spec_pv<MODE, " << oa + 1
      << ">(), z = ez - spec_pv<MODE, " << oa + 2 << ">();\n    const double r2 = x * x + y * y + z * z;\n"
      << "    const double gd = spec_gd<MODE, " << rt << ">(fj, x, y, z);\n";
    const bool with_n = rt == QMCB_GTO || rt == QMCB_STO;
    const bool arg_r2 = rt == QMCB_GTO_PURE || rt == QMCB_GTO;
    if (rt != QMCB_GTO_PURE) o << "    const double rinv = fast_rsqrt(r2), r = r2 * rinv;\n";
    else o << "    const double rinv = 0.0;\n";
    o << "    spec_ven<MODE, " << oa + 3 << ", " << rt << ">(r2, rinv, ven);\n";
    og << "  {  // atom " << A << "\n    const double x = ex - spec_pv<MODE, " << oa << ">(), y = ey - spec_pv<MODE, " << oa + 1
       << ">(), z = ez - spec_pv<MODE, " << oa + 2 << ">();\n    const double r2 = x * x + y * y + z * z;\n";
    if (rt != QMCB_GTO_PURE) og << "    const double rinv = fast_rsqrt(r2), r = r2 * rinv;\n";
    og << "    const double dfac = " << (arg_r2 ? "-r2" : "-r") << ";\n";
    // exponentials of this atom by exponent value: primitives with bitwise equal exponents (the s
    // and p functions of an SP shell) share one; a parameter update that separates them changes the
    // generated text and therefore selects (compiles) another module
    std::map<uint64_t, int> known;
    for (int s = 0; s < ns; ++s) {
      int nprim, ngrp;
      ints(rec[0], nprim, ngrp);
      rec += 2;
      std::vector<int> ev(nprim), iv(nprim);
      const double *prec = rec;
      for (int q = 0; q < nprim; ++q, rec += with_n ? 4 : 2) {
        // derived per-primitive constants, see spec_prim* (spec_kernel.cuh)
        const double a = rec[0], c = rec[1];
        iv[q] = push(-a);
        push(c);
        if (rt == QMCB_GTO_PURE) { push(-2.0 * a * c); push(-6.0 * a * c); push(4.0 * a * a * c); }
        else if (rt == QMCB_STO_PURE) { push(-a * c); push(a * a * c); push(0.0); }
        else { push(a); push(0.0); push(0.0); }
        uint64_t bits;
        memcpy(&bits, &a, sizeof(bits));
        auto it = known.find(bits);
        if (it == known.end()) {
          it = known.emplace(bits, nexp++).first;
          o << "    const double e" << it->second << " = spec_exp<MODE, " << iv[q] << ">(P, et, " << (arg_r2 ? "r2" : "r") << ");\n";
          og << "    const double e" << it->second << " = spec_exp<MODE, " << iv[q] << ">(P, et, " << (arg_r2 ? "r2" : "r") << ");\n";
        }
        ev[q] = it->second;
      }
      {
        // basis-parameter gradients of this shell: R_q = r^n e_q, dR_q/dalpha = -(r^2 | r) R_q; per component
        // X = Y_k Gao[a_k]; accumulators by FLAT primitive index (plan tables o_pfo / o_pflat)
        const int k0 = hi[S.o_sco + sidx], ncomp_s = hi[S.o_sco + sidx + 1] - k0, pf0 = hi[S.o_pfo + sidx];
        og << "    {\n";
        for (int q = 0; q < nprim; ++q) {
          og << "      const double R" << q << " = e" << ev[q];
          if (with_n) {
            const int pn = (int)prec[4 * q + 2];
            for (int i = 0; i < pn; ++i) og << " * r";
          }
          og << ", D" << q << " = dfac * R" << q << ";\n";
        }
        for (int k = 0; k < ncomp_s; ++k) {
          const int kk = hi[S.o_ck + k0 + k], ao = hi[S.o_cao + k0 + k];
          og << "      { const double X = gao[" << ao << "] * (" << mono(kk & 255, (kk >> 8) & 255, (kk >> 16) & 255) << ");\n";
          for (int q = 0; q < nprim; ++q) {
            const int flat = hi[S.o_pflat + pf0 + q * ncomp_s + k];
            og << "        aC[" << flat << "] = fma(R" << q << ", X, aC[" << flat << "]); aE[" << flat << "] = fma(D" << q
               << ", X, aE[" << flat << "]);\n";
          }
          og << "      }\n";
        }
        og << "    }\n";
      }
      o << "    {\n      double S0 = 0.0, S1 = 0.0, S2 = 0.0, T2 = 0.0;\n";
      for (int q = 0; q < nprim; ++q) {
        if (rt == QMCB_GTO_PURE)
          o << "      spec_prim<MODE, " << (q == 0 ? "true" : "false") << ", " << iv[q] << ">(e" << ev[q] << ", S0, S1, T2);\n";
        else if (rt == QMCB_STO_PURE)
          o << "      spec_prim_sto_pure<MODE, " << iv[q] << ">(e" << ev[q] << ", S0, S1, S2);\n";
        else
          o << "      spec_prim_power<MODE, " << iv[q] << ", " << (rt == QMCB_GTO ? "true" : "false") << ", "
            << (int)prec[4 * q + 2] << ">(e" << ev[q] << ", r2, r, rinv, S0, S1, S2);\n";
      }
      o << "      const double Wf = spec_shell_end<MODE, " << rt << ">(r2, rinv, gd, fj.lp, S0, S1, S2, T2);\n";
      for (int g = 0; g < ngrp; ++g, rec += 2) {
        int kk, ao;
        ints(rec[0], kk, ao);
        const int is = push(rec[1]);
        if (kk == 0) o << "      spec_s<MODE, " << ao << ", " << is << ">";
        else if (kk == (1 << 24)) o << "      spec_p<MODE, " << ao << ", " << is << ">";
        else o << "      spec_g<MODE, " << ao << ", " << is << ", " << kk << ">";
        o << "(mw, x, y, z, S0, S1, Wf, fj, acc);\n";
      }
      o << "    }\n";
      ++sidx;
    }
    o << "  }\n";
    og << "  }\n";
  }
  o << "}\n\n";
  og << "}\n\n";
  if ((int)(rec - (hd.data() + S.o_stream)) != 2 * S.nrec) return false;   // walked exactly the program
  // MO weights of the occupied columns only (the generic kernels pad the column count to a power of
  // two; a specialised kernel is compiled for the exact count)
  L.off_mow = nv;
  if (kind == KIND_TILE) {
    // warp-tile kernels: weights in the parameter block only when they are read as constant-bank
    // operands; CI coefficients and the index tables are staged from the plan's device tables
    if (!tile_mow_smem(S))
      for (int a = 0; a < S.nao; ++a)
        for (int j = 0; j < S.nmu; ++j) push(hd[S.o_mow + a * S.nmup + j]);
    L.off_ci = nv;
    if (nv == 0) push(0.0);
    L.nv = nv;
    if (code) *code = o.str();
    return true;
  }
  for (int a = 0; a < S.nao; ++a)
    for (int j = 0; j < S.nmu; ++j) push(hd[S.o_mow + a * S.nmup + j]);
  L.off_ci = nv;
  for (int c = 0; c < S.nconf; ++c) push(hd[S.o_ci + c]);
  // basis-parameter gradient factors by flat primitive: norm (-> d/d bas_coeffs), coef_q * component scale (-> d/d bas_exp)
  const int off_gfac = nv;
  {
    std::vector<double> fc(S.nbas, 0.0), fe(S.nbas, 0.0);
    for (int sh = 0; sh < S.nshell; ++sh) {
      const int q0 = hi[S.o_spo + sh], nq = hi[S.o_spo + sh + 1] - q0;
      const int k0 = hi[S.o_sco + sh], nk = hi[S.o_sco + sh + 1] - k0, pf0 = hi[S.o_pfo + sh];
      for (int q = 0; q < nq; ++q)
        for (int k = 0; k < nk; ++k) {
          const int flat = hi[S.o_pflat + pf0 + q * nk + k];
          fc[flat] = hd[S.o_fnorm + flat];
          fe[flat] = hd[S.o_coef + q0 + q] * hd[S.o_cscale + k0 + k];
        }
    }
    for (int f = 0; f < S.nbas; ++f) push(fc[f]);
    for (int f = 0; f < S.nbas; ++f) push(fe[f]);
  }
  L.nv = nv;
  o << og.str();
  o << "template <int MODE>\n__device__ __forceinline__ void spec_grad_scale(double (&aC)[SPEC_NBAS], double (&aE)[SPEC_NBAS]) {\n";
  for (int f = 0; f < S.nbas; ++f)
    o << "  aC[" << f << "] *= spec_pv<MODE, " << off_gfac + f << ">(); aE[" << f << "] *= spec_pv<MODE, " << off_gfac + S.nbas + f
      << ">();\n";
  o << "}\n\n";
  // determinants of the unique occupations
  const int nun = S.nuu + S.nud;
  o << "template <bool WB>\n__device__ __forceinline__ void spec_dets(const double *A, const double *B, double (&det)["
    << nun << "], double (&tr)[" << nun << "]) {\n";
  for (int u = 0; u < nun; ++u) {
    const bool up = u < S.nuu;
    const int n = up ? S.nup : S.ndown;
    if (n == 0) { o << "  det[" << u << "] = 1.0; tr[" << u << "] = 0.0;\n"; continue; }
    const int *cols = hi.data() + (up ? S.o_ucu + u * S.nup : S.o_ucd + (u - S.nuu) * S.ndown);
    const int row0 = (up ? 0 : S.nup) * S.nmu;
    o << "  { const int cols[" << n << "] = {";
    for (int j = 0; j < n; ++j) o << (j ? ", " : "") << cols[j];
    o << "}; det_trace_small(" << n << ", A + " << row0 << ", B + " << row0 << ", SPEC_NMUP, cols, WB, det[" << u
      << "], tr[" << u << "]); }\n";
  }
  o << "}\n\n";
  o << "template <int MODE>\n__device__ __forceinline__ void spec_ci(const double *mw, const double (&det)[" << nun
    << "], const double (&tr)[" << nun << "], double &sig, double &ksig) {\n  constexpr bool WB = MODE == MODE_ELOC;\n"
       "  sig = 0.0; ksig = 0.0;\n";
  for (int c = 0; c < S.nconf; ++c) {
    const int iu = hi[S.o_ciu + c], id = S.nuu + hi[S.o_cid + c];
    o << "  { const double d = (SPEC_MOW_SMEM ? mw[" << L.off_ci - L.off_mow + c << "] : spec_pv<MODE, " << L.off_ci + c
      << ">()) * det[" << iu << "] * det[" << id
      << "]; sig += d; if (WB) ksig += d * (tr[" << iu << "] + tr[" << id << "]); }\n";
  }
  o << "}\n";
  // gradient assembly (MODE_GRAD): inverse of every unique spin block, its CI weight, then per
  // electron and direction the trace against the derivative rows
  o << "\ntemplate <int MODE>\n__device__ __forceinline__ void spec_grad(const double *mw, const double *A, const double *G, "
       "const double *jv, double J, int pdf, double *out) {\n  constexpr int NE = SPEC_NE, NM = SPEC_NMUP, NENM = SPEC_NE * SPEC_NMUP;\n"
    << "  double det[" << nun << "];\n";
  for (int u = 0; u < nun; ++u) {
    const bool up = u < S.nuu;
    const int n = up ? S.nup : S.ndown;
    if (n == 0) { o << "  det[" << u << "] = 1.0;\n"; continue; }
    const int *cols = hi.data() + (up ? S.o_ucu + u * S.nup : S.o_ucd + (u - S.nuu) * S.ndown);
    o << "  double inv" << u << "[" << n * n << "];\n  { const int cols[" << n << "] = {";
    for (int j = 0; j < n; ++j) o << (j ? ", " : "") << cols[j];
    o << "}; det[" << u << "] = inverse_small(" << n << ", A + " << (up ? 0 : S.nup) * S.nmu << ", NM, cols, inv" << u
      << ", 1); }\n";
  }
  o << "  double sig = 0.0;\n";
  for (int u = 0; u < nun; ++u) o << "  double cw" << u << " = 0.0;\n";
  for (int c = 0; c < S.nconf; ++c) {
    const int iu = hi[S.o_ciu + c], id = S.nuu + hi[S.o_cid + c];
    o << "  { const double ci = SPEC_MOW_SMEM ? mw[" << L.off_ci - L.off_mow + c << "] : spec_pv<MODE, " << L.off_ci + c
      << ">(); sig += ci * det[" << iu << "] * det[" << id << "]; cw" << iu << " += ci * det[" << id << "]; cw" << id
      << " += ci * det[" << iu << "]; }\n";
  }
  for (int u = 0; u < nun; ++u) o << "  cw" << u << " *= det[" << u << "];\n";
  o << "  const double f = pdf ? 2.0 * sig * J : 1.0;\n";
  for (int e = 0; e < S.nelec; ++e) {
    const bool up = e < S.nup;
    const int n = up ? S.nup : S.ndown, el = up ? e : e - S.nup;
    o << "  {  // electron " << e << "\n    double gx = 0.0, gy = 0.0, gz = 0.0;\n";
    const int u0 = up ? 0 : S.nuu, u1 = up ? S.nuu : nun;
    for (int u = u0; u < u1; ++u) {
      const int *cols = hi.data() + (up ? S.o_ucu + u * S.nup : S.o_ucd + (u - S.nuu) * S.ndown);
      o << "    if (cw" << u << " != 0.0) {\n      double tx = 0.0, ty = 0.0, tz = 0.0;\n";
      for (int j = 0; j < n; ++j) {
        const int at = e * S.nmu + cols[j];
        o << "      { const double iv = inv" << u << "[" << j * n + el << "]; tx = fma(iv, G[" << at << "], tx); ty = fma(iv, G[NENM + "
          << at << "], ty); tz = fma(iv, G[2 * NENM + " << at << "], tz); }\n";
      }
      o << "      gx = fma(cw" << u << ", tx, gx); gy = fma(cw" << u << ", ty, gy); gz = fma(cw" << u << ", tz, gz);\n    }\n";
    }
    o << "    out[" << 3 * e << "] = f * J * (gx + jv[" << e << "] * sig); out[" << 3 * e + 1 << "] = f * J * (gy + jv[NE + " << e
      << "] * sig); out[" << 3 * e + 2 << "] = f * J * (gz + jv[2 * NE + " << e << "] * sig);\n  }\n";
  }
  o << "}\n";
  // parameter-gradient backward (MODE_BWD, spec_kernel.cuh: spec_bwd_body): inverses and CI weights as in
  // spec_grad, then G[e][m] = w J sum_u C_u inv_u[j(m)][e] and dW[a][m] += AO[e][a] G[e][m]
  o << "\ntemplate <int MODE>\n__device__ __forceinline__ void spec_bwd(const double *A, const double *sao, double wJ, "
       "double (&dW)[SPEC_NAO][SPEC_NMUP], double (&dci)[SPEC_NCONF], double &sig_out) {\n"
       "  constexpr int NM = SPEC_NMUP, NAO = SPEC_NAO;\n"
    << "  double det[" << nun << "];\n";
  for (int u = 0; u < nun; ++u) {
    const bool up = u < S.nuu;
    const int n = up ? S.nup : S.ndown;
    if (n == 0) { o << "  det[" << u << "] = 1.0;\n"; continue; }
    const int *cols = hi.data() + (up ? S.o_ucu + u * S.nup : S.o_ucd + (u - S.nuu) * S.ndown);
    o << "  double inv" << u << "[" << n * n << "];\n  { const int cols[" << n << "] = {";
    for (int j = 0; j < n; ++j) o << (j ? ", " : "") << cols[j];
    o << "}; det[" << u << "] = inverse_small(" << n << ", A + " << (up ? 0 : S.nup) * S.nmu << ", NM, cols, inv" << u
      << ", 1); }\n";
  }
  o << "  double sig = 0.0;\n";
  for (int u = 0; u < nun; ++u) o << "  double cw" << u << " = 0.0;\n";
  for (int c = 0; c < S.nconf; ++c) {
    const int iu = hi[S.o_ciu + c], id = S.nuu + hi[S.o_cid + c];
    o << "  { const double ci = spec_pv<MODE, " << L.off_ci + c << ">(), dd = det[" << iu << "] * det[" << id
      << "]; sig = fma(ci, dd, sig); cw" << iu << " = fma(ci, det[" << id << "], cw" << iu << "); cw" << id
      << " = fma(ci, det[" << iu << "], cw" << id << "); dci[" << c << "] = fma(wJ, dd, dci[" << c << "]); }\n";
  }
  for (int u = 0; u < nun; ++u) o << "  cw" << u << " *= det[" << u << "] * wJ;\n";
  for (int e = 0; e < S.nelec; ++e) {
    const bool up = e < S.nup;
    const int n = up ? S.nup : S.ndown, el = up ? e : e - S.nup;
    o << "  {  // electron " << e << "\n    double g[NM];\n#pragma unroll\n    for (int m = 0; m < NM; ++m) g[m] = 0.0;\n";
    const int u0 = up ? 0 : S.nuu, u1 = up ? S.nuu : nun;
    for (int u = u0; u < u1; ++u) {
      const int *cols = hi.data() + (up ? S.o_ucu + u * S.nup : S.o_ucd + (u - S.nuu) * S.ndown);
      for (int j = 0; j < n; ++j)
        o << "    g[" << cols[j] << "] = fma(cw" << u << ", inv" << u << "[" << j * n + el << "], g[" << cols[j] << "]);\n";
    }
    o << "#pragma unroll\n    for (int a = 0; a < NAO; ++a) {\n      const double v = sao[" << e << " * NAO + a];\n"
         "#pragma unroll\n      for (int m = 0; m < NM; ++m) dW[a][m] = fma(v, g[m], dW[a][m]);\n    }\n";
    // MODE_BWD_ALL: the AO row has been consumed - it becomes Gao[e][a] = sum_m G[e][m] W[a][m]
    o << "    if constexpr (MODE == MODE_BWD_ALL) {\n";
    for (int a = 0; a < S.nao; ++a) {
      o << "      const_cast<double *>(sao)[" << e << " * NAO + " << a << "] = ";
      for (int m = 0; m < S.nmu; ++m) o << (m ? " + " : "") << "g[" << m << "] * spec_pv<MODE, " << L.off_mow + a * S.nmu + m << ">()";
      o << ";\n";
    }
    o << "    }\n  }\n";
  }
  o << "  sig_out = sig;\n}\n";
  if (code) *code = o.str();
  return true;
}

std::string prelude(const qmcb_plan *p, const Layout &L, Kind kind = KIND_THREAD) {
  const DevSys &S = p->sys;
  std::ostringstream o;
  if (kind == KIND_THREAD) {
    o << "#define QMCB_ETAB_REP " << spec_etab_rep() << "\n#define SPEC_NAO " << S.nao << "\n#define SPEC_NCONF " << S.nconf << "\n#define SPEC_NBAS " << S.nbas << "\n";
    // two-electron systems: both trips of the electron loop in one basic block (measured, H2 1e6 walkers:
    // E_L 0.041 -> 0.039 ms, psi 0.023 -> 0.021 ms; for LiH the unrolled loop is slower - registers)
    if (S.nelec <= 2) o << "#ifndef SPEC_EUNROLL\n#define SPEC_EUNROLL 2\n#endif\n";
  }
  if (kind == KIND_TILE) {
    o << "#define SPEC_TILE 1\n#define SPEC_KPFX \"spect_\"\n#define SPEC_EEN_NTERM " << S.een_nterm
      << "\n#define SPEC_NAO " << S.nao << "\n#define SPEC_NCONF " << S.nconf << "\n#define SPEC_MOW_SMEM "
      << (tile_mow_smem(S) ? 1 : 0) << "\n#define SPEC_MWLD " << S.nmup << "\n#define SPEC_O_MOW " << S.o_mow
      << "\n#define SPEC_O_CI " << S.o_ci << "\n#define SPEC_O_UCU " << S.o_ucu << "\n#define SPEC_O_UCD " << S.o_ucd
      << "\n#define SPEC_O_CIU " << S.o_ciu << "\n#define SPEC_O_CID " << S.o_cid << "\n";
  }
  o << "typedef signed char int8_t;\ntypedef unsigned char uint8_t;\ntypedef int int32_t;\n"
       "typedef unsigned int uint32_t;\ntypedef long long int64_t;\ntypedef unsigned long long uint64_t;\n"
       "typedef unsigned long size_t;\n"
    << "#define QMCB_GTO_PURE 0\n#define QMCB_GTO 1\n#define QMCB_STO_PURE 2\n#define QMCB_STO 3\n"
    << "#define SPEC_NE " << S.nelec << "\n#define SPEC_NUP " << S.nup << "\n#define SPEC_NDOWN " << S.ndown
    << "\n#define SPEC_NATOM " << S.natom << "\n#define SPEC_NMUP " << S.nmu << "\n#define SPEC_NUU " << S.nuu
    << "\n#define SPEC_NUD " << S.nud << "\n#define SPEC_USE_JEE " << (S.use_jee ? 1 : 0) << "\n#define SPEC_USE_JEN "
    << (S.use_jen ? 1 : 0) << "\n#define SPEC_GRAM_FMA " << (S.gram_fma ? 1 : 0) << "\n#define SPEC_NV " << L.nv
    << "\n#define SPEC_OFF_ATOM " << L.off_atom << "\n#define SPEC_OFF_MOW " << L.off_mow << "\n#define SPEC_OFF_CI "
    << L.off_ci << "\n#define SPEC_THREADS " << (kind == KIND_TILE ? tile_threads(S) : SPEC_THREADS) << "\n#define SPEC_MINB "
    << (kind == KIND_TILE ? tile_minb(S) : SPEC_MINB) << "\n";
  return o.str();
}

std::string full_source(const qmcb_plan *p, const Layout &L, const std::string &code, Kind kind) {
  std::string k = kind == KIND_TILE ? SPEC_SRC_TILE : SPEC_SRC_KERNEL;
  const std::string mark = "SPEC_GENERATED_CODE\n";
  const size_t at = k.find("\n" + mark);
  if (at == std::string::npos) return std::string();
  k.replace(at + 1, mark.size(), code);
  return prelude(p, L, kind) + SPEC_SRC_ARGS + SPEC_SRC_DEVICE + SPEC_SRC_PHILOX + SPEC_SRC_COMMON + k;
}

// ---------------------------------------------------------------------------------------
// compile + module cache
// ---------------------------------------------------------------------------------------
struct Module {
  std::vector<char> cubin;
  std::string log;
  CUmodule mod = nullptr;
  CUfunction fn[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // psi, eloc, mh, grad, backward, backward (all)
  int occ[6] = {0, 0, 0, 0, 0, 0};
  int smem[6] = {0, 0, 0, 0, 0, 0};
  bool loaded = false;
  bool from_disk = false;
};

// ---- on-disk cache of compiled modules: <dir of libqmcb.so>/jitcache/<hash of source + options>.cubin
// (QMCB_JIT_CACHE=<dir> overrides, QMCB_JIT_CACHE=0 disables).  The straight-line program of a large
// molecule takes ptxas tens of seconds; the cache makes that a one-off per structure and lets modules
// built ahead of time (python -m qmctorch_b200.build --prebuild) travel with the library.
std::string cache_dir() {
  const char *e = getenv("QMCB_JIT_CACHE");
  if (e && !strcmp(e, "0")) return std::string();
  if (e && *e) return e;
  Dl_info info;
  if (dladdr((void *)&cache_dir, &info) && info.dli_fname) {
    std::string path = info.dli_fname;
    const size_t at = path.rfind('/');
    if (at != std::string::npos) return path.substr(0, at) + "/jitcache";
  }
  return std::string();
}
std::string hash_hex(const std::string &text) {
  uint64_t h1 = 1469598103934665603ull, h2 = 0x9E3779B97F4A7C15ull;
  for (unsigned char c : text) {
    h1 = (h1 ^ c) * 1099511628211ull;
    h2 = (h2 + c) * 0xD6E8FEB86659FD93ull; h2 ^= h2 >> 32;
  }
  char buf[40];
  snprintf(buf, sizeof(buf), "%016llx%016llx", (unsigned long long)h1, (unsigned long long)h2);
  return buf;
}
bool disk_load(const std::string &key, std::vector<char> &cubin) {
  const std::string dir = cache_dir();
  if (dir.empty()) return false;
  FILE *f = fopen((dir + "/" + key + ".cubin").c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  bool ok = n > 0;
  if (ok) { cubin.resize((size_t)n); ok = fread(cubin.data(), 1, (size_t)n, f) == (size_t)n; }
  fclose(f);
  if (!ok) cubin.clear();
  return ok;
}
void disk_store(const std::string &key, const std::vector<char> &cubin) {
  const std::string dir = cache_dir();
  if (dir.empty()) return;
  mkdir(dir.c_str(), 0755);
  const std::string tmp = dir + "/" + key + ".tmp" + std::to_string((long)getpid());
  FILE *f = fopen(tmp.c_str(), "wb");
  if (!f) return;
  const bool ok = fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
  fclose(f);
  if (ok) rename(tmp.c_str(), (dir + "/" + key + ".cubin").c_str()); else remove(tmp.c_str());
}
std::map<std::string, Module> &cache() { static std::map<std::string, Module> c; return c; }
std::mutex &cache_mu() { static std::mutex m; return m; }

int compile(const std::string &src, const std::string &arch, Module &m, std::string &err) {
  Dyn &d = dyn();
  if (!d.ok) { err = d.why; return -1; }
  nvrtcProgram prog = nullptr;
  if (d.CreateProgram(&prog, src.c_str(), "qmcb_spec.cu", 0, nullptr, nullptr) != 0) { err = "nvrtcCreateProgram failed"; return -1; }
  const std::string a = "--gpu-architecture=" + arch;
  std::vector<std::string> extra = extra_defs();
  std::vector<const char *> opts = {a.c_str(), "--std=c++17", "-lineinfo", "-DQMCB_SPEC"};
  for (auto &e : extra) opts.push_back(e.c_str());
  const int rc = d.CompileProgram(prog, (int)opts.size(), opts.data());
  size_t n = 0;
  d.GetProgramLogSize(prog, &n);
  if (n > 1) { m.log.resize(n); d.GetProgramLog(prog, &m.log[0]); }
  if (rc != 0) {
    err = "nvrtcCompileProgram failed: " + m.log.substr(0, 2000);
    d.DestroyProgram(&prog);
    return -1;
  }
  size_t cs = 0;
  if (d.GetCUBINSize(prog, &cs) != 0 || cs == 0) { err = "nvrtcGetCUBIN: no cubin"; d.DestroyProgram(&prog); return -1; }
  m.cubin.resize(cs);
  d.GetCUBIN(prog, m.cubin.data());
  if (const char *dump = getenv("QMCB_JIT_DUMP")) {
    size_t ps = 0;
    if (d.GetPTXSize && d.GetPTX && d.GetPTXSize(prog, &ps) == 0 && ps > 1) {
      std::vector<char> ptx(ps);
      d.GetPTX(prog, ptx.data());
      FILE *f = fopen((std::string(dump) + ".ptx").c_str(), "w");
      if (f) { fwrite(ptx.data(), 1, ps - 1, f); fclose(f); }
    }
  }
  d.DestroyProgram(&prog);
  return 0;
}

std::string drv_err(int rc) {
  const char *s = nullptr;
  Dyn &d = dyn();
  if (d.GetErrorString) d.GetErrorString(rc, &s);
  return s ? s : ("CUresult " + std::to_string(rc));
}

// dynamic shared memory of one CTA (doubles): etab | optional MO weights + CI | per-thread slices
// register accumulators of the specialised backward: dW[nao][nmu] + CI + 2 Jastrow weights
bool bwd_eligible(const DevSys &S) { return S.nao * S.nmu <= 48 && S.nconf <= 16; }

size_t smem_doubles(const DevSys &S, int mode) {
  if (mode == MODE_BWD || mode == MODE_BWD_ALL) {
    const int slice = (6 * S.nelec + S.nelec * S.nmu + S.nelec * S.nao) | 1;
    const size_t red = (size_t)(SPEC_THREADS / 32) * (S.nao * S.nmu + S.nconf + 2 + (mode == MODE_BWD_ALL ? 2 * S.nbas : 0));
    const size_t body = (size_t)SPEC_THREADS * slice;
    return (size_t)QMCB_ETAB * spec_etab_rep() + (body > red ? body : red);
  }
  const bool deriv = mode == MODE_ELOC || mode == MODE_GRAD;
  const int nrow = mode == MODE_ELOC ? 2 : (mode == MODE_GRAD ? 4 : 1);
  const int ne3 = 3 * S.nelec;
  int slice = (ne3 + (deriv ? 4 * S.nelec : 0) + nrow * S.nelec * S.nmu) | 1;
  if (def_value("SPEC_PREFETCH", SPEC_DEFAULT_PREFETCH) ||
      (mode == MODE_ELOC && def_value("SPEC_PREFETCH_ELOC", SPEC_DEFAULT_PREFETCH_ELOC)))
    slice += ne3 + (ne3 & 1);
  const int nmw = def_value("SPEC_MOW_SMEM", SPEC_DEFAULT_MOW_SMEM) ? ((S.nao * S.nmu + S.nconf + 1) & ~1) : 0;
  return (size_t)QMCB_ETAB * spec_etab_rep() + (size_t)nmw + (size_t)SPEC_THREADS * slice;
}

// warp-tile kernels (spec_tile.cuh): exp table | MO weights | CI | int tables | per-warp work areas
size_t tile_smem_doubles(const DevSys &S, int mode) {
  const int ne = S.nelec, per = 32 / ne, nm = S.nmu, ldm = nm | 1, nun = S.nuu + S.nud;
  const int nrow = mode == MODE_ELOC ? 2 : 1;
  const size_t nmw = tile_mow_smem(S) ? (size_t)S.nao * S.nmup : 0;
  const size_t nci = (S.nconf + 1) & ~1;
  const size_t nit = ((size_t)S.nuu * S.nup + (size_t)S.nud * S.ndown + 2 * (size_t)S.nconf + 1) & ~(size_t)1;
  const size_t npos = ((size_t)per * 3 * ne + 1) & ~(size_t)1;
  // psi / E_L double-buffer the coordinates (cp.async prefetch of the next tile)
  const size_t o_jv = npos * ((mode != MODE_MH && def_value("SPEC_TILE_PREFETCH", 1)) ? 2 : 1);
  const size_t rows = (size_t)nrow * per * ne * ldm;
  const size_t tls = S.een_nterm > 0 ? (size_t)((S.natom * (1 + 3 * S.een_nterm)) | 1) : 0;   // een_table_doubles
  const size_t ws = (o_jv + 96 + (rows > 32 * tls ? rows : 32 * tls) + 2 * (size_t)per * nun + 1) & ~(size_t)1;
  return QMCB_ETAB + nmw + nci + nit / 2 + (size_t)(tile_threads(S) / 32) * ws;
}

}  // namespace

struct qmcb_spec_state {
  int level = -1;             // QMCB_JIT as read when this plan was first prepared
  bool failed = false;
  std::string why;
  Module *mod = nullptr;
  Layout lay;
  Kind kind = KIND_NONE;
  std::vector<char> params;   // host image of SpecParams
  uint64_t params_version = ~0ull;
};

void qmcb_spec_free(qmcb_plan *p) {
  delete p->spec;
  p->spec = nullptr;
}

// "sm_100a" for a host-only plan, else the device's own architecture (queried once per device:
// plans are re-prepared after every parameter update)
static std::string device_arch(int device) {
  if (device < 0) return "sm_100a";
  static std::mutex mu;
  static std::map<int, std::string> known;
  std::lock_guard<std::mutex> lk(mu);
  auto it = known.find(device);
  if (it != known.end()) return it->second;
  int major = 10, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  std::string arch = "sm_" + std::to_string(major) + std::to_string(minor);
  if (major >= 9) arch += "a";
  known[device] = arch;
  return arch;
}

// Builds (or fetches) the specialised module of this plan.  load = false: compile only (no driver).
// Returns 0 when the module is ready, 1 when the generic kernels must be used (st.why says why).
static int spec_prepare(const qmcb_plan *p, bool load) {
  if (!p->spec) p->spec = new qmcb_spec_state();
  qmcb_spec_state &st = *p->spec;
  if (st.level < 0) st.level = jit_level();
  if (st.failed) return 1;
  if (st.mod && (!load || st.mod->loaded)) return 0;
  auto fail = [&](const std::string &why) { st.failed = true; st.why = why; return 1; };
  if (st.level <= 0) return fail("disabled (QMCB_JIT=0)");
  std::string why;
  const Kind kind = choose_kind(p, &why);
  if (kind == KIND_NONE) return fail("not eligible: " + why);
  st.kind = kind;
  std::string code;
  Layout L;
  if (!walk(p, &code, nullptr, L, kind)) return fail("program walk failed");
  if (L.nv > (kind == KIND_TILE ? TILE_MAX_VALUES : SPEC_MAX_VALUES)) return fail("parameter block too large");
  st.lay = L;
  const std::string arch = device_arch(p->device);
  std::string key = std::to_string(p->device) + "|" + arch + "|" + prelude(p, L, kind) + code;
  for (auto &e : extra_defs()) key += "|" + e;
  std::lock_guard<std::mutex> lk(cache_mu());
  Module &m = cache()[key];
  if (m.cubin.empty()) {
    const std::string src = full_source(p, L, code, kind);
    if (src.empty()) { cache().erase(key); return fail("embedded source has no insertion mark"); }
    std::string opts = arch;
    for (auto &e : extra_defs()) opts += "|" + e;
    const std::string dkey = hash_hex(opts + "\n" + src);
    if (!getenv("QMCB_JIT_DUMP") && disk_load(dkey, m.cubin)) m.from_disk = true;
    const char *dump = getenv("QMCB_JIT_DUMP");
    if (dump) {
      FILE *f = fopen((std::string(dump) + ".cu").c_str(), "w");
      if (f) { fwrite(src.data(), 1, src.size(), f); fclose(f); }
    }
    std::string err;
    if (m.cubin.empty()) {
      if (compile(src, arch, m, err) != 0) { cache().erase(key); return fail(err); }
      disk_store(dkey, m.cubin);
    }
    if (dump) {
      FILE *f = fopen((std::string(dump) + ".cubin").c_str(), "wb");
      if (f) { fwrite(m.cubin.data(), 1, m.cubin.size(), f); fclose(f); }
    }
  }
  if (load && !m.loaded) {
    Dyn &d = dyn();
    if (!d.drv) return fail("libcuda.so.1 not available");
    DeviceGuard guard(p->device);
    cudaFree(nullptr);   // primary context current on this thread
    int rc = d.ModuleLoadData(&m.mod, m.cubin.data());
    if (rc != 0) return fail("cuModuleLoadData: " + drv_err(rc));
    const bool tile = kind == KIND_TILE;
    const char *names[6] = {tile ? "spect_psi" : "spec_psi", tile ? "spect_eloc" : "spec_eloc",
                            tile ? "spect_mh" : "spec_mh", tile ? nullptr : "spec_grad_psi",
                            (tile || !bwd_eligible(p->sys)) ? nullptr : "spec_backward",
                            (tile || !bwd_eligible(p->sys) || 2 * p->sys.nbas > 64 || p->multi_component) ? nullptr
                                                                                                           : "spec_backward_all"};
    const int modes[6] = {MODE_PSI, MODE_ELOC, MODE_MH, MODE_GRAD, MODE_BWD, MODE_BWD_ALL};
    const int threads = tile ? tile_threads(p->sys) : SPEC_THREADS;
    for (int i = 0; i < 6; ++i) {
      if (!names[i]) { m.fn[i] = nullptr; continue; }     // grad psi / backward of tile structures: generic kernels
      rc = d.ModuleGetFunction(&m.fn[i], m.mod, names[i]);
      if (rc != 0) return fail(std::string("cuModuleGetFunction ") + names[i] + ": " + drv_err(rc));
      m.smem[i] = (int)((tile ? tile_smem_doubles(p->sys, modes[i]) : smem_doubles(p->sys, modes[i])) * sizeof(double));
      if ((size_t)m.smem[i] > (tile ? TILE_SMEM_BUDGET : SPEC_SMEM_BUDGET)) { m.fn[i] = nullptr; continue; }   // this mode stays generic
      rc = d.FuncSetAttribute(m.fn[i], 8 /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES */, m.smem[i]);
      if (rc != 0) return fail("cuFuncSetAttribute: " + drv_err(rc));
      int occ = 0;
      rc = d.OccupancyMaxActiveBlocks(&occ, m.fn[i], threads, (size_t)m.smem[i]);
      m.occ[i] = (rc == 0 && occ > 0) ? occ : 1;
    }
    m.loaded = true;
  }
  st.mod = &m;
  return 0;
}

static void refresh_params(const qmcb_plan *p, qmcb_spec_state &st) {
  if (st.params_version == p->version && !st.params.empty()) return;
  const DevSys &S = p->sys;
  std::vector<double> vals;
  Layout L;
  walk(p, nullptr, &vals, L, st.kind);
  // SpecParams (spec_common.cuh): expc[8] | jee_w jen_w vnn | etab, dtab, itab pointers | een a b a2 b2 c [8] | v[NV]
  constexpr int HDR = 8 + 3 + 3 + 5 * 8;      // = SPEC_V_BYTE0 / 8
  static_assert(HDR * 8 == 432, "SpecParams header layout (SPEC_V_BYTE0)");
  static_assert(QMCB_EEN_MAXTERM == 8, "SPEC_EEN_MAX");
  std::vector<char> buf((HDR + (size_t)L.nv) * sizeof(double));
  double *d = reinterpret_cast<double *>(buf.data());
  for (int i = 0; i < 8; ++i) d[i] = S.expc[i];
  d[8] = S.jee_w; d[9] = S.jen_w; d[10] = S.vnn;
  const double *et = p->d_dbl ? p->d_dbl + S.o_etab : nullptr;
  memcpy(&d[11], &et, sizeof(et));
  memcpy(&d[12], &p->d_dbl, sizeof(p->d_dbl));
  memcpy(&d[13], &p->d_int, sizeof(p->d_int));
  for (int m = 0; m < QMCB_EEN_MAXTERM; ++m) {
    d[14 + m] = S.een_a[m]; d[22 + m] = S.een_b[m]; d[30 + m] = S.een_a2[m]; d[38 + m] = S.een_b2[m];
    d[46 + m] = S.een_c[m];
  }
  for (int i = 0; i < L.nv; ++i) d[HDR + i] = vals[i];
  st.params.swap(buf);
  st.params_version = p->version;
}

int qmcb_spec_launch(const qmcb_plan *p, int mode, const FusedArgs &a, void *stream, int *grid_out) {
  const int slot = mode == MODE_PSI ? 0 : (mode == MODE_ELOC ? 1 : (mode == MODE_MH ? 2 : (mode == MODE_GRAD ? 3 :
                   (mode == MODE_BWD ? 4 : (mode == MODE_BWD_ALL ? 5 : -1)))));
  if (slot < 0 || p->device < 0) return QMCB_SPEC_SKIP;
  if (spec_prepare(p, true) != 0) {
    if (p->spec->level >= 2) {
      qmcb_set_error("qmcb: structure-specialised kernel unavailable: " + p->spec->why);
      return QMCB_EINVAL;
    }
    return QMCB_SPEC_SKIP;
  }
  qmcb_spec_state &st = *p->spec;
  Module &m = *st.mod;
  if (!m.fn[slot]) return QMCB_SPEC_SKIP;
  refresh_params(p, st);
  const bool tile = st.kind == KIND_TILE;
  const int threads = tile ? tile_threads(p->sys) : SPEC_THREADS;
  int64_t grid = (int64_t)p->sm_count * m.occ[slot];
  // work units per CTA: walkers (one per thread) or warp tiles of 32 / Ne walkers
  const int64_t per_cta = tile ? (int64_t)(threads / 32) * (32 / p->sys.nelec) : threads;
  const int64_t need = (a.W + per_cta - 1) / per_cta;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  if (grid_out) *grid_out = (int)grid;
  FusedArgs args = a;
  void *kp[2] = {st.params.data(), &args};
  const int rc = dyn().LaunchKernel(m.fn[slot], (unsigned)grid, 1, 1, threads, 1, 1, (unsigned)m.smem[slot],
                                    (CUstream)stream, kp, nullptr);
  if (rc != 0) {
    qmcb_set_error("qmcb: cuLaunchKernel(spec): " + drv_err(rc));
    return QMCB_EINVAL;
  }
  return 0;
}

// 1: specialised kernels compiled (and loaded when the plan has a device), 0: generic kernels
int qmcb_spec_status(const qmcb_plan *p, std::string *why) {
  const int rc = spec_prepare(p, p->device >= 0);
  if (why) *why = p->spec ? p->spec->why : std::string();
  return rc == 0 ? 1 : 0;
}

int qmcb_spec_eligible(const qmcb_plan *p) { return jit_level() > 0 ? (int)choose_kind(p, nullptr) : 0; }

// 0: generic kernels, 1: one walker per thread, 2: warp tiles (after qmcb_spec_status)
int qmcb_spec_kind(const qmcb_plan *p) { return p->spec && !p->spec->failed ? (int)p->spec->kind : 0; }
