// Host side of the plan: regroup the reference's flat primitive list into shells
// (one exp per (atom, exponent) shared by all cartesian components), collect the
// MO columns that some configuration occupies, deduplicate spin occupations, and
// choose the CTA tiling.  Reference data contract: atomic_orbitals.py:27-94,
// molecular_orbitals.py:40-74, orbital_configurations.py:14-214, orbital_projector.py:29-51.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>

#include "plan.h"
#include "spec.h"

static thread_local std::string g_err;
void qmcb_set_error(const std::string &msg) { g_err = msg; }
int qmcb_cuda_rc(int cuda_error, const char *where) {
  if (cuda_error != 0)
    g_err = std::string(where) + ": " + cudaGetErrorString((cudaError_t)cuda_error);
  return cuda_error;
}
extern "C" const char *qmcb_last_error(void) { return g_err.c_str(); }

unsigned *qmcb_ticket_slot(const qmcb_plan *p, void *stream) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (!p->d_ticket) return nullptr;
  int free_slot = -1;
  for (int i = 0; i < qmcb_plan::kTicketSlots; ++i) {
    if (p->ticket_used[i] && p->ticket_owner[i] == stream) return p->d_ticket + 2 * i;
    if (!p->ticket_used[i] && free_slot < 0) free_slot = i;
  }
  if (free_slot < 0) return nullptr;
  p->ticket_used[free_slot] = true;
  p->ticket_owner[free_slot] = stream;
  return p->d_ticket + 2 * free_slot;
}
extern "C" int qmcb_abi_version(void) { return QMCB_ABI_VERSION; }

namespace {

struct AoDesc {
  int ao = 0;           // AO index this component emits into
  int atom, kx, ky, kz;
  std::vector<double> alpha, pn, cn;   // per primitive
  std::vector<int> flat;
};

struct ShellDesc {
  int atom;
  std::vector<double> alpha, pn, coef;
  std::vector<int> comp_ao, comp_k;
  std::vector<double> comp_scale;
  std::vector<std::vector<int>> flat;  // [prim][comp]
};

bool proportional(const std::vector<double> &a, const std::vector<double> &ref, double *ratio) {
  double r = 0.0;
  bool have = false;
  for (size_t i = 0; i < a.size(); ++i) {
    if (ref[i] == 0.0) {
      if (a[i] != 0.0) return false;
      continue;
    }
    double q = a[i] / ref[i];
    if (!have) { r = q; have = true; }
    else if (std::fabs(q - r) > 1e-13 * std::fabs(r)) return false;
  }
  if (!have) r = 1.0;
  *ratio = r;
  return true;
}

}  // namespace

int qmcb_build_tables(const qmcb_system *s, qmcb_plan *p) {
  if (!s || s->nelec <= 0 || s->nup + s->ndown != s->nelec || s->natom <= 0 || s->nbas <= 0 ||
      s->nao <= 0 || s->nmo <= 0 || s->nconf <= 0) {
    qmcb_set_error("qmcb_plan: inconsistent sizes");
    return QMCB_EINVAL;
  }
  if (s->radial_type < 0 || s->radial_type > 3) {
    qmcb_set_error("qmcb_plan: unknown radial_type");
    return QMCB_EINVAL;
  }
  // ---- AOs from flat primitives.  An AO is normally ONE cartesian monomial x^kx y^ky z^kz times a
  // contracted radial function; real spherical harmonics of l = 2 (x^2 - y^2, 2z^2 - x^2 - y^2) arrive as
  // several monomials with the same AO index (the host expands them, wavefunction/orbitals/atomic_orbitals.py).
  // Every (AO, monomial) pair becomes a "component" of its own that EMITS INTO THE SAME AO: the projection
  // onto the MOs is linear, so the fused kernels need nothing else.
  std::vector<AoDesc> aos;
  std::map<std::vector<int>, int> ao_key;           // (AO, kx, ky, kz) -> component descriptor
  std::vector<char> seen(s->nao, 0);
  bool multi = false;
  for (int i = 0; i < s->nbas; ++i) {
    int a = s->index_ctr[i];
    if (a < 0 || a >= s->nao) {
      qmcb_set_error("qmcb_plan: index_ctr out of range");
      return QMCB_EINVAL;
    }
    const std::vector<int> key = {a, s->bas_kx[i], s->bas_ky[i], s->bas_kz[i]};
    auto it = ao_key.find(key);
    if (it == ao_key.end()) {
      if (seen[a]) multi = true;
      seen[a] = 1;
      it = ao_key.emplace(key, (int)aos.size()).first;
      aos.emplace_back();
      AoDesc &n = aos.back();
      n.ao = a;
      n.atom = s->bas_atom[i];
      n.kx = s->bas_kx[i]; n.ky = s->bas_ky[i]; n.kz = s->bas_kz[i];
    }
    AoDesc &d = aos[it->second];
    if (d.atom != s->bas_atom[i]) {
      qmcb_set_error("qmcb_plan: an AO mixes primitives of different centres (unsupported)");
      return QMCB_EINVAL;
    }
    if (d.kx < 0 || d.ky < 0 || d.kz < 0 || d.kx > 15 || d.ky > 15 || d.kz > 15) {
      qmcb_set_error("qmcb_plan: cartesian power out of range");
      return QMCB_EINVAL;
    }
    double cg = s->bas_norm[i] * s->bas_coeffs[i];
    double cv = s->contract ? cg : s->bas_norm[i];
    if (cv != cg) {
      // atomic_orbitals.py:236-249 drops bas_coeffs from values/laplacians when nothing is
      // contracted while :344 keeps them in gradients; identical unless coeffs != 1.
      qmcb_set_error("qmcb_plan: uncontracted basis with coefficients != 1 is not supported");
      return QMCB_EINVAL;
    }
    d.alpha.push_back(s->bas_exp[i]);
    d.pn.push_back((double)s->bas_kr[i]);
    d.cn.push_back(cg);
    d.flat.push_back(i);
  }
  for (int a = 0; a < s->nao; ++a)
    if (!seen[a]) {
      qmcb_set_error("qmcb_plan: AO without primitives");
      return QMCB_EINVAL;
    }
  p->multi_component = multi;
  // components in AO order (stable: the monomials of one AO stay together)
  std::stable_sort(aos.begin(), aos.end(), [](const AoDesc &x, const AoDesc &y) { return x.ao < y.ao; });
  // ---- shells
  std::vector<ShellDesc> shells;
  for (size_t ic = 0; ic < aos.size(); ++ic) {
    const AoDesc &d = aos[ic];
    const int a = d.ao;
    bool placed = false;
    for (auto &sh : shells) {
      if (sh.atom != d.atom || sh.alpha != d.alpha || sh.pn != d.pn) continue;
      double ratio;
      if (!proportional(d.cn, sh.coef, &ratio)) continue;
      sh.comp_ao.push_back(a);
      sh.comp_k.push_back(d.kx | (d.ky << 8) | (d.kz << 16));
      sh.comp_scale.push_back(ratio);
      for (size_t q = 0; q < d.flat.size(); ++q) sh.flat[q].push_back(d.flat[q]);
      placed = true;
      break;
    }
    if (!placed) {
      ShellDesc sh;
      sh.atom = d.atom;
      sh.alpha = d.alpha; sh.pn = d.pn; sh.coef = d.cn;
      sh.comp_ao = {a};
      sh.comp_k = {d.kx | (d.ky << 8) | (d.kz << 16)};
      sh.comp_scale = {1.0};
      sh.flat.resize(d.flat.size());
      for (size_t q = 0; q < d.flat.size(); ++q) sh.flat[q] = {d.flat[q]};
      shells.push_back(sh);
    }
  }
  std::stable_sort(shells.begin(), shells.end(),
                   [](const ShellDesc &x, const ShellDesc &y) { return x.atom < y.atom; });
  // ---- used MO columns and unique spin occupations
  std::vector<int> used;
  for (int c = 0; c < s->nconf; ++c) {
    for (int j = 0; j < s->nup; ++j) used.push_back(s->cfg_up[c * s->nup + j]);
    for (int j = 0; j < s->ndown; ++j) used.push_back(s->cfg_down[c * s->ndown + j]);
  }
  std::sort(used.begin(), used.end());
  used.erase(std::unique(used.begin(), used.end()), used.end());
  for (int m : used)
    if (m < 0 || m >= s->nmo) {
      qmcb_set_error("qmcb_plan: configuration refers to an MO outside [0,nmo)");
      return QMCB_EINVAL;
    }
  std::map<int, int> upos;
  for (size_t i = 0; i < used.size(); ++i) upos[used[i]] = (int)i;
  std::vector<std::vector<int>> uu, ud;
  std::vector<int> ciu(s->nconf), cid(s->nconf);
  for (int c = 0; c < s->nconf; ++c) {
    std::vector<int> u(s->nup), d(s->ndown);
    for (int j = 0; j < s->nup; ++j) u[j] = upos[s->cfg_up[c * s->nup + j]];
    for (int j = 0; j < s->ndown; ++j) d[j] = upos[s->cfg_down[c * s->ndown + j]];
    auto iu = std::find(uu.begin(), uu.end(), u);
    if (iu == uu.end()) { uu.push_back(u); ciu[c] = (int)uu.size() - 1; }
    else ciu[c] = (int)(iu - uu.begin());
    auto id = std::find(ud.begin(), ud.end(), d);
    if (id == ud.end()) { ud.push_back(d); cid[c] = (int)ud.size() - 1; }
    else cid[c] = (int)(id - ud.begin());
  }
  // ---- sizes
  DevSys &S = p->sys;
  S = DevSys{};
  S.nelec = s->nelec; S.nup = s->nup; S.ndown = s->ndown; S.natom = s->natom;
  S.nshell = (int)shells.size();
  S.nao = s->nao; S.nmo = s->nmo; S.nbas = s->nbas;
  S.nmu = (int)used.size();
  int mb = 1;
  while (mb < S.nmu && mb < 8) mb *= 2;
  S.nmup = ((S.nmu + mb - 1) / mb) * mb;
  S.nconf = s->nconf; S.nuu = (int)uu.size(); S.nud = (int)ud.size();
  S.radial_type = s->radial_type; S.use_jee = s->use_jee; S.use_jen = s->use_jen;
  S.gram_fma = s->gram_fma;
  S.jee_w = s->jee_w; S.jen_w = s->jen_w;
  S.een_nterm = s->een_nterm;
  if (s->een_nterm < 0 || s->een_nterm > QMCB_EEN_MAXTERM) {
    qmcb_set_error("qmcb_plan: een_nterm out of range");
    return QMCB_EINVAL;
  }
  for (int m = 0; m < s->een_nterm; ++m) {
    S.een_a[m] = s->een_num[m]; S.een_a2[m] = s->een_num[s->een_nterm + m];
    S.een_b[m] = s->een_denom[m]; S.een_b2[m] = s->een_denom[s->een_nterm + m];
    S.een_c[m] = s->een_fc[m];
  }
  {
    // exp(x) = 2^m * 2^(j/128) * exp(r), x = (128 m + j) ln2/128 + r, |r| <= ln2/256:
    // 128/ln2, the 1.5*2^52 rounding constant, -ln2/128 (one constant: its rounding error 4e-19
    // times |128 m + j| stays below 1e-15 for |x| < 14, i.e. for every term above 1e-6 of a sum),
    // Taylor 1/24, 1/6 (degree-4 truncation r^5/120 <= 1.2e-15 on that interval)
    // (QMCB_ETAB_LOG2 >= 10: |r| <= ln2/2048, the degree-3 polynomial is enough: r^4/24 <= 6e-16)
    const double c[16] = {std::ldexp(0x1.71547652b82fep+0, QMCB_ETAB_LOG2), 6755399441055744.0,
                          -std::ldexp(0x1.62e42fefa39efp-1, -QMCB_ETAB_LOG2),
                          0.0, 1.0 / 24.0, 1.0 / 6.0, 0.0, 0.5,
                          0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 16; ++i) S.expc[i] = c[i];
  }
  // nuclear repulsion, wf_base.py:97-116
  double vnn = 0.0;
  for (int a = 0; a < s->natom - 1; ++a)
    for (int b = a + 1; b < s->natom; ++b) {
      double dx = s->atom_coords[3 * a] - s->atom_coords[3 * b];
      double dy = s->atom_coords[3 * a + 1] - s->atom_coords[3 * b + 1];
      double dz = s->atom_coords[3 * a + 2] - s->atom_coords[3 * b + 2];
      vnn += s->atomic_number[a] * s->atomic_number[b] / std::sqrt(dx * dx + dy * dy + dz * dz);
    }
  S.vnn = vnn;
  // ---- blobs
  std::vector<double> &hd = p->hd;
  std::vector<int> &hi = p->hi;
  hd.clear(); hi.clear();
  S.o_atoms = (int)hd.size();
  for (int a = 0; a < s->natom; ++a) {
    hd.push_back(s->atom_coords[3 * a]); hd.push_back(s->atom_coords[3 * a + 1]);
    hd.push_back(s->atom_coords[3 * a + 2]); hd.push_back(s->atomic_number[a]);
  }
  S.o_alpha = (int)hd.size();
  for (auto &sh : shells) hd.insert(hd.end(), sh.alpha.begin(), sh.alpha.end());
  S.nprim = (int)hd.size() - S.o_alpha;
  S.o_coef = (int)hd.size();
  for (auto &sh : shells) hd.insert(hd.end(), sh.coef.begin(), sh.coef.end());
  S.o_pn = (int)hd.size();
  for (auto &sh : shells) hd.insert(hd.end(), sh.pn.begin(), sh.pn.end());
  S.o_cscale = (int)hd.size();
  for (auto &sh : shells) hd.insert(hd.end(), sh.comp_scale.begin(), sh.comp_scale.end());
  S.ncomp = (int)hd.size() - S.o_cscale;
  if (hd.size() & 1) hd.push_back(0.0);   // 16-byte aligned rows for vector loads
  S.o_mow = (int)hd.size();
  for (int a = 0; a < s->nao; ++a)
    for (int j = 0; j < S.nmup; ++j)
      hd.push_back(j < S.nmu ? s->mo[(size_t)a * s->nmo + used[j]] : 0.0);
  S.o_ci = (int)hd.size();
  for (int c = 0; c < s->nconf; ++c) hd.push_back(s->ci[c]);
  S.o_fnorm = (int)hd.size();
  for (int i = 0; i < s->nbas; ++i) hd.push_back(s->bas_norm[i]);
  if (hd.size() & 1) hd.push_back(0.0);
  S.o_etab = (int)hd.size();
  for (int j = 0; j < QMCB_ETAB; ++j) hd.push_back(std::exp2((double)j / (double)QMCB_ETAB));
  // ---- packed shell program: per shell a header record {nprim, ngroup | -}, then nprim records
  // {alpha, coef (, n as third field in the next record for gto/sto)}, then ngroup records
  // {kk | type<<24, ao | scale}.  A "P" group (type 1) stands for three consecutive AOs x,y,z
  // with a common scale; type 0 is a single component with explicit powers.
  if (hd.size() & 1) hd.push_back(0.0);
  S.o_stream = (int)hd.size();
  {
    auto push_ints = [&](int lo, int hi2, double y) {
      union { double d; int i[2]; } u;
      u.i[0] = lo; u.i[1] = hi2;
      hd.push_back(u.d); hd.push_back(y);
    };
    const bool with_n = (s->radial_type == QMCB_GTO || s->radial_type == QMCB_STO);
    for (auto &sh : shells) {
      // group components
      struct Grp { int kk, ao; double sc; };
      std::vector<Grp> grps;
      size_t k = 0;
      const size_t nc = sh.comp_ao.size();
      while (k < nc) {
        const int kx = 1, ky = 1 << 8, kz = 1 << 16;
        if (k + 2 < nc && sh.comp_k[k] == kx && sh.comp_k[k + 1] == ky && sh.comp_k[k + 2] == kz &&
            sh.comp_ao[k + 1] == sh.comp_ao[k] + 1 && sh.comp_ao[k + 2] == sh.comp_ao[k] + 2 &&
            sh.comp_scale[k + 1] == sh.comp_scale[k] && sh.comp_scale[k + 2] == sh.comp_scale[k]) {
          grps.push_back({(1 << 24), sh.comp_ao[k], sh.comp_scale[k]});
          k += 3;
        } else {
          grps.push_back({sh.comp_k[k], sh.comp_ao[k], sh.comp_scale[k]});
          k += 1;
        }
      }
      push_ints((int)sh.alpha.size(), (int)grps.size(), 0.0);
      for (size_t q = 0; q < sh.alpha.size(); ++q) {
        hd.push_back(sh.alpha[q]); hd.push_back(sh.coef[q]);
        if (with_n) { hd.push_back(sh.pn[q]); hd.push_back(0.0); }
      }
      for (auto &g : grps) push_ints(g.kk, g.ao, g.sc);
    }
  }
  S.nrec = ((int)hd.size() - S.o_stream) / 2;
  S.ndbl = (int)hd.size();

  S.o_ash = (int)hi.size();
  {
    int sidx = 0;
    for (int a = 0; a < s->natom; ++a) {
      hi.push_back(sidx);
      while (sidx < S.nshell && shells[sidx].atom == a) ++sidx;
    }
    hi.push_back(sidx);
    if (sidx != S.nshell) {
      qmcb_set_error("qmcb_plan: bas_atom out of range");
      return QMCB_EINVAL;
    }
  }
  S.o_spo = (int)hi.size();
  { int o = 0; for (auto &sh : shells) { hi.push_back(o); o += (int)sh.alpha.size(); } hi.push_back(o); }
  S.o_sco = (int)hi.size();
  { int o = 0; for (auto &sh : shells) { hi.push_back(o); o += (int)sh.comp_ao.size(); } hi.push_back(o); }
  S.o_ck = (int)hi.size();
  for (auto &sh : shells) hi.insert(hi.end(), sh.comp_k.begin(), sh.comp_k.end());
  S.o_cao = (int)hi.size();
  for (auto &sh : shells) hi.insert(hi.end(), sh.comp_ao.begin(), sh.comp_ao.end());
  S.o_used = (int)hi.size();
  hi.insert(hi.end(), used.begin(), used.end());
  S.o_ucu = (int)hi.size();
  for (auto &u : uu) hi.insert(hi.end(), u.begin(), u.end());
  S.o_ucd = (int)hi.size();
  for (auto &d : ud) hi.insert(hi.end(), d.begin(), d.end());
  S.o_ciu = (int)hi.size();
  hi.insert(hi.end(), ciu.begin(), ciu.end());
  S.o_cid = (int)hi.size();
  hi.insert(hi.end(), cid.begin(), cid.end());
  S.o_pfo = (int)hi.size();
  { int o = 0; for (auto &sh : shells) { hi.push_back(o); o += (int)(sh.alpha.size() * sh.comp_ao.size()); } hi.push_back(o); }
  S.o_pflat = (int)hi.size();
  for (auto &sh : shells)
    for (size_t q = 0; q < sh.alpha.size(); ++q)
      for (size_t k = 0; k < sh.comp_ao.size(); ++k) hi.push_back(sh.flat[q][k]);
  if (hi.size() & 1) hi.push_back(0);
  S.nint = (int)hi.size();

  // flat primitive tables of the E_L adjoint (plan.h)
  {
    const int nb = s->nbas;
    p->flat_dbl.assign(3 * (size_t)nb, 0.0);
    p->flat_int.assign(5 * (size_t)nb + 2 * (size_t)s->nao + 1, 0);
    for (int i = 0; i < nb; ++i) {
      p->flat_dbl[i] = s->bas_exp[i];
      p->flat_dbl[nb + i] = s->bas_norm[i] * s->bas_coeffs[i];
      p->flat_dbl[2 * nb + i] = s->bas_norm[i];
      p->flat_int[i] = s->bas_atom[i];
      p->flat_int[nb + i] = s->bas_kx[i] | (s->bas_ky[i] << 8) | (s->bas_kz[i] << 16);
      p->flat_int[2 * nb + i] = s->bas_kr[i];
      p->flat_int[3 * nb + i] = s->index_ctr[i];
    }
    int *start = p->flat_int.data() + 4 * nb, *list = start + s->nao + 1;
    int o = 0;
    for (int a = 0; a < s->nao; ++a) {
      start[a] = o;
      for (int i = 0; i < nb; ++i)
        if (s->index_ctr[i] == a) list[o++] = i;
    }
    start[s->nao] = o;
    // AOs by decreasing contraction length (stable): lockstep lanes of the adjoint kernel get equal work
    int *order = list + nb;
    for (int a = 0; a < s->nao; ++a) order[a] = a;
    std::stable_sort(order, order + s->nao, [&](int x, int y) { return start[x + 1] - start[x] > start[y + 1] - start[y]; });
  }
  p->index_ctr.assign(s->index_ctr, s->index_ctr + s->nbas);
  p->mo_full.assign(s->mo, s->mo + (size_t)s->nao * s->nmo);
  return 0;
}

static int upload(qmcb_plan *p) {
  cudaError_t e;
  size_t nd = p->hd.size() * sizeof(double), ni = p->hi.size() * sizeof(int);
  if (nd > p->cap_dbl) {
    if (p->d_dbl) cudaFree(p->d_dbl);
    if ((e = cudaMalloc(&p->d_dbl, nd)) != cudaSuccess) return (int)e;
    p->cap_dbl = nd;
  }
  if (ni > p->cap_int) {
    if (p->d_int) cudaFree(p->d_int);
    if ((e = cudaMalloc(&p->d_int, ni)) != cudaSuccess) return (int)e;
    p->cap_int = ni;
  }
  size_t nm = p->mo_full.size() * sizeof(double);
  if (nm > p->cap_mo_full) {
    if (p->d_mo_full) cudaFree(p->d_mo_full);
    if ((e = cudaMalloc(&p->d_mo_full, nm)) != cudaSuccess) return (int)e;
    p->cap_mo_full = nm;
  }
  // the default stream orders these copies before any later kernel on a blocking stream;
  // callers on non-blocking streams get a full sync here as well.
  if ((e = cudaMemcpy(p->d_dbl, p->hd.data(), nd, cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
  if ((e = cudaMemcpy(p->d_int, p->hi.data(), ni, cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
  if ((e = cudaMemcpy(p->d_mo_full, p->mo_full.data(), nm, cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
  size_t nfd = p->flat_dbl.size() * sizeof(double), nfi = p->flat_int.size() * sizeof(int);
  if (nfd > p->cap_flat_dbl) {
    if (p->d_flat_dbl) cudaFree(p->d_flat_dbl);
    p->d_flat_dbl = nullptr;
    if ((e = cudaMalloc(&p->d_flat_dbl, nfd)) != cudaSuccess) return (int)e;
    p->cap_flat_dbl = nfd;
  }
  if (nfi > p->cap_flat_int) {
    if (p->d_flat_int) cudaFree(p->d_flat_int);
    p->d_flat_int = nullptr;
    if ((e = cudaMalloc(&p->d_flat_int, nfi)) != cudaSuccess) return (int)e;
    p->cap_flat_int = nfi;
  }
  if ((e = cudaMemcpy(p->d_flat_dbl, p->flat_dbl.data(), nfd, cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
  if ((e = cudaMemcpy(p->d_flat_int, p->flat_int.data(), nfi, cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
  size_t nt = p->bwd_tiles.size() * sizeof(int);
  if (nt > p->cap_bwd_tiles) {
    if (p->d_bwd_tiles) cudaFree(p->d_bwd_tiles);
    p->d_bwd_tiles = nullptr;
    if ((e = cudaMalloc(&p->d_bwd_tiles, nt)) != cudaSuccess) return (int)e;
    p->cap_bwd_tiles = nt;
  }
  if (nt && (e = cudaMemcpy(p->d_bwd_tiles, p->bwd_tiles.data(), nt, cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
  if (!p->d_ticket) {
    const size_t nb = 2 * qmcb_plan::kTicketSlots * sizeof(unsigned);
    if ((e = cudaMalloc(&p->d_ticket, nb)) != cudaSuccess) return (int)e;
    if ((e = cudaMemset(p->d_ticket, 0, nb)) != cudaSuccess) return (int)e;
  }
  p->sys.dblob = p->d_dbl;
  p->sys.iblob = p->d_int;
  return 0;
}

extern "C" int qmcb_plan_create(const qmcb_system *sys, int device, qmcb_plan **out) {
  if (!out) return QMCB_EINVAL;
  *out = nullptr;
  // device < 0: host-only plan (tables + tiling, no upload) for CPU-side tests of the
  // grouping logic; every compute call on it fails with QMCB_EINVAL.
  if (device >= 0) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device >= ndev) {
      qmcb_set_error(std::string("qmcb_plan_create: no such CUDA device: ") +
                     (e != cudaSuccess ? cudaGetErrorString(e) : std::to_string(device).c_str()));
      return e != cudaSuccess ? (int)e : QMCB_EINVAL;
    }
  }
  DeviceGuard guard(device);
  qmcb_plan *p = new qmcb_plan();
  p->device = device;
  if (device >= 0) {
    cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&p->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  }
  int rc = qmcb_build_tables(sys, p);
  if (rc == 0) rc = qmcb_choose_launch(p);
  if (rc == 0) rc = qmcb_choose_backward(p);
  if (rc == 0 && device >= 0) {
    rc = upload(p);
    if (rc) qmcb_set_error(std::string("plan upload: ") + cudaGetErrorString((cudaError_t)rc));
  }
  if (rc) { qmcb_plan_destroy(p); return rc; }
  *out = p;
  return 0;
}

extern "C" int qmcb_plan_update(qmcb_plan *p, const qmcb_system *sys) {
  if (!p) return QMCB_EINVAL;
  DeviceGuard guard(p->device);
  qmcb_spec_free(p);        // the structure may have changed; compiled modules stay cached by structure
  ++p->version;
  int rc = qmcb_build_tables(sys, p);
  if (rc == 0) rc = qmcb_choose_launch(p);
  if (rc == 0) rc = qmcb_choose_backward(p);
  if (rc == 0 && p->device >= 0) {
    rc = upload(p);
    if (rc) qmcb_set_error(std::string("plan upload: ") + cudaGetErrorString((cudaError_t)rc));
  }
  return rc;
}

extern "C" void qmcb_plan_destroy(qmcb_plan *p) {
  if (!p) return;
  DeviceGuard guard(p->device);
  if (p->d_dbl) cudaFree(p->d_dbl);
  if (p->d_int) cudaFree(p->d_int);
  if (p->d_mo_full) cudaFree(p->d_mo_full);
  if (p->d_bwd_tiles) cudaFree(p->d_bwd_tiles);
  if (p->d_ticket) cudaFree(p->d_ticket);
  if (p->d_flat_dbl) cudaFree(p->d_flat_dbl);
  if (p->d_flat_int) cudaFree(p->d_flat_int);
  qmcb_spec_free(p);
  delete p;
}

extern "C" int qmcb_plan_info(const qmcb_plan *p, int what) {
  if (!p) return QMCB_EINVAL;
  switch (what) {
    case 0: return p->sys.nshell;
    case 1: return p->sys.nprim;
    case 2: return p->sys.ncomp;
    case 3: return p->sys.nmu;
    case 4: return p->sys.nuu;
    case 5: return p->sys.nud;
    case 6: return p->cfg_eloc.tw;
    case 7: return p->cfg_eloc.threads;
    case 8: return p->cfg_eloc.smem;
    case 9: return p->cfg_psi.tw;
    case 10: return p->bwd.tw;
    case 11: return p->bwd.smem;
    case 12: return p->bwd.ntile_mo + p->bwd.ntile_ao;
    case 13: {   // structure-specialised kernels: compiles (and loads) them now; 1 = in use
      std::string why;
      const int on = qmcb_spec_status(p, &why);
      if (!on) qmcb_set_error("qmcb: generic kernels in use: " + why);
      return on;
    }
    case 14: return qmcb_spec_eligible(p);     // 0: none, 1: one walker per thread, 2: warp tiles
    case 15: {                                   // kind of the specialised kernels in use (compiles them now)
      std::string why;
      return qmcb_spec_status(p, &why) ? qmcb_spec_kind(p) : 0;
    }
    default: return QMCB_EINVAL;
  }
}
