// ORACLE — TEST INFRASTRUCTURE ONLY (see tables.hpp).
#include "tables.hpp"

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <vector>

namespace oracle {

namespace {

// Fortran list-directed numeric token -> double ("0.1D+01" uses D exponents).
double fnum(std::string s) {
  for (auto& ch : s)
    if (ch == 'D' || ch == 'd') ch = 'E';
  char* end = nullptr;
  double v = std::strtod(s.c_str(), &end);
  if (end == s.c_str()) throw std::runtime_error("tables: bad numeric token '" + s + "'");
  return v;
}

std::vector<std::string> split(const std::string& line) {
  std::vector<std::string> out;
  std::string tok;
  std::istringstream is(line);
  while (is >> tok) out.push_back(tok);
  return out;
}

struct Lines {
  std::ifstream f;
  std::string name;
  explicit Lines(const std::string& path) : f(path), name(path) {
    if (!f) throw std::runtime_error("tables: cannot open " + path);
  }
  std::vector<std::string> next(size_t need) {
    std::string line;
    while (std::getline(f, line)) {
      auto t = split(line);
      if (t.empty()) continue;
      if (t.size() < need) throw std::runtime_error("tables: short line in " + name + ": " + line);
      return t;
    }
    throw std::runtime_error("tables: unexpected EOF in " + name);
  }
};

void zero(CcpolTables& t) { std::memset(&t, 0, sizeof(t)); }

// data1, proc_sapt5sf_new_ncd.f:1266-1356
void read_sapt_text(const std::string& path, CcpolTables& t) {
  Lines L(path);
  int nparm = (int)fnum(L.next(1)[0]);
  for (int i = 0; i < nparm; ++i) {
    auto k = L.next(4);
    int ityp = (int)fnum(k[0]), inumpar = (int)fnum(k[1]);
    double val = fnum(k[2]);
    t.PARAM(inumpar, ityp) = val;
    if (inumpar <= 9) t.PARAM(inumpar, ityp) = 18.22262373 * t.PARAM(inumpar, ityp);  // :1315-1317
  }
  int nparab = (int)fnum(L.next(1)[0]);
  for (int i = 0; i < nparab; ++i) {
    auto k = L.next(5);
    int t1 = (int)fnum(k[0]), t2 = (int)fnum(k[1]), ip = (int)fnum(k[2]);
    double val = fnum(k[3]);
    t.PARAB(ip, t1, t2) = val;  // :1326-1327 symmetrised
    t.PARAB(ip, t2, t1) = val;
  }
  L.next(5);  // ntpot, idonl, iopt, iweight, iasdone
  L.next(9);  // TOLF ... SAFETL
  L.next(3);  // R_0, isyst, npowers
  L.next(1);  // RCOND
  t.numlin = (int)fnum(L.next(1)[0]);
  if (t.numlin > 1000) throw std::runtime_error("tables: numlin > 1000");
  for (int i = 0; i < t.numlin; ++i) t.c[i] = fnum(L.next(1)[0]);
}

// ccpol8s_dimer(imode=-1), proc_ccpol8s-dimer_xyz_ncd.f:40-58
void read_ccpol8s_text(const std::string& path, CcpolTables& t) {
  Lines L(path);
  t.nparsall = (int)fnum(L.next(1)[0]);
  if (t.nparsall > 1000) throw std::runtime_error("stop 010");
  for (int i = 1; i <= t.nparsall; ++i) {
    auto k = L.next(2);
    if ((int)fnum(k[0]) != i) throw std::runtime_error("stop 020");
    t.params[i - 1] = fnum(k[1]);
  }
  t.nlin0 = (int)fnum(L.next(1)[0]);
  if (t.nlin0 > 2000) throw std::runtime_error("stop 030");
  for (int i = 1; i <= t.nlin0; ++i) {
    auto k = L.next(2);
    if ((int)fnum(k[0]) != i) throw std::runtime_error("stop 040");
    t.cc[i - 1] = fnum(k[1]);
  }
}

// read_cc_data, main_CCpol-8sf.f:822-923
void read_ccdata_text(const std::string& path, CcpolTables& t) {
  Lines L(path);
  L.next(1);  // "sites"
  for (int i = 1; i <= 25; ++i) {
    auto k = L.next(3);
    for (int j = 1; j <= 3; ++j) t.sites[(i - 1) * 3 + (j - 1)] = fnum(k[j - 1]);
  }
  L.next(1);
  {
    auto k = L.next(5);
    for (int j = 0; j < 5; ++j) t.chrg[j] = fnum(k[j]);
  }
  L.next(1);
  {
    auto k = L.next(5);
    for (int j = 0; j < 5; ++j) t.ind_charge[j] = (int)fnum(k[j]);
  }
  auto block = [&](int* dst, int nrow) {
    L.next(1);
    for (int i = 1; i <= nrow; ++i) {
      auto k = L.next(nrow);
      for (int j = 1; j <= nrow; ++j) dst[CcpolTables::IJ(i, j)] = (int)fnum(k[j - 1]);
    }
  };
  block(t.ind_beta, 25);
  block(t.ind_d1, 5);
  block(t.ind_d6, 3);
  block(t.ind_d8, 3);
  block(t.ind_d10, 3);
  block(t.ind_c6, 3);
  block(t.ind_c8, 3);
  block(t.ind_c10, 3);
}

std::map<std::string, std::vector<double>> read_packed(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("tables: cannot open " + path);
  std::map<std::string, std::vector<double>> m;
  std::string line;
  while (std::getline(f, line)) {
    if (line.empty() || line[0] == '#') continue;
    auto h = split(line);
    if (h.size() != 2) throw std::runtime_error("tables: bad packed header: " + line);
    size_t cnt = (size_t)std::strtoul(h[1].c_str(), nullptr, 10);
    std::vector<double> v;
    v.reserve(cnt);
    std::string tok;
    while (v.size() < cnt && (f >> tok)) v.push_back(std::strtod(tok.c_str(), nullptr));
    if (v.size() != cnt) throw std::runtime_error("tables: truncated packed block " + h[0]);
    std::getline(f, line);  // rest of the last value line
    m[h[0]] = v;
  }
  return m;
}

const std::vector<double>& need(const std::map<std::string, std::vector<double>>& m, const char* key,
                                size_t n) {
  auto it = m.find(key);
  if (it == m.end() || it->second.size() != n)
    throw std::runtime_error(std::string("tables: packed key missing or wrong size: ") + key);
  return it->second;
}

}  // namespace

bool surface_switches(int isurf, CcpolTables& t, std::string& saptfile) {
  // main_CCpol-8sf.f:28-107
  struct Row { int iembed, ipotparts; const char* f; int icc; };
  static const Row rows[10] = {
      {1, 1, "data_SAPT5spf_2014", 1},   {1, 1, "data_SAPT5spfIR_2014", 1},
      {2, 1, "data_SAPT5spfIR_2006", 1}, {1, 1, "data_SAPT5spfIR_2006", 1},
      {1, 1, "data_SAPT5spf_2014", 0},   {1, 1, "data_SAPT5spfIR_2014", 0},
      {1, 1, "data_SAPT5spfIR_2006", 0}, {1, 0, "data_SAPT5spf_2006", 0},
      {1, 0, "data_SAPT5spfIR_2006", 0}, {2, 1, "data_SAPT5spfIR_2014", 1}};
  if (isurf < 1 || isurf > 10) return false;
  const Row& r = rows[isurf - 1];
  t.iembed = r.iembed;
  t.ipotparts = r.ipotparts;
  t.icc = r.icc;
  saptfile = r.f;
  return true;
}

void load_text(const std::string& dir, int isurf, int iemon, CcpolTables& t) {
  zero(t);
  std::string sapt;
  if (!surface_switches(isurf, t, sapt)) throw std::runtime_error("wrong value of isurf");
  if (iemon != 0 && iemon != 1) throw std::runtime_error("wrong value of iemonomer");
  t.iemonomer = iemon;
  t.pjt2_r8 = 1;
  read_sapt_text(dir + "/" + sapt, t);
  read_ccpol8s_text(dir + "/data_CCpol8s", t);
  read_ccdata_text(dir + "/data_ccdata", t);
}

void load_packed(const std::string& sapt_tbl, const std::string& ccpol8s_tbl, int isurf, int iemon,
                 CcpolTables& t) {
  zero(t);
  std::string sapt;
  if (!surface_switches(isurf, t, sapt)) throw std::runtime_error("wrong value of isurf");
  if (iemon != 0 && iemon != 1) throw std::runtime_error("wrong value of iemonomer");
  t.iemonomer = iemon;
  t.pjt2_r8 = 1;
  auto a = read_packed(sapt_tbl);
  auto b = read_packed(ccpol8s_tbl);
  const auto& p = need(a, "param", 108);
  for (int i = 0; i < 108; ++i) t.param[i] = p[i];
  const auto& pb = need(a, "parab", 3024);
  for (int i = 0; i < 3024; ++i) t.parab[i] = pb[i];
  auto it = a.find("c");
  if (it == a.end() || it->second.size() > 1000) throw std::runtime_error("tables: packed key c");
  t.numlin = (int)it->second.size();
  for (int i = 0; i < t.numlin; ++i) t.c[i] = it->second[i];
  auto ip = b.find("params");
  auto ic = b.find("cc");
  if (ip == b.end() || ic == b.end()) throw std::runtime_error("tables: packed keys params/cc");
  t.nparsall = (int)ip->second.size();
  t.nlin0 = (int)ic->second.size();
  for (int i = 0; i < t.nparsall; ++i) t.params[i] = ip->second[i];
  for (int i = 0; i < t.nlin0; ++i) t.cc[i] = ic->second[i];
  const auto& s = need(b, "sites", 75);
  for (int i = 0; i < 75; ++i) t.sites[i] = s[i];
  const auto& q = need(b, "chrg", 25);
  for (int i = 0; i < 25; ++i) t.chrg[i] = q[i];
  const auto& ich = need(b, "ind_charge", 25);
  for (int i = 0; i < 25; ++i) t.ind_charge[i] = (int)ich[i];
  struct { const char* k; int* d; } blocks[] = {
      {"ind_beta", t.ind_beta}, {"ind_d1", t.ind_d1},   {"ind_d6", t.ind_d6}, {"ind_d8", t.ind_d8},
      {"ind_d10", t.ind_d10},   {"ind_c6", t.ind_c6},   {"ind_c8", t.ind_c8}, {"ind_c10", t.ind_c10}};
  for (auto& bl : blocks) {
    const auto& v = need(b, bl.k, 625);
    for (int i = 0; i < 625; ++i) bl.d[i] = (int)v[i];
  }
}

}  // namespace oracle
