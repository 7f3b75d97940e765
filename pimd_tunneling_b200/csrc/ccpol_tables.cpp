// Host-side loader for the CCpol-8sf parameter tables.
// Replaces, for the drop-in library, the reference's three readers:
//   data1            proc_sapt5sf_new_ncd.f:1266-1356   (unit 55, "./data_SAPT5spfIR_2006" for isurf=3,
//                                                        main_CCpol-8sf.f:44-50)
//   ccpol8s_dimer(-1) proc_ccpol8s-dimer_xyz_ncd.f:40-58 ("data_CCpol8s")
//   read_cc_data     main_CCpol-8sf.f:822-923            ("data_ccdata")
#include "ccpol_tables.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

namespace pimdk {
namespace {

std::string g_msg;

// One record of a Fortran list-directed READ: the leading whitespace-separated fields of a line.
struct Record {
  std::vector<std::string> f;
  double num(size_t i) const {
    std::string s = f.at(i);
    for (char& ch : s)
      if (ch == 'D' || ch == 'd') ch = 'e';
    return std::strtod(s.c_str(), nullptr);
  }
  long integer(size_t i) const { return std::strtol(f.at(i).c_str(), nullptr, 10); }
};

class RecordFile {
 public:
  explicit RecordFile(const std::string& path) : fp_(std::fopen(path.c_str(), "r")), path_(path) {}
  ~RecordFile() {
    if (fp_) std::fclose(fp_);
  }
  bool ok() const { return fp_ != nullptr; }
  // next non-blank line split into fields; false at EOF
  bool read(Record* r, size_t min_fields) {
    char buf[4096];
    while (std::fgets(buf, sizeof buf, fp_)) {
      r->f.clear();
      for (char* tok = std::strtok(buf, " \t\r\n,"); tok; tok = std::strtok(nullptr, " \t\r\n,")) r->f.emplace_back(tok);
      if (r->f.empty()) continue;
      if (r->f.size() < min_fields) {
        g_msg = "short record in " + path_;
        return false;
      }
      return true;
    }
    g_msg = "unexpected end of " + path_;
    return false;
  }

 private:
  FILE* fp_;
  std::string path_;
};

bool exists(const std::string& p) {
  FILE* f = std::fopen(p.c_str(), "r");
  if (!f) return false;
  std::fclose(f);
  return true;
}

#define NEED(cond)              \
  do {                          \
    if (!(cond)) return false;  \
  } while (0)

bool sapt_from_text(const std::string& path, CcpolHost* h) {
  RecordFile in(path);
  Record r;
  NEED(in.ok());
  NEED(in.read(&r, 1));
  long n1 = r.integer(0);
  for (long i = 0; i < n1; ++i) {
    NEED(in.read(&r, 4));
    long t = r.integer(0), k = r.integer(1);
    if (t < 1 || t > 6 || k < 1 || k > 18) { g_msg = "one-site index out of range in " + path; return false; }
    double v = r.num(2);
    if (k <= 9) v = 18.22262373 * v;  // charge rows to kcal/mol units, :1315-1317
    h->param[(t - 1) * 18 + (k - 1)] = v;
  }
  NEED(in.read(&r, 1));
  long n2 = r.integer(0);
  for (long i = 0; i < n2; ++i) {
    NEED(in.read(&r, 5));
    long ta = r.integer(0), tb = r.integer(1), k = r.integer(2);
    if (ta < 1 || ta > 6 || tb < 1 || tb > 6 || k < 1 || k > 84) { g_msg = "two-site index out of range in " + path; return false; }
    double v = r.num(3);
    h->parab[((tb - 1) * 6 + (ta - 1)) * 84 + (k - 1)] = v;  // symmetrised, :1326-1327
    h->parab[((ta - 1) * 6 + (tb - 1)) * 84 + (k - 1)] = v;
  }
  NEED(in.read(&r, 5));  // ntpot idonl iopt iweight iasdone
  NEED(in.read(&r, 9));  // optimiser settings
  NEED(in.read(&r, 3));
  NEED(in.read(&r, 1));
  NEED(in.read(&r, 1));
  h->numlin = (int)r.integer(0);
  if (h->numlin < 0 || h->numlin > 1000) { g_msg = "numlin out of range in " + path; return false; }
  for (int i = 0; i < h->numlin; ++i) {
    NEED(in.read(&r, 1));
    h->c[i] = r.num(0);
  }
  return true;
}

bool cc8s_from_text(const std::string& ppath, const std::string& dpath, CcpolHost* h) {
  {
    RecordFile in(ppath);
    Record r;
    NEED(in.ok());
    NEED(in.read(&r, 1));
    h->nparsall = (int)r.integer(0);
    if (h->nparsall > 1000) { g_msg = "stop 010"; return false; }
    for (int i = 1; i <= h->nparsall; ++i) {
      NEED(in.read(&r, 2));
      if (r.integer(0) != i) { g_msg = "stop 020"; return false; }
      h->params[i - 1] = r.num(1);
    }
    NEED(in.read(&r, 1));
    h->nlin0 = (int)r.integer(0);
    if (h->nlin0 > 2000) { g_msg = "stop 030"; return false; }
    for (int i = 1; i <= h->nlin0; ++i) {
      NEED(in.read(&r, 2));
      if (r.integer(0) != i) { g_msg = "stop 040"; return false; }
      h->cc[i - 1] = r.num(1);
    }
  }
  RecordFile in(dpath);
  Record r;
  NEED(in.ok());
  NEED(in.read(&r, 1));  // "sites"
  for (int s = 0; s < 25; ++s) {
    NEED(in.read(&r, 3));
    for (int j = 0; j < 3; ++j) h->sites[s * 3 + j] = r.num(j);
  }
  NEED(in.read(&r, 1));
  NEED(in.read(&r, 5));
  for (int j = 0; j < 5; ++j) h->chrg[j] = r.num(j);
  NEED(in.read(&r, 1));
  NEED(in.read(&r, 5));
  for (int j = 0; j < 5; ++j) h->ind_charge[j] = (int)r.integer(j);
  struct Blk { int* dst; int rows; };
  Blk blks[] = {{h->ind_beta, 25}, {h->ind_d1, 5}, {h->ind_d6, 3}, {h->ind_d8, 3},
                {h->ind_d10, 3},   {h->ind_c6, 3}, {h->ind_c8, 3}, {h->ind_c10, 3}};
  for (const Blk& b : blks) {
    NEED(in.read(&r, 1));
    for (int i = 0; i < b.rows; ++i) {
      NEED(in.read(&r, (size_t)b.rows));
      for (int j = 0; j < b.rows; ++j) b.dst[j * 25 + i] = (int)r.integer(j);
    }
  }
  return true;
}

// packed "key count\nvalues..." (tools/pack_ccpol_tables.py)
bool read_block_file(const std::string& path, std::vector<std::pair<std::string, std::vector<double>>>* out) {
  FILE* fp = std::fopen(path.c_str(), "r");
  if (!fp) { g_msg = "cannot open " + path; return false; }
  char key[64];
  char line[4096];
  long cnt;
  while (std::fgets(line, sizeof line, fp)) {
    if (line[0] == '#' || line[0] == '\n') continue;
    if (std::sscanf(line, "%63s %ld", key, &cnt) != 2 || cnt < 0) { std::fclose(fp); g_msg = "bad block header in " + path; return false; }
    std::vector<double> v((size_t)cnt);
    for (long i = 0; i < cnt; ++i) {
      char tok[64];
      if (std::fscanf(fp, "%63s", tok) != 1) { std::fclose(fp); g_msg = "truncated block in " + path; return false; }
      v[(size_t)i] = std::strtod(tok, nullptr);
    }
    if (!std::fgets(line, sizeof line, fp)) line[0] = 0;  // rest of last line
    out->emplace_back(key, std::move(v));
  }
  std::fclose(fp);
  return true;
}

const std::vector<double>* find(const std::vector<std::pair<std::string, std::vector<double>>>& m, const char* k) {
  for (auto& kv : m)
    if (kv.first == k) return &kv.second;
  return nullptr;
}

bool from_packed(const std::string& sapt, const std::string& cc, CcpolHost* h) {
  std::vector<std::pair<std::string, std::vector<double>>> a, b;
  NEED(read_block_file(sapt, &a));
  NEED(read_block_file(cc, &b));
  auto copy_d = [&](const std::vector<std::pair<std::string, std::vector<double>>>& m, const char* k, double* dst,
                    size_t cap, int* count) {
    const std::vector<double>* v = find(m, k);
    if (!v || v->size() > cap) { g_msg = std::string("packed tables: bad block ") + k; return false; }
    for (size_t i = 0; i < v->size(); ++i) dst[i] = (*v)[i];
    if (count) *count = (int)v->size();
    return true;
  };
  auto copy_i = [&](const char* k, int* dst, size_t n) {
    const std::vector<double>* v = find(b, k);
    if (!v || v->size() != n) { g_msg = std::string("packed tables: bad block ") + k; return false; }
    for (size_t i = 0; i < n; ++i) dst[i] = (int)(*v)[i];
    return true;
  };
  NEED(copy_d(a, "param", h->param, 108, nullptr));
  NEED(copy_d(a, "parab", h->parab, 3024, nullptr));
  NEED(copy_d(a, "c", h->c, 1000, &h->numlin));
  NEED(copy_d(b, "params", h->params, 1000, &h->nparsall));
  NEED(copy_d(b, "cc", h->cc, 2000, &h->nlin0));
  NEED(copy_d(b, "sites", h->sites, 75, nullptr));
  NEED(copy_d(b, "chrg", h->chrg, 25, nullptr));
  NEED(copy_i("ind_charge", h->ind_charge, 25));
  NEED(copy_i("ind_beta", h->ind_beta, 625));
  NEED(copy_i("ind_d1", h->ind_d1, 625));
  NEED(copy_i("ind_d6", h->ind_d6, 625));
  NEED(copy_i("ind_d8", h->ind_d8, 625));
  NEED(copy_i("ind_d10", h->ind_d10, 625));
  NEED(copy_i("ind_c6", h->ind_c6, 625));
  NEED(copy_i("ind_c8", h->ind_c8, 625));
  NEED(copy_i("ind_c10", h->ind_c10, 625));
  return true;
}

}  // namespace

const char* load_ccpol_tables(const char* dir, int isurf, CcpolHost* out) {
  std::memset(out, 0, sizeof(*out));
  g_msg.clear();
  // main_CCpol-8sf.f:28-107: embedding, potparts variant, SAPT data file, CCpol-8s correction on/off
  struct Row { int iembed, ipotparts; const char* f; int icc; };
  static const Row rows[10] = {
      {1, 1, "SAPT5spf_2014", 1},   {1, 1, "SAPT5spfIR_2014", 1}, {2, 1, "SAPT5spfIR_2006", 1}, {1, 1, "SAPT5spfIR_2006", 1},
      {1, 1, "SAPT5spf_2014", 0},   {1, 1, "SAPT5spfIR_2014", 0}, {1, 1, "SAPT5spfIR_2006", 0}, {1, 0, "SAPT5spf_2006", 0},
      {1, 0, "SAPT5spfIR_2006", 0}, {2, 1, "SAPT5spfIR_2014", 1}};
  if (isurf < 1 || isurf > 10) {
    g_msg = "wrong value of isurf";
    return g_msg.c_str();
  }
  const Row& r = rows[isurf - 1];
  std::string d(dir ? dir : "."), f(r.f);
  bool ok;
  if (exists(d + "/data_" + f) && exists(d + "/data_CCpol8s") && exists(d + "/data_ccdata")) {
    ok = sapt_from_text(d + "/data_" + f, out) && cc8s_from_text(d + "/data_CCpol8s", d + "/data_ccdata", out);
  } else if (exists(d + "/sapt_" + f + ".tbl") && exists(d + "/ccpol8s.tbl")) {
    ok = from_packed(d + "/sapt_" + f + ".tbl", d + "/ccpol8s.tbl", out);
  } else {
    g_msg = "no CCpol-8sf data files (data_" + f + "/data_CCpol8s/data_ccdata or *.tbl) in " + d;
    ok = false;
  }
  if (!ok && g_msg.empty()) g_msg = "failed to read CCpol-8sf data files in " + d;
  out->isurf = isurf;
  out->iembed = r.iembed;
  out->ipotparts = r.ipotparts;
  out->icc = r.icc;
  return ok ? "" : g_msg.c_str();
}

const char* build_ccpol_dev(const CcpolHost& h, int iemonomer, CcpolDev* o) {
  std::memset(o, 0, sizeof(*o));
  g_msg.clear();
  if (h.numlin < 1 || h.numlin > 568 || h.nlin0 != 144 || h.nparsall != 134) {
    g_msg = "unexpected CCpol-8sf table sizes (need <= 568 / 144 / 134)";
    return g_msg.c_str();
  }
  for (int t = 0; t < kNType; ++t)
    for (int k = 0; k < 18; ++k) o->param[t * 18 + k] = h.param[t * 18 + k];
  for (int tb = 0; tb < kNType; ++tb)
    for (int ta = 0; ta < kNType; ++ta)
      for (int k = 0; k < 84; ++k) o->parab[(tb * kNType + ta) * 84 + k] = h.parab[(tb * 6 + ta) * 84 + k];
  // type 6 must be unused for the trimming above to be valid
  for (int k = 0; k < 18; ++k)
    if (h.param[5 * 18 + k] != 0.0) { g_msg = "site type 6 carries parameters; not supported"; return g_msg.c_str(); }
  std::memcpy(o->c, h.c, 568 * sizeof(double));   // (zero beyond numlin)
  o->iembed = (uint8_t)h.iembed;
  o->icc = (uint8_t)h.icc;
  o->potparts_old = (uint8_t)(h.ipotparts == 0);
  std::memcpy(o->cc, h.cc, 144 * sizeof(double));
  std::memcpy(o->params, h.params, 134 * sizeof(double));
  std::memcpy(o->sites, h.sites, 75 * sizeof(double));
  for (int i = 0; i < 5; ++i) o->chrg[i] = h.chrg[i];
  for (int i = 5; i < 25; ++i)
    if (h.chrg[i] != 0.0 || h.ind_charge[i] != 0) { g_msg = "charged CCpol-8s site beyond 5; not supported"; return g_msg.c_str(); }
  // static replay of poten's itypus bookkeeping for the fixed type vector (set_sites :1748-1755);
  // a pair type owns basis functions iff its exponent can be non-zero (beta>0 branch, potparts :427)
  static const int types[8] = {1, 2, 2, 3, 3, 4, 4, 5};
  int next = 1;
  for (int ia = 0; ia < 8; ++ia)
    for (int ib = 0; ib < 8; ++ib) {
      int ta = types[ia], tb = types[ib];
      const double* pb = &h.parab[((tb - 1) * 6 + (ta - 1)) * 84];
      bool has_exp = pb[0] != 0.0 || pb[40] != 0.0 || pb[41] != 0.0 || pb[45] != 0.0 || pb[46] != 0.0;
      if (!has_exp) continue;
      int e = (tb - 1) * kNType + (ta - 1), et = (ta - 1) * kNType + (tb - 1);
      if (o->itu_s[e] == 0) {
        o->itu_s[e] = o->itu_s[et] = (int16_t)next;
        next += 40;
      }
      if (ta != tb && o->itu_a[e] == 0) {
        o->itu_a[e] = o->itu_a[et] = (int16_t)next;
        next += 28;
      }
    }
  for (int tb = 0; tb < kNType; ++tb)
    for (int ta = 0; ta < kNType; ++ta) {
      const double* pb = &h.parab[(tb * 6 + ta) * 84];
      uint8_t f = 0;
      if (pb[0] != 0.0 || pb[40] != 0.0 || pb[41] != 0.0 || pb[45] != 0.0 || pb[46] != 0.0) f |= 1;
      if (pb[5] != 0.0) f |= 2;                                  // dmp1
      if (pb[6] != 0.0 || pb[7] != 0.0 || pb[8] != 0.0) f |= 4;  // dmp6, dmp8, dmp10
      o->pairflags[tb * kNType + ta] = f;
    }
  for (int e = 0; e < kNType * kNType; ++e)   // the kernels read coefficient groups with 16-byte loads
    if ((o->itu_s[e] && (o->itu_s[e] - 1) % 4) || (o->itu_a[e] && (o->itu_a[e] - 1) % 4)) {
      g_msg = "linear-coefficient blocks are not 4-aligned";
      return g_msg.c_str();
    }
  if (next - 1 != h.numlin) {
    g_msg = "linear-coefficient index map does not cover the coefficient table";
    return g_msg.c_str();
  }
  for (int b = 0; b < 25; ++b)
    for (int a = 0; a < 25; ++a) {
      int v = h.ind_beta[b * 25 + a];
      if (v < 0 || v > 134) { g_msg = "ind_beta out of range"; return g_msg.c_str(); }
      o->ind_beta[b * 25 + a] = (uint8_t)v;
    }
  for (int i = 0; i < 5; ++i) o->ind_charge[i] = (uint8_t)h.ind_charge[i];
  for (int b = 0; b < 5; ++b)
    for (int a = 0; a < 5; ++a) o->ind_d1[b * 5 + a] = (uint8_t)h.ind_d1[b * 25 + a];
  for (int b = 0; b < 3; ++b)
    for (int a = 0; a < 3; ++a) {
      o->ind_d6[b * 3 + a] = (uint8_t)h.ind_d6[b * 25 + a];
      o->ind_d8[b * 3 + a] = (uint8_t)h.ind_d8[b * 25 + a];
      o->ind_d10[b * 3 + a] = (uint8_t)h.ind_d10[b * 25 + a];
      o->ind_c6[b * 3 + a] = (uint8_t)h.ind_c6[b * 25 + a];
      o->ind_c8[b * 3 + a] = (uint8_t)h.ind_c8[b * 25 + a];
      o->ind_c10[b * 3 + a] = (uint8_t)h.ind_c10[b * 25 + a];
    }
  // dispersion/electrostatic index blocks must vanish outside their 3x3 / 5x5 corners
  for (int b = 0; b < 25; ++b)
    for (int a = 0; a < 25; ++a) {
      if ((a >= 3 || b >= 3) && h.ind_d6[b * 25 + a] != 0) { g_msg = "ind_d6 outside 3x3"; return g_msg.c_str(); }
    }
  // site classes of the rigid model: consecutive sites whose ind_beta rows AND columns coincide
  o->ncls = 0;
  for (int a = 0; a < 25; ++a) {
    bool same = a > 0;
    for (int b = 0; same && b < 25; ++b)
      same = h.ind_beta[b * 25 + a] == h.ind_beta[b * 25 + a - 1] && h.ind_beta[a * 25 + b] == h.ind_beta[(a - 1) * 25 + b];
    if (!same) o->cls_start[o->ncls++] = (uint8_t)a;
  }
  o->cls_start[o->ncls] = 25;
  for (int c = 0; c < o->ncls; ++c) {
    const int sz = o->cls_start[c + 1] - o->cls_start[c];
    if (sz != 1 && sz != 2 && sz != 4) { g_msg = "CCpol-8s site classes must have 1, 2 or 4 sites"; return g_msg.c_str(); }
  }
  for (int b = 0; b < 25; ++b)
    for (int a = 0; a < 25; ++a)
      if (h.ind_beta[b * 25 + a] == 0) { g_msg = "ind_beta has empty entries; not supported"; return g_msg.c_str(); }
  // ---- U0 sweep: the 36 bins (site-class pairs) as warp tasks, largest first; a bin's pairs are walked in the
  // reference's order: block (A-class ca) x (B-class cb), then block (A-class cb) x (B-class ca)
  {
    struct Bin { int ca, cb, i0, npairs; };
    std::vector<Bin> bins;
    for (int ca = 0; ca < o->ncls; ++ca)
      for (int cb = ca; cb < o->ncls; ++cb) {
        const int a0 = o->cls_start[ca], na = o->cls_start[ca + 1] - a0;
        const int b0 = o->cls_start[cb], nb = o->cls_start[cb + 1] - b0;
        const int ib = h.ind_beta[b0 * 25 + a0];
        int indlin = ib - 98;
        if (indlin < 0) indlin += 65;
        if (indlin < 1 || indlin > 36) { g_msg = "ind_beta maps outside the 36 bins"; return g_msg.c_str(); }
        // every pair of the two blocks must carry the same ind_beta (same bin, same beta)
        for (int i = 0; i < na; ++i)
          for (int j = 0; j < nb; ++j)
            if (h.ind_beta[(b0 + j) * 25 + a0 + i] != ib || h.ind_beta[(a0 + i) * 25 + b0 + j] != ib) {
              g_msg = "ind_beta is not constant on site-class blocks";
              return g_msg.c_str();
            }
        bins.push_back({ca, cb, indlin - 1, na * nb * (ca == cb ? 1 : 2)});
        o->bin_beta[indlin - 1] = h.params[ib - 1];
        if (!(h.params[ib - 1] >= 0.0)) { g_msg = "negative exponent parameter in the CCpol-8s sweep (kernels assume beta >= 0)"; return g_msg.c_str(); }
      }
    if (bins.size() != 36) { g_msg = "expected 36 site-class pair bins"; return g_msg.c_str(); }
    {
      std::vector<int> big(bins.size());
      for (size_t i = 0; i < bins.size(); ++i) big[i] = (int)i;
      std::stable_sort(big.begin(), big.end(), [&](int x, int y) { return bins[x].npairs > bins[y].npairs; });
      for (size_t k = 0; k < big.size(); ++k) {
        const Bin& b = bins[big[k]];
        const int a0 = o->cls_start[b.ca], na = o->cls_start[b.ca + 1] - a0;
        const int b0 = o->cls_start[b.cb], nb = o->cls_start[b.cb + 1] - b0;
        o->tbins[k] = (uint32_t)a0 | (uint32_t)na << 5 | (uint32_t)b0 << 8 | (uint32_t)nb << 13 | (uint32_t)b.i0 << 16;
      }
    }
  }
  o->iemonomer = iemonomer;
  o->V0 = 0.0;
  return "";
}

}  // namespace pimdk
