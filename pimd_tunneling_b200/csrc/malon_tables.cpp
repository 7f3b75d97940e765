// Host loader of the malonaldehyde tables (pes_malonaldehyde.f90:45-9347, repacked by tools/pack_malon_tables.py into
// "key count / values" blocks): builds the device image of malon.cuh (records padded to 16-byte multiples, 0-based packed
// distance indices).  C stdio only.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "malon.cuh"

namespace pimdk {

const char* load_malon_tab(const char* dir, MalonTab* out) {
  static std::string msg;
  const std::string path = std::string(dir && dir[0] ? dir : ".") + "/malonaldehyde.tbl";
  FILE* f = std::fopen(path.c_str(), "r");
  if (!f) {
    msg = "cannot open " + path + " (tools/pack_malon_tables.py writes it from pes_malonaldehyde.f90)";
    return msg.c_str();
  }
  int c;
  while ((c = std::fgetc(f)) != EOF && c != '\n') {}   // header line
  double shift = 0.0;
  std::vector<int> imorse, ig1d, ig2d, ig3d;
  std::vector<double> morse, g1d, g2d, g3d;
  char key[64];
  long cnt;
  bool ok = true, have_shift = false;
  while (ok && std::fscanf(f, "%63s %ld", key, &cnt) == 2) {
    const std::string k(key);
    if (cnt < 0 || cnt > 100000) { ok = false; break; }
    if (k == "shift") {
      ok = std::fscanf(f, "%lf", &shift) == 1;
      have_shift = ok;
    } else if (k[0] == 'i') {
      std::vector<int>* v = k == "imorse" ? &imorse : k == "ig1d" ? &ig1d : k == "ig2d" ? &ig2d : k == "ig3d" ? &ig3d : nullptr;
      if (!v) { ok = false; break; }
      v->resize(cnt);
      for (long i = 0; i < cnt && ok; ++i) ok = std::fscanf(f, "%d", &(*v)[i]) == 1 && (*v)[i] >= 1 && (*v)[i] <= kMalDist;
    } else {
      std::vector<double>* v = k == "morse" ? &morse : k == "g1d" ? &g1d : k == "g2d" ? &g2d : k == "g3d" ? &g3d : nullptr;
      if (!v) { ok = false; break; }
      v->resize(cnt);
      for (long i = 0; i < cnt && ok; ++i) ok = std::fscanf(f, "%lf", &(*v)[i]) == 1;
    }
  }
  std::fclose(f);
  if (!ok || !have_shift || (int)imorse.size() != kMalMorse || (int)morse.size() != 3 * kMalMorse || (int)ig1d.size() != kMalG1 ||
      (int)g1d.size() != 4 * kMalG1 || (int)ig2d.size() != 2 * kMalG2 || (int)g2d.size() != 6 * kMalG2 ||
      (int)ig3d.size() != 3 * kMalG3 || (int)g3d.size() != 8 * kMalG3) {
    msg = "malonaldehyde table file is malformed or incomplete: " + path;
    return msg.c_str();
  }
  std::memset(out, 0, sizeof(MalonTab));
  out->shift = shift;
  for (int i = 0; i < kMalMorse; ++i) {
    for (int k = 0; k < 3; ++k) out->morse[4 * i + k] = morse[3 * i + k];
    out->imorse[i] = (uint32_t)(imorse[i] - 1);
  }
  std::memcpy(out->g1d, g1d.data(), sizeof(double) * g1d.size());
  std::memcpy(out->g2d, g2d.data(), sizeof(double) * g2d.size());
  std::memcpy(out->g3d, g3d.data(), sizeof(double) * g3d.size());
  for (int i = 0; i < kMalG1; ++i) out->ig1d[i] = (uint32_t)(ig1d[i] - 1);
  for (int i = 0; i < kMalG2; ++i) out->ig2d[i] = (uint32_t)(ig2d[2 * i] - 1) | ((uint32_t)(ig2d[2 * i + 1] - 1) << 8);
  for (int i = 0; i < kMalG3; ++i)
    out->ig3d[i] = (uint32_t)(ig3d[3 * i] - 1) | ((uint32_t)(ig3d[3 * i + 1] - 1) << 8) | ((uint32_t)(ig3d[3 * i + 2] - 1) << 16);
  return "";
}

}  // namespace pimdk
