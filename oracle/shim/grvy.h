/* oracle/shim/grvy.h -- TEST INFRASTRUCTURE ONLY.
 * Stand-in for libGRVY's input parser (the reference needs GRVY >= 0.32, configure.ac:51; it is
 * not in this image).  Only the surface the reference uses is provided:
 *   GRVY::GRVY_Input_Class::{Open, Read_Var (6 overloads), Fdump, Close},
 *   grvy_check_file_path, grvy_log_setlevel, GRVY_NOLOG, GRVY_INFO.
 * File syntax handled: `key = value  # comment`, `[Section]` -> key "Section/key",
 * booleans True/False (case-insensitive, also 1/0), 'quoted strings'. */
#ifndef LP_ORACLE_SHIM_GRVY_H
#define LP_ORACLE_SHIM_GRVY_H
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <sys/stat.h>
#include <vector>
#define GRVY_NOLOG 0
#define GRVY_INFO 1
static inline void grvy_log_setlevel(int) {}
static inline int grvy_check_file_path(const char *path)
{
  std::string p(path);
  size_t pos = 0;
  while ((pos = p.find('/', pos + 1)) != std::string::npos) {
    std::string dir = p.substr(0, pos);
    if (!dir.empty()) mkdir(dir.c_str(), 0755);
  }
  return 0;
}
namespace GRVY {
class GRVY_Input_Class {
  std::map<std::string, std::string> kv_;
  std::vector<std::string> order_;
  static std::string trim(const std::string &s)
  {
    size_t a = s.find_first_not_of(" \t\r\n");
    if (a == std::string::npos) return "";
    size_t b = s.find_last_not_of(" \t\r\n");
    return s.substr(a, b - a + 1);
  }
  bool find(const char *key, std::string &out) const
  {
    std::map<std::string, std::string>::const_iterator it = kv_.find(key);
    if (it == kv_.end()) return false;
    out = it->second;
    return true;
  }
public:
  int Open(const char *fname)
  {
    std::ifstream in(fname);
    if (!in.good()) return 0;
    std::string line, section;
    while (std::getline(in, line)) {
      bool inq = false; size_t cut = std::string::npos;
      for (size_t i = 0; i < line.size(); i++) {
        if (line[i] == '\'' || line[i] == '"') inq = !inq;
        if (line[i] == '#' && !inq) { cut = i; break; }
      }
      if (cut != std::string::npos) line = line.substr(0, cut);
      line = trim(line);
      if (line.empty()) continue;
      if (line[0] == '[') { size_t e = line.find(']'); section = trim(line.substr(1, e - 1)); continue; }
      size_t eq = line.find('=');
      if (eq == std::string::npos) continue;
      std::string k = trim(line.substr(0, eq)), v = trim(line.substr(eq + 1));
      if (v.size() >= 2 && (v[0] == '\'' || v[0] == '"') && v[v.size() - 1] == v[0]) v = v.substr(1, v.size() - 2);
      std::string full = section.empty() ? k : section + "/" + k;
      if (kv_.find(full) == kv_.end()) order_.push_back(full);
      kv_[full] = v;
    }
    return 1;
  }
  int Close() { return 1; }
  int Fdump(const char *prefix)
  {
    for (size_t i = 0; i < order_.size(); i++) printf("%s%s = %s\n", prefix, order_[i].c_str(), kv_[order_[i]].c_str());
    return 1;
  }
  int Read_Var(const char *key, int *v) { std::string s; if (!find(key, s)) return 0; *v = atoi(s.c_str()); return 1; }
  int Read_Var(const char *key, double *v) { std::string s; if (!find(key, s)) return 0; *v = atof(s.c_str()); return 1; }
  int Read_Var(const char *key, std::string *v) { std::string s; if (!find(key, s)) return 0; *v = s; return 1; }
  int Read_Var(const char *key, int *v, int dflt) { if (!Read_Var(key, v)) *v = dflt; return 1; }
  int Read_Var(const char *key, double *v, double dflt) { if (!Read_Var(key, v)) *v = dflt; return 1; }
  int Read_Var(const char *key, bool *v, bool dflt)
  {
    std::string s;
    if (!find(key, s)) { *v = dflt; return 1; }
    std::transform(s.begin(), s.end(), s.begin(), ::tolower);
    *v = (s == "true" || s == "1" || s == "yes" || s == ".true.");
    return 1;
  }
};
}
#endif
