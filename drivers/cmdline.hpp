// cmdline.hpp -- command-line options for the fidib200 drivers.
//
// Same interface and behaviour as the reference's CmdLineArgParser
// (ref: cxx/CmdLineArgParser.h:22-259, cxx/CmdLineArgParser.cpp:17-59) so that the
// reference's command lines work unchanged on the CUDA drivers:
//   * set(name, default, help) for double / int / std::string / bool; "-h" always exists;
//   * parse(): an unknown "-option" prints "<opt> is not a valid option." and returns
//     false; tokens that are not option names are ignored unless they follow a valued
//     option; a token starting with '-' followed by a digit is a value, not an option;
//     bool options TOGGLE each time they appear;
//   * get<T>(name) returns the value, or the reference's sentinel for unknown names;
//   * help() prints purpose, "<exec> [options]", "Usage:" and one line per option grouped
//     by type (double, int, string, bool), each group in name order.
// Written from scratch around one tagged option table instead of eight maps.
#pragma once

#include <cctype>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <map>
#include <string>

class CmdLineArgParser {
 public:
  CmdLineArgParser() {
    set("-h", false, "Print help.");
    footer_ = "\nfidib200: B200-native FiDiBench drivers (same options as the reference)\n";
  }

  void set(const std::string& name, double dflt, const std::string& help) { put(name, kDouble, help).d = dflt; }
  void set(const std::string& name, int dflt, const std::string& help) { put(name, kInt, help).i = dflt; }
  void set(const std::string& name, const std::string& dflt, const std::string& help) { put(name, kString, help).s = dflt; }
  void set(const std::string& name, const char* dflt, const std::string& help) { put(name, kString, help).s = dflt; }
  void set(const std::string& name, bool dflt, const std::string& help) { put(name, kBool, help).b = dflt; }

  bool parse(int argc, char* argv[]) {
    exec_ = argc > 0 ? argv[0] : "";
    // first pass: every token that looks like an option name must be known
    for (int a = 1; a < argc; ++a) {
      const std::string tok(argv[a]);
      if (looksLikeOption(tok) && opts_.find(tok) == opts_.end()) {
        std::cout << tok << " is not a valid option.\n";
        return false;
      }
    }
    // second pass: a valued option takes the token that follows it; the last
    // occurrence wins; a bool flips once per occurrence
    for (int a = 1; a < argc; ++a) {
      std::map<std::string, Opt>::iterator it = opts_.find(argv[a]);
      if (it == opts_.end()) continue;
      Opt& o = it->second;
      if (o.kind == kBool) {
        o.b = !o.b;
      } else if (a + 1 < argc) {
        const char* val = argv[a + 1];
        if (o.kind == kDouble) o.d = std::atof(val);
        if (o.kind == kInt) o.i = std::atoi(val);
        if (o.kind == kString) o.s = val;
      }
    }
    return true;
  }

  void setPurpose(const std::string& purpose) { purpose_ = purpose; }
  void addFootnote(const std::string& note) { footer_ = note + "\n" + footer_; }

  void help() const {
    std::cout << purpose_ << std::endl;
    std::cout << exec_ << " [options]\n";
    std::cout << "Usage:\n";
    static const char* tag[] = {" <double#> ", " <int#> ", " <string> ", " "};
    for (int kind = kDouble; kind <= kBool; ++kind) {
      for (std::map<std::string, Opt>::const_iterator it = opts_.begin(); it != opts_.end(); ++it) {
        const Opt& o = it->second;
        if (o.kind != kind) continue;
        std::cout << "\t" << it->first << tag[kind] << o.help << " (";
        if (kind == kDouble) std::cout << o.d;
        if (kind == kInt) std::cout << o.i;
        if (kind == kString) std::cout << o.s;
        if (kind == kBool) std::cout << o.b;
        std::cout << ")\n";
      }
    }
    std::cout << footer_ << std::endl;
  }

  template <class T>
  T get(const std::string& name) const;

 private:
  enum Kind { kDouble = 0, kInt = 1, kString = 2, kBool = 3 };
  struct Opt {
    Kind kind;
    double d;
    int i;
    std::string s;
    bool b;
    std::string help;
    Opt() : kind(kBool), d(0), i(0), b(false) {}
  };

  Opt& put(const std::string& name, Kind kind, const std::string& help) {
    Opt& o = opts_[name];
    o.kind = kind;
    o.help = help;
    return o;
  }
  const Opt* find(const std::string& name, Kind kind) const {
    std::map<std::string, Opt>::const_iterator it = opts_.find(name);
    return (it != opts_.end() && it->second.kind == kind) ? &it->second : 0;
  }
  static bool looksLikeOption(const std::string& tok) {
    return tok.size() >= 2 && tok[0] == '-' && !std::isdigit(static_cast<unsigned char>(tok[1]));
  }

  std::map<std::string, Opt> opts_;
  std::string exec_, purpose_, footer_;
};

template <>
inline double CmdLineArgParser::get<double>(const std::string& name) const {
  const Opt* o = find(name, kDouble);
  return o ? o->d : -std::numeric_limits<double>::max();
}
template <>
inline int CmdLineArgParser::get<int>(const std::string& name) const {
  const Opt* o = find(name, kInt);
  return o ? o->i : -std::numeric_limits<int>::max();
}
template <>
inline std::string CmdLineArgParser::get<std::string>(const std::string& name) const {
  const Opt* o = find(name, kString);
  return o ? o->s : std::string();
}
template <>
inline bool CmdLineArgParser::get<bool>(const std::string& name) const {
  const Opt* o = find(name, kBool);
  return o ? o->b : false;
}
