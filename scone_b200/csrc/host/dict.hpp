// SCONE dictionary input: insertion-ordered nested dictionaries parsed from the
// OpenFOAM-like ASCII grammar.
//
// Mirrors the behaviour (not the code) of
//   DataStructures/dictionary_class.f90      (get / getOrDefault / keys / isPresent)
//   DataStructures/dictParser_func.f90:26-110 (comments `!` and `//`, tabs/newlines -> space)
//   docs/Dictionary Input.rst                 (grammar; reals must contain a dot; int32 only)
//
// Header-only, host-side. This is input handling, not hot-path arithmetic.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace sb {

struct FatalError : std::runtime_error {
  explicit FatalError(const std::string& where, const std::string& what)
      : std::runtime_error(where + ": " + what) {}
};

class Dict {
 public:
  enum Kind { SCALAR, LIST, TOKENS, DICT };
  struct Entry {
    std::string key;
    Kind kind = SCALAR;
    std::vector<std::string> tok;      // SCALAR: 1 token; LIST/TOKENS: many
    std::shared_ptr<Dict> sub;          // DICT
  };

  // ---- construction -------------------------------------------------------
  static Dict fromFile(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw FatalError("fileToDict", "cannot open file: " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return fromString(ss.str());
  }

  static Dict fromString(const std::string& text) {
    std::string s = stripComments(text);
    s.push_back('}');
    size_t pos = 0;
    Dict d;
    parse(d, s, pos);
    // pos is one past the closing '}' that we appended
    for (size_t i = pos; i < s.size(); ++i)
      if (s[i] != ' ') throw FatalError("fileToDict", "extra '}' bracket somewhere");
    return d;
  }

  // ---- enquiry ------------------------------------------------------------
  bool isPresent(const std::string& key) const { return find(key) != nullptr; }

  // keys of a given kind in insertion order ("dict" -> sub-dictionaries, "all")
  std::vector<std::string> keys(const std::string& type = "all") const {
    std::vector<std::string> out;
    for (auto& e : entries_) {
      if (type == "all" || (type == "dict" && e.kind == DICT)) out.push_back(e.key);
    }
    return out;
  }

  const Dict& getDict(const std::string& key) const {
    const Entry* e = need(key);
    if (e->kind != DICT) throw FatalError("dictionary%get", "entry '" + key + "' is not a dictionary");
    return *e->sub;
  }

  int getInt(const std::string& key) const {
    const Entry* e = need(key);
    if (e->kind != SCALAR || !isInt(e->tok[0]))
      throw FatalError("dictionary%get", "entry '" + key + "' is not an integer");
    return toInt(e->tok[0]);
  }
  int getInt(const std::string& key, int def) const { return isPresent(key) ? getInt(key) : def; }

  double getReal(const std::string& key) const {
    const Entry* e = need(key);
    if (e->kind != SCALAR || !(isInt(e->tok[0]) || isReal(e->tok[0])))
      throw FatalError("dictionary%get", "entry '" + key + "' is not a real");
    return toReal(e->tok[0]);
  }
  double getReal(const std::string& key, double def) const { return isPresent(key) ? getReal(key) : def; }

  // logicals are stored as integers 0/1 in SCONE decks
  bool getBool(const std::string& key) const { return getInt(key) != 0; }
  bool getBool(const std::string& key, bool def) const { return isPresent(key) ? getBool(key) : def; }

  std::string getWord(const std::string& key) const {
    const Entry* e = need(key);
    if (e->kind != SCALAR) throw FatalError("dictionary%get", "entry '" + key + "' is not a word");
    return e->tok[0];
  }
  std::string getWord(const std::string& key, const std::string& def) const {
    return isPresent(key) ? getWord(key) : def;
  }

  std::vector<int> getIntArray(const std::string& key) const {
    const Entry* e = need(key);
    std::vector<int> out;
    for (auto& t : listTokens(e, key)) {
      if (!isInt(t)) throw FatalError("dictionary%get", "entry '" + key + "' is not an integer array");
      out.push_back(toInt(t));
    }
    return out;
  }

  std::vector<double> getRealArray(const std::string& key) const {
    const Entry* e = need(key);
    std::vector<double> out;
    for (auto& t : listTokens(e, key)) {
      if (!(isInt(t) || isReal(t)))
        throw FatalError("dictionary%get", "entry '" + key + "' is not a real array");
      out.push_back(toReal(t));
    }
    return out;
  }

  std::vector<std::string> getWordArray(const std::string& key) const {
    const Entry* e = need(key);
    return listTokens(e, key);
  }

  // ---- programmatic build (used by hosts that synthesise dictionaries) ------
  void store(const std::string& key, const std::string& scalarToken) {
    Entry e; e.key = key; e.kind = SCALAR; e.tok = {scalarToken};
    add(std::move(e));
  }
  void storeList(const std::string& key, std::vector<std::string> toks) {
    Entry e; e.key = key; e.kind = LIST; e.tok = std::move(toks);
    add(std::move(e));
  }
  void storeDict(const std::string& key, const Dict& d) {
    Entry e; e.key = key; e.kind = DICT; e.sub = std::make_shared<Dict>(d);
    add(std::move(e));
  }
  // replace-or-insert of a scalar (deck overrides: pop, active, ...)
  void setScalar(const std::string& key, const std::string& tok) {
    for (auto& e : entries_) if (e.key == key) { e.kind = SCALAR; e.tok = {tok}; e.sub.reset(); return; }
    store(key, tok);
  }
  void setDict(const std::string& key, const Dict& d) {
    for (auto& e : entries_) if (e.key == key) { e.kind = DICT; e.tok.clear(); e.sub = std::make_shared<Dict>(d); return; }
    storeDict(key, d);
  }

  static bool isInt(const std::string& t) {
    if (t.empty()) return false;
    size_t i = (t[0] == '-' || t[0] == '+') ? 1 : 0;
    if (i >= t.size()) return false;
    for (; i < t.size(); ++i) if (t[i] < '0' || t[i] > '9') return false;
    // only 32-bit integers are integers; longer ones are read as words
    char* end = nullptr;
    long long v = std::strtoll(t.c_str(), &end, 10);
    return v >= INT32_MIN && v <= INT32_MAX;
  }
  // What a Fortran formatted READ with an ES edit descriptor accepts (dictParser_func.f90:781, tried after the I20 read failed):
  // [sign] digits [. digits] | [sign] . digits, then optionally an exponent: a letter e E d D with an optional sign, or a bare sign,
  // followed by digits ("1E-11", "1.0d3", "2.5+3").  Integers that do not fit 32 bits fail the integer read and are reals too.
  static bool isReal(const std::string& t) {
    size_t i = 0, n = t.size();
    if (i < n && (t[i] == '-' || t[i] == '+')) ++i;
    size_t nd = 0; bool dot = false;
    for (; i < n; ++i) {
      if (t[i] >= '0' && t[i] <= '9') ++nd;
      else if (t[i] == '.' && !dot) dot = true;
      else break;
    }
    if (nd == 0) return false;
    if (i == n) return true;
    if (t[i] == 'e' || t[i] == 'E' || t[i] == 'd' || t[i] == 'D') { ++i; if (i < n && (t[i] == '-' || t[i] == '+')) ++i; }
    else if (t[i] == '-' || t[i] == '+') ++i;
    else return false;
    if (i >= n) return false;
    for (; i < n; ++i) if (t[i] < '0' || t[i] > '9') return false;
    return true;
  }
  static int toInt(const std::string& t) { return (int)std::strtol(t.c_str(), nullptr, 10); }
  static double toReal(const std::string& t) {
    std::string s = t;
    for (auto& c : s) if (c == 'd' || c == 'D') c = 'e';
    for (size_t i = 1; i < s.size(); ++i)                                   // exponent given by a bare sign: 2.5+3
      if ((s[i] == '+' || s[i] == '-') && ((s[i - 1] >= '0' && s[i - 1] <= '9') || s[i - 1] == '.')) { s.insert(i, "e"); break; }
    return std::strtod(s.c_str(), nullptr);   // correctly rounded, as the formatted READ
  }

 private:
  std::vector<Entry> entries_;

  void add(Entry e) {
    if (find(e.key)) throw FatalError("dictionary%store", "keyword '" + e.key + "' is already present");
    entries_.push_back(std::move(e));
  }
  const Entry* find(const std::string& key) const {
    for (auto& e : entries_) if (e.key == key) return &e;
    return nullptr;
  }
  const Entry* need(const std::string& key) const {
    const Entry* e = find(key);
    if (!e) throw FatalError("dictionary%get", "keyword '" + key + "' is not present");
    return e;
  }
  static std::vector<std::string> listTokens(const Entry* e, const std::string& key) {
    if (e->kind == DICT) throw FatalError("dictionary%get", "entry '" + key + "' is a dictionary");
    return e->tok;   // a scalar is accepted as a size-1 array
  }

  static std::string stripComments(const std::string& text) {
    std::string s;
    s.reserve(text.size());
    size_t i = 0, n = text.size();
    while (i < n) {
      char c = text[i];
      if (c == '!' || (c == '/' && i + 1 < n && text[i + 1] == '/')) {
        while (i < n && text[i] != '\n') ++i;
        s.push_back(' ');
        continue;
      }
      if (c == '\t' || c == '\n' || c == '\r') c = ' ';
      s.push_back(c);
      ++i;
    }
    return s;
  }

  static void skipSpace(const std::string& s, size_t& pos) {
    while (pos < s.size() && s[pos] == ' ') ++pos;
  }

  static std::vector<std::string> split(const std::string& s) {
    std::vector<std::string> out;
    std::istringstream is(s);
    std::string t;
    while (is >> t) out.push_back(t);
    return out;
  }

  // parse items until the matching '}' ; leaves pos one past it
  static void parse(Dict& d, const std::string& s, size_t& pos) {
    for (;;) {
      skipSpace(s, pos);
      if (pos >= s.size()) throw FatalError("parseDict", "missing '}'");
      if (s[pos] == '}') { ++pos; return; }
      // keyword
      size_t b = pos;
      while (pos < s.size() && s[pos] != ' ' && s[pos] != '{' && s[pos] != ';' && s[pos] != '}') ++pos;
      std::string key = s.substr(b, pos - b);
      if (key.empty()) throw FatalError("parseDict", "empty keyword near: " + s.substr(b, 30));
      skipSpace(s, pos);
      if (pos >= s.size()) throw FatalError("parseDict", "unexpected end after keyword " + key);
      if (s[pos] == '{') {
        ++pos;
        Entry e; e.key = key; e.kind = DICT; e.sub = std::make_shared<Dict>();
        parse(*e.sub, s, pos);
        d.add(std::move(e));
        continue;
      }
      // content up to ';'
      size_t semi = s.find(';', pos);
      if (semi == std::string::npos) throw FatalError("parseDict", "missing ';' after keyword " + key);
      std::string content = s.substr(pos, semi - pos);
      pos = semi + 1;
      Entry e; e.key = key;
      size_t p0 = content.find_first_not_of(' ');
      if (p0 == std::string::npos) throw FatalError("parseDict", "empty content for keyword " + key);
      if (content[p0] == '(') {
        size_t p1 = content.rfind(')');
        if (p1 == std::string::npos) throw FatalError("parseDict", "missing ')' for keyword " + key);
        e.kind = LIST;
        e.tok = split(content.substr(p0 + 1, p1 - p0 - 1));
      } else if (content[p0] == '[') {
        size_t p1 = content.rfind(']');
        if (p1 == std::string::npos) throw FatalError("parseDict", "missing ']' for keyword " + key);
        e.kind = TOKENS;
        e.tok = split(content.substr(p0 + 1, p1 - p0 - 1));
      } else {
        e.kind = SCALAR;
        e.tok = split(content);
        if (e.tok.size() != 1) throw FatalError("parseDict", "scalar entry '" + key + "' has spaces in content");
      }
      d.add(std::move(e));
    }
  }
};

}  // namespace sb
