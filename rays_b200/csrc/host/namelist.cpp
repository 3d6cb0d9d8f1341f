// namelist.cpp — see namelist.hpp
#include "namelist.hpp"

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace rays_host {

std::string NamelistGroup::lower(const std::string &s) {
    std::string r = s;
    for (auto &ch : r) ch = (char)std::tolower((unsigned char)ch);
    return r;
}

bool NamelistFile::load(const std::string &path) {
    std::ifstream in(path);
    if (!in) { err_ = "cannot open namelist file '" + path + "'"; return false; }
    std::stringstream ss;
    ss << in.rdbuf();
    return load_text(ss.str());
}

namespace {
struct Tok { enum Kind { NAME, VALUE, STR, EQ, LP, RP, STAR, END } kind; std::string text; };

// tokenise the body of one group (between `&name` and `/`)
bool tokenize(const std::string &s, std::vector<Tok> &out, std::string &err) {
    size_t i = 0, n = s.size();
    while (i < n) {
        char ch = s[i];
        if (ch == '!') { while (i < n && s[i] != '\n') ++i; continue; }
        if (std::isspace((unsigned char)ch) || ch == ',') { ++i; continue; }
        if (ch == '\'' || ch == '"') {
            char q = ch; ++i; std::string v;
            for (;;) {
                if (i >= n) { err = "unterminated string in namelist"; return false; }
                if (s[i] == q) { if (i + 1 < n && s[i + 1] == q) { v += q; i += 2; continue; } ++i; break; }
                v += s[i++];
            }
            out.push_back({Tok::STR, v});
            continue;
        }
        if (ch == '=') { out.push_back({Tok::EQ, "="}); ++i; continue; }
        if (ch == '(') { out.push_back({Tok::LP, "("}); ++i; continue; }
        if (ch == ')') { out.push_back({Tok::RP, ")"}); ++i; continue; }
        if (ch == '*') { out.push_back({Tok::STAR, "*"}); ++i; continue; }
        // bare word: name, number or logical
        size_t j = i;
        while (j < n && !std::isspace((unsigned char)s[j]) && s[j] != ',' && s[j] != '=' && s[j] != '(' && s[j] != ')' &&
               s[j] != '*' && s[j] != '!' && s[j] != '\'' && s[j] != '"')
            ++j;
        out.push_back({Tok::VALUE, s.substr(i, j - i)});
        i = j;
    }
    out.push_back({Tok::END, ""});
    return true;
}
}  // namespace

bool NamelistFile::load_text(const std::string &text) {
    groups_.clear();
    size_t i = 0, n = text.size();
    while (i < n) {
        // find '&' at start of a token outside comments
        if (text[i] == '!') { while (i < n && text[i] != '\n') ++i; continue; }
        if (text[i] != '&') { ++i; continue; }
        size_t j = i + 1;
        while (j < n && (std::isalnum((unsigned char)text[j]) || text[j] == '_')) ++j;
        std::string gname = NamelistGroup::lower(text.substr(i + 1, j - i - 1));
        // body ends at a '/' that is outside quotes and comments
        size_t k = j;
        bool inq = false; char q = 0;
        for (; k < n; ++k) {
            char ch = text[k];
            if (inq) { if (ch == q) inq = false; continue; }
            if (ch == '\'' || ch == '"') { inq = true; q = ch; continue; }
            if (ch == '!') { while (k < n && text[k] != '\n') ++k; continue; }
            if (ch == '/') break;
        }
        if (k >= n) { err_ = "namelist group &" + gname + " not terminated by '/'"; return false; }
        std::vector<Tok> toks;
        if (!tokenize(text.substr(j, k - j), toks, err_)) return false;
        std::vector<NmlAssign> assigns;
        size_t t = 0;
        while (toks[t].kind != Tok::END) {
            // assignment:  NAME [ ( int ) ] = values...
            if (toks[t].kind != Tok::VALUE) { err_ = "namelist &" + gname + ": expected a variable name near '" + toks[t].text + "'"; return false; }
            NmlAssign a;
            a.name = NamelistGroup::lower(toks[t].text);
            ++t;
            if (toks[t].kind == Tok::LP) {
                if (toks[t + 1].kind != Tok::VALUE || toks[t + 2].kind != Tok::RP) { err_ = "namelist &" + gname + ": bad subscript on " + a.name; return false; }
                a.has_index = true;
                a.index = std::atoi(toks[t + 1].text.c_str());
                t += 3;
            }
            if (toks[t].kind != Tok::EQ) { err_ = "namelist &" + gname + ": expected '=' after " + a.name; return false; }
            ++t;
            // values until the next `NAME =` or `NAME (`...`) =` or END
            for (;;) {
                if (toks[t].kind == Tok::END) break;
                if (toks[t].kind == Tok::VALUE) {
                    // look ahead: is this the next variable name?
                    if (toks[t + 1].kind == Tok::EQ) break;
                    if (toks[t + 1].kind == Tok::LP && toks[t + 2].kind == Tok::VALUE && toks[t + 3].kind == Tok::RP && toks[t + 4].kind == Tok::EQ) break;
                }
                int rep = 1;
                if (toks[t].kind == Tok::VALUE && toks[t + 1].kind == Tok::STAR) {
                    rep = std::atoi(toks[t].text.c_str());
                    t += 2;
                }
                if (toks[t].kind != Tok::VALUE && toks[t].kind != Tok::STR) { err_ = "namelist &" + gname + ": bad value for " + a.name; return false; }
                for (int r = 0; r < rep; ++r) { a.values.push_back(toks[t].text); a.quoted.push_back(toks[t].kind == Tok::STR); }
                ++t;
            }
            assigns.push_back(a);
        }
        groups_[gname] = assigns;  // a later group of the same name replaces the earlier (first match wins in Fortran; inputs have one)
        i = k + 1;
    }
    return true;
}

bool NamelistFile::has_group(const std::string &g) const { return groups_.count(NamelistGroup::lower(g)) > 0; }
const std::vector<NmlAssign> *NamelistFile::group(const std::string &g) const {
    auto it = groups_.find(NamelistGroup::lower(g));
    return it == groups_.end() ? nullptr : &it->second;
}

static bool parse_real(const std::string &s, double &v) {
    std::string t = s;
    for (auto &ch : t) if (ch == 'd' || ch == 'D') ch = 'e';
    char *end = nullptr;
    v = std::strtod(t.c_str(), &end);
    return end && *end == '\0' && end != t.c_str();
}
static bool parse_logical(const std::string &s, bool &v) {
    std::string t = NamelistGroup::lower(s);
    if (t == ".true." || t == "t" || t == ".t." || t == "true") { v = true; return true; }
    if (t == ".false." || t == "f" || t == ".f." || t == "false") { v = false; return true; }
    return false;
}

bool NamelistGroup::read(const NamelistFile &f, std::string &err) const {
    const std::vector<NmlAssign> *g = f.group(name_);
    if (!g) { err = "namelist group &" + name_ + " not found"; return false; }
    for (const NmlAssign &a : *g) {
        auto it = slots_.find(a.name);
        if (it == slots_.end()) { err = "namelist &" + name_ + ": unknown variable '" + a.name + "'"; return false; }
        const Slot &s = it->second;
        int pos = a.has_index ? a.index - s.lb : 0;
        for (size_t iv = 0; iv < a.values.size(); ++iv, ++pos) {
            if (pos < 0 || pos >= s.n) { err = "namelist &" + name_ + ": too many values / bad subscript for '" + a.name + "'"; return false; }
            const std::string &val = a.values[iv];
            switch (s.kind) {
                case 'd': { double v; if (!parse_real(val, v)) { err = "namelist &" + name_ + ": bad real '" + val + "' for " + a.name; return false; } s.pd[pos] = v; } break;
                case 'i': { double v; if (!parse_real(val, v)) { err = "namelist &" + name_ + ": bad integer '" + val + "' for " + a.name; return false; } s.pi[pos] = (int)v; } break;
                case 'b': { bool v; if (!parse_logical(val, v)) { err = "namelist &" + name_ + ": bad logical '" + val + "' for " + a.name; return false; } s.pb[pos] = v; } break;
                case 's': s.ps[pos] = val; break;
            }
        }
    }
    return true;
}

}  // namespace rays_host
