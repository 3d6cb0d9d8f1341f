// namelist.hpp — reader for the Fortran namelist input files RAYS uses (`rays.in`, one group per
// module, re-opened per module: e.g. RAYS_lib/ode_m.f90:130-132).  Supports what the reference's
// inputs use: `&group ... /`, `name = v`, `name(i) = v`, value lists, repeat counts `2*'zero'`,
// quoted strings, .true./.false., d/e exponents, `!` comments, text after the closing `/`.
// Like a Fortran READ, an unknown variable name in a group is an error.
#pragma once
#include <map>
#include <string>
#include <vector>

namespace rays_host {

struct NmlAssign {
    std::string name;                // lower case
    bool has_index = false;
    int index = 0;                   // the (i) subscript if present
    std::vector<std::string> values; // repeat counts expanded; strings unquoted
    std::vector<bool> quoted;
};

class NamelistFile {
  public:
    // returns false and sets error() if the file cannot be read or parsed
    bool load(const std::string &path);
    bool load_text(const std::string &text);
    bool has_group(const std::string &group) const;
    const std::vector<NmlAssign> *group(const std::string &group) const;
    const std::string &error() const { return err_; }

  private:
    std::map<std::string, std::vector<NmlAssign>> groups_;
    std::string err_;
};

// Binds namelist variables of one group to C++ storage, then applies a NamelistFile group.
class NamelistGroup {
  public:
    explicit NamelistGroup(std::string name) : name_(std::move(name)) {}
    void add(const std::string &var, double *p) { add_arr(var, p, 0, 1); }
    void add(const std::string &var, int *p) { add_arr(var, p, 0, 1); }
    void add(const std::string &var, bool *p) { Slot s; s.kind = 'b'; s.pb = p; s.lb = 0; s.n = 1; slots_[lower(var)] = s; }
    void add(const std::string &var, std::string *p) { add_arr(var, p, 0, 1); }
    // arrays: lb = Fortran lower bound, n = element count (column-major flattening for 2-D)
    void add_arr(const std::string &var, double *p, int lb, int n) { Slot s; s.kind = 'd'; s.pd = p; s.lb = lb; s.n = n; slots_[lower(var)] = s; }
    void add_arr(const std::string &var, int *p, int lb, int n) { Slot s; s.kind = 'i'; s.pi = p; s.lb = lb; s.n = n; slots_[lower(var)] = s; }
    void add_arr(const std::string &var, std::string *p, int lb, int n) { Slot s; s.kind = 's'; s.ps = p; s.lb = lb; s.n = n; slots_[lower(var)] = s; }
    // apply; group missing from the file is an error (Fortran: end-of-file on READ)
    bool read(const NamelistFile &f, std::string &err) const;
    static std::string lower(const std::string &s);

  private:
    struct Slot {
        char kind = 'd';
        double *pd = nullptr; int *pi = nullptr; bool *pb = nullptr; std::string *ps = nullptr;
        int lb = 0, n = 1;
    };
    std::string name_;
    std::map<std::string, Slot> slots_;
};

}  // namespace rays_host
