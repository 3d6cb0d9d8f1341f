// netcdf3.cpp — see netcdf3.hpp.  Format: "The NetCDF Classic Format Specification" (CDF-1/2/5).
#include "netcdf3.hpp"

#include <cstdio>
#include <cstring>

namespace rays_host {

namespace {
const uint32_t NC_DIMENSION = 10, NC_VARIABLE = 11, NC_ATTRIBUTE = 12;
int type_size(int t) {
    switch (t) { case NC_BYTE: case NC_CHAR: return 1; case NC_SHORT: return 2; case NC_INT: case NC_FLOAT: return 4; case NC_DOUBLE: return 8; }
    return 0;
}
struct Cursor {
    const std::vector<uint8_t> &b; size_t p = 0; bool ok = true; int version;
    Cursor(const std::vector<uint8_t> &bb, int v) : b(bb), version(v) {}
    uint32_t u32() { if (p + 4 > b.size()) { ok = false; return 0; } uint32_t v = ((uint32_t)b[p] << 24) | ((uint32_t)b[p + 1] << 16) | ((uint32_t)b[p + 2] << 8) | b[p + 3]; p += 4; return v; }
    uint64_t u64() { uint64_t hi = u32(), lo = u32(); return (hi << 32) | lo; }
    uint64_t count() { return version == 5 ? u64() : u32(); }     // NON_NEG: 64-bit in CDF-5
    std::string name() { uint64_t n = count(); if (p + n > b.size()) { ok = false; return ""; } std::string s((const char *)&b[p], n); p += (n + 3) & ~(uint64_t)3; return s; }
};
}  // namespace

bool NcReader::open(const std::string &path) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) { err_ = "cannot open '" + path + "'"; return false; }
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    buf_.resize((size_t)sz);
    size_t got = std::fread(buf_.data(), 1, (size_t)sz, f);
    std::fclose(f);
    if (got != (size_t)sz || sz < 8 || std::memcmp(buf_.data(), "CDF", 3) != 0) { err_ = "'" + path + "' is not a netCDF classic file"; return false; }
    version_ = buf_[3];
    if (version_ != 1 && version_ != 2 && version_ != 5) { err_ = "unsupported netCDF version byte"; return false; }
    Cursor c(buf_, version_);
    c.p = 4;
    numrecs_ = (int64_t)c.count();
    // dim_list
    uint32_t tag = c.u32(); uint64_t n = c.count();
    if (tag == NC_DIMENSION) for (uint64_t i = 0; i < n; ++i) { std::string nm = c.name(); int64_t len = (int64_t)c.count(); dim_index_[nm] = (int)dim_names_.size(); dim_names_.push_back(nm); dim_lens_.push_back(len); }
    auto read_atts = [&](std::map<std::string, std::string> *text) {
        uint32_t t = c.u32(); uint64_t na = c.count();
        if (t != NC_ATTRIBUTE) return;
        for (uint64_t i = 0; i < na; ++i) {
            std::string nm = c.name(); uint32_t ty = c.u32(); uint64_t ne = c.count();
            size_t bytes = (size_t)ne * type_size((int)ty);
            if (text && ty == NC_CHAR && c.p + bytes <= buf_.size()) (*text)[nm] = std::string((const char *)&buf_[c.p], bytes);
            c.p += (bytes + 3) & ~(size_t)3;
        }
    };
    read_atts(&gatt_text_);
    tag = c.u32(); n = c.count();
    if (tag == NC_VARIABLE) for (uint64_t i = 0; i < n; ++i) {
        NcVarInfo v; v.name = c.name();
        uint64_t nd = c.count();
        for (uint64_t d = 0; d < nd; ++d) v.dimids.push_back((int)c.count());
        read_atts(nullptr);
        v.type = (int)c.u32();
        v.vsize = c.count();
        v.begin = (version_ == 1) ? c.u32() : c.u64();
        v.is_record = !v.dimids.empty() && dim_lens_[v.dimids[0]] == 0;
        var_index_[v.name] = (int)vars_.size();
        vars_.push_back(v);
    }
    if (!c.ok) { err_ = "truncated netCDF header in '" + path + "'"; return false; }
    return true;
}

int64_t NcReader::dim_len(const std::string &n) const {
    auto it = dim_index_.find(n);
    if (it == dim_index_.end()) return -1;
    int64_t l = dim_lens_[it->second];
    return l == 0 ? numrecs_ : l;
}
std::string NcReader::get_att_text(const std::string &n) const {
    auto it = gatt_text_.find(n);
    return it == gatt_text_.end() ? std::string() : it->second;
}

bool NcReader::get_var_double(const std::string &n, std::vector<double> &out) {
    auto it = var_index_.find(n);
    if (it == var_index_.end()) { err_ = "netCDF variable '" + n + "' not found"; return false; }
    const NcVarInfo &v = vars_[it->second];
    if (v.is_record) { err_ = "record variable '" + n + "' not supported by this reader"; return false; }
    size_t cnt = 1;
    for (int d : v.dimids) cnt *= (size_t)dim_lens_[d];
    int ts = type_size(v.type);
    if (v.begin + cnt * ts > buf_.size()) { err_ = "netCDF variable '" + n + "' runs past end of file"; return false; }
    out.resize(cnt);
    const uint8_t *p = &buf_[v.begin];
    for (size_t i = 0; i < cnt; ++i, p += ts) {
        if (v.type == NC_DOUBLE) { uint64_t u = 0; for (int k = 0; k < 8; ++k) u = (u << 8) | p[k]; double d; std::memcpy(&d, &u, 8); out[i] = d; }
        else if (v.type == NC_FLOAT) { uint32_t u = 0; for (int k = 0; k < 4; ++k) u = (u << 8) | p[k]; float d; std::memcpy(&d, &u, 4); out[i] = d; }
        else if (v.type == NC_INT) { uint32_t u = 0; for (int k = 0; k < 4; ++k) u = (u << 8) | p[k]; out[i] = (double)(int32_t)u; }
        else { err_ = "netCDF variable '" + n + "' has a non-numeric type"; return false; }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
int NcWriter::def_dim(const std::string &name, int64_t len) { dims_.push_back({name, len}); return (int)dims_.size() - 1; }
int NcWriter::def_var(const std::string &name, int type, const std::vector<int> &dimids) { vars_.push_back({name, type, dimids, {}}); return (int)vars_.size() - 1; }
void NcWriter::put_att_text(const std::string &name, const std::string &value) {
    Att a; a.name = name; a.type = NC_CHAR; a.data.assign(value.begin(), value.end()); a.nelems = (int64_t)value.size(); gatts_.push_back(a);
}
static void be32(std::vector<uint8_t> &o, uint32_t v) { o.push_back(v >> 24); o.push_back(v >> 16); o.push_back(v >> 8); o.push_back(v); }
static void be64(std::vector<uint8_t> &o, uint64_t v) { be32(o, (uint32_t)(v >> 32)); be32(o, (uint32_t)v); }
void NcWriter::put_att_int(const std::string &name, const std::vector<int32_t> &v) {
    Att a; a.name = name; a.type = NC_INT; a.nelems = (int64_t)v.size();
    for (int32_t x : v) be32(a.data, (uint32_t)x);
    gatts_.push_back(a);
}
void NcWriter::put_double(int varid, const double *d, size_t n) {
    Var &v = vars_[varid];
    v.data.clear();
    if (v.type == NC_DOUBLE) { v.data.reserve(n * 8); for (size_t i = 0; i < n; ++i) { uint64_t u; std::memcpy(&u, &d[i], 8); be64(v.data, u); } }
    else { v.data.reserve(n * 4); for (size_t i = 0; i < n; ++i) { float f = (float)d[i]; uint32_t u; std::memcpy(&u, &f, 4); be32(v.data, u); } }
}
void NcWriter::put_int(int varid, const int32_t *d, size_t n) { Var &v = vars_[varid]; v.data.clear(); for (size_t i = 0; i < n; ++i) be32(v.data, (uint32_t)d[i]); }
void NcWriter::put_char(int varid, const char *d, size_t n) { Var &v = vars_[varid]; v.data.assign(d, d + n); }

bool NcWriter::close(const std::string &path, std::string &err) {
    // CDF-2 (64-bit offsets).  Every variable here is < 4 GiB or the file is refused.
    // Layout (netCDF classic): header, the fixed-size variables in definition order, then numrecs records, each holding
    // one slab of every record variable in definition order.
    auto put_name = [](std::vector<uint8_t> &o, const std::string &s) { be32(o, (uint32_t)s.size()); o.insert(o.end(), s.begin(), s.end()); while (o.size() % 4) o.push_back(0); };
    auto is_rec = [&](const Var &v) { return !v.dimids.empty() && dims_[v.dimids[0]].len == 0; };
    auto slab_count = [&](const Var &v) { uint64_t cnt = 1; for (size_t k = is_rec(v) ? 1 : 0; k < v.dimids.size(); ++k) cnt *= (uint64_t)dims_[v.dimids[k]].len; return cnt; };
    size_t n_rec_vars = 0;
    for (auto &v : vars_) n_rec_vars += is_rec(v) ? 1 : 0;
    auto header = [&](const std::vector<uint64_t> &begins, std::vector<uint64_t> &vsizes) {
        std::vector<uint8_t> h = {'C', 'D', 'F', 2};
        be32(h, (uint32_t)numrecs_);
        if (dims_.empty()) { be32(h, 0); be32(h, 0); } else { be32(h, NC_DIMENSION); be32(h, (uint32_t)dims_.size()); for (auto &d : dims_) { put_name(h, d.name); be32(h, (uint32_t)d.len); } }
        if (gatts_.empty()) { be32(h, 0); be32(h, 0); } else {
            be32(h, NC_ATTRIBUTE); be32(h, (uint32_t)gatts_.size());
            for (auto &a : gatts_) { put_name(h, a.name); be32(h, (uint32_t)a.type); be32(h, (uint32_t)a.nelems); h.insert(h.end(), a.data.begin(), a.data.end()); while (h.size() % 4) h.push_back(0); }
        }
        if (vars_.empty()) { be32(h, 0); be32(h, 0); } else {
            be32(h, NC_VARIABLE); be32(h, (uint32_t)vars_.size());
            for (size_t i = 0; i < vars_.size(); ++i) {
                auto &v = vars_[i];
                put_name(h, v.name); be32(h, (uint32_t)v.dimids.size());
                for (int d : v.dimids) be32(h, (uint32_t)d);
                be32(h, 0); be32(h, 0);  // no variable attributes
                be32(h, (uint32_t)v.type);
                uint64_t vs = slab_count(v) * type_size(v.type);
                if (!(is_rec(v) && n_rec_vars == 1)) vs = (vs + 3) & ~(uint64_t)3;   // a lone record variable is not padded
                vsizes[i] = vs;
                be32(h, vs > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)vs);
                be64(h, begins.empty() ? 0 : begins[i]);
            }
        }
        return h;
    };
    std::vector<uint64_t> vsizes(vars_.size()), begins;
    std::vector<uint8_t> h0 = header(begins, vsizes);
    begins.resize(vars_.size());
    uint64_t off = h0.size();
    for (size_t i = 0; i < vars_.size(); ++i) if (!is_rec(vars_[i])) { begins[i] = off; off += vsizes[i]; }
    for (size_t i = 0; i < vars_.size(); ++i) if (is_rec(vars_[i])) { begins[i] = off; off += vsizes[i]; }
    std::vector<uint8_t> h = header(begins, vsizes);
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) { err = "cannot create '" + path + "'"; return false; }
    std::fwrite(h.data(), 1, h.size(), f);
    for (size_t i = 0; i < vars_.size(); ++i) {
        auto &v = vars_[i];
        if (is_rec(v)) continue;
        const uint64_t need = slab_count(v) * type_size(v.type);
        if (v.data.size() < need) v.data.resize(need, 0);
        std::fwrite(v.data.data(), 1, need, f);
        for (uint64_t p = need; p < vsizes[i]; ++p) std::fputc(0, f);
    }
    for (int64_t r = 0; r < numrecs_; ++r)
        for (size_t i = 0; i < vars_.size(); ++i) {
            auto &v = vars_[i];
            if (!is_rec(v)) continue;
            const uint64_t slab = slab_count(v) * type_size(v.type);
            if (v.data.size() < slab * (uint64_t)numrecs_) v.data.resize(slab * (uint64_t)numrecs_, 0);
            std::fwrite(v.data.data() + slab * (uint64_t)r, 1, slab, f);
            for (uint64_t p = slab; p < vsizes[i]; ++p) std::fputc(0, f);
        }
    std::fclose(f);
    return true;
}

}  // namespace rays_host
