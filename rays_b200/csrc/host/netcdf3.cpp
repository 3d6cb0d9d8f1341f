// netcdf3.cpp — see netcdf3.hpp.  Format: "The NetCDF Classic Format Specification" (CDF-1/2/5).
#include "netcdf3.hpp"

#include <algorithm>
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
    if (sz < 0) { std::fclose(f); err_ = "cannot size '" + path + "'"; return false; }
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
        for (int d : v.dimids) if (d < 0 || (size_t)d >= dim_lens_.size()) { err_ = "corrupt netCDF header in '" + path + "' (dimension id out of range)"; return false; }
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
    size_t slab = 1;      // elements per record (record variable) or of the whole variable
    for (size_t k = v.is_record ? 1 : 0; k < v.dimids.size(); ++k) slab *= (size_t)dim_lens_[v.dimids[k]];
    const size_t nrec = v.is_record ? (size_t)numrecs_ : 1;
    uint64_t recsize = 0;   // distance between two records: the vsize of every record variable
    if (v.is_record) for (const NcVarInfo &q : vars_) if (q.is_record) recsize += q.vsize;
    const size_t cnt = slab * nrec;
    int ts = type_size(v.type);
    if (ts == 0 || (nrec && v.begin + (nrec - 1) * recsize + (uint64_t)slab * ts > buf_.size())) { err_ = "netCDF variable '" + n + "' runs past end of file"; return false; }
    out.resize(cnt);
    for (size_t i = 0; i < cnt; ++i) {
        const uint8_t *p = &buf_[v.begin + (i / slab) * recsize + (i % slab) * ts];
        if (v.type == NC_DOUBLE) { uint64_t u = 0; for (int k = 0; k < 8; ++k) u = (u << 8) | p[k]; double d; std::memcpy(&d, &u, 8); out[i] = d; }
        else if (v.type == NC_FLOAT) { uint32_t u = 0; for (int k = 0; k < 4; ++k) u = (u << 8) | p[k]; float d; std::memcpy(&d, &u, 4); out[i] = d; }
        else if (v.type == NC_INT) { uint32_t u = 0; for (int k = 0; k < 4; ++k) u = (u << 8) | p[k]; out[i] = (double)(int32_t)u; }
        else { err_ = "netCDF variable '" + n + "' has a non-numeric type"; return false; }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
int NcWriter::def_dim(const std::string &name, int64_t len) { dims_.push_back({name, len, false}); return (int)dims_.size() - 1; }
int NcWriter::def_unlimited_dim(const std::string &name) { dims_.push_back({name, 0, true}); return (int)dims_.size() - 1; }
int NcWriter::def_var(const std::string &name, int type, const std::vector<int> &dimids) {
    Var v; v.name = name; v.type = type; v.dimids = dimids;
    vars_.push_back(v);
    return (int)vars_.size() - 1;
}
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
void NcWriter::put_double(int varid, const double *d, size_t n) { Var &v = vars_[varid]; v.data.clear(); v.dsrc = d; v.isrc = nullptr; v.nsrc = n; }
void NcWriter::put_int(int varid, const int32_t *d, size_t n) { Var &v = vars_[varid]; v.data.clear(); v.isrc = d; v.dsrc = nullptr; v.nsrc = n; }
void NcWriter::put_char(int varid, const char *d, size_t n) { Var &v = vars_[varid]; v.data.assign(d, d + n); v.dsrc = nullptr; v.isrc = nullptr; v.nsrc = 0; }

bool NcWriter::close(const std::string &path, std::string &err) {
    // Layout (netCDF classic): header, the fixed-size variables in definition order, then numrecs records, each holding
    // one slab of every record variable in definition order.
    auto is_rec = [&](const Var &v) { return !v.dimids.empty() && dims_[v.dimids[0]].unlimited; };
    auto slab_count = [&](const Var &v) { uint64_t cnt = 1; for (size_t k = is_rec(v) ? 1 : 0; k < v.dimids.size(); ++k) cnt *= (uint64_t)dims_[v.dimids[k]].len; return cnt; };
    size_t n_rec_vars = 0, n_unlimited = 0;
    for (auto &d : dims_) {
        n_unlimited += d.unlimited ? 1 : 0;
        if (!d.unlimited && d.len < 1) { err = "netCDF dimension '" + d.name + "' has length " + std::to_string(d.len) + " (fixed dimensions need length >= 1)"; return false; }
    }
    if (n_unlimited > 1) { err = "more than one unlimited dimension"; return false; }
    for (auto &v : vars_) {
        n_rec_vars += is_rec(v) ? 1 : 0;
        for (size_t k = 1; k < v.dimids.size(); ++k) if (dims_[v.dimids[k]].unlimited) { err = "variable '" + v.name + "': the unlimited dimension must come first"; return false; }
    }
    // CDF-2 holds 32-bit counts and vsize: go to CDF-5 as soon as one of them does not fit
    bool big = force5_ || numrecs_ > 0xFFFFFFFFll;
    for (auto &d : dims_) big = big || d.len > 0xFFFFFFFFll;
    for (auto &v : vars_) big = big || slab_count(v) * (uint64_t)type_size(v.type) + 3 > 0xFFFFFFFFull;
    const int version = big ? 5 : 2;
    auto cnt = [&](std::vector<uint8_t> &o, uint64_t x) { if (version == 5) be64(o, x); else be32(o, (uint32_t)x); };   // NON_NEG
    auto put_name = [&](std::vector<uint8_t> &o, const std::string &s) { cnt(o, s.size()); o.insert(o.end(), s.begin(), s.end()); while (o.size() % 4) o.push_back(0); };
    auto header = [&](const std::vector<uint64_t> &begins, std::vector<uint64_t> &vsizes) {
        std::vector<uint8_t> h = {'C', 'D', 'F', (uint8_t)version};
        cnt(h, (uint64_t)numrecs_);
        if (dims_.empty()) { be32(h, 0); cnt(h, 0); } else { be32(h, NC_DIMENSION); cnt(h, dims_.size()); for (auto &d : dims_) { put_name(h, d.name); cnt(h, d.unlimited ? 0 : (uint64_t)d.len); } }
        if (gatts_.empty()) { be32(h, 0); cnt(h, 0); } else {
            be32(h, NC_ATTRIBUTE); cnt(h, gatts_.size());
            for (auto &a : gatts_) { put_name(h, a.name); be32(h, (uint32_t)a.type); cnt(h, (uint64_t)a.nelems); h.insert(h.end(), a.data.begin(), a.data.end()); while (h.size() % 4) h.push_back(0); }
        }
        if (vars_.empty()) { be32(h, 0); cnt(h, 0); } else {
            be32(h, NC_VARIABLE); cnt(h, vars_.size());
            for (size_t i = 0; i < vars_.size(); ++i) {
                auto &v = vars_[i];
                put_name(h, v.name); cnt(h, v.dimids.size());
                for (int d : v.dimids) cnt(h, (uint64_t)d);
                be32(h, 0); cnt(h, 0);  // no variable attributes
                be32(h, (uint32_t)v.type);
                uint64_t vs = slab_count(v) * type_size(v.type);
                if (!(is_rec(v) && n_rec_vars == 1)) vs = (vs + 3) & ~(uint64_t)3;   // a lone record variable is not padded
                vsizes[i] = vs;
                cnt(h, vs);
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
    bool ok = std::fwrite(h.data(), 1, h.size(), f) == h.size();
    // `need` bytes of variable v starting at element `first`: borrowed arrays are byte-swapped block by block
    std::vector<uint8_t> blk;
    auto write_elems = [&](const Var &v, uint64_t first_byte, uint64_t need) {
        if (!v.dsrc && !v.isrc) {   // owned bytes (char data); short data is zero-padded
            const uint64_t have = v.data.size() > first_byte ? std::min<uint64_t>(v.data.size() - first_byte, need) : 0;
            if (have && std::fwrite(v.data.data() + first_byte, 1, have, f) != have) ok = false;
            for (uint64_t p = have; p < need && ok; ++p) if (std::fputc(0, f) == EOF) ok = false;
            return;
        }
        const int ts = type_size(v.type);
        uint64_t e0 = first_byte / ts, ne = need / ts;
        const uint64_t chunk = 1u << 20;
        while (ne && ok) {
            const uint64_t m = std::min(ne, chunk);
            blk.clear();
            blk.reserve(m * ts);
            for (uint64_t i = 0; i < m; ++i) {
                const uint64_t e = e0 + i;
                if (v.dsrc) {
                    const double x = e < v.nsrc ? v.dsrc[e] : 0.0;
                    if (v.type == NC_DOUBLE) { uint64_t u; std::memcpy(&u, &x, 8); be64(blk, u); }
                    else { float fl = (float)x; uint32_t u; std::memcpy(&u, &fl, 4); be32(blk, u); }
                } else be32(blk, (uint32_t)(e < v.nsrc ? v.isrc[e] : 0));
            }
            if (std::fwrite(blk.data(), 1, blk.size(), f) != blk.size()) ok = false;
            e0 += m; ne -= m;
        }
    };
    for (size_t i = 0; i < vars_.size() && ok; ++i) {
        auto &v = vars_[i];
        if (is_rec(v)) continue;
        const uint64_t need = slab_count(v) * type_size(v.type);
        write_elems(v, 0, need);
        for (uint64_t p = need; p < vsizes[i] && ok; ++p) if (std::fputc(0, f) == EOF) ok = false;
    }
    for (int64_t r = 0; r < numrecs_ && ok; ++r)
        for (size_t i = 0; i < vars_.size() && ok; ++i) {
            auto &v = vars_[i];
            if (!is_rec(v)) continue;
            const uint64_t slab = slab_count(v) * type_size(v.type);
            write_elems(v, slab * (uint64_t)r, slab);
            for (uint64_t p = slab; p < vsizes[i] && ok; ++p) if (std::fputc(0, f) == EOF) ok = false;
        }
    if (std::fflush(f) != 0) ok = false;
    if (std::fclose(f) != 0) ok = false;
    if (!ok) { err = "write error on '" + path + "' (disk full?)"; std::remove(path.c_str()); return false; }
    return true;
}

}  // namespace rays_host
