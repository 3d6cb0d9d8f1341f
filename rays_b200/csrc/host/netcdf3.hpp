// netcdf3.hpp — minimal reader/writer for the netCDF classic file format (CDF-1 / CDF-2 / CDF-5
// headers; fixed-size variables, and record variables in the writer and in get_var_double).  The reference links netCDF-Fortran (`use netcdf`:
// mirror_magnetics_lib/mirror_magnetics_m.f90:382,456, RAYS_lib/ray_results_m.f90:175); this image
// has no netCDF library, and the files involved are classic-format, so the format is implemented here.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace rays_host {

enum NcType { NC_BYTE = 1, NC_CHAR = 2, NC_SHORT = 3, NC_INT = 4, NC_FLOAT = 5, NC_DOUBLE = 6 };

struct NcVarInfo {
    std::string name;
    std::vector<int> dimids;
    int type = 0;
    uint64_t vsize = 0, begin = 0;
    bool is_record = false;
};

class NcReader {
  public:
    bool open(const std::string &path);
    const std::string &error() const { return err_; }
    bool has_dim(const std::string &n) const { return dim_index_.count(n) > 0; }
    int64_t dim_len(const std::string &n) const;
    bool has_var(const std::string &n) const { return var_index_.count(n) > 0; }
    // read a whole NC_DOUBLE / NC_FLOAT / NC_INT variable as doubles, in file (C) order
    bool get_var_double(const std::string &n, std::vector<double> &out);
    std::string get_att_text(const std::string &n) const;

  private:
    std::vector<uint8_t> buf_;
    std::string err_;
    std::vector<std::string> dim_names_;
    std::vector<int64_t> dim_lens_;
    std::map<std::string, int> dim_index_, var_index_;
    std::vector<NcVarInfo> vars_;
    std::map<std::string, std::string> gatt_text_;
    int64_t numrecs_ = 0;
    int version_ = 1;
};

// Writer: define dims/vars/atts, then put whole variables, then close().  def_unlimited_dim defines the record (NC_UNLIMITED)
// dimension; variables whose first dimension it is are record variables, written for set_numrecs(n) records.  Fixed dimensions
// have length >= 1 (the classic format has no zero-length fixed dimension): close() refuses anything else.
// Format: CDF-2 (64-bit offsets) while every variable is < 4 GiB and every count < 2^32; otherwise CDF-5 (64-bit counts), which
// netCDF-C >= 4.4 and therefore netCDF-Fortran read transparently (ray_vec of a 1M-ray run is 60 GB).  put_double / put_int
// BORROW the caller's array (it must stay alive until close()); the byte-swapped copy is made block by block while writing.
class NcWriter {
  public:
    int def_dim(const std::string &name, int64_t len);
    int def_unlimited_dim(const std::string &name);
    void set_numrecs(int64_t n) { numrecs_ = n; }
    void force_cdf5(bool on) { force5_ = on; }
    int def_var(const std::string &name, int type, const std::vector<int> &dimids);
    void put_att_text(const std::string &name, const std::string &value);
    void put_att_int(const std::string &name, const std::vector<int32_t> &v);
    // data in C order of the dims given to def_var
    void put_double(int varid, const double *d, size_t n);   // NC_DOUBLE or NC_FLOAT (converted)
    void put_int(int varid, const int32_t *d, size_t n);
    void put_char(int varid, const char *d, size_t n);
    bool close(const std::string &path, std::string &err);

  private:
    struct Dim { std::string name; int64_t len; bool unlimited; };
    struct Var { std::string name; int type; std::vector<int> dimids; std::vector<uint8_t> data; const double *dsrc = nullptr; const int32_t *isrc = nullptr; size_t nsrc = 0; };
    struct Att { std::string name; int type; std::vector<uint8_t> data; int64_t nelems; };
    std::vector<Dim> dims_;
    std::vector<Var> vars_;
    std::vector<Att> gatts_;
    int64_t numrecs_ = 0;
    bool force5_ = false;
};

}  // namespace rays_host
