// ORACLE -- test infrastructure only.  C entry points around the reference's OWN object code for the two translation units
// that compile without Eigen (see oracle/ref_build.mk): hrm::Interval (src/datastructure/Interval.cpp:10-100) and
// hrm::parse2DCsvFile (src/util/Parse2dCsvFile.cpp:3-43).  tests/test_ref_pin.py drives the restatement (oracle/hrmo_*.hpp),
// the product's host reader (hrm_b200/host/hrm_io.hpp) and this library with the same inputs.
#include "hrm/datastructure/Interval.h"
#include "hrm/util/Parse2dCsvFile.h"

#include <cstring>

namespace {
std::vector<hrm::Interval> make(const double* se, int n) {
    std::vector<hrm::Interval> v;
    for (int i = 0; i < n; ++i) v.emplace_back(se[2 * i], se[2 * i + 1]);
    return v;
}
int put(const std::vector<hrm::Interval>& v, double* out, int cap) {
    const int n = static_cast<int>(v.size());
    for (int i = 0; i < n && i < cap; ++i) out[2 * i] = v[i].s(), out[2 * i + 1] = v[i].e();
    return n;
}
}  // namespace

extern "C" {
// op: 0 unions(a), 1 intersects(a), 2 complements(a = outer, b = inner); intervals as (start, end) pairs; returns the count
int hrmref_interval_op(int op, const double* a, int na, const double* b, int nb, double* out, int cap) {
    if (op == 0) return put(hrm::Interval::unions(make(a, na)), out, cap);
    if (op == 1) return put(hrm::Interval::intersects(make(a, na)), out, cap);
    return put(hrm::Interval::complements(make(a, na), make(b, nb)), out, cap);
}
// rows flattened with `stride` values per row; row_len[r] = cells of row r; returns the number of rows
long hrmref_parse_csv(const char* filename, double* out, int stride, int* row_len, long cap_rows) {
    const auto data = hrm::parse2DCsvFile(filename);
    long r = 0;
    for (const auto& row : data) {
        if (r >= cap_rows) break;
        row_len[r] = static_cast<int>(row.size());
        for (size_t c = 0; c < row.size() && c < static_cast<size_t>(stride); ++c) out[r * stride + c] = row[c];
        ++r;
    }
    return static_cast<long>(data.size());
}
}
