// edge-file-writer.cpp -- see include/edge-file-writer.hpp.
#include "edge-file-writer.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <iomanip>

namespace {

// right-aligned unsigned in a field of `width` (std::setw semantics: never truncates)
inline char* put_uint(char* p, unsigned long long v, int width)
{
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (int k = n; k < width; ++k) *p++ = ' ';
    while (n) *p++ = tmp[--n];
    return p;
}

// std::fixed << std::setprecision(4) << std::setw(8) << (double)v for a float v.
// v * 10^4 is exact in double (24-bit significand times 5^4 * 2^4), so rounding it to
// nearest-even is the correctly rounded decimal, which is what glibc's printf (and
// therefore libstdc++'s num_put) produces.
inline char* put_fixed4(char* p, float v)
{
    const double a = std::fabs((double)v);
    if (!(a < 1.0e11)) {                       // huge, inf, nan: rare, defer to printf
        p += std::snprintf(p, 48, "%8.4f", (double)v);
        return p;
    }
    const unsigned long long scaled = (unsigned long long)std::llrint(a * 10000.0);
    const unsigned long long ip = scaled / 10000ull;
    unsigned fp = (unsigned)(scaled % 10000ull);
    char tmp[24];
    int n = 0;
    unsigned long long t = ip;
    do { tmp[n++] = (char)('0' + t % 10); t /= 10; } while (t);
    const int neg = std::signbit(v) ? 1 : 0;
    for (int k = n + neg + 5; k < 8; ++k) *p++ = ' ';
    if (neg) *p++ = '-';
    while (n) *p++ = tmp[--n];
    *p++ = '.';
    p[3] = (char)('0' + fp % 10); fp /= 10;
    p[2] = (char)('0' + fp % 10); fp /= 10;
    p[1] = (char)('0' + fp % 10); fp /= 10;
    p[0] = (char)('0' + fp);
    return p + 4;
}

}  // namespace

EdgeFileWriter::EdgeFileWriter(std::ostream& os, std::vector<unsigned> ids) : os_(os), ids_(std::move(ids)) {}

void EdgeFileWriter::writeHeader()
{
    os_ << std::fixed << std::setprecision(4);
    os_ << std::setw(6) << "#Id" << std::setw(8) << "AtmId" << std::setw(6) << "ResId" << std::setw(6) << "Id"
        << std::setw(8) << "AtmId" << std::setw(6) << "ResId" << std::setw(8) << "Energy" << std::setw(8)
        << "Length" << std::setw(8) << "Cos\n";
    os_.flush();
}

void EdgeFileWriter::formatRows(const wn_edge* edges, size_t n_edges, std::string& out) const
{
    const size_t start = out.size();
    out.resize(start + n_edges * 96);            // upper bound per row
    char* const base = &out[0] + start;
    char* p = base;
    for (size_t k = 0; k < n_edges; ++k) {
        const wn_edge& e = edges[k];
        const unsigned long long id1 = ids_.empty() ? (unsigned long long)e.i : ids_.at((size_t)e.i);
        const unsigned long long id2 = ids_.empty() ? (unsigned long long)e.j : ids_.at((size_t)e.j);
        p = put_uint(p, id1, 6);
        p = put_uint(p, id2, 6);
        p = put_fixed4(p, e.energy);
        p = put_fixed4(p, e.length);
        p = put_fixed4(p, e.cosine);
        *p++ = '\n';
    }
    out.resize(start + (size_t)(p - base));
}

void EdgeFileWriter::writeFrame(int frnr, const wn_edge* edges, size_t n_edges)
{
    buf_.clear();
    buf_ += "#--- frame ";
    buf_ += std::to_string(frnr);
    buf_ += " ---\n";
    formatRows(edges, n_edges, buf_);
    os_.write(buf_.data(), (std::streamsize)buf_.size());
    os_.flush();                                  // the reference flushes per frame (:651)
}

void EdgeFileWriter::writeNodesHeader(std::ostream& os)
{
    os << std::setw(7) << "#Id" << std::setw(7) << "AtmId" << std::setw(7) << "Name" << std::setw(7) << "ResId"
       << std::setw(7) << "Res" << std::setw(7) << "Type" << std::setw(7) << "Hyd\n";
    os.flush();
}

void EdgeFileWriter::writeNode(std::ostream& os, const SiteInfo& site)
{
    os << std::setw(7) << site.id << std::setw(7) << site.atmIndex << std::setw(7) << site.name << std::setw(7)
       << site.resIndex << std::setw(7) << site.resName << std::setw(7) << site.type << std::setw(7)
       << site.nbHydrogen << "\n";
}
