// test_writer.cpp -- EdgeFileWriter against std::ostream driven exactly as the reference
// drives it (src/waternetwork.cpp:485-499 header, :638-652 rows).  CPU only.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iomanip>
#include <limits>
#include <random>
#include <sstream>

#include "edge-file-writer.hpp"

// the reference's row statement, values held as doubles like HydrogenBond's fields
static void reference_rows(std::ostream& os, const std::vector<unsigned>& ids, const wn_edge* e, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        os << std::setw(6) << ids.at(e[i].i) << std::setw(6) << ids.at(e[i].j) << std::setw(8) << (double)e[i].energy
           << std::setw(8) << (double)e[i].length << std::setw(8) << (double)e[i].cosine << "\n";
    }
}

int main()
{
    std::mt19937 rng(42);
    std::vector<unsigned> ids(2000000);
    for (size_t s = 0; s < ids.size(); ++s) ids[s] = (unsigned)(s + 1 + 37);      // protein offset + 1
    std::vector<wn_edge> edges;
    std::uniform_real_distribution<float> ue(-4.1f, 0.3f), ul(2.0f, 6.5f), uc(0.0f, 1.0f);
    std::uniform_int_distribution<int> ui(0, 99999);
    for (int k = 0; k < 200000; ++k) edges.push_back(wn_edge{ui(rng), ui(rng), ue(rng), ul(rng), uc(rng)});
    // rounding ties, signed zeros, wide fields, non-finite values
    const float specials[] = {0.0f, -0.0f, 0.00005f, -0.00005f, 0.00015f, 0.00025f, 0.99995f, 0.99994999f, 9.99995f,
                              -9.99995f, 999.99995f, 1234.5678f, -1234.5678f, 99999.9f, 1e7f, -1e9f, 3.4e38f,
                              1e-30f, -1e-30f, 6.5f, 5.5f, 0.5f, 0.125f, 0.0625f, 0.03125f, 1.00005f, 2.00015f,
                              std::numeric_limits<float>::infinity(), -std::numeric_limits<float>::infinity(),
                              std::numeric_limits<float>::quiet_NaN(), std::numeric_limits<float>::denorm_min()};
    for (float a : specials)
        for (float b : specials) edges.push_back(wn_edge{1999999, 5, a, b, a});
    // every float step around a few rounding boundaries
    for (float x = 0.12344f; x < 0.12347f; x = std::nextafter(x, 1.0f)) edges.push_back(wn_edge{0, 1, -x, x, x});

    std::ostringstream ref;
    ref << std::fixed << std::setprecision(4);
    ref << std::setw(6) << "#Id" << std::setw(8) << "AtmId" << std::setw(6) << "ResId" << std::setw(6) << "Id"
        << std::setw(8) << "AtmId" << std::setw(6) << "ResId" << std::setw(8) << "Energy" << std::setw(8) << "Length"
        << std::setw(8) << "Cos\n";
    ref << "#--- frame " << 17 << " ---\n";
    reference_rows(ref, ids, edges.data(), edges.size());

    std::ostringstream got;
    EdgeFileWriter w(got, ids);
    w.writeHeader();
    w.writeFrame(17, edges.data(), edges.size());
    if (got.str() != ref.str()) {
        const std::string a = got.str(), b = ref.str();
        size_t k = 0;
        while (k < a.size() && k < b.size() && a[k] == b[k]) ++k;
        std::fprintf(stderr, "mismatch at byte %zu:\n got: %.60s\n ref: %.60s\n", k, a.c_str() + (k > 20 ? k - 20 : 0),
                     b.c_str() + (k > 20 ? k - 20 : 0));
        return 1;
    }
    // nodes file
    std::ostringstream n1, n2;
    SiteInfo s; s.type = WATER; s.name = "OW"; s.id = 38; s.atmIndex = 114; s.resIndex = 12; s.resName = "SOL"; s.nbHydrogen = 2;
    EdgeFileWriter::writeNodesHeader(n1);
    EdgeFileWriter::writeNode(n1, s);
    n2 << std::setw(7) << "#Id" << std::setw(7) << "AtmId" << std::setw(7) << "Name" << std::setw(7) << "ResId"
       << std::setw(7) << "Res" << std::setw(7) << "Type" << std::setw(7) << "Hyd\n";
    n2 << std::setw(7) << s.id << std::setw(7) << s.atmIndex << std::setw(7) << s.name << std::setw(7) << s.resIndex
       << std::setw(7) << s.resName << std::setw(7) << s.type << std::setw(7) << s.nbHydrogen << "\n";
    if (n1.str() != n2.str()) { std::fprintf(stderr, "nodes mismatch\n"); return 1; }

    // throughput of the formatter vs iostream (informative)
    std::string out;
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < 5; ++r) { out.clear(); w.formatRows(edges.data(), 200000, out); }
    auto t1 = std::chrono::steady_clock::now();
    std::ostringstream slow;
    slow << std::fixed << std::setprecision(4);
    reference_rows(slow, ids, edges.data(), 200000);
    auto t2 = std::chrono::steady_clock::now();
    const double fast = 5.0 * out.size() / std::chrono::duration<double>(t1 - t0).count() / 1e6;
    const double ios = (double)slow.str().size() / std::chrono::duration<double>(t2 - t1).count() / 1e6;
    std::printf("test_writer OK: %zu rows identical; formatter %.0f MB/s, iostream %.0f MB/s\n", edges.size(), fast, ios);
    return 0;
}
