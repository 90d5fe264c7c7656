#include "cwr_topology.h"
#include <cstdio>
#include <cstdint>
#include <vector>
#include <chrono>
using namespace cwr;
template <typename T> uint64_t h(const std::vector<T>& v) {
    uint64_t x = 1469598103934665603ull;
    const unsigned char* p = reinterpret_cast<const unsigned char*>(v.data());
    for (size_t i = 0; i < v.size() * sizeof(T); ++i) { x ^= p[i]; x *= 1099511628211ull; }
    return x ^ v.size();
}
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb");
    int32_t hdr[3]; fread(hdr, 4, 3, f);
    int n = hdr[0], F = hdr[1], E = hdr[2];
    std::vector<int32_t> f1(E), f2(E); std::vector<float> hint(E);
    fread(f1.data(), 4, E, f); fread(f2.data(), 4, E, f); fread(hint.data(), 4, E, f); fclose(f);
    struct Cfg { bool rcm; int nc; bool hint; int parts; int strips; int cap; };
    Cfg cfgs[] = {{true, 0, false, 1, 0, 0}, {true, 12, true, 1, 0, 256}, {true, 15, true, 1, 296, 256}, {true, 8, true, 4, 37, 1024},
                  {false, 11, false, 2, 0, 0}, {true, 24, false, 1, 16, 0}};
    for (auto& c : cfgs) {
        Topology T;
        auto t0 = std::chrono::steady_clock::now();
        std::string err = build_topology(n, F, E, f1.data(), f2.data(), c.rcm, c.nc, c.hint ? hint.data() : nullptr, c.parts, T, c.strips, c.cap);
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (!err.empty()) { printf("ERR %s\n", err.c_str()); continue; }
        fprintf(stderr, "cfg nc=%d parts=%d strips=%d: %.3f s\n", c.nc, c.parts, c.strips, s);
        printf("%llx %llx %llx %llx %llx %llx %llx %llx %llx %llx %llx %llx %llx %llx %llx %llx %llx %llx %llx %llx %llx %d %d %lld %d\n",
            (unsigned long long)h(T.old_of_new), (unsigned long long)h(T.new_of_old), (unsigned long long)h(T.eperm), (unsigned long long)h(T.f1p), (unsigned long long)h(T.f2p),
            (unsigned long long)h(T.rowptr), (unsigned long long)h(T.col), (unsigned long long)h(T.slot_edge), (unsigned long long)h(T.ell_col), (unsigned long long)h(T.ell_code),
            (unsigned long long)h(T.bcell), (unsigned long long)h(T.bptr), (unsigned long long)h(T.bedge), (unsigned long long)h(T.color_ptr), (unsigned long long)h(T.color_of),
            (unsigned long long)h(T.part_ptr), (unsigned long long)h(T.strip_cptr), (unsigned long long)h(T.strip_nptr), (unsigned long long)h(T.strip_nbr), (unsigned long long)h(T.send_rows),
            (unsigned long long)h(T.iedge_ptr), T.W, T.n_levels, (long long)T.bandwidth, T.max_strip_nbr);
    }
}
