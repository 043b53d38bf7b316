// examples/grid_search.cpp — the reference's grid greedy-search test
// (diskann/src/graph/test/cases/grid_search.rs:86-207) written against the C++ host mirror.
// Usage: grid_search <dims> <size> <beam> <query-value>   -> prints "id:distance" pairs, cmps, hops
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../include/diskann_b200.hpp"

using namespace diskann_b200;

int main(int argc, char** argv) {
    if (argc != 5) {
        std::fprintf(stderr, "usage: %s dims size beam query_value\n", argv[0]);
        return 2;
    }
    const int dims = std::atoi(argv[1]), size = std::atoi(argv[2]);
    const uint32_t beam = (uint32_t)std::atoi(argv[3]);
    const float qv = (float)std::atof(argv[4]);
    try {
        uint32_t n = 1;
        for (int a = 0; a < dims; ++a) n *= (uint32_t)size;
        // synthetic.rs:102-348: lattice points, last coordinate fastest; start point (size,..,size)
        std::vector<float> data((size_t)(n + 1) * dims);
        std::vector<uint32_t> adj((size_t)(n + 1) * (2 * dims + 1), 0);
        std::vector<uint32_t> stride(dims);
        for (int a = 0; a < dims; ++a) {
            stride[a] = 1;
            for (int b = a + 1; b < dims; ++b) stride[a] *= (uint32_t)size;
        }
        for (uint32_t i = 0; i < n; ++i) {
            uint32_t deg = 0;
            uint32_t* row = &adj[(size_t)i * (2 * dims + 1)];
            for (int a = 0; a < dims; ++a) {
                const uint32_t c = (i / stride[a]) % (uint32_t)size;
                data[(size_t)i * dims + a] = (float)c;
                if (c > 0) row[1 + deg++] = i - stride[a];
                if (c + 1 < (uint32_t)size) row[1 + deg++] = i + stride[a];
            }
            row[0] = deg;
        }
        for (int a = 0; a < dims; ++a) data[(size_t)n * dims + a] = (float)size;
        adj[(size_t)n * (2 * dims + 1)] = 1;
        adj[(size_t)n * (2 * dims + 1) + 1] = n - 1;

        Provider<float> provider(Metric::L2, (uint32_t)dims, n, 1, (uint32_t)(2 * dims));
        provider.set_elements(data.data(), 0, n + 1);
        provider.set_neighbors(adj.data(), (uint32_t)(2 * dims + 1), 0, n + 1);
        GpuKNN<float> knn(provider, /*l_value=*/10, beam);
        std::vector<float> query(dims, qv);
        KnnResults r = knn.search(query.data(), 1, 10);
        for (uint32_t j = 0; j < r.stats[0].result_count; ++j) std::printf("%u:%g ", r.ids[j], r.distances[j]);
        std::printf("| cmps=%u hops=%u count=%u\n", r.stats[0].cmps, r.stats[0].hops, r.stats[0].result_count);
        return 0;
    } catch (const ANNError& e) {
        std::fprintf(stderr, "ANNError %d: %s\n", e.code(), e.what());
        return 10 + e.code();
    }
}
