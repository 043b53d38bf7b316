// diskann_b200.hpp — C++ host-side mirror of the reference's operator interface for the distance
// hot path, layered on the C ABI in diskann_b200.h (header-only; link libdiskann_b200.so).
//
// The reference is compiled Rust and this image has no Rust toolchain, so the host layer above
// the C ABI is C++ with the reference's names, argument meaning and error behaviour:
//
//   Metric                         diskann-vector/src/distance/metric.rs:8-20
//   Distance<T,U>::call            diskann-vector/src/distance/distance_provider.rs:62-100
//   distance_comparer<T,U>         DistanceProvider::distance_comparer, distance_provider.rs:44-46
//   Provider                       diskann_inmem::Provider (diskann-inmem/src/provider.rs:71-131):
//                                  set_element / set_neighbors / search accessor creation
//   QueryDistance / expand_beam    layers::QueryDistance::evaluate + SearchAccessor::expand_beam
//                                  (diskann-inmem/src/layers/mod.rs:59-77; provider.rs:436-479),
//                                  batched over queries
//   GpuKNN::search                 benchmark_core::search::graph::KNN::search
//                                  (diskann-benchmark-core/src/search/graph/knn.rs:208-238) for a
//                                  whole query batch (the 3' boundary of SURVEY.md §8b)
//   SearchStats                    diskann/src/graph/index.rs:90 (cmps, hops, result_count)
//   MinMaxQuantizer                diskann-quantization/src/minmax/quantizer.rs:69-228 (Transform::Null) + the
//                                  MinMax distance functors over compressed rows (vectors.rs:231-455)
//
// Errors: every non-zero status becomes ANNError (the inmem layer returns Err on length / type
// mismatch, layers/full.rs:203-213, 306-314; it never panics across the boundary).
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "diskann_b200.h"

namespace diskann_b200 {

enum class Metric : int { Cosine = DAB_COSINE, InnerProduct = DAB_INNER_PRODUCT, L2 = DAB_L2, CosineNormalized = DAB_COSINE_NORMALIZED };

// IEEE binary16 storage type (the reference uses half::f16)
struct f16 {
    uint16_t bits;
};

class ANNError : public std::runtime_error {
   public:
    ANNError(int code, const std::string& msg) : std::runtime_error(msg), code_(code) {}
    int code() const { return code_; }

   private:
    int code_;
};

inline void check(int status) {
    if (status != DAB_OK) throw ANNError(status, dab_last_error());
}

template <class T>
struct ElementType;
template <>
struct ElementType<float> {
    static constexpr int value = DAB_F32;
};
template <>
struct ElementType<f16> {
    static constexpr int value = DAB_F16;
};
template <>
struct ElementType<int8_t> {
    static constexpr int value = DAB_I8;
};
template <>
struct ElementType<uint8_t> {
    static constexpr int value = DAB_U8;
};

// Distance<T, U>: what T::distance_comparer(metric, Some(dim)) returns.
template <class T, class U = T>
class Distance {
   public:
    Distance(Metric metric, size_t dim, int device = 0) : metric_(metric), dim_(dim), device_(device) {}
    // distance_comparer.call(x, y): both slices must have the comparer's dimension
    float call(const T* x, size_t xlen, const U* y, size_t ylen) const {
        if (xlen != dim_ || ylen != dim_)
            throw ANNError(DAB_ERR_INVALID_ARGUMENT, "expected slices of length " + std::to_string(dim_) + " - instead got " +
                                                         std::to_string(xlen) + " and " + std::to_string(ylen));
        float out = 0.0f;
        check(dab_pair_distances(ElementType<T>::value, ElementType<U>::value, (int)metric_, (uint32_t)dim_, x, y, 1, &out, device_));
        return out;
    }
    // n pairs at once (rows dense, dim elements each)
    std::vector<float> call_batch(const T* x, const U* y, size_t n) const {
        std::vector<float> out(n);
        check(dab_pair_distances(ElementType<T>::value, ElementType<U>::value, (int)metric_, (uint32_t)dim_, x, y, n, out.data(), device_));
        return out;
    }

   private:
    Metric metric_;
    size_t dim_;
    int device_;
};

template <class T, class U = T>
Distance<T, U> distance_comparer(Metric metric, size_t dim, int device = 0) {
    return Distance<T, U>(metric, dim, device);
}

struct SearchStats {
    uint32_t cmps, hops, result_count;
};

struct KnnResults {
    uint32_t nq, k;
    std::vector<uint32_t> ids;     // [nq][k], padded UINT32_MAX
    std::vector<float> distances;  // [nq][k], padded +inf
    std::vector<SearchStats> stats;
};

// Device-resident snapshot of an in-memory provider for element type T.
template <class T>
class Provider {
   public:
    Provider(Metric metric, uint32_t dim, uint64_t n_points, uint32_t n_start, uint32_t max_degree, int device = 0)
        : dim_(dim), n_points_(n_points), n_start_(n_start), max_degree_(max_degree) {
        check(dab_create(&h_, ElementType<T>::value, (int)metric, dim, n_points, n_start, max_degree, device));
    }
    ~Provider() { dab_destroy(h_); }
    Provider(const Provider&) = delete;
    Provider& operator=(const Provider&) = delete;

    uint32_t dim() const { return dim_; }
    uint64_t n_points() const { return n_points_; }
    uint32_t n_start() const { return n_start_; }
    dab_index* raw() { return h_; }

    // SetElement: rows [first, first + count)
    void set_elements(const T* rows, uint64_t first, uint64_t count) { check(dab_upload_vectors(h_, rows, first, count)); }
    // neighbors().set_neighbors for rows [first, first + count): row = [len, ids...]
    void set_neighbors(const uint32_t* adj, uint32_t stride, uint64_t first, uint64_t count) {
        check(dab_upload_graph(h_, adj, stride, first, count));
    }
    std::vector<uint32_t> get_neighbors(uint64_t first, uint64_t count) {
        std::vector<uint32_t> adj(count * (size_t)(max_degree_ + 1));
        check(dab_download_graph(h_, adj.data(), max_degree_ + 1, first, count));
        return adj;
    }

    // QueryDistance::evaluate for every (query q, id ids[q][j]): expand_beam's distance stage
    std::vector<float> query_distances(const T* queries, uint32_t nq, const uint32_t* ids, uint32_t c) {
        std::vector<float> out((size_t)nq * c);
        check(dab_distances(h_, queries, nq, ids, c, out.data()));
        return out;
    }
    // Distance<T,T> between stored rows (prune closure)
    std::vector<float> row_distances(const uint32_t* a, const uint32_t* b, uint64_t n) {
        std::vector<float> out(n);
        check(dab_row_pair_distances(h_, a, b, n, out.data()));
        return out;
    }

    // index construction on the device (multi_insert semantics)
    void build(uint32_t pruned_degree, uint32_t l_build, float alpha = 1.2f, uint32_t batch = 0) {
        check(dab_build(h_, pruned_degree, l_build, alpha, batch));
    }

   private:
    dab_index* h_ = nullptr;
    uint32_t dim_;
    uint64_t n_points_;
    uint32_t n_start_, max_degree_;
};

// KNN::search for a whole batch: Knn::new(l_value, beam_width) + k results per query.
template <class T>
class GpuKNN {
   public:
    GpuKNN(Provider<T>& provider, uint32_t l_value, uint32_t beam_width = 1) : p_(provider), l_(l_value), beam_(beam_width) {
        if (l_value == 0) throw ANNError(DAB_ERR_INVALID_ARGUMENT, "l_value cannot be zero");       // KnnSearchError::LZero
        if (beam_width == 0) throw ANNError(DAB_ERR_INVALID_ARGUMENT, "beam_width cannot be zero");  // BeamWidthZero
    }
    KnnResults search(const T* queries, uint32_t nq, uint32_t k) {
        KnnResults r;
        r.nq = nq;
        r.k = k;
        r.ids.resize((size_t)nq * k);
        r.distances.resize((size_t)nq * k);
        std::vector<uint32_t> counts(nq), cmps(nq), hops(nq);
        check(dab_search_batch(p_.raw(), queries, nq, k, l_, beam_, r.ids.data(), r.distances.data(), counts.data(), cmps.data(),
                               hops.data()));
        r.stats.resize(nq);
        for (uint32_t i = 0; i < nq; ++i) r.stats[i] = SearchStats{cmps[i], hops[i], counts[i]};
        return r;
    }

    // Batches in flight (search_all's one task per query partition, api.rs:410-419, as slots of the device):
    // `search_async` queues a batch on `slot` and returns at once, `wait` joins it and returns its results.
    // `queries` must stay valid until then.
    void search_async(uint32_t slot, const T* queries, uint32_t nq, uint32_t k) {
        if (slot >= DAB_MAX_SLOTS) throw ANNError(DAB_ERR_INVALID_ARGUMENT, "slot out of range");
        Pending& s = pending_[slot];
        s.r.nq = nq;
        s.r.k = k;
        s.r.ids.resize((size_t)nq * k);
        s.r.distances.resize((size_t)nq * k);
        s.counts.resize(nq), s.cmps.resize(nq), s.hops.resize(nq);
        check(dab_search_batch_async(p_.raw(), slot, queries, nq, k, l_, beam_, s.r.ids.data(), s.r.distances.data(), s.counts.data(),
                                     s.cmps.data(), s.hops.data()));
    }
    KnnResults wait(uint32_t slot) {
        check(dab_wait(p_.raw(), slot));
        Pending& s = pending_[slot];
        s.r.stats.resize(s.r.nq);
        for (uint32_t i = 0; i < s.r.nq; ++i) s.r.stats[i] = SearchStats{s.cmps[i], s.hops[i], s.counts[i]};
        return std::move(s.r);
    }

   private:
    struct Pending {
        KnnResults r;
        std::vector<uint32_t> counts, cmps, hops;
    };
    Provider<T>& p_;
    uint32_t l_, beam_;
    Pending pending_[DAB_MAX_SLOTS];
};

// MinMaxQuantizer (diskann-quantization/src/minmax/quantizer.rs:69-110, Transform::Null) and the MinMax distance
// functors over compressed rows (vectors.rs:231-455).  Rows are the reference's canonical-front Data<NBITS> bytes.
class MinMaxQuantizer {
   public:
    MinMaxQuantizer(uint32_t dim, float grid_scale, int device = 0) : dim_(dim), grid_scale_(grid_scale), device_(device) {}
    uint32_t dim() const { return dim_; }
    uint32_t output_dim() const { return dim_; }
    // Data::<NBITS>::canonical_bytes(dim)
    size_t canonical_bytes(int nbits) const { return dab_minmax_row_bytes(dim_, nbits); }
    // CompressInto<&[f32], DataMutRef<NBITS>> for n vectors; throws ANNError on NaN input (InputContainsNaN)
    std::vector<uint8_t> compress(const float* vectors, uint64_t n, int nbits, std::vector<float>* loss = nullptr) const {
        std::vector<uint8_t> rows(n * canonical_bytes(nbits));
        if (loss) loss->resize(n);
        check(dab_minmax_compress(device_, grid_scale_, dim_, nbits, vectors, n, rows.data(), loss ? loss->data() : nullptr));
        return rows;
    }
    // MinMax{L2Squared, IP, Cosine, CosineNormalized}::evaluate(DataRef<N>, DataRef<M>) row by row (N x N, 8 x N)
    std::vector<float> distances(Metric metric, int nbits_x, int nbits_y, const uint8_t* x_rows, const uint8_t* y_rows, uint64_t n) const {
        std::vector<float> out(n);
        check(dab_minmax_distances(device_, static_cast<int>(metric), nbits_x, nbits_y, dim_, x_rows, y_rows, n, out.data()));
        return out;
    }

    // CompressInto<&[f32], FullQueryMut> + MinMax*::evaluate(FullQueryRef, DataRef<NBITS>) for every (query, row): [nq][n]
    std::vector<float> query_distances(Metric metric, int nbits, const float* queries, uint32_t nq, const uint8_t* rows, uint64_t n) const {
        std::vector<float> out(static_cast<size_t>(nq) * n);
        check(dab_minmax_query_distances(device_, static_cast<int>(metric), nbits, dim_, queries, nq, rows, n, out.data()));
        return out;
    }

   private:
    uint32_t dim_;
    float grid_scale_;
    int device_;
};

}  // namespace diskann_b200
