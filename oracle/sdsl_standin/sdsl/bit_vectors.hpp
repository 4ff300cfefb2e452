// TEST INFRASTRUCTURE ONLY (oracle/). Minimal stand-in for the part of SDSL
// (simongog/sdsl-lite, version unpinned by the reference: only a find_path for
// sdsl/suffix_arrays.hpp in src/cpp/CMakeLists.txt:46) that the reference's
// msa_transforms.cpp touches: bit_vector(size, fill), operator[], size(), and
// select_1_type / select_0_type (k-th bit equal to b, k is 1-based).
// The semantics are fully determined by the published rank/select definition,
// so this header lets the UNMODIFIED reference source compile into oracle/_ref.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace sdsl {

template <int BIT>
struct standin_select;

class bit_vector {
    std::vector<uint8_t> bits_;

public:
    struct proxy {
        uint8_t& cell;
        operator bool() const { return cell != 0; }
        proxy& operator=(int v) {
            cell = static_cast<uint8_t>(v != 0);
            return *this;
        }
        proxy& operator=(const proxy& other) {
            cell = other.cell;
            return *this;
        }
    };

    using select_1_type = standin_select<1>;
    using select_0_type = standin_select<0>;

    bit_vector() = default;
    bit_vector(size_t n, int fill) : bits_(n, static_cast<uint8_t>(fill != 0)) {}

    size_t size() const { return bits_.size(); }
    proxy operator[](size_t i) { return proxy{bits_[i]}; }
    bool operator[](size_t i) const { return bits_[i] != 0; }
};

template <int BIT>
struct standin_select {
    std::vector<size_t> where_;

    standin_select() = default;
    explicit standin_select(const bit_vector* bv) {
        for (size_t i = 0; i < bv->size(); ++i)
            if (static_cast<int>((*bv)[i]) == BIT) where_.push_back(i);
    }
    // k-th (1-based) position holding BIT
    size_t operator()(size_t k) const { return where_.at(k - 1); }
};

}  // namespace sdsl
