// Fixed-size numeric vectors for the PathTrace host API (B200 edition).
//
// API-compatible with the reference's include/PathTrace/util/vector.h (impl::rt_vector, vec2, vec3, vec4, dot, min,
// max, cross, reflect) so that code written against the reference compiles unchanged.  The host side only uses these
// for scene set-up; sums run in index order in fp32 so that set-up arithmetic rounds like the reference's.
#ifndef PATHTRACE_VECTOR_H
#define PATHTRACE_VECTOR_H

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <type_traits>
#include <utility>

namespace impl {

    template<typename TYPE, int SIZE>
    struct rt_vector {
        using value_type = TYPE;
        using self = rt_vector<TYPE, SIZE>;

        TYPE elements[SIZE];

        constexpr std::size_t size() const noexcept { return static_cast<std::size_t>(SIZE); }

        TYPE &operator[](std::size_t i) noexcept { return elements[i]; }
        constexpr TYPE operator[](std::size_t i) const noexcept { return elements[i]; }

        TYPE *data() noexcept { return elements; }
        constexpr const TYPE *data() const noexcept { return elements; }

        // element-wise combination helpers; every arithmetic operator below is one of these two shapes
        template<typename F>
        constexpr self zip(const self &rhs, F f) const noexcept {
            self out{};
            for(int i = 0; i < SIZE; i++) {
                out.elements[i] = f(elements[i], rhs.elements[i]);
            }
            return out;
        }

        template<typename F>
        constexpr self map(F f) const noexcept {
            self out{};
            for(int i = 0; i < SIZE; i++) {
                out.elements[i] = f(elements[i]);
            }
            return out;
        }

        constexpr bool operator==(const self &rhs) const noexcept {
            bool same = true;
            for(int i = 0; i < SIZE; i++) {
                same = same && (elements[i] == rhs.elements[i]);
            }
            return same;
        }
        constexpr bool operator!=(const self &rhs) const noexcept { return !(*this == rhs); }

        constexpr self operator+(const self &rhs) const noexcept {
            return zip(rhs, [](TYPE a, TYPE b) { return a + b; });
        }
        constexpr self operator-(const self &rhs) const noexcept {
            return zip(rhs, [](TYPE a, TYPE b) { return a - b; });
        }
        constexpr self operator*(const self &rhs) const noexcept {
            return zip(rhs, [](TYPE a, TYPE b) { return a * b; });
        }
        constexpr self operator*(TYPE factor) const noexcept {
            return map([factor](TYPE a) { return a * factor; });
        }
        constexpr self operator/(TYPE divisor) const noexcept {
            return map([divisor](TYPE a) { return a / divisor; });
        }
        constexpr self operator-() const noexcept {
            return map([](TYPE a) { return -a; });
        }

        self &operator+=(const self &rhs) noexcept { return *this = *this + rhs; }
        self &operator-=(const self &rhs) noexcept { return *this = *this - rhs; }
        self &operator*=(TYPE factor) noexcept { return *this = *this * factor; }
        self &operator/=(TYPE divisor) noexcept { return *this = *this / divisor; }

        //! squared euclidean length, summed in index order
        constexpr TYPE getLengthSquared() const noexcept {
            TYPE sum = static_cast<TYPE>(0);
            for(int i = 0; i < SIZE; i++) {
                sum += elements[i] * elements[i];
            }
            return sum;
        }

        TYPE getLength() const noexcept { return std::sqrt(getLengthSquared()); }

        //! scales by the reciprocal of the length; unspecified for the zero vector
        self normalize() const noexcept {
            const TYPE reciprocal = static_cast<TYPE>(1) / getLength();
            return *this * reciprocal;
        }

        self normalizeSafely() noexcept {
            if(std::abs(getLength()) > static_cast<TYPE>(0)) {
                return normalize();
            }
            return *this;
        }
    };

}

template<typename TYPE, int SIZE>
constexpr TYPE dot(const impl::rt_vector<TYPE, SIZE> &a, const impl::rt_vector<TYPE, SIZE> &b) noexcept {
    TYPE sum = static_cast<TYPE>(0);
    for(int i = 0; i < SIZE; i++) {
        sum += a[i] * b[i];
    }
    return sum;
}

template<typename TYPE, int SIZE>
constexpr impl::rt_vector<TYPE, SIZE> min(const impl::rt_vector<TYPE, SIZE> &a, const impl::rt_vector<TYPE, SIZE> &b) noexcept {
    return a.zip(b, [](TYPE x, TYPE y) { return std::min(x, y); });
}

template<typename TYPE, int SIZE>
constexpr impl::rt_vector<TYPE, SIZE> max(const impl::rt_vector<TYPE, SIZE> &a, const impl::rt_vector<TYPE, SIZE> &b) noexcept {
    return a.zip(b, [](TYPE x, TYPE y) { return std::max(x, y); });
}

template<typename TYPE>
constexpr impl::rt_vector<TYPE, 3> cross(const impl::rt_vector<TYPE, 3> &a, const impl::rt_vector<TYPE, 3> &b) noexcept {
    return {{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}};
}

//! mirrors v at the plane with unit normal n
template<typename TYPE, int SIZE>
constexpr impl::rt_vector<TYPE, SIZE> reflect(const impl::rt_vector<TYPE, SIZE> &v, const impl::rt_vector<TYPE, SIZE> &n) noexcept {
    const TYPE d = dot(v, n);
    return v - n * static_cast<TYPE>(2) * d;
}

namespace impl {

    // shared shape of vec2 / vec3 / Color: an rt_vector with element-list and rt_vector converting constructors
    template<typename TYPE, int N>
    struct named_vector : public rt_vector<TYPE, N> {
        using T = TYPE;
        static constexpr int SIZE = N;

        named_vector() noexcept = default;

        template<typename... ARGS, typename = std::enable_if_t<(sizeof...(ARGS) > 0) && (std::is_convertible_v<ARGS, TYPE> && ...)>>
        named_vector(ARGS... values) noexcept : rt_vector<TYPE, N>{{static_cast<TYPE>(values)...}} {}

        named_vector(const rt_vector<TYPE, N> &other) noexcept : rt_vector<TYPE, N>(other) {}
    };

}

template<typename TYPE>
struct vec2 final : public impl::named_vector<TYPE, 2> {
    using impl::named_vector<TYPE, 2>::named_vector;
    vec2() noexcept = default;
    vec2(const impl::rt_vector<TYPE, 2> &other) noexcept : impl::named_vector<TYPE, 2>(other) {}

    TYPE &x() noexcept { return this->elements[0]; }
    TYPE &y() noexcept { return this->elements[1]; }
    TYPE &u() noexcept { return this->elements[0]; }
    TYPE &v() noexcept { return this->elements[1]; }
    constexpr TYPE x() const noexcept { return this->elements[0]; }
    constexpr TYPE y() const noexcept { return this->elements[1]; }
    constexpr TYPE u() const noexcept { return this->elements[0]; }
    constexpr TYPE v() const noexcept { return this->elements[1]; }
};

template<typename TYPE>
struct vec3 final : public impl::named_vector<TYPE, 3> {
    using impl::named_vector<TYPE, 3>::named_vector;
    vec3() noexcept = default;
    vec3(const impl::rt_vector<TYPE, 3> &other) noexcept : impl::named_vector<TYPE, 3>(other) {}

    TYPE &x() noexcept { return this->elements[0]; }
    TYPE &y() noexcept { return this->elements[1]; }
    TYPE &z() noexcept { return this->elements[2]; }
    TYPE &u() noexcept { return this->elements[0]; }
    TYPE &v() noexcept { return this->elements[1]; }
    TYPE &w() noexcept { return this->elements[2]; }
    constexpr TYPE x() const noexcept { return this->elements[0]; }
    constexpr TYPE y() const noexcept { return this->elements[1]; }
    constexpr TYPE z() const noexcept { return this->elements[2]; }
    constexpr TYPE u() const noexcept { return this->elements[0]; }
    constexpr TYPE v() const noexcept { return this->elements[1]; }
    constexpr TYPE w() const noexcept { return this->elements[2]; }
};

//! homogeneous coordinates for 3D affine maps
template<typename T>
using vec4 = impl::rt_vector<T, 4>;

#endif /* PATHTRACE_VECTOR_H */
