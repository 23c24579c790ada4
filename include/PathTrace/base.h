// Core value types of the PathTrace host API (B200 edition): Ray, the xorshift engine and its RandomEngine wrapper,
// and the debug assertion helpers.  API of the reference's include/PathTrace/base.h.
//
// PATHTRACE_B200 is defined so that client code can detect the GPU-backed implementation and use its batch
// extensions (Scene::getIntersections, ptb::renderSamples, ...).
#ifndef PATHTRACE_BASE_H
#define PATHTRACE_BASE_H

#define PATHTRACE_B200 1

#include <PathTrace/util/matrix.h>
#include <PathTrace/util/vector.h>

#include <cassert>
#include <cmath>
#include <cstdint>
#include <limits>
#include <random>

//! origin + unit direction
struct Ray {
    vec3<float> origin;
    vec3<float> dir;
};

/**
 * 64-bit xorshift-multiply generator producing 32-bit outputs.  The device code carries the identical engine
 * (cpupathtrace_b200/csrc/rng.cuh) so that a host RandomEngine and a device path consume one stream:
 * state() / setState() hand the raw state across the C-ABI and back.
 */
class xorshift {
  public:
    using result_type = uint32_t;

    xorshift(uint64_t seed) noexcept : m_state(seed ^ (~seed << 32)) {}

    uint32_t operator()() noexcept {
        const uint64_t product = m_state * 0xD989BCACC137DCD5ULL;
        m_state ^= m_state >> 11;
        m_state ^= m_state << 31;
        m_state ^= m_state >> 18;
        return static_cast<uint32_t>(product >> 32);
    }

    static constexpr uint32_t min() noexcept { return std::numeric_limits<uint32_t>::min(); }
    static constexpr uint32_t max() noexcept { return std::numeric_limits<uint32_t>::max(); }

    uint64_t state() const noexcept { return m_state; }
    void setState(uint64_t state) noexcept { m_state = state; }

  private:
    uint64_t m_state;
};

//! thin wrapper handing out random bits; usable as a UniformRandomBitGenerator
class RandomEngine {
  private:
    xorshift engine;

  public:
    using result_type = uint32_t;

    RandomEngine(auto seed) noexcept : engine(static_cast<uint64_t>(seed)) {}

    auto operator()() noexcept { return engine(); }

    static constexpr auto min() noexcept { return xorshift::min(); }
    static constexpr auto max() noexcept { return xorshift::max(); }

    uint64_t state() const noexcept { return engine.state(); }
    void setState(uint64_t state) noexcept { engine.setState(state); }
};

template<typename T, int SIZE>
bool isNormalized(impl::rt_vector<T, SIZE> vec) noexcept {
    return std::abs(vec.getLengthSquared() - static_cast<T>(1)) < static_cast<T>(1E-4);
}

template<typename T, int SIZE>
bool isNonNegative(impl::rt_vector<T, SIZE> vec) noexcept {
    bool ok = true;
    for(int i = 0; i < SIZE; i++) {
        ok = ok && (vec[i] >= static_cast<T>(0)); // false for NaN
    }
    return ok;
}

#define assertNormalized(x) assert(isNormalized(x)) // NOLINT
#define assertNonNegative(x) assert(isNonNegative(x)) // NOLINT
#define assertFinite(x) assert(std::isfinite(x)) // NOLINT

#endif /* PATHTRACE_BASE_H */
