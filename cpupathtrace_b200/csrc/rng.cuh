// Random numbers for the path sampler.
//
// Two engines behind one draw() interface, selected per launch (ptb_rng_mode):
//
//   * PTB_RNG_REFERENCE_XORSHIFT — the reference engine (include/PathTrace/base.h:24-41): 64-bit state seeded as
//     `seed ^ (~seed << 32)`, output = high 32 bits of `state * 0xD989BCACC137DCD5` taken BEFORE the state update
//     `s ^= s >> 11; s ^= s << 31; s ^= s >> 18`.  One engine per (pixel, sample); the validation oracle runs
//     processItem on a 1x1 WorkItem with RandomEngine(seed), which consumes the identical stream.
//   * PTB_RNG_COUNTER — production: stateless, counter-based.  key = mix(job seed, pixel, sample); the n-th draw of
//     bounce b is mix(key, b << 8 | n): any (pixel, sample, bounce) can be generated independently of launch shape,
//     pool size or GPU count, which is what makes multi-GPU sharding result-invariant.
//
// The distributions restate libstdc++ 13 arithmetic bit for bit (bits/random.tcc generate_canonical, bits/random.h
// uniform_real_distribution::operator() and bernoulli_distribution::operator()):
//   uniform_real<float>(a,b): ONE draw x, u = float(x) / 2^32 (round to nearest; if u >= 1 then nextafter(1,0)),
//                             result u * (b - a) + a
//   bernoulli(p)            : TWO draws x0, x1 combined in double: u = (double(x0) + double(x1) * 2^32) / 2^64
//                             (clamped below 1), result u < double(p)
#ifndef PTB_RNG_CUH
#define PTB_RNG_CUH

#include "device_math.cuh"

namespace ptb {

    PTB_DEV uint64_t mix64(uint64_t z) {
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }

    // XORSHIFT = true: the reference engine (validation); false: the counter-based production generator.
    // The kind is a compile-time property of the kernels so that neither pays for the other's arithmetic.
    template<bool XORSHIFT>
    struct RngT {
        static constexpr bool kXorshift = XORSHIFT;
        uint64_t state;   // xorshift state, or the (pixel, sample) key in counter mode
        uint32_t counter; // counter mode: (bounce << 8) | draw-in-bounce

        PTB_DEV uint32_t draw() {
            if constexpr(XORSHIFT) {
                const uint64_t result = state * 0xD989BCACC137DCD5ULL;
                state ^= state >> 11;
                state ^= state << 31;
                state ^= state >> 18;
                return static_cast<uint32_t>(result >> 32);
            }
            else {
                const uint64_t z = mix64(state + 0x9E3779B97F4A7C15ULL * (static_cast<uint64_t>(counter) + 1ULL));
                counter++;
                return static_cast<uint32_t>(z >> 32);
            }
        }

        // generate_canonical<float, 24>
        PTB_DEV float canonical() {
            const float u = __uint2float_rn(draw()) / 4294967296.0F;
            return u >= 1.0F ? 0.99999994F : u;
        }

        // two consecutive uniform_real_distribution<float>(0, 1) draws, `first` drawn first.  The reference engine draws
        // twice; the counter generator takes both halves of ONE 64-bit mix (its two 64-bit multiplies are the most
        // expensive integer work of the shade kernel: 9 mixes per path vertex in round 1, 6 with the pairs)
        PTB_DEV void canonicalPair(float &first, float &second) {
            if constexpr(XORSHIFT) {
                first = canonical();
                second = canonical();
            }
            else {
                const uint64_t z = mix64(state + 0x9E3779B97F4A7C15ULL * (static_cast<uint64_t>(counter) + 1ULL));
                counter++;
                const float u = __uint2float_rn(static_cast<uint32_t>(z >> 32)) / 4294967296.0F;
                const float v = __uint2float_rn(static_cast<uint32_t>(z)) / 4294967296.0F;
                first = u >= 1.0F ? 0.99999994F : u;
                second = v >= 1.0F ? 0.99999994F : v;
            }
        }

        // uniform_real_distribution<float>(a, b)
        PTB_DEV float uniform(float a, float b) {
            return canonical() * (b - a) + a;
        }

        // uniform_real_distribution<float>(0, 1): u * (1 - 0) + 0 == u
        PTB_DEV float uniform01() {
            return canonical();
        }

        // bernoulli_distribution(p)
        PTB_DEV bool bernoulli(double p) {
#if defined(PTB_FAST_MATH)
            if constexpr(!XORSHIFT) {
                // production build: one 32-bit draw compared in fp32 (the reference's two-draw fp64 form buys nothing here)
                return canonical() < static_cast<float>(p);
            }
#endif
            const double x0 = static_cast<double>(draw());
            const double x1 = static_cast<double>(draw());
            const double sum = x0 + x1 * 4294967296.0;
            double u = sum / 18446744073709551616.0;
            if(u >= 1.0) {
                u = 0.99999999999999988898;
            }
            return u < p * 1.0;
        }

        // counter mode: the draws of path vertex `bounce` start at counter bounce << 8
        PTB_DEV void startBounce(int bounce) {
            if constexpr(!XORSHIFT) {
                counter = static_cast<uint32_t>(bounce) << 8;
            }
        }
    };

    using ReferenceRng = RngT<true>;
    using CounterRng = RngT<false>;

    PTB_DEV uint64_t xorshiftSeed(uint64_t seed) {
        return seed ^ (~seed << 32);
    }

    PTB_DEV uint64_t counterKey(uint64_t job_seed, uint32_t pixel_x, uint32_t pixel_y, uint32_t sample) {
        uint64_t k = mix64(job_seed ^ 0xA0761D6478BD642FULL);
        k = mix64(k ^ ((static_cast<uint64_t>(pixel_y) << 32) | pixel_x));
        k = mix64(k ^ (static_cast<uint64_t>(sample) * 0xE7037ED1A0B428DBULL + 1ULL));
        return k;
    }

}

#endif
