// RGBA colour = 4-vector with channel accessors (API of the reference's include/PathTrace/util/color.h).
#ifndef PATHTRACE_COLOR_H
#define PATHTRACE_COLOR_H

#include <PathTrace/util/vector.h>

template<typename TYPE>
struct Color : public impl::named_vector<TYPE, 4> {
    using impl::named_vector<TYPE, 4>::named_vector;
    Color() noexcept = default;
    Color(const impl::rt_vector<TYPE, 4> &other) noexcept : impl::named_vector<TYPE, 4>(other) {}

    TYPE &r() noexcept { return this->elements[0]; }
    TYPE &g() noexcept { return this->elements[1]; }
    TYPE &b() noexcept { return this->elements[2]; }
    TYPE &a() noexcept { return this->elements[3]; }
    constexpr TYPE r() const noexcept { return this->elements[0]; }
    constexpr TYPE g() const noexcept { return this->elements[1]; }
    constexpr TYPE b() const noexcept { return this->elements[2]; }
    constexpr TYPE a() const noexcept { return this->elements[3]; }
};

#endif /* PATHTRACE_COLOR_H */
