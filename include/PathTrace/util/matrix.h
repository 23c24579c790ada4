// Small row-major matrices (API of the reference's include/PathTrace/util/matrix.h): mat3, mat4 with the affine
// vec3 product (w-divide) that io::loadMesh and the demo use for vertex transforms.
#ifndef PATHTRACE_MATRIX_H
#define PATHTRACE_MATRIX_H

#include <PathTrace/util/vector.h>

namespace impl {

    template<typename TYPE, int WIDTH, int HEIGHT>
    struct matrix {
        rt_vector<TYPE, WIDTH> rows[HEIGHT];

        rt_vector<TYPE, WIDTH> &operator[](std::size_t row) noexcept { return rows[row]; }
        constexpr const rt_vector<TYPE, WIDTH> &operator[](std::size_t row) const noexcept { return rows[row]; }

        constexpr matrix<TYPE, WIDTH, HEIGHT> operator*(TYPE factor) const noexcept {
            matrix<TYPE, WIDTH, HEIGHT> scaled{};
            for(int r = 0; r < HEIGHT; r++) {
                scaled.rows[r] = rows[r] * factor;
            }
            return scaled;
        }

        constexpr rt_vector<TYPE, HEIGHT> operator*(const rt_vector<TYPE, WIDTH> &vec) const noexcept {
            rt_vector<TYPE, HEIGHT> out{};
            for(int r = 0; r < HEIGHT; r++) {
                out[r] = dot(rows[r], vec);
            }
            return out;
        }
    };

}

template<typename T>
using mat3 = impl::matrix<T, 3, 3>;

template<typename T>
struct mat4 final : public impl::matrix<T, 4, 4> {
    using impl::matrix<T, 4, 4>::operator*;

    //! affine transform of a point: (x, y, z, 1) is multiplied, then divided by the resulting w
    constexpr impl::rt_vector<T, 3> operator*(const impl::rt_vector<T, 3> &point) const noexcept {
        const impl::rt_vector<T, 4> lifted{{point[0], point[1], point[2], static_cast<T>(1)}};
        impl::rt_vector<T, 4> image = impl::matrix<T, 4, 4>::operator*(lifted);
        image = image * (static_cast<T>(1) / image[3]);
        return {{image[0], image[1], image[2]}};
    }
};

template<typename T>
const mat4<T> mat4_identity{vec4<T>{{1, 0, 0, 0}}, vec4<T>{{0, 1, 0, 0}}, vec4<T>{{0, 0, 1, 0}}, vec4<T>{{0, 0, 0, 1}}};

#endif // PATHTRACE_MATRIX_H
